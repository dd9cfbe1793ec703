#!/bin/bash
# r06zg: L2 warmed with the floating base's rows before pass three (u1) against u0
mkdir -p gpurun_out
AB_ALGOS=aba AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06zg_ab.jsonl 2> gpurun_out/r06zg_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06zg_ab.jsonl'):
    d = json.loads(l); print("%-8s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
