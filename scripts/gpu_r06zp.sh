#!/bin/bash
# r06zp: the floating base's quaternion rows asked for in L2 already at its DESCEND (d2) against d0
mkdir -p gpurun_out
AB_ALGOS=aba AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06zp_ab.jsonl 2> gpurun_out/r06zp_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06zp_ab.jsonl'):
    d = json.loads(l); print("%-16s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
