#!/bin/bash
# r06zi: no wait for the ring in the prologue of a state whose first op is not a one-DoF DESCEND (w1) against w0
mkdir -p gpurun_out
AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06zi_ab.jsonl 2> gpurun_out/r06zi_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06zi_ab.jsonl'):
    d = json.loads(l); print("%-10s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
