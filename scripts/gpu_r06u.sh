#!/bin/bash
# r06u: plain ABA kinds again, now that pass three and DESCEND have one loop body per joint type: ASCEND (2), pass three (8), both (10)
mkdir -p gpurun_out
AB_ALGOS=aba AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06u_ab.jsonl 2> gpurun_out/r06u_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06u_ab.jsonl'):
    d = json.loads(l); print("%-10s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
