#!/bin/bash
# r06zc: discard predicate folded into the context, rotation-only transform for the floating base in pass three (t1) against t0
mkdir -p gpurun_out
AB_ALGOS=aba AB_ROUNDS=3 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06zc_ab.jsonl 2> gpurun_out/r06zc_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06zc_ab.jsonl'):
    d = json.loads(l); print("%-6s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
