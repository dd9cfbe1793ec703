#!/bin/bash
# r06f: ABA variants: v6 (sin/cos in records), v7 (= v6 without the empty plain cases / NEXT1 folding), v7_next1, v8 (32-bit record slot stride)
mkdir -p gpurun_out
AB_ROUNDS=1 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06f_ab.jsonl 2> gpurun_out/r06f_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06f_ab.jsonl'):
    d = json.loads(l); print("%-12s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
