#!/bin/bash
# r06n: uniform base pointers + 32-bit per-thread offsets (x1) against per-thread 64-bit pointers (x0)
mkdir -p gpurun_out
AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06n_ab.jsonl 2> gpurun_out/r06n_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06n_ab.jsonl'):
    d = json.loads(l); print("%-12s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
