#!/bin/bash
# r06s: fast quaternion reciprocal in the ABA kernels only (z6) against z0; full GPU suite
mkdir -p gpurun_out
AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06s_ab.jsonl 2> gpurun_out/r06s_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06s_ab.jsonl'):
    d = json.loads(l); print("%-20s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r06s_pytest.log 2>&1; tail -3 gpurun_out/r06s_pytest.log
