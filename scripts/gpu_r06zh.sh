#!/bin/bash
# r06zh: draws two ahead, next state's first rows warmed in L2 (v2) against u1; full GPU suite
mkdir -p gpurun_out
AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06zh_ab.jsonl 2> gpurun_out/r06zh_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06zh_ab.jsonl'):
    d = json.loads(l); print("%-9s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r06zh_pytest.log 2>&1; tail -3 gpurun_out/r06zh_pytest.log
