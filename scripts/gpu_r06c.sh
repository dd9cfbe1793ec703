#!/bin/bash
# r06c: plain run kinds (compile-time flags) for RNEA and ABA against the previous builds; parity tests
mkdir -p gpurun_out
timeout 900 python scripts/gpu_ab.py > gpurun_out/r06c_ab.jsonl 2> gpurun_out/r06c_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06c_ab.jsonl'):
    d = json.loads(l); print("%-12s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "beyond or kernels_match or source_modes or golden" > gpurun_out/r06c_pytest.log 2>&1; tail -5 gpurun_out/r06c_pytest.log
