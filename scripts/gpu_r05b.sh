#!/bin/bash
# r05b: full GPU parity suite (team kernels, three-DoF joints), racecheck of the body-parallel kernels, config-5 script on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r05b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05b_pytest.log
tail -15 gpurun_out/r05b_pytest.log
timeout 900 compute-sanitizer --tool racecheck python scripts/gpu_memcheck.py body > gpurun_out/r05b_racecheck.txt 2>&1; tail -5 gpurun_out/r05b_racecheck.txt
timeout 900 python scripts/gpu_config5.py gpurun_out/r05b_config5_1gpu.jsonl tree > gpurun_out/r05b_config5.log 2>&1; tail -3 gpurun_out/r05b_config5.log
