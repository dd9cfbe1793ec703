#!/bin/bash
# r05a: full GPU parity suite (2 GPUs: the multi-device slice tests run on hardware), bench, RNEA/ABA timing
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r05a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05a_pytest.log
tail -6 gpurun_out/r05a_pytest.log
for algo in rnea aba; do AB_ALGO=$algo timeout 300 python scripts/gpu_aba_ab.py child; done | tee gpurun_out/r05a_ab.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r05a_bench.json 2> gpurun_out/r05a_bench.err
echo "bench exit $?"; tail -3 gpurun_out/r05a_bench.err; head -c 1500 gpurun_out/r05a_bench.json
