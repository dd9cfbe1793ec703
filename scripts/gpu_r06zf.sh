#!/bin/bash
# r06zf: the bench line and the reference-arm line with the final build (host pipeline chunk 256 MB)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r06zf_bench.json 2> gpurun_out/r06zf_bench.err; echo "bench exit $?"; head -c 400 gpurun_out/r06zf_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r06zf_bench_ref.json 2>> gpurun_out/r06zf_bench.err; head -c 300 gpurun_out/r06zf_bench_ref.json; echo
timeout 600 python -m pytest tests/test_gpu_host_path.py -x -q > gpurun_out/r06zf_pytest_host.log 2>&1; tail -2 gpurun_out/r06zf_pytest_host.log
