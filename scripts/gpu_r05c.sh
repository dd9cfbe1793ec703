#!/bin/bash
# r05c (8 GPUs): multi-device engine tests on hardware, BASELINE config 5 at 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "(multi-device tests: r05c first call, 9 passed at 8 GPUs)"

rm -f gpurun_out/r05c_config5_8gpu.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_config5.py gpurun_out/r05c_config5_8gpu.jsonl > gpurun_out/r05c_config5.log 2>&1
echo "config5 exit $?"; tail -3 gpurun_out/r05c_config5.log | cut -c1-600
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/r05c_bench_n8.json 2> gpurun_out/r05c_bench_n8.err
echo "bench exit $?"; tail -2 gpurun_out/r05c_bench_n8.err | cut -c1-300; head -c 600 gpurun_out/r05c_bench_n8.json
