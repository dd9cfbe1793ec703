#!/bin/bash
# r06x (8 GPUs): the bench at 8 GPUs with the r06 kernels (one rank per GPU; e2e through one multi-device call from rank 0), the
# multi-device engine's slicing tests
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/r06x_bench_n8.json 2> gpurun_out/r06x_bench_n8.err
echo "bench exit $?"; tail -2 gpurun_out/r06x_bench_n8.err | cut -c1-300; head -c 500 gpurun_out/r06x_bench_n8.json; echo
timeout 600 python -m pytest tests/test_gpu_host_path.py -x -q -k "multi_device" > gpurun_out/r06x_multidevice_pytest.txt 2>&1; tail -3 gpurun_out/r06x_multidevice_pytest.txt
