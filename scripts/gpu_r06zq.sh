#!/bin/bash
# r06zq (N GPUs): the device-resident bench line at N GPUs with the final kernels (weak scaling; no e2e / extras)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/r06zq_bench_n$N.json 2> gpurun_out/r06zq_bench_n$N.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r06zq_bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
