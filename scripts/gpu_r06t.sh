#!/bin/bash
# r06t: CRBA as a persistent grid whose warps draw their states (MECANO_B200_PERSIST=1) against one block per tile
mkdir -p gpurun_out
for p in 0 1; do
  if [ $p = 1 ]; then export MECANO_B200_PERSIST=1; else unset MECANO_B200_PERSIST; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-extras 2>gpurun_out/r06t_bench_$p.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('persist=$p', 'step ms', round(d['ms_per_step'],4), {k:(round(v['ms'],4), round(v.get('fp64_frac',0),4), round(v['hbm_frac'],4)) for k,v in d['kernels'].items()})
"
done
