#!/bin/bash
# r06zk: CRBA prologue without the wait for the ring when op 0 is a floating base (c1) against c0: bench kernels, alternating
mkdir -p gpurun_out
for rep in 1 2; do for v in c0 c1; do
  MECANO_B200_LIB=mecano_b200/variants/$v.so timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', {k:round(v['ms'],4) for k,v in d['kernels'].items()}, {k:round(v['ms'],4) for k,v in d['extras'].items() if isinstance(v,dict) and 'ms' in v and k.startswith('c')})
"
done; done
