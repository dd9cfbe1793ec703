#!/bin/bash
# r06zl: does the length of the timed region move the per-kernel times (clocks under sustained FP64 load)?
mkdir -p gpurun_out
for k in 10 50 10 100 20; do
  timeout 300 python bench.py --steps $k --warmup 3 --no-e2e --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('steps $k', round(d['ms_per_step'],4), {k:(round(v['ms'],4), round(v.get('fp64_frac',0),4)) for k,v in d['kernels'].items()}, 'peak', round(d['roofline']['fp64_peak_tflops_measured_live'],2), d['clocks'])
"
done
