#!/bin/bash
# r06zm: burst and sustained FP64 denominators next to the per-kernel times, at the driver's 20 + 5 steps and at 100
mkdir -p gpurun_out
for k in 20 100 20; do
  timeout 300 python bench.py --steps $k --warmup 5 --no-e2e --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('steps $k', round(d['ms_per_step'],4), {k:(round(v['ms'],4), round(v.get('fp64_frac',0),4), round(v.get('fp64_frac_of_sustained_peak') or 0,4)) for k,v in d['kernels'].items()}, 'burst', round(d['roofline']['fp64_peak_tflops_measured_live'],2), 'sustained', round(d['roofline']['fp64_peak_tflops_sustained_live'],2), d['clocks'])
"
done
