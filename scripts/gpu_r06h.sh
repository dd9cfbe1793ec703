#!/bin/bash
# r06h: pass four of the source-mode forward dynamics inside the ABA kernel (fused) against the two-launch path; parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_welds.py -x -q -k "source_modes or welds or fixed or ignored or kernels_match" > gpurun_out/r06h_pytest.log 2>&1; tail -4 gpurun_out/r06h_pytest.log
for f in 1 0; do
  MECANO_B200_ABA_P4_FUSED=$f timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/r06h_bench_$f.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fused=$f', 'step ms', d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, 'aba_source_modes', d['extras']['aba_source_modes']['ms'], 'rnea_byproducts', d['extras']['rnea_byproducts']['ms'])
"
done
