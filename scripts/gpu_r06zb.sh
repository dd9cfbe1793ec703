#!/bin/bash
# r06zb: the mass-matrix kernels keep their plain tile loop; full GPU suite; a bench line with extras
# batches of deep trees; a bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r06zb_pytest.log 2>&1; tail -4 gpurun_out/r06zb_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/r06zb_bench.err > gpurun_out/r06zb_bench.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r06zb_bench.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], {k:(round(v['ms'],4), round(v.get('fp64_frac',0),4)) for k,v in d['kernels'].items()})
print({k:round(v['ms'],4) for k,v in d['extras'].items() if isinstance(v,dict) and 'ms' in v})
PY
