#!/bin/bash
# r06zn: the bench line and the reference-arm line exactly as the driver runs them (--gpus 1 --steps 20 --warmup 5), final build
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r06zn_bench.json 2> gpurun_out/r06zn_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r06zn_bench_ref.json 2>> gpurun_out/r06zn_bench.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r06zn_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],4), {k:(round(v['ms'],4), round(v.get('fp64_frac',0),4), round(v.get('fp64_frac_at_sampled_clock') or 0,4)) for k,v in d['kernels'].items()}, d['clocks'])
print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), round(d['e2e']['roofline']['frac'],3))
r=json.loads(open('gpurun_out/r06zn_bench_ref.json').read().strip().splitlines()[-1]); print('ref', round(r['value']), r['config'].get('sample_states_per_step'))
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
