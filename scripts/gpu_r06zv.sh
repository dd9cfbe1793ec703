#!/bin/bash
# r06zv: final session of the round -- the whole GPU suite, smoke() and the bench line as the driver runs it
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r06zv_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r06zv_pytest.log
tail -3 gpurun_out/r06zv_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r06zv_smoke.log 2>&1; echo "smoke exit $?"
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r06zv_bench.json 2> gpurun_out/r06zv_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r06zv_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],4), {k:(round(v['ms'],4), round(v.get('fp64_frac',0),4), round(v.get('fp64_frac_at_sampled_clock') or 0,4)) for k,v in d['kernels'].items()}, d['clocks'])
print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), round(d['e2e']['roofline']['frac'],3))
x=d['extras']
for k in ('e2e_dense_kept','e2e_packed'):
    print(k, round(x[k]['value']), round(x[k]['ms_per_step'],1), round(x[k]['pcie_d2h_frac'],3))
PY
