#!/bin/bash
# r06zt: the whole GPU suite and the bench line as the driver runs it, after the centre-of-mass-only entry point and the
# ZEROS_PRESENT fix of the fused host step (extras: center_of_mass, centroidal_convective_term, e2e_dense_kept)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r06zt_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r06zt_pytest.log
tail -4 gpurun_out/r06zt_pytest.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r06zt_bench.json 2> gpurun_out/r06zt_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r06zt_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],4), {k:(round(v['ms'],4), round(v.get('fp64_frac',0),4)) for k,v in d['kernels'].items()}, d['clocks'])
print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), round(d['e2e']['roofline']['frac'],3))
x=d['extras']
for k in ('e2e_dense_kept','e2e_packed'):
    print(k, round(x[k]['value']), round(x[k]['ms_per_step'],1), round(x[k]['pcie_d2h_frac'],3), x[k].get('zero_entries_still_zero'))
for k in ('center_of_mass','centroidal_convective_term','crba_centroidal'):
    print(k, round(x[k]['ms'],3))
PY
