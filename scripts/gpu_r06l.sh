#!/bin/bash
# r06l: self re-arming work counter (w4) against memset per launch (w3); full GPU suite; a bench line
mkdir -p gpurun_out
AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py > gpurun_out/r06l_ab.jsonl 2> gpurun_out/r06l_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06l_ab.jsonl'):
    d = json.loads(l); print("%-12s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r06l_pytest.log 2>&1; tail -4 gpurun_out/r06l_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/r06l_bench.err > gpurun_out/r06l_bench.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r06l_bench.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], {k:(round(v['ms'],4), round(v.get('fp64_frac',0),4)) for k,v in d['kernels'].items()}, 'aba_source_modes', d['extras']['aba_source_modes']['ms'])
PY
