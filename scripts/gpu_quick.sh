#!/bin/bash
# Quick GPU check: parity tests + one bench line.  Usage: bash scripts/gpu_quick.sh <tag> [extra bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g states/s  ms/step %.3f  e2e %s" % (d["value"], d["ms_per_step"], d.get("e2e") and d["e2e"]["value"]))
for k,v in d["kernels"].items(): print(k, "ms %.3f  states/s %.4g  GB/s %.0f  block %d regs %d smem %d bps %d  fp64frac %s" % (v["ms"], v["states_per_s"], v["achieved_gbs"], v["block_threads"], v["regs"], v["smem_bytes"], v["blocks_per_sm"], v.get("fp64_frac")))
print(d["clocks"])
PY
