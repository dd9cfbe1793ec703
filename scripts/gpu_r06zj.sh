#!/bin/bash
# r06zj: the round's evidence session (scripts/gpu_session.sh: parity, bench + reference arm, ncu launch list, full captures), smoke(),
# memcheck / initcheck over every entry point
bash scripts/gpu_session.sh r06zj > gpurun_out/r06zj_session.log 2>&1
tail -4 gpurun_out/r06zj_pytest.log; head -c 600 gpurun_out/r06zj_bench.json; echo; tail -3 gpurun_out/r06zj_ncu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r06zj_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r06zj_smoke.log
timeout 1200 compute-sanitizer --tool memcheck python scripts/gpu_memcheck.py > gpurun_out/r06zj_memcheck.txt 2>&1; tail -3 gpurun_out/r06zj_memcheck.txt
timeout 1200 compute-sanitizer --tool initcheck python scripts/gpu_memcheck.py > gpurun_out/r06zj_initcheck.txt 2>&1; tail -3 gpurun_out/r06zj_initcheck.txt
