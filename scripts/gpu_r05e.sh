#!/bin/bash
# r05e: the round's evidence session (scripts/gpu_session.sh: parity, bench + reference arm, ncu launch list, full captures), smoke(),
# memcheck / initcheck over every entry point with the round-2 kernels, config-5 batch sweep on one GPU
bash scripts/gpu_session.sh r05e > gpurun_out/r05e_session.log 2>&1
tail -4 gpurun_out/r05e_pytest.log; head -c 700 gpurun_out/r05e_bench.json; echo; tail -3 gpurun_out/r05e_ncu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r05e_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r05e_smoke.log
timeout 1200 compute-sanitizer --tool memcheck python scripts/gpu_memcheck.py > gpurun_out/r05e_memcheck.txt 2>&1; tail -3 gpurun_out/r05e_memcheck.txt
timeout 1200 compute-sanitizer --tool initcheck python scripts/gpu_memcheck.py > gpurun_out/r05e_initcheck.txt 2>&1; tail -3 gpurun_out/r05e_initcheck.txt
rm -f gpurun_out/r05e_config5_1gpu_batch.jsonl
timeout 900 python scripts/gpu_config5.py gpurun_out/r05e_config5_1gpu_batch.jsonl batch > gpurun_out/r05e_config5.log 2>&1; tail -2 gpurun_out/r05e_config5.log | cut -c1-500
