#!/bin/bash
# r06a: A/B of the round's per-op latency changes against the previous build (mecano_b200/variants/*.so), new large-angle tests
mkdir -p gpurun_out
timeout 600 python scripts/gpu_ab.py > gpurun_out/r06a_ab.jsonl 2> gpurun_out/r06a_ab.err; cat gpurun_out/r06a_ab.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "beyond or kernels_match or source_modes" > gpurun_out/r06a_pytest.log 2>&1; tail -5 gpurun_out/r06a_pytest.log
