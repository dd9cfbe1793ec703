#!/bin/bash
# r05d: welded-system GPU tests, ABA at 13 / 14 warps per SM (cfg 16 / 15), host topology probe, config-5 batch sweep on one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_welds.py -m gpu -q > gpurun_out/r05d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05d_pytest.log
tail -5 gpurun_out/r05d_pytest.log
for cfg in "" "aba=15" "aba=16" ""; do
  MECANO_B200_CFG="$cfg" AB_ALGO=aba timeout 300 python scripts/gpu_aba_ab.py child
done | tee gpurun_out/r05d_aba_cfg.txt
{
  echo "== lscpu"; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"
  echo "== numa nodes"; ls /sys/devices/system/node/ | grep node; for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  echo "== gpu pci numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then echo "$(basename $d) numa=$(cat $d/numa_node) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
  echo "== topo"; nvidia-smi topo -m
  echo "== libnuma"; ldconfig -p | grep -i numa; which numactl
  echo "== affinity"; python -c "import os; print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8])"
  echo "== meminfo"; grep -E "MemTotal|MemAvailable|Huge" /proc/meminfo
} > gpurun_out/r05d_topology.txt 2>&1
tail -30 gpurun_out/r05d_topology.txt
timeout 900 python scripts/gpu_config5.py gpurun_out/r05d_config5_1gpu_batch.jsonl batch > gpurun_out/r05d_config5.log 2>&1; tail -2 gpurun_out/r05d_config5.log | cut -c1-400
