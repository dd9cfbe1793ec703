#!/bin/bash
# ncu --set full capture of one launch per algorithm.  Usage: bash scripts/gpu_prof.sh <tag> <algos, e.g. rnea,aba,crba> [n_states]
TAG=$1; ALGOS=${2:-rnea,aba,crba}; N=${3:-1048576}
mkdir -p gpurun_out
NA=$(echo $ALGOS | tr ',' '\n' | wc -l)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:thread_kernel|mb_spec_kernel" -s $NA -c $NA -f -o gpurun_out/${TAG}_prof \
   python scripts/prof_run.py $N 2 $ALGOS > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
