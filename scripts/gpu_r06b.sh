#!/bin/bash
# r06b: start stagger between the warps of a scheduler (MECANO_B200_STAGGER_NS sweep), next-op record carry (RNEA)
mkdir -p gpurun_out
V=mecano_b200/variants
AB_ROUNDS=1 timeout 1200 python scripts/gpu_ab.py base:$V/base.so s0:$V/v1.so s500:$V/v1.so:MECANO_B200_STAGGER_NS=500 s1000:$V/v1.so:MECANO_B200_STAGGER_NS=1000 \
   s2000:$V/v1.so:MECANO_B200_STAGGER_NS=2000 s4000:$V/v1.so:MECANO_B200_STAGGER_NS=4000 s8000:$V/v1.so:MECANO_B200_STAGGER_NS=8000 \
   s16000:$V/v1.so:MECANO_B200_STAGGER_NS=16000 carry:$V/v2carry.so carry_s2000:$V/v2carry.so:MECANO_B200_STAGGER_NS=2000 \
   > gpurun_out/r06b_ab.jsonl 2> gpurun_out/r06b_ab.err
cut -c1-40,90-175 gpurun_out/r06b_ab.jsonl
