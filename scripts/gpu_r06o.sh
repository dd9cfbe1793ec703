#!/bin/bash
# r06o: ABA DESCEND runs split by SC (y0) or not (y1); record discard on / off with the four-double2 records
mkdir -p gpurun_out
V=mecano_b200/variants
AB_ALGOS=aba AB_ROUNDS=2 timeout 900 python scripts/gpu_ab.py y0:$V/y0.so y1_dunsplit:$V/y1_dunsplit.so y0_nodiscard:$V/y0.so:MECANO_B200_ABA_DISCARD=0 > gpurun_out/r06o_ab.jsonl 2> gpurun_out/r06o_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r06o_ab.jsonl'):
    d = json.loads(l); print("%-14s %-5s median %.4f min %.4f %s" % (d['tag'], d['algo'], d['ms_median'], d['ms_min'], d['sha']))
PY
