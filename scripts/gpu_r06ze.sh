#!/bin/bash
# r06ze: host pipeline chunk size (MECANO_B200_HOST_CHUNK_MB) and slots on this box: e2e dense / packed
mkdir -p gpurun_out
for cfg in "128 3" "256 3" "512 3" "256 4" "64 4"; do
  set -- $cfg
  MECANO_B200_HOST_CHUNK_MB=$1 MECANO_B200_HOST_SLOTS=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r06ze_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d['e2e']; p=d['extras'].get('e2e_packed',{})
print('chunk_mb $1 slots $2', 'dense', round(e['value']), round(e['ms_per_step'],1), 'ms', round(e['roofline']['frac'],3), 'packed', round(p.get('value',0)), round(p.get('ms_per_step',0),1))
"
done
