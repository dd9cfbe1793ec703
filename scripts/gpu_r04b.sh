#!/bin/bash
# r04b: parity tests with the 3-double2 ABA records, A/B of pass-three order x L2 discard (timing + DRAM bytes), host pipeline sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r04b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r04b_pytest.log
tail -4 gpurun_out/r04b_pytest.log
timeout 600 python scripts/gpu_aba_ab.py > gpurun_out/r04b_aba_ab.jsonl 2> gpurun_out/r04b_aba_ab.err
cat gpurun_out/r04b_aba_ab.jsonl
for fwd in 1 0; do for disc in 0 1; do
  export MECANO_B200_ABA_DISCARD=$disc; if [ $fwd = 1 ]; then export MECANO_B200_ABA_P3_FORWARD=1; else unset MECANO_B200_ABA_P3_FORWARD; fi
  AB_REPS=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:thread_kernel --csv \
     --log-file gpurun_out/r04b_aba_dram_f${fwd}_d${disc}.csv python scripts/gpu_aba_ab.py child > /dev/null 2>&1
  echo "fwd=$fwd disc=$disc"; grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/r04b_aba_dram_f${fwd}_d${disc}.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done; done
unset MECANO_B200_ABA_DISCARD MECANO_B200_ABA_P3_FORWARD
timeout 900 python scripts/gpu_host_pipe.py > gpurun_out/r04b_host_pipe.jsonl 2> gpurun_out/r04b_host_pipe.err
cat gpurun_out/r04b_host_pipe.jsonl
