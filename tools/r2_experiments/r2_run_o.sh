#!/bin/bash
mkdir -p gpurun_out
METRICS=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_issued.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum
for C in 0 1; do
  RT_B200_OPTIONS=pipeline_lanes=1,classify_rays=$C timeout 600 ncu --metrics $METRICS --clock-control none -k regex:k_wf_ -s 36 -c 9 -f -o /tmp/prof_c$C python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others > gpurun_out/ncu_r2o_c$C.log 2>&1; tail -1 gpurun_out/ncu_r2o_c$C.log
  ncu -i /tmp/prof_c$C.ncu-rep --page raw --csv > gpurun_out/prof_r2o_c$C.raw.csv 2>/dev/null
done
