#!/bin/bash
mkdir -p gpurun_out
METRICS=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
export RT_B200_OPTIONS=pipeline_lanes=1
for S in 1 8; do
  timeout 600 ncu --metrics $METRICS --clock-control none -k regex:k_wf_ -s 36 -c 9 -f -o /tmp/prof_s$S python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others --slice $S > gpurun_out/ncu_r2s_$S.log 2>&1; tail -1 gpurun_out/ncu_r2s_$S.log
  ncu -i /tmp/prof_s$S.ncu-rep --page raw --csv > gpurun_out/prof_r2s_s$S.raw.csv 2>/dev/null
done
