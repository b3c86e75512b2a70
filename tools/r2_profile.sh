#!/bin/bash
# ncu captures of one frame's wavefront launches per workload -> gpurun_out/prof_r2_<workload>.raw.csv (tools/make_profiles_r2.py).
# K3 gets --set full (the roofline's `traffic` comes from it), the others a metric list that needs few replay passes.
mkdir -p gpurun_out
export RT_B200_OPTIONS=pipeline_lanes=1
METRICS=gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
cap() { # workload launches-per-frame mode
  local WL=$1 N=$2 MODE=$3
  local CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others --workload $WL"
  if [ "$MODE" = full ]; then SEL="--set full"; else SEL="--metrics $METRICS"; fi
  timeout 900 ncu $SEL --clock-control none -k regex:k_wf_ -s $((4 * N)) -c $N -f -o /tmp/prof_$WL $CMD > gpurun_out/ncu_r2_$WL.log 2>&1; tail -1 gpurun_out/ncu_r2_$WL.log
  ncu -i /tmp/prof_$WL.ncu-rep --page raw --csv > gpurun_out/prof_r2_$WL.raw.csv 2>/dev/null
}
for spec in "$@"; do cap ${spec%%:*} $(echo $spec | cut -d: -f2) $(echo $spec | cut -d: -f3); done
ls -la gpurun_out/prof_r2_*.raw.csv
