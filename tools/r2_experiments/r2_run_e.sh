#!/bin/bash
# ncu full captures of one frame's wavefront launches: whole frame, rank-0-of-8 slice, headline (single lane).
# The reports are turned into raw-page CSV on the box (gpurun_out/ is limited to 64 MiB) and removed.
mkdir -p gpurun_out
export RT_B200_OPTIONS=pipeline_lanes=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
cap() { # tag skip count extra-args
  ncu --set full --clock-control none -k regex:k_wf_ -s $2 -c $3 -f -o /tmp/prof_$1 $CMD $4 > gpurun_out/ncu_r2e_$1.log 2>&1; tail -1 gpurun_out/ncu_r2e_$1.log
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_r2e_$1.raw.csv 2>/dev/null
}
cap s8 27 9 "--slice 8"
cap full 27 9 ""
cap head 21 7 "--workload K3headline"
ls -la gpurun_out/ | tail -8
