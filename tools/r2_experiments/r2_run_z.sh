#!/bin/bash
# source-level ncu captures: the first two shade launches and the second traversal launch of a K3 frame
mkdir -p gpurun_out
export RT_B200_OPTIONS=pipeline_lanes=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others --no-verify --workload K3"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_wf_shade -s 12 -c 2 -f -o /tmp/src_shade $CMD > gpurun_out/ncu_r2z_shade.log 2>&1; tail -1 gpurun_out/ncu_r2z_shade.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_wf_traverse -s 17 -c 1 -f -o /tmp/src_trav $CMD > gpurun_out/ncu_r2z_trav.log 2>&1; tail -1 gpurun_out/ncu_r2z_trav.log
for n in shade trav; do
  ncu -i /tmp/src_$n.ncu-rep --page source --csv > gpurun_out/src_r2z_$n.sass.csv 2> gpurun_out/src_r2z_$n.err
  ncu -i /tmp/src_$n.ncu-rep --page source --csv --print-source cuda > gpurun_out/src_r2z_$n.cuda.csv 2>> gpurun_out/src_r2z_$n.err
  ncu -i /tmp/src_$n.ncu-rep --page raw --csv > gpurun_out/src_r2z_$n.raw.csv 2>> gpurun_out/src_r2z_$n.err
done
ls -la gpurun_out/src_r2z_* /tmp/*.ncu-rep
