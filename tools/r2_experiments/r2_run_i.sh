#!/bin/bash
# one k_wf_traverse launch (bounce 1 + shadow 0 of a timed K3 frame) with source correlation
mkdir -p gpurun_out
export RT_B200_OPTIONS=pipeline_lanes=1
ncu --set full --clock-control none --import-source on -k regex:k_wf_traverse -s 13 -c 1 -f -o gpurun_out/prof_r2i_bounce python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others > gpurun_out/ncu_r2i.log 2>&1
tail -2 gpurun_out/ncu_r2i.log; ls -la gpurun_out/*.ncu-rep
