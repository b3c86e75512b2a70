#!/bin/bash
mkdir -p gpurun_out
export RT_B200_OPTIONS=pipeline_lanes=1
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 13 -c 1 -f -o gpurun_out/prof_r2r_shade_c1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others > gpurun_out/ncu_r2r.log 2>&1
RT_B200_OPTIONS=pipeline_lanes=1,classify_rays=0 ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 13 -c 1 -f -o gpurun_out/prof_r2r_shade_c0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others >> gpurun_out/ncu_r2r.log 2>&1
ls -la gpurun_out/prof_r2r*
