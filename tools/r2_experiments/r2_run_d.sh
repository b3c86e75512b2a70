#!/bin/bash
mkdir -p gpurun_out
run() { # tag options workload-args
  RT_B200_OPTIONS=$2 timeout 300 python bench.py $3 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2d_$1.json 2> gpurun_out/r2d_$1.err; echo "$1 rc=$?"
}
S8="--steps 10 --warmup 3 --slice 8"; HD="--steps 20 --warmup 3 --workload K3headline"; K3="--steps 5 --warmup 3"
run s8_l2_b3 pipeline_lanes=2,blocks_per_sm=3 "$S8"
run s8_l2_b3_st pipeline_lanes=2,blocks_per_sm=3,lane_stagger=1 "$S8"
run s8_l2_b4_st pipeline_lanes=2,blocks_per_sm=4,lane_stagger=1 "$S8"
run s8_l3_b2_st pipeline_lanes=3,blocks_per_sm=2,lane_stagger=1 "$S8"
run s8_l2_b6_st pipeline_lanes=2,blocks_per_sm=6,lane_stagger=1 "$S8"
run s8_l1_b3 pipeline_lanes=1,blocks_per_sm=3 "$S8"
run hd_l2_b3 pipeline_lanes=2,blocks_per_sm=3 "$HD"
run hd_l2_b3_st pipeline_lanes=2,blocks_per_sm=3,lane_stagger=1 "$HD"
run hd_l1_b3 pipeline_lanes=1,blocks_per_sm=3 "$HD"
run hd_l1_b4 pipeline_lanes=1,blocks_per_sm=4 "$HD"
run K3_l2_b3_st pipeline_lanes=2,blocks_per_sm=3,lane_stagger=1 "$K3"
run K3_l2_b6_st pipeline_lanes=2,blocks_per_sm=6,lane_stagger=1 "$K3"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
