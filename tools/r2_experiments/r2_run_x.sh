#!/bin/bash
mkdir -p gpurun_out
for R in 8 32 64 128; do
  RT_B200_OPTIONS=ploc_radius=$R timeout 300 python bench.py --steps 3 --warmup 3 --workload K4 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2x_K4_ploc$R.json 2> gpurun_out/r2x_K4_ploc$R.err
  tail -1 gpurun_out/r2x_K4_ploc$R.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ploc_radius $R', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done
