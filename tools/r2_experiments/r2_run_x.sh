#!/bin/bash
mkdir -p gpurun_out
for O in "tlas_ploc_radius=512" "tlas_ploc_radius=1024" "tlas_ploc_radius=4096"; do
  RT_B200_OPTIONS=$O timeout 300 python bench.py --steps 3 --warmup 3 --workload K4 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2x_K4.json 2> gpurun_out/r2x_K4.err
  tail -1 gpurun_out/r2x_K4.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$O', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done
