#!/bin/bash
# traversal CTA shape after the TriSetup packing: 128 x 6 (default), 128 x 7 at 72 registers, 64 x 12, 64 x 14
mkdir -p gpurun_out
run() { # lib options workload extra
  RT_B200_LIBNAME=$1 RT_B200_OPTIONS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --workload $3 $4 --no-others --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $3 $4', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
}
for spec in "librt_b200.so blocks_per_sm=6" "librt_b200_mb7.so blocks_per_sm=7" "librt_b200_b64.so blocks_per_sm=12" "librt_b200_b64m14.so blocks_per_sm=14"; do
  set -- $spec
  run $1 $2 K3 ""
  run $1 $2 K3headline ""
  run $1 $2 K3 "--slice 8"
  run $1 $2,pipeline_lanes=2 K3 "--slice 8"
  run $1 $2 K4 ""
done
