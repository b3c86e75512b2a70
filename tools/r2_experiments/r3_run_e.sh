#!/bin/bash
# shade kernel CTA shape (128 x 6 / 7 / 9 against 256 x 4), traversal at 9 CTAs per SM, refill / lanes options at the new occupancy
mkdir -p gpurun_out
run() { # lib options workload extra
  RT_B200_LIBNAME=$1 RT_B200_OPTIONS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --workload $3 $4 --no-others --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $3 $4', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
}
for LIB in librt_b200.so librt_b200_sh7.so librt_b200_sh6.so librt_b200_sh9.so; do
  run $LIB pipeline_lanes=0 K3 ""
  run $LIB pipeline_lanes=0 K2 ""
  run $LIB pipeline_lanes=0 K4 ""
  run $LIB pipeline_lanes=0 K3headline ""
done
run librt_b200_mb9.so blocks_per_sm=9 K3 ""
run librt_b200_mb9.so blocks_per_sm=9 K4 ""
run librt_b200.so pipeline_lanes=1 K3 ""
run librt_b200.so traversal_variant=2 K3 ""
run librt_b200.so traversal_variant=2 K4 ""
run librt_b200.so pipeline_lanes=3 K3 ""
