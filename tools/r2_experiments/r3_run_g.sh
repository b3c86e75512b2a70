#!/bin/bash
# cooperative triangle stage re-measured at 7 / 8 CTAs per SM: always (coop), only when a lane has >= 2 triangles (hyb), with one pass (hyb1)
mkdir -p gpurun_out
run() { # lib workload extra
  RT_B200_LIBNAME=$1 timeout 300 python bench.py --steps 5 --warmup 3 --workload $2 $3 --no-others --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $3', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
}
RT_B200_LIBNAME=librt_b200_hyb.so timeout 600 python -m pytest tests -m gpu -x -q -k "full_size or golden or intersect or glass or instancing" 2>&1 | tail -2
for LIB in librt_b200.so librt_b200_coop.so librt_b200_hyb.so librt_b200_hyb1.so; do
  run $LIB K3 ""
  run $LIB K4 ""
  run $LIB K3headline ""
  run $LIB K3glass ""
done
