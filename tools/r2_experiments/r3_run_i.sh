#!/bin/bash
# last sweeps: refill check every 1 / 3 iterations at 8 CTAs per SM; cooperative closest-hit stage in the real-TLAS kernels only (K4)
mkdir -p gpurun_out
run() { # lib workload extra
  RT_B200_LIBNAME=$1 timeout 300 python bench.py --steps 5 --warmup 3 --workload $2 $3 --no-others --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $3', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
}
for LIB in librt_b200.so librt_b200_spc3.so librt_b200_spc1.so; do
  run $LIB K3 ""
  run $LIB K4 ""
  run $LIB K3headline ""
done
for LIB in librt_b200_creal.so librt_b200_creal2.so librt_b200.so librt_b200_creal.so; do
  run $LIB K4 ""
done
RT_B200_LIBNAME=librt_b200_creal.so timeout 300 python -m pytest tests -m gpu -x -q -k "instancing or tlas_refit or intersect" 2>&1 | tail -2
