#!/bin/bash
# ray prefetch ring (cp.async): parity tests with the ring in every launch (pf7) and in the default build, then timings
mkdir -p gpurun_out
RT_B200_LIBNAME=librt_b200_pf7.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_pytest_pf7.log 2>&1; tail -3 gpurun_out/r3d_pytest_pf7.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_pytest.log 2>&1; tail -3 gpurun_out/r3d_pytest.log
run() { # lib workload extra
  RT_B200_LIBNAME=$1 timeout 300 python bench.py --steps 5 --warmup 3 --workload $2 $3 --no-others --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 $3', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
}
for LIB in librt_b200.so librt_b200_nopf.so librt_b200_pf7.so librt_b200_pf8.so; do
  run $LIB K3 ""
  run $LIB K3headline ""
  run $LIB K3 "--slice 8"
  run $LIB K4 ""
  run $LIB K2 ""
done
