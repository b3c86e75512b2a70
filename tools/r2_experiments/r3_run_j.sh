#!/bin/bash
# shade kernel specialised for plain PBR + environment light (kPlain = 2) against the general build (envgen)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_pytest.log 2>&1; tail -2 gpurun_out/r3j_pytest.log
run() { # lib workload
  RT_B200_LIBNAME=$1 timeout 300 python bench.py --steps 5 --warmup 3 --workload $2 --no-others --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
}
for LIB in librt_b200.so librt_b200_envgen.so librt_b200.so librt_b200_envgen.so; do run $LIB K3env; done
run librt_b200.so K3
