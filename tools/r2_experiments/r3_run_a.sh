#!/bin/bash
# cooperative triangle stage: parity tests on the default build, then K3 / headline / K4 / slice 8 on three builds
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; tail -3 gpurun_out/r3a_pytest.log
for LIB in librt_b200.so librt_b200_nocoop.so librt_b200_coop1.so; do
  for W in K3 K3headline K4 K2; do
    RT_B200_LIBNAME=$LIB timeout 300 python bench.py --steps 5 --warmup 3 --workload $W --no-others --no-cpu-baseline --no-e2e > gpurun_out/r3a_${LIB}_$W.json 2> gpurun_out/r3a_${LIB}_$W.err
    tail -1 gpurun_out/r3a_${LIB}_$W.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$LIB $W', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
  done
  RT_B200_LIBNAME=$LIB timeout 300 python bench.py --steps 5 --warmup 3 --workload K3 --slice 8 --no-others --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$LIB K3 slice8', d['value'], d['ms_per_step'])"
done
