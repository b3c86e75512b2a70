#!/bin/bash
mkdir -p gpurun_out
run() { # tag lib workload extra
  RT_B200_LIBNAME=$2 timeout 300 python bench.py --steps 5 --warmup 3 --workload $3 --no-others --no-cpu-baseline --no-e2e $4 > gpurun_out/r2p_$1.json 2> gpurun_out/r2p_$1.err; echo "$1 rc=$?"
}
for V in "" _ns; do
  for WL in K3 K2 K3headline K5; do run ${WL}$V librt_b200$V.so $WL; done
  run K3s8$V librt_b200$V.so K3 "--slice 8"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p_K*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
