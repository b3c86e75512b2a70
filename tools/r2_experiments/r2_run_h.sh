#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
run() { # tag lib workload extra
  RT_B200_LIBNAME=$2 timeout 300 python bench.py --steps 5 --warmup 3 --workload $3 --no-others --no-cpu-baseline $4 > gpurun_out/r2h_$1.json 2> gpurun_out/r2h_$1.err; echo "$1 rc=$?"
}
for WL in K3 K3headline K2 K5 K4; do
  run ${WL}_prev librt_b200_prev.so $WL
  run ${WL}_new librt_b200.so $WL
done
run K3s8_prev librt_b200_prev.so K3 "--slice 8 --no-e2e"
run K3s8_new librt_b200.so K3 "--slice 8 --no-e2e"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h_K*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
