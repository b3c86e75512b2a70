#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
run() { # tag options workload extra
  RT_B200_OPTIONS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --workload $3 --no-others --no-cpu-baseline --no-e2e $4 > gpurun_out/r2n_$1.json 2> gpurun_out/r2n_$1.err; echo "$1 rc=$?"
}
for C in 0 1; do
  for WL in K3 K2 K3headline K5 K3glass; do run ${WL}_c$C classify_rays=$C $WL; done
  run K3s8_c$C classify_rays=$C K3 "--slice 8"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2n_K*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
