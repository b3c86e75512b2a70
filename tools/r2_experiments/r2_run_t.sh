#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest.log
run() { # tag options workload extra
  RT_B200_OPTIONS=$2 timeout 200 python bench.py --steps 5 --warmup 3 --workload $3 --no-others --no-cpu-baseline --no-e2e $4 > gpurun_out/r2t_$1.json 2> gpurun_out/r2t_$1.err; echo "$1 rc=$?"
}
for L in 1 2; do
  for WL in K3 K4 K3glass; do run ${WL}_l$L pipeline_lanes=$L $WL; done
done
for WL in K2 K3headline K5; do run ${WL}_auto pipeline_lanes=0 $WL; done
run K3s8_auto pipeline_lanes=0 K3 "--slice 8"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t_K*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
