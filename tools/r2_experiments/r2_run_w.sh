#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2w_pytest.log
for WL in K3env; do
  timeout 300 python bench.py --steps 5 --warmup 3 --workload $WL --no-others --no-cpu-baseline > gpurun_out/r2w_$WL.json 2> gpurun_out/r2w_$WL.err; echo "$WL rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2w_K*.json')):
    d=json.loads(open(f).read().strip().split('\n')[-1])
    print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
PY
