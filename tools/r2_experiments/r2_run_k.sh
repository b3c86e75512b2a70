#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
for WL in K3 K2 K3headline K4 K5; do
  timeout 300 python bench.py --steps 5 --warmup 3 --workload $WL --no-others --no-cpu-baseline > gpurun_out/r2k_$WL.json 2> gpurun_out/r2k_$WL.err; echo "$WL rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_K*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
bash tools/r2_profile.sh K3:9:full K3headline:7:list K2:7:list K4:7:list K5:7:list K3glass:27:list
