#!/bin/bash
# Round 2, first GPU call: the whole GPU test suite, then the bench with 1 / 2 / 3 pipeline lanes (K3 and the 1-spp headline frame)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
for L in 2 1 3; do
  RT_B200_OPTIONS=pipeline_lanes=$L timeout 300 python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/r2a_bench_K3_l$L.json 2> gpurun_out/r2a_bench_K3_l$L.err; echo "K3 lanes=$L rc=$?"
  RT_B200_OPTIONS=pipeline_lanes=$L timeout 300 python bench.py --steps 20 --warmup 5 --workload K3headline --no-cpu-baseline > gpurun_out/r2a_bench_head_l$L.json 2> gpurun_out/r2a_bench_head_l$L.err; echo "headline lanes=$L rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
