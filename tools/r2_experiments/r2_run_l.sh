#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.log
timeout 900 python tools/fast_shade.py > gpurun_out/r2l_fast_shade.log 2>&1; echo "fast_shade rc=$?"; tail -22 gpurun_out/r2l_fast_shade.log
cp profiles/r2_fast_shade.md gpurun_out/r2_fast_shade.md
rm -f gpurun_out/fast_shade_*.npz
timeout 900 python bench.py > gpurun_out/r2l_bench_default.json 2> gpurun_out/r2l_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2l_bench_default.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline'].get('counted'), d['roofline'].get('issue',{}).get('thread_instruction_frac'))
for k,v in (d.get('others') or {}).items(): print(k, v)
PY
