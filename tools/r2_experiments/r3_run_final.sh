#!/bin/bash
# end of the round: whole GPU suite, smoke, every kernel class once, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest_final.log 2>&1; tail -3 gpurun_out/r3_pytest_final.log
timeout 600 python tools/sanitize_smoke.py > gpurun_out/r3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r3_smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], (d['roofline'].get('counted') or {}).get('frac'), (d['roofline'].get('issue') or {}).get('thread_instruction_frac'), d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
for k,v in (d.get('others') or {}).items(): print(k, {a:b for a,b in v.items() if a in ('mrays','ms','e2e')})
PY
