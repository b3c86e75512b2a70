#!/bin/bash
# final evidence of the round: smoke of the new paths, counted work, ncu summaries, the default bench line
mkdir -p gpurun_out
timeout 600 python tools/sanitize_smoke.py > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2f_smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
RT_B200_LIBNAME=librt_b200_count.so timeout 600 python tools/count_work.py > gpurun_out/r2f_counts.log 2>&1; echo "counts rc=$?"; cp profiles/work_counts.json gpurun_out/r2_work_counts.json
bash tools/r2_profile.sh K3:9:full K3headline:7:list K2:7:list K4:7:list K5:7:list K3glass:27:list > gpurun_out/r2f_profile.log 2>&1; tail -7 gpurun_out/r2f_profile.log
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], (d['roofline'].get('counted') or {}).get('frac'), (d['roofline'].get('issue') or {}).get('thread_instruction_frac'), d['cpu_baseline']['value'])
for k,v in (d.get('others') or {}).items(): print(k, {a:b for a,b in v.items() if a!='workload'})
PY
