
for v in 0 1 2; do RT_B200_OPTIONS="sort_rays=$v" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"; done 2>&1 | tee gpurun_out/sort_bench.log
