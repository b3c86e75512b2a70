python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 1 0; do python tools/tune.py "{\"fuse_traversal\": $v}" 2>&1 | tail -1; done | tee gpurun_out/tune19.log
for v in 1 0; do RT_B200_OPTIONS="fuse_traversal=$v" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"; done 2>&1 | tee gpurun_out/fuse_bench.log
python tools/render_gallery.py
