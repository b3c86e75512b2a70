#!/bin/bash
# GPU tests with the new TLAS leaf order / re-test, work counts, K4 / K3 / K5 before (librt_b200_prev.so) and after
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log
RT_B200_LIBNAME=librt_b200_count.so timeout 600 python tools/count_work.py > gpurun_out/r2f_counts.log 2>&1; echo "counts rc=$?"; cat gpurun_out/r2f_counts.log | cut -c1-400
cp profiles/work_counts.json gpurun_out/r2f_work_counts.json
run() { # tag lib workload
  RT_B200_LIBNAME=$2 timeout 300 python bench.py --steps 5 --warmup 3 --workload $3 --no-others --no-cpu-baseline > gpurun_out/r2f_$1.json 2> gpurun_out/r2f_$1.err; echo "$1 rc=$?"
}
for WL in K4 K3 K5 K2; do
  run ${WL}_prev librt_b200_prev.so $WL
  run ${WL}_new librt_b200.so $WL
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_K*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
# BVH quality: PLOC radius vs frame time and counted work (K3)
for R in 8 32 64 128; do
  RT_B200_OPTIONS=ploc_radius=$R timeout 300 python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2f_K3_ploc$R.json 2> gpurun_out/r2f_K3_ploc$R.err
  tail -1 gpurun_out/r2f_K3_ploc$R.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ploc_radius $R', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
  RT_B200_OPTIONS=ploc_radius=$R RT_B200_LIBNAME=librt_b200_count.so timeout 300 python - <<PY
import sys; sys.path.insert(0,'.')
import bench
from metal4_raytracing_b200 import device
ctx = device.Context(0)
sc,u,seeds,w,h = bench.build_scene("K3")
rnd = device.Renderer(ctx, sc, w, h, seeds=seeds)
rnd.draw(u, count_rays=True); c = rnd.read_ray_counters(); wk = c["work"]
info = ctx.as_info(rnd.blas_id(0))
print("  sah", round(info.sahCost,2), "levels", info.levelCount, "nodes", info.wideNodeCount, "nodes/ray", round((wk["closest"]["nodes"]+wk["any"]["nodes"])/c["rays"],2), "tris/ray", round((wk["closest"]["triangles"]+wk["any"]["triangles"])/c["rays"],2))
PY
done
