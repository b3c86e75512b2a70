#!/bin/bash
mkdir -p gpurun_out
for C in 1 0; do
RT_B200_LIBNAME=librt_b200_count.so RT_B200_OPTIONS=pipeline_lanes=1,classify_rays=$C timeout 600 python - <<'PY' 2>&1 | tee -a gpurun_out/r2g_hist2.log
import sys, os; sys.path.insert(0,'.')
import bench, json
from metal4_raytracing_b200 import device
ctx = device.Context(0)
print("options", os.environ["RT_B200_OPTIONS"])
for name, mod in (("K3",1),("K3",8),("K3headline",1)):
    sc,u,seeds,w,h = bench.build_scene(name)
    rnd = device.Renderer(ctx, sc, w, h, seeds=seeds)
    rnd.draw(u, count_rays=True, tile_modulo=mod, tile_remainder=0); c = rnd.read_ray_counters(); wk = c["work"]
    tot_iters = wk["iterations_histogram"]["sum"]
    print(name, "slice", mod, "rays", c["rays"], "ray-iterations", tot_iters, "max", wk["iterations_histogram"]["max"], json.dumps(wk["tail"]), "tail mean", round(wk["tail"]["warp_iterations_sum"]/max(1,wk["tail"]["warps"]),1), "nodes/ray", round((wk["closest"]["nodes"]+wk["any"]["nodes"])/c["rays"],2))
    rnd.close()
PY
done
