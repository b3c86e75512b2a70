#!/bin/bash
mkdir -p gpurun_out
RT_B200_LIBNAME=librt_b200_count.so RT_B200_OPTIONS=pipeline_lanes=1 timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2g_hist.log
import sys; sys.path.insert(0,'.')
import bench, json
from metal4_raytracing_b200 import device
ctx = device.Context(0)
for name, mod in (("K3",1),("K3",8),("K3headline",1),("K4",1),("K3glass",1)):
    sc,u,seeds,w,h = bench.build_scene(name)
    rnd = device.Renderer(ctx, sc, w, h, seeds=seeds)
    rnd.draw(u, count_rays=True, tile_modulo=mod, tile_remainder=0); c = rnd.read_ray_counters(); wk = c["work"]
    print(name, "slice", mod, "rays", c["rays"], json.dumps(wk["iterations_histogram"]), "mean", round(wk["iterations_histogram"]["sum"]/c["rays"],2), json.dumps(wk["tail"]), "tail mean", round(wk["tail"]["warp_iterations_sum"]/max(1,wk["tail"]["warps"]),1))
    rnd.close()
PY
