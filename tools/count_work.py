#!/usr/bin/env python
"""Counts what the traversal actually fetches per ray, with the counter build of the library
(make -C metal4_raytracing_b200/csrc count; RT_B200_LIBNAME=librt_b200_count.so): node steps (80 B each), triangle
tests (48 B), instance entries (64 B record), separately for closest-hit and any-hit rays, one frame per workload.
Writes profiles/work_counts.json, which bench.py turns into roofline.counted.

  RT_B200_LIBNAME=librt_b200_count.so python tools/count_work.py [workload ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RT_B200_LIBNAME", "librt_b200_count.so")

import bench  # noqa: E402
from metal4_raytracing_b200 import device, scene  # noqa: E402

names = sys.argv[1:] or ["K3", "K3headline", "K3glass", "K2", "K4", "K5"]
ctx = device.Context(0)
out_path = os.path.join(ROOT, "profiles", "work_counts.json")
out = json.load(open(out_path)) if os.path.isfile(out_path) else {}
for name in names:
    sc, u, seeds, w, h = bench.build_scene(name)
    rnd = device.Renderer(ctx, sc, w, h, seeds=seeds)
    u.frameIndex = 0
    rnd.draw(u, count_rays=True)
    c = rnd.read_ray_counters()
    rnd.close()
    work = c.get("work")
    if not work:
        raise SystemExit("the loaded library is not the counter build (RT_B200_LIBNAME=librt_b200_count.so)")
    rays = c["rays"]
    nodes = work["closest"]["nodes"] + work["any"]["nodes"]
    tris = work["closest"]["triangles"] + work["any"]["triangles"]
    entries = work["closest"]["entries"] + work["any"]["entries"]
    out[name] = {
        "rays": rays, "closest_rays": c["closest"], "any_rays": c["any"],
        "nodes_per_ray": round(nodes / rays, 3), "triangles_per_ray": round(tris / rays, 3),
        "entries_per_ray": round(entries / rays, 3),
        "closest": {k: round(v / max(1, c["closest"]), 3) for k, v in work["closest"].items()},
        "any": {k: round(v / max(1, c["any"]), 3) for k, v in work["any"].items()},
        "bytes_per_ray": round((nodes * bench.B_NODE + tris * bench.B_TRIANGLE + entries * bench.B_INSTANCE) / rays, 1),
        "model_bytes_per_ray": bench.B_RAY[name],
        "source": "tools/count_work.py, counter build (-DRT_COUNT_WORK), frame 0 of the workload",
    }
    print(name, json.dumps(out[name]))
json.dump(out, open(out_path, "w"), indent=1)
ctx.close()
