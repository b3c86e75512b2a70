import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from metal4_raytracing_b200 import _abi as A, device, scene
w, h = 255, 128
sc = scene.Scene()
m = sc.add_obj(os.path.join(scene.asset_dir(), "sphere.obj")); sc.add_instance(m)
p = sc.add_procedural("plane"); sc.add_instance(p, position=(0, -1, 0), scale=4.0)
sc.add_light(scene.make_light(A.LIGHT_POINT, position=(0, 5, 5), color=(9, 9, 9)))
u = scene.default_uniforms(w, h); u.lightCount, u.samplesPerPixel, u.maxBounces = 1, 1, 2
u.enableMotionAdaptiveSampling = u.enableMotionAdaptiveAccumulation = 0
u.camera = scene.orbit_camera(w, h, (0, 0, 0), 0.0, 0.0, 5.38); u.previousCamera = u.camera
seeds = np.zeros((h, w), np.uint32)
ctx = device.Context(0); rnd = device.Renderer(ctx, sc, w, h, seeds=seeds)
rnd.draw(u, want_ids=True); gid = rnd.read_ids()
orc = oracle.Oracle(sc); imgs = oracle.FrameImages(w, h, seeds); _, rid = orc.render(u, imgs, want_ids=True)
bad = (gid[..., :3] != rid[..., :3]).any(-1)
ys, xs = np.nonzero(bad)
print("mismatches", len(ys))
idx = sc.mesh_arrays(m)["submeshes"][0]; pos = sc.mesh_arrays(m)["positions"]
for y, x in list(zip(ys, xs))[:20]:
    g, r = gid[y, x], rid[y, x]
    print((x, y), "gpu", g[:3], np.array([g[3]], np.uint32).view(np.float32)[0], "orc", r[:3], np.array([r[3]], np.uint32).view(np.float32)[0])
    for pr in (g[2], r[2]):
        if pr != 0xFFFFFFFF and g[0] == 0: print("    tri", pr, pos[idx[pr]][:, :3].tolist())
