"""Small end-to-end run for compute-sanitizer: every kernel class once (build, PLOC, refit, skin, palette, trace,
shade, shadow, resolve, tonemap, environment, tiles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metal4_raytracing_b200 import _abi as A, device, scene
w, h = 96, 64
ctx = device.Context(0)
for name, kw in (("K3small", {}), ("K5small", {"gpu_skeleton": True}), ("K4small", {})):
    try:
        sc, u, seed = scene.Scene.named(name, w, h)
    except Exception as e:
        print("skip", name, e); continue
    u.samplesPerPixel, u.maxBounces = 3, 3
    u.enableDenoiseGBuffer = 1
    rnd = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, seed), **kw)
    rnd.set_environment(scene.procedural_sky(64, 32), 0.5)
    for f in range(3):
        u.frameIndex = f
        if f:
            sc.animate(f / 60.0); rnd.update()
        rnd.draw(u, want_ids=True, count_rays=True)
    img = ctx.tonemap(rnd.image_info(A.TEXTURE_ACCUMULATION))
    print(name, rnd.read_ray_counters(), int(img.sum()))
    ctx.set_trace_mode(0); rnd.draw(u); ctx.set_trace_mode(1)
    rnd.close()
ctx.close()
print("done")
