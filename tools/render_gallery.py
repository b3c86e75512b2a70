"""Renders small tonemapped PNGs of the named scenes into gpurun_out/gallery_*.png (visual sanity check)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metal4_raytracing_b200 import _abi as A, device, scene
ctx = device.Context(0)
w, h = 640, 360
for name, spp, mb, frames, env in (("K1", 8, 2, 4, False), ("K2", 8, 3, 4, False), ("K2tex", 8, 3, 4, False), ("K4", 4, 2, 4, False),
                                   ("K5", 8, 3, 6, False), ("K3", 8, 3, 4, True), ("appscene", 8, 3, 4, False)):
    try:
        sc, u, seed = scene.Scene.named(name, w, h)
    except Exception as e:
        print("skip", name, e); continue
    u.samplesPerPixel, u.maxBounces, u.accumulationWeight = spp, mb, 0.7
    rnd = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, seed))
    if env:
        rnd.set_environment(scene.procedural_sky(512, 256), 0.6)
    for f in range(frames):
        u.frameIndex = f
        if f and name == "K5":
            sc.animate(0.4 + f / 60.0); rnd.update()
        rnd.draw(u)
    img = ctx.tonemap(rnd.image_info(A.TEXTURE_ACCUMULATION))
    scene.write_png(f"gpurun_out/gallery_{name}{'_env' if env else ''}.png", img)
    print(name, img.mean())
    rnd.close()
ctx.close()
