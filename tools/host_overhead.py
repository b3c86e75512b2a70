"""Host-side cost of one e2e frame (update + draw + async read-back) on a frame so small that the GPU never limits it."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metal4_raytracing_b200 import _abi as A, device, scene
w, h = 64, 32
sc, u, seed = scene.Scene.named("K3", w, h)
u.samplesPerPixel, u.maxBounces = 16, 3
ctx = device.Context(0)
rnd = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, seed))
first = rnd.read_image(A.TEXTURE_ACCUMULATION)
host = [ctx.pinned_array(first.shape, first.dtype) for _ in range(2)]
def loop(n, what):
    inflight = None
    t0 = time.perf_counter()
    for k in range(n):
        u.frameIndex = k
        if "u" in what: rnd.update()
        if "d" in what: rnd.draw(u)
        if "r" in what:
            if inflight is not None: ctx.download_wait(inflight)
            inflight = rnd.read_image_async(A.TEXTURE_ACCUMULATION, host[k & 1])
    if inflight is not None: ctx.download_wait(inflight)
    ctx.sync()
    return 1e3 * (time.perf_counter() - t0) / n
loop(20, "udr")
for what in ("d", "ud", "udr"):
    print(what, "ms/frame", round(loop(200, what), 4))
