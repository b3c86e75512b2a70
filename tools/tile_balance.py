"""Per-rank device time of the K3 bench frame when one GPU renders each rank's tile set in turn (load balance of the
interleaved-tile partition at N = 2, 4, 8)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metal4_raytracing_b200 import device, scene
w, h = 1920, 1080
sc, u, seed = scene.Scene.named("K3", w, h)
u.samplesPerPixel, u.maxBounces = 16, 3
ctx = device.Context(0)
rnd = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, seed))
for n in (2, 4, 8):
    times = []
    for r in range(n):
        best = 1e9
        for rep in range(3):
            u.frameIndex = rep
            ctx.timer_begin(); rnd.draw(u, tile_modulo=n, tile_remainder=r); best = min(best, ctx.timer_end())
        times.append(round(best, 3))
    print(n, times, "max/mean", round(max(times) / (sum(times) / n), 3), flush=True)
