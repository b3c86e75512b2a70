import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metal4_raytracing_b200 import _abi as A, device, scene
w, h = 64, 32
sc, u, seed = scene.Scene.named("K3", w, h)
u.samplesPerPixel, u.maxBounces = 16, 3
ctx = device.Context(0)
rnd = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, seed))
for k in range(3):
    u.frameIndex = k; rnd.draw(u)
ctx.sync()
# host enqueue time vs total
t0 = time.perf_counter()
for k in range(100):
    u.frameIndex = k; rnd.draw(u)
t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
print("enqueue ms/frame", 10 * (t1 - t0), "total ms/frame", 10 * (t2 - t0))
ctx.kernel_timing(True)
for k in range(20):
    u.frameIndex = k; rnd.draw(u)
print({k: (round(v[0] / 20, 4), v[1] / 20) for k, v in ctx.kernel_times().items()})
