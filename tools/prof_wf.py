import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metal4_raytracing_b200 import device, scene
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
name = sys.argv[2] if len(sys.argv) > 2 else "K3"
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mb = int(sys.argv[4]) if len(sys.argv) > 4 else 2
w = int(sys.argv[5]) if len(sys.argv) > 5 else 1920
h = int(sys.argv[6]) if len(sys.argv) > 6 else 1080
sc, u, seed = scene.Scene.named(name, w, h)
u.samplesPerPixel, u.maxBounces = spp, mb
ctx = device.Context(0); ctx.set_trace_mode(mode)
rnd = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, seed))
for f in range(2):
    u.frameIndex = f
    ctx.timer_begin(); rnd.draw(u); print("frame", f, ctx.timer_end(), "ms")
