"""Megakernel vs wavefront: bit-equality of the outputs and device time per frame on several workloads."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metal4_raytracing_b200 import _abi as A, device, scene

def run(name, w, h, spp, mb, frames=3, glass=False):
    sc, u, seed = scene.Scene.named(name, w, h)
    u.samplesPerPixel, u.maxBounces = spp, mb
    if glass:
        m = sc.get_material(0); m.baseColor.set(0.95, 0.98, 1.0); m.refractionIndex, m.opacity = 1.52, 0.08
        sc.set_material(0, 0, m)
    seeds = scene.seed_image(w, h, seed)
    out = {"scene": name + ("+glass" if glass else ""), "size": [w, h], "spp": spp, "mb": mb}
    imgs = {}
    for mode in (0, 1):
        ctx = device.Context(0); ctx.set_trace_mode(mode)
        rnd = device.Renderer(ctx, sc, w, h, seeds=seeds)
        times = []
        for f in range(frames):
            u.frameIndex = f
            ctx.timer_begin(); rnd.draw(u, count_rays=True, want_ids=(f == 0)); ms = ctx.timer_end()
            times.append(ms)
            rays = rnd.read_ray_counters()
        imgs[mode] = (rnd.read_image(0), rnd.read_image(A.TEXTURE_DEPTH), rnd.read_image(A.TEXTURE_MOTION), rnd.read_ids())
        out[f"mode{mode}_ms"] = [round(t, 3) for t in times]
        out[f"mode{mode}_mrays"] = round(rays["rays"] / min(times[1:]) / 1e3, 1)
        out[f"mode{mode}_rays"] = rays["rays"]
        rnd.close(); ctx.close()
    out["equal"] = [bool(np.array_equal(a.view(np.uint8), b.view(np.uint8))) for a, b in zip(imgs[0], imgs[1])]
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    run("K3small", 256, 256, 2, 3)
    run("K3small", 256, 256, 2, 3, glass=True)
    run("K3", 1920, 1080, 1, 2)
    run("K3", 1920, 1080, 4, 3)
    run("K2", 1920, 1080, 4, 2)
    run("K4", 1920, 1080, 2, 2)
    run("K3", 1920, 1080, 2, 3, glass=True)
