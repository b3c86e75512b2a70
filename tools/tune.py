"""Times the wavefront layout under tuning options; all settings must give identical images."""
import sys, os, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metal4_raytracing_b200 import _abi as A, device, scene
import oracle

def parity(opts):
    w, h = 320, 200
    sc, u, seed = scene.Scene.named("K3small", w, h, assets=None); u.samplesPerPixel, u.maxBounces = 2, 3
    seeds = scene.seed_image(w, h, seed)
    orc = oracle.Oracle(sc); imgs = oracle.FrameImages(w, h, seeds); _, rid = orc.render(u, imgs, want_ids=True)
    ctx = device.Context(0)
    for k, v in opts.items(): ctx.set_option(k, v)
    rnd = device.Renderer(ctx, sc, w, h, seeds=seeds); rnd.draw(u, want_ids=True)
    ok = np.array_equal(rnd.read_image(0).view(np.uint16), imgs.output.view(np.uint16)) and np.array_equal(rnd.read_ids(), rid)
    rnd.close(); ctx.close()
    return bool(ok)

def timeit(name, w, h, spp, mb, opts, frames=4):
    sc, u, seed = scene.Scene.named(name, w, h); u.samplesPerPixel, u.maxBounces = spp, mb
    ctx = device.Context(0)
    for k, v in opts.items(): ctx.set_option(k, v)
    rnd = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, seed))
    ts = []
    for f in range(frames):
        u.frameIndex = f
        ctx.timer_begin(); rnd.draw(u, count_rays="accumulate" if f else True); ts.append(ctx.timer_end())
    rays = rnd.read_ray_counters()["rays"] / frames
    img = rnd.read_image(0)
    ctx.kernel_timing(True)
    for f in range(frames, frames + 2):
        u.frameIndex = f
        rnd.draw(u)
    kt = {k: round(v[0] / 2, 3) for k, v in ctx.kernel_times().items()}
    rnd.close(); ctx.close()
    return min(ts[1:]), rays, img, kt

if __name__ == "__main__":
    grid = [dict(trace_mode=0)] + [dict(trace_mode=1, traversal_variant=v, blocks_per_sm=b) for v in (0, 1, 2) for b in (2, 4, 8)]
    if len(sys.argv) > 1: grid = [json.loads(a) for a in sys.argv[1:]]
    ref = {}
    for opts in grid:
        row = {"opts": opts, "parity": parity(opts)}
        for (name, spp, mb) in (("K3", 1, 2), ("K3", 4, 3), ("K2", 4, 2), ("K4", 2, 2)):
            ms, rays, img, kt = timeit(name, 1920, 1080, spp, mb, opts)
            key = (name, spp, mb)
            if key not in ref: ref[key] = img
            row[f"{name}_{spp}_{mb}"] = {"ms": round(ms, 3), "mrays": round(rays / ms / 1e3), "same": bool(np.array_equal(img.view(np.uint16), ref[key].view(np.uint16))), "k": kt}
        print(json.dumps(row), flush=True)
