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
    for key, value in (("pipeline_lanes", 2), ("classify_rays", 0), ("pipeline_lanes", 0), ("classify_rays", 1)):
        ctx.set_option(key, value); rnd.draw(u)
    rnd.close()
    # round 2 paths: TLAS rebuild by one CTA vs refit, sample partition (fp32 shares), glass (all segment passes launched)
    reb = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, seed), fp32=True, tlas_rebuild=True, **kw)
    u.enableMotionAdaptiveSampling = u.enableMotionAdaptiveAccumulation = 0
    for r in range(2):
        reb.update(); reb.draw(u, sample_modulo=2, sample_remainder=r)
    reb.close()
# a TLAS over many instances (one-CTA PLOC + collapse, then refits) and a glass material
rng = np.random.default_rng(3)
sc = scene.Scene()
ball = sc.add_procedural("uvsphere", 10, 8)
m = sc.get_material(ball); m.refractionIndex, m.opacity = 1.5, 0.1; sc.set_material(ball, 0, m)
for i in range(300):
    sc.add_instance(ball, tuple(rng.uniform(-3, 3, 3)), (0, float(rng.uniform(0, 6)), 0), 0.15)
sc.add_instance(sc.add_procedural("plane"), (0, -3.2, 0), (0, 0, 0), 8.0)
sc.default_lights()
u = scene.default_uniforms(w, h)
u.camera = scene.orbit_camera(w, h, (0, 0, 0), 0.4, 0.3, 9.0, 45.0); u.previousCamera = u.camera
u.samplesPerPixel, u.maxBounces, u.lightCount = 2, 3, sc.desc().lightCount
for rebuild in (False, True):
    rnd = device.Renderer(ctx, sc, w, h, seeds=scene.seed_image(w, h, 5), tlas_rebuild=rebuild)
    for f in range(3):
        u.frameIndex = f
        if f:
            for i in range(300):
                sc.set_instance_transform(i, tuple(rng.uniform(-3, 3, 3)), (0, 0.1 * f, 0), 0.15)
            rnd.update()
        rnd.draw(u, count_rays=True)
    print("swarm rebuild" if rebuild else "swarm refit", rnd.read_ray_counters(), ctx.as_info(rnd.tlas_id()).wideNodeCount)
    # rt_intersect on the same TLAS: closest and any hit, un-normalised directions, tmin / tmax windows
    rays = np.zeros((1000, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-6, 6, (1000, 3)); rays[:, 4:7] = rng.uniform(-1, 1, (1000, 3)) - 0.2 * rays[:, 0:3]
    rays[:, 3] = 0.01; rays[:, 7] = np.where(rng.random(1000) < 0.5, 2.0, np.inf)
    hc, ha = ctx.intersect(rnd.tlas_id(), rays), ctx.intersect(rnd.tlas_id(), rays, any_hit=True)
    assert np.array_equal(np.isfinite(hc["t"]), np.isfinite(ha["t"]))
    print("intersect", int(np.isfinite(hc["t"]).sum()), "of 1000 rays hit")
    rnd.close()
ctx.close()
print("done")
