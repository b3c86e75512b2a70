"""First-contact GPU diagnostics: parity of ids / radiance / skinning for several small scenes + timings."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metal4_raytracing_b200 import _abi as A, device, scene
import oracle


def compare(name, w, h, spp=None, mb=None, frames=1, assets="auto", oracle_too=True, animate=False):
    sc, u, seed = scene.Scene.named(name, w, h, assets=assets)
    if spp: u.samplesPerPixel = spp
    if mb: u.maxBounces = mb
    seeds = scene.seed_image(w, h, seed)
    ctx = device.Context(0)
    t0 = time.time()
    rnd = device.Renderer(ctx, sc, w, h, seeds=seeds)
    ctx.sync()
    tb = time.time() - t0
    orc = oracle.Oracle(sc) if oracle_too else None
    imgs = oracle.FrameImages(w, h, seeds) if oracle_too else None
    res = {"scene": name, "size": [w, h], "spp": u.samplesPerPixel, "mb": u.maxBounces, "gpu_setup_s": round(tb, 3)}
    for f in range(frames):
        u.frameIndex = f
        if animate and f > 0:
            sc.animate(f / 60.0)
            rnd.update()
            if orc: orc.update()
        ctx.timer_begin()
        rnd.draw(u, want_ids=True, count_rays=True)
        ms = ctx.timer_end()
        g = rnd.read_image(A.TEXTURE_ACCUMULATION).astype(np.float32)
        gid = rnd.read_ids()
        rays = rnd.read_ray_counters()
        res[f"f{f}_ms"] = round(ms, 3)
        res[f"f{f}_mrays"] = round(rays["rays"] / ms / 1e3, 1)
        res[f"f{f}_rays"] = rays
        if orc:
            t0 = time.time()
            st, rid = orc.render(u, imgs, want_ids=True)
            res[f"f{f}_oracle_s"] = round(time.time() - t0, 2)
            r = imgs.output.astype(np.float32)
            imgs.swap()
            mism = (gid[..., :3] != rid[..., :3]).any(-1)
            tm = (gid[..., 3] != rid[..., 3]) & ~mism
            res[f"f{f}_id_mismatch"] = float(mism.mean())
            res[f"f{f}_t_mismatch"] = float(tm.mean())
            res[f"f{f}_rays_equal"] = (st["closest"] == rays["closest"], st["any"] == rays["any"], st["hits"] == rays["hits"])
            d = g[..., :3] - r[..., :3]
            res[f"f{f}_relrmse"] = float(np.sqrt((d ** 2).mean()) / max(1e-12, np.sqrt((r[..., :3] ** 2).mean())))
            res[f"f{f}_maxabs"] = float(np.abs(d).max())
            res[f"f{f}_exact_frac"] = float((g == r).all(-1).mean())
            dep_g = rnd.read_image(A.TEXTURE_DEPTH); mot_g = rnd.read_image(A.TEXTURE_MOTION).astype(np.float32)
            res[f"f{f}_depth_eq"] = float((dep_g == imgs.arrays[A.TEXTURE_DEPTH]).mean())
            res[f"f{f}_motion_maxabs"] = float(np.abs(mot_g - imgs.arrays[A.TEXTURE_MOTION].astype(np.float32)).max())
    if any(sc.desc().meshes[i].jointCount for i in range(sc.desc().meshCount)) and orc:
        for m in range(sc.desc().meshCount):
            if sc.desc().meshes[m].jointCount:
                n = sc.desc().meshes[m].vertexCount
                gp, gn, _ = rnd.mesh_streams(m, n)
                op, on, _ = orc.mesh_streams(m, n)
                res["skin_pos_exact"] = bool((gp == op).all()); res["skin_nrm_exact"] = bool((gn == on).all())
                res["skin_pos_maxrel"] = float(np.abs(gp - op).max() / max(1e-12, np.abs(op).max()))
    info = ctx.as_info(rnd.tlas_id())
    res["tlas"] = {"prims": info.primitiveCount, "nodes": info.wideNodeCount, "levels": info.levelCount}
    for m in range(min(3, sc.desc().meshCount)):
        bi = ctx.as_info(rnd.blas_id(m))
        res[f"blas{m}"] = {"prims": bi.primitiveCount, "nodes": bi.wideNodeCount, "levels": bi.levelCount,
                           "sah": round(bi.sahCost, 2), "MB": round(bi.bytes / 1e6, 2)}
    res["launches"] = ctx.launches
    rnd.close(); ctx.close()
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    which = sys.argv[1:] or ["K1", "K3small", "K4small", "K5small", "K2", "K3"]
    for n in which:
        if n == "K1": compare("K1", 512, 512)
        elif n == "K3small": compare("K3small", 256, 256, spp=2, mb=3, frames=2)
        elif n == "K3smallglass": compare("K3glass", 64, 64, spp=1, mb=2)
        elif n == "K4small": compare("K4small", 256, 256, spp=1, mb=2)
        elif n == "K5small": compare("K5small", 256, 256, spp=1, mb=2, frames=3, animate=True)
        elif n == "K2": compare("K2", 640, 360, spp=1, mb=2)
        elif n == "K2tex": compare("K2tex", 320, 180, spp=1, mb=2)
        elif n == "K3": compare("K3", 1920, 1080, spp=1, mb=2, oracle_too=False)
        elif n == "K3o": compare("K3", 480, 270, spp=1, mb=2)
