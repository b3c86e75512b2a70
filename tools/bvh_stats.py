"""Build time and SAH cost of the K3 dragon BLAS for several PLOC radii (0 = LBVH)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metal4_raytracing_b200 import device, scene
name = sys.argv[1] if len(sys.argv) > 1 else "K3"
for radius in (0, 4, 8, 16, 32, 64):
    sc, u, seed = scene.Scene.named(name, 1920, 1080)
    ctx = device.Context(0)
    ctx.set_option("ploc_radius", radius)
    t0 = time.perf_counter()
    rnd = device.Renderer(ctx, sc, 1920, 1080, seeds=scene.seed_image(1920, 1080, seed))
    ctx.sync()
    dt = time.perf_counter() - t0
    n_mesh = device.lib().rtr_mesh_count(rnd._h)
    infos = [ctx.as_info(rnd.blas_id(m)) for m in range(n_mesh)]
    big = max(infos, key=lambda i: i.primitiveCount)
    tl = ctx.as_info(rnd.tlas_id())
    print(json.dumps({"radius": radius, "create_s": round(dt, 3), "tris": big.primitiveCount, "nodes": big.wideNodeCount,
                      "levels": big.levelCount, "sah": round(big.sahCost, 2), "tlas_sah": round(tl.sahCost, 2),
                      "tlas_nodes": tl.wideNodeCount}), flush=True)
    rnd.close(); ctx.close()
