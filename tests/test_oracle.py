"""Known-answer tests that pin the CPU oracle (the reference ships no tests, so these are authored from the
reference's source: Raytracing.metal:28-57 halton, :61-74 barycentric convention, :150-166 BRDF terms, :421 sampler,
Skinning.metal:7-49) plus the committed golden frame. CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from brute_force import brute_force_hits, uv_sphere
from metal4_raytracing_b200 import _abi as A
from metal4_raytracing_b200 import scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def radical_inverse(i, b):
    f, r = np.float32(1.0), np.float32(0.0)
    inv = np.float32(1.0) / np.float32(b)
    while i > 0:
        f = np.float32(f * inv)
        r = np.float32(r + np.float32(f * np.float32(i % b)))
        i //= b
    return float(r)


def test_halton_known_values():
    # base 2: 0, 1/2, 1/4, 3/4, 7/8
    assert [oracle.halton(i, 0) for i in (0, 1, 2, 3, 7)] == [0.0, 0.5, 0.25, 0.75, 0.875]
    assert oracle.halton(1, 1) == pytest.approx(1 / 3, abs=1e-7)
    assert oracle.halton(5, 2) == pytest.approx(1 / 5 * 0 + 1 / 25, abs=1e-7)  # 5 = (1,0) base 5 -> 0/5 + 1/25


def test_halton_matches_float32_restatement():
    primes = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101]
    rng = np.random.default_rng(0)
    for i in list(rng.integers(0, 2 ** 21, 40)) + [0, 1, 1048575, 2 ** 24 + 5]:
        for d in (0, 1, 2, 7, 13, 25):
            assert oracle.halton(int(i), d) == radical_inverse(int(i), primes[d]), (i, d)
    # the reference would read past primes[100] for d > 99 (F9); the oracle wraps the index
    assert oracle.halton(12345, 103) == oracle.halton(12345, 3)
    vals = [oracle.halton(i, 3) for i in range(1, 2000)]
    assert 0.0 < min(vals) and max(vals) < 1.0


def test_halton_shortcuts_used_by_the_cuda_path_are_identities():
    """The CUDA kernels evaluate halton() with two shortcuts that must be exact identities of the reference loop
    (Raytracing.metal:42-57): base 2 as a bit reversal while i < 2^24, and i / p by a multiply-high with the
    round-up reciprocal of tools/gen_halton_table.py."""
    rng = np.random.default_rng(3)
    idx = np.concatenate([np.arange(1, 3000), rng.integers(1, 1 << 24, 4000), [(1 << 24) - 1, 1 << 23, 0xAAAAAA]])
    for i in idx:
        i = int(i)
        brev = int(f"{i:032b}"[::-1], 2)
        assert oracle.halton(i, 0) == float(np.float32(brev) * np.float32(2.0 ** -32)), i
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "gen_halton_table.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    n = np.concatenate([rng.integers(1, 1 << 31, 50000, dtype=np.uint64), np.array([1, (1 << 31) - 1], np.uint64)])
    for p in gen.PRIMES:
        _, magic, shift, inv = gen.entry(p)
        assert np.array_equal(((n * np.uint64(magic)) >> np.uint64(32)) >> np.uint64(shift), n // np.uint64(p)), p
        assert inv == int((np.float32(1) / np.float32(p)).view(np.uint32))


def test_triangle_barycentric_convention_and_bounds():
    v0, v1, v2 = (0, 0, 0), (1, 0, 0), (0, 1, 0)
    hit, (t, u, v) = oracle.intersect_triangle((0.25, 0.5, 1.0), (0, 0, -1), v0, v1, v2)
    # Metal: u weights vertex 1, v weights vertex 2 (Raytracing.metal:63-73)
    assert hit and t == 1.0 and u == 0.25 and v == 0.5
    # two-sided
    hit, (t, u, v) = oracle.intersect_triangle((0.25, 0.5, -2.0), (0, 0, 1), v0, v1, v2)
    assert hit and t == 2.0 and u == 0.25 and v == 0.5
    # direction is not renormalised: t scales inversely with |d|
    hit, (t, _, _) = oracle.intersect_triangle((0.25, 0.5, 1.0), (0, 0, -4), v0, v1, v2)
    assert hit and t == 0.25
    # tmin < t < tmax, both exclusive
    assert not oracle.intersect_triangle((0.25, 0.5, 1.0), (0, 0, -1), v0, v1, v2, 0.0, 1.0)[0]
    assert not oracle.intersect_triangle((0.25, 0.5, 1.0), (0, 0, -1), v0, v1, v2, 1.0, 5.0)[0]
    assert oracle.intersect_triangle((0.25, 0.5, 1.0), (0, 0, -1), v0, v1, v2, 0.5, 1.5)[0]
    # behind the origin, outside, parallel, degenerate
    assert not oracle.intersect_triangle((0.25, 0.5, 1.0), (0, 0, 1), v0, v1, v2)[0]
    assert not oracle.intersect_triangle((0.8, 0.8, 1.0), (0, 0, -1), v0, v1, v2)[0]
    assert not oracle.intersect_triangle((0.25, 0.5, 1.0), (1, 0, 0), v0, v1, v2)[0]
    assert not oracle.intersect_triangle((0.25, 0.5, 1.0), (0, 0, -1), v0, v0, v0)[0]


def test_triangle_edges_are_watertight():
    """A ray through a shared edge or vertex hits at least one of the adjacent triangles (Woop et al. 2013)."""
    a, b, c, d = (0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)
    rng = np.random.default_rng(1)
    for _ in range(200):
        s = float(rng.random())
        p = (s, s, 0.0)  # on the shared diagonal a-c
        o = (float(rng.normal()), float(rng.normal()), 3.0 + float(rng.random()))
        dirv = tuple(np.float32(p[k]) - np.float32(o[k]) for k in range(3))
        h1 = oracle.intersect_triangle(o, dirv, a, b, c)[0]
        h2 = oracle.intersect_triangle(o, dirv, a, c, d)[0]
        assert h1 or h2


def test_half_conversion_matches_ieee():
    rng = np.random.default_rng(2)
    vals = np.concatenate([rng.normal(size=2000).astype(np.float32) * np.float32(10.0) ** rng.integers(-8, 5, 2000),
                           np.array([0.0, -0.0, 1.0, 65504.0, 65520.0, 1e-8, 6e-8, 5.96e-8, np.inf, -np.inf, 1e9],
                                    np.float32)])
    L = oracle.lib()
    for f in vals:
        h = L.oracle_float_to_half(float(f))
        assert h == int(np.float32(f).astype(np.float16).view(np.uint16)), f
    for h in list(range(0, 65536, 97)) + [0x7C00, 0xFC00, 0x0001, 0x03FF, 0x0400]:
        got = L.oracle_half_to_float(h)
        ref = float(np.array([h], np.uint16).view(np.float16)[0])
        assert (np.isnan(got) and np.isnan(ref)) or got == ref, hex(h)


def test_invert_affine_is_an_inverse():
    rng = np.random.default_rng(3)
    for _ in range(20):
        m = rng.normal(size=(4, 3)).astype(np.float32)  # [column][row]
        m[:3] += np.eye(3, dtype=np.float32) * 2
        inv = oracle.invert_affine(m).reshape(4, 3)
        M = np.eye(4)
        M[:3, :3], M[:3, 3] = m[:3].T, m[3]
        I = np.eye(4)
        I[:3, :3], I[:3, 3] = inv[:3].T, inv[3]
        assert np.allclose(M @ I, np.eye(4), atol=2e-5)


def test_texture_sampling_bilinear_repeat_srgb():
    tex = np.zeros((2, 2, 4), np.uint8)
    tex[0, 0], tex[0, 1], tex[1, 0], tex[1, 1] = (0, 0, 0, 255), (255, 0, 0, 255), (0, 255, 0, 255), (255, 255, 255, 255)
    # texel centres
    assert oracle.sample_texture(tex, 0.25, 0.25) == (0.0, 0.0, 0.0, 1.0)
    assert oracle.sample_texture(tex, 0.75, 0.25)[:3] == (1.0, 0.0, 0.0)
    # halfway between the two texels of row 0
    r = oracle.sample_texture(tex, 0.5, 0.25)
    assert r[0] == pytest.approx(0.5) and r[1] == 0.0
    # repeat addressing: u = 0 sits between the last and first texel of the row
    r0, r1 = oracle.sample_texture(tex, 0.0, 0.25), oracle.sample_texture(tex, 1.0, 0.25)
    assert r0 == r1 and r0[0] == pytest.approx(0.5)
    assert oracle.sample_texture(tex, 2.75, -0.75) == oracle.sample_texture(tex, 0.75, 0.25)
    # sRGB decode applies to rgb only
    grey = np.full((1, 1, 4), 128, np.uint8)
    lin = oracle.sample_texture(grey, 0.5, 0.5, srgb=True)
    assert lin[0] == pytest.approx(((128 / 255 + 0.055) / 1.055) ** 2.4, rel=1e-6)
    assert lin[3] == pytest.approx(128 / 255)


def test_skinning_identity_and_blend():
    rng = np.random.default_rng(4)
    n = 257
    pos = np.zeros((n, 4), np.float32)
    pos[:, :3] = rng.normal(size=(n, 3))
    nrm = np.zeros((n, 4), np.float32)
    nrm[:, :3] = rng.normal(size=(n, 3))
    idx = rng.integers(0, 3, (n, 4)).astype(np.uint16)
    w = rng.random((n, 4)).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    ident = np.tile(np.eye(4, dtype=np.float32).reshape(16), (3, 1))
    p, q = oracle.skin(pos, nrm, idx, w, ident)
    assert np.allclose(p[:, :3], pos[:, :3], atol=1e-6) and np.allclose(q[:, :3], nrm[:, :3], atol=1e-6)
    # two-joint blend == lerp of the two transforms (translation by +x and +y)
    mats = ident.copy()
    mats[1, 12], mats[2, 13] = 2.0, 4.0  # column-major: translation in elements 12..14
    idx2 = np.tile(np.array([1, 2, 0, 0], np.uint16), (n, 1))
    w2 = np.tile(np.array([0.25, 0.75, 0, 0], np.float32), (n, 1))
    p, q = oracle.skin(pos, nrm, idx2, w2, mats)
    assert np.allclose(p[:, 0], pos[:, 0] + 0.5, atol=1e-6) and np.allclose(p[:, 1], pos[:, 1] + 3.0, atol=1e-6)
    assert np.allclose(q[:, :3], nrm[:, :3], atol=1e-6)  # normals ignore translation
    # all-zero weights fall back to joint indices.x with weight 1, weights are NOT renormalised otherwise
    w0 = np.zeros((n, 4), np.float32)
    p, _ = oracle.skin(pos, nrm, idx2, w0, mats)
    assert np.allclose(p[:, 0], pos[:, 0] + 2.0, atol=1e-6)
    wh = np.tile(np.array([0.5, 0, 0, 0], np.float32), (n, 1))
    p, _ = oracle.skin(pos, nrm, np.zeros((n, 4), np.uint16), wh, ident)
    assert np.allclose(p[:, :3], 0.5 * pos[:, :3], atol=1e-6)


def _single_triangle_scene(w=32, h=32):
    sc = scene.Scene()
    m = sc.add_raw([(-50, 0, 50), (50, 0, 50), (0, 0, -50)], [(0, 1, 2)], normals=[(0, 1, 0)] * 3)
    sc.add_instance(m)
    sc.add_light(scene.make_light(A.LIGHT_POINT, position=(0, 2, 0), color=(3, 3, 3)))
    u = scene.default_uniforms(w, h)
    u.lightCount, u.samplesPerPixel, u.maxBounces = 1, 1, 1
    u.enableMotionAdaptiveSampling = u.enableMotionAdaptiveAccumulation = 0
    return sc, u


def test_ray_hits_ground_at_known_point():
    sc, u = _single_triangle_scene()
    orc = oracle.Oracle(sc, threads=1)
    hit, ids, (t, uu, vv) = orc.trace_ray((1.0, 3.0, 2.0), (0.0, -1.0, 0.0))
    assert hit and ids == (0, 0, 0) and t == 3.0
    assert not orc.trace_ray((1.0, 3.0, 2.0), (0.0, 1.0, 0.0))[0]
    # instance transform: move the ground up by 1 -> t = 2
    sc.set_instance_transform(0, (0, 1, 0), (0, 0, 0), 1.0)
    orc.update()
    hit, ids, (t, _, _) = orc.trace_ray((1.0, 3.0, 2.0), (0.0, -1.0, 0.0))
    assert hit and t == pytest.approx(2.0, abs=1e-6)


def test_oracle_traversal_against_brute_force():
    """The oracle's two-level BVH traversal (oracle_bvh.cpp: SAH BVH2 per mesh, instance transforms, watertight test, tie
    rule) against a float64 brute force over the same triangles moved to world space by this test: ids, t and (u, v) agree
    wherever float64 puts the hit clear of the triangle's edges. The meshes are the test's own arrays (add_raw), so the OBJ
    loader and the procedural meshes play no part; the GPU traversal is checked against the same brute force in
    test_gpu_parity.py::test_intersect_against_brute_force."""
    rng = np.random.default_rng(29)
    sph_v, sph_t = uv_sphere(16, 20, 1.0)
    g = np.linspace(-1.5, 1.5, 9, dtype=np.float32)
    gx, gz = np.meshgrid(g, g, indexing="ij")
    grid_v = np.stack([gx.ravel(), np.full(81, -1.25, np.float32), gz.ravel()], 1)
    grid_t = np.array([t for i in range(8) for j in range(8)
                       for t in ((i * 9 + j, i * 9 + j + 1, (i + 1) * 9 + j), (i * 9 + j + 1, (i + 1) * 9 + j + 1, (i + 1) * 9 + j))], np.int32)
    sc = scene.Scene()
    meshes = [sc.add_raw(sph_v[:, :3], sph_t), sc.add_raw(grid_v, grid_t)]
    arrays = [(sph_v[:, :3], sph_t), (grid_v, grid_t)]
    count = 24
    for k in range(count):
        sc.add_instance(meshes[k % 2], tuple(rng.uniform(-5, 5, 3)), tuple(rng.uniform(0, 6.28, 3)), float(rng.uniform(0.5, 1.8)))
    sc.add_light(scene.make_light(A.LIGHT_POINT, position=(0, 9, 0)))
    d = sc.desc()
    tris_world, owner = [], []
    for k in range(count):
        inst = d.instances[k]
        M = np.array(inst.transform[:], np.float64).reshape(4, 4).T  # stored column-major
        v, t = arrays[inst.meshIndex]
        w = v.astype(np.float64) @ M[:3, :3].T + M[:3, 3]
        tris_world.append(w[t])
        owner += [(k, 0, p) for p in range(len(t))]
    tris_world, owner = np.concatenate(tris_world), np.array(owner, np.int64)
    n = 1500
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-8, 8, (n, 3))
    rays[:, 4:7] = rng.uniform(-5, 5, (n, 3)) - rays[:, 0:3]   # not normalised
    rays[: n // 10, 4:7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 10)]
    rays[:, 3] = np.where(rng.random(n) < 0.3, 0.2, 0.0)
    rays[:, 7] = np.where(rng.random(n) < 0.3, 1.2, np.inf)
    best, second, arg, margin = brute_force_hits(tris_world, owner, rays)
    orc = oracle.Oracle(sc, threads=1)
    checked = 0
    for i in range(n):
        hit, ids, (t, u, v) = orc.trace_ray(tuple(rays[i, 0:3]), tuple(rays[i, 4:7]), float(rays[i, 3]), float(rays[i, 7]))
        if np.isfinite(best[i]) and margin[i] > 1e-3:
            assert hit and abs(t - best[i]) <= 2e-5 * max(1.0, best[i])
            if second[i] - best[i] > 1e-4 * max(1.0, best[i]):
                assert ids == tuple(owner[arg[i]])
                w = tris_world[arg[i]]
                point = (1 - u - v) * w[0] + u * w[1] + v * w[2]
                on_ray = rays[i, 0:3].astype(np.float64) + t * rays[i, 4:7].astype(np.float64)
                assert np.abs(point - on_ray).max() < 2e-3
            checked += 1
        elif not np.isfinite(best[i]) and margin[i] > 0:
            assert not hit
    assert checked > n // 8


def test_lambert_point_light_radiance_value():
    """Centre pixel of a plane lit by one point light: closed form of the kernel's PBR branch with roughness 1,
    metallic 0 (Raytracing.metal:692-744): direct = (kD*albedo/pi + spec) * L * NdotL."""
    sc, u = _single_triangle_scene(33, 33)
    u.camera = scene.orbit_camera(33, 33, (0, 0, 0), 0.0, 1.2, 4.0)
    u.previousCamera = u.camera
    imgs = oracle.FrameImages(33, 33, np.zeros((33, 33), np.uint32), fp32=True)
    orc = oracle.Oracle(sc, threads=2)
    stats, ids = orc.render(u, imgs, want_ids=True)
    assert stats["closest"] == 33 * 33 and stats["any"] == stats["hits"]
    img = imgs.output
    c = img[16, 16, :3]
    assert np.all(c > 0) and c[0] == c[1] == c[2] and img[16, 16, 3] == 1.0
    # brightest near the point below the light, darker toward the edges
    assert img[16, 16, 0] > img[2, 2, 0]
    depth = imgs.arrays[A.TEXTURE_DEPTH][16, 16, 0]
    assert depth == pytest.approx(4.0, rel=2e-2)


def test_sphere_silhouette_area(assets):
    """sphere.obj (radius 1) seen from distance 5.38 with a 45 degree vertical fov at 256x256: the silhouette's
    pixel count matches the analytic projected disc within 2 %."""
    w = h = 256
    sc = scene.Scene()
    m = sc.add_obj(os.path.join(assets, "sphere.obj"))
    sc.add_instance(m)
    sc.add_light(scene.make_light(A.LIGHT_POINT, position=(0, 5, 5), color=(9, 9, 9)))
    u = scene.default_uniforms(w, h)
    u.lightCount, u.samplesPerPixel, u.maxBounces = 1, 1, 1
    u.enableMotionAdaptiveSampling = u.enableMotionAdaptiveAccumulation = 0
    u.camera = scene.orbit_camera(w, h, (0, 0, 0), 0.0, 0.0, 5.38)
    u.previousCamera = u.camera
    imgs = oracle.FrameImages(w, h, np.zeros((h, w), np.uint32))
    _, ids = oracle.Oracle(sc).render(u, imgs, want_ids=True)
    covered = int((ids[..., 0] != 0xFFFFFFFF).sum())
    d, r = 5.38, 1.0
    tan_half = np.tan(np.radians(22.5))
    # silhouette of a sphere under perspective: circle of angular radius asin(r/d) -> tan = r / sqrt(d^2 - r^2)
    rad_px = (r / np.sqrt(d * d - r * r)) / tan_half * (h / 2)
    assert covered == pytest.approx(np.pi * rad_px ** 2, rel=0.02)


def test_frame_index_and_ema(assets):
    """frameIndex 0 writes the plain mean; frameIndex > 0 blends with history at <= 0.95 (Raytracing.metal:796-817)."""
    sc, u, seed = scene.Scene.named("K1", 48, 48, assets=assets)
    seeds = scene.seed_image(48, 48, seed)
    orc = oracle.Oracle(sc)
    imgs = oracle.FrameImages(48, 48, seeds, fp32=True)
    u.accumulationWeight = 0.99  # clamped to 0.95
    u.frameIndex = 0
    orc.render(u, imgs)
    f0 = imgs.output.copy()
    imgs.swap()
    u.frameIndex = 1
    orc.render(u, imgs)
    f1 = imgs.output.copy()
    # recompute frame 1's own samples without history, then blend by hand
    imgs2 = oracle.FrameImages(48, 48, seeds, fp32=True)
    u2 = u.copy()
    u2.accumulationWeight = 0.0
    orc.render(u2, imgs2)
    cur = imgs2.output
    expect = cur[..., :3] + (f0[..., :3] - cur[..., :3]) * np.float32(0.95)
    assert np.allclose(f1[..., :3], expect, rtol=1e-6, atol=1e-7)


def test_max_bounces_counts_segments(assets):
    """maxBounces = 1 -> primary + shadow only (F13): closest rays == pixels, any-hit rays <= hits."""
    sc, u, seed = scene.Scene.named("K1", 64, 64, assets=assets)
    imgs = oracle.FrameImages(64, 64, scene.seed_image(64, 64, seed))
    orc = oracle.Oracle(sc)
    st1, _ = orc.render(u, imgs)
    assert st1["closest"] == 64 * 64 and st1["any"] <= st1["hits"]
    u.maxBounces = 2
    st2, _ = orc.render(u, imgs)
    assert st2["closest"] == 64 * 64 + st1["hits"]  # every primary hit spawns exactly one bounce ray (albedo > 0)


def test_golden_frame(assets):
    """The committed K1 frame (tests/golden/make_golden.py) pins the oracle against drift."""
    path = os.path.join(GOLDEN, "k1_128.npz")
    g = np.load(path)
    w = h = 128
    sc, u, seed = scene.Scene.named("K1", w, h, assets=assets)
    C.memmove(C.byref(u), g["uniforms"].tobytes(), C.sizeof(A.Uniforms))  # camera bytes from the fixture
    imgs = oracle.FrameImages(w, h, scene.seed_image(w, h, int(g["seed"])))
    stats, ids = oracle.Oracle(sc).render(u, imgs, want_ids=True)
    assert np.array_equal(ids, g["ids"])
    assert np.array_equal(imgs.output.view(np.uint16), g["image"].view(np.uint16))
    assert np.array_equal(imgs.arrays[A.TEXTURE_DEPTH], g["depth"])
    assert [stats["closest"], stats["any"], stats["hits"]] == list(g["stats"])


def test_environment_extension_known_answers(assets):
    """rt_environment (extension): a constant environment makes every escaping primary ray return intensity x value;
    the equirect mapping puts 'straight up' in row 0 and wraps columns; unbound = black (the reference behaviour)."""
    from metal4_raytracing_b200 import scene
    w = h = 32
    sc, u, seed = scene.Scene.named("K1", w, h, assets=assets)
    u.samplesPerPixel, u.maxBounces = 1, 1
    seeds = scene.seed_image(w, h, seed)
    orc = oracle.Oracle(sc)
    imgs = oracle.FrameImages(w, h, seeds, fp32=True)
    _, ids = orc.render(u, imgs, want_ids=True)
    miss = ids[..., 0] == 0xFFFFFFFF
    assert miss.any() and not miss.all()
    assert float(np.abs(imgs.output[miss][:, :3]).max()) == 0.0
    const = np.full((8, 16, 4), 0.5, np.float32)
    orc.set_environment(const, 3.0)
    orc.render(u, imgs)
    assert np.allclose(imgs.output[miss][:, :3], 1.5, rtol=0, atol=1e-6)
    # mapping: u = atan2(z, x) / 2pi + 0.5, v = acos(y) / pi; texel centres at (i + 0.5) / size; columns wrap
    grid = np.zeros((4, 8, 4), np.float32)
    grid[..., 0] = np.arange(8, dtype=np.float32)[None, :]  # red = column
    grid[..., 1] = np.arange(4, dtype=np.float32)[:, None]  # green = row
    probe = lambda d: oracle.sample_environment(grid, d)
    assert probe((0, 1, 0))[1] == 0.0 and probe((0, -1, 0))[1] == 3.0          # up = row 0, down = last row
    assert abs(probe((1, 0, 0))[0] - 3.5) < 1e-5 and abs(probe((1, 0, 0))[1] - 1.5) < 1e-5   # +x: u = v = 0.5
    assert abs(probe((0, 0, 1))[0] - 5.5) < 1e-5                                # +z: u = 0.75
    assert abs(probe((0, 0, -1))[0] - 1.5) < 1e-5                               # -z: u = 0.25
    assert abs(probe((-1, 0, 0))[0] - 3.5) < 1e-5                               # -x: u = 1 -> blend of columns 7 and 0
    assert np.allclose(oracle.sample_environment(grid, (0, 0, 1), 2.0), 2.0 * probe((0, 0, 1)))
    orc.set_environment(None)
    orc.render(u, imgs)
    assert float(np.abs(imgs.output[miss][:, :3]).max()) == 0.0


def test_environment_importance_tables():
    """RT_ENV_IMPORTANCE (include/rt_b200.h): the library's table builder and the oracle's restatement agree bit for
    bit; a constant map gives the sin(theta) marginal and uniform rows; a single bright texel takes all the mass."""
    from metal4_raytracing_b200 import device, scene
    sky = scene.procedural_sky(64, 32)
    a, b = device.environment_cdf(sky), oracle.environment_cdf(sky)
    n = (32 + 1) + 32 * (64 + 1)
    assert b.shape == (n,) and a.shape == (n + 65 * 33,) and np.array_equal(a[:n], b)
    # RT_ENV_GUIDED: behind the running sums, 65 guide entries for the marginal and for every row — guide[k] = the largest
    # cell index whose running sum is <= k / 64, i.e. where the plain search for any xi of that bucket would still be right
    guides = a[n:].view(np.uint32).reshape(33, 65)
    sums = [a[:33]] + [a[33 + y * 65: 33 + (y + 1) * 65] for y in range(32)]
    for g, c in zip(guides, sums):
        cells = len(c) - 1
        for k in (0, 1, 17, 40, 63, 64):
            assert g[k] == max(i for i in range(cells + 1) if c[i] <= np.float32(k) / np.float32(64)), (k, g[k])
        for xi in np.float32([0.0, 0.013, 0.31, 0.5, 0.77, 0.999]):
            k = min(int(xi * 64), 63)
            want = max(i for i in range(cells) if c[i] <= xi)  # what the search from [0, n) returns
            assert g[k] <= want < min(g[k + 1] + 1, cells) or want == g[k]
    h, w = 16, 8
    t = oracle.environment_cdf(np.full((h, w, 4), 0.25, np.float32))
    marginal, rows = t[:h + 1], t[h + 1:].reshape(h, w + 1)
    assert marginal[0] == 0.0 and marginal[-1] == 1.0 and np.all(np.diff(marginal) > 0)
    edges = (1.0 - np.cos(np.pi * np.arange(h + 1) / h)) / 2.0  # integral of sin over [0, theta] / 2
    assert np.abs(marginal - edges).max() < 2e-3
    assert np.allclose(rows, np.arange(w + 1, dtype=np.float32)[None, :] / w, atol=1e-7)
    one = np.zeros((h, w, 4), np.float32)
    one[5, 3, :3] = 7.0
    t = oracle.environment_cdf(one)
    marginal, rows = t[:h + 1], t[h + 1:].reshape(h, w + 1)
    assert np.array_equal(marginal, (np.arange(h + 1) > 5).astype(np.float32))
    assert np.array_equal(rows[5], (np.arange(w + 1) > 3).astype(np.float32))
    assert np.allclose(rows[0], np.arange(w + 1, dtype=np.float32) / w)  # a row without weight is uniform
    black = oracle.environment_cdf(np.zeros((h, w, 4), np.float32))    # so is a map without weight
    assert np.allclose(black[:h + 1], np.arange(h + 1, dtype=np.float32) / h)


def test_environment_importance_sampling_converges_to_the_same_image(assets):
    """Sampling the environment as a light (balance heuristic against the cosine bounce) must not change what the
    estimator converges to, only its noise: with a small bright sun the two means agree and the light-sampled frame
    is the less noisy one."""
    from metal4_raytracing_b200 import scene
    w = h = 40
    sc, u, seed = scene.Scene.named("K1", w, h, assets=assets)
    u.samplesPerPixel, u.maxBounces = 256, 2
    seeds = scene.seed_image(w, h, seed)
    sky = scene.procedural_sky(128, 64)
    orc = oracle.Oracle(sc)
    imgs = oracle.FrameImages(w, h, seeds, fp32=True)
    _, ids = orc.render(u, imgs, want_ids=True)
    hit = ids[..., 0] != 0xFFFFFFFF
    out = {}
    for importance in (False, True):
        orc.set_environment(sky, 1.0, importance=importance)
        orc.render(u, imgs)
        out[importance] = imgs.output[..., :3].astype(np.float64).copy()
    plain, sampled = out[False][hit], out[True][hit]
    assert not np.array_equal(plain, sampled)
    assert abs(sampled.mean() - plain.mean()) / plain.mean() < 0.02
    # noise: difference to the 256-spp mean of a 16-spp frame, light-sampled against not
    u.samplesPerPixel = 16
    err = {}
    for importance in (False, True):
        orc.set_environment(sky, 1.0, importance=importance)
        orc.render(u, imgs)
        ref = out[importance][hit]
        err[importance] = float(np.sqrt(np.mean((imgs.output[..., :3].astype(np.float64)[hit] - ref) ** 2)))
    assert err[True] < err[False]
    orc.set_environment(None)


def _golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["k3small_96x64", "k5small_64", "k2tex_80x48"])
def test_more_golden_frames(name):
    """Committed fixtures of three procedural scenes (bounces + EMA; skinning + refit + adaptive sampling over three
    frames; textured PBR + G-buffer) pin the oracle against drift: rendering them again gives the same bits."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    got = _golden_module().render_case(name)
    assert np.array_equal(got["uniforms"], g["uniforms"])
    for key in ("ids", "image", "depth", "motion", "stats") + (("normal",) if "normal" in g else ()):
        a, b = got[key], g[key]
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), key


def test_enable_ao_switch_in_the_oracle():
    """ENABLE_AO (ShaderTypes.h:155-157; Raytracing.metal:405-409,442-446,475-479,672,748): off — the reference's shipping
    build — a bound ambient-occlusion map changes nothing; on, it scales the bounce throughput (darker or equal
    everywhere) and debug view 5 shows the sampled value instead of magenta."""
    from metal4_raytracing_b200 import _abi as A
    w, h = 96, 64

    def render(bind, enable, debug=0):
        sc, u, seed = scene.Scene.named("K2tex", w, h, assets=None)
        if bind:
            sc.bind_texture(0, 0, A.SLOT_AO, sc.add_texture_procedural("valuenoise", 64, 64, seed=5, srgb=False))
        u.samplesPerPixel, u.maxBounces, u.debugTextureMode = 2, 3, debug
        orc = oracle.Oracle(sc)
        orc.set_enable_ao(enable)
        imgs = oracle.FrameImages(w, h, scene.seed_image(w, h, seed))
        orc.render(u, imgs)
        return imgs.output.astype(np.float32)[..., :3]

    plain = render(False, False)
    assert np.array_equal(render(True, False), plain)
    assert np.array_equal(render(False, True), plain)  # the switch alone, without a map, changes nothing
    on = render(True, True)
    assert not np.array_equal(on, plain) and on.sum() < plain.sum()
    dbg_off, dbg_on = render(True, False, A.DEBUG_AO), render(True, True, A.DEBUG_AO)
    hit = dbg_off.sum(-1) > 0
    assert hit.any() and np.all(dbg_off[hit][:, 1] == 0.0) and np.all(dbg_off[hit][:, 0] == dbg_off[hit][:, 2])  # magenta
    textured = (dbg_on[..., 0] == dbg_on[..., 1]) & (dbg_on[..., 1] == dbg_on[..., 2]) & hit
    assert textured.any()
