"""Host scene library: OBJ loader rules, transforms (Utilities.swift:302-355), orbit camera (Scene.swift:126-159),
default lights/uniforms (Scene.swift:82-93, Renderer.swift:117-192), procedural stand-ins, animation
(Model.swift:207-261). CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from metal4_raytracing_b200 import _abi as A
from metal4_raytracing_b200 import scene


def test_obj_loader_rules(assets, tmp_path):
    sc = scene.Scene()
    plane = sc.add_obj(os.path.join(assets, "plane.obj"))
    a = sc.mesh_arrays(plane)
    assert a["positions"].shape == (4, 4) and len(a["submeshes"]) == 1
    assert a["submeshes"][0].tolist() == [[0, 1, 2], [0, 2, 3]]  # fan triangulation of the quad
    assert a["uvs"] is not None and np.allclose(a["normals"][:, :3], [0, 1, 0])
    m = sc.get_material(plane)
    assert m.baseColor.tuple() == (0.5, 0.5, 0.5) and m.refractionIndex == 1.0 and m.opacity == 1.0
    assert m.specularExponent == 0.0  # Ns is never taken (SubMesh.swift:309-311)
    sphere = sc.add_obj(os.path.join(assets, "sphere.obj"))
    s = sc.mesh_arrays(sphere)
    assert sum(len(x) for x in s["submeshes"]) == 4900 and s["uvs"] is None
    train = sc.add_obj(os.path.join(assets, "train.obj"))
    t = sc.mesh_arrays(train)
    assert len(t["submeshes"]) == 6 and sum(len(x) for x in t["submeshes"]) == 3624  # one submesh per usemtl run
    teapot = sc.add_obj(os.path.join(assets, "teapot.obj"))
    tp = sc.mesh_arrays(teapot)
    assert sum(len(x) for x in tp["submeshes"]) == 15704 and np.all(tp["normals"] == 0)  # no vn -> zero normals
    # welding: one vertex per distinct v/vt/vn tuple, first appearance order; negative indices
    p = tmp_path / "w.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\nvn 0 0 -1\nf 1//1 2//1 3//1\nf -3//1 -1//2 -2//1\n")
    wm = sc.add_obj(str(p))
    wa = sc.mesh_arrays(wm)
    assert wa["positions"].shape[0] == 4 and wa["submeshes"][0].tolist() == [[0, 1, 2], [1, 3, 2]]
    assert wa["normals"][3, 2] == -1.0
    with pytest.raises(RuntimeError):
        sc.add_obj(str(tmp_path / "missing.obj"))


def test_glass_override(assets):
    sc = scene.Scene()
    m = sc.add_obj(os.path.join(assets, "sphere.obj"), glass=True)
    mat = sc.get_material(m)
    assert mat.baseColor.tuple() == pytest.approx((0.95, 0.98, 1.0)) and mat.refractionIndex == pytest.approx(1.52)
    assert mat.opacity == pytest.approx(0.08)


def test_transform_convention():
    """T * Rx*Ry*Rz * S, column-major (Mesh.swift:61-68)."""
    sc = scene.Scene()
    m = sc.add_procedural("plane")
    i = sc.add_instance(m, position=(1, 2, 3), rotation=(0, np.pi / 2, 0), scale=2.0)
    M = np.array(sc.desc().instances[i].transform[:], np.float32).reshape(4, 4).T  # rows
    p = M @ np.array([1, 0, 0, 1], np.float32)
    # rotateY(+90 deg) by the reference's matrix sends +x to -z
    assert np.allclose(p[:3], [1, 2, 3 - 2], atol=1e-6)
    assert np.allclose(M[:3, 3], [1, 2, 3])
    sc.set_instance_transform(i, (0, 0, 0), (0, 0, 0), 1.0)
    d = sc.desc().instances[i]
    assert np.allclose(np.array(d.transform[:]).reshape(4, 4), np.eye(4))
    assert np.allclose(np.array(d.previousTransform[:]).reshape(4, 4).T, M)  # previous transform kept


def test_orbit_camera_and_defaults():
    c = scene.default_camera(1920, 1080)
    assert c.position.tuple() == pytest.approx((0.0, 1.0, 5.38), abs=1e-5)
    f = np.array(c.forward.tuple())
    assert np.linalg.norm(f) == pytest.approx(1.0, abs=1e-6)
    assert np.allclose(f, -np.array([0, 1, 5.38]) / np.linalg.norm([0, 1, 5.38]), atol=1e-6)
    th = np.tan(np.radians(22.5))
    assert np.linalg.norm(c.up.tuple()) == pytest.approx(th, rel=1e-6)
    assert np.linalg.norm(c.right.tuple()) == pytest.approx(th * 1920 / 1080, rel=1e-6)
    assert abs(np.dot(c.right.tuple(), c.up.tuple())) < 1e-6
    u = scene.default_uniforms(640, 360)
    assert (u.samplesPerPixel, u.maxBounces, u.blocksWide) == (2, 2, 40)
    assert u.accumulationWeight == pytest.approx(0.9) and u.enableMotionAdaptiveSampling == 1
    assert u.motionSamplingMaxExtraSamples == 2 and u.motionAccumulationMinWeight == pytest.approx(0.1)
    sc = scene.Scene()
    sc.default_lights()
    d = sc.desc()
    assert d.lightCount == 2 and d.lights[0].type == A.LIGHT_AREA and d.lights[1].type == A.LIGHT_SPOT
    assert d.lights[1].coneAngle == pytest.approx(np.radians(25))


def test_seed_image_range_and_determinism():
    s = scene.seed_image(64, 48, 0xC0FFEE)
    assert s.shape == (48, 64) and s.max() < (1 << 20) and len(np.unique(s)) > 3000
    assert np.array_equal(s, scene.seed_image(64, 48, 0xC0FFEE))
    assert not np.array_equal(s, scene.seed_image(64, 48, 1))


def test_procedural_stand_in_sizes():
    sc = scene.Scene()
    b = sc.add_procedural("icosphere_bumpy", 6, 2)
    ba = sc.mesh_arrays(b)
    assert ba["positions"].shape[0] == 40962 and len(ba["submeshes"][0]) == 81920
    n = ba["normals"][:, :3]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    k = sc.add_procedural("torusknot", 1320, 330, 3)
    ka = sc.mesh_arrays(k)
    assert ka["positions"].shape[0] == 435600 and len(ka["submeshes"][0]) == 871200
    assert ka["positions"][:, 1].min() == pytest.approx(-0.3, abs=1e-6)
    idx = ka["submeshes"][0]
    assert idx.min() == 0 and idx.max() == 435599
    # closed surface: every edge is shared by exactly two triangles
    e = np.sort(np.concatenate([idx[:, [0, 1]], idx[:, [1, 2]], idx[:, [2, 0]]]), axis=1)
    _, counts = np.unique(e[:, 0].astype(np.int64) * 435600 + e[:, 1], return_counts=True)
    assert np.all(counts == 2)


def test_humanoid_skeleton_and_animation():
    sc = scene.Scene()
    m = sc.add_procedural("humanoid", 100000, 64)
    a = sc.mesh_arrays(m)
    assert a["positions"].shape[0] == 100000 and a["jointMatrices"].shape == (64, 16)
    assert 195000 <= len(a["submeshes"][0]) <= 200000
    assert a["jointIndices"].max() < 64
    assert np.allclose(a["jointWeights"].sum(1), 1.0, atol=1e-5)
    sc.animate(0.0)
    m0 = sc.mesh_arrays(m)["jointMatrices"].reshape(64, 4, 4)
    assert np.allclose(m0, np.eye(4), atol=1e-5)  # clip starts at the bind pose
    sc.animate(0.4)
    m1 = sc.mesh_arrays(m)["jointMatrices"].reshape(64, 4, 4).transpose(0, 2, 1)  # -> row-major
    rot = m1[:, :3, :3]
    assert np.allclose(rot @ rot.transpose(0, 2, 1), np.eye(3), atol=1e-4)  # rigid
    assert np.abs(m1 - np.eye(4)).max() > 0.05
    sc.animate(0.4 + 2.0)  # clip duration 2 s -> periodic
    m2 = sc.mesh_arrays(m)["jointMatrices"].reshape(64, 4, 4).transpose(0, 2, 1)
    assert np.allclose(m1, m2, atol=2e-3)


def test_keyed_skinned_mesh_and_file_round_trip(tmp_path, two_joint_arm):
    """rts_add_mesh_skinned + rts_set_animation_keys: the sampled clip gives the analytic palette
    (half way to a 90-degree bend is a 45-degree bend: the normalised-linear midpoint is the slerp midpoint), loops with
    the clip duration, hands the same TRS to the device-side palette evaluation, and survives the file format."""
    import oracle
    pos, tris, nrm, ji, jw, parents, rest, bind, times, keys = two_joint_arm
    sc = scene.Scene()
    m = sc.add_skinned(pos, tris, ji, jw, parents, rest, bind, normals=nrm)
    a = sc.mesh_arrays(m)
    assert np.allclose(a["jointMatrices"].reshape(2, 4, 4), np.eye(4), atol=1e-6)  # rest pose == bind pose
    sc.set_animation_keys(m, times, keys)
    sc.animate(0.5)
    a = sc.mesh_arrays(m)
    pal = a["jointMatrices"].reshape(2, 4, 4).transpose(0, 2, 1)  # -> row-major
    c = np.float32(np.cos(np.pi / 4))
    want = np.array([[c, -c, 0, c], [c, c, 0, 1 - c], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)  # T(0,1,0) Rz(45) T(0,-1,0)
    assert np.allclose(pal[0], np.eye(4), atol=1e-6) and np.allclose(pal[1], want, atol=1e-6)
    p4 = np.concatenate([pos, np.zeros((len(pos), 1), np.float32)], 1)
    n4 = np.concatenate([nrm, np.zeros((len(pos), 1), np.float32)], 1)
    skinned, _ = oracle.skin(p4, n4, ji, jw, a["jointMatrices"])
    tip = skinned[np.argmax(pos[:, 1] + 0.01 * pos[:, 0])]  # rest (0.1, 2, 0): elbow-local (0.1, 1, 0) turned by 45 degrees
    assert np.allclose(tip[:3], [0.1 * c - c, 1 + 0.1 * c + c, 0], atol=1e-5)
    assert np.allclose(skinned[pos[:, 1] <= 1.0][:, :3], pos[pos[:, 1] <= 1.0], atol=1e-7)
    # the TRS the device-side evaluation (rt_joint_palette) receives is the sampled key, and the clip loops
    trs = a["jointLocalTRS"]
    assert np.allclose(trs[1, :3], [0, 1, 0]) and np.allclose(trs[1, 7:], 1.0)
    q = trs[1, 3:7] / np.linalg.norm(trs[1, 3:7])
    assert np.allclose(q, [0, 0, np.sin(np.pi / 8), np.cos(np.pi / 8)], atol=1e-6)
    sc.animate(1.5)
    assert np.allclose(sc.mesh_arrays(m)["jointMatrices"], a["jointMatrices"], atol=1e-6)
    sc.animate(0.0)
    assert np.allclose(sc.mesh_arrays(m)["jointMatrices"].reshape(2, 4, 4), np.eye(4), atol=1e-6)
    # file round trip
    path = tmp_path / "arm.rtsk"
    sc.save_skinned(m, path)
    sc2 = scene.Scene()
    m2 = sc2.load_skinned(path)
    sc2.animate(0.5)
    b = sc2.mesh_arrays(m2)
    for key in ("positions", "normals", "jointIndices", "jointWeights", "jointMatrices", "jointLocalTRS", "jointParents"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(a["submeshes"][0], b["submeshes"][0])
    # errors are reported, not trapped
    with pytest.raises(RuntimeError):
        sc.add_skinned(pos, tris, ji, jw, [0, -1], rest, bind)  # a child before its parent
    with pytest.raises(RuntimeError):
        sc.set_animation_keys(m, [1.0, 0.5], keys)  # times must ascend
    (tmp_path / "bad.rtsk").write_bytes(b"RTSK1\0\0\0" + b"\1" * 20)
    with pytest.raises(RuntimeError):
        sc2.load_skinned(tmp_path / "bad.rtsk")


def test_keyed_skinned_mesh_renders_through_the_oracle(two_joint_arm_scene):
    """The raw skinned mesh goes through the same per-frame path as the stand-in (skin -> refit -> trace): the bent
    arm covers different pixels than the straight one."""
    import oracle
    sc, u, seeds, w, h = two_joint_arm_scene()
    orc = oracle.Oracle(sc)
    imgs = oracle.FrameImages(w, h, seeds)
    _, ids0 = orc.render(u, imgs, want_ids=True)
    sc.animate(0.75)  # (the clip lasts 1 s and loops: t = 1 would be the rest pose again)
    orc.update()
    _, ids1 = orc.render(u, imgs, want_ids=True)
    hit0, hit1 = ids0[..., 0] != 0xFFFFFFFF, ids1[..., 0] != 0xFFFFFFFF
    assert hit0.sum() > 20 and hit1.sum() > 20
    assert (hit0 != hit1).sum() > 10


def test_named_scenes(assets):
    for name, (w, h), tris, inst, spp, mb in [("K1", (512, 512), 4902, 2, 1, 1), ("K2", (1920, 1080), 81922, 2, 4, 2),
                                             ("K4small", (256, 256), 19682, 65, 8, 2),
                                             ("K5small", (256, 256), None, 2, 2, 2)]:
        sc, u, seed = scene.Scene.named(name, w, h, assets=assets)
        d = sc.desc()
        total = sum(d.meshes[i].submeshes[k].triangleCount for i in range(d.meshCount)
                    for k in range(d.meshes[i].submeshCount))
        if tris is not None:
            assert total == tris
        assert d.instanceCount == inst and (u.samplesPerPixel, u.maxBounces) == (spp, mb)
        assert u.lightCount == d.lightCount and u.width == w and u.height == h
    sc, u, _ = scene.Scene.named("K4small", 64, 64, assets=assets)
    assert sc.desc().maxSubmeshes == 6
    with pytest.raises(RuntimeError):
        scene.Scene.named("nope", 8, 8)
    with pytest.raises(RuntimeError):
        scene.Scene.named("K1", 8, 8, assets=None)  # K1 needs the OBJ files


def test_texture_binding_sets_flags():
    sc = scene.Scene()
    m = sc.add_procedural("uvsphere", 8, 8)
    t = sc.add_texture_procedural("checker", 16, 16, 1, srgb=True)
    sc.bind_texture(m, 0, A.SLOT_BASECOLOR, t)
    mat = sc.get_material(m)
    assert mat.textureFlags & 1 and mat.baseColor.tuple() == (1.0, 1.0, 1.0)  # SubMesh.swift:119-125
    d = sc.desc()
    assert d.meshes[m].submeshes[0].textureIndex[A.SLOT_BASECOLOR] == t
    assert d.textures[t].srgb == 1 and d.textures[t].width == 16
    # 1x1 fallbacks (SubMesh.swift:176-241): white, neutral normal, black
    assert bytes(d.textures[0].texels[0:4]) == b"\xff\xff\xff\xff" and bytes(d.textures[1].texels[0:4]) == b"\x80\x80\xff\xff"


def test_png_decoder(assets):
    from PIL import Image
    p = os.path.join(assets, "uv_test", "uv_test.png")
    if not os.path.isfile(p):
        pytest.skip("uv_test.png not staged")
    ref = np.asarray(Image.open(p).convert("RGBA"))
    sc = scene.Scene()
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "m.mtl"), "w") as f:
            f.write(f"newmtl a\nKd 1 1 1\nmap_Kd {os.path.relpath(p, d)}\n")
        with open(os.path.join(d, "m.obj"), "w") as f:
            f.write("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nusemtl a\nf 1/1 2/2 3/3\n")
        m = sc.add_obj(os.path.join(d, "m.obj"))
    desc = sc.desc()
    ti = desc.meshes[m].submeshes[0].textureIndex[A.SLOT_BASECOLOR]
    t = desc.textures[ti]
    got = np.ctypeslib.as_array(t.texels, shape=(t.height, t.width, 4))
    assert (t.width, t.height) == (ref.shape[1], ref.shape[0]) and np.array_equal(got, ref) and t.srgb == 1


def _png_chunk(kind, data):
    import struct
    import zlib
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)


def _load_png_through_mtl(tmp_path, blob, name):
    """The scene library decodes PNGs when an MTL names one: returns the bound texture record or None when the file was
    rejected (the material then keeps its 1x1 fallback)."""
    (tmp_path / f"{name}.png").write_bytes(blob)
    (tmp_path / f"{name}.mtl").write_text(f"newmtl a\nKd 1 1 1\nmap_Kd {name}.png\n")
    (tmp_path / f"{name}.obj").write_text(f"mtllib {name}.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nusemtl a\nf 1/1 2/2 3/3\n")
    sc = scene.Scene()
    m = sc.add_obj(str(tmp_path / f"{name}.obj"))
    d = sc.desc()
    t = d.textures[d.meshes[m].submeshes[0].textureIndex[A.SLOT_BASECOLOR]]
    return (t.width, t.height, bytes(t.texels[0:4])) if (d.meshes[m].submeshes[0].material.textureFlags & 1) else None


def test_png_decoder_rejects_hostile_files(tmp_path):
    """The decoder does not trust the file: short or repeated IHDR, dimensions whose byte counts would wrap
    size_t, truncated chunks and bad filter bytes are refused — the load fails cleanly, nothing is read or written out
    of bounds and nothing is thrown through the C boundary."""
    import struct
    import zlib
    sig = b"\x89PNG\r\n\x1a\n"

    def ihdr(w, h, depth=8, ctype=6):
        return _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))

    good_rows = b"".join(b"\x00" + bytes([10 * y, 20, 30, 255] * 2) for y in range(2))
    good = sig + ihdr(2, 2) + _png_chunk(b"IDAT", zlib.compress(good_rows)) + _png_chunk(b"IEND", b"")
    assert _load_png_through_mtl(tmp_path, good, "good") == (2, 2, bytes([0, 20, 30, 255]))
    hostile = {
        "short_ihdr": sig + _png_chunk(b"IHDR", b"\x00\x00\x00\x02") + _png_chunk(b"IEND", b""),
        "wrapping_dims": sig + ihdr(1 << 31, 1 << 30, 16, 6) + _png_chunk(b"IDAT", zlib.compress(b"\x00" * 64)) + _png_chunk(b"IEND", b""),
        "huge_dims": sig + ihdr(70000, 70000) + _png_chunk(b"IDAT", zlib.compress(b"\x00" * 64)) + _png_chunk(b"IEND", b""),
        "zero_dims": sig + ihdr(0, 4) + _png_chunk(b"IEND", b""),
        "two_headers": sig + ihdr(2, 2) + ihdr(4096, 4096) + _png_chunk(b"IDAT", zlib.compress(good_rows)) + _png_chunk(b"IEND", b""),
        "truncated_chunk": sig + ihdr(2, 2) + struct.pack(">I", 1 << 30) + b"IDAT" + b"\x00" * 8,
        "bad_filter": sig + ihdr(2, 2) + _png_chunk(b"IDAT", zlib.compress(b"\x09" + bytes(8) + b"\x00" + bytes(8))) + _png_chunk(b"IEND", b""),
        "short_data": sig + ihdr(2, 2) + _png_chunk(b"IDAT", zlib.compress(b"\x00" * 5)) + _png_chunk(b"IEND", b""),
    }
    for name, blob in hostile.items():
        assert _load_png_through_mtl(tmp_path, blob, name) is None, name


def _rgbe_rows(rng, w, h):
    px = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    px[..., 3] = rng.integers(120, 140, (h, w), dtype=np.uint8)
    px[0, :5] = [10, 20, 30, 0]          # exponent 0 decodes to black
    px[1, 8:40] = [200, 100, 50, 130]    # a long run for the run-length coder
    return px


def _rle_channel(values):
    out, i, n = bytearray(), 0, len(values)
    while i < n:
        run = 1
        while i + run < n and run < 127 and values[i + run] == values[i]:
            run += 1
        if run >= 4:
            out += bytes([128 + run, values[i]])
            i += run
        else:
            lit = [values[i]]
            i += 1
            while i < n and len(lit) < 128 and not (i + 3 < n and values[i] == values[i + 1] == values[i + 2] == values[i + 3]):
                lit.append(values[i])
                i += 1
            out += bytes([len(lit)]) + bytes(lit)
    return bytes(out)


def test_radiance_hdr_reader(tmp_path):
    """rts_load_hdr: flat and run-length-encoded RGBE files decode to (r, g, b) * 2^(e - 136); malformed files are
    refused with a message. The texels can go straight into the environment extension's table builder."""
    rng = np.random.default_rng(11)
    w, h = 48, 6
    px = _rgbe_rows(rng, w, h)
    want = np.ones((h, w, 4), np.float32)
    scale = np.where(px[..., 3] > 0, np.ldexp(np.float32(1.0), px[..., 3].astype(np.int32) - 136), 0).astype(np.float32)
    want[..., :3] = px[..., :3].astype(np.float32) * scale[..., None]
    header = b"#?RADIANCE\n# written by the test\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y %d +X %d\n" % (h, w)
    flat = tmp_path / "flat.hdr"
    flat.write_bytes(header + px.tobytes())
    assert np.array_equal(scene.load_hdr(flat), want)
    body = b""
    for y in range(h):
        body += bytes([2, 2, w >> 8, w & 255]) + b"".join(_rle_channel(px[y, :, c].tolist()) for c in range(4))
    rle = tmp_path / "rle.hdr"
    rle.write_bytes(header.replace(b"#?RADIANCE", b"#?RGBE") + body)
    assert len(body) < px.nbytes + 4 * h and np.array_equal(scene.load_hdr(rle), want)
    from metal4_raytracing_b200 import device
    table = device.environment_cdf(scene.load_hdr(rle))
    assert table.shape == ((h + 1) + h * (w + 1) + 65 * (h + 1),) and table[h] == 1.0  # CDF rows + 64-cell guide tables
    for name, data in (("magic.hdr", b"P6\n" + px.tobytes()), ("short.hdr", header + px.tobytes()[:100]),
                       ("format.hdr", header.replace(b"rgbe", b"xyze") + px.tobytes()),
                       ("run.hdr", header + bytes([2, 2, 0, w, 128 + 100, 7]))):
        (tmp_path / name).write_bytes(data)
        with pytest.raises(RuntimeError):
            scene.load_hdr(tmp_path / name)
    with pytest.raises(RuntimeError):
        scene.load_hdr(tmp_path / "absent.hdr")
    # writer: decode(encode(x)) within 1/128 of the pixel's largest channel; zero and negative values come back as 0
    img = np.ones((5, 7, 4), np.float32)
    img[..., :3] = rng.uniform(0, 1, (5, 7, 3)).astype(np.float32) * np.float32(10.0) ** rng.integers(-3, 4, (5, 7, 1))
    img[0, 0, :3] = 0.0
    img[0, 1, :3] = [-1.0, 0.5, 0.25]
    scene.write_hdr(tmp_path / "out.hdr", img)
    back = scene.load_hdr(tmp_path / "out.hdr")
    want = np.maximum(img[..., :3], 0)
    assert back.shape == img.shape and np.array_equal(back[0, 0, :3], [0, 0, 0])
    assert (np.abs(back[..., :3] - want) <= want.max(-1, keepdims=True) / 128 + 1e-30).all()


def test_png_writer_round_trip(tmp_path):
    """rts_write_png (post chain image writer): what it writes decodes to the same pixels with an independent
    decoder (PIL), including a ragged size."""
    from PIL import Image
    rng = np.random.default_rng(5)
    for (h, w) in ((37, 53), (1, 1), (128, 256)):
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        p = tmp_path / f"t_{w}x{h}.png"
        scene.write_png(p, img)
        back = np.asarray(Image.open(p))
        assert back.shape == img.shape and np.array_equal(back, img)


@pytest.mark.parametrize("name", ["K2", "K3small", "K5small"])
def test_procedural_stand_ins_face_outward(name):
    """The closed stand-in meshes (bunny, dragon, robot) must enclose positive volume with their winding and carry
    vertex normals on the same side — inward normals would send every bounce ray into the mesh and make the
    benchmark workload unrepresentative (this caught the torus knot being inside out)."""
    sc, _, _ = scene.Scene.named(name, 64, 64, assets=None)
    m = sc.mesh_arrays(0)
    P = m["positions"][:, :3].astype(np.float64)
    N = m["normals"][:, :3].astype(np.float64)
    idx = np.concatenate([np.asarray(s_).reshape(-1, 3) for s_ in m["submeshes"]])
    a, b, c = P[idx[:, 0]], P[idx[:, 1]], P[idx[:, 2]]
    assert float((a * np.cross(b, c)).sum(1).sum()) > 0.0
    assert float(((np.cross(b - a, c - a) * N[idx].mean(1)).sum(1) > 0).mean()) > 0.99
