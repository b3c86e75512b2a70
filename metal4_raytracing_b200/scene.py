"""Python face of the host scene library (librt_scene.so, include/rt_scene.h).

Plays the role of the reference's Scene/AppScene/Model classes (MetalRaytracing/Scene.swift,
AppScene.swift:11-28, Model.swift:45-261): builds the host-memory inputs of the hot path. Pure CPU.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "librt_scene.so")
_lib = None


def asset_dir():
    """Directory with the reference's OBJ/MTL/PNG assets: the read-only mount when present, else the copy
    that __graft_entry__.build() stages (git-ignored, travels to the GPU box)."""
    env = os.environ.get("RT_ASSET_DIR")
    cands = [env] if env else []
    cands += ["/root/reference/AssetResources",
              os.path.join(os.path.dirname(_HERE), "assets", "_ref", "AssetResources")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "plane.obj")):
            return c
    return None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(_LIB_PATH):
        raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(_LIB_PATH)
    L.rts_last_error.restype = C.c_char_p
    L.rts_scene_create.restype = C.c_void_p
    L.rts_scene_destroy.argtypes = [C.c_void_p]
    L.rts_add_mesh_obj.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.rts_add_mesh_procedural.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.rts_add_mesh_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.c_uint32]
    L.rts_add_mesh_skinned.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                                                         C.c_void_p, C.c_void_p, C.c_void_p]
    L.rts_set_animation_keys.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
    L.rts_save_skinned_mesh.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.rts_load_skinned_mesh.argtypes = [C.c_void_p, C.c_char_p]
    L.rts_set_material.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(A.Material)]
    L.rts_get_material.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(A.Material)]
    L.rts_add_texture_rgba8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.rts_add_texture_procedural.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.rts_bind_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.rts_add_instance.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float]
    L.rts_set_instance_transform.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                             C.c_float]
    L.rts_add_light.argtypes = [C.c_void_p, C.POINTER(A.Light)]
    L.rts_clear_lights.argtypes = [C.c_void_p]
    L.rts_default_lights.argtypes = [C.c_void_p]
    L.rts_make_orbit_camera.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float), C.c_float, C.c_float,
                                        C.c_float, C.c_float, C.POINTER(A.Camera)]
    L.rts_default_camera.argtypes = [C.c_float, C.c_float, C.POINTER(A.Camera)]
    L.rts_animate.argtypes = [C.c_void_p, C.c_double]
    L.rts_scene_create_named.restype = C.c_void_p
    L.rts_scene_create_named.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(A.Uniforms),
                                         C.POINTER(C.c_uint32)]
    L.rts_scene_get_desc.argtypes = [C.c_void_p, C.POINTER(A.SceneDesc)]
    L.rts_fill_seed_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32]
    L.rts_default_uniforms.argtypes = [C.c_int, C.c_int, C.POINTER(A.Uniforms)]
    L.rts_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    L.rts_load_hdr.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_float))]
    L.rts_free.argtypes = [C.c_void_p]
    L.rts_write_hdr.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    _lib = L
    return L


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _err():
    return lib().rts_last_error().decode()


class Scene:
    """Owns an rts_scene handle. `desc()` returns the flat view both the oracle and the GPU renderer consume."""

    def __init__(self, handle=None):
        self._h = handle if handle is not None else lib().rts_scene_create()
        self._desc = None

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rts_scene_destroy(self._h)
            self._h = None

    @classmethod
    def named(cls, name, width, height, assets="auto"):
        """Benchmark scenes of SURVEY.md §8(d): K1..K5 (+K2tex, K3glass, K3small, K4small, K5small, appscene).
        Returns (scene, uniforms, seed)."""
        u = A.Uniforms()
        seed = C.c_uint32(0)
        d = asset_dir() if assets == "auto" else assets
        h = lib().rts_scene_create_named(name.encode(), d.encode() if d else None, width, height, C.byref(u),
                                         C.byref(seed))
        if not h:
            raise RuntimeError(_err())
        return cls(h), u, seed.value

    def _check(self, r):
        if r < 0:
            raise RuntimeError(_err())
        self._desc = None
        return r

    def add_obj(self, path, glass=False):
        return self._check(lib().rts_add_mesh_obj(self._h, path.encode(), int(glass)))

    def add_procedural(self, kind, p0=0, p1=0, p2=0, p3=0):
        return self._check(lib().rts_add_mesh_procedural(self._h, kind.encode(), p0, p1, p2, p3))

    def add_raw(self, positions, indices, normals=None, uvs=None):
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(indices, np.int32).reshape(-1, 3)
        n = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        t = None if uvs is None else np.ascontiguousarray(uvs, np.float32).reshape(-1, 2)
        return self._check(lib().rts_add_mesh_raw(
            self._h, p.ctypes.data, n.ctypes.data if n is not None else None,
            t.ctypes.data if t is not None else None, len(p), i.ctypes.data, len(i)))

    def get_material(self, mesh, submesh=0):
        m = A.Material()
        self._check(lib().rts_get_material(self._h, mesh, submesh, C.byref(m)))
        return m

    def set_material(self, mesh, submesh, m):
        self._check(lib().rts_set_material(self._h, mesh, submesh, C.byref(m)))

    def add_texture(self, rgba8, srgb=False):
        t = np.ascontiguousarray(rgba8, np.uint8)
        assert t.ndim == 3 and t.shape[2] == 4
        return self._check(lib().rts_add_texture_rgba8(self._h, t.ctypes.data, t.shape[1], t.shape[0], int(srgb)))

    def add_texture_procedural(self, kind, width, height, seed=0, srgb=False):
        return self._check(lib().rts_add_texture_procedural(self._h, kind.encode(), width, height, seed, int(srgb)))

    def bind_texture(self, mesh, submesh, slot, texture):
        self._check(lib().rts_bind_texture(self._h, mesh, submesh, slot, texture))

    def add_instance(self, mesh, position=(0, 0, 0), rotation=(0, 0, 0), scale=1.0):
        return self._check(lib().rts_add_instance(self._h, mesh, _f3(position), _f3(rotation), float(scale)))

    def set_instance_transform(self, instance, position, rotation, scale):
        self._check(lib().rts_set_instance_transform(self._h, instance, _f3(position), _f3(rotation), float(scale)))

    def add_light(self, light):
        return self._check(lib().rts_add_light(self._h, C.byref(light)))

    def clear_lights(self):
        lib().rts_clear_lights(self._h)
        self._desc = None

    def default_lights(self):
        lib().rts_default_lights(self._h)
        self._desc = None

    def animate(self, t):
        self._check(lib().rts_animate(self._h, float(t)))

    def desc(self):
        d = A.SceneDesc()
        self._check(lib().rts_scene_get_desc(self._h, C.byref(d)))
        self._desc = d
        return d

    def add_skinned(self, positions, indices, joint_indices, joint_weights, parents, rest_trs, inverse_bind,
                    normals=None, uvs=None):
        """rts_add_mesh_skinned: a skinned mesh from arrays (joint_indices (V, 4) uint16, joint_weights (V, 4),
        parents (J,), rest_trs (J, 10), inverse_bind (J, 16) column-major)."""
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(indices, np.int32).reshape(-1, 3)
        n = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        t = None if uvs is None else np.ascontiguousarray(uvs, np.float32).reshape(-1, 2)
        ji = np.ascontiguousarray(joint_indices, np.uint16).reshape(-1, 4)
        jw = np.ascontiguousarray(joint_weights, np.float32).reshape(-1, 4)
        par = np.ascontiguousarray(parents, np.int32).reshape(-1)
        rest = np.ascontiguousarray(rest_trs, np.float32).reshape(-1, 10)
        bind = np.ascontiguousarray(inverse_bind, np.float32).reshape(-1, 16)
        assert len(ji) == len(jw) == len(p) and len(rest) == len(bind) == len(par)
        return self._check(lib().rts_add_mesh_skinned(
            self._h, p.ctypes.data, n.ctypes.data if n is not None else None, t.ctypes.data if t is not None else None,
            ji.ctypes.data, jw.ctypes.data, len(p), i.ctypes.data, len(i), len(par), par.ctypes.data, rest.ctypes.data,
            bind.ctypes.data))

    def set_animation_keys(self, mesh, times, trs):
        """rts_set_animation_keys: times (K,) ascending seconds, trs (K, J, 10); None / empty removes the clip."""
        if times is None or len(times) == 0:
            return self._check(lib().rts_set_animation_keys(self._h, mesh, 0, None, None))
        tm = np.ascontiguousarray(times, np.float32).reshape(-1)
        k = np.ascontiguousarray(trs, np.float32).reshape(len(tm), -1, 10)
        return self._check(lib().rts_set_animation_keys(self._h, mesh, len(tm), tm.ctypes.data, k.ctypes.data))

    def save_skinned(self, mesh, path):
        return self._check(lib().rts_save_skinned_mesh(self._h, mesh, str(path).encode()))

    def load_skinned(self, path):
        return self._check(lib().rts_load_skinned_mesh(self._h, str(path).encode()))

    # ---- numpy views (copies) for tests ---------------------------------------------------------------
    def mesh_arrays(self, mesh):
        d = self.desc()
        m = d.meshes[mesh]
        n = m.vertexCount
        out = {
            "positions": np.ctypeslib.as_array(m.positions, shape=(n, 4)).copy(),
            "normals": np.ctypeslib.as_array(m.normals, shape=(n, 4)).copy(),
            "uvs": np.ctypeslib.as_array(m.uvs, shape=(n, 2)).copy() if m.uvs else None,
            "jointIndices": np.ctypeslib.as_array(m.jointIndices, shape=(n, 4)).copy() if m.jointIndices else None,
            "jointWeights": np.ctypeslib.as_array(m.jointWeights, shape=(n, 4)).copy() if m.jointWeights else None,
            "jointMatrices": (np.ctypeslib.as_array(m.jointMatrices, shape=(m.jointCount, 16)).copy()
                              if m.jointCount else None),
            "jointLocalTRS": (np.ctypeslib.as_array(m.jointLocalTRS, shape=(m.jointCount, 10)).copy()
                              if m.jointCount and m.jointLocalTRS else None),
            "jointParents": (np.ctypeslib.as_array(m.jointParents, shape=(m.jointCount,)).copy()
                             if m.jointCount and m.jointParents else None),
            "submeshes": [],
        }
        for k in range(m.submeshCount):
            sm = m.submeshes[k]
            out["submeshes"].append(np.ctypeslib.as_array(sm.indices, shape=(sm.triangleCount, 3)).copy())
        return out


def load_hdr(path):
    """rts_load_hdr: a Radiance RGBE picture as (H, W, 4) float32 (alpha 1), e.g. for Renderer.set_environment."""
    w, h, p = C.c_int(), C.c_int(), C.POINTER(C.c_float)()
    if lib().rts_load_hdr(str(path).encode(), C.byref(w), C.byref(h), C.byref(p)) != 0:
        raise RuntimeError(_err())
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    finally:
        lib().rts_free(p)


def write_hdr(path, rgba):
    """rts_write_hdr: (H, W, 4) float32 as a Radiance RGBE picture."""
    a = np.ascontiguousarray(rgba, np.float32)
    assert a.ndim == 3 and a.shape[2] == 4
    if lib().rts_write_hdr(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0]) != 0:
        raise RuntimeError(_err())


def default_uniforms(width, height):
    u = A.Uniforms()
    lib().rts_default_uniforms(width, height, C.byref(u))
    return u


def default_camera(width, height):
    c = A.Camera()
    lib().rts_default_camera(float(width), float(height), C.byref(c))
    return c


def orbit_camera(width, height, target, azimuth, elevation, distance, fov_degrees=45.0):
    c = A.Camera()
    lib().rts_make_orbit_camera(float(width), float(height), _f3(target), float(azimuth), float(elevation),
                                float(distance), float(fov_degrees), C.byref(c))
    return c


def seed_image(width, height, seed):
    img = np.empty((height, width), np.uint32)
    lib().rts_fill_seed_image(img.ctypes.data, width, height, seed)
    return img


def make_light(kind, position=(0, 0, 0), color=(1, 1, 1), direction=(0, -1, 0), cone_angle=0.0,
               forward=(0, -1, 0), right=(1, 0, 0), up=(0, 0, 1)):
    l = A.Light()
    l.type = kind
    l.position.set(*position)
    l.color.set(*color)
    l.direction.set(*direction)
    l.coneAngle = cone_angle
    l.forward.set(*forward)
    l.right.set(*right)
    l.up.set(*up)
    return l


def procedural_sky(width=1024, height=512, sun_dir=(0.35, 0.8, -0.45), sun_radiance=40.0):
    """Deterministic HDR equirectangular sky (float32 RGBA, row 0 = straight up) standing in for the reference's
    vulture_hide_4k.hdr, which is absent from the mount: horizon-to-zenith gradient, a darker ground
    half and a soft sun disc. Same layout as rt_environment: u = atan2(z, x) / 2pi + 0.5, v = acos(y) / pi."""
    v = (np.arange(height, dtype=np.float64) + 0.5) / height
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    theta, phi = np.pi * v[:, None], 2.0 * np.pi * (u[None, :] - 0.5)
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), np.sin(theta) * np.sin(phi)], -1)
    up = np.clip(d[..., 1], 0.0, 1.0)[..., None]
    sky = (1.0 - up) * np.array([0.9, 0.95, 1.0]) + up * np.array([0.25, 0.45, 0.9])
    ground = np.array([0.18, 0.16, 0.14]) * np.ones_like(sky)
    img = np.where(d[..., 1:2] >= 0.0, sky, ground)
    s = np.asarray(sun_dir, np.float64)
    s = s / np.linalg.norm(s)
    cosang = np.clip((d * s).sum(-1), -1.0, 1.0)
    img = img + sun_radiance * np.exp(-((1.0 - cosang) / 0.0015))[..., None] * np.array([1.0, 0.93, 0.8])
    out = np.ones((height, width, 4), np.float32)
    out[..., :3] = img.astype(np.float32)
    return out


def write_png(path, rgba8):
    """(H, W, 4) uint8, rows top to bottom -> PNG file (rts_write_png, the post chain's image writer)."""
    a = np.ascontiguousarray(rgba8, np.uint8)
    assert a.ndim == 3 and a.shape[2] == 4
    if lib().rts_write_png(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0]) != 0:
        raise RuntimeError(f"rts_write_png failed for {path}")
