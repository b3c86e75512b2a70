"""CPU oracle bindings — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
PARITY UNPINNED: the reference (Swift + Metal, Apple frameworks) can neither be built nor run here and ships no
tests or golden vectors, so liboracle_rt.so — a restatement of MetalRaytracing/Raytracing.metal:220-831 and
MetalRaytracing/Skinning.metal:7-49 over a plain SAH BVH — is itself the specification (oracle/README.md).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from metal4_raytracing_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_rt.so")
_lib = None


def build(force=False):
    if force or not os.path.isfile(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.oracle_create.restype = C.c_void_p
    L.oracle_create.argtypes = [C.POINTER(A.SceneDesc), C.c_int]
    L.oracle_destroy.argtypes = [C.c_void_p]
    L.oracle_update.argtypes = [C.c_void_p, C.POINTER(A.SceneDesc)]
    L.oracle_render.argtypes = [C.c_void_p, C.POINTER(A.Uniforms), C.POINTER(A.Image), C.c_void_p,
                                C.POINTER(C.c_uint64), C.c_int, C.c_int]
    L.oracle_get_mesh_streams.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oracle_skin.argtypes = [C.POINTER(C.c_void_p), C.c_uint32]
    L.oracle_halton.restype = C.c_float
    L.oracle_halton.argtypes = [C.c_int, C.c_int]
    L.oracle_intersect_triangle.argtypes = [C.POINTER(C.c_float)] * 5 + [C.c_float, C.c_float, C.POINTER(C.c_float)]
    L.oracle_sample_texture.argtypes = [C.POINTER(A.Texture2D), C.c_float, C.c_float, C.POINTER(C.c_float)]
    L.oracle_invert_affine.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.oracle_float_to_half.restype = C.c_uint16
    L.oracle_float_to_half.argtypes = [C.c_float]
    L.oracle_half_to_float.restype = C.c_float
    L.oracle_half_to_float.argtypes = [C.c_uint16]
    L.oracle_trace_ray.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float,
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    L.oracle_thread_count.argtypes = [C.c_void_p]
    L.oracle_set_environment.argtypes = [C.c_void_p, C.c_void_p]
    L.oracle_sample_environment.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    _lib = L
    return L


_FORMAT_DTYPE = {
    A.FORMAT_R32_UINT: (np.uint32, 1), A.FORMAT_R32_FLOAT: (np.float32, 1), A.FORMAT_RG16_FLOAT: (np.float16, 2),
    A.FORMAT_RGBA16_FLOAT: (np.float16, 4), A.FORMAT_R16_FLOAT: (np.float16, 1),
    A.FORMAT_RG32_FLOAT: (np.float32, 2), A.FORMAT_RGBA32_FLOAT: (np.float32, 4),
}


def new_image(width, height, fmt):
    dt, ch = _FORMAT_DTYPE[fmt]
    return np.zeros((height, width, ch), dt)


def image_record(arr, fmt):
    h, w = arr.shape[:2]
    return A.Image(arr.ctypes.data, w, h, fmt, 0)


class FrameImages:
    """The nine images bound at TextureIndex 0..8, as numpy arrays (host memory)."""

    def __init__(self, width, height, seeds, fp32=False, gbuffer=False):
        rgba = A.FORMAT_RGBA32_FLOAT if fp32 else A.FORMAT_RGBA16_FLOAT
        rg = A.FORMAT_RG32_FLOAT if fp32 else A.FORMAT_RG16_FLOAT
        self.width, self.height = width, height
        self.formats = [rgba, rgba, A.FORMAT_R32_UINT, A.FORMAT_R32_FLOAT, rg, rgba, rgba, rgba,
                        A.FORMAT_R32_FLOAT if fp32 else A.FORMAT_R16_FLOAT]
        self.arrays = [new_image(width, height, f) for f in self.formats]
        self.arrays[A.TEXTURE_RANDOM][..., 0] = seeds
        self.gbuffer = gbuffer

    def records(self):
        recs = (A.Image * A.TEXTURE_COUNT)()
        for i, (arr, f) in enumerate(zip(self.arrays, self.formats)):
            recs[i] = image_record(arr, f)
        return recs

    def swap(self):
        """Renderer.swift:1492-1494: swap accumulation targets after each dispatch."""
        a = self.arrays
        a[A.TEXTURE_ACCUMULATION], a[A.TEXTURE_PREVIOUS_ACCUMULATION] = (a[A.TEXTURE_PREVIOUS_ACCUMULATION],
                                                                         a[A.TEXTURE_ACCUMULATION])

    @property
    def output(self):
        """Image written by the last dispatch (before swap)."""
        return self.arrays[A.TEXTURE_PREVIOUS_ACCUMULATION]


class Oracle:
    def __init__(self, scene, threads=0):
        self.scene = scene  # keep host memory alive
        self._desc = scene.desc()
        self._h = lib().oracle_create(C.byref(self._desc), threads)
        if not self._h:
            raise RuntimeError("oracle_create failed")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.oracle_destroy(self._h)
            self._h = None

    @property
    def threads(self):
        return lib().oracle_thread_count(self._h)

    def set_environment(self, texels, intensity=1.0, importance=False):
        """Environment extension (include/rt_b200.h rt_environment); texels (H, W, 4) float32 or None.
        importance=True: RT_ENV_IMPORTANCE with the oracle's own table."""
        from metal4_raytracing_b200.device import Environment
        if texels is None:
            self._env_texels = self._env_cdf = None
            lib().oracle_set_environment(self._h, None)
            return
        self._env_texels = np.ascontiguousarray(texels, np.float32)  # kept alive: the oracle reads it in place
        self._env_cdf = environment_cdf(self._env_texels) if importance else None
        env = Environment(self._env_texels.ctypes.data, self._env_texels.shape[1], self._env_texels.shape[0],
                          float(intensity), 1 if importance else 0,
                          self._env_cdf.ctypes.data if importance else None)
        if lib().oracle_set_environment(self._h, C.byref(env)) != 0:
            raise RuntimeError("oracle_set_environment failed")

    def set_enable_ao(self, enable=True):
        """The reference's compile-time ENABLE_AO switch (ShaderTypes.h:155-157), off by default."""
        L = lib()
        L.oracle_set_enable_ao.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_enable_ao.restype = None
        L.oracle_set_enable_ao(self._h, 1 if enable else 0)

    def update(self):
        self._desc = self.scene.desc()
        if lib().oracle_update(self._h, C.byref(self._desc)) != 0:
            raise RuntimeError("oracle_update failed")

    def render(self, uniforms, images, want_ids=False, tile_modulo=1, tile_remainder=0):
        """One kernel dispatch. Returns (stats dict, primary id array or None)."""
        ids = np.full((uniforms.height, uniforms.width, 4), 0xFFFFFFFF, np.uint32) if want_ids else None
        stats = (C.c_uint64 * 3)()
        recs = images.records()
        r = lib().oracle_render(self._h, C.byref(uniforms), recs, ids.ctypes.data if want_ids else None, stats,
                                tile_modulo, tile_remainder)
        if r != 0:
            raise RuntimeError("oracle_render failed")
        return {"closest": stats[0], "any": stats[1], "hits": stats[2], "rays": stats[0] + stats[1]}, ids

    def mesh_streams(self, mesh, vertex_count):
        p = np.zeros((vertex_count, 4), np.float32)
        n = np.zeros((vertex_count, 4), np.float32)
        q = np.zeros((vertex_count, 4), np.float32)
        lib().oracle_get_mesh_streams(self._h, mesh, p.ctypes.data, n.ctypes.data, q.ctypes.data)
        return p, n, q

    def trace_ray(self, origin, direction, tmin=0.0, tmax=float("inf")):
        ids = (C.c_uint32 * 4)()
        tuv = (C.c_float * 3)()
        o = (C.c_float * 3)(*origin)
        d = (C.c_float * 3)(*direction)
        hit = lib().oracle_trace_ray(self._h, o, d, tmin, tmax, ids, tuv)
        return bool(hit), tuple(ids)[1:], tuple(tuv)


def halton(i, d):
    return lib().oracle_halton(int(i), int(d))


def intersect_triangle(origin, direction, v0, v1, v2, tmin=0.0, tmax=float("inf")):
    f = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    out = (C.c_float * 3)()
    hit = lib().oracle_intersect_triangle(f(origin), f(direction), f(v0), f(v1), f(v2), tmin, tmax, out)
    return (bool(hit), tuple(out))


def sample_texture(rgba8, u, v, srgb=False):
    t = np.ascontiguousarray(rgba8, np.uint8)
    rec = A.Texture2D(t.ctypes.data, t.shape[1], t.shape[0], int(srgb), 0)
    out = (C.c_float * 4)()
    lib().oracle_sample_texture(C.byref(rec), u, v, out)
    return tuple(out)


def invert_affine(m4x3):
    m = np.ascontiguousarray(m4x3, np.float32).reshape(12)
    inv = np.zeros(12, np.float32)
    lib().oracle_invert_affine(m.ctypes.data_as(C.POINTER(C.c_float)), inv.ctypes.data_as(C.POINTER(C.c_float)))
    return inv


def skin(rest_pos4, rest_nrm4, joint_idx, joint_w, matrices):
    """skinningKernel over host arrays; returns (positions4, normals4)."""
    n = len(rest_pos4)
    rp = np.ascontiguousarray(rest_pos4, np.float32)
    rn = np.ascontiguousarray(rest_nrm4, np.float32)
    ji = np.ascontiguousarray(joint_idx, np.uint16)
    jw = np.ascontiguousarray(joint_w, np.float32)
    mt = np.ascontiguousarray(matrices, np.float32)
    op = np.zeros((n, 4), np.float32)
    on = np.zeros((n, 4), np.float32)
    bufs = (C.c_void_p * A.BUFFER_COUNT)()
    bufs[A.BUFFER_REST_POSITIONS] = rp.ctypes.data
    bufs[A.BUFFER_REST_NORMALS] = rn.ctypes.data
    bufs[A.BUFFER_JOINT_INDICES] = ji.ctypes.data
    bufs[A.BUFFER_JOINT_WEIGHTS] = jw.ctypes.data
    bufs[A.BUFFER_JOINT_MATRICES] = mt.ctypes.data
    bufs[A.BUFFER_SKINNED_POSITIONS] = op.ctypes.data
    bufs[A.BUFFER_SKINNED_NORMALS] = on.ctypes.data
    lib().oracle_skin(bufs, n)
    return op, on


def environment_cdf(texels):
    """The oracle's restatement of rt_environment_cdf (include/rt_b200.h): marginal + per-row conditional table."""
    t = np.ascontiguousarray(texels, np.float32)
    h, w = t.shape[0], t.shape[1]
    out = np.empty((h + 1) + h * (w + 1), np.float32)
    L = lib()
    L.oracle_environment_cdf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    if L.oracle_environment_cdf(t.ctypes.data, w, h, out.ctypes.data) != 0:
        raise RuntimeError("oracle_environment_cdf failed")
    return out


def sample_environment(texels, direction, intensity=1.0):
    """KAT probe of the environment extension's equirectangular lookup (include/rt_b200.h rt_environment)."""
    from metal4_raytracing_b200.device import Environment
    t = np.ascontiguousarray(texels, np.float32)
    env = Environment(t.ctypes.data, t.shape[1], t.shape[0], float(intensity), 0, None)
    d = (C.c_float * 3)(*[float(x) for x in direction])
    out = (C.c_float * 3)()
    lib().oracle_sample_environment(C.byref(env), d, out)
    return np.array(out[:], np.float32)
