"""ctypes bridge to librt_b200.so — the sm_100a ray-tracing hot path (include/rt_b200.h, include/rt_renderer.h).

`Context` wraps rt_create and the kernel-level entry points (the argument-table API that replaces the reference's
Metal bindings, MetalRaytracing/Renderer.swift:1453-1490 and SkinningPass.swift:160-211); `Renderer` wraps the host
orchestrator that mirrors Renderer.swift's createBuffers / updateSkinningAndBLAS / draw. There is no CPU fallback:
a missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", os.environ.get("RT_B200_LIBNAME", "librt_b200.so"))
_lib = None

RTR_FLAG_FP32_IMAGES = 1
RTR_FLAG_REBUILD_SKINNED = 2
RTR_FLAG_GPU_SKELETON = 4
RTR_FLAG_ENABLE_AO = 8
RTR_FLAG_TLAS_REBUILD = 16


class Environment(C.Structure):
    """rt_environment (include/rt_b200.h): RGBA32F equirectangular texels + intensity. An extension, off by default."""
    _fields_ = [("texelsDev", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("intensity", C.c_float),
                ("flags", C.c_uint32), ("cdfDev", C.c_void_p)]


ENV_IMPORTANCE = 1  # RT_ENV_IMPORTANCE
ENV_GUIDED = 2      # RT_ENV_GUIDED: the table carries rt_environment_cdf's guide tables


def environment_cdf(texels):
    """rt_environment_cdf: the sampling table of RT_ENV_IMPORTANCE for (H, W, 4) float32 texels (host side, no GPU)."""
    t = np.ascontiguousarray(texels, np.float32)
    h, w = t.shape[0], t.shape[1]
    L = lib()
    L.rt_environment_cdf_floats.restype = C.c_size_t
    L.rt_environment_cdf_floats.argtypes = [C.c_int32, C.c_int32]
    L.rt_environment_cdf.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    out = np.empty(L.rt_environment_cdf_floats(w, h), np.float32)
    if L.rt_environment_cdf(t.ctypes.data, w, h, out.ctypes.data) != 0:
        raise RuntimeError("rt_environment_cdf failed")
    return out


class DenoiseFrame(C.Structure):
    """rt_denoise_frame: colour, motion, depth, normal image records of one frame."""
    _fields_ = [("color", A.Image), ("motion", A.Image), ("depth", A.Image), ("normal", A.Image)]


class TraceOptions(C.Structure):
    _fields_ = [("tileModulo", C.c_int32), ("tileRemainder", C.c_int32), ("primaryIdsDev", C.c_void_p),
                ("rayCountersDev", C.c_void_p), ("peerAccumulation", C.POINTER(C.c_void_p)),
                ("environment", C.POINTER(Environment)), ("hints", C.c_uint32), ("_pad", C.c_uint32),
                ("sampleModulo", C.c_int32), ("sampleRemainder", C.c_int32)]


class AsInfo(C.Structure):
    _fields_ = [("primitiveCount", C.c_uint32), ("wideNodeCount", C.c_uint32), ("levelCount", C.c_uint32),
                ("_pad", C.c_uint32), ("bytes", C.c_uint64), ("boundsMin", C.c_float * 3),
                ("boundsMax", C.c_float * 3), ("sahCost", C.c_float), ("_pad2", C.c_float)]


# every symbol include/rt_b200.h and include/rt_renderer.h declare (tests/test_abi.py checks the library exports them)
EXPORTS = [
    "rt_create", "rt_destroy", "rt_last_error", "rt_set_stream", "rt_get_stream", "rt_sync", "rt_timer_begin", "rt_timer_end",
    "rt_malloc", "rt_free", "rt_malloc_host", "rt_free_host", "rt_upload", "rt_download", "rt_copy", "rt_memset",
    "rt_blas_build", "rt_blas_refit", "rt_blas_destroy", "rt_tlas_build", "rt_tlas_update", "rt_tlas_refit",
    "rt_tlas_destroy", "rt_host_sync_count",
    "rt_as_get_info", "rt_skin", "rt_trace", "rt_texture_create", "rt_texture_destroy", "rt_launch_count",
    "rt_set_trace_mode", "rt_set_option", "rt_kernel_timing_enable", "rt_kernel_timing_read", "rt_joint_palette",
    "rt_tonemap", "rt_temporal_filter", "rt_download_async", "rt_download_wait", "rt_fence", "rt_fence_wait",
    "rtr_last_error", "rtr_create", "rtr_destroy", "rtr_set_seeds", "rtr_update", "rtr_draw", "rtr_read_image",
    "rtr_image_info", "rtr_reset_accumulation", "rtr_read_mesh_streams", "rtr_get_blas_id", "rtr_get_tlas_id",
    "rtr_mesh_count", "rt_environment_cdf_floats", "rt_environment_cdf", "rt_spatial_filter", "rt_intersect",
]


RAY_HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("instance", np.uint32),
                          ("geometry", np.uint32), ("primitive", np.uint32)])  # rt_ray_hit, 24 bytes


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "this package has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t
    L.rt_last_error.restype = C.c_char_p
    L.rtr_last_error.restype = C.c_char_p
    L.rt_create.argtypes = [i32, C.POINTER(vp)]
    L.rt_destroy.argtypes = [vp]
    L.rt_set_stream.argtypes = [vp, vp]
    L.rt_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.rt_sync.argtypes = [vp]
    L.rt_timer_begin.argtypes = [vp]
    L.rt_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    L.rt_malloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.rt_free.argtypes = [vp, vp]
    L.rt_malloc_host.argtypes = [vp, sz, C.POINTER(vp)]
    L.rt_free_host.argtypes = [vp, vp]
    L.rt_upload.argtypes = [vp, vp, vp, sz]
    L.rt_download.argtypes = [vp, vp, vp, sz]
    L.rt_copy.argtypes = [vp, vp, vp, sz]
    L.rt_download_async.argtypes = [vp, vp, vp, sz, C.POINTER(u64)]
    L.rt_download_wait.argtypes = [vp, u64]
    L.rt_fence.argtypes = [vp, C.POINTER(u64)]
    L.rt_fence_wait.argtypes = [vp, u64]
    L.rt_memset.argtypes = [vp, vp, i32, sz]
    L.rt_blas_build.argtypes = [vp, C.POINTER(A.TriangleGeometry), u32, u32, C.POINTER(u64)]
    L.rt_blas_refit.argtypes = [vp, u64, C.POINTER(A.TriangleGeometry), u32]
    L.rt_blas_destroy.argtypes = [vp, u64]
    L.rt_tlas_build.argtypes = [vp, vp, u32, C.POINTER(u64)]
    L.rt_tlas_update.argtypes = [vp, u64, vp, u32]
    L.rt_tlas_refit.argtypes = [vp, u64, vp, u32]
    L.rt_host_sync_count.restype = u64
    L.rt_host_sync_count.argtypes = [vp]
    L.rt_tlas_destroy.argtypes = [vp, u64]
    L.rt_as_get_info.argtypes = [vp, u64, C.POINTER(AsInfo)]
    L.rt_intersect.argtypes = [vp, u64, vp, u32, u32, vp]
    L.rt_skin.argtypes = [vp, C.POINTER(vp), u32]
    L.rt_trace.argtypes = [vp, C.POINTER(vp), C.POINTER(A.Image), i32, i32, C.POINTER(TraceOptions)]
    L.rt_texture_create.argtypes = [vp, vp, i32, i32, i32, C.POINTER(vp)]
    L.rt_texture_destroy.argtypes = [vp, vp]
    L.rt_launch_count.restype = u64
    L.rt_launch_count.argtypes = [vp]
    L.rt_set_trace_mode.argtypes = [vp, i32]
    L.rt_set_option.argtypes = [vp, C.c_char_p, i32]
    L.rt_joint_palette.argtypes = [vp, vp, vp, vp, u32, vp]
    L.rt_tonemap.argtypes = [vp, C.POINTER(A.Image), vp, u32]
    L.rt_temporal_filter.argtypes = [vp, C.POINTER(DenoiseFrame), C.POINTER(DenoiseFrame), C.POINTER(A.Image),
                                     C.c_float, C.c_float, C.c_float]
    L.rt_spatial_filter.argtypes = [vp, C.POINTER(DenoiseFrame), C.POINTER(A.Image), i32, C.c_float, i32, C.c_float]
    L.rt_kernel_timing_enable.argtypes = [vp, i32]
    L.rt_kernel_timing_read.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(u32)]
    L.rtr_create.argtypes = [vp, C.POINTER(A.SceneDesc), i32, i32, u32, C.POINTER(vp)]
    L.rtr_destroy.argtypes = [vp]
    L.rtr_set_seeds.argtypes = [vp, vp]
    L.rtr_update.argtypes = [vp, C.POINTER(A.SceneDesc)]
    L.rtr_draw.argtypes = [vp, C.POINTER(A.Uniforms), C.POINTER(TraceOptions)]
    L.rtr_read_image.argtypes = [vp, i32, vp, sz]
    L.rtr_image_info.argtypes = [vp, i32, C.POINTER(A.Image)]
    L.rtr_reset_accumulation.argtypes = [vp]
    L.rtr_read_mesh_streams.argtypes = [vp, i32, vp, vp, vp]
    L.rtr_get_blas_id.argtypes = [vp, i32, C.POINTER(u64)]
    L.rtr_get_tlas_id.argtypes = [vp, C.POINTER(u64)]
    L.rtr_mesh_count.argtypes = [vp]
    _lib = L
    return L


class RtError(RuntimeError):
    pass


def _check(rc, renderer=False):
    if rc != 0:
        L = lib()
        msg = (L.rtr_last_error() if renderer else L.rt_last_error()).decode() or L.rt_last_error().decode()
        raise RtError(f"rc={rc}: {msg}")


_FORMAT_DTYPE = {
    A.FORMAT_R32_UINT: (np.uint32, 1), A.FORMAT_R32_FLOAT: (np.float32, 1), A.FORMAT_RG16_FLOAT: (np.float16, 2),
    A.FORMAT_RGBA16_FLOAT: (np.float16, 4), A.FORMAT_R16_FLOAT: (np.float16, 1),
    A.FORMAT_RG32_FLOAT: (np.float32, 2), A.FORMAT_RGBA32_FLOAT: (np.float32, 4),
}


class Context:
    """rt_context: one CUDA device + stream. Raises RtError when no GPU is present."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().rt_create(device, C.byref(h)))
        self._h = h
        self.device = device
        self._owned = []

    def close(self):
        if getattr(self, "_h", None):
            for p in getattr(self, "_owned", []):
                lib().rt_free_host(self._h, p)
            self._owned = []
            lib().rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- memory -------------------------------------------------------------------------------------------
    def malloc(self, nbytes):
        p = C.c_void_p()
        _check(lib().rt_malloc(self._h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr):
        _check(lib().rt_free(self._h, ptr))

    def upload(self, array, ptr=None):
        a = np.ascontiguousarray(array)
        if ptr is None:
            ptr = self.malloc(a.nbytes)
        _check(lib().rt_upload(self._h, ptr, a.ctypes.data, a.nbytes))
        return ptr

    def download(self, ptr, shape, dtype):
        out = np.empty(shape, dtype)
        _check(lib().rt_download(self._h, out.ctypes.data, ptr, out.nbytes))
        return out

    def pinned_array(self, shape, dtype):
        """numpy view of page-locked host memory (rt_malloc_host), freed with the context."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _check(lib().rt_malloc_host(self._h, n, C.byref(p)))
        self._owned.append(p.value)
        buf = (C.c_uint8 * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def download_async(self, ptr, out):
        """rt_download_async into a (pinned) numpy array; returns the ticket for download_wait."""
        t = C.c_uint64()
        _check(lib().rt_download_async(self._h, out.ctypes.data, ptr, out.nbytes, C.byref(t)))
        return t.value

    def download_wait(self, ticket):
        _check(lib().rt_download_wait(self._h, ticket))

    def memset(self, ptr, value, nbytes):
        _check(lib().rt_memset(self._h, ptr, value, nbytes))

    def copy(self, dst, src, nbytes):
        _check(lib().rt_copy(self._h, dst, src, nbytes))

    def sync(self):
        _check(lib().rt_sync(self._h))

    def set_stream(self, cuda_stream):
        _check(lib().rt_set_stream(self._h, cuda_stream))

    @property
    def stream(self):
        """cudaStream_t (int) the context enqueues on (rt_get_stream)."""
        out = C.c_void_p()
        _check(lib().rt_get_stream(self._h, C.byref(out)))
        return out.value or 0

    def set_trace_mode(self, mode):
        """0 = megakernel, 1 = wavefront (default). Both produce identical images."""
        _check(lib().rt_set_trace_mode(self._h, int(mode)))

    def set_option(self, key, value):
        _check(lib().rt_set_option(self._h, key.encode(), int(value)))

    def timer_begin(self):
        _check(lib().rt_timer_begin(self._h))

    def timer_end(self):
        ms = C.c_float()
        _check(lib().rt_timer_end(self._h, C.byref(ms)))
        return ms.value

    @property
    def launches(self):
        return lib().rt_launch_count(self._h)

    @property
    def host_syncs(self):
        """rt_host_sync_count: implicit host<->device synchronisations inside the library so far."""
        return lib().rt_host_sync_count(self._h)

    def tonemap(self, image, srgb=True, flip_y=True):
        """rt_tonemap of a device image record (A.Image): returns (H, W, 4) uint8, Reinhard [+ sRGB] [+ row flip]."""
        n = image.width * image.height * 4
        dst = self.malloc(n)
        _check(lib().rt_tonemap(self._h, C.byref(image), dst, (1 if srgb else 0) | (2 if flip_y else 0)))
        out = self.download(dst, (image.height, image.width, 4), np.uint8)
        self.free(dst)
        return out

    def image_from_array(self, array, fmt):
        """Uploads an (H, W[, C]) array and returns its A.Image record (caller frees image.data)."""
        a = np.ascontiguousarray(array)
        img = A.Image()
        img.data, img.width, img.height, img.format = self.upload(a), a.shape[1], a.shape[0], fmt
        return img

    def temporal_filter(self, current, history, out, history_weight=0.9, depth_tolerance=0.05, normal_threshold=0.9):
        """rt_temporal_filter on DenoiseFrame records; `history` may be None."""
        _check(lib().rt_temporal_filter(self._h, C.byref(current), C.byref(history) if history is not None else None,
                                        C.byref(out), history_weight, depth_tolerance, normal_threshold))

    def spatial_filter(self, frame, out, step=1, depth_sigma=0.02, normal_squarings=5, color_sigma=0.0):
        """rt_spatial_filter: one a-trous pass over frame.color guided by frame.depth / frame.normal."""
        _check(lib().rt_spatial_filter(self._h, C.byref(frame), C.byref(out), step, depth_sigma, normal_squarings,
                                       color_sigma))

    def joint_palette(self, local_trs, parents, inverse_bind):
        """rt_joint_palette on numpy inputs; returns the (J, 16) palette."""
        trs = np.ascontiguousarray(local_trs, np.float32).reshape(-1, 10)
        par = np.ascontiguousarray(parents, np.int32)
        ib = np.ascontiguousarray(inverse_bind, np.float32).reshape(-1, 16)
        n = trs.shape[0]
        d_trs, d_par, d_ib = self.upload(trs), self.upload(par), self.upload(ib)
        d_out = self.malloc(n * 64)
        _check(lib().rt_joint_palette(self._h, d_trs, d_par, d_ib, n, d_out))
        out = self.download(d_out, (n, 16), np.float32)
        for p in (d_trs, d_par, d_ib, d_out):
            self.free(p)
        return out

    KERNEL_CLASSES = ("generate", "trace", "shade", "shadow", "resolve", "skin", "refit", "build", "megakernel", "other")

    def kernel_timing(self, enable=True):
        """Per-kernel-class CUDA-event timing of the library's own launches (rt_kernel_timing_enable)."""
        _check(lib().rt_kernel_timing_enable(self._h, 1 if enable else 0))

    def kernel_times(self):
        """{class: (milliseconds, launches)} since the last read; synchronises the stream."""
        ms = (C.c_float * len(self.KERNEL_CLASSES))()
        n = (C.c_uint32 * len(self.KERNEL_CLASSES))()
        _check(lib().rt_kernel_timing_read(self._h, ms, n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES) if n[i]}

    # -- acceleration structures -----------------------------------------------------------------------------
    def blas_build(self, geoms, flags=A.AS_FLAG_COMPACT):
        arr = (A.TriangleGeometry * max(1, len(geoms)))(*geoms)
        out = C.c_uint64()
        _check(lib().rt_blas_build(self._h, arr, len(geoms), flags, C.byref(out)))
        return out.value

    def blas_refit(self, blas, geoms):
        arr = (A.TriangleGeometry * max(1, len(geoms)))(*geoms)
        _check(lib().rt_blas_refit(self._h, blas, arr, len(geoms)))

    def blas_destroy(self, blas):
        _check(lib().rt_blas_destroy(self._h, blas))

    def tlas_build(self, descriptors_dev, count):
        out = C.c_uint64()
        _check(lib().rt_tlas_build(self._h, descriptors_dev, count, C.byref(out)))
        return out.value

    def tlas_update(self, tlas, descriptors_dev, count):
        _check(lib().rt_tlas_update(self._h, tlas, descriptors_dev, count))

    def tlas_refit(self, tlas, descriptors_dev, count):
        _check(lib().rt_tlas_refit(self._h, tlas, descriptors_dev, count))

    def tlas_destroy(self, tlas):
        _check(lib().rt_tlas_destroy(self._h, tlas))

    def as_info(self, as_id):
        info = AsInfo()
        _check(lib().rt_as_get_info(self._h, as_id, C.byref(info)))
        return info

    def intersect(self, tlas, rays, any_hit=False):
        """rt_intersect: rays = float32 array (n, 8) of origin, tmin, direction, tmax. Returns a structured array with
        the fields of rt_ray_hit (t, u, v, instance, geometry, primitive)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        out = np.zeros(n, RAY_HIT_DTYPE)
        if n == 0:
            return out
        rdev, hdev = self.upload(rays), self.malloc(n * RAY_HIT_DTYPE.itemsize)
        try:
            _check(lib().rt_intersect(self._h, tlas, rdev, n, A.INTERSECT_ANY if any_hit else 0, hdev))
            out = self.download(hdev, (n,), RAY_HIT_DTYPE)
        finally:
            self.free(rdev)
            self.free(hdev)
        return out

    # -- kernels -----------------------------------------------------------------------------------------------
    def skin(self, table, vertex_count):
        """table: dict BufferIndex -> device pointer (Skinning.metal:7-15 bindings)."""
        bufs = (C.c_void_p * A.BUFFER_COUNT)()
        for k, v in table.items():
            bufs[k] = v
        _check(lib().rt_skin(self._h, bufs, vertex_count))

    def trace(self, table, images, uniforms, max_submeshes, options=None):
        """table: dict BufferIndex -> device pointer / TLAS id; images: (A.Image * 9)."""
        bufs = (C.c_void_p * A.BUFFER_COUNT)()
        for k, v in table.items():
            bufs[k] = v
        bufs[A.BUFFER_UNIFORMS] = C.addressof(uniforms)
        _check(lib().rt_trace(self._h, bufs, images, C.sizeof(A.Resource), max_submeshes,
                              C.byref(options) if options is not None else None))

    def texture_create(self, rgba8, srgb=False):
        t = np.ascontiguousarray(rgba8, np.uint8)
        out = C.c_void_p()
        _check(lib().rt_texture_create(self._h, t.ctypes.data, t.shape[1], t.shape[0], int(srgb), C.byref(out)))
        return out.value


class Renderer:
    """rtr_renderer: scene resident in HBM + per-frame update/draw (Renderer.swift's hot-path half)."""

    def __init__(self, ctx, scene, width, height, seeds=None, fp32=False, rebuild_skinned=False, gpu_skeleton=False,
                 enable_ao=False, tlas_rebuild=False):
        self.ctx = ctx
        self.scene = scene
        self.width, self.height = width, height
        flags = ((RTR_FLAG_FP32_IMAGES if fp32 else 0) | (RTR_FLAG_REBUILD_SKINNED if rebuild_skinned else 0) |
                 (RTR_FLAG_GPU_SKELETON if gpu_skeleton else 0) | (RTR_FLAG_ENABLE_AO if enable_ao else 0) |
                 (RTR_FLAG_TLAS_REBUILD if tlas_rebuild else 0))
        desc = scene.desc()
        h = C.c_void_p()
        _check(lib().rtr_create(ctx._h, C.byref(desc), width, height, flags, C.byref(h)), True)
        self._h = h
        self._ids_dev = None
        self._counters_dev = None
        if seeds is not None:
            self.set_seeds(seeds)

    def close(self):
        if getattr(self, "_h", None):
            if self._ids_dev:
                self.ctx.free(self._ids_dev)
            if self._counters_dev:
                self.ctx.free(self._counters_dev)
            lib().rtr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_seeds(self, seeds):
        s = np.ascontiguousarray(seeds, np.uint32)
        assert s.shape == (self.height, self.width)
        _check(lib().rtr_set_seeds(self._h, s.ctypes.data), True)

    def update(self):
        desc = self.scene.desc()
        _check(lib().rtr_update(self._h, C.byref(desc)), True)

    def set_environment(self, texels, intensity=1.0, importance=False):
        """Binds an (H, W, 4) float32 equirectangular environment for the following draws; None unbinds it
        (the reference's behaviour: a miss contributes nothing). importance=True also samples it as a light
        (RT_ENV_IMPORTANCE)."""
        for name in ("_env_dev", "_env_cdf_dev"):
            if getattr(self, name, None):
                self.ctx.free(getattr(self, name))
        self._env_dev, self._env_cdf_dev, self._env = None, None, None
        if texels is not None:
            t = np.ascontiguousarray(texels, np.float32)
            assert t.ndim == 3 and t.shape[2] == 4
            self._env_dev = self.ctx.upload(t)
            flags = 0
            if importance:
                self._env_cdf_dev = self.ctx.upload(environment_cdf(t))
                flags = ENV_IMPORTANCE | ENV_GUIDED
            self._env = Environment(self._env_dev, t.shape[1], t.shape[0], float(intensity), flags, self._env_cdf_dev)

    def draw(self, uniforms, want_ids=False, count_rays=False, tile_modulo=1, tile_remainder=0, peers=None, hints=0,
             sample_modulo=1, sample_remainder=0):
        opt = TraceOptions()
        opt.sampleModulo, opt.sampleRemainder = sample_modulo, sample_remainder
        opt.hints = hints  # RT_TRACE_HINT_*; rtr_draw adds RT_TRACE_HINT_UNTEXTURED itself when the scene has no maps
        opt.tileModulo, opt.tileRemainder = tile_modulo, tile_remainder
        if want_ids:
            if self._ids_dev is None:
                self._ids_dev = self.ctx.malloc(self.width * self.height * 16)
            self.ctx.memset(self._ids_dev, 0xFF, self.width * self.height * 16)
            opt.primaryIdsDev = self._ids_dev
        if count_rays:  # True: reset then count this frame; "accumulate": keep adding to the counters
            if self._counters_dev is None:
                self._counters_dev = self.ctx.malloc(192)
                self.ctx.memset(self._counters_dev, 0, 192)
            if count_rays is True:
                self.ctx.memset(self._counters_dev, 0, 192)
            opt.rayCountersDev = self._counters_dev
        if peers is not None:
            arr = (C.c_void_p * len(peers))(*peers)
            opt.peerAccumulation = arr
        if getattr(self, "_env", None) is not None:
            opt.environment = C.pointer(self._env)
        _check(lib().rtr_draw(self._h, C.byref(uniforms), C.byref(opt)), True)

    def read_ids(self):
        return self.ctx.download(self._ids_dev, (self.height, self.width, 4), np.uint32)

    def reset_ray_counters(self):
        if self._counters_dev is None:
            self._counters_dev = self.ctx.malloc(192)
        self.ctx.memset(self._counters_dev, 0, 192)

    def read_ray_counters(self):
        c = self.ctx.download(self._counters_dev, (24,), np.uint64)
        out = {"closest": int(c[0]), "any": int(c[1]), "hits": int(c[2]), "rays": int(c[0] + c[1])}
        if c[3:].any():  # counter build of the library (RT_COUNT_WORK, tools/count_work.py)
            out["work"] = {"closest": {"nodes": int(c[3]), "triangles": int(c[4]), "entries": int(c[5])},
                           "any": {"nodes": int(c[6]), "triangles": int(c[7]), "entries": int(c[8])},
                           "iterations_histogram": {"bins": [4, 8, 16, 32, 64, 128, 256, "more"],
                                                    "rays": [int(x) for x in c[9:17]], "max": int(c[17]), "sum": int(c[18])},
                           "tail": {"warp_iterations_sum": int(c[19]), "warp_iterations_max": int(c[20]), "warps": int(c[21])}}
        return out

    def image_info(self, index):
        img = A.Image()
        _check(lib().rtr_image_info(self._h, index, C.byref(img)), True)
        return img

    def read_image(self, index=A.TEXTURE_ACCUMULATION, out=None):
        """Host copy of the image bound at `index`; after draw(), index 0 is the frame just rendered. `out` may be
        a pinned array from Context.pinned_array (the copy then runs at full PCIe rate, no staging)."""
        info = self.image_info(index)
        dt, ch = _FORMAT_DTYPE[info.format]
        if out is None:
            out = np.empty((self.height, self.width, ch), dt)
        assert out.dtype == dt and out.shape == (self.height, self.width, ch) and out.flags.c_contiguous
        _check(lib().rtr_read_image(self._h, index, out.ctypes.data, out.nbytes), True)
        return out

    def read_image_async(self, index, out):
        """Starts the read-back of the image bound at `index` into the pinned array `out` without waiting; later draws
        overlap the copy (they write the other accumulation target). Returns a ticket for Context.download_wait."""
        info = self.image_info(index)
        dt, ch = _FORMAT_DTYPE[info.format]
        assert out.dtype == dt and out.shape == (self.height, self.width, ch) and out.flags.c_contiguous
        return self.ctx.download_async(info.data, out)

    def reset_accumulation(self):
        _check(lib().rtr_reset_accumulation(self._h), True)

    def mesh_streams(self, mesh, vertex_count):
        p = np.zeros((vertex_count, 4), np.float32)
        n = np.zeros((vertex_count, 4), np.float32)
        q = np.zeros((vertex_count, 4), np.float32)
        _check(lib().rtr_read_mesh_streams(self._h, mesh, p.ctypes.data, n.ctypes.data, q.ctypes.data), True)
        return p, n, q

    def blas_id(self, mesh):
        out = C.c_uint64()
        _check(lib().rtr_get_blas_id(self._h, mesh, C.byref(out)), True)
        return out.value

    def tlas_id(self):
        out = C.c_uint64()
        _check(lib().rtr_get_tlas_id(self._h, C.byref(out)), True)
        return out.value
