"""metal4_raytracing_b200 — B200-native ray-tracing hot path behind the reference's ShaderTypes.h ABI.

`scene` (pure CPU) builds the host inputs; `device` (added with the CUDA library) is the C-ABI bridge to the
sm_100a kernels and raises if librt_b200.so is missing — there is no CPU fallback in this package.
"""
from . import _abi  # noqa: F401
