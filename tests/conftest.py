import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the native pieces once if a fresh checkout has none (nvcc cross-compiles without a GPU)."""
    need = [os.path.join(ROOT, "metal4_raytracing_b200", "lib", "librt_b200.so"),
            os.path.join(ROOT, "metal4_raytracing_b200", "lib", "librt_scene.so"),
            os.path.join(ROOT, "oracle", "liboracle_rt.so")]
    if not all(os.path.isfile(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def assets():
    from metal4_raytracing_b200 import scene
    d = scene.asset_dir()
    if d is None:
        pytest.skip("reference OBJ assets not staged (run __graft_entry__.build() where /root/reference exists)")
    return d


@pytest.fixture(scope="session")
def gpu_ctx():
    from metal4_raytracing_b200 import device
    ctx = device.Context(0)  # raises if there is no GPU or no library: no CPU fallback
    yield ctx
    ctx.close()
