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


import numpy as np  # noqa: E402


def _two_joint_arm():
    """A bar along +y, two joints: the root at the origin and an elbow at (0, 1, 0); vertices above y = 1 follow the
    elbow. Keys: rest pose at t = 0, elbow bent 90 degrees about z at t = 1."""
    ys = np.linspace(0.0, 2.0, 9, dtype=np.float32)
    pos = np.array([[x, y, 0.0] for y in ys for x in (-0.1, 0.1)], np.float32)
    tris = np.array([[2 * i, 2 * i + 1, 2 * i + 2] for i in range(8)] + [[2 * i + 1, 2 * i + 3, 2 * i + 2] for i in range(8)],
                    np.int32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (len(pos), 1))
    upper = pos[:, 1] > 1.0
    ji = np.zeros((len(pos), 4), np.uint16)
    ji[upper, 0] = 1
    jw = np.tile(np.array([1, 0, 0, 0], np.float32), (len(pos), 1))
    ident = [0, 0, 0, 1]
    rest = np.array([[0, 0, 0] + ident + [1, 1, 1], [0, 1, 0] + ident + [1, 1, 1]], np.float32)
    bind = np.stack([np.eye(4, dtype=np.float32).T.reshape(16), np.eye(4, dtype=np.float32).T.reshape(16)])
    bind[1][13] = -1.0  # column-major translation (0, -1, 0): inverse of the elbow's global bind transform
    bent = rest.copy()
    bent[1, 3:7] = [0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)]
    return pos, tris, nrm, ji, jw, [-1, 0], rest, bind, np.array([0.0, 1.0], np.float32), np.stack([rest, bent])


@pytest.fixture
def two_joint_arm():
    return _two_joint_arm()


@pytest.fixture
def two_joint_arm_scene():
    """Factory: a scene with the keyed two-joint arm (scaled x4, one instance, default lights) + uniforms + seeds."""
    def make(w=48, h=48):
        from metal4_raytracing_b200 import scene
        pos, tris, nrm, ji, jw, parents, rest, bind, times, keys = _two_joint_arm()
        trs_scale = np.float32([4, 4, 4, 1, 1, 1, 1, 1, 1, 1])
        sc = scene.Scene()
        m = sc.add_skinned(pos * np.float32(4.0), tris, ji, jw, parents, rest * trs_scale,
                           bind * np.float32([1] * 12 + [4, 4, 4, 1]), normals=nrm)
        sc.set_animation_keys(m, times, keys * trs_scale)
        sc.add_instance(m, (0, -1.5, 0), (0, 0, 0), 0.5)
        sc.default_lights()
        u = scene.default_uniforms(w, h)
        u.camera = scene.default_camera(w, h)
        u.lightCount, u.samplesPerPixel, u.maxBounces = 2, 1, 1
        return sc, u, scene.seed_image(w, h, 5), w, h
    return make
