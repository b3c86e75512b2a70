import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metal4_raytracing_b200 import device, scene
L = device.lib()
L.rt_selftest_child_boxes.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
ctx = device.Context(0)
for name in ("K1", "K3small", "K2"):
    sc, u, seed = scene.Scene.named(name, 64, 64)
    rnd = device.Renderer(ctx, sc, 64, 64, seeds=scene.seed_image(64, 64, seed))
    for m in range(sc.desc().meshCount):
        out = (C.c_uint64 * 11)()
        device._check(L.rt_selftest_child_boxes(ctx._h, rnd.blas_id(m), 64, 1234, out))
        o = list(out)
        print(name, "mesh", m, "missed", o[0], "extra", o[1], "tests", o[2],
              "first:", "node", o[3], "fast %08x ref %08x" % (o[4], o[5]), "ox", np.array([o[6]], np.uint32).view(np.float32)[0],
              "dx", np.array([o[7]], np.uint32).view(np.float32)[0], "w0.w %08x meta %08x %08x" % (o[8], o[9], o[10]))
    rnd.close()
