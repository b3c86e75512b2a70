"""Regenerates tests/golden/k1_128.npz from the CPU oracle (run where /root/reference or assets/_ref exists):
    python tests/golden/make_golden.py
The reference has no golden vectors of its own, so the oracle's output on config K1 — the reference's
own CPU-runnable case: plane.obj + sphere.obj, one point light, 1 spp, primary + shadow rays — is the fixture.
The camera bytes are stored so the test does not depend on libm's tanf/asinf on the machine that runs it."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from metal4_raytracing_b200 import _abi as A, scene  # noqa: E402

w = h = 128
sc, u, seed = scene.Scene.named("K1", w, h)
imgs = oracle.FrameImages(w, h, scene.seed_image(w, h, seed))
stats, ids = oracle.Oracle(sc).render(u, imgs, want_ids=True)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "k1_128.npz")
np.savez_compressed(out, uniforms=np.frombuffer(bytes(u), np.uint8), seed=np.uint32(seed), ids=ids,
                    image=imgs.output, depth=imgs.arrays[A.TEXTURE_DEPTH],
                    stats=np.array([stats["closest"], stats["any"], stats["hits"]], np.uint64))
print("wrote", out, os.path.getsize(out), "bytes", stats)
