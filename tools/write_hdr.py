"""Writes the procedural sky (scene.procedural_sky) as a flat Radiance RGBE file: write_hdr.py out.hdr [width height]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metal4_raytracing_b200 import scene


def to_rgbe(rgb):
    m = rgb.max(-1)
    mant, exp = np.frexp(m)                      # m = mant * 2^exp, mant in [0.5, 1)
    scale = np.where(m > 1e-32, 256.0 / np.ldexp(1.0, exp), 0.0)
    out = np.zeros(rgb.shape[:-1] + (4,), np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, exp + 128, 0).astype(np.uint8)
    return out


if __name__ == "__main__":
    path = sys.argv[1]
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 256)
    sky = scene.procedural_sky(w, h)[..., :3].astype(np.float64)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w))
        f.write(to_rgbe(sky).tobytes())
    back = scene.load_hdr(path)[..., :3]
    print("wrote", path, "max rel err", float(np.abs(back - sky).max() / sky.max()))
