"""Independent reference for intersection tests: a tessellated sphere and a float64 brute force (Moeller-Trumbore over
every world-space triangle). It shares no code with the library, the scene loader or the oracle; used by
test_gpu_parity.py::test_intersect_against_brute_force (GPU traversal) and test_oracle.py::test_oracle_traversal_against_brute_force."""
import numpy as np


def uv_sphere(rings, sectors, radius):
    """Vertices (n, 4) float32 and triangles (m, 3) int32 of a latitude / longitude sphere, poles included."""
    v = [(0.0, radius, 0.0)]
    for r in range(1, rings):
        th = np.pi * r / rings
        for s in range(sectors):
            ph = 2.0 * np.pi * s / sectors
            v.append((radius * np.sin(th) * np.cos(ph), radius * np.cos(th), radius * np.sin(th) * np.sin(ph)))
    v.append((0.0, -radius, 0.0))
    tri = []
    for s in range(sectors):
        tri.append((0, 1 + (s + 1) % sectors, 1 + s))
    for r in range(rings - 2):
        a, b = 1 + r * sectors, 1 + (r + 1) * sectors
        for s in range(sectors):
            s1 = (s + 1) % sectors
            tri += [(a + s, a + s1, b + s), (a + s1, b + s1, b + s)]
    last, a = len(v) - 1, 1 + (rings - 2) * sectors
    for s in range(sectors):
        tri.append((last, a + s, a + (s + 1) % sectors))
    verts = np.zeros((len(v), 4), np.float32)
    verts[:, :3] = np.array(v, np.float32)
    return verts, np.array(tri, np.int32)


def brute_force_hits(tris_world, owner, rays):
    """float64 Moeller-Trumbore of every ray against every world-space triangle: per ray the two smallest t in
    (tmin, tmax) and the index of the smallest; independent of the library's BVH, transforms and triangle test."""
    o, d = rays[:, 0:3].astype(np.float64), rays[:, 4:7].astype(np.float64)
    tmin, tmax = rays[:, 3].astype(np.float64), rays[:, 7].astype(np.float64)
    v0, e1, e2 = tris_world[:, 0], tris_world[:, 1] - tris_world[:, 0], tris_world[:, 2] - tris_world[:, 0]
    best = np.full(len(rays), np.inf)
    second = np.full(len(rays), np.inf)
    arg = np.full(len(rays), -1, np.int64)
    margin = np.full(len(rays), np.inf)  # how far inside the winning triangle's edges the hit lies (barycentric units)
    for lo in range(0, len(rays), 256):
        O, D = o[lo:lo + 256, None, :], d[lo:lo + 256, None, :]
        p = np.cross(D, e2[None])
        det = np.einsum("rtk,tk->rt", p, e1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            s = O - v0[None]
            u = np.einsum("rtk,rtk->rt", s, p) * inv
            q = np.cross(s, e1[None])
            v = np.einsum("rtk,rtk->rt", np.broadcast_to(D, q.shape), q) * inv
            t = np.einsum("rtk,tk->rt", q, e2) * inv
        t_all = t
        ok = (np.abs(det) > 0) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > tmin[lo:lo + 256, None]) & (t < tmax[lo:lo + 256, None])
        t = np.where(ok, t, np.inf)
        order = np.argsort(t, axis=1)[:, :2]
        rows = np.arange(t.shape[0])
        best[lo:lo + 256] = t[rows, order[:, 0]]
        second[lo:lo + 256] = t[rows, order[:, 1]]
        arg[lo:lo + 256] = np.where(np.isfinite(t[rows, order[:, 0]]), order[:, 0], -1)
        uu, vv = u[rows, order[:, 0]], v[rows, order[:, 0]]
        inside = np.where(arg[lo:lo + 256] >= 0, np.minimum(np.minimum(uu, vv), 1.0 - uu - vv), 1.0)
        # a ray that passes within a hair of an edge of ANY triangle in range may legitimately come out differently in
        # float32 and float64: margin 0 takes it out of the comparison
        in_range = (np.abs(det) > 0) & (t_all > tmin[lo:lo + 256, None] - 1e-4) & (t_all < tmax[lo:lo + 256, None] + 1e-4)
        edge_any = (in_range & (np.abs(np.minimum(np.minimum(u, v), 1 - u - v)) < 1e-4)).any(axis=1)
        margin[lo:lo + 256] = np.where(edge_any, 0.0, inside)
    return best, second, arg, margin
