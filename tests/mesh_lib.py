"""Procedural closed triangle meshes and signed-distance shapes for the SDF-mesh extension tests
(no mesh file of the reference is copied; resources/bunny.obj is only read where /root/reference exists)."""
import numpy as np

import oracle_lib
from oracle_lib import _f, _i


def icosphere(subdiv=2):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m)); cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v, np.float32), np.array(f, np.int32)


def blob(subdiv=2, radius=0.5, bumps=0.2):
    """Star-shaped closed mesh: an icosphere with a smooth radial displacement."""
    v, f = icosphere(subdiv)
    d = v.astype(np.float64)
    r = radius * (1.0 + bumps * np.sin(3.0 * d[:, 0]) * np.sin(2.0 * d[:, 1] + 1.0) + 0.5 * bumps * np.cos(4.0 * d[:, 2]))
    return (d * r[:, None]).astype(np.float32), f


def build_sdf(verts, tris, cell=0.06, margin=3):
    """Grid over the vertex bounding box plus `margin` cells, built by the restatement's orc_build_sdf."""
    L = oracle_lib.lib()
    lo, hi = verts.min(0) - margin * cell, verts.max(0) + margin * cell
    dims = tuple(int(np.ceil((hi[k] - lo[k]) / cell)) + 1 for k in range(3))
    grid = np.zeros(dims[0] * dims[1] * dims[2], np.float32)
    origin = lo.astype(np.float32)
    v = np.ascontiguousarray(verts, np.float32); t = np.ascontiguousarray(tris, np.int32)
    L.orc_build_sdf(len(v), _f(v), len(t), _i(t), dims[0], dims[1], dims[2], _f(origin), float(cell), _f(grid))
    return dict(dims=dims, origin=origin, cell=float(cell), grid=grid, verts=v, tris=t)


def load_obj(path, max_v=100000, max_t=200000):
    L = oracle_lib.lib()
    v = np.zeros((max_v, 3), np.float32); t = np.zeros((max_t, 3), np.int32)
    nv = np.zeros(1, np.int32); nt = np.zeros(1, np.int32)
    if L.orc_load_obj(path.encode(), _f(v), max_v, _i(t), max_t, _i(nv), _i(nt)) != 0:
        raise FileNotFoundError(path)
    return v[:nv[0]].copy(), t[:nt[0]].copy()
