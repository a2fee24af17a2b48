"""Synthetic scenes for benchmarks that the reference's Scenarios.h does not have (numpy only; no oracle, no reference file).

mixed_pile: BASELINE configs[3] -- thousands of mixed spheres / boxes / closed meshes poured into a five-box container, ONE
scene. The mesh is a procedural star-shaped blob (the reference's resources/bunny.obj is not copied into this repo and does
not exist on the GPU box; tests read it where /root/reference is present). Body tables follow slrbs_upload_bodies:
geometry types and parameters as include/collision/Geometry.h, inertia of a mesh = that of the solid box around its
vertices (the convention of collision/SDFMesh.h)."""
import numpy as np

SPHERE, BOX, PLANE, CYLINDER, SDF = 0, 1, 2, 3, 4


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


def blob(subdiv=2, radius=0.35, bumps=0.2):
    """Star-shaped closed mesh: an icosphere with a smooth radial displacement (162 vertices, 320 triangles at subdiv 2)."""
    v, f = icosphere(subdiv)
    d = v.astype(np.float64)
    r = radius * (1.0 + bumps * np.sin(3.0 * d[:, 0]) * np.sin(2.0 * d[:, 1] + 1.0) + 0.5 * bumps * np.cos(4.0 * d[:, 2]))
    return (d * r[:, None]).astype(np.float32), f


def sdf_shape(verts, tris, cell=0.05, margin=3, device=0):
    """Signed-distance grid of a closed mesh, built ON THE DEVICE (slrbs_sdf_build)."""
    from .capi import sdf_build
    lo, hi = verts.min(0) - margin * cell, verts.max(0) + margin * cell
    dims = tuple(int(np.ceil((hi[k] - lo[k]) / cell)) + 1 for k in range(3))
    origin = lo.astype(np.float32)
    grid = sdf_build(verts, tris, dims, origin, cell, device=device)
    return dict(dims=dims, origin=origin, cell=float(cell), grid=grid, verts=np.ascontiguousarray(verts, np.float32))


def _diag_inertia(d):
    I = np.zeros((len(d), 9), np.float32)
    I[:, 0], I[:, 4], I[:, 8] = d[:, 0], d[:, 1], d[:, 2]
    inv = np.zeros_like(I)
    with np.errstate(divide="ignore"):
        inv[:, 0], inv[:, 4], inv[:, 8] = 1.0 / d[:, 0], 1.0 / d[:, 1], 1.0 / d[:, 2]
    return I, inv


def mixed_pile(n, mesh_verts, shape_id=0, seed=7, spacing=1.0):
    """n bodies -- spheres (r 0.35), boxes (0.55^3) and meshes in equal parts, random orientations -- on a lattice above a slab
    inside four walls (all fixed boxes). Returns the arrays of slrbs_upload_bodies for ONE scene."""
    rng = np.random.default_rng(seed)
    side = max(2, int(np.ceil((n / 12.0) ** 0.5)))           # about twelve layers deep
    nlay = (n + side * side - 1) // (side * side)
    i = np.arange(n)
    ix, iz, iy = i % side, (i // side) % side, i // (side * side)
    x = np.stack([(ix - 0.5 * (side - 1)) * spacing, 0.9 + spacing * iy, (iz - 0.5 * (side - 1)) * spacing], axis=1).astype(np.float32)
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    kind = i % 3
    q[kind == 0] = [0, 0, 0, 1]
    gtype = np.where(kind == 0, SPHERE, np.where(kind == 1, BOX, SDF)).astype(np.int32)
    geom = np.zeros((n, 6), np.float32)
    geom[kind == 0, 0] = 0.35
    geom[kind == 1, :3] = 0.55
    geom[kind == 2, 0] = float(shape_id)
    mass = np.ones(n, np.float32)
    bb = (mesh_verts.max(0) - mesh_verts.min(0)).astype(np.float32)
    d = np.zeros((n, 3), np.float32)
    d[kind == 0] = (2.0 / 5.0) * 0.35 * 0.35
    for sel, dim in ((kind == 1, np.array([0.55, 0.55, 0.55], np.float32)), (kind == 2, bb)):
        d[sel] = (1.0 / 12.0) * np.array([dim[1] ** 2 + dim[2] ** 2, dim[0] ** 2 + dim[2] ** 2, dim[0] ** 2 + dim[1] ** 2], np.float32)
    ext = side * spacing + 1.0
    H = nlay * spacing + 3.0
    walls = [([0, 0, 0], [ext + 1.0, 0.4, ext + 1.0])]
    for sx, sz in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        walls.append(([sx * 0.5 * (ext + 0.4), 0.5 * H, sz * 0.5 * (ext + 0.4)], [0.4 if sx else ext + 1.0, H, 0.4 if sz else ext + 1.0]))
    nw = len(walls)
    wx = np.array([w[0] for w in walls], np.float32); wd = np.array([w[1] for w in walls], np.float32)
    wI = (1.0 / 12.0) * np.stack([wd[:, 1] ** 2 + wd[:, 2] ** 2, wd[:, 0] ** 2 + wd[:, 2] ** 2, wd[:, 0] ** 2 + wd[:, 1] ** 2], axis=1)
    I, Iinv = _diag_inertia(np.concatenate([d, wI.astype(np.float32)]))
    wgeom = np.zeros((nw, 6), np.float32); wgeom[:, :3] = wd
    wq = np.tile(np.array([0, 0, 0, 1], np.float32), (nw, 1))
    N = n + nw
    return dict(x=np.concatenate([x, wx]), q=np.concatenate([q.astype(np.float32), wq]), v=np.zeros((N, 3), np.float32),
                w=np.zeros((N, 3), np.float32), mass=np.concatenate([mass, np.ones(nw, np.float32)]), ibody=I, ibodyinv=Iinv,
                fixed=np.concatenate([np.zeros(n, np.int32), np.ones(nw, np.int32)]),
                geom_type=np.concatenate([gtype, np.full(nw, BOX, np.int32)]), geom=np.concatenate([geom, wgeom]),
                jtype=np.zeros(0, np.int32))
