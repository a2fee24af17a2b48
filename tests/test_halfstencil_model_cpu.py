"""Host model of the pair search k_detect_rows runs (csrc/detect_kernels.inl, pass 1a): a hashed uniform grid walked with a
HALF stencil. Sphere i visits its own cell and the 13 cells after it in (z, y, x) order; a candidate from the visited BUCKET
counts only if it really lives in the visited CELL (buckets are hashed, so they also hold spheres of other cells); same-cell
pairs are taken from the lower index. Claim checked here against brute force, with a table small enough to force bucket
collisions: every touching pair is found exactly once. (The kernel itself is compared with the all-pairs oracle, contact
list for contact list, in the GPU tests.)"""
import numpy as np
import pytest


def cell_hash(c, mask):
    x, y, z = (np.uint32(int(v) & 0xFFFFFFFF) for v in c)
    with np.errstate(over="ignore"):
        return int((x * np.uint32(73856093)) ^ (y * np.uint32(19349663)) ^ (z * np.uint32(83492791))) & mask


def touching(ci, cj, ri, rj):
    d = ci - cj                                                # CollisionDetect.cpp:111-115: dist < r0 + r1 + 1e-3
    return float(np.sqrt(np.float32(d @ d))) < float(ri + rj) + 1e-3


@pytest.mark.parametrize("n,H,seed", [(300, 16, 1), (600, 64, 2), (600, 1024, 3)])
def test_every_touching_pair_once(n, H, seed):
    rng = np.random.default_rng(seed)
    r = rng.uniform(0.3, 0.5, n).astype(np.float32)
    side = 0.8 * n ** (1.0 / 3.0)                              # dense: several touching partners per sphere
    x = rng.uniform(-0.5 * side, 0.5 * side, (n, 3)).astype(np.float32)
    inv_h = np.float32(1.0) / (np.float32(2.0) * r.max() + np.float32(2e-3))
    cell = np.floor(x * inv_h).astype(np.int64)
    buckets = {}
    for j in range(n):
        buckets.setdefault(cell_hash(cell[j], H - 1), []).append(j)

    found = []
    for i in range(n):
        for cc in range(13, 27):                               # 13 = own cell, 14 .. 26 = the cells after it
            c = cell[i] + np.array([cc % 3 - 1, (cc // 3) % 3 - 1, cc // 9 - 1])
            for j in buckets.get(cell_hash(c, H - 1), []):
                if j == i or (cc == 13 and j < i):
                    continue
                if not touching(x[i], x[j], r[i], r[j]):
                    continue
                if not np.array_equal(cell[j], c):            # lives elsewhere: found from the cell it does live in
                    continue
                found.append((min(i, j), max(i, j)))

    brute = {(i, j) for i in range(n) for j in range(i + 1, n) if touching(x[i], x[j], r[i], r[j])}
    assert len(brute) > n                                      # the scene is dense enough to mean something
    assert len(found) == len(set(found)), "a pair was recorded twice"
    assert set(found) == brute
