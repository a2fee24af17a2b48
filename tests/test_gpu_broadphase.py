"""GPU: the hashed-grid broad phase must not change the contact list at all: same pairs, same order, same
bits as the oracle's all-pairs loop (CollisionDetect.cpp:32-34), also with hash collisions, mixed radii,
negative / far coordinates, non-sphere bodies and capacity fall-backs; and it must agree with the device's own
all-pairs path bit for bit over whole trajectories."""
import os

import numpy as np
import pytest

from oracle_lib import Oracle, SPHERE, BOX, PLANE, CYLINDER
from gpu_helpers import context_from_oracles, push_state, assert_bit_equal

pytestmark = pytest.mark.gpu
DT = 0.01


def make(oracles, broadphase, mode=None, **kw):
    os.environ["SLRBS_BROADPHASE"] = str(broadphase)
    if mode:
        os.environ["SLRBS_PGS_MODE"] = mode
    try:
        return context_from_oracles(oracles, **kw)
    finally:
        os.environ.pop("SLRBS_BROADPHASE", None)
        os.environ.pop("SLRBS_PGS_MODE", None)


def check_contacts(ctx, oracles):
    ctx.stage_prepare(); ctx.stage_detect(); ctx.sync()
    cnt = ctx.contact_counts()
    for s, o in enumerate(oracles):
        o.stage_prepare(); o.stage_detect()
        assert cnt[s] == o.n_contacts, (s, cnt[s], o.n_contacts)
        if o.n_contacts:
            g, c = ctx.contacts(s), o.contacts()
            np.testing.assert_array_equal(g["body0"], c["body0"]); np.testing.assert_array_equal(g["body1"], c["body1"])
            for k in ("p", "n", "phi"):
                assert_bit_equal(g[k], c[k], k)


@pytest.mark.parametrize("scene,kw,steps", [("marble_box", {}, [0, 60, 150]), ("marble_box_n", dict(nx=8, nz=7, ny=6), [0, 80]),
                                            ("car", {}, [0, 40]), ("cylinder_on_plane", {}, [40]), ("sphere_stack", {"k": 8}, [30])])
def test_grid_contacts_equal_oracle(scene, kw, steps):
    oracles, o, done = [], Oracle(scene, **kw), 0
    for st in steps:
        o.step(DT, st - done); done = st
        c = Oracle(scene, **kw); c.set_state(**o.state())
        oracles.append(c)
    ctx = make(oracles, 1, max_contacts=4096)
    push_state(ctx, oracles)
    check_contacts(ctx, oracles)
    ctx.close()


def test_grid_random_soup_mixed_radii_and_shapes():
    rng = np.random.default_rng(11)
    oracles = []
    for trial in range(3):
        o = Oracle()
        n = 400
        for i in range(n):
            kind = rng.integers(0, 20)
            x = rng.uniform(-6, 6, 3) + [0, 6, 0] + (1000.0 * trial if trial == 2 else 0.0)    # also far from the origin
            if kind == 0:
                q = rng.normal(size=4); q /= np.linalg.norm(q)
                o.add_body(1.0, BOX, rng.uniform(0.5, 3.0, 3), fixed=bool(rng.integers(0, 2)), x=x, q=q)
            elif kind == 1:
                q = rng.normal(size=4); q /= np.linalg.norm(q)
                o.add_body(1.0, CYLINDER, [rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0)], x=x, q=q)
            else:
                o.add_body(1.0, SPHERE, [float(rng.choice([0.25, 0.5, 0.8]))], x=x, fixed=bool(rng.integers(0, 10) == 0))
        o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True, x=[0, 1000.0 * 0, 0])
        oracles.append(o)
    ctx = make(oracles, 1, max_contacts=8192)
    check_contacts(ctx, oracles)
    assert ctx.contact_counts().min() > 50
    ctx.close()


def test_grid_dense_clump_falls_back_to_all_pairs_rows():
    # > CAND_CAP spheres in one cell neighbourhood: those rows use the all-pairs walk, result unchanged
    rng = np.random.default_rng(3)
    o = Oracle()
    for i in range(300):
        o.add_body(1.0, SPHERE, [0.5], x=rng.uniform(-0.6, 0.6, 3))
    for i in range(300):
        o.add_body(1.0, SPHERE, [0.5], x=rng.uniform(-20, 20, 3) + [60, 0, 0])
    ctx = make([o], 1, max_contacts=60000)
    check_contacts(ctx, [o])
    ctx.close()


@pytest.mark.parametrize("mode", ["thread", "level"])
def test_grid_trajectory_identical_to_all_pairs(mode):
    o = [Oracle("marble_box_n", nx=6, nz=6, ny=8) for _ in range(2)]
    a = make(o, 0, mode, max_contacts=4096)
    b = make(o, 1, mode, max_contacts=4096)
    for k in range(60):
        a.step(DT); b.step(DT)
    a.sync(); b.sync()
    np.testing.assert_array_equal(a.contact_counts(), b.contact_counts())
    for key in "xqvw":
        assert_bit_equal(a.download_state()[key], b.download_state()[key], key)
    a.close(); b.close()
