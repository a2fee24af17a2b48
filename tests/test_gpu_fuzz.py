"""GPU: seeded random scenes mixing everything the path supports -- spheres, boxes, cylinders, planes, fixed and
dynamic, hinges and spherical joints between random pairs, contacts and joints on the same bodies -- against the oracle:
contact lists, Jacobian rows, diagonal blocks and RHS bit-exact; impulses within 5e-4 of the largest impulse and the
state after one step within 1e-4 m / 1e-2 m/s (deeply overlapping random bodies produce large, cancelling impulses)."""
import numpy as np
import pytest

from oracle_lib import Oracle, BOX, CYLINDER, PLANE, SPHERE
from gpu_helpers import assert_bit_equal, context_from_oracles, push_state

pytestmark = pytest.mark.gpu
DT = 0.01


def random_scene(seed, n_bodies=40, n_joints=8):
    rng = np.random.default_rng(seed)
    o = Oracle()
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    o.add_body(1.0, BOX, [14, 0.4, 14], fixed=True, x=[0, 0.2, 0])
    pos, ident = {}, set(rng.choice(np.arange(2, n_bodies), 2 * n_joints, replace=False).tolist())   # jointed bodies: identity pose
    for i in range(2, n_bodies):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        if i in ident:
            q = np.array([0, 0, 0, 1.0])
        x = rng.uniform(-3.0, 3.0, 3) * [1, 0, 1] + [0, rng.uniform(0.75, 2.2), 0]
        kind = rng.integers(0, 10)
        fixed = bool(rng.integers(0, 14) == 0) and i not in ident
        v, w = 0.5 * rng.normal(size=3), rng.normal(size=3)
        if kind < 6:
            o.add_body(0.5 + rng.random(), SPHERE, [0.3 + 0.3 * rng.random()], fixed=fixed, x=x, q=q, v=v, w=w)
        elif kind < 8:
            o.add_body(0.5 + 2 * rng.random(), BOX, 0.5 + rng.random(3), fixed=fixed, x=x, q=q, v=v, w=w)
        else:
            x[1] = 0.9 + 0.3 * rng.random()                    # cylinders near the ground plane so that they touch it
            o.add_body(0.5 + rng.random(), CYLINDER, [0.3 + 0.5 * rng.random(), 0.3 + 0.4 * rng.random()], fixed=fixed, x=x, q=q, v=v, w=w)
        pos[i] = x
    pairs = np.array(sorted(ident)).reshape(-1, 2)
    for j, (a, b) in enumerate(pairs):
        p = 0.5 * (pos[int(a)] + pos[int(b)])                  # anchors meet at the midpoint: no initial joint error
        r0, r1 = p - pos[int(a)], p - pos[int(b)]
        if j % 2:
            o.add_hinge(int(a), int(b), r0, [0, 0, 0, 1], r1, [0, 0, 0, 1])
        else:
            o.add_spherical(int(a), int(b), r0, r1)
    return o


@pytest.mark.parametrize("seed0", [100, 200])
def test_random_mixed_scenes(seed0):
    oracles, k = [], 0
    while len(oracles) < 6:                                   # skip seeds with a sphere centre inside a box: that contact has a
        o = random_scene(seed0 + k); k += 1                   # NaN normal in the reference (CollisionDetect.cpp:144), which is
        probe = random_scene(seed0 + k - 1)                   # reproduced (tests/test_gpu_parity.py) but turns every impulse of
        probe.stage_prepare(); probe.stage_detect()           # the scene into NaN, leaving nothing to compare here
        if probe.n_contacts and np.isfinite(probe.contacts()["n"]).all():
            oracles.append(o)
    ctx = context_from_oracles(oracles, max_contacts=512)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.stage_jacobians(DT); ctx.sync()
    counts = ctx.contact_counts()
    for s, o in enumerate(oracles):
        o.stage_prepare(); o.stage_detect(); o.stage_jacobians()
        assert counts[s] == o.n_contacts and o.n_contacts > 5, (s, counts[s], o.n_contacts)
        g, c = ctx.contacts(s), o.contacts()
        np.testing.assert_array_equal(g["body0"], c["body0"]); np.testing.assert_array_equal(g["body1"], c["body1"])
        for k in ("p", "n", "phi", "t", "b"):
            assert_bit_equal(g[k], c[k], f"scene {s} contact {k}")
        gr, orow, ob = ctx.contact_rows(s), o.contact_rows(), o.solver_blocks(DT)
        for k in ("J0", "J1", "J0Minv", "J1Minv"):
            assert_bit_equal(gr[k], orow[k], f"scene {s} contact rows {k}")
        assert_bit_equal(gr["A"], ob["A"], "contact A"); assert_bit_equal(gr["rhs"], ob["rhs"], "contact rhs")
        gj, oj = ctx.joint_rows(s), o.joint_rows()
        for k in ("J0", "J1", "J0Minv", "J1Minv"):
            assert_bit_equal(gj[k], oj[k], f"scene {s} joint rows {k}")
        assert_bit_equal(gj["rhs"], ob["rhsj"], "joint rhs")
    ctx.stage_solve(DT); ctx.sync()
    lamj = ctx.joint_lambda()
    for s, o in enumerate(oracles):
        o.stage_solve(DT); o.stage_scatter(DT)
        ol = o.contacts()["lam"]
        scale = max(float(np.abs(ol).max()), float(np.abs(o.joint_rows()["lam"]).max()), 1e-3)
        np.testing.assert_allclose(ctx.contacts(s)["lam"], ol, rtol=0, atol=5e-4 * scale + 1e-6, err_msg=f"scene {s} contact lambda")
        np.testing.assert_allclose(lamj[s], o.joint_rows()["lam"], rtol=0, atol=5e-4 * scale + 1e-6, err_msg=f"scene {s} joint lambda")
    ctx.stage_integrate(DT); ctx.sync()
    st = ctx.download_state()
    for s, o in enumerate(oracles):
        o.stage_integrate(DT)
        np.testing.assert_allclose(st["x"][s], o.state()["x"], rtol=0, atol=1e-4)
        np.testing.assert_allclose(st["v"][s], o.state()["v"], rtol=0, atol=1e-2)
    ctx.close()
