"""Pins the C restatement (oracle/slrbs_oracle.c) against the REFERENCE'S OWN SOURCES.

oracle/_ref/libslrbs_ref.so is the reference's unmodified src/{rigidbody,collision,contact,joint,solvers}
compiled where they lie under /root/reference against the Eigen API shim oracle/ref_shim (see
oracle/Makefile, target `ref`).  Every comparison below is BIT-EXACT: states, contact lists (pairs, order,
p, n, t, b, pene, lambda), Jacobian rows and joint impulses, for all five Scenarios.h scenes, the two
synthetic scenes, the three solvers, save/restore and the pre-step force hook.  What this pins is the
reference's control flow, orderings, thresholds and data flow; what it cannot pin is Eigen's internal
summation order (the shim and the restatement share the choices documented in oracle/slrbs_oracle.h).
"""
import numpy as np
import pytest

from oracle_lib import Oracle, Reference, ref_available
from gpu_helpers import assert_bit_equal

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference; make -C oracle ref)")
DT = 0.01

SCENES = [("marble_box", {}, 90), ("sphere_on_box", {}, 150), ("swinging_boxes", {}, 120),
          ("cylinder_on_plane", {}, 150), ("car", {}, 150), ("sphere_stack", {"k": 8}, 120),
          ("marble_box_n", dict(nx=4, nz=3, ny=5), 120)]


def assert_same_system(o, r, what):
    so, sr = o.state(), r.state()
    for k in so:
        assert_bit_equal(so[k], sr[k], f"{what} state {k}")
    assert o.n_contacts == r.n_contacts, (what, o.n_contacts, r.n_contacts)
    co, cr = o.contacts(), r.contacts()
    np.testing.assert_array_equal(co["body0"], cr["body0"], err_msg=what)
    np.testing.assert_array_equal(co["body1"], cr["body1"], err_msg=what)
    for k in ("p", "n", "t", "b", "phi", "lam"):
        assert_bit_equal(co[k], cr[k], f"{what} contact {k}")
    ro, rr = o.contact_rows(), r.contact_rows()
    for k in ro:
        assert_bit_equal(ro[k], rr[k], f"{what} contact rows {k}")
    jo, jr = o.joint_rows(), r.joint_rows()
    for k in jo:
        assert_bit_equal(jo[k], jr[k], f"{what} joint rows {k}")
    do, dr = o.body_dyn(), r.body_dyn()
    dyn = o.body_params()["fixed"] == 0
    for k in ("f", "tau", "fc", "tauc", "Iinv"):
        assert_bit_equal(do[k], dr[k], f"{what} body {k}")
    assert_bit_equal(do["I"][dyn], dr["I"][dyn], f"{what} body I")     # I of fixed bodies is uninitialised in the reference


@pytest.mark.parametrize("scene,kw,steps", SCENES)
def test_scene_trajectory_bit_identical(scene, kw, steps):
    o, r = Oracle(scene, **kw), Reference(scene, **kw)
    po, pr = o.body_params(), r.body_params()
    for k in po:
        a, b = np.asarray(po[k]), np.asarray(pr[k])
        assert np.array_equal(a, b, equal_nan=True), (scene, k)       # plane IbodyInv is inf/NaN in both
    jo, jr = o.joints(), r.joints()
    for k in jo:
        np.testing.assert_array_equal(jo[k], jr[k])
    for i in range(steps):
        o.step(DT); r.step(DT)
        if i % 10 == 0 or i == steps - 1:
            assert_same_system(o, r, f"{scene} step {i + 1}")


@pytest.mark.parametrize("solver_id", [1, 2])
@pytest.mark.parametrize("scene", ["swinging_boxes", "car"])
def test_krylov_solvers_bit_identical(scene, solver_id):
    o, r = Oracle(scene), Reference(scene)
    for s in (o, r):
        s.set_params(0.8, solver_id, 10, True)
    for i in range(60):
        o.step(DT); r.step(DT)
        if i % 6 == 0:
            assert_same_system(o, r, f"{scene} solver {solver_id} step {i + 1}")


def test_params_save_restore_and_prestep_forces():
    o, r = Oracle("marble_box"), Reference("marble_box")
    for s in (o, r):
        s.set_params(0.4, 0, 7, True); s.step(DT, 20); s.save_state(); s.step(DT, 15); s.restore_state()
    assert_bit_equal(o.state()["x"], r.state()["x"], "restored x")
    for _ in range(5):
        # reference: forces added by the pre-step callback (RigidBodySystem.cpp:70-73); oracle: after its prepare stage
        r.add_force(3, [1, 2, 3], [0.1, 0.2, 0.3]); r.add_force_at_pos(5, [0.5, 2, 0.5], [0, 30, 0]); r.step(DT)
        o.stage_prepare(); o.add_force(3, [1, 2, 3], [0.1, 0.2, 0.3]); o.add_force_at_pos(5, [0.5, 2, 0.5], [0, 30, 0])
        o.stage_detect(); o.stage_jacobians(); o.stage_solve(DT); o.stage_scatter(DT); o.stage_integrate(DT)
    assert_same_system(o, r, "forces")


def test_spherical_joints_and_mixed_bodies():
    def build(cls):
        s = cls()
        a = s.add_body(1.0, 1, [1, 1, 1], fixed=True, x=[0, 5, 0])
        b = s.add_body(2.0, 0, [0.5], x=[1.5, 5, 0], w=[0, 1, 0])
        c = s.add_body(1.0, 1, [1, 0.5, 1], x=[3, 5, 0], v=[0, 0, 2])
        d = s.add_body(1.5, 3, [0.6, 0.4], x=[4.5, 5, 0], q=[0, 0, 0.6, 0.8])
        s.add_spherical(a, b, [0.75, 0, 0], [-0.75, 0, 0]); s.add_spherical(b, c, [0.75, 0, 0], [-0.75, 0, 0])
        s.add_hinge(c, d, [0.75, 0, 0], [0, 0, 0, 1], [-0.75, 0, 0], [0, 0, 0, 1])
        s.add_body(1.0, 2, [0, 0, 0, 0, 1, 0], fixed=True)
        return s
    o, r = build(Oracle), build(Reference)
    for i in range(150):
        o.step(DT); r.step(DT)
        if i % 15 == 0:
            assert_same_system(o, r, f"mixed step {i + 1}")


def test_random_sphere_box_piles():
    """Seeded random piles: overlapping spheres, tilted fixed boxes (incl. a sphere centre inside a box:
    dist == 0 -> NaN normal contact is emitted by both, CollisionDetect.cpp:144)."""
    rng = np.random.default_rng(42)
    for trial in range(4):
        bodies = []
        for i in range(40):
            bodies.append((1.0 + rng.random(), 0, [0.3 + 0.4 * rng.random()], False, rng.uniform(-2, 2, 3) + [0, 3, 0], None))
        for i in range(3):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            bodies.append((1.0, 1, list(2 + 3 * rng.random(3)), True, rng.uniform(-2, 2, 3), q))
        o, r = Oracle(), Reference()
        for s in (o, r):
            for m, t, g, fx, x, q in bodies:
                s.add_body(m, t, g, fixed=fx, x=x, q=q)
        for i in range(12):
            o.step(DT); r.step(DT)
            so, sr = o.state(), r.state()
            assert o.n_contacts == r.n_contacts
            co, cr = o.contacts(), r.contacts()
            np.testing.assert_array_equal(co["body0"], cr["body0"]); np.testing.assert_array_equal(co["body1"], cr["body1"])
            for k in ("p", "n", "phi"):
                assert_bit_equal(co[k], cr[k], f"pile {trial} step {i} {k}")
            for k in so:
                assert_bit_equal(so[k], sr[k], f"pile {trial} step {i} {k}")
