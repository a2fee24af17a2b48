"""Pins the CPU oracle against the known answers of SURVEY.md Appendix B (derived from the
reference's constants; the reference ships no tests or golden vectors of its own) and against
analytic invariants.  CPU only."""
import numpy as np
import pytest

from oracle_lib import Oracle, SPHERE, BOX, CYLINDER, PLANE

DT = 0.01


def test_B1_B2_B3_inertia():
    o = Oracle()
    o.add_body(1.0, SPHERE, [0.5])
    o.add_body(1.0, BOX, [10.0, 0.4, 10.0])
    o.add_body(1.0, CYLINDER, [0.2, 0.5])
    ib = o.body_params()["ibody"].reshape(-1, 3, 3)
    np.testing.assert_allclose(np.diag(ib[0]), [0.1] * 3, rtol=1e-6)
    np.testing.assert_allclose(np.diag(ib[1]), [100.16 / 12, 200.0 / 12, 100.16 / 12], rtol=1e-6)
    np.testing.assert_allclose(np.diag(ib[2]), [(0.75 + 0.04) / 12, 0.125, (0.75 + 0.04) / 12], rtol=1e-6)
    ibi = o.body_params()["ibodyinv"].reshape(-1, 3, 3)
    np.testing.assert_allclose(np.diag(ibi[0]), [10.0] * 3, rtol=1e-5)


def test_B4_marble_box_first_step():
    o = Oracle("marble_box")
    assert o.n_bodies == 167
    bp = o.body_params()
    assert bp["fixed"].sum() == 5 and (bp["geom_type"][:162] == SPHERE).all() and (bp["geom_type"][162:] == BOX).all()
    x = o.state()["x"]
    # index 2*(9i+j) lower, +1 upper
    assert x[0, 1] == 2.0 and x[1, 1] == 3.0 and x[2, 2] == -3.0 and x[18, 0] == -3.0
    o.step(DT)
    assert o.pairs_tested == 13851
    c = o.contacts()
    assert o.n_contacts == 369
    assert (c["body0"] < 162).all() and (c["body1"] < 162).all()          # all sphere-sphere
    assert (c["phi"] == 0.0).all()
    vertical = (c["body1"] - c["body0"] == 1) & (c["body0"] % 2 == 0)
    assert vertical.sum() == 81 and (~vertical).sum() == 288
    # lexicographic (i, j) detection order
    key = c["body0"].astype(np.int64) * 1000 + c["body1"]
    assert (np.diff(key) > 0).all()


def test_B5_sphere_on_box_first_contact():
    o = Oracle("sphere_on_box")
    for k in range(1, 83):
        o.step(DT)
        assert o.n_contacts == 0, k
        assert abs(o.state()["v"][0, 1] - (-0.0981 * k)) < 1e-4
    y_before = o.state()["x"][0, 1]
    o.step(DT)
    assert o.n_contacts == 1
    c = o.contacts()
    assert c["body0"][0] == 0 and c["body1"][0] == 1
    assert abs(y_before - 0.661656) < 2e-4


def test_B6_car_first_step():
    o = Oracle("car")
    assert o.n_bodies == 6 and o.n_joints == 4 and o.n_joint_rows == 20
    o.step(DT)
    c = o.contacts()
    assert o.n_contacts == 8
    assert list(c["body0"]) == [1, 1, 2, 2, 3, 3, 4, 4]
    assert (c["body1"] == 5).all()
    np.testing.assert_array_equal(c["n"], np.tile(np.array([0, 1, 0], np.float32), (8, 1)))
    assert (np.abs(c["phi"]) < 1e-5).all()


def test_B7_cylinder_on_plane_first_contact():
    o = Oracle("cylinder_on_plane")
    first = None
    for k in range(1, 60):
        o.step(DT)
        if o.n_contacts > 0:
            first = k
            break
    assert first == 36
    c = o.contacts()
    assert o.n_contacts == 1
    np.testing.assert_allclose(c["p"][0], [-0.3023, 0.0004, 0.0], atol=2e-3)


def test_B8_swinging_boxes():
    o = Oracle("swinging_boxes")
    assert o.n_bodies == 20 and o.n_joints == 19 and o.n_joint_rows == 95
    bp = o.body_params()
    assert bp["fixed"][0] == 1 and bp["mass"][19] == 100.0
    assert o.state()["v"][19, 2] == 10.0
    for _ in range(50):
        o.step(DT)
        assert o.n_contacts == 0
    # hinge position error stays small (Baumgarte-stabilised PGS)
    assert np.abs(o.joint_rows()["phi"][:, :3]).max() < 0.05
    assert np.isfinite(o.state()["x"]).all()


def test_free_fall_symplectic_euler():
    o = Oracle()
    o.add_body(2.0, SPHERE, [0.5], x=[0, 100, 0])
    y = np.float32(100.0); v = np.float32(0.0)
    for _ in range(100):
        o.step(DT)
        v = np.float32(v + np.float32(np.float32(DT) * np.float32(1.0 / 2.0)) * np.float32(np.float32(2.0) * np.float32(-9.81)))
        y = np.float32(y + np.float32(DT) * v)
    s = o.state()
    assert s["v"][0, 1] == v and s["x"][0, 1] == y


def test_sphere_sphere_contact_normal_and_momentum():
    # two equal spheres approaching head-on along x, gravity acts equally: relative momentum conserved
    o = Oracle()
    o.add_body(1.0, SPHERE, [0.5], x=[-0.49, 0, 0], v=[1, 0, 0])
    o.add_body(1.0, SPHERE, [0.5], x=[0.49, 0, 0], v=[-1, 0, 0])
    o.stage_prepare(); o.stage_detect()
    c = o.contacts()
    assert o.n_contacts == 1
    np.testing.assert_allclose(c["n"][0], [-1, 0, 0])          # from body1 toward body0
    np.testing.assert_allclose(c["phi"][0], -0.02, atol=1e-6)
    np.testing.assert_allclose(c["p"][0], [0, 0, 0], atol=1e-6)
    o.stage_jacobians(); o.stage_solve(DT); o.stage_scatter(DT); o.stage_integrate(DT)
    v = o.state()["v"]
    np.testing.assert_allclose(v[0, 0] + v[1, 0], 0.0, atol=1e-5)
    assert v[0, 0] <= 0.0 + 1e-3 and v[1, 0] >= 0.0 - 1e-3      # no longer approaching


def test_sphere_box_nan_normal_when_centre_inside():
    o = Oracle()
    o.add_body(1.0, SPHERE, [0.5], x=[0, 0, 0])
    o.add_body(1.0, BOX, [2, 2, 2], fixed=True)
    o.stage_prepare(); o.stage_detect()
    assert o.n_contacts == 1 and np.isnan(o.contacts()["n"]).all()   # CollisionDetect.cpp:144 quirk kept


def test_unsupported_pairs_are_ignored():
    o = Oracle()
    o.add_body(1.0, SPHERE, [0.5], x=[0, 0.2, 0])
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    o.add_body(1.0, BOX, [1, 1, 1], x=[0, 0.2, 0])
    o.add_body(1.0, CYLINDER, [1, 1], x=[0.1, 5.0, 0])
    o.stage_prepare(); o.stage_detect()
    c = o.contacts()
    # only sphere-box (0,2) is supported; sphere-plane, box-plane, sphere-cyl, box-cyl are not; cylinder is far
    assert o.n_contacts == 1 and c["body0"][0] == 0 and c["body1"][0] == 2


def test_cylinder_plane_cap_branch_four_points():
    o = Oracle()
    o.add_body(1.0, CYLINDER, [1.0, 0.5], x=[0, 0.5, 0])     # axis = y, aligned with plane normal
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    o.stage_prepare(); o.stage_detect()
    c = o.contacts()
    assert o.n_contacts == 4
    assert (c["body0"] == 0).all() and (c["body1"] == 1).all()
    np.testing.assert_allclose(np.linalg.norm(c["p"][:, [0, 2]], axis=1), 0.5, atol=1e-6)


def test_save_restore_keeps_joint_lambda():
    o = Oracle("swinging_boxes")
    o.save_state()
    s0 = o.state()
    o.step(DT, 5)
    lam = o.joint_rows()["lam"].copy()
    assert np.abs(lam).max() > 0
    o.restore_state()
    s1 = o.state()
    for k in s0:
        np.testing.assert_array_equal(s0[k], s1[k])
    np.testing.assert_array_equal(o.joint_rows()["lam"], lam)   # RigidBodyState.cpp:26-40 does not reset lambda


@pytest.mark.parametrize("solver_id", [1, 2])
def test_krylov_solvers_hold_hinges(solver_id):
    o = Oracle("swinging_boxes")
    o.set_params(solver_id=solver_id, solver_iter=50)
    o.step(DT, 20)
    assert np.isfinite(o.state()["x"]).all()
    assert np.abs(o.joint_rows()["phi"][:, :3]).max() < 0.2
