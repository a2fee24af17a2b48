"""GPU parity: the sm_100a kernels (through the C ABI) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): contact sets bit-exact (pair indices, counts, order) -- here also
p, n, pene bit-exact; Jacobian blocks, J M^-1, diagonal blocks, RHS, world inertia and the integrator
bit-exact (same fp32 operation order, no FMA contraction on either side); PGS impulses and multi-step
trajectories within the fp32 tolerances stated in each test (the sweep keeps the reference's row
order but carries per-body accumulators instead of re-summing neighbours, so rounding differs).
"""
import numpy as np
import pytest

from oracle_lib import Oracle, SPHERE, BOX, PLANE, CYLINDER
from gpu_helpers import context_from_oracles, push_state, assert_bit_equal

pytestmark = pytest.mark.gpu
DT = 0.01

SCENE_STEPS = [("marble_box", {}, [0, 1, 40, 120]), ("sphere_on_box", {}, [0, 82, 83, 90, 150]),
               ("cylinder_on_plane", {}, [0, 35, 36, 60, 200]), ("car", {}, [0, 1, 50, 200]),
               ("swinging_boxes", {}, [0, 10, 100]), ("sphere_stack", {"k": 8}, [0, 5, 100])]


def harvest(scene, kw, steps):
    """Oracle systems advanced to the requested step counts (states harvested from its trajectory)."""
    out, o, done = [], Oracle(scene, **kw), 0
    for s in steps:
        o.step(DT, s - done); done = s
        c = Oracle(scene, **kw)
        c.set_state(**o.state())
        if o.n_joints:
            c.set_joint_lambda(o.joint_rows()["lam"])
        out.append(c)
    return out


@pytest.mark.parametrize("scene,kw,steps", SCENE_STEPS)
def test_stages_bit_exact(scene, kw, steps):
    oracles = harvest(scene, kw, steps)
    ctx = context_from_oracles(oracles)
    push_state(ctx, oracles)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.stage_jacobians(DT); ctx.sync()
    counts = ctx.contact_counts()
    for s, o in enumerate(oracles):
        o.stage_prepare(); o.stage_detect(); o.stage_jacobians()
        # S1/S2: forces and world inertia
        gd, od = ctx.body_dyn(s), o.body_dyn()
        dyn = o.body_params()["fixed"] == 0
        assert_bit_equal(gd["f"], od["f"], "f"); assert_bit_equal(gd["tau"], od["tau"], "tau")
        assert_bit_equal(gd["Iinv"], od["Iinv"], "Iinv"); assert_bit_equal(gd["I"][dyn], od["I"][dyn], "I")
        # S5-S9: contact set, bit-exact including order
        assert counts[s] == o.n_contacts, (scene, steps[s], counts[s], o.n_contacts)
        gc, oc = ctx.contacts(s), o.contacts()
        if o.n_contacts:
            np.testing.assert_array_equal(gc["body0"], oc["body0"]); np.testing.assert_array_equal(gc["body1"], oc["body1"])
            for k in ("p", "n", "phi", "t", "b"):
                assert_bit_equal(gc[k], oc[k], f"contact {k} scene {scene} step {steps[s]}")
            # S10/S11/S14/S15: rows, diagonal blocks, RHS
            gr, orow, ob = ctx.contact_rows(s), o.contact_rows(), o.solver_blocks(DT)
            for k in ("J0", "J1", "J0Minv", "J1Minv"):
                assert_bit_equal(gr[k], orow[k], f"contact {k}")
            assert_bit_equal(gr["A"], ob["A"], "contact A"); assert_bit_equal(gr["rhs"], ob["rhs"], "contact rhs")
        if o.n_joints:
            gj, oj, ob = ctx.joint_rows(s), o.joint_rows(), o.solver_blocks(DT)
            for k in ("J0", "J1", "J0Minv", "J1Minv"):
                assert_bit_equal(gj[k], oj[k], f"joint {k}")
            assert_bit_equal(gj["rhs"], ob["rhsj"], "joint rhs")
    ctx.close()


@pytest.mark.parametrize("scene,kw,steps", SCENE_STEPS)
def test_solve_and_integrate_one_step(scene, kw, steps):
    """One full step from identical states: impulses within 1e-4 relative to the largest impulse
    (+1e-6 abs), state after the step within 2e-6 abs (x, q) / 2e-4 abs (velocities)."""
    oracles = harvest(scene, kw, steps)
    ctx = context_from_oracles(oracles)
    push_state(ctx, oracles)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.stage_jacobians(DT); ctx.stage_solve(DT); ctx.sync()
    lamj = ctx.joint_lambda()
    for s, o in enumerate(oracles):
        o.stage_prepare(); o.stage_detect(); o.stage_jacobians(); o.stage_solve(DT); o.stage_scatter(DT)
        if o.n_contacts:
            gl, ol = ctx.contacts(s)["lam"], o.contacts()["lam"]
            scale = max(np.abs(ol).max(), 1e-3)
            np.testing.assert_allclose(gl, ol, rtol=0, atol=1e-4 * scale + 1e-6, err_msg=f"contact lambda {scene}@{steps[s]}")
        if o.n_joints:
            ol = o.joint_rows()["lam"]
            scale = max(np.abs(ol).max(), 1e-3)
            np.testing.assert_allclose(lamj[s], ol, rtol=0, atol=2e-4 * scale + 1e-6, err_msg=f"joint lambda {scene}@{steps[s]}")
        gd, od = ctx.body_dyn(s), o.body_dyn()
        fscale = max(np.abs(od["fc"]).max(), 1.0)
        np.testing.assert_allclose(gd["fc"], od["fc"], rtol=0, atol=3e-4 * fscale)
        np.testing.assert_allclose(gd["tauc"], od["tauc"], rtol=0, atol=3e-4 * fscale)
    ctx.stage_integrate(DT); ctx.sync()
    gs = ctx.download_state()
    for s, o in enumerate(oracles):
        o.stage_integrate(DT)
        os_ = o.state()
        np.testing.assert_allclose(gs["x"][s], os_["x"], rtol=0, atol=2e-6 * max(1.0, np.abs(os_["x"]).max()))
        np.testing.assert_allclose(gs["q"][s], os_["q"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(gs["v"][s], os_["v"], rtol=0, atol=2e-4 * max(1.0, np.abs(os_["v"]).max()))
        np.testing.assert_allclose(gs["w"][s], os_["w"], rtol=0, atol=2e-4 * max(1.0, np.abs(os_["w"]).max()))
    ctx.close()


def test_integrator_bit_exact_given_forces():
    """No constraints: the whole step is prepare + integrate and must match to the bit."""
    rng = np.random.default_rng(7)
    o = Oracle()
    for i in range(40):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        g = [[0.3 + rng.random()], list(0.2 + rng.random(3)), None, [0.5 + rng.random(), 0.2 + rng.random()]][i % 4]
        t = [SPHERE, BOX, None, CYLINDER][i % 4]
        if t is None:
            t, g = SPHERE, [0.25]
        o.add_body(0.5 + rng.random() * 3, t, g, x=rng.normal(size=3) * 50 + [0, 500 * i, 0], q=q, v=rng.normal(size=3), w=rng.normal(size=3) * 4)
    o.set_params(collisions=False)
    ctx = context_from_oracles([o], max_contacts=8)
    ctx.set_params(collisions=False)
    for _ in range(25):
        ctx.step(DT); o.step(DT)
    ctx.sync()
    g, s = ctx.download_state(), o.state()
    for k in "xqvw":
        assert_bit_equal(g[k][0], s[k], k)
    ctx.close()


TRAJ = [("sphere_on_box", {}, 200, 2e-3), ("car", {}, 150, 2e-3), ("swinging_boxes", {}, 120, 5e-3),
        ("cylinder_on_plane", {}, 120, 2e-3), ("sphere_stack", {"k": 8}, 150, 2e-3), ("marble_box", {}, 30, 2e-3)]


@pytest.mark.parametrize("scene,kw,n,tol", TRAJ)
def test_trajectory_within_tolerance(scene, kw, n, tol):
    """N steps from the scene's initial state, every step on the device (slrbs_step), compared with the
    oracle: positions within tol (abs, metres), contact counts identical along the way."""
    o = Oracle(scene, **kw)
    ctx = context_from_oracles([o])
    for k in range(n):
        ctx.step(DT); o.step(DT)
        if k % 10 == 9 or k == n - 1:
            ctx.sync()
            assert ctx.contact_counts()[0] == o.n_contacts, (scene, k)
    g, s = ctx.download_state(), o.state()
    assert np.isfinite(g["x"]).all()
    np.testing.assert_allclose(g["x"][0], s["x"], rtol=0, atol=tol)
    np.testing.assert_allclose(g["q"][0], s["q"], rtol=0, atol=tol)
    np.testing.assert_allclose(g["v"][0], s["v"], rtol=0, atol=20 * tol)
    ctx.close()


def test_batched_scenes_are_independent():
    """64 car scenes with different chassis speeds in one context == 64 separate oracle runs."""
    B = 64
    oracles = []
    for e in range(B):
        o = Oracle("car")
        st = o.state(); st["v"][0, 2] = -10.0 + 0.05 * e
        o.set_state(**st)
        oracles.append(o)
    ctx = context_from_oracles(oracles, max_contacts=32)
    ctx.step(DT, 40); ctx.sync()
    g = ctx.download_state()
    cnt = ctx.contact_counts()
    for e, o in enumerate(oracles):
        o.step(DT, 40)
        assert cnt[e] == o.n_contacts
        np.testing.assert_allclose(g["x"][e], o.state()["x"], rtol=0, atol=1e-3)
    assert np.abs(g["x"][0] - g["x"][B - 1]).max() > 0.1      # scenes really differ
    ctx.close()


def test_edge_cases_contacts():
    # unsupported pairs ignored, NaN normal emitted for a sphere centre inside a box, fixed-fixed skipped
    o = Oracle()
    o.add_body(1.0, SPHERE, [0.5], x=[0, 0.2, 0])
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    o.add_body(1.0, BOX, [1, 1, 1], x=[0, 0.2, 0])
    o.add_body(1.0, CYLINDER, [1, 1], x=[0.1, 5.0, 0])
    o.add_body(1.0, SPHERE, [0.5], x=[0.3, 0.2, 0], fixed=True)
    o.add_body(1.0, SPHERE, [0.5], x=[0.3, 0.2, 0.1], fixed=True)
    ctx = context_from_oracles([o], max_contacts=16)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.sync()
    o.stage_prepare(); o.stage_detect()
    gc, oc = ctx.contacts(0), o.contacts()
    assert ctx.contact_counts()[0] == o.n_contacts
    np.testing.assert_array_equal(gc["body0"], oc["body0"]); np.testing.assert_array_equal(gc["body1"], oc["body1"])
    assert_bit_equal(gc["n"], oc["n"], "n (with NaN)")
    assert np.isnan(gc["n"]).any()
    ctx.close()


def test_cylinder_cap_branch_and_empty_scene():
    o = Oracle()
    o.add_body(1.0, CYLINDER, [1.0, 0.5], x=[0, 0.5, 0])
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    e = Oracle()
    e.add_body(1.0, SPHERE, [0.5], x=[0, 9, 0])
    e.add_body(1.0, SPHERE, [0.5], x=[5, 9, 0])
    ctx = context_from_oracles([o, e], max_contacts=8)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.stage_jacobians(DT); ctx.sync()
    o.stage_prepare(); o.stage_detect()
    assert list(ctx.contact_counts()) == [4, 0]
    assert_bit_equal(ctx.contacts(0)["p"], o.contacts()["p"], "cap p")
    ctx.step(DT, 3); ctx.sync()
    ctx.close()


def test_contact_capacity_overflow_is_reported():
    from slrbs_b200 import SlrbsError
    o = Oracle("marble_box")
    ctx = context_from_oracles([o], max_contacts=100)
    ctx.stage_prepare(); ctx.stage_detect()
    with pytest.raises(SlrbsError):
        ctx.sync()
    assert ctx.contact_counts()[0] == 100
    ctx.close()


def test_collisions_disabled_and_solver_iters():
    o = Oracle("car"); o.set_params(collisions=False, solver_iter=4)
    ctx = context_from_oracles([o]); ctx.set_params(collisions=False, solver_iters=4)
    ctx.step(DT, 20); o.step(DT, 20); ctx.sync()
    assert ctx.contact_counts()[0] == 0
    np.testing.assert_allclose(ctx.download_state()["x"][0], o.state()["x"], rtol=0, atol=1e-3)
    ctx.close()


def test_external_forces_match_prestep_callback():
    o = Oracle("sphere_on_box")
    ctx = context_from_oracles([o])
    f = np.zeros((2, 3), np.float32); f[0] = [3.0, 0.0, -1.0]
    tau = np.zeros((2, 3), np.float32); tau[0] = [0.0, 0.2, 0.0]
    ctx.set_external_forces(f, tau)
    for _ in range(30):
        ctx.step(DT)
        o.stage_prepare(); o.add_force(0, f[0], tau[0]); o.stage_detect(); o.stage_jacobians()
        o.stage_solve(DT); o.stage_scatter(DT); o.stage_integrate(DT)
    ctx.sync()
    g = ctx.download_state()
    for k in "xqvw":
        assert_bit_equal(g[k][0], o.state()[k], k)       # no contact yet: bit-exact
    ctx.close()


def test_save_restore_scene_range():
    oracles = [Oracle("car") for _ in range(4)]
    ctx = context_from_oracles(oracles, max_contacts=32)
    ctx.save_state()
    s0 = ctx.download_state()
    ctx.step(DT, 10)
    ctx.restore_state(1, 2)
    ctx.sync()
    s1 = ctx.download_state()
    for k in "xqvw":
        assert_bit_equal(s1[k][1:3], s0[k][1:3], k)
    assert np.abs(s1["x"][0] - s0["x"][0]).max() > 1e-3 and np.abs(s1["x"][3] - s0["x"][3]).max() > 1e-3
    ctx.close()


@pytest.mark.parametrize("solver_id", [1, 2])
@pytest.mark.parametrize("scene,steps", [("swinging_boxes", [0, 5, 40]), ("car", [0, 3, 30])])
def test_krylov_solvers_match_oracle(solver_id, scene, steps):
    """solverId 1 (CG, with the reference's p = r - beta p) and 2 (CR) on the joint-only system: one
    step from identical states; lambda within 2e-3 of the largest impulse, state within 1e-4."""
    o0 = Oracle(scene); o0.set_params(solver_id=solver_id, solver_iter=20)
    oracles, done = [], 0
    for s in steps:
        o0.step(DT, s - done); done = s
        c = Oracle(scene); c.set_params(solver_id=solver_id, solver_iter=20)
        c.set_state(**o0.state()); c.set_joint_lambda(o0.joint_rows()["lam"])
        oracles.append(c)
    ctx = context_from_oracles(oracles, max_contacts=32)
    ctx.set_params(solver_id=solver_id, solver_iters=20)
    push_state(ctx, oracles)
    ctx.step(DT, 1); ctx.sync()
    lam, st = ctx.joint_lambda(), ctx.download_state()
    for k, o in enumerate(oracles):
        o.step(DT, 1)
        ol = o.joint_rows()["lam"]
        np.testing.assert_allclose(lam[k], ol, rtol=0, atol=2e-3 * max(np.abs(ol).max(), 1e-3), err_msg=f"{scene}@{steps[k]}")
        np.testing.assert_allclose(st["x"][k], o.state()["x"], rtol=0, atol=1e-4)
        assert (ctx.contacts(k)["lam"] == 0).all() if ctx.contact_counts()[k] else True     # contacts are ignored
    ctx.close()


def test_krylov_trajectory_swinging_boxes():
    """The reference's CG (p = r - beta p, SolverConjGradient.cpp:150) and CR iterations amplify last-bit
    differences by orders of magnitude per step on this scene (mass ratio 100): a 1-ulp change of one input
    moves the ORACLE's own result by ~1e-2 m within a few steps. The tolerance is therefore a small multiple
    (20x) of the oracle's measured 1-ulp self-sensitivity over the same horizon, not an absolute number."""
    for solver_id, n in ((1, 6), (2, 40)):
        o = Oracle("swinging_boxes"); o.set_params(solver_id=solver_id, solver_iter=30)
        p = Oracle("swinging_boxes"); p.set_params(solver_id=solver_id, solver_iter=30)
        st = p.state(); st["v"][19, 2] = np.nextafter(np.float32(10.0), np.float32(11.0)); p.set_state(**st)
        ctx = context_from_oracles([o], max_contacts=8); ctx.set_params(solver_id=solver_id, solver_iters=30)
        ctx.step(DT, n); o.step(DT, n); p.step(DT, n); ctx.sync()
        self_sens = np.abs(p.state()["x"] - o.state()["x"]).max()
        d = np.abs(ctx.download_state()["x"][0] - o.state()["x"]).max()
        assert np.isfinite(d) and d < 20 * max(self_sens, 1e-4), (solver_id, n, d, self_sens)
        ctx.close()


def test_spherical_joints_match_oracle():
    """Spherical (ball) joints have no scene in the reference but are public API (Spherical.cpp:34-52): a chain
    of spheres hanging from a fixed box, rows bit-exact, trajectory within tolerance."""
    def build():
        o = Oracle()
        o.add_body(1.0, BOX, [1, 1, 1], fixed=True, x=[0, 10, 0])
        prev = 0
        for k in range(6):
            b = o.add_body(1.0 + 0.5 * k, SPHERE, [0.3], x=[0.8 * (k + 1), 10, 0], v=[0, 0, 0.5 * k])
            o.add_spherical(prev, b, [0.4, 0, 0], [-0.4, 0, 0])
            prev = b
        o.set_params(collisions=False)
        return o
    o = build()
    ctx = context_from_oracles([o], max_contacts=8); ctx.set_params(collisions=False)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.stage_jacobians(DT); ctx.sync()
    o.stage_prepare(); o.stage_detect(); o.stage_jacobians()
    gj, oj = ctx.joint_rows(0), o.joint_rows()
    for k in ("J0", "J1", "J0Minv", "J1Minv"):
        assert_bit_equal(gj[k], oj[k], k)
    assert_bit_equal(gj["rhs"], o.solver_blocks(DT)["rhsj"], "rhs")
    o2 = build()
    ctx2 = context_from_oracles([o2], max_contacts=8); ctx2.set_params(collisions=False)
    ctx2.step(DT, 100); o2.step(DT, 100); ctx2.sync()
    np.testing.assert_allclose(ctx2.download_state()["x"][0], o2.state()["x"], rtol=0, atol=3e-3)
    assert np.abs(o2.joint_rows()["phi"][:, :3]).max() < 0.05
    ctx.close(); ctx2.close()


def test_ragged_batch_with_per_scene_counts():
    """Scenes of different sizes in one context (counts argument): each must equal its own oracle run, and
    unused body / joint slots must be ignored."""
    from slrbs_b200 import Context
    from gpu_helpers import scene_arrays
    specs = [("car", {}), ("sphere_on_box", {}), ("sphere_stack", {"k": 4}), ("swinging_boxes", {}), ("cylinder_on_plane", {})]
    oracles = [Oracle(s, **kw) for s, kw in specs]
    NBm = max(o.n_bodies for o in oracles); NJm = max(o.n_joints for o in oracles)
    B = len(oracles)
    ctx = Context(B, NBm, NJm, 64)
    def pad(a, n, fill=0):
        out = np.full((n,) + a.shape[1:], fill, a.dtype); out[: a.shape[0]] = a; return out
    arrs = [scene_arrays(o) for o in oracles]
    st = lambda k, n: np.stack([pad(a[k], n) for a in arrs])
    q = st("q", NBm); q[..., 3][np.abs(q).sum(-1) == 0] = 1.0          # padding slots: identity quaternion
    mass = st("mass", NBm); mass[mass == 0] = 1.0
    ctx.upload_bodies(st("x", NBm), q, st("v", NBm), st("w", NBm), mass, st("ibody", NBm), st("ibodyinv", NBm),
                      st("fixed", NBm), st("geom_type", NBm), st("geom", NBm), counts=[o.n_bodies for o in oracles])
    jt = st("jtype", NJm); jt[jt == 0] = 2
    q0 = st("q0", NJm); q0[..., 3][np.abs(q0).sum(-1) == 0] = 1.0
    q1 = st("q1", NJm); q1[..., 3][np.abs(q1).sum(-1) == 0] = 1.0
    ctx.upload_joints(jt, st("jb0", NJm), st("jb1", NJm), st("r0", NJm), q0, st("r1", NJm), q1, counts=[o.n_joints for o in oracles])
    ctx.step(DT, 60); ctx.sync()
    g = ctx.download_state(); cnt = ctx.contact_counts()
    for s, o in enumerate(oracles):
        o.step(DT, 60)
        assert cnt[s] == o.n_contacts, specs[s]
        np.testing.assert_allclose(g["x"][s, : o.n_bodies], o.state()["x"], rtol=0, atol=3e-3, err_msg=str(specs[s]))
    ctx.close()


def test_degenerate_scenes_empty_single_body_and_no_contact_capacity():
    """Ragged extremes in one batch: a scene with no live bodies, a scene with one body, and a full scene; then a
    context created with max_contacts = 0 (collision stage skipped like setEnableCollisionDetection(false))."""
    full = Oracle("sphere_on_box")
    one = Oracle(); one.add_body(2.0, SPHERE, [0.3], x=[0, 5, 0], v=[1, 0, 0], w=[0, 2, 0])
    from slrbs_b200 import Context
    from gpu_helpers import scene_arrays
    a, b = scene_arrays(full), scene_arrays(one)
    nb = 2
    def pad(arr, k):
        out = np.zeros((nb,) + arr.shape[1:], arr.dtype); out[:arr.shape[0]] = arr; return out
    keys = ("x", "q", "v", "w", "mass", "ibody", "ibodyinv", "fixed", "geom_type", "geom")
    stacked = {k: np.stack([pad(a[k], 2), pad(b[k], 1), pad(b[k], 1)]) for k in keys}
    stacked["mass"][stacked["mass"] == 0] = 1.0
    ctx = Context(3, nb, 0, 8)
    ctx.upload_bodies(*[stacked[k] for k in keys], counts=np.array([2, 1, 0], np.int32))
    for _ in range(100):
        ctx.step(DT); full.step(DT); one.step(DT)
    ctx.sync()
    g = ctx.download_state()
    for k in "xqvw":
        assert_bit_equal(g[k][1, :1], one.state()[k], k)                 # free flight: bit-exact
    np.testing.assert_allclose(g["x"][0], full.state()["x"], rtol=0, atol=1e-4)
    assert list(ctx.contact_counts()) == [full.n_contacts, 0, 0]
    ctx.close()
    nocap = Context(1, 2, 0, 0)
    nocap.upload_bodies(*[stacked[k][0] for k in keys])
    nocap.step(DT, 5); nocap.sync()
    assert np.isfinite(nocap.download_state()["x"]).all()
    nocap.close()
