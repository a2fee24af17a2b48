"""GPU parity against the committed golden vectors of the reference's own code (tests/golden/ref_*.npz).

From each checkpoint state of the reference the device takes one step through the C ABI: the contact list
(count, order, pairs, p, n, pene) must match BIT FOR BIT; impulses within 1e-4 * max|lambda| + 1e-6; the
state after the step within 2e-6 (x, q) / 2e-4 (velocities) relative to max(1, |.|max).  A free-running
device trajectory from the initial state is compared at `total` (150-300) steps with the per-scene tolerance below.
"""
import numpy as np
import pytest

from golden_lib import SCENE_KW, contacts, load, state
from gpu_helpers import assert_bit_equal, context_from_oracles
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

# |x_gpu - x_ref|max after `total` free-running steps. Measured on B200: 0 (bit-identical positions) for the four
# sphere scenes, 3e-10 cylinder, 1.4e-6 car, 7.8e-5 swinging boxes (200 steps of a 19-hinge chain).
TRAJ_TOL = {"marble_box": 1e-5, "marble_box_n": 1e-5, "sphere_on_box": 1e-6, "cylinder_on_plane": 1e-6,
            "car": 5e-5, "swinging_boxes": 2e-3, "sphere_stack": 1e-6}


@pytest.mark.parametrize("scene", sorted(SCENE_KW))
def test_one_step_from_reference_checkpoints(scene):
    g = load(scene)
    dt = float(g["dt"])
    cps = [int(k) for k in g["checkpoints"]]
    builders = [Oracle(scene, **SCENE_KW[scene]) for _ in cps]          # scene tables only (masses, shapes, joints)
    ctx = context_from_oracles(builders, max_contacts=1024)
    pre = [state(g, f"pre_{k}") for k in cps]
    ctx.upload_state(*[np.stack([p[key] for p in pre]) for key in ("x", "q", "v", "w")])
    if builders[0].n_joints:
        ctx.set_joint_lambda(np.stack([g[f"pre_{k}_jlam"] for k in cps]))
    ctx.step(dt, 1); ctx.sync()
    counts, gs = ctx.contact_counts(), ctx.download_state()
    for s, k in enumerate(cps):
        ref = contacts(g, k)
        assert counts[s] == ref["body0"].size, (scene, k, counts[s], ref["body0"].size)
        if counts[s]:
            c = ctx.contacts(s)
            np.testing.assert_array_equal(c["body0"], ref["body0"]); np.testing.assert_array_equal(c["body1"], ref["body1"])
            for key in ("p", "n", "phi"):
                assert_bit_equal(c[key], ref[key], f"{scene} step {k + 1} contact {key}")
            scale = max(float(np.abs(ref["lam"]).max()), 1e-3)
            np.testing.assert_allclose(c["lam"], ref["lam"], rtol=0, atol=1e-4 * scale + 1e-6)
        post = state(g, f"post_{k}")
        for key, tol in (("x", 2e-6), ("q", 2e-6), ("v", 2e-4), ("w", 2e-4)):
            np.testing.assert_allclose(gs[key][s], post[key], rtol=0, atol=tol * max(1.0, float(np.abs(post[key]).max())),
                                       err_msg=f"{scene} post {k} {key}")
    ctx.close()


@pytest.mark.parametrize("scene", sorted(SCENE_KW))
def test_free_running_trajectory(scene):
    g = load(scene)
    o = Oracle(scene, **SCENE_KW[scene])
    ctx = context_from_oracles([o], max_contacts=1024)
    ctx.step(float(g["dt"]), int(g["total"])); ctx.sync()
    gs, fin = ctx.download_state(), state(g, "final")
    assert np.abs(gs["x"][0] - fin["x"]).max() < TRAJ_TOL[scene], (scene, np.abs(gs["x"][0] - fin["x"]).max())
    ctx.close()


@pytest.mark.parametrize("name", ["car_cg", "car_cr", "swinging_boxes_cg", "swinging_boxes_cr"])
def test_krylov_one_step_from_reference_checkpoints(name):
    """solverId 1 / 2 on the device against the reference's own CG / CR: one step from each checkpoint state; joint
    impulses within 2e-3 of the largest impulse, positions within 1e-4 (the reference's Krylov iterations amplify
    last-bit differences, so longer horizons are compared in tests/test_gpu_parity.py against the oracle's own
    1-ulp sensitivity)."""
    g = load(name)
    scene = name.rsplit("_", 1)[0]
    cps = [int(k) for k in g["checkpoints"]]
    builders = [Oracle(scene) for _ in cps]
    ctx = context_from_oracles(builders, max_contacts=32)
    ctx.set_params(0.8, int(g["solver_id"]), int(g["solver_iter"]), True)
    ctx.upload_state(*[np.stack([g[f"pre_{k}_{key}"] for k in cps]) for key in "xqvw"])
    ctx.set_joint_lambda(np.stack([g[f"pre_{k}_jlam"] for k in cps]))
    ctx.step(float(g["dt"]), 1); ctx.sync()
    lam, st = ctx.joint_lambda(), ctx.download_state()
    for s, k in enumerate(cps):
        ref = g[f"post_{k}_jlam"]
        np.testing.assert_allclose(lam[s], ref, rtol=0, atol=2e-3 * max(float(np.abs(ref).max()), 1e-3), err_msg=f"{name}@{k}")
        np.testing.assert_allclose(st["x"][s], g[f"post_{k}_x"], rtol=0, atol=1e-4)
    ctx.close()


def test_spherical_chain_trajectory_against_reference():
    from test_golden_cpu import spherical_chain
    g = load("spherical_chain")
    o = spherical_chain()
    ctx = context_from_oracles([o], max_contacts=8); ctx.set_params(0.8, 0, 10, False)
    done = 0
    for k in (1, 50, 150):
        ctx.step(float(g["dt"]), k - done); done = k
        ctx.sync()
        st = ctx.download_state()
        np.testing.assert_allclose(st["x"][0], g[f"at_{k}_x"], rtol=0, atol=2e-3, err_msg=f"step {k}")
    ctx.close()
