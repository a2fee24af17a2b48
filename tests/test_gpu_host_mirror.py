"""GPU: the C++ mirror of the reference API (RigidBodySystem::step through libslrbs_host.so) against
the oracle: same contact sets, trajectories within tolerance, callbacks, save/restore, solver knobs."""
import numpy as np
import pytest

from oracle_lib import Oracle
from gpu_helpers import assert_bit_equal
from slrbs_b200 import host_api

pytestmark = pytest.mark.gpu
DT = 0.01


@pytest.mark.parametrize("scene,n,tol", [("car", 100, 2e-3), ("sphere_on_box", 120, 2e-3), ("swinging_boxes", 60, 5e-3),
                                         ("cylinder_on_plane", 80, 2e-3), ("marble_box", 20, 2e-3)])
def test_reference_api_step_matches_oracle(scene, n, tol):
    h, o = host_api.HostSystem(scene), Oracle(scene)
    for k in range(n):
        h.step(DT); o.step(DT)
        assert h.n_contacts == o.n_contacts, (scene, k)
    hs, os_ = h.state(), o.state()
    np.testing.assert_allclose(hs["x"], os_["x"], rtol=0, atol=tol)
    np.testing.assert_allclose(hs["q"], os_["q"], rtol=0, atol=tol)
    hc, oc = h.contacts(), o.contacts()
    np.testing.assert_array_equal(hc["body0"], oc["body0"]); np.testing.assert_array_equal(hc["body1"], oc["body1"])
    if o.n_contacts:
        np.testing.assert_allclose(hc["p"], oc["p"], rtol=0, atol=10 * tol)
        scale = max(np.abs(oc["lam"]).max(), 1e-3)
        np.testing.assert_allclose(hc["lam"], oc["lam"], rtol=0, atol=0.05 * scale)
    if o.n_joints:
        ol = o.joint_rows()["lam"]
        np.testing.assert_allclose(h.joint_lambda(), ol, rtol=0, atol=0.02 * max(np.abs(ol).max(), 1e-3))
    h.close()


def test_prestep_callback_forces():
    h, o = host_api.HostSystem("sphere_on_box"), Oracle("sphere_on_box")
    f, tau = np.array([2.0, 0.0, 1.0], np.float32), np.array([0.0, 0.1, 0.0], np.float32)
    h.set_prestep_force(0, f, tau)
    for _ in range(40):
        h.step(DT)
        o.stage_prepare(); o.add_force(0, f, tau); o.stage_detect(); o.stage_jacobians()
        o.stage_solve(DT); o.stage_scatter(DT); o.stage_integrate(DT)
    for k in "xqvw":
        assert_bit_equal(h.state()[k], o.state()[k], k)      # free flight: bit-exact
    h.close()


def test_save_restore_and_direct_field_writes():
    h, o = host_api.HostSystem("car"), Oracle("car")
    h.save(); o.save_state()
    h.step(DT, 15); o.step(DT, 15)
    h.restore(); o.restore_state()
    for k in "xqvw":
        assert_bit_equal(h.state()[k], o.state()[k], k)
    # joint lambda survives the restore on both sides (RigidBodyState.cpp:26-40)
    h.step(DT, 10); o.step(DT, 10)
    np.testing.assert_allclose(h.state()["x"], o.state()["x"], rtol=0, atol=1e-3)
    # poke a public field between steps
    st = h.state(); st["v"][0] = [0.0, 2.0, -3.0]
    h.set_state(**st); o.set_state(**st)
    h.step(DT, 10); o.step(DT, 10)
    np.testing.assert_allclose(h.state()["x"], o.state()["x"], rtol=0, atol=1e-3)
    h.close()


def test_solver_iterations_and_friction_knobs():
    h, o = host_api.HostSystem("sphere_on_box"), Oracle("sphere_on_box")
    h.set_solver(0, 3, 0.2, True); o.set_params(mu=0.2, solver_id=0, solver_iter=3)
    h.step(DT, 150); o.step(DT, 150)
    np.testing.assert_allclose(h.state()["x"], o.state()["x"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(h.state()["w"], o.state()["w"], rtol=0, atol=2e-2)
    h.set_solver(0, 10, 0.8, True)
    h.close()


def test_headless_program_written_against_the_reference_headers():
    """examples/headless.cpp uses only the reference's class surface (RigidBodySystem, Scenarios, public fields);
    built against the mirror headers it must reproduce the oracle's car after 120 steps."""
    import os, subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "slrbs_b200", "slrbs_headless")
    out = subprocess.run([exe, "car", "120"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l.split() for l in out.stdout.splitlines() if l and l[0].isdigit()]
    x = np.array([[float(v) for v in l[1:4]] for l in lines], np.float32)
    o = Oracle("car"); o.step(DT, 120)
    np.testing.assert_allclose(x, o.state()["x"], rtol=0, atol=2e-3)
    assert f"contacts {o.n_contacts}" in out.stdout


def test_batch_matches_single_systems():
    """RigidBodySystemBatch: every scene of a batch evolves exactly as a RigidBodySystem built the same way (bit for bit:
    same kernels, same order); reset by scene index, per-scene forces, readScene."""
    proto = host_api.HostSystem("car")
    batch = host_api.HostBatch(proto, 5)
    single = host_api.HostSystem("car")
    batch.save()
    batch.step(DT, 40); batch.sync()
    single.step(DT, 40)
    st = batch.state()
    for s in range(5):
        for k in "xqvw":
            assert_bit_equal(st[k][s], single.state()[k], f"scene {s} {k}")
    assert (batch.contact_counts() == single.n_contacts).all()
    # reset scenes 1..2 to the saved state: they restart (joint lambda kept, RigidBodyState.cpp:26-40), the others go on
    batch.reset(1, 2)
    batch.step(DT, 10); batch.sync()
    single.step(DT, 10)
    st = batch.state()
    for s in (0, 3, 4):
        assert_bit_equal(st["x"][s], single.state()["x"], f"scene {s} after the partial reset")
    o = Oracle("car"); o.step(DT, 10)
    np.testing.assert_allclose(st["x"][1], o.state()["x"], rtol=0, atol=5e-3)       # warm-started joints: close to a fresh run
    assert_bit_equal(st["x"][1], st["x"][2], "the two reset scenes agree")
    # a force on scene 4 only
    f = np.zeros((1, batch.n_bodies, 3), np.float32); f[0, 0] = [0.0, 30.0, 0.0]
    batch.set_forces(f=f, scene0=4, n_scenes=1)
    batch.step(DT, 5); batch.sync()
    single.step(DT, 5)
    st = batch.state()
    assert_bit_equal(st["x"][0], single.state()["x"], "scene 0 unaffected by scene 4's force")
    assert st["x"][4][0, 1] > st["x"][0][0, 1]
    target = host_api.HostSystem("car")
    batch.read_scene(0, target)
    assert_bit_equal(target.state()["x"], single.state()["x"], "readScene")
    for h in (proto, single, target):
        h.close()
    batch.close()


@pytest.mark.parametrize("scene,steps,tol", [("sphere_on_box", 120, 2e-3), ("car", 80, 5e-3), ("marble_box", 15, 5e-3)])
def test_solver_subclass_without_device_kernel_runs_on_the_host(scene, steps, tol):
    """A Solver subclass that only overrides solve(h) (as one written against the reference's Solver.h would) is run on the
    host: rows mirrored into the public fields, lambdas sent back, force scatter and integration on the device."""
    h, o = host_api.HostSystem(scene), Oracle(scene)
    h.use_host_solver(0)
    for k in range(steps):
        h.step(DT); o.step(DT)
    assert h.n_contacts == o.n_contacts
    np.testing.assert_allclose(h.state()["x"], o.state()["x"], rtol=0, atol=tol)
    np.testing.assert_allclose(h.state()["v"], o.state()["v"], rtol=0, atol=20 * tol)
    h.close()


def test_headless_batch_program():
    """`slrbs_headless batch car 64 120`: RigidBodySystemBatch from C++; the last scene equals the single-system run."""
    import os, subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "slrbs_b200", "slrbs_headless")
    out = subprocess.run([exe, "batch", "car", "64", "120"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l.split() for l in out.stdout.splitlines() if l and l[0].isdigit()]
    x = np.array([[float(v) for v in l[1:4]] for l in lines], np.float32)
    o = Oracle("car"); o.step(DT, 120)
    np.testing.assert_allclose(x, o.state()["x"], rtol=0, atol=2e-3)
