"""GPU: block principal pivoting (solver_id 3, csrc/bpp.cu) against the restatement's solve_bpp. EXTENSION -- the
reference has no such solver, so nothing is pinned to reference output. The kernel follows the
restatement operation for operation; the assembled rows it reads differ from the restatement's only where the row
builders use fused multiply-adds, so impulses agree to fp32 round-off, and both must solve the box LCP they pose."""
import os

import numpy as np
import pytest

from oracle_lib import Oracle
from gpu_helpers import context_from_oracles, push_state
from slrbs_b200 import SlrbsError

pytestmark = pytest.mark.gpu
DT = 0.01
SCENES = [("car", {}, False, 32), ("sphere_stack", {"k": 8}, False, 32), ("box_stack", {"k": 5}, True, 64),
          ("swinging_boxes", {}, False, 8), ("cylinder_on_plane", {}, False, 8), ("sphere_on_box", {}, False, 8)]


def snapshots(name, kw, ext, at):
    """Systems frozen along a BPP-driven trajectory of the restatement."""
    out, o, done = [], Oracle(name, **kw), 0
    o.set_params(solver_id=3, solver_iter=10)
    if ext:
        o.set_extensions(1)
    for st in at:
        o.step(DT, st - done); done = st
        c = Oracle(name, **kw); c.set_params(solver_id=3, solver_iter=10)
        if ext:
            c.set_extensions(1)
        c.set_state(**o.state())
        if o.n_joints:
            c.set_joint_lambda(o.joint_rows()["lam"])
        out.append(c)
    return out


def make(mode, oracles, **kw):
    os.environ["SLRBS_PGS_MODE"] = mode
    try:
        return context_from_oracles(oracles, **kw)
    finally:
        os.environ.pop("SLRBS_PGS_MODE", None)


@pytest.mark.parametrize("mode", ["thread", "level"])
@pytest.mark.parametrize("name,kw,ext,nc", SCENES, ids=[s[0] for s in SCENES])
def test_bpp_impulses_match_the_restatement(name, kw, ext, nc, mode):
    oracles = snapshots(name, kw, ext, [0, 1, 40, 90, 160])
    ctx = make(mode, oracles, max_contacts=nc)
    ctx.set_params(solver_id=3, solver_iters=10)
    if ext:
        ctx.set_extensions(1)
    push_state(ctx, oracles)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.stage_jacobians(DT); ctx.stage_solve(DT); ctx.sync()
    jl = ctx.joint_lambda()
    for s, o in enumerate(oracles):
        o.stage_prepare(); o.stage_detect(); o.stage_jacobians(); o.stage_solve(DT)
        assert o.blcp_residual(DT) < 1e-4
        assert ctx.contact_counts()[s] == o.n_contacts
        scale = 1.0
        if o.n_contacts:
            ref = o.contacts()["lam"]
            scale = max(scale, float(np.abs(ref).max()))
            np.testing.assert_allclose(ctx.contacts(s)["lam"], ref, atol=2e-4 * scale, rtol=0)
        if o.n_joints:
            ref = o.joint_rows()["lam"]
            np.testing.assert_allclose(jl[s], ref, atol=2e-4 * max(scale, float(np.abs(ref).max())), rtol=0)
    ctx.close()


@pytest.mark.parametrize("name,kw,ext,nc", SCENES, ids=[s[0] for s in SCENES])
def test_bpp_trajectories_follow_the_restatement(name, kw, ext, nc):
    o = Oracle(name, **kw); o.set_params(solver_id=3, solver_iter=10)
    if ext:
        o.set_extensions(1)
    ctx = context_from_oracles([o, o, o], max_contacts=nc)
    ctx.set_params(solver_id=3, solver_iters=10)
    if ext:
        ctx.set_extensions(1)
    ctx.step(DT, 100); ctx.sync()
    o.step(DT, 100)
    g, c = ctx.download_state(), o.state()
    tol = 2e-2 if name == "swinging_boxes" else 2e-3
    for s in range(3):
        np.testing.assert_allclose(g["x"][s], c["x"], atol=tol)
        np.testing.assert_allclose(g["v"][s], c["v"], atol=10 * tol)
    ctx.close()


def test_bpp_solves_the_reference_marble_box_from_global_scratch():
    """The reference's own marble box has 369+ contacts = 1100+ scalar rows, beyond what shared memory holds (288): such scenes
    run the same algorithm with the row records and the Cholesky triangle in a global scratch area (k_bpp<BIG>)."""
    o = Oracle("marble_box"); o.set_params(solver_id=3, solver_iter=10)
    o.step(DT, 3)
    snap = Oracle("marble_box"); snap.set_params(solver_id=3, solver_iter=10); snap.set_state(**o.state())
    ctx = context_from_oracles([snap], max_contacts=1024)
    ctx.set_params(solver_id=3, solver_iters=10)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.stage_jacobians(DT); ctx.stage_solve(DT); ctx.sync()
    snap.stage_prepare(); snap.stage_detect(); snap.stage_jacobians(); snap.stage_solve(DT)
    assert snap.n_contacts == ctx.contact_counts()[0] and 3 * snap.n_contacts > 288
    assert snap.blcp_residual(DT) < 1e-4
    ref = snap.contacts()["lam"]
    np.testing.assert_allclose(ctx.contacts(0)["lam"], ref, atol=5e-4 * max(1.0, float(np.abs(ref).max())), rtol=0)
    # and a short trajectory stays with the restatement's
    ctx.stage_integrate(DT); snap.stage_scatter(DT); snap.stage_integrate(DT)
    ctx.step(DT, 10); ctx.sync(); snap.step(DT, 10)
    np.testing.assert_allclose(ctx.download_state()["x"][0], snap.state()["x"], atol=5e-3)
    ctx.close()


def test_bpp_refuses_a_scene_with_more_rows_than_it_holds():
    o = Oracle("marble_box_n", nx=12, nz=12, ny=16)      # ~7000 contacts = 21 000 rows > 4096
    ctx = context_from_oracles([o], max_contacts=8192)
    ctx.set_params(solver_id=3, solver_iters=10)
    ctx.step(DT, 2)
    with pytest.raises(SlrbsError):
        ctx.sync()
    ctx.set_params(solver_id=0, solver_iters=10)       # reported once: the context is usable again with Box-PGS
    ctx.step(DT, 2); ctx.sync()
    ctx.close()


def test_bpp_through_the_host_mirror():
    from slrbs_b200 import host_api
    sys_ = host_api.HostSystem("car")
    sys_.set_solver(solver_id=3, solver_iter=10)
    sys_.step(DT, 100)
    o = Oracle("car"); o.set_params(solver_id=3, solver_iter=10); o.step(DT, 100)
    np.testing.assert_allclose(sys_.state()["x"], o.state()["x"], atol=2e-3)
