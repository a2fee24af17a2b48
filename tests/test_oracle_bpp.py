"""Block principal pivoting (solver_id 3) in the oracle. EXTENSION: the reference has no such solver
(SURVEY.md 8(f)-3), so nothing here is pinned to reference output; the checks are the box-LCP residual of
the problem the solver poses and agreement with a long Box-PGS run (SolverBoxPGS.cpp:217-248, 500 sweeps)."""
import numpy as np
import pytest

from oracle_lib import Oracle

DT = 0.01
SCENES = [("car", {}, False), ("sphere_stack", {"k": 8}, False), ("box_stack", {"k": 5}, True),
          ("swinging_boxes", {}, False), ("cylinder_on_plane", {}, False), ("sphere_on_box", {}, False)]


def _run(name, kw, ext, solver_id, iters, steps):
    o = Oracle(name, **kw)
    o.set_params(solver_id=solver_id, solver_iter=iters)
    if ext:
        o.set_extensions(1)
    worst = 0.0
    for _ in range(steps):
        o.stage_prepare(); o.stage_detect(); o.stage_jacobians(); o.stage_solve(DT)
        worst = max(worst, o.blcp_residual(DT))
        o.stage_scatter(DT); o.stage_integrate(DT)
    return o, worst


@pytest.mark.parametrize("name,kw,ext", SCENES, ids=[s[0] for s in SCENES])
def test_bpp_solves_the_box_lcp(name, kw, ext):
    o, worst = _run(name, kw, ext, 3, 10, 120)
    assert worst < 1e-4, worst
    st = o.state()
    assert all(np.isfinite(st[k]).all() for k in st)


@pytest.mark.parametrize("name,kw,ext", SCENES, ids=[s[0] for s in SCENES])
def test_bpp_agrees_with_long_pgs(name, kw, ext):
    a, _ = _run(name, kw, ext, 3, 10, 120)
    b, _ = _run(name, kw, ext, 0, 500, 120)
    assert a.n_contacts == b.n_contacts
    # the 100:1 mass-ratio hinge chain is where 500 PGS sweeps are still 3e-2 from the solution
    np.testing.assert_allclose(a.state()["x"], b.state()["x"], atol=0.15 if name == "swinging_boxes" else 5e-3)


def test_bpp_marble_box_short():
    o, worst = _run("marble_box", {}, False, 3, 10, 80)
    assert worst < 1e-4, worst
    # marbles rest on the floor and on each other: lower layer at y = 0.7, upper at 1.7 (Appendix B geometry)
    y = o.state()["x"][:162, 1]
    assert abs(y.min() - 0.7) < 2e-3 and y.max() < 2.01


def test_bpp_zero_iterations_leaves_impulses_alone():
    o = Oracle("car")
    o.set_params(solver_id=3, solver_iter=0)
    o.stage_prepare(); o.stage_detect(); o.stage_jacobians(); o.stage_solve(DT)
    assert np.all(o.contacts()["lam"] == 0.0)
