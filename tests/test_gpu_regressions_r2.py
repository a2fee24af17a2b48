"""GPU regressions for the round-1 review findings: state that the level-scheduled kernels walk must not
survive the thing it described -- contact levels after collisions are switched off, joint levels after the joints are
removed -- and the external-force mode is per scene, not per context."""
import os

import numpy as np
import pytest

from oracle_lib import Oracle
from gpu_helpers import context_from_oracles, assert_bit_equal
from slrbs_b200 import SlrbsError

pytestmark = pytest.mark.gpu
DT = 0.01


def make(mode, oracles, **kw):
    os.environ["SLRBS_PGS_MODE"] = mode
    try:
        return context_from_oracles(oracles, **kw)
    finally:
        os.environ.pop("SLRBS_PGS_MODE", None)


@pytest.mark.parametrize("solver_id", [0, 2])
def test_collisions_switched_off_mid_run_leave_no_phantom_contacts(solver_id):
    """RigidBodySystem.cpp:75-81: contacts are cleared every step and only re-detected when collisions are enabled."""
    o = Oracle("marble_box")
    o.step(DT, 40)
    assert o.n_contacts > 300
    a, b = make("thread", [o], max_contacts=1024), make("level", [o], max_contacts=1024)
    for c in (a, b):
        c.step(DT, 3)
    o.step(DT, 3)
    for c in (a, b):
        c.set_params(collisions=False, solver_id=solver_id)
        c.step(DT, 5); c.sync()
    o.set_params(collisions=False, solver_id=solver_id)
    o.step(DT, 5)
    assert a.contact_counts()[0] == b.contact_counts()[0] == o.n_contacts == 0
    sa, sb, so = a.download_state(), b.download_state(), o.state()
    for k in "xqvw":
        assert_bit_equal(sa[k], sb[k], f"thread vs level {k}")
    # after the switch the spheres are in free fall: the last five steps are exact, the three before carry PGS rounding
    np.testing.assert_allclose(sb["v"][0], so["v"], rtol=0, atol=2e-4)
    assert_bit_equal(b.body_dyn(0)["fc"], np.zeros_like(b.body_dyn(0)["fc"]), "constraint forces without contacts")
    a.close(); b.close()


def test_removed_joints_stop_acting_in_level_mode():
    """slrbs_upload_joints(n = 0) is RigidBodySystem::clear's joint half (RigidBodySystem.cpp:133-137) for a scene range."""
    o = [Oracle("swinging_boxes") for _ in range(2)]
    a, b = make("thread", o, max_contacts=8), make("level", o, max_contacts=8)
    for c in (a, b):
        c.step(DT, 20)
        c.upload_joints(np.zeros((2, 0), np.int32), None, None, None, None, None, None)
        c.step(DT, 10); c.sync()
    sa, sb = a.download_state(), b.download_state()
    for k in "xqvw":
        assert_bit_equal(sa[k], sb[k], f"thread vs level {k}")
    # nothing holds the boxes any more: every dynamic box accelerates at g (10 free-fall steps)
    st0 = [Oracle("swinging_boxes")]
    dyn = st0[0].body_params()["fixed"] == 0
    fc = b.body_dyn(0)["fc"]
    assert_bit_equal(fc, np.zeros_like(fc), "joint forces after the joints were removed")
    assert dyn.sum() > 0
    a.close(); b.close()


def test_external_force_mode_is_per_scene():
    """A sub-range call in absolute mode must not switch gravity off in the other scenes (RigidBodySystem.cpp:58-73)."""
    o = [Oracle("sphere_on_box") for _ in range(3)]
    ctx = make("level", o, max_contacts=8)
    nb = ctx.n_bodies
    f = np.zeros((1, nb, 3), np.float32); f[0, :, 0] = 2.0            # absolute: no gravity, push along x
    ctx.set_external_forces(f=f, scene0=1, n_scenes=1, absolute=True)
    ctx.step(DT, 5); ctx.sync()
    free = Oracle("sphere_on_box"); free.step(DT, 5)
    st = ctx.download_state()
    assert_bit_equal(st["v"][0], free.state()["v"], "scene 0 keeps gravity")
    assert_bit_equal(st["v"][2], free.state()["v"], "scene 2 keeps gravity")
    dyn = free.body_params()["fixed"] == 0
    assert np.all(st["v"][1][dyn, 1] == 0.0) and np.all(st["v"][1][dyn, 0] > 0.0)
    # additive mode on scene 2 does not reinterpret scene 1
    ctx.set_external_forces(f=f, scene0=2, n_scenes=1, absolute=False)
    ctx.step(DT, 1); ctx.sync()
    st2 = ctx.download_state()
    assert np.all(st2["v"][1][dyn, 1] == 0.0)
    assert np.all(st2["v"][2][dyn, 0] > 0.0) and np.all(st2["v"][2][dyn, 1] < st["v"][2][dyn, 1])
    # clearing the sub-range restores gravity there
    ctx.set_external_forces(scene0=1, n_scenes=1)
    ctx.step(DT, 1); ctx.sync()
    assert np.all(ctx.download_state()["v"][1][dyn, 1] < 0.0)
    ctx.close()


def test_contact_overflow_is_reported_once_and_upload_is_validated():
    o = Oracle("marble_box")
    ctx = make("level", [o], max_contacts=64)                         # the scene makes 369 contacts
    ctx.step(DT, 1)
    with pytest.raises(SlrbsError):
        ctx.sync()
    ctx.set_params(collisions=False)
    ctx.step(DT, 1)
    ctx.sync()                                                        # the flag was cleared: no stale error
    bp, st = o.body_params(), o.state()
    bad = bp["geom_type"].copy(); bad[0] = 9
    with pytest.raises(SlrbsError):
        ctx.upload_bodies(st["x"][None], st["q"][None], st["v"][None], st["w"][None], bp["mass"][None], bp["ibody"][None],
                          bp["ibodyinv"][None], bp["fixed"][None], bad[None], bp["geom"][None])
    with pytest.raises(SlrbsError):
        ctx.upload_bodies(st["x"][None], st["q"][None], st["v"][None], st["w"][None], bp["mass"][None], bp["ibody"][None],
                          bp["ibodyinv"][None], bp["fixed"][None], bp["geom_type"][None], bp["geom"][None],
                          counts=np.array([o.n_bodies + 1], np.int32))
    ctx.close()
