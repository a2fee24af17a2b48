"""GPU: the EXTENSION contact pairs (sphere-plane, box-plane, box-box; slrbs_set_extensions) against the restatement.
No reference behaviour exists for these pairs; the device and the C restatement implement the same routines, so the
contact lists (pairs, order, p, n, pene) must still agree bit for bit, and trajectories within fp32 tolerance."""
import numpy as np
import pytest

from oracle_lib import Oracle, BOX, PLANE, SPHERE, CYLINDER
from gpu_helpers import assert_bit_equal, context_from_oracles, push_state

pytestmark = pytest.mark.gpu
DT = 0.01


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
                assert_bit_equal(g[k], c[k], f"scene {s} {k}")


def test_box_stack_contacts_bit_exact_along_its_trajectory():
    steps, oracles, o, done = [0, 1, 30, 200], [], Oracle("box_stack", k=5), 0
    for st in steps:
        o.step(DT, st - done); done = st
        c = Oracle("box_stack", k=5); c.set_state(**o.state())
        oracles.append(c)
    ctx = context_from_oracles(oracles, max_contacts=64)
    ctx.set_extensions(1)
    push_state(ctx, oracles)
    check_contacts(ctx, oracles)
    assert ctx.contact_counts().min() >= 16
    ctx.close()


def test_extension_off_reproduces_the_reference_dispatch():
    o = Oracle("box_stack", k=4); o.set_extensions(0)
    ctx = context_from_oracles([o], max_contacts=64)
    ctx.step(DT, 5); ctx.sync()
    assert ctx.contact_counts()[0] == 0


def test_random_soup_of_boxes_spheres_and_a_plane():
    rng = np.random.default_rng(5)
    oracles = []
    for trial in range(4):
        o = Oracle(); o.set_extensions(1)
        for i in range(60):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            x = rng.uniform(-2.5, 2.5, 3) + [0, 2.0, 0]
            kind = rng.integers(0, 4)
            if kind == 0:
                o.add_body(1.0, SPHERE, [0.3 + 0.3 * rng.random()], x=x)
            elif kind == 3:
                o.add_body(1.0, CYLINDER, [0.5 + rng.random(), 0.3 + 0.3 * rng.random()], x=x, q=q)
            else:
                o.add_body(1.0 + rng.random(), BOX, 0.4 + 1.2 * rng.random(3), x=x, q=q, fixed=bool(rng.integers(0, 8) == 0))
        o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
        oracles.append(o)
    ctx = context_from_oracles(oracles, max_contacts=2048)
    ctx.set_extensions(1)
    check_contacts(ctx, oracles)
    assert ctx.contact_counts().min() > 30
    ctx.close()


def test_many_box_on_box_poses_face_and_edge_cases():
    rng = np.random.default_rng(17)
    oracles = []
    for trial in range(64):
        o = Oracle(); o.set_extensions(1)
        o.add_body(1.0, BOX, [2.0, 1.0, 1.5], fixed=True)
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        if trial % 4 == 0:
            q = np.array([0, 0, 0, 1.0])                                    # axis aligned: face-face, 4 points
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        o.add_body(1.0, BOX, 0.5 + rng.random(3), x=d * rng.uniform(0.6, 1.6), q=q)
        oracles.append(o)
    ctx = context_from_oracles(oracles, max_contacts=16)
    ctx.set_extensions(1)
    check_contacts(ctx, oracles)
    cnt = ctx.contact_counts()
    assert (cnt > 0).sum() > 20 and cnt.max() == 4
    ctx.close()


def test_box_stack_trajectory_and_batch():
    B = 32
    oracles = [Oracle("box_stack", k=5) for _ in range(B)]
    for e, o in enumerate(oracles):                                           # each stack gets its own push on the top box
        st = o.state(); st["v"][4, 0] = 0.02 * e
        o.set_state(**st)
    ctx = context_from_oracles(oracles, max_contacts=64)
    ctx.set_extensions(1)
    ctx.step(DT, 150); ctx.sync()
    g = ctx.download_state()
    for e in (0, 7, 31):
        oracles[e].step(DT, 150)
        s = oracles[e].state()
        assert ctx.contact_counts()[e] == oracles[e].n_contacts
        np.testing.assert_allclose(g["x"][e], s["x"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(g["x"][0, :5, 1], 0.5 + np.arange(5), atol=0.01)  # the unpushed stack stays up
    ctx.close()


def test_many_cylinder_poses_against_box_sphere_and_cylinder():
    """The cylinder pairs the reference ignores (sphere-cylinder, cylinder-box, cylinder-cylinder): device and restatement
    implement the same sampled-surface-against-distance-field routines; lists must agree bit for bit."""
    rng = np.random.default_rng(23)
    oracles = []
    for trial in range(96):
        o = Oracle(); o.set_extensions(1)
        q0 = rng.normal(size=4); q0 /= np.linalg.norm(q0)
        if trial % 3 == 0:
            o.add_body(1.0, BOX, [2.0, 1.0, 1.5], q=q0, fixed=True)
        elif trial % 3 == 1:
            o.add_body(1.0, CYLINDER, [1.5, 0.8], q=q0, fixed=True)
        else:
            o.add_body(1.0, SPHERE, [0.7], fixed=True)
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        if trial % 5 == 0:
            q = np.array([0, 0, 0, 1.0])
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        o.add_body(1.0, CYLINDER, [0.4 + rng.random(), 0.2 + 0.4 * rng.random()], x=d * rng.uniform(0.5, 1.7), q=q)
        oracles.append(o)
    ctx = context_from_oracles(oracles, max_contacts=16)
    ctx.set_extensions(1)
    check_contacts(ctx, oracles)
    cnt = ctx.contact_counts()
    assert (cnt > 0).sum() > 30 and cnt.max() == 4
    ctx.close()


def test_cylinders_resting_on_a_box_trajectory():
    def scene():
        o = Oracle(); o.set_extensions(1)
        o.add_body(1.0, BOX, [8, 0.4, 8], fixed=True)
        o.add_body(1.0, CYLINDER, [1.0, 0.5], x=[-2, 1.0, 0])
        o.add_body(1.0, CYLINDER, [1.0, 0.5], x=[2, 1.0, 0], q=[np.sin(np.pi / 4), 0, 0, np.cos(np.pi / 4)])
        o.add_body(1.0, SPHERE, [0.3], x=[-2, 2.2, 0])
        return o
    o = scene()
    ctx = context_from_oracles([scene(), scene()], max_contacts=32)
    ctx.set_extensions(1)
    ctx.step(DT, 300); ctx.sync()
    o.step(DT, 300)
    assert ctx.contact_counts()[0] == o.n_contacts
    np.testing.assert_allclose(ctx.download_state()["x"][0], o.state()["x"], rtol=0, atol=3e-3)
    assert abs(ctx.download_state()["x"][1][1, 1] - 0.7) < 5e-3
    ctx.close()
