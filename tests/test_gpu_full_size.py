"""GPU, at the bench's FULL default size (marble_box_1k x 2664 scenes, 2.7 M bodies, ~6.9 M contacts): properties that
do not need the oracle to run the whole batch.

1. After 200 device-only steps, the contact list the device builds from the state of sampled scenes equals, bit for bit
   (pairs, order, p, n, pene), what the oracle's all-pairs CollisionDetect builds from the same state; one further step
   from that state agrees with the oracle within the fp32 tolerances of tests/test_gpu_parity.py.
2. Structural invariants over ALL scenes: contact counts within capacity, every body state finite, nobody below the
   floor, unit quaternions.
3. Determinism: two contexts fed the same batch produce identical bits after 20 steps (no atomics-order or
   scheduling dependence anywhere in the step).
"""
import numpy as np
import pytest

import bench
from gpu_helpers import assert_bit_equal
from oracle_lib import Oracle
from slrbs_b200 import Context

pytestmark = pytest.mark.gpu
DT = 0.01


def build(B, rank=0):
    s, x, nb, nj, nc, settle = bench.build_batch("marble_box_1k", B, rank)
    ctx = Context(B, nb, nj, nc)
    ctx.set_params(bench.MU, 0, bench.SOLVER_ITERS, True)
    ctx.upload_bodies(x, s["q"], s["v"], s["w"], s["mass"], s["ibody"], s["ibodyinv"], s["fixed"], s["geom_type"], s["geom"])
    return ctx, nb, settle


def test_full_size_batch_contacts_match_oracle_on_sampled_scenes():
    B = bench.WORKLOADS["marble_box_1k"][3]
    ctx, nb, settle = build(B)
    ctx.step(DT, settle); ctx.sync()
    st = ctx.download_state()
    cnt = ctx.contact_counts()
    # 2. invariants over every scene
    assert cnt.max() <= ctx.max_contacts and cnt.min() > 1500
    for k in "xqvw":
        assert np.isfinite(st[k]).all()
    assert st["x"][:, :1000, 1].min() > 0.2 + 0.5 - 0.08                       # floor top at 0.2, radius 0.5
    np.testing.assert_allclose(np.linalg.norm(st["q"], axis=2), 1.0, atol=1e-5)
    # 1. sampled scenes against the oracle
    ctx.stage_prepare(); ctx.stage_detect(); ctx.sync()
    for s in (0, B // 3, B - 1):
        o = Oracle("marble_box_n", nx=10, nz=10, ny=10)
        o.set_state(x=st["x"][s], q=st["q"][s], v=st["v"][s], w=st["w"][s])
        o.stage_prepare(); o.stage_detect()
        g, c = ctx.contacts(s), o.contacts()
        assert ctx.contact_counts()[s] == o.n_contacts
        np.testing.assert_array_equal(g["body0"], c["body0"]); np.testing.assert_array_equal(g["body1"], c["body1"])
        for k in ("p", "n", "phi"):
            assert_bit_equal(g[k], c[k], f"scene {s} {k}")
    ctx.stage_jacobians(DT); ctx.stage_solve(DT); ctx.stage_integrate(DT); ctx.sync()
    after = ctx.download_state()
    for s in (0, B - 1):
        o = Oracle("marble_box_n", nx=10, nz=10, ny=10)
        o.set_state(x=st["x"][s], q=st["q"][s], v=st["v"][s], w=st["w"][s])
        o.step(DT)
        np.testing.assert_allclose(after["x"][s], o.state()["x"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(after["v"][s], o.state()["v"], rtol=0, atol=2e-3)
    ctx.close()


def test_full_size_step_is_deterministic():
    B = 1332
    a, nb, _ = build(B)
    b, _, _ = build(B)
    a.step(DT, 20); b.step(DT, 20); a.sync(); b.sync()
    sa, sb = a.download_state(), b.download_state()
    np.testing.assert_array_equal(a.contact_counts(), b.contact_counts())
    for k in "xqvw":
        assert_bit_equal(sa[k], sb[k], k)
    a.close(); b.close()
