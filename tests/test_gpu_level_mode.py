"""GPU: the level-scheduled PGS kernel (rows that share no dynamic body run in parallel) must give the
SAME bits as the thread-per-scene kernel that walks the rows sequentially: same impulses, same constraint
forces, same trajectories. Also exercises every lane-group size."""
import os

import numpy as np
import pytest

from oracle_lib import Oracle
from gpu_helpers import context_from_oracles, assert_bit_equal

pytestmark = pytest.mark.gpu
DT = 0.01


# variant -> (SLRBS_PGS_SPLIT, SLRBS_PGS_PIPE, SLRBS_PGS_GACC, SLRBS_PGS_STAGE, kernel name in slrbs_describe)
VARIANTS = {"lv": ("0", "0", "0", "0", "k_pgs_lv"), "split": ("1", "0", "0", "0", "k_pgs_sp"), "pipe": ("0", "1", "0", "0", "k_pgs_pipe"),
            "gacc": ("0", "1", "1", "0", "k_pgs_gacc"), "stage": ("0", "1", "0", "1", "k_pgs_stage")}


def make(mode, oracles, group=None, variant=None, **kw):
    """variant: None = the library's choice; for the wide-scene groups "lv" = one lane per row (k_pgs_lv), "split" = two
    lanes per contact row (k_pgs_sp, groups 32 / 64), "pipe" = one warp per contact-only scene with register prefetch
    (k_pgs_pipe, group 32), "gacc" = one warp per contact-only scene with the accumulators in global memory (k_pgs_gacc)."""
    os.environ["SLRBS_PGS_MODE"] = mode
    if group:
        os.environ["SLRBS_LEVEL_GROUP"] = str(group)
    if variant is not None:
        (os.environ["SLRBS_PGS_SPLIT"], os.environ["SLRBS_PGS_PIPE"], os.environ["SLRBS_PGS_GACC"],
         os.environ["SLRBS_PGS_STAGE"]) = VARIANTS[variant][:4]
    try:
        return context_from_oracles(oracles, **kw)
    finally:
        for k in ("SLRBS_PGS_MODE", "SLRBS_LEVEL_GROUP", "SLRBS_PGS_SPLIT", "SLRBS_PGS_PIPE", "SLRBS_PGS_GACC", "SLRBS_PGS_STAGE"):
            os.environ.pop(k, None)


def jittered(scene, n, **kw):
    out = []
    rng = np.random.default_rng(5)
    for _ in range(n):
        o = Oracle(scene, **kw)
        st = o.state(); dyn = o.body_params()["fixed"] == 0
        st["x"][dyn, 0] += rng.uniform(-0.01, 0.01, dyn.sum()).astype(np.float32)
        st["x"][dyn, 2] += rng.uniform(-0.01, 0.01, dyn.sum()).astype(np.float32)
        o.set_state(**st)
        out.append(o)
    return out


@pytest.mark.parametrize("scene,kw,n_scenes,steps,group,variant", [
    ("marble_box", {}, 5, 60, 8, None), ("marble_box", {}, 3, 25, 4, None), ("marble_box", {}, 3, 25, 16, None),
    ("marble_box", {}, 2, 25, 32, "lv"), ("marble_box", {}, 2, 25, 32, "split"), ("marble_box", {}, 2, 25, 64, "split"),
    ("marble_box", {}, 2, 25, 32, "pipe"), ("sphere_on_box", {}, 2, 120, 32, "pipe"), ("cylinder_on_plane", {}, 2, 80, 32, "pipe"),
    ("marble_box", {}, 2, 25, 32, "gacc"), ("sphere_on_box", {}, 2, 120, 32, "gacc"), ("cylinder_on_plane", {}, 2, 80, 32, "gacc"),
    ("marble_box", {}, 2, 25, 32, "stage"), ("sphere_on_box", {}, 2, 120, 32, "stage"), ("cylinder_on_plane", {}, 2, 80, 32, "stage"),
    ("car", {}, 7, 80, 8, None), ("swinging_boxes", {}, 3, 60, 8, None), ("cylinder_on_plane", {}, 2, 80, 8, None),
    # joints and cylinder-plane contacts through the side-split kernel (its joint rows run on the even lanes)
    ("car", {}, 3, 80, 32, "split"), ("swinging_boxes", {}, 2, 60, 32, "split"), ("cylinder_on_plane", {}, 2, 80, 64, "split"),
    ("marble_box_n", dict(nx=5, nz=4, ny=6), 3, 40, 16, None)])
def test_level_mode_is_bit_identical_to_sequential(scene, kw, n_scenes, steps, group, variant):
    oracles = jittered(scene, n_scenes, **kw)
    a = make("thread", oracles, max_contacts=1024)
    b = make("level", oracles, group=group, variant=variant, max_contacts=1024)
    if variant is not None:
        assert VARIANTS[variant][4] in b.describe(), b.describe()
    for k in range(steps):
        a.step(DT); b.step(DT)
        if k in (0, 1, steps // 2, steps - 1):
            a.sync(); b.sync()
            np.testing.assert_array_equal(a.contact_counts(), b.contact_counts())
            sa, sb = a.download_state(), b.download_state()
            for key in "xqvw":
                assert_bit_equal(sa[key], sb[key], f"{scene} step {k} {key}")
            for s in range(n_scenes):
                ca, cb = a.contacts(s), b.contacts(s)
                np.testing.assert_array_equal(ca["body0"], cb["body0"])
                assert_bit_equal(ca["lam"], cb["lam"], f"{scene} step {k} contact lambda")
                da, db = a.body_dyn(s), b.body_dyn(s)
                assert_bit_equal(da["fc"], db["fc"], "fc"); assert_bit_equal(da["tauc"], db["tauc"], "tauc")
                if a.n_joints:
                    ra, rb = a.joint_rows(s), b.joint_rows(s)
                    assert_bit_equal(ra["J0Minv"], rb["J0Minv"], "joint rows")
            if a.n_joints:
                assert_bit_equal(a.joint_lambda(), b.joint_lambda(), "joint lambda")
    a.close(); b.close()


def test_level_mode_matches_oracle_marble_box():
    """Level mode against the oracle directly, including the un-permuted row export (rows are stored in
    level order on the device)."""
    o = Oracle("marble_box")
    ctx = make("level", [o], max_contacts=1024)
    ctx.step(DT, 30); o.step(DT, 30); ctx.sync()
    assert ctx.contact_counts()[0] == o.n_contacts
    np.testing.assert_allclose(ctx.download_state()["x"][0], o.state()["x"], rtol=0, atol=2e-3)
    ctx.upload_state(**o.state())
    ctx.stage_prepare(); ctx.stage_detect(); ctx.stage_jacobians(DT); ctx.sync()
    o.stage_prepare(); o.stage_detect(); o.stage_jacobians()
    rows, orow = ctx.contact_rows(0), o.contact_rows()
    for k in ("J0", "J1", "J0Minv", "J1Minv"):
        assert_bit_equal(rows[k], orow[k], k)
    assert_bit_equal(rows["rhs"], o.solver_blocks(DT)["rhs"], "rhs")
    ctx.close()


def test_krylov_in_level_mode():
    o = [Oracle("swinging_boxes") for _ in range(2)]
    for x in o:
        x.set_params(solver_id=2, solver_iter=20)
    a = make("thread", o, max_contacts=8); b = make("level", o, max_contacts=8)
    a.set_params(solver_id=2, solver_iters=20); b.set_params(solver_id=2, solver_iters=20)
    a.step(DT, 10); b.step(DT, 10); a.sync(); b.sync()
    for key in "xqvw":
        assert_bit_equal(a.download_state()[key], b.download_state()[key], key)
    a.close(); b.close()


@pytest.mark.parametrize("compact", ["0", "1"], ids=["rows256", "rows208"])
@pytest.mark.parametrize("variant", ["lv", "split", "pipe", "gacc", "stage"])
def test_big_scene_variant_with_packed_accumulators(variant, compact, monkeypatch):
    """Thousand-body scenes take the kernel variants with 24 B packed accumulators (9 instead of 7 scenes per SM): one
    lane per row (k_pgs_lv<32>), two lanes per contact row (k_pgs_sp<32>) or the register-pipelined warp (k_pgs_pipe, the
    default for contact-only scenes). All must still give the bits of the sequential kernel."""
    if compact == "1" and variant in ("lv", "split"):
        pytest.skip("13-field row records are read by k_pgs_pipe / k_pgs_gacc / k_pgs_stage only")
    if compact == "0" and variant == "stage":
        pytest.skip("k_pgs_stage always reads the 13-field records")
    monkeypatch.setenv("SLRBS_ROWS_COMPACT", compact)    # 208 B records: a0d, a1d recomputed from rr0, rr1, bit-identically
    oracles = jittered("marble_box_n", 2, nx=10, nz=10, ny=10)
    for o in oracles:
        o.step(DT, 60)                                   # let contacts with the walls and floor form
    a = make("thread", oracles, max_contacts=4096)
    b = make("level", oracles, group=32, variant=variant, max_contacts=4096)
    assert VARIANTS[variant][4] in b.describe(), b.describe()
    for k in range(12):
        a.step(DT); b.step(DT)
    a.sync(); b.sync()
    np.testing.assert_array_equal(a.contact_counts(), b.contact_counts())
    assert a.contact_counts().min() > 1000
    sa, sb = a.download_state(), b.download_state()
    for key in "xqvw":
        assert_bit_equal(sa[key], sb[key], f"big scene {key}")
    for s in range(2):
        assert_bit_equal(a.contacts(s)["lam"], b.contacts(s)["lam"], "big scene contact lambda")
        da, db = a.body_dyn(s), b.body_dyn(s)
        assert_bit_equal(da["fc"], db["fc"], "fc"); assert_bit_equal(da["tauc"], db["tauc"], "tauc")
        ra, rb = a.contact_rows(s), b.contact_rows(s)        # the exported Jacobian blocks, whichever record form is stored
        for key in ("J0", "J1", "J0Minv", "J1Minv", "A", "rhs"):
            assert_bit_equal(ra[key], rb[key], f"big scene rows {key}")
    a.close(); b.close()
