"""GPU: SDF-mesh contacts (geom_type 4, csrc/sdf.cuh / sdf.cu) against the restatement. EXTENSION -- the reference has
no SDF code; the device and the C restatement implement the same routines, so the grid built on the
device and the contact lists (pairs, order, p, n, pene) must agree bit for bit, trajectories within fp32 tolerance."""
import numpy as np
import pytest

import mesh_lib
from oracle_lib import Oracle, BOX, PLANE, SDF, SPHERE
from gpu_helpers import assert_bit_equal, context_from_oracles, push_state
from slrbs_b200 import capi

pytestmark = pytest.mark.gpu
DT = 0.01


@pytest.fixture(scope="module")
def shape():
    v, f = mesh_lib.blob()
    return mesh_lib.build_sdf(v, f)


def test_device_grid_equals_the_restatements(shape):
    g = capi.sdf_build(shape["verts"], shape["tris"], shape["dims"], shape["origin"], shape["cell"])
    assert_bit_equal(g, shape["grid"], "sdf grid")
    assert (g < 0).sum() > 100


def pile(shape, rng, n_mesh=6, n_sphere=3):
    o = Oracle(); o.set_extensions(1)
    sid = o.add_sdf_shape(shape)
    for i in range(n_mesh):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        o.add_body(1.0, SDF, [sid], x=[rng.uniform(-0.4, 0.4), 0.9 + 0.95 * i, rng.uniform(-0.4, 0.4)], q=q)
    for i in range(n_sphere):
        o.add_body(1.0, SPHERE, [0.3], x=[rng.uniform(-0.3, 0.3), 1.5 + 0.95 * n_mesh + 0.7 * i, rng.uniform(-0.3, 0.3)])
    o.add_body(1.0, BOX, [0.5, 0.4, 0.6], x=[0.2, 0.9 + 0.95 * n_mesh, -0.1])          # a loose box on top of the meshes
    o.add_body(1.0, BOX, [1.2, 0.3, 1.2], x=[0.0, 0.15, 0.0], fixed=True)                   # a fixed slab on the plane
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    return o


def frozen(shape, src):
    c = Oracle(); c.set_extensions(1)
    sid = c.add_sdf_shape(shape)
    bp = src.body_params(); st = src.state()
    for i in range(src.n_bodies):
        c.add_body(float(bp["mass"][i]), int(bp["geom_type"][i]), [sid] if bp["geom_type"][i] == SDF else bp["geom"][i].tolist(),
                   fixed=bool(bp["fixed"][i]))
    c.set_state(**st)
    return c


def make_ctx(oracles, shape, nc=256):
    ctx = context_from_oracles(oracles, max_contacts=nc)
    ctx.set_extensions(1)
    assert ctx.add_sdf_shape(shape["dims"], shape["origin"], shape["cell"], shape["grid"], shape["verts"]) == 0
    return ctx


def test_contact_lists_bit_exact_along_a_pile_trajectory(shape):
    rng = np.random.default_rng(3)
    o = pile(shape, rng)
    oracles, done = [], 0
    for st in [0, 40, 80, 126, 200, 300]:
        o.step(DT, st - done); done = st
        oracles.append(frozen(shape, o))
    ctx = make_ctx(oracles, shape)
    push_state(ctx, oracles)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.sync()
    cnt = ctx.contact_counts()
    kinds = set()
    for s, c in enumerate(oracles):
        c.stage_prepare(); c.stage_detect()
        assert cnt[s] == c.n_contacts, (s, cnt[s], c.n_contacts)
        if c.n_contacts:
            g, r = ctx.contacts(s), c.contacts()
            np.testing.assert_array_equal(g["body0"], r["body0"]); np.testing.assert_array_equal(g["body1"], r["body1"])
            for k in ("p", "n", "phi"):
                assert_bit_equal(g[k], r[k], f"scene {s} {k}")
            gt = c.body_params()["geom_type"]
            kinds |= {(int(gt[a]), int(gt[b])) for a, b in zip(r["body0"], r["body1"])}
    assert {(SDF, PLANE), (SDF, SDF), (SPHERE, SDF), (SDF, BOX)} <= kinds, kinds
    ctx.close()


def test_pile_trajectory_follows_the_restatement(shape):
    rng = np.random.default_rng(11)
    o = pile(shape, rng, n_mesh=3, n_sphere=1)
    ctx = make_ctx([o, o], shape)
    ctx.step(DT, 60); ctx.sync()
    o.step(DT, 60)
    g, c = ctx.download_state(), o.state()
    for s in range(2):
        np.testing.assert_allclose(g["x"][s], c["x"], atol=5e-3)
    ctx.close()


def write_obj(path, v, f):
    with open(path, "w") as fh:
        fh.write("# procedural blob\no blob\n")
        for p in v:
            fh.write(f"v {p[0]:.6f} {p[1]:.6f} {p[2]:.6f}\n")
        for t in f:
            fh.write(f"f {t[0] + 1}//1 {t[1] + 1}//1 {t[2] + 1}//1\n")


def test_mesh_pile_through_the_headless_program(tmp_path):
    """examples/headless.cpp, written against the reference's class names + the SDFMesh extension: OBJ file -> SDFShape
    (grid built on the device) -> Scenarios::createMeshPile -> RigidBodySystem::step."""
    import os, subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "slrbs_b200", "slrbs_headless")
    assert os.path.exists(exe), "build the host mirror first (make -C slrbs_b200/host)"
    v, f = mesh_lib.blob()
    obj = str(tmp_path / "blob.obj")
    write_obj(obj, v, f)
    n, steps = 3, 120
    dump = str(tmp_path / "frames.txt")
    out = subprocess.run([exe, "mesh_pile", str(steps), "0", obj, str(n)], capture_output=True, text=True, check=True,
                         env=dict(os.environ, SLRBS_DUMP=dump)).stdout.split("\n")
    frames = np.loadtxt(dump)
    assert frames.shape == (steps * (n + 1), 9) and frames[-1, 0] == steps - 1
    rows = np.array([[float(t) for t in ln.split()[1:]] for ln in out if ln and ln[0].isdigit()])
    assert rows.shape == (n + 1, 7) and np.isfinite(rows).all()
    # the same scene in the restatement: vertices as the OBJ text holds them, re-centred on their bounding box
    vr, fr = mesh_lib.load_obj(obj)
    vr = (vr - 0.5 * (vr.min(0) + vr.max(0))).astype(np.float32)
    sh = mesh_lib.build_sdf(vr, fr, cell=0.06, margin=3)
    o = Oracle(); o.set_extensions(1)
    sid = o.add_sdf_shape(sh)
    h = float(vr[:, 1].max() - vr[:, 1].min())
    for i in range(n):
        a = 0.7 * i
        o.add_body(1.0, SDF, [sid], x=[0.05 * (i % 3 - 1), 0.6 * h + 1.1 * h * i, 0.05 * (i % 2)],
                   q=[0.0, np.sin(0.5 * a), 0.0, np.cos(0.5 * a)])
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    o.step(DT, steps)
    np.testing.assert_allclose(rows[:, :3], o.state()["x"], atol=2e-2)
    np.testing.assert_array_equal(frames[-(n + 1):, 2:5].astype(np.float32), rows[:, :3].astype(np.float32))      # last frame = final state
    assert int([ln for ln in out if ln.startswith("contacts")][0].split()[1]) > 0


def test_mesh_pile_with_the_pivoting_solver(shape):
    """BASELINE config 3 asks for the mesh pile under the BPP solver: both extensions together, device vs restatement."""
    rng = np.random.default_rng(11)
    o = pile(shape, rng, n_mesh=3, n_sphere=1)
    o.set_params(solver_id=3, solver_iter=10)
    ctx = make_ctx([o, o], shape)
    ctx.set_params(solver_id=3, solver_iters=10)
    ctx.step(DT, 60); ctx.sync()
    worst = 0.0
    for _ in range(60):
        o.stage_prepare(); o.stage_detect(); o.stage_jacobians(); o.stage_solve(DT)
        worst = max(worst, o.blcp_residual(DT))
        o.stage_scatter(DT); o.stage_integrate(DT)
    assert worst < 1e-4
    g, c = ctx.download_state(), o.state()
    for s in range(2):
        np.testing.assert_allclose(g["x"][s], c["x"], atol=5e-3)
    ctx.close()


def test_mesh_body_without_an_uploaded_shape_makes_no_contacts(shape):
    o = Oracle(); o.set_extensions(1)
    sid = o.add_sdf_shape(shape)
    o.add_body(1.0, SDF, [sid], x=[0, 0.3, 0])
    o.add_body(1.0, SPHERE, [0.3], x=[0.0, 0.9, 0.0])
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    ctx = context_from_oracles([o], max_contacts=32)          # the shape is never given to the context
    ctx.set_extensions(1)
    ctx.step(DT, 3); ctx.sync()
    assert ctx.contact_counts()[0] == 0
    ctx.close()


def mixed_scene(shape, n=150, seed=7):
    """Spheres, boxes and meshes in equal parts on a lattice inside four walls on a slab (dev_measure's mixed_pile)."""
    o = Oracle(); o.set_extensions(1)
    sid = o.add_sdf_shape(shape)
    side = max(2, int(round((n / 6.0) ** (1.0 / 3.0))))
    nlay = (n + side * side - 1) // (side * side)
    rr = np.random.default_rng(seed)
    for i in range(n):
        ix, iz, iy = i % side, (i // side) % side, i // (side * side)
        pos = [(ix - 0.5 * (side - 1)) * 1.1, 1.0 + 1.15 * iy, (iz - 0.5 * (side - 1)) * 1.1]
        q = rr.normal(size=4); q /= np.linalg.norm(q)
        if i % 3 == 0:
            o.add_body(1.0, SPHERE, [0.35], x=pos)
        elif i % 3 == 1:
            o.add_body(1.0, BOX, [0.55, 0.55, 0.55], x=pos, q=q)
        else:
            o.add_body(1.0, SDF, [sid], x=pos, q=q)
    ext = side * 1.1 + 1.0
    o.add_body(1.0, BOX, [ext + 1.0, 0.4, ext + 1.0], x=[0, 0, 0], fixed=True)
    for sx, sz in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        o.add_body(1.0, BOX, [0.4 if sx else ext + 1.0, nlay + 3.0, 0.4 if sz else ext + 1.0],
                   x=[sx * 0.5 * (ext + 0.4), 0.5 * (nlay + 3.0), sz * 0.5 * (ext + 0.4)], fixed=True)
    o.add_body(1.0, PLANE, [0, -0.2, 0, 0, 1, 0], fixed=True)
    return o


@pytest.mark.parametrize("pairs", ["1", "0"], ids=["pair-parallel", "row-parallel"])
def test_mixed_scene_of_160_bodies_bit_exact(shape, pairs, monkeypatch):
    """>= 128 bodies with the extension pairs: detect_pairs.cu (candidates -> warp narrow phase -> per-scene sort) and,
    with SLRBS_DETECT_PAIRS=0, the row-parallel kernels; both must give the restatement's contact list bit for bit."""
    monkeypatch.setenv("SLRBS_DETECT_PAIRS", pairs)
    o = mixed_scene(shape)
    oracles, done = [], 0
    for st in [0, 50, 90, 140, 220]:
        o.step(DT, st - done); done = st
        oracles.append(frozen(shape, o))
    ctx = make_ctx(oracles, shape, nc=4096)
    assert ("k_pair" in ctx.describe()) == (pairs == "1"), ctx.describe()
    push_state(ctx, oracles)
    ctx.stage_prepare(); ctx.stage_detect(); ctx.sync()
    cnt = ctx.contact_counts()
    for s, c in enumerate(oracles):
        c.stage_prepare(); c.stage_detect()
        assert cnt[s] == c.n_contacts, (s, cnt[s], c.n_contacts)
        g, r = ctx.contacts(s), c.contacts()
        np.testing.assert_array_equal(g["body0"], r["body0"]); np.testing.assert_array_equal(g["body1"], r["body1"])
        for k in ("p", "n", "phi"):
            assert_bit_equal(g[k], r[k], f"scene {s} {k}")
    assert cnt.max() > 300
    # and a few whole steps stay together
    ctx.step(DT, 20); ctx.sync()
    for c in oracles:
        c.step(DT, 20)
    g = ctx.download_state()
    for s, c in enumerate(oracles):
        np.testing.assert_allclose(g["x"][s], c.state()["x"], atol=2e-2)      # 20 steps of a tumbling pile, fp32 solver
    ctx.close()
