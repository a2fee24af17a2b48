"""Developer measurement tool (not the bench): batched copies of an oracle-built scene on the GPU."""
import argparse, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle_lib import Oracle
from gpu_helpers import scene_arrays
from slrbs_b200 import Context

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="marble_box"); ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--nc", type=int, default=1024); ap.add_argument("--settle", type=int, default=200)
ap.add_argument("--steps", type=int, default=20); ap.add_argument("--k", type=int, default=8)
ap.add_argument("--mode", default=""); ap.add_argument("--group", default="")
ap.add_argument("--solver", type=int, default=0); ap.add_argument("--ext", type=int, default=0)
a = ap.parse_args()
if a.mode: os.environ["SLRBS_PGS_MODE"] = a.mode
if a.group: os.environ["SLRBS_LEVEL_GROUP"] = a.group
shape = None
if a.scene == "mesh_pile":
    import mesh_lib
    from oracle_lib import SDF, PLANE
    mv, mf = mesh_lib.blob(); shape = mesh_lib.build_sdf(mv, mf)
    o = Oracle(); o.set_extensions(1); sid = o.add_sdf_shape(shape)
    for i in range(a.k):
        ang = 0.7 * i
        o.add_body(1.0, SDF, [sid], x=[0.05 * (i % 3 - 1), 0.6 + 1.1 * i, 0.05 * (i % 2)], q=[0.0, np.sin(0.5 * ang), 0.0, np.cos(0.5 * ang)])
    o.add_body(1.0, PLANE, [0, 0, 0, 0, 1, 0], fixed=True)
    a.ext = 1
elif a.scene == "mixed_pile":
    # BASELINE config 3: mixed spheres / boxes / meshes poured into a box container, a.k bodies per scene
    import mesh_lib
    from oracle_lib import SDF, BOX, SPHERE
    mv, mf = mesh_lib.blob(radius=0.35); shape = mesh_lib.build_sdf(mv, mf, cell=0.05)
    o = Oracle(); o.set_extensions(1); sid = o.add_sdf_shape(shape)
    side = max(2, int(round((a.k / 6.0) ** (1.0 / 3.0))))
    nlay = (a.k + side * side - 1) // (side * side)
    rr = np.random.default_rng(7)
    for i in range(a.k):
        ix, iz, iy = i % side, (i // side) % side, i // (side * side)
        pos = [(ix - 0.5 * (side - 1)) * 1.0, 0.9 + 1.0 * iy, (iz - 0.5 * (side - 1)) * 1.0]
        q = rr.normal(size=4); q /= np.linalg.norm(q)
        kind = i % 3
        if kind == 0: o.add_body(1.0, SPHERE, [0.35], x=pos)
        elif kind == 1: o.add_body(1.0, BOX, [0.55, 0.55, 0.55], x=pos, q=q)
        else: o.add_body(1.0, SDF, [sid], x=pos, q=q)
    ext = side * 1.0 + 1.0
    o.add_body(1.0, BOX, [ext + 1.0, 0.4, ext + 1.0], x=[0, 0, 0], fixed=True)
    for sx, sz in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        o.add_body(1.0, BOX, [0.4 if sx else ext + 1.0, nlay + 3.0, 0.4 if sz else ext + 1.0], x=[sx * 0.5 * (ext + 0.4), 0.5 * (nlay + 3.0), sz * 0.5 * (ext + 0.4)], fixed=True)
    a.ext = 1
else:
  o = Oracle(a.scene, k=a.k) if a.scene in ("sphere_stack", "box_stack") else (Oracle("marble_box_n", nx=10, nz=10, ny=10) if a.scene == "mb1k" else Oracle(a.scene))
s = scene_arrays(o)
B, nb, nj = a.B, o.n_bodies, o.n_joints
ctx = Context(B, nb, nj, a.nc)
ctx.set_params(solver_id=a.solver)
if a.ext: ctx.set_extensions(1)
if shape is not None: ctx.add_sdf_shape(shape["dims"], shape["origin"], shape["cell"], shape["grid"], shape["verts"])
rng = np.random.default_rng(1234)
x = np.broadcast_to(s["x"], (B, nb, 3)).copy()
dyn = s["fixed"] == 0
x[:, dyn, 0] += rng.uniform(-0.01, 0.01, (B, dyn.sum())).astype(np.float32)
x[:, dyn, 2] += rng.uniform(-0.01, 0.01, (B, dyn.sum())).astype(np.float32)
ctx.upload_bodies(x, s["q"], s["v"], s["w"], s["mass"], s["ibody"], s["ibodyinv"], s["fixed"], s["geom_type"], s["geom"])
if nj: ctx.upload_joints(s["jtype"], s["jb0"], s["jb1"], s["r0"], s["q0"], s["r1"], s["q1"])
t0 = time.time(); ctx.step(0.01, a.settle); ctx.sync(); t1 = time.time()
print(f"settle {a.settle} steps: {t1-t0:.3f}s  contacts/scene mean {ctx.contact_counts().mean():.1f} max {ctx.contact_counts().max()}")
if a.solver:
    t0 = time.time(); ctx.step(0.01, a.steps); ctx.sync(); t1 = time.time()
    dt = (t1 - t0) / a.steps
    print(f"{a.scene} B={B} solver {a.solver}: {dt*1e3:.3f} ms/step  {B*nb/dt/1e6:.1f} M body-steps/s")
    if a.solver == 3:
        ctx.profile_enable(True); ctx.profile_read(True); ctx.step(0.01, 5); ctx.sync()
        p = ctx.profile_read(True)
        print(f"  pivot steps per scene-solve: {p['contact_visits'] / (5 * B):.2f}  solve {p['solve_ms'] / 5:.3f} ms/step")
    sys.exit(0)
ctx.profile_enable(True); ctx.profile_read(True)
t0 = time.time(); ctx.step(0.01, a.steps); ctx.sync(); t1 = time.time()
p = ctx.profile_read(True)
dt = (t1 - t0) / a.steps
alg = 464 * p["contact_visits"] + 744 * p["joint_visits"]
print(f"{a.scene} B={B}: {dt*1e3:.3f} ms/step  {B*nb/dt/1e6:.1f} M body-steps/s | pgs {p['solve_ms']/p['solve_launches']:.3f} ms/launch "
      f"share {p['solve_ms']/1e3/(t1-t0):.2f}  alg {alg/ (p['solve_ms']/1e3)/1e9:.1f} GB/s  visits/launch c={p['contact_visits']/p['solve_launches']:.0f} j={p['joint_visits']/p['solve_launches']:.0f}")
st = ctx.download_state(); bad = (~np.isfinite(st['x']).all(axis=(1, 2))).sum()
if bad: print(f"  scenes with non-finite positions: {bad} of {B}")
