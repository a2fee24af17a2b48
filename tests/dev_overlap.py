"""Developer experiment: the default workload split over K contexts (one stream each) stepping concurrently, so that the
issue-bound kernels of one part (detection, row builders) overlap the latency-bound sweep of another."""
import argparse, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle_lib import Oracle
from gpu_helpers import scene_arrays
from slrbs_b200 import Context

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=2664); ap.add_argument("--parts", type=int, default=2)
ap.add_argument("--steps", type=int, default=30); ap.add_argument("--settle", type=int, default=200)
a = ap.parse_args()
o = Oracle("marble_box_n", nx=10, nz=10, ny=10)
s = scene_arrays(o)
nb = o.n_bodies
Bp = a.B // a.parts
rng = np.random.default_rng(1234)
ctxs = []
for k in range(a.parts):
    ctx = Context(Bp, nb, 0, 4096)
    x = np.broadcast_to(s["x"], (Bp, nb, 3)).copy()
    dyn = s["fixed"] == 0
    x[:, dyn, 0] += rng.uniform(-0.01, 0.01, (Bp, dyn.sum())).astype(np.float32)
    x[:, dyn, 2] += rng.uniform(-0.01, 0.01, (Bp, dyn.sum())).astype(np.float32)
    ctx.upload_bodies(x, s["q"], s["v"], s["w"], s["mass"], s["ibody"], s["ibodyinv"], s["fixed"], s["geom_type"], s["geom"])
    ctxs.append(ctx)
for c in ctxs: c.step(0.01, a.settle)
for c in ctxs: c.sync()
for rep in range(2):
    t0 = time.time()
    for i in range(a.steps):
        for c in ctxs: c.step(0.01, 1)
    for c in ctxs: c.sync()
    dt = (time.time() - t0) / a.steps
    print(f"parts {a.parts} x {Bp} scenes: {dt*1e3:.3f} ms/step  {a.parts*Bp*nb/dt/1e6:.1f} M body-steps/s")
