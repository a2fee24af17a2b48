"""Generates tests/golden/ref_*.npz from oracle/_ref = the REFERENCE'S OWN SOURCES compiled against the
Eigen API shim (oracle/Makefile, target `ref`; needs /root/reference, so it runs in the build container
only).  The fixtures travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py

Per scene and checkpoint k: `pre_*` = state (and joint impulses, which persist across steps) after k
reference steps; `c_*` = the contact list the reference builds during step k+1 (order, pairs, p, n, pene
and the solved impulses); `post_*` = state after step k+1.  `final_*` = state after `total` steps.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Reference  # noqa: E402

DT = 0.01
SCENES = {
    "marble_box": ({}, [0, 40, 120], 150),
    "sphere_on_box": ({}, [0, 82, 90], 200),
    "cylinder_on_plane": ({}, [35, 60, 200], 260),
    "car": ({}, [0, 50, 200], 300),
    "swinging_boxes": ({}, [0, 100], 200),
    "sphere_stack": ({"k": 8}, [0, 5, 100], 200),
    "marble_box_n": (dict(nx=5, nz=5, ny=6), [0, 60, 150], 200),
}


# solverId 1 / 2 (CG / CR, joint-only: SolverConjGradient.cpp, SolverConjResidual.cpp) and a spherical-joint chain
# (Spherical.cpp; no reference scene uses it): one-step checkpoints only, because the reference's Krylov iterations
# amplify last-bit differences by orders of magnitude per step (tests/test_gpu_parity.py)
VARIANTS = {
    "car_cg": ("car", 1, 20, [0, 3, 30]), "car_cr": ("car", 2, 20, [0, 3, 30]),
    "swinging_boxes_cg": ("swinging_boxes", 1, 20, [0, 5, 40]), "swinging_boxes_cr": ("swinging_boxes", 2, 20, [0, 5, 40]),
}


def spherical_chain(cls):
    o = cls()
    o.add_body(1.0, 1, [1, 1, 1], fixed=True, x=[0, 10, 0])
    prev = 0
    for k in range(6):
        b = o.add_body(1.0 + 0.5 * k, 0, [0.3], x=[0.8 * (k + 1), 10, 0], v=[0, 0, 0.5 * k])
        o.add_spherical(prev, b, [0.4, 0, 0], [-0.4, 0, 0])
        prev = b
    o.set_params(0.8, 0, 10, False)
    return o


def variants():
    for name, (scene, solver_id, iters, checkpoints) in VARIANTS.items():
        r = Reference(scene)
        r.set_params(0.8, solver_id, iters, True)
        out = {"checkpoints": np.asarray(checkpoints, np.int32), "solver_id": np.int32(solver_id), "solver_iter": np.int32(iters),
               "dt": np.float32(DT)}
        done = 0
        for k in checkpoints:
            r.step(DT, k - done); done = k
            st = r.state()
            for key in st:
                out[f"pre_{k}_{key}"] = st[key]
            out[f"pre_{k}_jlam"] = r.joint_rows()["lam"]
            r.step(DT); done += 1
            st = r.state()
            for key in st:
                out[f"post_{k}_{key}"] = st[key]
            out[f"post_{k}_jlam"] = r.joint_rows()["lam"]
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **out)
        print(name, "ok")
    r = spherical_chain(Reference)
    out = {"dt": np.float32(DT), "total": np.int32(150)}
    for k in (1, 50, 150):
        r.step(DT, k - (0 if k == 1 else (1 if k == 50 else 50)))
        st = r.state()
        for key in st:
            out[f"at_{k}_{key}"] = st[key]
        out[f"at_{k}_jlam"] = r.joint_rows()["lam"]
    np.savez_compressed(os.path.join(HERE, "ref_spherical_chain.npz"), **out)
    print("spherical_chain ok")


def main():
    variants()
    for scene, (kw, checkpoints, total) in SCENES.items():
        r = Reference(scene, **kw)
        out = {"checkpoints": np.asarray(checkpoints, np.int32), "total": np.int32(total), "dt": np.float32(DT)}
        done = 0
        for k in checkpoints:
            r.step(DT, k - done); done = k
            st = r.state()
            for key in st:
                out[f"pre_{k}_{key}"] = st[key]
            if r.n_joints:
                out[f"pre_{k}_jlam"] = r.joint_rows()["lam"]
            r.step(DT); done += 1
            c = r.contacts()
            for key in ("body0", "body1", "p", "n", "phi", "lam"):
                out[f"c_{k}_{key}"] = c[key]
            st = r.state()
            for key in st:
                out[f"post_{k}_{key}"] = st[key]
            if r.n_joints:
                out[f"post_{k}_jlam"] = r.joint_rows()["lam"]
        r.step(DT, total - done)
        st = r.state()
        for key in st:
            out[f"final_{key}"] = st[key]
        path = os.path.join(HERE, f"ref_{scene}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {os.path.getsize(path)} bytes, contacts at checkpoints "
              f"{[int(out[f'c_{k}_body0'].size) for k in checkpoints]}")


if __name__ == "__main__":
    main()
