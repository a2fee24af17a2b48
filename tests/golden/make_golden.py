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


def main():
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
