"""Loader for tests/golden/ref_*.npz (made by tests/golden/make_golden.py from oracle/_ref, i.e. from the
reference's own sources; see that script for the layout)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENE_KW = {"marble_box": {}, "sphere_on_box": {}, "cylinder_on_plane": {}, "car": {}, "swinging_boxes": {},
            "sphere_stack": {"k": 8}, "marble_box_n": dict(nx=5, nz=5, ny=6)}


def load(scene):
    z = np.load(os.path.join(GOLDEN_DIR, f"ref_{scene}.npz"))
    return {k: z[k] for k in z.files}


def state(g, prefix):
    return {k: g[f"{prefix}_{k}"] for k in ("x", "q", "v", "w")}


def contacts(g, k):
    return {key: g[f"c_{k}_{key}"] for key in ("body0", "body1", "p", "n", "phi", "lam")}
