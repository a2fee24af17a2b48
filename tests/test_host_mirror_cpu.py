"""CPU-only: the C++ mirror's Scenarios.h builders produce exactly the systems the oracle's
restatement of include/rigidbody/Scenarios.h does (bitwise: positions, orientations, inertia and its
inverse, masses, geometry, joint frames, body/joint order); the host C library exports every symbol
include/slrbs_host.h declares."""
import os
import re

import numpy as np
import pytest

from oracle_lib import Oracle
from gpu_helpers import scene_arrays as oracle_arrays, assert_bit_equal
from slrbs_b200 import host_api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [("marble_box", (0, 0, 0), {}), ("sphere_on_box", (0, 0, 0), {}), ("swinging_boxes", (0, 0, 0), {}),
         ("cylinder_on_plane", (0, 0, 0), {}), ("car", (0, 0, 0), {}), ("sphere_stack", (8, 0, 0), {"k": 8}),
         ("marble_box_n", (4, 3, 5), {"nx": 4, "nz": 3, "ny": 5})]


@pytest.fixture(scope="module", autouse=True)
def _build():
    if not os.path.exists(host_api.host_lib_path()):
        import __graft_entry__ as ge
        ge.build()


@pytest.mark.parametrize("scene,abc,kw", CASES)
def test_scene_builders_match_reference_restatement(scene, abc, kw):
    h = host_api.scene_arrays(scene, *abc)
    o = oracle_arrays(Oracle(scene, **kw))
    assert h["x"].shape == o["x"].shape
    for k in ("x", "q", "v", "w", "mass", "geom", "r0", "q0", "r1", "q1"):
        assert_bit_equal(h[k], o[k], f"{scene}:{k}")
    for k in ("fixed", "geom_type", "jtype", "jb0", "jb1"):
        np.testing.assert_array_equal(h[k], o[k], err_msg=f"{scene}:{k}")
    assert_bit_equal(h["ibody"], o["ibody"], "ibody")
    # plane inertia inverse is NaN/inf on both sides (RigidBody.cpp:28 on a zero matrix)
    fin = np.isfinite(o["ibodyinv"]).all(axis=1)
    assert_bit_equal(h["ibodyinv"][fin], o["ibodyinv"][fin], "ibodyinv")
    assert (~np.isfinite(h["ibodyinv"][~fin])).any(axis=1).all()


def test_host_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "slrbs_host.h")).read()
    declared = sorted(set(re.findall(r"\b(slrbs_host_[a-z_0-9]+)\s*\(", hdr)))
    L = host_api.load_host_library()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert declared == sorted(host_api.SYMBOLS)


def test_host_step_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    s = host_api.HostSystem("car")
    with pytest.raises(host_api.SlrbsError):
        s.step(0.01, 1)
