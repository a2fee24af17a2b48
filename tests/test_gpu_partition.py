"""GPU: one pile cut into overlapping blocks (slrbs_b200/partition.py) against the same pile stepped as a single
scene. Across blocks the coupling is block-Jacobi with one layer of overlap, so the comparison is on penetration,
mean drift and energy, not on trajectories. The cross-GPU path (NCCL all_to_all of packed halo state) runs when
the box has two GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from slrbs_b200.partition import PartitionedPile, sphere_pile

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_blocks_match_single_scene_on_penetration_and_drift():
    x, r, m, fixed = sphere_pile(12, 6, 10)
    n = x.shape[0]
    cut = PartitionedPile(x, r, m, fixed, block=4.0, rebuild_every=8)
    one = PartitionedPile(x, r, m, fixed, block=64.0, rebuild_every=10 ** 9)
    assert cut.stats["scenes"] >= 12 and one.stats["scenes"] == 1 and cut.stats["halo_bodies"] > 0
    cut.step(120); one.step(120)
    a, b = cut.gather_state(), one.gather_state()
    assert np.isfinite(a).all()
    drift = np.linalg.norm(a[:, :3] - b[:, :3], axis=1)
    assert cut.stats["rebuilds"] >= 10
    assert cut.max_penetration() < 0.05 and one.max_penetration() < 0.05
    assert drift.mean() < 0.15, drift.mean()                      # sphere diameter 1 m, pile 6 m high; the jittered pile is chaotic
    assert abs(a[:, 1].mean() - b[:, 1].mean()) < 0.02             # same settling height
    assert a[:, 1].min() > 0.2 + 0.5 - 0.06                        # nobody fell through the floor
    cut.close(); one.close()


def test_block_run_is_deterministic_and_owns_every_body():
    x, r, m, fixed = sphere_pile(8, 4, 8)
    runs = []
    for _ in range(2):
        p = PartitionedPile(x, r, m, fixed, block=4.0, rebuild_every=5)
        p.step(23)
        runs.append(p.gather_state()); p.close()
    np.testing.assert_array_equal(runs[0], runs[1])
    assert (np.abs(runs[0][:, 3:7]).sum(axis=1) > 0.5).all()       # every body was written by its owner


def test_two_gpus_nccl_halo_exchange():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "run_partition_multi.py"), "--steps", "40"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["finite"] and d["stats"]["remote_halos"] > 0
    assert d["max_penetration"] < 0.05 and d["mean_drift"] < 0.15
