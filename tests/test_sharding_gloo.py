"""World-size-2 gloo test of the N>1 host logic (CPU): contiguous scene blocks partition the batch,
synthetic data is independent of the rank count, and the max/sum aggregation over ranks works."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slrbs_b200 import sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = sharding.scene_block(total, rank, world)
    jit = sharding.jitter_for_scenes(first, n, 5, 0.01)
    ms = 10.0 + 3.0 * rank                                     # rank 1 is the slow one
    mx = sharding.reduce_over_ranks([ms], "max")[0]
    sm = sharding.reduce_over_ranks([float(n)], "sum")[0]
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([first], dtype=torch.int64))
    ret[rank] = dict(first=first, n=n, jit=jit, mx=mx, sm=sm, firsts=[int(g) for g in gathered])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scene_sharding():
    world, total = 2, 2500                                      # not a multiple of the jitter block or of world
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), total, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["first"] == 0 and r0["n"] + r1["n"] == total and r1["first"] == r0["n"]
    assert r0["mx"] == r1["mx"] == 13.0 and r0["sm"] == r1["sm"] == float(total)
    assert r0["firsts"] == [0, r0["n"]]
    whole = sharding.jitter_for_scenes(0, total, 5, 0.01)
    np.testing.assert_array_equal(np.concatenate([r0["jit"], r1["jit"]]), whole)
    assert np.abs(whole).max() <= 0.01 and whole.std() > 0.005
    assert abs(sharding.job_throughput(100.0, 2, 13.0) - 200.0 / 0.013) < 1e-6


def test_blocks_partition_for_many_world_sizes():
    for total in (1, 7, 4096, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharding.scene_block(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
            for (a, n), (b, _) in zip(blocks, blocks[1:]):
                assert a + n == b
