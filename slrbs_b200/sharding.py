"""Scene sharding for multi-GPU runs: independent scenes are split by scene index into contiguous
blocks, one block per rank, with no data-path collective. Only timings and counters
cross ranks (torch.distributed all_reduce), never body state."""
import numpy as np

JITTER_BLOCK = 1024     # scenes per jitter block: synthetic data is generated per global block so that
                        # a scene's initial state does not depend on how many ranks share the batch


def scene_block(total_scenes, rank, world):
    """(first scene, count) of the contiguous block of global scene indices owned by `rank`."""
    base, rem = divmod(int(total_scenes), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def jitter_for_scenes(first_scene, n_scenes, n_items, amplitude, seed=1234):
    """Deterministic U(-a, a) x/z offsets [n_scenes, n_items, 2] keyed by GLOBAL scene index."""
    out = np.empty((n_scenes, n_items, 2), np.float32)
    s = first_scene
    while s < first_scene + n_scenes:
        blk = s // JITTER_BLOCK
        lo = blk * JITTER_BLOCK
        rng = np.random.default_rng([seed, blk])
        full = rng.uniform(-amplitude, amplitude, (JITTER_BLOCK, n_items, 2)).astype(np.float32)
        e = min(lo + JITTER_BLOCK, first_scene + n_scenes)
        out[s - first_scene:e - first_scene] = full[s - lo:e - lo]
        s = e
    return out


def reduce_over_ranks(values, op="max"):
    """all_reduce a list of floats over the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return [float(v) for v in t.cpu()]


def job_throughput(units_per_rank, world, ms_max):
    """Whole-job units/s: every rank processed `units_per_rank` units in at most `ms_max` ms."""
    return float(units_per_rank) * world / (ms_max * 1e-3)
