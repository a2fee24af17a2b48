"""CPU: the block partition of a large pile (slrbs_b200/partition.py). Every body is owned exactly once, every
body that can touch an owned body is present in that block's scene, and the halo send / receive lists that
the ranks derive independently agree (checked in-process for 3 ranks and over gloo for 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slrbs_b200.partition import build_partition, halo_maps, sphere_pile


def pile(n=(14, 6, 9), seed=3):
    x, r, m, fixed = sphere_pile(*n, jitter=0.05, seed=seed)
    return x


def test_every_body_owned_once_and_neighbours_present():
    x = pile()
    n = x.shape[0]
    block, halo_w = 4.0, 1.3
    part = build_partition(x, block, halo_w, world=3)
    owned_count = np.zeros(n, int)
    present = {}
    for r, scenes in enumerate(part["ranks"]):
        for s, sc in enumerate(scenes):
            g = sc["ids"]
            assert (np.diff(g) > 0).all()                               # sorted by global id, no duplicates
            owned_count[g[sc["owned"]]] += 1
            present[(r, s)] = set(g.tolist())
            for slot in np.nonzero(sc["owned"])[0]:
                b = g[slot]
                assert part["owner_rank"][b] == r and part["owner_scene"][b] == s and part["owner_slot"][b] == slot
    assert (owned_count == 1).all()
    # neighbours within the halo width (a superset of "can touch") are in the owner's scene
    d = np.linalg.norm(x[:, None, :] - x[None, :, :], axis=-1)
    ii, jj = np.nonzero((d < halo_w - 1e-6) & (d > 0))
    for i, j in zip(ii[::7], jj[::7]):
        assert j in present[(part["owner_rank"][i], part["owner_scene"][i])]
    # slabs are contiguous in x and balanced
    sizes = [sum(int(sc["owned"].sum()) for sc in scenes) for scenes in part["ranks"]]
    assert sum(sizes) == n and min(sizes) > 0.15 * n


def test_halo_maps_agree_between_ranks():
    x = pile()
    world = 3
    part = build_partition(x, 4.0, 1.3, world)
    NB = part["max_bodies"] + 5
    maps = [halo_maps(part, r, NB) for r in range(world)]

    def gid(rank, flat):
        return np.array([part["ranks"][rank][f // NB]["ids"][f % NB] for f in flat], np.int64)

    for r in range(world):
        # local copies: source is owned here, destination is a halo slot of the same body
        assert (gid(r, maps[r]["local_src"]) == gid(r, maps[r]["local_dst"])).all()
        for f in maps[r]["local_src"]:
            assert part["ranks"][r][f // NB]["owned"][f % NB]
        for p in range(world):
            if p == r:
                continue
            assert maps[r]["send"][p].size == maps[p]["recv"][r].size
            assert (gid(r, maps[r]["send"][p]) == gid(p, maps[p]["recv"][r])).all()
    assert sum(m["send"][p].size for m in maps for p in m["send"]) > 0       # there is cross-rank traffic


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = pile()
    part = build_partition(x, 4.0, 1.3, world)
    NB = part["max_bodies"] + 5
    m = halo_maps(part, rank, NB)
    peer = 1 - rank
    ids_out = torch.tensor([part["ranks"][rank][f // NB]["ids"][f % NB] for f in m["send"][peer]], dtype=torch.int64)
    ids_in = torch.zeros(m["recv"][peer].size, dtype=torch.int64)
    reqs = [dist.isend(ids_out, peer), dist.irecv(ids_in, peer)]
    for q in reqs:
        q.wait()
    want = torch.tensor([part["ranks"][rank][f // NB]["ids"][f % NB] for f in m["recv"][peer]], dtype=torch.int64)
    ret[rank] = bool((ids_in == want).all()) and ids_in.numel() > 0
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_exchange_matching_halo_lists_over_gloo():
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


# ---- fixed-grid partition (host mirror of the device re-partition, csrc/pile.cu) -------------------------------
def test_fixed_grid_mirror_owns_every_body_once_and_keeps_neighbours():
    from slrbs_b200.partition import fixed_grid, grid_partition_numpy, slab_bounds
    x = pile()
    n = x.shape[0]
    block, halo_w = 4.0, 1.3
    origin, ncell = fixed_grid(x, block, 2 * halo_w)
    bounds = slab_bounds(x, origin, block, int(ncell[0]), 3)
    assert bounds[0] == 0 and bounds[-1] == ncell[0] and all(b > a for a, b in zip(bounds, bounds[1:]))
    owned = np.zeros(n, int)
    scene_of = {}
    members = []
    for r in range(3):
        for s, keys in enumerate(grid_partition_numpy(x, origin, block, halo_w, ncell, bounds[r], bounds[r + 1])):
            assert (np.diff(keys >> 1) > 0).all()                        # sorted by global id, unique
            owned[keys[(keys & 1) == 0] >> 1] += 1
            for g in keys[(keys & 1) == 0] >> 1:
                scene_of[int(g)] = len(members)
            members.append(set((keys >> 1).tolist()))
    assert (owned == 1).all()
    d = np.linalg.norm(x[:, None, :] - x[None, :, :], axis=-1)
    ii, jj = np.nonzero((d < halo_w - 1e-5) & (d > 0))
    for i, j in zip(ii[::5], jj[::5]):
        assert int(j) in members[scene_of[int(i)]]                       # whoever can touch an owned body is in its scene


def _device_protocol_worker(rank, world, port, ret):
    """One rank of the device-side re-partition's exchange protocol, with the device kernels replaced by their host
    mirror: requests for remote halo copies -> ids the neighbours ask of me -> the per-step state exchange in those
    orders. Every halo copy must end up with its owner's state."""
    from slrbs_b200.partition import exchange_requests, fixed_grid, grid_partition_numpy, slab_bounds, STATE_W
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = pile((16, 5, 6))
    n = x.shape[0]
    block, halo_w = 3.0, 1.3
    origin, ncell = fixed_grid(x, block, 2 * halo_w)
    bounds = slab_bounds(x, origin, block, int(ncell[0]), world)
    cx0, cx1 = bounds[rank], bounds[rank + 1]
    scenes = grid_partition_numpy(x, origin, block, halo_w, ncell, cx0, cx1)
    cx_of = np.clip(np.floor((x[:, 0] - origin[0]) / np.float32(block)).astype(np.int64), 0, ncell[0] - 1)
    # what k_pile_halo_map produces: (scene, slot, gid) of the halo copies owned by the left / right neighbour
    left, right, owned_at = [], [], {}
    for s, keys in enumerate(scenes):
        for slot, key in enumerate(keys):
            g = int(key) >> 1
            if not (key & 1):
                owned_at[g] = (s, slot)
            elif cx_of[g] < cx0:
                left.append((s, slot, g))
            elif cx_of[g] >= cx1:
                right.append((s, slot, g))
    req = torch.tensor([g for _, _, g in left + right] or [0], dtype=torch.int32)
    gids, ask, give = exchange_requests(dist, rank, world, req, len(left), len(right))
    gids = gids[: sum(give)].numpy()
    assert all(int(g) in owned_at for g in gids)                         # k_pile_lookup would flag anything else
    # per-step refresh: pack my owners' states in the asked order, exchange, unpack in my request order
    state = np.arange(n * STATE_W, dtype=np.float32).reshape(n, STATE_W) + 0.5 * rank      # "owner's state" = f(gid, owner rank)
    sbuf = torch.from_numpy(np.ascontiguousarray(state[gids]).reshape(-1) if gids.size else np.zeros(0, np.float32))
    rbuf = torch.zeros((len(left) + len(right)) * STATE_W, dtype=torch.float32)
    dist.all_to_all_single(rbuf, sbuf, [a * STATE_W for a in ask], [g * STATE_W for g in give])
    got = rbuf.numpy().reshape(-1, STATE_W)
    ok = True
    for k, (_, _, g) in enumerate(left + right):
        owner = rank - 1 if k < len(left) else rank + 1
        ok &= bool(np.array_equal(got[k], np.arange(g * STATE_W, (g + 1) * STATE_W, dtype=np.float32) + 0.5 * owner))
    ret[rank] = (ok, len(left), len(right))
    dist.barrier()
    dist.destroy_process_group()


def test_device_repartition_protocol_over_gloo_three_ranks():
    ret = mp.Manager().dict()
    mp.spawn(_device_protocol_worker, args=(3, _free_port(), ret), nprocs=3, join=True)
    assert all(ret[r][0] for r in range(3))
    assert ret[0][1] == 0 and ret[2][2] == 0 and ret[1][1] > 0 and ret[1][2] > 0      # the middle rank has both neighbours
