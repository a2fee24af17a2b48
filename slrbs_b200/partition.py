"""One large sphere pile cut into overlapping blocks and spread over the GPUs of a box (SURVEY.md 8(e),
BASELINE.json config 5). No reference counterpart: the reference is one process with an O(n^2) pair loop.

Space is cut into cubic blocks. Every block is ONE SCENE of the batched pipeline and holds
  * the bodies whose centre lies in the block ("owned"),
  * copies of the bodies of the 26 neighbouring blocks that can reach an owned body before the next
    re-partition ("halo": distance to the block <= 2 r_max + margin + 2 v_max K dt), simulated as ordinary
    dynamic bodies so that contacts across a block face are solved on both sides,
  * the fixed container bodies.
Blocks are dealt to the ranks as contiguous slabs of block-x indices balanced by body count. After every
step each halo copy is overwritten with its owner's new state: on the same GPU by an indexed device copy
(slrbs_halo_sync_local), across GPUs by slrbs_halo_pack -> NCCL all_to_all_single over NVLink ->
slrbs_halo_unpack. Every K steps the owners' state is all-reduced, the partition rebuilt (bodies migrate) and
the scenes re-uploaded. Inside a block the solver keeps the reference's sequential row order (bodies sorted by
global index); ACROSS blocks the coupling is block-Jacobi with one layer of overlap, so results are validated
on penetration / residual against a single-scene run, not on trajectories (north_star).

The partition itself (build_partition) is pure numpy and identical on every rank, so ranks need not
exchange index lists: each derives its own send / receive lists from the same global picture.
"""
import os

import numpy as np

from .capi import Context, SlrbsError, SPHERE

STATE_W = 13      # x3 q4 v3 w3


def build_partition(x, block, halo_w, world):
    """Partition N points. Returns a dict with, for every rank, its scenes (arrays of global body ids, sorted,
    and an owned mask) and, for every body, where it is owned (rank, scene, slot)."""
    x = np.asarray(x, np.float64)
    n = x.shape[0]
    assert halo_w <= block, "halo wider than a block"
    origin = x.min(axis=0) - 1e-3
    cell = np.floor((x - origin) / block).astype(np.int64)
    ncell = cell.max(axis=0) + 1
    # slabs of block-x indices balanced by body count
    per_x = np.bincount(cell[:, 0], minlength=ncell[0])
    before = np.cumsum(per_x) - per_x
    rank_of_cx = np.minimum(((2 * before + per_x) * world) // max(2 * int(per_x.sum()), 1), world - 1).astype(np.int64)
    rank_of_cx = np.maximum.accumulate(rank_of_cx)
    key = (cell[:, 0] * ncell[1] + cell[:, 1]) * ncell[2] + cell[:, 2]
    ukeys = np.unique(key)                                         # existing blocks, lexicographic
    block_rank = rank_of_cx[ukeys // (ncell[1] * ncell[2])]
    # scene index of a block within its rank
    scene_of_block = np.zeros(len(ukeys), np.int64)
    for r in range(world):
        m = block_rank == r
        scene_of_block[m] = np.arange(m.sum())
    bidx = np.searchsorted(ukeys, key)                             # block index of every body
    # halo candidates: (body, neighbour block) pairs. A body is copied into the neighbour at offset d when, on every
    # axis where d is non-zero, it lies within halo_w of that face (a box-shaped, slightly generous neighbourhood:
    # cheaper than the exact distance to the neighbouring cube and still a superset of what can touch)
    rel = x - origin - cell * block                                # position inside the own cell, in [0, block)
    near_lo, near_hi = rel <= halo_w, (block - rel) <= halo_w      # [n, 3]
    border = np.nonzero((near_lo | near_hi).any(axis=1))[0]        # interior bodies are nobody's halo
    nlo, nhi, bcell = near_lo[border], near_hi[border], cell[border]
    pairs_body, pairs_block = [np.arange(n)], [bidx]
    owned_flag = [np.ones(n, bool)]
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                m = np.ones(len(border), bool)
                for ax, dd in enumerate((dx, dy, dz)):
                    if dd < 0: m &= nlo[:, ax]
                    elif dd > 0: m &= nhi[:, ax]
                if not m.any():
                    continue
                nc = bcell[m] + np.array([dx, dy, dz])
                ok = ((nc >= 0) & (nc < ncell)).all(axis=1)
                ids = border[m][ok]
                nk = (nc[ok, 0] * ncell[1] + nc[ok, 1]) * ncell[2] + nc[ok, 2]
                pos = np.minimum(np.searchsorted(ukeys, nk), len(ukeys) - 1)
                exists = ukeys[pos] == nk
                pairs_body.append(ids[exists]); pairs_block.append(pos[exists]); owned_flag.append(np.zeros(int(exists.sum()), bool))
    pb, pk, po = np.concatenate(pairs_body), np.concatenate(pairs_block), np.concatenate(owned_flag)
    order = np.argsort(pk.astype(np.int64) * n + pb, kind="stable")      # by block, then global body id
    pb, pk, po = pb[order], pk[order], po[order]
    starts = np.searchsorted(pk, np.arange(len(ukeys)))
    ends = np.searchsorted(pk, np.arange(len(ukeys)), side="right")
    slot = np.arange(len(pk)) - starts[pk]
    owner_rank = block_rank[bidx]
    owner_scene = scene_of_block[bidx]
    owner_slot = np.zeros(n, np.int64)
    owner_slot[pb[po]] = slot[po]
    ranks = []
    for r in range(world):
        blocks = np.nonzero(block_rank == r)[0]
        scenes = [dict(ids=pb[starts[b]:ends[b]], owned=po[starts[b]:ends[b]]) for b in blocks]
        ranks.append(scenes)
    return dict(ranks=ranks, owner_rank=owner_rank, owner_scene=owner_scene, owner_slot=owner_slot,
                max_bodies=int((ends - starts).max()), n_blocks=len(ukeys))


def halo_maps(part, rank, max_bodies):
    """Index lists of one rank: local (src_flat, dst_flat) copies and, per peer, what to send / where to unpack.
    Lists towards / from a peer are ordered by (receiving scene, receiving slot), which both sides can compute."""
    world = len(part["ranks"])
    loc_src, loc_dst = [], []
    recv = {p: [] for p in range(world)}          # flat destinations on this rank, by source peer
    for s, sc in enumerate(part["ranks"][rank]):
        halo = np.nonzero(~sc["owned"])[0]
        g = sc["ids"][halo]
        orank, oscene, oslot = part["owner_rank"][g], part["owner_scene"][g], part["owner_slot"][g]
        dst = s * max_bodies + halo
        m = orank == rank
        loc_src.append(oscene[m] * max_bodies + oslot[m]); loc_dst.append(dst[m])
        for p in np.unique(orank[~m]):
            recv[int(p)].append(dst[orank == p])
    send = {p: [] for p in range(world)}          # flat sources on this rank, by destination peer
    for p in range(world):
        if p == rank:
            continue
        for s, sc in enumerate(part["ranks"][p]):
            halo = np.nonzero(~sc["owned"])[0]
            g = sc["ids"][halo]
            m = part["owner_rank"][g] == rank
            send[p].append(part["owner_scene"][g[m]] * max_bodies + part["owner_slot"][g[m]])
    cat = lambda L: np.concatenate(L).astype(np.int32) if L else np.zeros(0, np.int32)
    return dict(local_src=cat(loc_src), local_dst=cat(loc_dst),
                send={p: cat(v) for p, v in send.items()}, recv={p: cat(v) for p, v in recv.items()})


class PartitionedPile:
    """Driver of one rank. `comm` is None (single process, world 1) or a torch.distributed-like object used through
    all_reduce_sum(tensor) and all_to_all(out, inp, out_splits, in_splits) (see TorchComm)."""

    def __init__(self, x, radius, mass, fixed_bodies, block=8.0, rebuild_every=0, vmax=6.0, dt=0.01, rank=0, world=1,
                 device=0, comm=None, contact_factor=4, mu=0.8, iters=10, slack=0.5):
        """rebuild_every = K > 0 fixes the re-partition interval and sizes the halo for speeds up to `vmax`;
        K = 0 (default) adapts it: the halo gets `slack` metres of room and the next re-partition is scheduled for
        when the fastest body (current speed plus free-fall acceleration) could have used half of it."""
        self.n = x.shape[0]
        self.radius = np.asarray(radius, np.float32); self.mass = np.asarray(mass, np.float32)
        self.fixed = fixed_bodies                                 # list of dict(x, q, dim)
        self.block, self.K, self.vmax, self.dt = float(block), int(rebuild_every), float(vmax), float(dt)
        self.rank, self.world, self.device, self.comm = rank, world, device, comm
        self.contact_factor, self.mu, self.iters = contact_factor, mu, iters
        self.adaptive, self.slack = self.K <= 0, float(slack)
        if self.adaptive:
            self.halo_w = 2.0 * float(self.radius.max()) + 2e-3 + self.slack
            self.K = 1
        else:
            self.halo_w = 2.0 * float(self.radius.max()) + 2e-3 + 2.0 * self.vmax * self.K * self.dt
        self.state = np.zeros((self.n, STATE_W), np.float32)
        self.state[:, :3] = x
        self.state[:, 6] = 1.0
        self.ctx = None
        self.steps_since_rebuild = 0
        self.stats = dict(rebuilds=0, halo_bodies=0, owned_bodies=0, scenes=0, remote_halos=0)
        self.rebuild()

    # ---- partition + upload ------------------------------------------------------------------------------------
    def rebuild(self):
        import time
        t0 = time.perf_counter()
        part = build_partition(self.state[:, :3], self.block, min(self.halo_w, self.block), self.world)
        if part["n_blocks"] > 1 and self.halo_w > self.block:
            raise ValueError(f"halo width {self.halo_w:.2f} exceeds the block size {self.block}: rebuild more often or use larger blocks")
        nf = len(self.fixed)
        NBd = part["max_bodies"]
        NB = NBd + nf
        scenes = part["ranks"][self.rank]
        B = max(len(scenes), 1)
        if self.ctx is None or B != self.ctx.n_scenes or NB > self.ctx.max_bodies:
            if self.ctx is not None:
                self.ctx.close()
            NBcap = int(NB * 1.15) + 8
            self.ctx = Context(B, NBcap, 0, self.contact_factor * NBcap, device=self.device)
            self.ctx.set_params(self.mu, 0, self.iters, True)
        NBc = self.ctx.max_bodies
        xs = np.zeros((B, NBc, 3), np.float32); q = np.zeros((B, NBc, 4), np.float32); q[..., 3] = 1
        v = np.zeros((B, NBc, 3), np.float32); w = np.zeros((B, NBc, 3), np.float32)
        mass = np.ones((B, NBc), np.float32); fixed = np.zeros((B, NBc), np.int32); gtype = np.zeros((B, NBc), np.int32)
        geom = np.zeros((B, NBc, 6), np.float32); geom[..., 0] = 0.5
        ib = np.zeros((B, NBc, 9), np.float32); ibi = np.zeros((B, NBc, 9), np.float32)
        counts = np.zeros(B, np.int32)
        self.scene_ids, self.scene_owned = [], []
        for s, sc in enumerate(scenes):
            g = sc["ids"]; k = len(g)
            st = self.state[g]
            xs[s, :k] = st[:, 0:3]; q[s, :k] = st[:, 3:7]; v[s, :k] = st[:, 7:10]; w[s, :k] = st[:, 10:13]
            mass[s, :k] = self.mass[g]; geom[s, :k, 0] = self.radius[g]
            inertia = 0.4 * self.mass[g] * self.radius[g] ** 2    # Sphere::computeInertia, Geometry.h:41
            for d in (0, 4, 8):
                ib[s, :k, d] = inertia; ibi[s, :k, d] = 1.0 / inertia
            for f, fb in enumerate(self.fixed):                     # container boxes behind the dynamic bodies
                j = k + f
                xs[s, j] = fb["x"]; q[s, j] = fb["q"]; fixed[s, j] = 1; gtype[s, j] = 1; geom[s, j, :3] = fb["dim"]
                dim = np.asarray(fb["dim"], np.float32)
                diag = np.array([dim[1] ** 2 + dim[2] ** 2, dim[0] ** 2 + dim[2] ** 2, dim[0] ** 2 + dim[1] ** 2], np.float32) / 12.0
                ib[s, j, [0, 4, 8]] = diag; ibi[s, j, [0, 4, 8]] = 1.0 / diag
            counts[s] = k + nf
            self.scene_ids.append(g); self.scene_owned.append(sc["owned"])
        gtype[:] = np.where(fixed == 1, 1, SPHERE)
        self.ctx.upload_bodies(xs, q, v, w, mass, ib, ibi, fixed, gtype, geom, counts=counts)
        maps = halo_maps(part, self.rank, NBc)
        self.ctx.halo_set_local(maps["local_src"], maps["local_dst"])
        peers = [p for p in range(self.world) if p != self.rank]
        send = np.concatenate([maps["send"][p] for p in peers]) if peers else np.zeros(0, np.int32)
        recv = np.concatenate([maps["recv"][p] for p in peers]) if peers else np.zeros(0, np.int32)
        self.ctx.halo_set_remote(send, recv)
        self.send_splits = [int(maps["send"][p].size) if p != self.rank else 0 for p in range(self.world)]
        self.recv_splits = [int(maps["recv"][p].size) if p != self.rank else 0 for p in range(self.world)]
        if self.comm is not None:
            self.comm.prepare(sum(self.send_splits), sum(self.recv_splits))
        self.steps_since_rebuild = 0
        if self.adaptive:
            # two bodies may approach each other: each may use slack / 2 before the next re-partition.
            # displacement after k steps <= v0 k dt + g (k dt)^2 / 2  (v0 = fastest body now)
            v0 = float(np.sqrt((self.state[:, 7:10] ** 2).sum(axis=1).max())) if self.n else 0.0
            k = 1
            while k < 200 and v0 * (k + 1) * self.dt + 4.905 * ((k + 1) * self.dt) ** 2 <= 0.5 * self.slack:
                k += 1
            self.K = k
        owned = sum(int(o.sum()) for o in self.scene_owned)
        total = sum(len(g) for g in self.scene_ids)
        self.stats["rebuild_s"] = self.stats.get("rebuild_s", 0.0) + time.perf_counter() - t0
        self.stats["K"] = self.K
        self.stats.update(rebuilds=self.stats["rebuilds"] + 1, halo_bodies=total - owned, owned_bodies=owned,
                          scenes=len(scenes), remote_halos=sum(self.recv_splits), max_bodies=NBc)

    # ---- stepping ----------------------------------------------------------------------------------------------
    def step(self, n=1):
        for _ in range(n):
            if self.steps_since_rebuild >= self.K:
                self.gather_state()
                self.rebuild()
            self.ctx.step(self.dt, 1)
            self.ctx.halo_sync_local()
            if self.comm is not None and self.world > 1:
                self.ctx.halo_pack(self.comm.send_ptr())
                self.ctx.sync()                                    # the pack kernel runs on the context's stream
                self.comm.exchange(self.send_splits, self.recv_splits)   # NCCL all_to_all over NVLink
                self.ctx.halo_unpack(self.comm.recv_ptr())
            self.steps_since_rebuild += 1

    def gather_state(self):
        """Owners' state of ALL bodies on every rank (all-reduce of disjoint contributions)."""
        import time
        t0 = time.perf_counter()
        self.ctx.sync()
        st = self.ctx.download_state()
        full = np.zeros((self.n, STATE_W), np.float32)
        for s, (g, own) in enumerate(zip(self.scene_ids, self.scene_owned)):
            k = len(g)
            rows = np.concatenate([st["x"][s, :k], st["q"][s, :k], st["v"][s, :k], st["w"][s, :k]], axis=1)
            full[g[own]] = rows[own]
        if self.comm is not None and self.world > 1:
            full = self.comm.all_reduce_sum(full)
        self.state = full
        self.stats["gather_s"] = self.stats.get("gather_s", 0.0) + time.perf_counter() - t0
        return full

    def max_penetration(self):
        """Deepest contact over this rank's scenes (contacts between two halo copies are someone else's)."""
        self.ctx.sync()
        worst = 0.0
        for s in range(len(self.scene_ids)):
            c = self.ctx.contacts(s)
            if c["phi"].size:
                worst = max(worst, float(-c["phi"].min()))
        return worst

    def close(self):
        if self.ctx is not None:
            self.ctx.close(); self.ctx = None


# ---- device-side re-partition (fixed grid of blocks; slrbs_pile_* in the C ABI, csrc/pile.cu) ---------------------------
def fixed_grid(x, block, margin):
    """Origin and cell counts of the fixed block grid that covers the initial pile plus `margin` on every side
    (bodies that leave it are clamped into the boundary blocks)."""
    x = np.asarray(x, np.float32)
    lo = (x.min(axis=0) - np.float32(margin)).astype(np.float32)
    hi = (x.max(axis=0) + np.float32(margin)).astype(np.float32)
    ncell = np.maximum(np.ceil((hi - lo) / np.float32(block)).astype(np.int64), 1)
    return lo, ncell.astype(np.int32)


def slab_bounds(x, origin, block, ncx, world):
    """Contiguous slabs of block-x indices, balanced by body count, at least one block column per rank."""
    assert ncx >= world, f"{world} ranks need at least {world} block columns along x (have {ncx}): use smaller blocks"
    cx = np.clip(np.floor((np.asarray(x, np.float32)[:, 0] - origin[0]) / np.float32(block)).astype(np.int64), 0, ncx - 1)
    per_x = np.bincount(cx, minlength=ncx)
    cum = np.concatenate([[0], np.cumsum(per_x)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        b = int(np.searchsorted(cum, target, side="left"))
        b = max(b, bounds[-1] + 1)                       # at least one column for the previous rank
        b = min(b, ncx - (world - r))                    # and one for each rank still to come
        bounds.append(b)
    bounds.append(ncx)
    return bounds


def grid_partition_numpy(x, origin, block, halo_w, ncell, cx0, cx1):
    """Host mirror (float32, same expressions) of k_pile_classify + k_pile_sort: for every scene of the slab
    [cx0, cx1) the sorted keys global id * 2 + halo flag. Used by the tests to check the device partition."""
    x = np.asarray(x, np.float32)
    origin = np.asarray(origin, np.float32); block = np.float32(block); halo_w = np.float32(halo_w)
    ncell = np.asarray(ncell, np.int64)
    rel = x - origin
    cell = np.clip(np.floor(rel / block).astype(np.int64), 0, ncell - 1)
    inside = rel - cell.astype(np.float32) * block
    lo, hi = inside <= halo_w, (block - inside) <= halo_w
    nscenes = (cx1 - cx0) * int(ncell[1]) * int(ncell[2])
    scene_l, key_l = [], []
    gid = np.arange(x.shape[0], dtype=np.int64)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                m = np.ones(x.shape[0], bool)
                for ax, dd in enumerate((dx, dy, dz)):
                    if dd < 0: m &= lo[:, ax]
                    elif dd > 0: m &= hi[:, ax]
                nc = cell + np.array([dx, dy, dz])
                m &= (nc[:, 0] >= cx0) & (nc[:, 0] < cx1) & (nc[:, 1] >= 0) & (nc[:, 1] < ncell[1]) & (nc[:, 2] >= 0) & (nc[:, 2] < ncell[2])
                sc = ((nc[m, 0] - cx0) * ncell[1] + nc[m, 1]) * ncell[2] + nc[m, 2]
                scene_l.append(sc); key_l.append(gid[m] * 2 + (1 if (dx or dy or dz) else 0))
    sc, key = np.concatenate(scene_l), np.concatenate(key_l)
    order = np.lexsort((key, sc))
    sc, key = sc[order], key[order]
    starts = np.searchsorted(sc, np.arange(nscenes)); ends = np.searchsorted(sc, np.arange(nscenes), side="right")
    return [key[a:b].astype(np.int32) for a, b in zip(starts, ends)]


def contact_velocity_residual(ctx, n_scenes, owned=None):
    """Velocity-level constraint residual of the contacts a context currently holds: the largest approach speed
    max(0, -(v0 + w0 x r0 - v1 - w1 x r1) . n) over the contacts (0 for a converged non-penetration solve), and the
    number of contacts counted. `owned(scene)` -> boolean array over the scene's body slots: only contacts that touch
    an owned body count (contacts between two halo copies belong to other blocks, where they are counted)."""
    st = ctx.download_state()
    cnt = ctx.contact_counts()
    worst, total = 0.0, 0
    for s in np.nonzero(cnt[:n_scenes] > 0)[0]:
        c = ctx.contacts(int(s))
        b0, b1 = c["body0"], c["body1"]
        x, v, w = st["x"][s], st["v"][s], st["w"][s]
        r0, r1 = c["p"] - x[b0], c["p"] - x[b1]
        rel = (v[b0] + np.cross(w[b0], r0)) - (v[b1] + np.cross(w[b1], r1))
        vn = np.einsum("ij,ij->i", rel, c["n"])
        if owned is not None:
            own = owned(int(s))
            keep = own[b0] | own[b1]
            if not keep.any():
                continue
            vn = vn[keep]
        worst = max(worst, float(np.maximum(0.0, -vn).max()))
        total += int(vn.size)
    return worst, total


def exchange_requests(dist, rank, world, req, n_left, n_right):
    """Request protocol of the device-side re-partition, for slabs along x (a rank's only neighbours are rank - 1 and
    rank + 1). `req` holds the global ids of the halo copies this rank needs refreshed, the left neighbour's first
    (n_left) and then the right neighbour's (n_right), in the order the rank will unpack them. Every rank tells its
    neighbours how many it asks for, then the id lists travel (all_to_all_single). Returns (gids, ask, give): the ids
    the neighbours asked of this rank, ordered by asking rank (so the left neighbour's first), and the per-rank
    counts asked / given, which are also the split sizes of the per-step state exchange. Works on CUDA tensors over
    NCCL and on CPU tensors over gloo (tests/test_partition_cpu.py)."""
    import torch
    assert (n_left == 0 or rank > 0) and (n_right == 0 or rank < world - 1), "a halo copy's owner must be an adjacent slab"
    ask = [0] * world
    if rank > 0: ask[rank - 1] = int(n_left)
    if rank < world - 1: ask[rank + 1] = int(n_right)
    ask_t = torch.tensor(ask, dtype=torch.int32, device=req.device); give_t = torch.empty_like(ask_t)
    dist.all_to_all_single(give_t, ask_t)
    give = [int(v) for v in give_t.cpu()]
    gids = torch.empty(max(sum(give), 1), dtype=torch.int32, device=req.device)
    dist.all_to_all_single(gids[: sum(give)], req[: n_left + n_right], give, ask)
    return gids, ask, give


class DevicePile:
    """One rank of a pile whose re-partition runs on the device (no host round trip of the body states).

    Space is a fixed grid of blocks; the rank owns a slab of block columns, one scene per block. Every K steps the
    owners' states are scattered into a dense device table (slrbs_pile_export), summed over the ranks with an NCCL
    all-reduce, and every rank rebuilds its scenes from the table on its GPU (slrbs_pile_rebuild). Halo copies owned
    by a neighbour rank are requested from it once per rebuild (an all-to-all of global ids) and refreshed after every
    step by slrbs_halo_pack -> NCCL all_to_all_single -> slrbs_halo_unpack; halo copies owned locally by an indexed
    device copy. K follows the fastest body so that nothing can cross the halo between two rebuilds."""

    def __init__(self, x, radius, mass, fixed_bodies, block=8.0, dt=0.01, rank=0, world=1, device=0, slack=0.5,
                 contact_factor=4, mu=0.8, iters=10, rebuild_every=0, max_bodies=0, dist=None, native=None):
        """dist: the collectives (default torch.distributed; tests pass an EmulatedGroup handle to run several ranks on one
        GPU). native: a NativeHaloComm (libslrbs_pile.so) -- the per-step halo refresh then runs pack -> ncclSend/ncclRecv ->
        unpack on the context's own stream with no host synchronisation."""
        import torch
        self.torch = torch
        self.dist = dist
        self.native = native
        if world > 1 and dist is None:
            import torch.distributed as tdist
            self.dist = tdist
        self.n = int(x.shape[0])
        self.radius = np.ascontiguousarray(radius, np.float32); self.mass = np.ascontiguousarray(mass, np.float32)
        self.block, self.dt, self.slack = float(block), float(dt), float(slack)
        self.rank, self.world, self.device = rank, world, device
        self.contact_factor, self.mu, self.iters = contact_factor, mu, iters
        self.fixed_K = int(rebuild_every)
        self.halo_w = min(2.0 * float(self.radius.max()) + 2e-3 + self.slack, self.block)
        self.fixed10 = np.array([list(f["x"]) + list(f["q"]) + list(f["dim"]) for f in fixed_bodies], np.float32).reshape(-1, 10)
        self.origin, self.ncell = fixed_grid(x, self.block, 2.0 * self.halo_w)
        self.bounds = slab_bounds(x, self.origin, self.block, int(self.ncell[0]), world)
        self.cx0, self.cx1 = self.bounds[rank], self.bounds[rank + 1]
        self.B = (self.cx1 - self.cx0) * int(self.ncell[1]) * int(self.ncell[2])
        dev = torch.device("cuda", device)
        state = np.zeros((self.n, STATE_W), np.float32)
        state[:, :3] = x; state[:, 6] = 1.0
        self.table = torch.from_numpy(state).to(dev).reshape(-1)
        self.vmax2 = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ctx = None
        nf = self.fixed10.shape[0]
        if max_bodies <= 0:                                        # uniform-density guess; grown on overflow
            cells = float(np.prod(self.ncell.astype(np.float64)))
            grow = ((self.block + 2.0 * self.halo_w) / self.block) ** 3
            max_bodies = int(1.5 * self.n / cells * grow) + nf + 64
        self.NB = max_bodies
        self.stats = dict(rebuilds=0, rebuild_s=0.0, grows=0)
        self.steps_since_rebuild, self.K = 0, 1
        self.sbuf = self.rbuf = None
        self._rebuild(first=True)

    # ---- context + rebuild --------------------------------------------------------------------------------------
    def _make_ctx(self):
        if self.ctx is not None:
            self.ctx.close()
        self.ctx = Context(self.B, self.NB, 0, self.contact_factor * self.NB, device=self.device)
        self.ctx.set_params(self.mu, 0, self.iters, True)
        self.ctx.pile_init(self.radius, self.mass, self.origin, self.block, self.halo_w, self.ncell, self.cx0, self.cx1, self.fixed10)

    def _rebuild(self, first=False):
        import time
        torch = self.torch
        grow = 0
        if self.ctx is not None and not first:
            # the steps since the last re-partition are still in flight (nothing synchronises them any more): wait for them
            # here so that the rebuild's own time is what gets counted below
            try:
                self.ctx.sync()
            except SlrbsError as e:
                # a block produced more contacts than its scene holds since the last re-partition (they were dropped for
                # those steps): the states are complete all the same; rebuild into a context with more contact slots
                if "more contacts" not in str(e):
                    raise
                grow = 1
        t0 = time.perf_counter()
        if self.ctx is None:
            self._make_ctx()
        if not first:
            self.table.zero_(); self.vmax2.zero_()
            torch.cuda.synchronize(self.device)
            self.ctx.pile_export(self.table.data_ptr(), self.vmax2.data_ptr())
            self.ctx.sync()
            if self.world > 1:
                self.dist.all_reduce(self.table)                                   # disjoint owners: sum = gather
                self.dist.all_reduce(self.vmax2, op=self.dist.ReduceOp.MAX)        # float bits of v^2 >= 0 order like ints
                g = torch.tensor([grow], dtype=torch.int32, device=self.table.device)
                self.dist.all_reduce(g, op=self.dist.ReduceOp.MAX)                 # every rank keeps the same capacity
                grow = int(g.cpu()[0])
                torch.cuda.synchronize(self.device)
            if grow:
                self.contact_factor = int(self.contact_factor * 1.5) + 1
                self.stats["contact_grows"] = self.stats.get("contact_grows", 0) + 1
                self._make_ctx()
        while True:
            try:
                info = self.ctx.pile_rebuild(self.table.data_ptr())
                break
            except SlrbsError:
                info = self.ctx.pile_info
                if not info[0]:
                    raise
                self.NB = int(info[1] * 1.25) + self.fixed10.shape[0] + 64         # a block outgrew the scenes: larger context
                self.stats["grows"] += 1
                self._make_ctx()
        self.info = info
        nl, nr = int(info[3]), int(info[4])
        if self.world > 1:
            dev = self.table.device
            req = torch.empty(max(nl + nr, 1), dtype=torch.int32, device=dev)
            self.ctx.pile_requests(req.data_ptr())
            gids, ask, give = exchange_requests(self.dist, self.rank, self.world, req, nl, nr)
            torch.cuda.synchronize(self.device)
            self.ctx.pile_set_send(gids.data_ptr(), sum(give))
            self.ask13, self.give13 = [a * STATE_W for a in ask], [g * STATE_W for g in give]
            self.sbuf = torch.empty(max(sum(give), 1) * STATE_W, dtype=torch.float32, device=dev)
            self.rbuf = torch.empty(max(nl + nr, 1) * STATE_W, dtype=torch.float32, device=dev)
            self.n_give, self.n_ask = sum(give), nl + nr
            # what goes to / comes from the left and the right neighbour (slabs along x), for the native exchange
            self.give_lr = (give[self.rank - 1] if self.rank > 0 else 0, give[self.rank + 1] if self.rank < self.world - 1 else 0)
            self.ask_lr = (nl, nr)
        if self.fixed_K > 0:
            self.K = self.fixed_K
        else:
            # two bodies may approach each other: each may use slack / 2 before the next rebuild;
            # displacement after k steps <= v0 k dt + g (k dt)^2 / 2  (v0 = fastest body now)
            v2 = float(np.array([int(self.vmax2.cpu()[0])], np.int32).view(np.float32)[0]) if not first else 0.0
            v0 = float(np.sqrt(max(v2, 0.0)))
            k = 1
            while k < 200 and v0 * (k + 1) * self.dt + 4.905 * ((k + 1) * self.dt) ** 2 <= 0.5 * self.slack:
                k += 1
            self.K = k
        self.steps_since_rebuild = 0
        self.stats["rebuilds"] += 1
        self.stats["rebuild_s"] += time.perf_counter() - t0
        if not first:
            self.stats["steady_rebuild_s"] = self.stats.get("steady_rebuild_s", 0.0) + time.perf_counter() - t0
        self.stats.update(K=self.K, scenes=self.B, max_bodies=self.NB, largest_scene=int(info[1]), local_halos=int(info[2]),
                          remote_halos=nl + nr, slots=int(info[5]), owned=int(info[6]))

    # ---- stepping ------------------------------------------------------------------------------------------------
    def step(self, n=1):
        for _ in range(n):
            if self.steps_since_rebuild >= self.K:
                self._rebuild()
            self.ctx.step(self.dt, 1)
            self.ctx.halo_sync_local()
            if self.world > 1 and self.native is not None:
                # stream-ordered: pack -> grouped ncclSend / ncclRecv with the two neighbours -> unpack, all on the context's stream
                self.native.halo_exchange(self.ctx, self.sbuf.data_ptr(), self.rbuf.data_ptr(), self.give_lr, self.ask_lr)
            elif self.world > 1:
                self.ctx.halo_pack(self.sbuf.data_ptr())
                self.ctx.sync()
                self.dist.all_to_all_single(self.rbuf[: self.n_ask * STATE_W], self.sbuf[: self.n_give * STATE_W], self.ask13, self.give13)
                self.torch.cuda.synchronize(self.device)
                self.ctx.halo_unpack(self.rbuf.data_ptr())
            self.steps_since_rebuild += 1

    def gather_state(self):
        """[n, 13] states of ALL bodies (owners' values) on every rank."""
        torch = self.torch
        self.table.zero_(); self.vmax2.zero_()
        torch.cuda.synchronize(self.device)
        self.ctx.pile_export(self.table.data_ptr(), self.vmax2.data_ptr())
        self.ctx.sync()
        if self.world > 1:
            self.dist.all_reduce(self.table)
            torch.cuda.synchronize(self.device)
        return self.table.cpu().numpy().reshape(self.n, STATE_W).copy()

    def max_penetration(self):
        self.ctx.sync()
        cnt = self.ctx.contact_counts()
        worst = 0.0
        for s in np.nonzero(cnt > 0)[0]:
            c = self.ctx.contacts(int(s))
            worst = max(worst, float(-c["phi"].min()))
        return worst

    def velocity_residual(self):
        """Residual over the contacts that touch a body this rank owns (halo-halo contacts are other blocks' business;
        the halo copies carry their owners' post-step velocities, so owned-halo contacts measure the coupling error)."""
        self.ctx.sync()

        def owned(s):
            keys = self.ctx.pile_scene_keys(s)
            m = np.zeros(self.ctx.max_bodies, bool)
            m[: keys.size] = (keys & 1) == 0
            return m
        return contact_velocity_residual(self.ctx, self.B, owned)

    def close(self):
        if self.ctx is not None:
            self.ctx.close(); self.ctx = None


class EmulatedGroup:
    """Several ranks of a DevicePile on ONE GPU in ONE process, one Python thread per rank: the subset of
    torch.distributed the pile uses (all_reduce, all_to_all_single), done with device copies between the ranks' tensors
    at thread barriers. Lets the whole multi-rank path -- exported state table, requested-id protocol, packed halo
    refresh, rebuild -- run and be checked on a one-GPU box; NCCL itself refuses two ranks on one device."""

    class ReduceOp:
        SUM, MAX = 0, 1

    def __init__(self, world):
        import threading
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.result = None

    def handle(self, rank):
        return _EmulatedDist(self, rank)


class _EmulatedDist:
    def __init__(self, group, rank):
        self.g, self.rank, self.ReduceOp = group, rank, EmulatedGroup.ReduceOp

    def all_reduce(self, t, op=EmulatedGroup.ReduceOp.SUM):
        import torch
        g = self.g
        torch.cuda.synchronize()
        g.slots[self.rank] = t
        g.barrier.wait()
        if self.rank == 0:
            acc = g.slots[0].clone()
            for other in g.slots[1:]:
                acc = torch.maximum(acc, other) if op == EmulatedGroup.ReduceOp.MAX else acc + other
            g.result = acc
            torch.cuda.synchronize()
        g.barrier.wait()
        t.copy_(g.result)
        torch.cuda.synchronize()
        g.barrier.wait()

    def all_to_all_single(self, out, inp, out_splits=None, in_splits=None):
        import torch
        g, w = self.g, self.g.world
        if in_splits is None:
            in_splits = [inp.numel() // w] * w
        if out_splits is None:
            out_splits = [out.numel() // w] * w
        torch.cuda.synchronize()
        g.slots[self.rank] = (inp, list(in_splits))
        g.barrier.wait()
        o = 0
        for src in range(w):
            sin, ssp = g.slots[src]
            a = sum(ssp[: self.rank]); n = ssp[self.rank]
            assert n == out_splits[src], (self.rank, src, n, out_splits[src])
            if n:
                out[o:o + n].copy_(sin[a:a + n])
            o += n
        torch.cuda.synchronize()
        g.barrier.wait()


class NativeHaloComm:
    """libslrbs_pile.so: NCCL point-to-point between slab neighbours, issued on the simulation context's own stream.
    The 128-byte NCCL unique id is made on rank 0 and handed to the other ranks by the caller (any side channel;
    run_partition_multi.py broadcasts it with torch.distributed)."""

    def __init__(self, rank, world, device, unique_id):
        import ctypes as C
        from .capi import load_library
        load_library()
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libslrbs_pile.so")
        if not os.path.exists(path):
            raise SlrbsError(f"{path} is missing: build it with `make -C slrbs_b200/csrc pile`")
        self.L = C.CDLL(path)
        self.L.slrbs_pile_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_char_p]
        self.L.slrbs_pile_comm_halo_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        self.L.slrbs_pile_comm_destroy.argtypes = [C.c_void_p]
        self.L.slrbs_pile_comm_last_error.restype = C.c_char_p
        self.h = C.c_void_p()
        if self.L.slrbs_pile_comm_create(C.byref(self.h), rank, world, device, bytes(unique_id)) != 0:
            raise SlrbsError(self.L.slrbs_pile_comm_last_error().decode())

    @staticmethod
    def unique_id():
        import ctypes as C
        from .capi import load_library
        load_library()
        L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libslrbs_pile.so"))
        buf = C.create_string_buffer(128)
        if L.slrbs_pile_comm_unique_id(buf) != 0:
            raise SlrbsError("ncclGetUniqueId failed")
        return buf.raw

    def halo_exchange(self, ctx, send_ptr, recv_ptr, give_lr, ask_lr):
        import ctypes as C
        rc = self.L.slrbs_pile_comm_halo_exchange(self.h, ctx.h, C.c_void_p(send_ptr), C.c_void_p(recv_ptr),
                                                  int(give_lr[0]), int(give_lr[1]), int(ask_lr[0]), int(ask_lr[1]))
        if rc != 0:
            raise SlrbsError(self.L.slrbs_pile_comm_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.slrbs_pile_comm_destroy(self.h)
            self.h = None


class TorchComm:
    """NCCL plumbing through torch.distributed: device buffers for the packed halo state, all_to_all_single."""

    def __init__(self, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, torch.device("cuda", device)
        self.sbuf = self.rbuf = None

    def prepare(self, n_send, n_recv):
        self.sbuf = self.torch.zeros(max(n_send, 1) * STATE_W, dtype=self.torch.float32, device=self.device)
        self.rbuf = self.torch.zeros(max(n_recv, 1) * STATE_W, dtype=self.torch.float32, device=self.device)
        self.n_send, self.n_recv = n_send, n_recv

    def send_ptr(self): return self.sbuf.data_ptr()
    def recv_ptr(self): return self.rbuf.data_ptr()

    def exchange(self, send_splits, recv_splits):
        inp = self.sbuf[: self.n_send * STATE_W]
        out = self.rbuf[: self.n_recv * STATE_W]
        self.dist.all_to_all_single(out, inp, [r * STATE_W for r in recv_splits], [s * STATE_W for s in send_splits])
        self.torch.cuda.synchronize(self.device)

    def all_reduce_sum(self, arr):
        t = self.torch.from_numpy(arr).to(self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()


def sphere_pile(nx, ny, nz, radius=0.5, jitter=0.01, seed=1234):
    """nx x ny x nz lattice of touching spheres (spacing 2r) resting in an open box of five fixed boxes, the
    marble-box construction (Scenarios.h:18-80) scaled up; returns (x, radius, mass, fixed_bodies)."""
    rng = np.random.default_rng(seed)
    d = 2.0 * radius
    ii, kk, jj = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    x = np.stack([(ii - 0.5 * (nx - 1)) * d, 2.0 + kk * d, (jj - 0.5 * (nz - 1)) * d], axis=-1).reshape(-1, 3).astype(np.float32)
    x[:, [0, 2]] += rng.uniform(-jitter, jitter, (x.shape[0], 2)).astype(np.float32)
    ex, ez, hy = nx * d + d, nz * d + d, ny * d + 2.0
    s, c = np.sin(0.5 * 1.57), np.cos(0.5 * 1.57)
    qi, qy = [0, 0, 0, 1], [0, s, 0, c]
    fixed = [dict(x=[0.5 * ex - 0.25, 0.5 * hy, 0], q=qi, dim=[0.4, hy, ez]), dict(x=[-0.5 * ex + 0.25, 0.5 * hy, 0], q=qi, dim=[0.4, hy, ez]),
             dict(x=[0, 0.5 * hy, 0.5 * ez - 0.25], q=qy, dim=[0.4, hy, ex]), dict(x=[0, 0.5 * hy, -0.5 * ez + 0.25], q=qy, dim=[0.4, hy, ex]),
             dict(x=[0, 0, 0], q=qi, dim=[ex, 0.4, ez])]
    n = x.shape[0]
    return x, np.full(n, radius, np.float32), np.ones(n, np.float32), fixed
