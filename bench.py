#!/usr/bin/env python
"""bench.py -- body-steps/sec of the per-step rigid-body pipeline on N B200s (one process per GPU).

A "step" is one RigidBodySystem::step(dt) of EVERY scene in the batch (slrbs_step(ctx, dt, 1)):
detection, Jacobian assembly, 10 Box-PGS sweeps, integration. Scenes are independent, so ranks
shard the batch by scene index with no data-path collective (weak scaling: fixed scenes per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload marble_box|car|swinging_boxes|sphere_stack|marble_box_1k|box_stack] [--scenes B]

Prints ONE JSON line (rank 0). Keys beyond the base contract:
  roofline      the PGS sweep kernel (dominant): algorithmic bytes (464 B / contact visit, 744 B / hinge
                visit) / its CUDA-event duration, against the measured HBM copy peak
  cpu_baseline  the reference's own sources (oracle/_ref: compiled unmodified against the Eigen API shim,
                kind "reference"; the C restatement, kind "port", only if that build is absent) timed
                single-threaded on this box's host CPU
  e2e           the same metric through the C ABI with HOST buffers: every step uploads x,q,v,omega
                of every body from pinned host memory and downloads them again
The --impl reference arm times the same CPU implementation on all host cores, one independent scene per
process (the reference is single-threaded and keeps process-wide solver state).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 0.01           # SimViewer.cpp:113
SOLVER_ITERS = 10   # RigidBodySystem.cpp:22
MU = 0.8            # Contact.cpp:4
BYTES_PER_CONTACT_VISIT = 464
BYTES_PER_HINGE_VISIT = 744
E2E_CHUNKS = int(os.environ.get("SLRBS_E2E_CHUNKS", "4"))                  # contexts the batch is split into for the end-to-end arm (transfer / compute overlap)

# --workload mixed_pile: BASELINE configs[3], 10 002 mixed spheres / boxes / closed meshes in ONE scene (slrbs_b200/scenes.py):
# extension contact pairs, SDF meshes, pair-parallel detection, graph-coloured solver; handled by build_batch()
MIXED_PILE_BODIES = 10002
# workload -> (Scenarios builder, (a,b,c), contact capacity / scene, scenes / GPU, settle steps, jitter)
WORKLOADS = {
    # the reference scene itself, 167 bodies; 49152 scenes = 1536 warps of the thread-per-scene solver, 10.4 per SM (12 fit;
    # 32768 scenes leave it at 6.9 warps per SM: 3.79e8 vs 4.06e8 body-steps/s; beyond 53248 a second, ragged wave starts)
    "marble_box":     ("marble_box", (0, 0, 0), 768, 49152, 150, 0.01),
    "car":            ("car", (0, 0, 0), 16, 262144, 50, 0.0),
    "swinging_boxes": ("swinging_boxes", (0, 0, 0), 0, 65536, 50, 0.0),
    "sphere_stack":   ("sphere_stack", (8, 0, 0), 32, 262144, 50, 0.0),
    # BASELINE configs[1]: ~1k spheres in a box. 2664 scenes = 2 x (9 scenes per SM x 148 SMs): with the packed 24 B
    # accumulators 9 warp-scenes fit an SM, so this batch is exactly two full waves of the level-scheduled solver
    "marble_box_1k":  ("marble_box_n", (10, 10, 10), 4096, 2664, 200, 0.01),
    # BASELINE configs[0] "box-stack on plane": NOT a reference scene and needs box-box / box-plane contacts, which the
    # reference does not have -> the extension pairs (slrbs_set_extensions), no reference parity for this workload
    "box_stack":      ("box_stack", (5, 0, 0), 32, 262144, 50, 0.0),
    # ONE scene no SM can hold (9605 bodies, ~28 000 touching pairs): pair-parallel detection over the whole GPU and the
    # graph-coloured solver (pgs_color.cu); order differs from the reference, validated on the residual (tests/test_gpu_color.py)
    "big_scene":      ("marble_box_n", (20, 20, 24), 40000, 1, 100, 0.01),
    "mixed_pile":     ("mixed_pile", (MIXED_PILE_BODIES, 0, 0), 60000, 1, 150, 0.0),
}
# --workload pile_1m / pile: one pile cut into blocks over the ranks (slrbs_b200/partition.py), handled by run_pile()
PILE_DIMS = {"pile_1m": (100, 100, 100), "pile": (48, 24, 48)}



SOLVER_SOURCES = ("pgs_pipe.cu", "pgs_stage.cu", "pgs_gacc.cu", "pgs_color.cu", "pgs_split.cu", "pgs_level.cu", "pgs_level.cuh", "pgs_math.cuh", "layout.cuh", "kernels.cu")


def solver_source_hash():
    """Hash of the sources that decide the solver kernels' memory traffic: a measured-DRAM record in
    profiles/pgs_traffic.json only counts while it was taken on the same kernels."""
    import hashlib
    h = hashlib.sha256()
    for f in SOLVER_SOURCES:
        with open(os.path.join(ROOT, "slrbs_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def common_config(workload, B, nb, nj):
    """The workload description BOTH arms print (identical keys and values)."""
    scene, abc, nc, _, settle, jitter = WORKLOADS[workload]
    return {"workload": f"{workload} x {B} scenes/GPU", "scene": scene, "scene_args": list(abc), "scenes_per_gpu": B,
            "bodies_per_scene": nb, "joints_per_scene": nj, "dt": DT, "solver": "Box-PGS", "solver_iters": SOLVER_ITERS, "mu": MU,
            "settle_steps": settle, "jitter": jitter,
            "l2": "no flush: the per-step working set (constraint rows of all scenes) is larger than the 126 MB L2"
                  if workload in ("marble_box", "marble_box_1k") else "no flush: every step re-detects and rebuilds all rows; inputs of a step are the previous step's outputs"}


def bind_rank_to_cores(local, world):
    """One contiguous share of the host's cores per rank (its own threads, its pinned buffers' first touch): 8 ranks of
    one box otherwise float over the same cores. Uses the GPU's local CPU list when the platform exposes one."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        share = max(1, len(cpus) // max(world, 1))
        mine = cpus[local * share:(local + 1) * share] or cpus
        os.sched_setaffinity(0, mine)
        return f"cores {mine[0]}-{mine[-1]}"
    except Exception as e:       # not fatal: binding is an optimisation
        return f"unbound ({e})"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region. The sampler is started ahead of the warm-up
    (nvidia-smi needs ~0.1 s before its first line) and only the lines that arrive between begin() and end() count; a timed
    region shorter than the sampling interval is followed by untimed steps of the same load until two samples exist, and
    the result says so."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.t0, self.t1 = [], None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def _inside(self):
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        return [ln for t, ln in self.lines if t0 <= t <= t1]

    def stop(self, load=None):
        """load: callable that keeps the GPU under the timed region's load for a moment (used when the region was too short)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.end()
        note = None
        if len(self._inside()) < 2 and load is not None:
            t_end = time.perf_counter() + 2.0
            while len(self._inside()) < 2 and time.perf_counter() < t_end:
                load()
                self.t1 = time.perf_counter()
            note = "timed region shorter than the sampling interval: sampled under the same load right after it"
        else:
            time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        lines = self._inside()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def build_batch(workload, B, rank):
    """Synthetic batch: this rank's contiguous block of B scenes (global indices rank*B ..), each a copy of
    a Scenarios.h scene whose dynamic bodies are jittered in x,z by a generator keyed on the GLOBAL scene
    index (the data does not depend on the number of ranks)."""
    from slrbs_b200 import host_api, sharding
    scene, abc, nc, _, settle, jitter = WORKLOADS[workload]
    if scene == "mixed_pile":
        from slrbs_b200 import scenes
        mv, mf = scenes.blob()
        s = scenes.mixed_pile(abc[0], mv)
        s["mesh"] = (mv, mf)
    else:
        s = host_api.scene_arrays(scene, *abc)
    nb = s["x"].shape[0]
    x = np.broadcast_to(s["x"], (B, nb, 3)).copy()
    if jitter > 0:
        dyn = s["fixed"] == 0
        j = sharding.jitter_for_scenes(rank * B, B, nb, jitter)
        x[:, dyn, 0] += j[:, dyn, 0]
        x[:, dyn, 2] += j[:, dyn, 1]
    return s, x, nb, s["jtype"].shape[0], nc, settle


def run_ours(args):
    import torch
    import torch.distributed as dist
    from slrbs_b200 import Context

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    binding = bind_rank_to_cores(local, world)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.scenes or WORKLOADS[args.workload][3]
    s, x, nb, nj, nc, settle = build_batch(args.workload, B, rank)

    ctx = Context(B, nb, nj, nc, device=local)
    ctx.set_params(MU, 0, SOLVER_ITERS, True)
    if args.workload in ("box_stack", "mixed_pile"):
        ctx.set_extensions(1)
    if "mesh" in s:
        from slrbs_b200 import scenes
        sh = scenes.sdf_shape(*s["mesh"], device=local)
        ctx.add_sdf_shape(sh["dims"], sh["origin"], sh["cell"], sh["grid"], sh["verts"])
    variants = ctx.describe()
    ctx.upload_bodies(x, s["q"], s["v"], s["w"], s["mass"], s["ibody"], s["ibodyinv"], s["fixed"], s["geom_type"], s["geom"])
    if nj:
        ctx.upload_joints(s["jtype"], s["jb0"], s["jb1"], s["r0"], s["q0"], s["r1"], s["q1"])
    del x
    ctx.step(DT, settle); ctx.sync()                       # untimed: let the scene reach its contact-dense state
    sampler = ClockSampler(local) if rank == 0 else None   # started ahead of the warm-up; only lines inside the timed region count
    ctx.step(DT, args.warmup); ctx.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region (CUDA-graph replay, the path slrbs_step takes for a caller) ------
    ctx.profile_read(True)                                 # zero the launch counter
    barrier()
    if sampler:
        sampler.begin()
    ctx.timer_start()
    ctx.step(DT, args.steps)
    ms = ctx.timer_stop()
    barrier()
    if sampler:
        sampler.end()
    launches_timed = int(ctx.profile_read(True)["total_launches"])       # kernels launched (graph nodes replayed) in the timed region

    def same_load():
        ctx.step(DT, 10); ctx.sync()
    clocks = sampler.stop(same_load) if sampler else None
    ctx.profile_read(True)
    # ---- separate profiled pass: CUDA events around every solver launch + visit counters (profiling replaces the graph by
    # plain launches, so it is kept out of the headline region) -------------------------------------
    psteps = max(3, min(args.steps, 10))
    ctx.profile_enable(True); ctx.profile_read(True)
    ctx.timer_start()
    ctx.step(DT, psteps)
    prof_ms = ctx.timer_stop()
    prof = ctx.profile_read(True)
    ctx.profile_enable(False)
    ctx.sync()
    contacts_mean = float(ctx.contact_counts().mean())

    # ---- end-to-end through the C ABI with host buffers ------------------------------------------
    # The scenes are split over E2E_CHUNKS contexts (one stream each) so that one chunk's PCIe transfers overlap with
    # another chunk's step: every step still uploads x,q,v,omega of EVERY body from pinned host memory, advances every
    # scene once and downloads every body's state again before the step counts as done.
    # pinned buffers: cudaHostAlloc by this rank after its core binding, first touched here (fill_) by the bound thread
    hx, hq, hv, hw = [torch.empty((B, nb, k), dtype=torch.float32, pin_memory=True).fill_(0) for k in (3, 4, 3, 3)]
    ptr = [t.data_ptr() for t in (hx, hq, hv, hw)]
    ctx.download_state_raw(B, nb, *ptr)
    # what the host link gives this rank while ALL ranks copy at once (explains the end-to-end number at N > 1)
    pcie = {}
    dbuf = torch.empty_like(hx, device="cuda")
    for name, (dst, src) in (("h2d", (dbuf, hx)), ("d2h", (hx, dbuf))):
        dst.copy_(src, non_blocking=True); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        pcie[name] = 5 * hx.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    ctx.download_state_raw(B, nb, *ptr)
    del dbuf
    nch = E2E_CHUNKS if B % E2E_CHUNKS == 0 and B // E2E_CHUNKS >= 64 else 1
    if nch == 1:
        chunks = [(ctx, 0, B)]
    else:
        ctx.close(); ctx = None                             # free the big context's memory for the chunk contexts
        Bc = B // nch
        chunks = []
        for k in range(nch):
            c = Context(Bc, nb, nj, nc, device=local)
            c.set_params(MU, 0, SOLVER_ITERS, True)
            if args.workload in ("box_stack", "mixed_pile"):
                c.set_extensions(1)
            sl = slice(k * Bc, (k + 1) * Bc)
            c.upload_bodies(hx[sl].numpy(), hq[sl].numpy(), hv[sl].numpy(), hw[sl].numpy(), s["mass"], s["ibody"], s["ibodyinv"], s["fixed"],
                            s["geom_type"], s["geom"])
            if nj:
                c.upload_joints(s["jtype"], s["jb0"], s["jb1"], s["r0"], s["q0"], s["r1"], s["q1"])
            chunks.append((c, k * Bc, Bc))
    sizes = (3, 4, 3, 3)
    cptr = [[p + 4 * sz * nb * s0 for p, sz in zip(ptr, sizes)] for _, s0, _ in chunks]

    def e2e_run(nsteps):
        # per chunk and step: wait until the chunk's previous result is in host memory, upload its state again, step, download.
        # The wait is per chunk (not a barrier over all chunks), so chunk 0 starts step k + 1 while chunk 7 still finishes step k.
        for _ in range(nsteps):
            for (c, _, n_sc), pp in zip(chunks, cptr):
                c.sync()
                c.upload_state_raw_async(n_sc, nb, *pp)     # H2D: 13 floats / body
                c.step(DT, 1)
                c.download_state_raw_async(n_sc, nb, *pp)   # D2H: 13 floats / body
        for c, _, _ in chunks:
            c.sync()                                        # every body's new state is in host memory

    e2e_steps = max(3, min(args.steps, 10))
    e2e_run(3)                                              # warm the staging buffers (and capture the step graphs)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    # the reference viewer's per-frame traffic (SimViewer.cpp:40-45 reads every body's x and q; a PreStepFunc writes f and tau,
    # RigidBodySystem.cpp:70-73): forces up (24 B / body), poses down (28 B / body), the state itself stays on the device
    hf, ht = [torch.zeros((B, nb, 3), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    fptr = [[t.data_ptr() + 4 * 3 * nb * s0 for t in (hf, ht)] for _, s0, _ in chunks]

    def poses_run(nsteps):
        for _ in range(nsteps):
            for (c, _, n_sc), pp, fp in zip(chunks, cptr, fptr):
                c.set_external_forces_raw(n_sc, nb, fp[0], fp[1])      # H2D (waits for the chunk's previous download too)
                c.step(DT, 1)
                c.download_state_raw_async(n_sc, nb, pp[0], pp[1], 0, 0)   # D2H: x and q
        for c, _, _ in chunks:
            c.sync()

    poses_run(2)
    barrier()
    t0 = time.perf_counter()
    poses_run(e2e_steps)
    torch.cuda.synchronize()
    poses_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    for c, _, _ in chunks:
        if c is not ctx:
            c.close()

    from slrbs_b200 import sharding
    poses_ms = sharding.reduce_over_ranks([poses_ms], "max")[0]
    ms, e2e_ms, h2d_inv, d2h_inv = sharding.reduce_over_ranks([ms, e2e_ms, 1.0 / pcie["h2d"], 1.0 / pcie["d2h"]], "max")   # device time: max over ranks
    agg = [prof["solve_ms"], float(prof["solve_launches"]), float(prof["contact_visits"]), float(prof["joint_visits"]),
           float(prof["total_launches"])]                                    # roofline / launches: rank 0's GPU
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = sharding.job_throughput(float(B) * nb * args.steps, world, ms)
    e2e_value = sharding.job_throughput(float(B) * nb * e2e_steps, world, e2e_ms)
    peak, peak_src = load_peaks()
    solve_ms, launches = float(agg[0]), max(float(agg[1]), 1.0)
    alg_bytes = BYTES_PER_CONTACT_VISIT * float(agg[2]) + BYTES_PER_HINGE_VISIT * float(agg[3])
    achieved = alg_bytes / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
    # measured DRAM bytes of the solver kernel (one `ncu --set full` capture, profiles/pgs_traffic.json): used only while the
    # record was taken on these very kernel sources, workload and batch
    traffic, traffic_note = None, "no ncu capture recorded for this workload / batch"
    tp = os.path.join(ROOT, "profiles", "pgs_traffic.json")
    if os.path.exists(tp):
        try:
            for rec in json.load(open(tp)):
                if rec.get("workload") == args.workload and rec.get("scenes") == B:
                    if rec.get("source_hash") == solver_source_hash():
                        traffic, traffic_note = rec.get("dram_bytes_per_launch"), rec.get("source")
                    else:
                        traffic_note = "stale: the recorded capture was taken on different solver sources"
        except Exception:
            pass
    kernel = next((k for k in ("k_pgs_pipe", "k_pgs_color", "k_pgs_sp", "k_pgs_lv") if k in variants), "k_pgs")
    cfg = common_config(args.workload, B, nb, nj)
    out = {
        "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "details": {"kernels": variants, "contacts_per_scene": round(contacts_mean, 1),
                    "sharding": f"scene index, {world} ranks, no data-path collective", "host_binding": binding,
                    "timed_region": "CUDA-graph replay of slrbs_step; roofline from a separate profiled pass of "
                                    f"{psteps} steps ({prof_ms / psteps:.3f} ms/step with per-launch events)"},
        "roofline": {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_note,
                     # measured DRAM bytes / event time / peak: what the memory system actually moved (the contract figure
                     # `frac` counts 464 B per contact visit, more than a row record holds)
                     "frac_dram": (traffic / (solve_ms / launches * 1e-3) / 1e9 / peak) if traffic else None,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes / launches, "ms_per_launch": solve_ms / launches,
                     # SURVEY 8(d), informational: the same visits at 244 B per contact visit (rows recomputed from compact
                     # data instead of streamed); the fraction above uses the 464 B contract figure
                     "achieved_at_244B_per_contact_visit": ((244.0 * float(agg[2]) + BYTES_PER_HINGE_VISIT * float(agg[3])) / (solve_ms * 1e-3) / 1e9
                                                           if solve_ms > 0 else 0.0),
                     "share_of_step": solve_ms / prof_ms if prof_ms else None},
        "e2e": {"value": e2e_value, "unit": "body-steps/s", "h2d_bytes_per_step": int(B * nb * 13 * 4),
                "d2h_bytes_per_step": int(B * nb * 13 * 4), "steps": e2e_steps, "contexts": len(chunks),
                "payload": "x,q,v,omega of every body up and down every step (13 floats each way)",
                "host_link_gbs_per_rank_all_ranks_copying": {"h2d": 1.0 / h2d_inv, "d2h": 1.0 / d2h_inv},
                # informational second protocol, NOT the headline: what the reference's own viewer loop moves per frame
                "forces_up_poses_down": {"value": sharding.job_throughput(float(B) * nb * e2e_steps, world, poses_ms), "unit": "body-steps/s",
                                         "h2d_bytes_per_step": int(B * nb * 6 * 4), "d2h_bytes_per_step": int(B * nb * 7 * 4),
                                         "payload": "f, tau of every body up (what a PreStepFunc produces), x, q of every body down (what "
                                                    "SimViewer reads); the state stays on the device"}},
        "gpu_launches": launches_timed,
        "clocks": clocks,
    }
    if world == 1:
        out["cpu_baseline"] = cpu_baseline_single(args.workload, settle)
    print(json.dumps(out), flush=True)
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_pile(args):
    """One pile of touching spheres in a five-box container, cut into blocks; ranks own slabs of block columns. Per step:
    slrbs_step on every block scene, local halo copy, packed halo exchange with the slab neighbours (NCCL point-to-point on
    the context's stream, libslrbs_pile.so); every K steps a device-side re-partition (all-reduced state table)."""
    import torch
    import torch.distributed as dist
    from slrbs_b200.partition import DevicePile, NativeHaloComm, sphere_pile
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    binding = bind_rank_to_cores(local, world)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = PILE_DIMS[args.workload]
    x, r, m, fixed = sphere_pile(*dims)
    native = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(NativeHaloComm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        native = NativeHaloComm(rank, world, local, bytes(idt.cpu().numpy().tobytes()))
    # block edge: 8 m up to 4 ranks; 6 m on 8 (measured on 10^6 spheres, 8 GPUs: 2.53 / 2.14 / 2.81 ms per step for 8 / 6 / 5 m --
    # smaller blocks shorten a scene's sweep and even out the slabs, until the halo copies outweigh it)
    block = float(os.environ.get("SLRBS_PILE_BLOCK", "6" if world >= 8 else "8"))
    pile = DevicePile(x, r, m, fixed, block=block, rank=rank, world=world, device=local, native=native)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pile.step(max(args.warmup, 3))
    pile._rebuild()                                         # one full re-partition before the clock starts: the first one sets up
    pile.step(2)                                            # the collectives' communicators and may grow the context
    pile.ctx.sync()
    r0, s0 = pile.stats["rebuilds"], pile.stats.get("steady_rebuild_s", 0.0)
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.2 if sampler else 0.0)                     # nvidia-smi needs ~0.1 s before its first line
    pile.ctx.profile_read(True)
    barrier()
    if sampler:
        sampler.begin()
    t0 = time.perf_counter()
    pile.ctx.timer_start()
    pile.step(args.steps)
    ms = pile.ctx.timer_stop()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    if sampler:
        sampler.end()
    launches = int(pile.ctx.profile_read(True)["total_launches"])

    def same_load():                                        # one rank only: a step of the pile is collective across ranks
        pile.step(2); pile.ctx.sync()
    clocks = sampler.stop(same_load if world == 1 else None) if sampler else None
    pile.ctx.profile_read(True)
    rebuilds, rebuild_ms = pile.stats["rebuilds"] - r0, (pile.stats.get("steady_rebuild_s", 0.0) - s0) * 1e3
    # profiled pass for the solver share (per-launch events)
    psteps = 5
    pile.ctx.profile_enable(True); pile.ctx.profile_read(True)
    pile.step(psteps); pile.ctx.sync()
    prof = pile.ctx.profile_read(True)
    pile.ctx.profile_enable(False)
    # end to end: every step also brings the owners' states to pinned host memory (13 floats per body)
    host = torch.empty((x.shape[0], 13), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pile.step(1)
        pile.table.zero_(); pile.vmax2.zero_(); torch.cuda.synchronize(local)
        pile.ctx.pile_export(pile.table.data_ptr(), pile.vmax2.data_ptr()); pile.ctx.sync()
        host.copy_(pile.table.view(-1, 13), non_blocking=True); torch.cuda.synchronize(local)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    from slrbs_b200 import sharding
    ms, wall_ms, e2e_ms = sharding.reduce_over_ranks([ms, wall_ms, e2e_ms], "max")
    variants = pile.ctx.describe()
    stats = dict(pile.stats)
    pen = pile.max_penetration()
    if rank == 0:
        n = int(x.shape[0])
        peak, peak_src = load_peaks()
        solve_ms, nl = float(prof["solve_ms"]), max(float(prof["solve_launches"]), 1.0)
        alg = BYTES_PER_CONTACT_VISIT * float(prof["contact_visits"])
        achieved = alg / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
        out = {
            "metric": "body_steps_per_sec", "value": n * args.steps / (ms * 1e-3), "unit": "body-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: one pile of {dims[0]}x{dims[1]}x{dims[2]} = {n} spheres in a five-box container",
                       "bodies": n, "dt": DT, "solver": "Box-PGS inside a block (reference row order), block-Jacobi across block faces",
                       "solver_iters": SOLVER_ITERS, "mu": MU, "block_edge": block,
                       "l2": "no flush: the row records of all block scenes exceed the 126 MB L2"},
            "details": {"kernels": variants, "host_binding": binding, "stats": stats, "wall_ms_per_step": wall_ms / args.steps,
                        "rebuilds_in_timed_region": rebuilds, "rebuild_ms_total": rebuild_ms,
                        "halo_exchange": "ncclSend/ncclRecv with the slab neighbours on the context stream (libslrbs_pile.so)" if native else "local indexed copy only (one rank)",
                        "max_penetration": pen,
                        "phases_ms_per_step": {"solver": solve_ms / psteps, "rebuild": rebuild_ms / args.steps,
                                               "other_kernels_and_exchange": max(ms / args.steps - solve_ms / psteps - rebuild_ms / args.steps, 0.0)}},
            "roofline": {"kernel": next((k for k in ("k_pgs_pipe", "k_pgs_color", "k_pgs_sp", "k_pgs_lv") if k in variants), "k_pgs"),
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": None, "frac_dram": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg / nl,
                         "ms_per_launch": solve_ms / nl},
            "e2e": {"value": n * e2e_steps / (e2e_ms * 1e-3), "unit": "body-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(n * 13 * 4), "steps": e2e_steps,
                    "payload": "every step downloads x,q,v,omega of every body from its owner rank; nothing is uploaded (the pile has no per-step host input)"},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(out), flush=True)
    pile.close()
    if native is not None:
        native.close()
    if world > 1:
        dist.destroy_process_group()


# ---- CPU legs: the only place bench.py touches oracle/ -------------------------------------------
def _oracle_system(workload, seed, settle):
    """One CPU scene of the workload, settled.  Returns (system, kind): the reference's own sources
    (oracle/_ref, kind "reference") when that library is present, else the C restatement (kind "port").
    The settle always runs on the restatement (it is bit-identical to oracle/_ref, tests/test_ref_vs_oracle.py,
    and ~2x faster); the settled state is then handed to the system that gets timed."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Reference, ref_available
    scene, abc, _, _, _, jitter = WORKLOADS[workload]
    if scene == "mixed_pile":
        # extension pairs + mesh bodies exist only in the restatement; same body tables as the GPU arm (slrbs_b200/scenes.py),
        # the distance grid built by the restatement's own orc_build_sdf (bit-identical to the device's, tests/test_gpu_sdf.py)
        import mesh_lib
        from slrbs_b200 import scenes
        mv, mf = scenes.blob()
        t = scenes.mixed_pile(abc[0], mv)
        o = Oracle(); o.set_extensions(1)
        sid = o.add_sdf_shape(mesh_lib.build_sdf(mv, mf, cell=0.05))
        for i in range(t["x"].shape[0]):
            g = [sid] if t["geom_type"][i] == 4 else t["geom"][i]
            o.add_body(float(t["mass"][i]), int(t["geom_type"][i]), g, fixed=bool(t["fixed"][i]), x=t["x"][i], q=t["q"][i])
        o.set_params(MU, 0, SOLVER_ITERS, True)
        o.step(DT, min(settle, 2))
        return o, "port"
    kw = {"k": abc[0]} if scene in ("sphere_stack", "box_stack") else ({"nx": abc[0], "nz": abc[1], "ny": abc[2]} if scene == "marble_box_n" else {})
    o = Oracle(scene, **kw)
    if jitter > 0:
        st = o.state(); fixed = o.body_params()["fixed"] == 1
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter, jitter, (o.n_bodies, 2)).astype(np.float32); d[fixed] = 0
        st["x"][:, 0] += d[:, 0]; st["x"][:, 2] += d[:, 1]
        o.set_state(**st)
    o.set_params(MU, 0, SOLVER_ITERS, True)
    o.step(DT, settle)
    if not ref_available() or scene == "box_stack":          # the extension pairs exist only in the restatement
        return o, "port"
    r = Reference(scene, **kw)
    r.set_params(MU, 0, SOLVER_ITERS, True)
    r.set_state(**o.state())
    if o.n_joints:
        r.set_joint_lambda(o.joint_rows()["lam"])
    r.step(DT, 1)                                           # builds its contact list
    return r, "reference"


def cpu_baseline_single(workload, settle, budget_s=12.0):
    """The reference's CPU implementation, one thread, one scene of the same workload; bounded to ~budget_s."""
    o, kind = _oracle_system(workload, 1234, settle)
    n, t = 0, 0.0
    chunk = 10 if o.n_bodies < 5000 else 1
    while t < budget_s:
        t += o.time_steps(DT, chunk); n += chunk
    cores = os.cpu_count()
    out = {"value": o.n_bodies * n / t, "unit": "body-steps/s", "cores": 1, "kind": kind}
    what = "C restatement of the reference (oracle/_ref not built or not applicable)"
    if kind == "reference":
        # the same scene on the C restatement (bit-identical results, no Eigen-style temporaries): the faster CPU figure
        from oracle_lib import Oracle
        scene, abc = WORKLOADS[workload][0], WORKLOADS[workload][1]
        kw = {"k": abc[0]} if scene in ("sphere_stack", "box_stack") else ({"nx": abc[0], "nz": abc[1], "ny": abc[2]} if scene == "marble_box_n" else {})
        p = Oracle(scene, **kw)
        p.set_params(MU, 0, SOLVER_ITERS, True)
        p.set_state(**o.state())
        if o.n_joints:
            p.set_joint_lambda(o.joint_rows()["lam"])
        pn, pt = 0, 0.0
        while pt < 0.4 * budget_s:
            pt += p.time_steps(DT, chunk); pn += chunk
        out["port_value"] = p.n_bodies * pn / pt
        what = (f"reference sources compiled against the Eigen API shim (oracle/_ref); port_value = the C restatement on the same "
                f"scene ({pn} steps)")
    if workload == "mixed_pile":
        settle = min(settle, 2)         # the all-pairs CPU loop over 10^4 mixed bodies takes seconds per step: the sample starts from the lattice
    out["sample"] = (f"1 scene of {workload} ({o.n_bodies} bodies, {o.n_contacts} contacts), {n} steps after a {settle}-step settle, "
                     f"single thread like the reference; {what}; host has {cores} logical cores")
    return out


def _ref_worker(job):
    workload, seed, settle, steps = job
    o, kind = _oracle_system(workload, seed, settle)
    o.time_steps(DT, 2)
    return o.time_steps(DT, steps), o.n_bodies, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    workload = args.workload
    settle = WORKLOADS[workload][4]
    cores = os.cpu_count() or 1
    sub = {"marble_box": 20, "marble_box_1k": 2}.get(workload, 1000)    # CPU steps per worker per bench step
    total = sub * (args.steps + args.warmup)
    with ProcessPoolExecutor(max_workers=cores) as ex:
        t0 = time.perf_counter()
        res = list(ex.map(_ref_worker, [(workload, 1234 + i, settle, total) for i in range(cores)]))
        wall = time.perf_counter() - t0
    nb = res[0][1]
    # the library the workers timed, loaded once in this process too (the driver's loaded-library record looks at the parent)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import lib, ref_available
    loaded = lib("ref" if res[0][2] == "reference" and ref_available() else "oracle")
    nj = _oracle_system(workload, 1234, 0)[0].n_joints
    del loaded
    tmax = max(r[0] for r in res)
    value = cores * nb * total / tmax
    B = args.scenes or WORKLOADS[workload][3]
    out = {
        "impl": "reference", "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tmax * 1e3 / (args.steps + args.warmup),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(workload, B, nb, nj),
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": cores, "kind": res[0][2],
                         "sample": f"{cores} independent scenes of {workload}, one process per logical core ("
                                   + ("reference sources + Eigen API shim, oracle/_ref" if res[0][2] == "reference" else "C restatement")
                                   + f"), {sub} reference steps per "
                                   f"bench step ({total} per scene) after a {settle}-step settle; wall {wall:.1f} s incl. settle"},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="marble_box_1k", choices=sorted(WORKLOADS) + sorted(PILE_DIMS))
    ap.add_argument("--scenes", type=int, default=0, help="scenes per GPU (default: per workload)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload in PILE_DIMS:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference's all-pairs loop cannot run a pile of this size "
                              "(5e11 pair tests per step for 1e6 spheres); see cpu_baseline.extrapolated in the other arm"}), flush=True)
        else:
            run_pile(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
