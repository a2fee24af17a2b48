"""Multi-GPU run of the partitioned pile (launched by torchrun, one rank per GPU). Prints one JSON line on rank 0:
penetration / drift of the partitioned run against the same pile stepped as ONE scene on rank 0, and throughput."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from slrbs_b200 import Context
from slrbs_b200.partition import DevicePile, NativeHaloComm, PartitionedPile, TorchComm, contact_velocity_residual, sphere_pile

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=3, default=[24, 8, 12]); ap.add_argument("--block", type=float, default=6.0)
ap.add_argument("--steps", type=int, default=60); ap.add_argument("--rebuild", type=int, default=10)
ap.add_argument("--check", type=int, default=1); ap.add_argument("--device-rebuild", type=int, default=0)
ap.add_argument("--slack", type=float, default=0.5); ap.add_argument("--mixed-radii", type=int, default=0)
ap.add_argument("--native", type=int, default=1, help="per-step halo refresh through libslrbs_pile.so (ncclSend/ncclRecv on the context stream)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, r, m, fixed = sphere_pile(*a.dims)
if a.mixed_radii:                                  # BASELINE config 4 substitute: mixed spheres r in {0.25, 0.5}, mass ~ r^3
    rng = np.random.default_rng(7)
    small = rng.random(x.shape[0]) < 0.5
    r = np.where(small, 0.25, 0.5).astype(np.float32); m = np.where(small, 0.125, 1.0).astype(np.float32)
comm = TorchComm(local) if world > 1 else None
native = None
if a.device_rebuild and a.native and world > 1:
    # the NCCL unique id is made on rank 0 and broadcast over the torch process group (any side channel would do)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(NativeHaloComm.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    native = NativeHaloComm(rank, world, local, bytes(idt.cpu().numpy().tobytes()))
if a.device_rebuild:
    pile = DevicePile(x, r, m, fixed, block=a.block, rank=rank, world=world, device=local, slack=a.slack, native=native)
else:
    pile = PartitionedPile(x, r, m, fixed, block=a.block, rebuild_every=a.rebuild, rank=rank, world=world, device=local, comm=comm)
pile.step(5)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
pile.step(a.steps)
pile.ctx.sync(); torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
pen = pile.max_penetration()
res = pile.velocity_residual() if a.device_rebuild else contact_velocity_residual(pile.ctx, pile.ctx.n_scenes)
state = pile.gather_state()
out = dict(world=world, halo_exchange=("native ncclSend/ncclRecv on the context stream" if native else "torch all_to_all_single"), bodies=int(x.shape[0]), steps=a.steps, ms_per_step=1e3 * dt / a.steps, body_steps_per_s=x.shape[0] * a.steps / dt,
           max_penetration=pen, approach_speed_residual=res[0], contacts=res[1], stats=pile.stats, finite=bool(np.isfinite(state).all()))
if rank == 0 and a.check and x.shape[0] <= 5000:
    one = PartitionedPile(x, r, m, fixed, block=max(a.dims) * 2.0 + 8.0, rebuild_every=10 ** 9, rank=0, world=1, device=local)
    one.step(5 + a.steps)
    ref = one.gather_state()
    d = np.linalg.norm(state[:, :3] - ref[:, :3], axis=1)
    mres = contact_velocity_residual(one.ctx, one.ctx.n_scenes)
    out.update(mono_max_penetration=one.max_penetration(), mono_approach_speed_residual=mres[0], mono_contacts=mres[1], mean_drift=float(d.mean()), max_drift=float(d.max()),
               mean_height=float(state[:, 1].mean()), mono_mean_height=float(ref[:, 1].mean()),
               kinetic=float((state[:, 7:10] ** 2).sum()), mono_kinetic=float((ref[:, 7:10] ** 2).sum()))
    one.close()
if rank == 0:
    print(json.dumps(out), flush=True)
pile.close()
if native is not None:
    native.close()
if world > 1:
    dist.destroy_process_group()
