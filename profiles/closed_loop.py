"""BASELINE config 4: closed-loop batch of rollouts (MPC solve warm-started by the shifted previous solution, RK4 plant, pushes,
actual footsteps fed back).  usage: python profiles/closed_loop.py [rollouts] [ticks] [tol]
Under torchrun (one rank per GPU) the rollouts are sharded over the ranks (contiguous ranges, own seeds), every rank rolls out
its shard without any exchange, and the per-rollout results are gathered ONCE at the end over NCCL: rank 0
prints rollouts, ticks, converged fraction, mean iterations per tick, MPC ticks per second of device time (max over ranks)
and of wall time."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
import torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module(PKG)
R = importlib.import_module(PKG + ".rollout")
sharding = importlib.import_module(PKG + ".sharding")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 100
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-8
lo, hi = sharding.shard_bounds(B, rank, world)
sizes = [b - a for a, b in (sharding.shard_bounds(B, r, world) for r in range(world))]
for strat in ("mehrotra", "monotone"):
    s = pkg.BatchedCentroidalMPC(pkg.ergocub_config(ipopt_tolerance=tol, device=local,
                                                    mu_strategy=pkg.MU_MEHROTRA if strat == "mehrotra" else pkg.MU_MONOTONE))
    R.closed_loop_rollout(s, B=min(hi - lo, 64), ticks=3, seed=1)          # warm-up (library load, allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = R.closed_loop_rollout(s, B=hi - lo, ticks=ticks, seed=1 + rank, push_range=(1.0, 3.0), time_device=True)
    dev = torch.device("cuda", local)
    local_res = torch.stack([torch.from_numpy(out[k].astype(np.float64)) for k in ("converged_ticks", "iterations", "com_err_max", "com_z_min")], dim=1).to(dev)
    allres = sharding.gather_results(local_res, world, sizes).cpu().numpy()     # the one exchange of the job
    tt = torch.tensor([out["device_ms"], (time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    s.close()
    if rank == 0:
        dev_ms, wall_ms = tt.tolist()
        conv = allres[:, 0].sum() / (B * ticks)
        print(f"closed loop {strat}: {B} rollouts x {ticks} ticks on {world} GPU(s), tol {tol:g}: converged ticks {100 * conv:.3f} %, "
              f"iterations / tick {allres[:, 1].sum() / (B * ticks):.2f}, device {dev_ms:.0f} ms "
              f"= {B * ticks / dev_ms * 1e3:.0f} MPC ticks/s (wall {wall_ms / 1e3:.2f} s = {B * ticks / wall_ms * 1e3:.0f} ticks/s end to end: schedule "
              f"from the device table, no host synchronisation inside the loop, one gather at the end), "
              f"CoM error max {allres[:, 2].max():.3f} m, min CoM height {allres[:, 3].min():.3f} m, "
              f"rollouts with every tick converged {int((allres[:, 0] == ticks).sum())}/{B}", flush=True)
if world > 1:
    dist.destroy_process_group()
