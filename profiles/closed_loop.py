"""BASELINE config 4 on one GPU: closed-loop batch of rollouts (MPC solve warm-started by the shifted previous solution, RK4
plant, pushes, actual footsteps fed back).  usage: python profiles/closed_loop.py [rollouts] [ticks] [tol]
Prints rollouts, ticks, converged fraction, mean iterations per tick, device time of the kernels and MPC ticks per second."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
pkg = importlib.import_module(PKG)
R = importlib.import_module(PKG + ".rollout")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 100
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-8
for strat in ("mehrotra", "monotone"):
    s = pkg.BatchedCentroidalMPC(pkg.ergocub_config(ipopt_tolerance=tol, mu_strategy=pkg.MU_MEHROTRA if strat == "mehrotra" else pkg.MU_MONOTONE))
    t0 = time.perf_counter()
    out = R.closed_loop_rollout(s, B=B, ticks=ticks, seed=1, push_range=(1.0, 3.0), time_device=True)
    wall = time.perf_counter() - t0
    s.close()
    conv = out["converged_ticks"].sum() / (B * ticks)
    print(f"closed loop {strat}: {B} rollouts x {ticks} ticks, tol {tol:g}: converged ticks {100 * conv:.3f} %, "
          f"iterations / tick {out['iterations'].sum() / (B * ticks):.2f}, device {out['device_ms']:.0f} ms "
          f"= {B * ticks / out['device_ms'] * 1e3:.0f} MPC ticks/s (wall {wall:.1f} s = {B * ticks / wall:.0f} ticks/s end to end: schedule from the device table, no host synchronisation inside the loop), "
          f"CoM error max {out['com_err_max'].max():.3f} m, min CoM height {out['com_z_min'].min():.3f} m, "
          f"rollouts with every tick converged {int((out['converged_ticks'] == ticks).sum())}/{B}", flush=True)
