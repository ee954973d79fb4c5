"""Convergence stress of the default configuration: seeds x robots x horizons x tolerances; prints non-converged counts.
usage: python profiles/stress.py"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
pkg = importlib.import_module(PKG); wl = importlib.import_module(PKG + ".workloads")
bad = 0
def run(name, cfg, w):
    global bad
    s = pkg.BatchedCentroidalMPC(cfg)
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    s.close()
    nb = int((status != 0).sum()); bad += nb
    print(f"{name:58s} B {len(status):5d} not converged {nb:3d} status {np.bincount(status, minlength=5).tolist()} iters mean {iters.mean():5.2f} max {iters.max():3d}", flush=True)
for tol in (1e-8, 1e-4):
    for seed in (1, 2, 3):
        run(f"icub3 N=15 no step adjustment seed {seed} tol {tol:g}", pkg.icub3_config(ipopt_tolerance=tol),
            wl.walk_batch(N=15, B=4096, seed=seed, state_noise=1.0, step_adjust=False))
        run(f"ergocub N=12 noise 2 yaw .3 seed {seed} tol {tol:g}", pkg.ergocub_config(ipopt_tolerance=tol),
            wl.walk_batch(N=12, B=4096, seed=seed, state_noise=2.0, yaw_range=0.3))
    for N in (10, 20, 30, 50):
        run(f"ergocub N={N} noise 1.5 yaw .2 tol {tol:g}", pkg.ergocub_config(horizon=N, ipopt_tolerance=tol),
            wl.walk_batch(N=N, B=1036, seed=N, state_noise=1.5, yaw_range=0.2))
    rng = np.random.default_rng(7)
    push = rng.uniform(-3, 3, size=(4096, 3)) * [1, 1, 0]
    run(f"ergocub N=12 pushes up to 3 m/s^2 tol {tol:g}", pkg.ergocub_config(ipopt_tolerance=tol),
        wl.walk_batch(N=12, B=4096, seed=9, state_noise=1.0, yaw_range=0.2, push=push))
    run(f"ergocub N=12 noise 3 tol {tol:g}", pkg.ergocub_config(ipopt_tolerance=tol),
        wl.walk_batch(N=12, B=4096, seed=11, state_noise=3.0, yaw_range=0.3))
print("TOTAL not converged:", bad)
