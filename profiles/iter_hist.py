"""Iteration / status histogram of a workload for both barrier updates.  usage: python profiles/iter_hist.py [workload] [batch]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
pkg = importlib.import_module(PKG)
wl = importlib.import_module(PKG + ".workloads")
import torch
name = sys.argv[1] if len(sys.argv) > 1 else "ergocub"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8288
for strat in ("mehrotra", "monotone"):
    kw = dict(mu_strategy=pkg.MU_MEHROTRA if strat == "mehrotra" else pkg.MU_MONOTONE)
    if name == "ergocub":
        cfg = pkg.ergocub_config(**kw)
        w = wl.walk_batch(N=12, dT=0.1, B=B, seed=0, state_noise=2.0, yaw_range=0.3, step_adjust=True)
    else:
        cfg = pkg.icub3_config(**kw)
        w = wl.walk_batch(N=15, dT=0.1, B=B, seed=0, state_noise=1.0, step_adjust=False)
    s = pkg.BatchedCentroidalMPC(cfg)
    for rep in range(2):
        t0 = time.perf_counter()
        x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
        dt = time.perf_counter() - t0
    print(name, B, strat, f"{B/dt:.0f} solves/s (host call)", "status", np.bincount(status, minlength=5).tolist(),
          "iters mean %.2f p50 %d p99 %d max %d" % (iters.mean(), np.median(iters), np.percentile(iters, 99), iters.max()),
          "hist>30:", np.sort(iters[iters > 30]).tolist()[-20:], "failed idx", np.where(status != 0)[0].tolist()[:10], "idx iters>45:", np.where(iters > 45)[0].tolist()[:40], flush=True)
    s.close()
