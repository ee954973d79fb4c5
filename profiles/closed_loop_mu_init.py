"""Closed loop (2048 rollouts x 60 ticks, predictor-corrector) against cmpc_config.mu_init: the value doubles as the multiplier floor of
warm-started ticks when warm_start_mu_init is left at it; the experiment behind the default warm_start_mu_init = 0.01.
usage: python profiles/closed_loop_mu_init.py"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
import torch
pkg = importlib.import_module(PKG); R = importlib.import_module(PKG + ".rollout")
for tol in (1e-4, 1e-8):
    for mi in (0.1, 1e-2, 1e-3, 1e-4, 1e-5):
        s = pkg.BatchedCentroidalMPC(pkg.ergocub_config(ipopt_tolerance=tol, mu_init=mi, warm_start_mu_init=mi))
        out = R.closed_loop_rollout(s, B=2048, ticks=60, seed=1, push_range=(1.0, 3.0), time_device=True)
        s.close()
        print("tol", tol, "mu_init", mi, "iters/tick %.2f" % (out["iterations"].sum() / (2048 * 60)), "converged %.4f" % (out["converged_ticks"].sum() / (2048 * 60)), "device ms %.0f" % out["device_ms"], "ticks/s %.0f" % (2048 * 60 / out["device_ms"] * 1e3), flush=True)
