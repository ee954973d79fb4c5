"""Does the L2 residency of the per-team scratch bound the loaded SM?  Horizon sweep at one instance per resident team (batch
1036) down to horizons whose whole scratch (iterate vectors + factors) fits the 126 MB L2: cost per knot and iteration.
usage: python profiles/l2_fit.py [library.so] [N1,N2,...]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
pkg = importlib.import_module(PKG)
if len(sys.argv) > 1:
    pkg.LIB_PATH = os.path.abspath(sys.argv[1])
wl = importlib.import_module(PKG + ".workloads")
import torch  # noqa: E402


def footprint(N):   # works_doubles of csrc/cmpc_ipm.cuh, bytes: (vectors, factors)
    vec = 5 * (N + 1) * 48 + 6 * (N + 1) * 16 + 17 * N * 40 + (N + 1) * 100 + N * 80 + (N + 1) * 124
    return 8 * vec, 8 * N * 1312


print(f"{'N':>3} {'ms':>8} {'iters':>6} {'us / (knot iteration)':>22} {'vectors MB':>11} {'factors MB':>11}   (1036 resident teams, L2 = 126 MB)")
HORIZONS = tuple(int(a) for a in sys.argv[2].split(",")) if len(sys.argv) > 2 else (3, 4, 5, 6, 8, 10, 12, 15, 20)
for N in HORIZONS:
    cfg = pkg.ergocub_config(horizon=N)
    s = pkg.BatchedCentroidalMPC(cfg)
    w = wl.walk_batch(N=N, dT=0.1, B=1036, seed=0, state_noise=1.0, yaw_range=0.2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    p, lb, ub, x0 = t(w["p"]), t(w["lbg"]), t(w["ubg"]), t(w["x0"])
    lam = torch.zeros(1036, s.L.m, dtype=torch.float64, device="cuda")
    x = torch.empty_like(x0)
    ms = []
    for r in range(4):
        x.copy_(x0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        obj, st, it, _ = s.solve(p, lb, ub, x, lam)
        b.record()
        torch.cuda.synchronize()
        if r:
            ms.append(a.elapsed_time(b))
    m = float(np.median(ms))
    iters = float(it.double().mean().item())
    mx = float(it.max().item())
    v, f = footprint(N)
    print(f"{N:3d} {m:8.3f} {iters:6.2f} {1e3 * m / (N * mx):10.2f} (max it) {1e3 * m / (N * iters):8.2f} (mean) {1036 * v / 1e6:11.1f} {1036 * f / 1e6:11.1f}  conv {(st == 0).sum().item()}", flush=True)
    s.close()
