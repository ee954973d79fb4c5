"""BASELINE config 5: horizon sweep and batch sweep of the batched solve on one B200 (CUDA-event timing, device-resident
inputs, cold start, ergoCub weights, step adjustment on).  usage: python profiles/sweep.py > profiles/r1_sweep.txt"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
pkg = importlib.import_module(PKG)
wl = importlib.import_module(PKG + ".workloads")
import torch


def run(N, B, reps=3):
    cfg = pkg.ergocub_config(horizon=N)
    s = pkg.BatchedCentroidalMPC(cfg)
    w = wl.walk_batch(N=N, dT=0.1, B=B, seed=0, state_noise=1.0, yaw_range=0.2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    p, lb, ub, x0 = t(w["p"]), t(w["lbg"]), t(w["ubg"]), t(w["x0"])
    ms = []
    for r in range(reps + 1):
        x = x0.clone()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        obj, st, it, _ = s.solve(p, lb, ub, x)
        b.record()
        torch.cuda.synchronize()
        if r:
            ms.append(a.elapsed_time(b))
    conv = int((st == 0).sum().item())
    res = (N, B, float(np.median(ms)), B / (np.median(ms) * 1e-3), conv, float(it.double().mean().item()))
    s.close()
    return res


print("horizon sweep (batch 1036 = one instance per resident team)")
print(f"{'N':>4} {'batch':>8} {'ms':>10} {'solves/s':>12} {'converged':>10} {'mean iters':>11}")
for N in (10, 12, 15, 20, 25, 30, 40, 50):
    print("{:4d} {:8d} {:10.2f} {:12.0f} {:10d} {:11.2f}".format(*run(N, 1036)), flush=True)
print("batch sweep (N = 12)")
for B in (1, 8, 64, 512, 4096, 32768, 262144):
    print("{:4d} {:8d} {:10.2f} {:12.0f} {:10d} {:11.2f}".format(*run(12, B, reps=2)), flush=True)
