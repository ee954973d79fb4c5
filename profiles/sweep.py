"""BASELINE config 5: horizon sweep (10 - 50 knots) and batch sweep (1 - 10^6 instances) of the batched solve on one B200,
with the CPU oracle (the restatement of the reference's IPOPT path, one instance per host core) timed beside every line.
CUDA-event timing, device-resident inputs, cold start, ergoCub weights, step adjustment on, tol 1e-8, library defaults.
GPU column: solves/s and executed-FP64 roofline fraction; CPU column: oracle solves/s on a bounded sample of the same instances
(all host cores); latency: batch 1.  Batches above 65536 repeat a 65536-instance set on the device (generating and holding 10^6
distinct 25 KB instances on the host is the only reason; the solver sees independent instances either way).
usage: python profiles/sweep.py > profiles/r2_sweep.txt"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
pkg = importlib.import_module(PKG)
wl = importlib.import_module(PKG + ".workloads")
import torch  # noqa: E402

from oracle import oracle as om  # noqa: E402  (reported CPU baseline, never the thing measured on the GPU side)

FLOP_PER_ITER_KNOT = 1.839e5   # executed FP64 flop per iteration and knot (ncu capture r2_prof_final3, see bench.py)
om.build()
O = om.Oracle()
CORES = os.cpu_count() or 1


def cpu_rate(N, w, sample):
    cfg = om.make_cfg(N=N, w_pos=2000.0)
    t0 = time.perf_counter()
    x, lam, st = O.solve_batch(cfg, w["p"][:sample], w["lbg"][:sample], w["ubg"][:sample], w["x0"][:sample], threads=CORES)
    dt = time.perf_counter() - t0
    return sample / dt, float(np.mean([s.iters for s in st])), sum(1 for s in st if s.status == 0)


def run(N, B, reps=3, peak=None, cpu_sample=64):
    cfg = pkg.ergocub_config(horizon=N)
    s = pkg.BatchedCentroidalMPC(cfg)
    gen = min(B, 65536)
    w = wl.walk_batch(N=N, dT=0.1, B=gen, seed=0, state_noise=1.0, yaw_range=0.2)
    rep = (B + gen - 1) // gen
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda().repeat(rep, 1)[:B].contiguous()  # noqa: E731
    p, lb, ub, x0 = t(w["p"]), t(w["lbg"]), t(w["ubg"]), t(w["x0"])
    lam = torch.zeros(B, s.L.m, dtype=torch.float64, device="cuda")
    x = torch.empty_like(x0)
    ms = []
    for r in range(reps + 1):
        x.copy_(x0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        obj, st, it, _ = s.solve(p, lb, ub, x, lam)
        b.record()
        torch.cuda.synchronize()
        if r:
            ms.append(a.elapsed_time(b))
    t_s = float(np.median(ms)) * 1e-3
    conv = int((st == 0).sum().item())
    iters = float(it.double().sum().item())
    if peak is None:
        peak = s.measure_fp64_peak()
    frac = iters * N * FLOP_PER_ITER_KNOT / t_s / 1e12 / peak
    s.close()
    del p, lb, ub, x0, x, lam
    torch.cuda.empty_cache()
    c_rate, c_it, c_ok = cpu_rate(N, w, min(gen, cpu_sample))
    return dict(N=N, B=B, ms=1e3 * t_s, rate=B / t_s, conv=conv, iters=iters / B, frac=frac, cpu=c_rate, cpu_it=c_it, peak=peak)


HDR = f"{'N':>4} {'batch':>8} {'ms':>11} {'GPU solves/s':>13} {'converged':>10} {'iters':>6} {'FP64 frac':>9} | {'CPU solves/s':>12} {'iters':>6} {'GPU/CPU':>8}"


def line(r):
    return (f"{r['N']:4d} {r['B']:8d} {r['ms']:11.3f} {r['rate']:13.0f} {r['conv']:10d} {r['iters']:6.2f} {r['frac']:9.4f} | "
            f"{r['cpu']:12.1f} {r['cpu_it']:6.1f} {r['rate'] / r['cpu']:8.1f}")


print(f"CPU column: oracle (IPOPT restatement, monotone = IPOPT's default path) on {CORES} host cores, sample of <= 64 instances; "
      "GPU: library defaults (predictor-corrector), tol 1e-8")
print("horizon sweep (batch 1036 = one instance per resident team)")
print(HDR)
peak = None
for N in (10, 12, 15, 20, 25, 30, 40, 50):
    r = run(N, 1036, peak=peak)
    peak = r["peak"]
    print(line(r), flush=True)
print("batch sweep (N = 12); batch 1 = single-solve latency")
print(HDR)
for B in (1, 8, 64, 512, 4096, 32768, 262144, 1000000):
    print(line(run(12, B, reps=2 if B < 262144 else 1, peak=peak)), flush=True)
print(f"measured FP64 FMA peak used for the fraction: {peak:.2f} TFLOP/s")
