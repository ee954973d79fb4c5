"""Launch time of small and mid-size batches (iCub3, N = 15, cold start) with the default dispatch: up to 2 / up to 4 instances per
SM on independent single-team CTAs, larger batches on the seven-team lock-step CTAs (round 1 switched between the two with environment variables;
the library no longer reads the environment: force the seven-team kernel with teams_per_cta = 7).  usage: [CMPC_B200_LIB=...] python profiles/lat_sweep.py B1 B2 ..."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
pkg = importlib.import_module(PKG); wl = importlib.import_module(PKG + ".workloads")
import torch
def run(B, reps=3):
    s = pkg.BatchedCentroidalMPC(pkg.icub3_config())
    w = wl.walk_batch(N=15, dT=0.1, B=B, seed=0, state_noise=1.0, step_adjust=False)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    p, lb, ub, x0 = t(w["p"]), t(w["lbg"]), t(w["ubg"]), t(w["x0"])
    ms = []
    for r in range(reps + 1):
        x = x0.clone(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); obj, st, it, _ = s.solve(p, lb, ub, x); b.record(); torch.cuda.synchronize()
        if r: ms.append(a.elapsed_time(b))
    s.close(); return float(np.median(ms)), int((st == 0).sum().item())
for B in [int(a) for a in sys.argv[1:]] or [148, 296]:
    ms, conv = run(B)
    print("lib", os.path.basename(os.environ.get("CMPC_B200_LIB", "default")), "batch", B, f"{ms:.2f} ms", f"{B / ms:.1f} solves/ms", "converged", conv, flush=True)
