"""Cycle accounting of the solve phases with the -DCMPC_PROFILE build (profiles/_build/libcmpc_prof.so): thread 0's clock64()
per phase, summed over instances.  usage: python profiles/phase_cycles.py [batch] [team]"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"
pkg = importlib.import_module(PKG)
pkg.LIB_PATH = os.environ.get("CMPC_PROF_LIB", os.path.join(ROOT, "profiles", "_build", "libcmpc_prof.so"))
wl = importlib.import_module(PKG + ".workloads")
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
team = int(sys.argv[2]) if len(sys.argv) > 2 else 0
strategy = os.environ.get("CMPC_MU", "mehrotra")
cfg = pkg.icub3_config(mu_strategy=pkg.MU_MEHROTRA if strategy == "mehrotra" else pkg.MU_MONOTONE); cfg.threads_per_instance = team
s = pkg.BatchedCentroidalMPC(cfg)
w = wl.walk_batch(N=15, dT=0.1, B=B, seed=0, state_noise=1.0, step_adjust=False)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
p, lb, ub, x0 = t(w["p"]), t(w["lbg"]), t(w["ubg"]), t(w["x0"])
out = (C.c_longlong * 16)()
for rep in range(3):
    x = x0.clone()
    s.lib.cmpc_debug_profile(out)
    obj, st, it, _ = s.solve(p, lb, ub, x)
    torch.cuda.synchronize()
s.lib.cmpc_debug_profile(out)
names = ["setup", "kkt_pass", "barrier_pass", "backward", "forward (x2 with corrector)", "recover (+ affine pass)", "step", "refine", "line search", "accept",
         "corrector backward sweep", "  bw: F1 load", "  bw: F2 form", "  bw: F3 form", "  bw: factor", "  bw: syrk+store"]
mode = os.environ.get("CMPC_PROF_MODE", "1")   # 2: slots 10..14 = sub-phases of the forward sweeps, 3: of the corrector backward sweep
if mode == "2":
    names[11:] = ["  fw: wait for the block", "  fw: t = Y dxi", "  fw: substitution chain + barrier", "  fw: next dxi", "  fw: -"]
elif mode == "5":
    names[11:] = ["  bw: F3a", "  bw: F3b + first diagonal block", "  bw: SYRK tiles (warp 0)", "  bw: factor stores + barrier", "  bw: copy + barrier"]
elif mode == "4":
    names[11:] = ["  fa: first diagonal block", "  fa: 4 panels", "  fa: 3 trailing updates + look-ahead", "  fa: -", "  fa: -"]
elif mode == "3":
    names[11:] = ["  cb: wait + stage tables", "  cb: h_u", "  cb: substitution chain + barrier", "  cb: Y'z", "  cb: copy"]
v = np.array(list(out[:10]) + [out[15]], dtype=np.float64)
vb = np.array(out[10:15], dtype=np.float64)
iters = float(it.sum().item())
print(f"strategy {strategy} batch {B} team {team or 'default'} iterations {iters:.0f}  total cycles/iteration {v.sum()/iters:.0f}")
for n, c in zip(names, list(v) + list(vb)):
    print(f"{n:28s} {c/iters:12.0f} cycles/iteration  {100*c/v.sum():5.1f} %")
