"""Smallest cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), one kernel family per run:
  lockstep : 16 instances on the seven-team lock-step CTAs (named barriers, bar.red votes, cp.async double buffers), 3 groups
  single4  : 3 instances per SM on the single-team kernel compiled for 4 CTAs per SM (<128,1,4>)
  single1  : one instance on the single-team latency kernel (<128,1,1>)
  aux      : populate, shift, plant, rollout tick / feedback, resample, zmp, eval / jac / hess kernels
usage: compute-sanitizer --tool <tool> python profiles/sanitize.py <case>"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = importlib.import_module("paper_romualdi_2022_icra_centroidal-mpc-walking_b200")
wl = importlib.import_module("paper_romualdi_2022_icra_centroidal-mpc-walking_b200.workloads")
R = importlib.import_module("paper_romualdi_2022_icra_centroidal-mpc-walking_b200.rollout")
case = sys.argv[1]
sms = torch.cuda.get_device_properties(0).multi_processor_count
if case == "lockstep":
    s = P.BatchedCentroidalMPC(P.ergocub_config(teams_per_cta=7))
    w = wl.walk_batch(N=12, B=16, seed=1, state_noise=1.5, yaw_range=0.2)
elif case == "single4":
    s = P.BatchedCentroidalMPC(P.ergocub_config())
    w = wl.walk_batch(N=12, B=2 * sms + 8, seed=1, state_noise=1.5, yaw_range=0.2)
elif case == "single1":
    s = P.BatchedCentroidalMPC(P.ergocub_config())
    w = wl.walk_batch(N=12, B=1, seed=1, state_noise=1.5, yaw_range=0.2)
elif case == "aux":
    s = P.BatchedCentroidalMPC(P.ergocub_config(ipopt_tolerance=1e-4))
    out = R.closed_loop_rollout_device(s, B=8, ticks=3, seed=0, use_graph=False, yaw_range=0.2)
    w = wl.walk_batch(N=12, B=4, seed=1, ticks=True)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    p, lbg, ubg, x0 = s.populate(dev(w["ticks"]))
    f, grad, g, jnz = s.eval_jac_fg(x0, p)
    h = s.eval_hess_l(x0, p, 1.0, torch.zeros(4, s.L.m, dtype=torch.float64, device="cuda"))
    z, v = s.desired_zmp(x0, p)
    tk = dev(w["ticks"])
    s.resample_references(tk, dev(np.arange(10) * 0.2), dev(np.zeros((4, 10, 3))), dev(np.zeros((4, 10, 3))), dev(np.arange(13) * 0.1), 50.0, 0.7)
    torch.cuda.synchronize()
    print("aux ok", out["converged_ticks"].tolist())
    sys.exit(0)
else:
    raise SystemExit(case)
x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
print(case, "status", np.bincount(status), "iterations", iters[:8].tolist(), "geometry", s.geometry())
assert (status == 0).all()
