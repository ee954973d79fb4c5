#!/usr/bin/env python
"""bench.py -- batched centroidal-MPC solves/sec on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload icub3_b1024|ergocub_b65536]

A "step" = one pass of the hot path (cmpc_solve_batched: the full interior-point solve of every instance, cold start
from the same x0) over one batch of synthetic MPC instances.  Default workload = BASELINE.json configs[1]:
batch 1024 iCub3 instances (iCubGazeboV3/centroidal_mpc.ini: N = 15, dT = 0.1), randomised CoM / momentum initial
states, step adjustment off, ipopt tolerance 1e-8 -- per GPU (weak scaling over --gpus).  --tol defaults to the
ipopt_tolerance of the workload's robot ini (iCub3: 1e-8; ergoCubGazeboV1_1: 1e-4); --mu-strategy selects the barrier update
(mehrotra = library default, monotone = IPOPT's default path).
  value    : solves/s with inputs resident in HBM, CUDA-event time of the K steps (max over ranks)
  e2e      : solves/s through cmpc_solve_ticks_host, the per-tick call of the drop-in host class, with pinned HOST buffers: the
             compact tick records (state, references, contact windows) go up, the formal input is expanded on the device, the
             solution comes back (H2D + populate + solve + D2H inside the timed region); e2e.formal_input = the same through
             cmpc_solve_host with the full (p, lbg, ubg, x0) arrays
  roofline : executed FP64 flop / kernel time against the measured FP64 FMA peak (the kernel is FP64 CUDA-core
             bound; HBM numbers are given beside it to show that HBM is not the bound)
  cpu_baseline : the CPU oracle (restatement of the IPOPT solve) on the host cores, bounded sample, rank 0, N = 1 only
  north_star_configs : BASELINE.json configs[2] (65536 ergoCub instances in total, STRONG-scaled over the ranks, at tol 1e-8 and
             at the ini's 1e-4) and configs[3] (4096 closed-loop rollouts x 100 ticks in total), measured in the same run
  per_rank : kernel time and iteration statistics of every rank (the solve needs no exchange; ONE all_gather of the
             per-instance results ends the timed region)
--impl reference times the reference's CPU path (the oracle port: IPOPT/CasADi are not installable here).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "paper_romualdi_2022_icra_centroidal-mpc-walking_b200"

# FP64 flop executed per interior-point iteration and knot by cmpc_solve_team_kernel (ncu
# smsp__sass_thread_inst_executed_op_{dfma x 2, dadd, dmul}_pred_on of one launch / (iterations x knots), see profiles/ and
# DESIGN.md "flop accounting"; the per-iteration-and-knot work does not depend on the workload: the same code runs for every
# robot, horizon and batch), and the canonical dense figure of SURVEY.md 8(d)
# Round 2: the stage factorisation and the SYRK run on the FP64 tensor cores: 1.796e5 = (2 x 5.811e9 DFMA + 1.510e9 DADD +
# 1.969e9 DMUL thread instructions + 1.7288e10 tensor-core flop [sm__ops_path_tensor_src_fp64.sum]) / (12024 iterations x 15
# knots) of the round's last capture r2_prof_final3 (profiles/r2_notes.md; the capture before the last chain / flat-loop changes
# gave 1.839e5); 9.6e4 of it are DMMA flop on padded 8 x 8 tiles (round 1 executed 1.19e5 on 3 x 3 tiles that skipped
# structural zeros).  monotone: round 1's figure + the same DMMA surplus (not re-captured).
FLOP_EXEC_PER_ITER_KNOT = {"monotone": 1.87e5, "mehrotra": 1.796e5}
FLOP_CANON_PER_ITER_KNOT = 422275.0
# measured DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum), keyed by (workload, batch, strategy)
DRAM_TRAFFIC_PER_LAUNCH = {("icub3_b1024", 1024, "mehrotra"): 1.219e10}   # r2_prof_final3: 7.95 GB read + 4.24 GB written
# ipopt_tolerance of the robot ini each workload is built from (config/robots/<robot>/centroidal_mpc.ini)
INI_TOLERANCE = {"icub3_b1024": 1e-8, "ergocub_b65536": 1e-4}
FP64_PEAK_FALLBACK_TFLOPS = 37.0  # vendor figure (HGX B200 296 TF / 8); used only if the live DFMA probe fails


def workload(name: str, pkg, wl, seed: int, batch: int = 0):
    if name == "icub3_b1024":
        cfg = pkg.icub3_config()
        w = wl.walk_batch(N=15, dT=0.1, B=batch or 1024, seed=seed, state_noise=1.0, step_adjust=False, ticks=True)
        ocfg = dict(N=15, w_com=(1.0, 1.0, 200.0), w_pos=200.0, w_sym=0.0,
                    corners=[[(0.08, 0.03, 0), (0.08, -0.03, 0), (-0.08, -0.03, 0), (-0.08, 0.03, 0)]] * 2)
    elif name == "ergocub_b65536":
        cfg = pkg.ergocub_config()
        w = wl.walk_batch(N=12, dT=0.1, B=batch or 65536, seed=seed, state_noise=2.0, yaw_range=0.3, step_adjust=True, ticks=True)
        ocfg = dict(N=12, w_pos=2000.0)
    else:
        raise SystemExit(f"unknown workload {name}")
    return cfg, w, ocfg


def config_dict(name: str, B: int, N: int, tol: float):
    """static description of the workload: the same dict on both arms (--impl ours / reference)"""
    return {"workload": name, "instances_per_gpu": B, "horizon_knots": N, "ipopt_tolerance": tol,
            "robot_ini": "iCubGazeboV3/centroidal_mpc.ini" if name.startswith("icub3") else "ergoCubGazeboV1_1/centroidal_mpc.ini",
            "cold_start": True,
            "l2": "GPU arm: the timed steps cycle through 8 different input batches (8 x 25 MB) and one 332 MB scratch arena per "
                  "handle, a working set several times the 126 MB L2 (with --pipeline 1 a 256 MB flush write runs between the "
                  "steps as well); CPU arm: not applicable"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): one nvidia-smi process in loop
    mode (-lms 50), its lines are collected while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.strip().split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass
        finally:
            if self.proc is not None:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(ocfg, w, sample: int, threads: int, tol: float):
    from oracle import oracle as om
    om.build()
    O = om.Oracle()
    cfg = om.make_cfg(**ocfg)
    t0 = time.perf_counter()
    x, lam, st = O.solve_batch(cfg, w["p"][:sample], w["lbg"][:sample], w["ubg"][:sample], w["x0"][:sample],
                               threads=threads, opts=O.default_opts(tol=tol))
    dt = time.perf_counter() - t0
    ok = sum(1 for s in st if s.status == 0)
    its = float(np.mean([s.iters for s in st]))
    cpu_oracle_rate.last = (x, np.array([s.obj for s in st]), np.array([s.status for s in st]))   # for the parity report
    return sample / dt, dt, ok, its


def foot_wrenches(L, x, p, corners):
    """(B, N, 2, 6) per-foot resultant force and torque (about the foot origin) of the corner forces at every knot"""
    B, N = x.shape[0], L.N
    out = np.zeros((B, N, 2, 6))
    cr = np.asarray(corners, dtype=np.float64)
    for k in range(N):
        for c in range(2):
            R = p[:, L.p_rot(c, k):L.p_rot(c, k) + 9].reshape(B, 3, 3).transpose(0, 2, 1)
            en = p[:, L.p_en(c, k)][:, None]
            for j in range(4):
                f = en * x[:, L.x_frc(c, j, k):L.x_frc(c, j, k) + 3]
                arm = R @ cr[j]
                out[:, k, c, :3] += f
                out[:, k, c, 3:] += np.cross(arm, f)
    return out


def parity_report(x_gpu, obj_gpu, status_gpu, L, p, corners):
    """north_star: "the fraction of instances landing in the same local optimum is reported" -- GPU solutions of the timed
    workload (default barrier update) against the oracle's solutions of the same instances (IPOPT's default path)."""
    xo, obj_o, st_o = cpu_oracle_rate.last
    n, N = xo.shape[0], L.N
    both = (st_o == 0) & (status_gpu[:n] == 0)
    rel = lambda a, b: np.max(np.abs(a - b), axis=1) / np.maximum(1.0, np.max(np.abs(b), axis=1))  # noqa: E731
    st = np.arange(0, 9 * (N + 1))                                # com, dcom, h trajectories
    fs = np.concatenate([np.arange(L.x_pos(c, 0), L.x_pos(c, 0) + 3 * (N + 1)) for c in range(2)])   # footsteps
    frc = np.concatenate([np.arange(L.x_frc(c, 0, 0), L.x_frc(c, 0, 0) + 12 * N) for c in range(2)])  # corner forces
    xg = x_gpu[:n]
    d_obj = np.abs(obj_gpu[:n] - obj_o) / np.maximum(1.0, np.abs(obj_o))
    same_state = both & (d_obj <= 1e-6) & (rel(xg[:, st], xo[:, st]) <= 1e-5) & (rel(xg[:, fs], xo[:, fs]) <= 1e-5)
    wg, wo = foot_wrenches(L, xg, p[:n], corners).reshape(n, -1), foot_wrenches(L, xo, p[:n], corners).reshape(n, -1)
    same_wrench = same_state & (rel(wg, wo) <= 1e-5)
    same_all = same_state & (rel(xg[:, frc], xo[:, frc]) <= 1e-5)
    return {"instances": int(n), "both_converged": int(both.sum()),
            "same_objective_trajectories_footsteps": int(same_state.sum()),
            "same_per_foot_wrenches_too": int(same_wrench.sum()), "same_corner_forces_too": int(same_all.sum()),
            "tolerances": "objective 1e-6 relative, scaled inf-norm 1e-5", "against": "oracle, IPOPT's default (monotone) path",
            "note": "per-foot wrench = resultant force and torque of a foot's four corner forces at every knot (what the dynamics "
                    "and the reference's whole-body layer see); with contact_force_symmetry_weight 0 (iCub3) the split over the "
                    "corners is not unique",
            "max_rel_objective_difference": float(d_obj[both].max()) if both.any() else None}


def run_reference(args):
    """--impl reference: the reference's CPU path for this metric = the oracle port (IPOPT restatement) on all host
    cores.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    pkg_wl = importlib.import_module(PKG + ".workloads")

    class _P:  # config factory without touching the CUDA library
        @staticmethod
        def icub3_config():
            return None

        @staticmethod
        def ergocub_config():
            return None
    _, w, ocfg = workload(args.workload, _P, pkg_wl, seed=0, batch=getattr(args, "batch", 0))
    cores = os.cpu_count() or 1
    sample = min(w["p"].shape[0], max(64, 16 * cores))
    times = []
    for s in range(args.warmup + args.steps):
        rate, dt, ok, its = cpu_oracle_rate(ocfg, w, sample, cores, args.tol)
        if s >= args.warmup:
            times.append(dt)
    t = float(np.sum(times))
    value = sample * len(times) / t
    line = {"impl": "reference", "metric": "batched centroidal-MPC solves/sec", "value": value, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, w["p"].shape[0], ocfg["N"], args.tol),
            "solve_stats": {"mu_strategy": "monotone (IPOPT's default)", "converged": f"{ok}/{sample}", "mean_iterations": its,
                            "note": "each step = bounded sample of the workload (the first instances)"},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} instances of {args.workload} per step, tol {args.tol:g}, "
                                       f"{ok}/{sample} converged, mean {its:.1f} iterations"},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def max_over_ranks(v: float, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()


def run_config3(pkg, wl, dev, local, rank, world, flush, total=65536):
    """BASELINE configs[2]: 65536 ergoCub instances (step adjustment and friction cones active, state noise x 2, footstep yaw
    +-0.3) STRONG-scaled: rank r solves total / world of them.  Device-resident and end-to-end (cmpc_solve_host from pinned
    host buffers) solves/s at tol 1e-8 and at the ini's ipopt_tolerance 1e-4."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    B = total // world
    w = wl.walk_batch(N=12, dT=0.1, B=B, seed=1000 + rank, state_noise=2.0, yaw_range=0.3, step_adjust=True, ticks=True)
    tens = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    d_p, d_lbg, d_ubg, d_x0 = tens(w["p"]), tens(w["lbg"]), tens(w["ubg"]), tens(w["x0"])
    d_x = d_x0.clone()
    h_tk = pin(w["ticks"])
    m = d_lbg.shape[1]
    d_lam = torch.zeros(B, m, dtype=torch.float64, device=dev)
    h_obj = torch.zeros(B, dtype=torch.float64).pin_memory()
    h_st, h_it = torch.zeros(B, dtype=torch.int32).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    out = {"workload": "ergocub_b65536", "instances_total": B * world, "instances_per_gpu": B, "scaling": "strong",
           "horizon_knots": 12, "robot_ini": "ergoCubGazeboV1_1/centroidal_mpc.ini", "steps": 2, "warmup": 1}
    for tol in (1e-8, 1e-4):
        cfg = pkg.ergocub_config()
        cfg.device, cfg.ipopt_tolerance = local, tol
        solver = pkg.BatchedCentroidalMPC(cfg)
        ev = []
        for s in range(3):
            flush.zero_()
            d_x.copy_(d_x0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if s == 1:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
            a.record()
            obj, status, iters, _ = solver.solve(d_p, d_lbg, d_ubg, d_x, d_lam)
            b.record()
            if s >= 1:
                ev.append((a, b))
        torch.cuda.synchronize()
        t_own = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
        t_dev = max_over_ranks(t_own, dev, world)
        st_h, it_h = status.cpu().numpy(), iters.cpu().numpy()
        # end to end: cmpc_solve_ticks_host with pinned host buffers (tick records up, populate + solve, x / obj / status /
        # iterations back, synchronisation: all inside the timed region)
        h_xs = [torch.zeros(B, d_x0.shape[1], dtype=torch.float64).pin_memory() for _ in range(3)]
        solver.lib.cmpc_solve_ticks_host(solver.handle, B, vp(h_tk), 0, vp(h_xs[0]), None, vp(h_obj), vp(h_st), vp(h_it))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for j in (1, 2):
            rc = solver.lib.cmpc_solve_ticks_host(solver.handle, B, vp(h_tk), 0, vp(h_xs[j]), None, vp(h_obj), vp(h_st), vp(h_it))
            assert rc == 0, rc
        assert int((h_st.numpy() == 0).sum()) == int((st_h == 0).sum())
        t_e2e = max_over_ranks(time.perf_counter() - t0, dev, world)
        conv, it_sum = sum_over_ranks([float((st_h == 0).sum()), float(it_h.sum())], dev, world)
        it_max = max_over_ranks(float(it_h.max()), dev, world)
        out[f"tol_{tol:g}"] = {"solves_per_s": world * B * len(ev) / t_dev, "ms_per_step": 1e3 * t_dev / len(ev),
                               "e2e_solves_per_s": world * B * 2 / t_e2e, "converged": f"{int(conv)}/{B * world}",
                               "mean_iterations": it_sum / (B * world), "max_iterations": int(it_max),
                               "kernel_ms_this_rank": 1e3 * t_own / len(ev)}
        solver.close()
    return out


def run_config4(pkg, dev, local, rank, world, total=4096, ticks=100):
    """BASELINE configs[3]: 4096 closed-loop rollouts (10 s of walking = 100 MPC ticks, one push per rollout, warm-started shift
    every tick) in total, rank r rolls out total / world of them; no exchange during the rollouts, ONE gather at the end."""
    import torch
    import torch.distributed as dist
    R = importlib.import_module(PKG + ".rollout")
    sharding = importlib.import_module(PKG + ".sharding")
    B = total // world
    out = {"rollouts_total": B * world, "rollouts_per_gpu": B, "ticks": ticks, "scaling": "strong", "robot_ini": "ergoCubGazeboV1_1/centroidal_mpc.ini",
           "driver": "rollout.closed_loop_rollout_device: six library kernels per tick (tick records, populate, shift, solve, plant, "
                     "feedback), ticks 1 .. T - 1 = one captured CUDA graph replayed; wall time includes set-up, capture and the final gather"}
    for tol in (1e-4, 1e-8):
        cfg = pkg.ergocub_config()
        cfg.device, cfg.ipopt_tolerance = local, tol
        solver = pkg.BatchedCentroidalMPC(cfg)
        R.closed_loop_rollout_device(solver, B=B, ticks=4, seed=rank)   # warm-up (allocator, kernels, graph capture)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # the same rollout three times (bitwise identical results): 2.5 s of back-to-back launches are long enough to catch
        # stalls that do not come from this code (profiles/r2_notes.md section 10), so every run is listed and the median counts
        runs = []
        for _ in range(3):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            res = R.closed_loop_rollout_device(solver, B=B, ticks=ticks, seed=100 + rank, time_device=True)
            local_stats = torch.from_numpy(np.stack([res["converged_ticks"], res["iterations"]], axis=1).astype(np.float64)).to(dev)
            allres = sharding.gather_results(local_stats, world)
            torch.cuda.synchronize()
            wall = max_over_ranks(time.perf_counter() - t0, dev, world)
            runs.append((max_over_ranks(res["device_ms"] * 1e-3, dev, world), wall))
        runs.sort()
        t_dev, wall = runs[1]
        allres = allres.cpu().numpy()
        n = B * world
        out[f"tol_{tol:g}"] = {"mpc_ticks_per_s_device": n * ticks / t_dev, "mpc_ticks_per_s_wall": n * ticks / wall,
                               "wall_s": wall, "device_s": t_dev, "device_s_runs": [round(r[0], 4) for r in runs],
                               "converged_ticks": f"{int(allres[:, 0].sum())}/{n * ticks}",
                               "iterations_per_tick": float(allres[:, 1].sum() / (n * ticks))}
        solver.close()
    return out


def host_operator_latency(samples=24):
    """BASELINE configs[0] through the drop-in class: one CentroidalMPC object initialised from the iCub3 ini, straight-walk
    contact list, cold start every tick (the ini does not enable the warm start): wall time of advance() = input population,
    H2D, solve, D2H, output unpacking."""
    H = importlib.import_module(PKG + ".host")
    wl = importlib.import_module(PKG + ".workloads")
    lay = importlib.import_module(PKG + ".layout")
    ini = os.path.join(ROOT, "tests", "data", "icub3", "centroidal_mpc_walking.ini")
    m = H.CentroidalMPCHost(ini, "CENTROIDAL_MPC")
    N, L = m.N, lay.Layout(m.N)
    lists = H.walk_contact_lists(0, n_steps=8 + samples // 8)
    lat, its = [], []
    for j in range(samples + 3):
        w = wl.walk_batch(N=N, dT=m.dT, B=1, seed=j, phase=j, state_noise=0.0, step_adjust=True)
        p, g0 = w["p"][0], L.p_glob()
        m.set_state(p[g0:g0 + 3], p[g0 + 3:g0 + 6], p[g0 + 6:g0 + 9])
        m.set_reference_trajectory(p[L.p_comref(0):L.p_comref(0) + 3 * (N + 1)], p[L.p_href(0):L.p_href(0) + 3 * (N + 1)])
        m.set_contact_phase_list(lists)
        t0 = time.perf_counter()
        ok = m.advance()
        dt = time.perf_counter() - t0
        if not ok:
            return {"error": m.last_error()}
        if j >= 3:
            lat.append(1e3 * dt)
            its.append(m.stats()[1])
    m.close()
    return {"p50_ms": float(np.median(lat)), "p95_ms": float(np.percentile(lat, 95)), "samples": len(lat),
            "mean_iterations": float(np.mean(its)),
            "api": "CentroidalMPC::advance() of the C++ drop-in class (iCub3 ini, straight-walk contact list, state on the reference, cold start)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="icub3_b1024")
    ap.add_argument("--tol", type=float, default=None,
                    help="ipopt_tolerance; default = the value of the workload's robot ini (iCubGazeboV3: 1e-8, the BLF default, "
                         "its ini leaves the key commented out; ergoCubGazeboV1_1: 1e-4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the north-star configs 3 / 4 and the host-operator latency")
    ap.add_argument("--gather", default="end", choices=["end", "per-step"],
                    help="the one exchange of the sharded job: a single all_gather after the last step (default) or one per step on a "
                         "side stream (round 1; kept for the A/B of DESIGN.md section 7)")
    ap.add_argument("--team", type=int, default=0, help="threads per instance (32/64/96/128), 0 = library default")
    ap.add_argument("--lockstep", type=int, default=0, help="teams per CTA walking in lock-step (1, 3, 7; 0 = library default)")
    ap.add_argument("--groups", type=int, default=0, help="independent lock-step groups per CTA (0 = library default)")
    ap.add_argument("--identical", action="store_true", help="experiment: every instance is a copy of instance 0")
    ap.add_argument("--ctas", type=int, default=0, help="resident teams per SM (0 = occupancy)")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (0 = the workload's own size)")
    ap.add_argument("--batches", type=int, default=8, help="different synthetic batches per rank the timed steps cycle through")
    ap.add_argument("--pipeline", type=int, default=3,
                    help="handles (each with its own stream and scratch arena) the timed steps alternate between: 3 (default; 2 and 4 measured too) lets the "
                         "straggler tail of one batch overlap the start of the next and the copies of the end-to-end path overlap the "
                         "solves; 1 = one handle, one stream, L2 flush between steps (rounds 1 and 2a)")
    ap.add_argument("--mu-strategy", default="mehrotra", choices=["mehrotra", "monotone"],
                    help="barrier update of the solve: Mehrotra predictor-corrector (library default) or IPOPT's monotone update")
    args = ap.parse_args()
    if args.tol is None:
        args.tol = INI_TOLERANCE[args.workload]
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pkg = importlib.import_module(PKG)
    wl = importlib.import_module(PKG + ".workloads")
    sharding = importlib.import_module(PKG + ".sharding")
    # every rank its own instances; the timed steps cycle through NB different batches (seeds rank * NB .. rank * NB + NB - 1):
    # a single-wave launch lasts as long as its unluckiest SM (seven instances sharing lock-step groups), which makes the time
    # of ONE batch of 1024 vary by +-5 % with the draw (16.5 ms for seed 0, 18.2 ms for seed 1 at the same mean and maximum
    # iteration count); a number measured on one draw is luck, and the ratio of two such numbers is not a scaling efficiency
    NB = max(1, min(args.steps, args.batches))
    ws = [workload(args.workload, pkg, wl, seed=rank * NB + j, batch=args.batch) for j in range(NB)]
    cfg, w, ocfg = ws[0]
    if args.identical:
        for _, wj, _ in ws:
            for key in ("p", "lbg", "ubg", "x0"):
                wj[key] = np.repeat(wj[key][:1], wj[key].shape[0], axis=0)
    cfg.device = local
    cfg.ipopt_tolerance = args.tol
    cfg.threads_per_instance = args.team
    cfg.ctas_per_sm = args.ctas
    cfg.teams_per_cta = args.lockstep
    cfg.lockstep_groups = args.groups
    cfg.mu_strategy = pkg.MU_MEHROTRA if args.mu_strategy == "mehrotra" else pkg.MU_MONOTONE
    solver = pkg.BatchedCentroidalMPC(cfg)
    P = max(1, args.pipeline)
    solvers = [solver] + [pkg.BatchedCentroidalMPC(cfg) for _ in range(P - 1)]
    B, N = w["p"].shape[0], cfg.horizon
    n, m, npar = solver.L.n, solver.L.m, solver.L.np
    tens = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    dws = [tuple(tens(wj[k]) for k in ("p", "lbg", "ubg", "x0")) for _, wj, _ in ws]   # device-resident inputs of every batch
    d_p, d_lbg, d_ubg, d_x0 = dws[0]
    d_x = d_x0.clone()
    d_lam = torch.zeros(B, m, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2
    d_xs = [d_x] + [d_x0.clone() for _ in range(P - 1)]
    d_lams = [d_lam] + [torch.zeros_like(d_lam) for _ in range(P - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(P)] if P > 1 else [torch.cuda.current_stream()]

    # The only exchange of the sharded job is the gather of the per-instance results.  Default: the results of
    # every step stay on the device and ONE all_gather follows the last step ("once per batch / rollout"): no NCCL kernel is
    # resident while a solve runs.  --gather per-step = round 1: one all_gather per step on a side stream; its kernel waits on
    # the SMs for the slowest peer while the next solve's persistent CTAs (one per SM, 222 KB of shared memory each) want every
    # SM: that was the unexplained 8-10 % of round 1's multi-GPU runs.
    per_step = args.gather == "per-step"
    gather = sharding.AsyncGather(world) if per_step else None
    results = torch.zeros(args.steps, B, 3, dtype=torch.float64, device=dev)

    for j in range(max(args.warmup, P)):
        flush.zero_()
        bp, bl, bu, bx = dws[j % NB]
        with torch.cuda.stream(streams[j % P]):    # the stream the handle is timed on (first use of a stream is not free)
            d_xs[j % P].copy_(bx)
            obj, status, iters, _ = solvers[j % P].solve(bp, bl, bu, d_xs[j % P], d_lams[j % P])
            if not per_step:
                results[0] = sharding.pack_results(obj, status, iters)
        torch.cuda.synchronize()
        if per_step:
            gather.submit(sharding.pack_results(obj, status, iters))
    if per_step:
        gather.wait()
    else:
        sharding.gather_results(sharding.pack_results(obj, status, iters), world)   # NCCL communicator warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # nvidia-smi start-up (NVML initialisation perturbs the device for a few ms: kept out of the timed region) -- spent on more
    # untimed warm-up launches, not asleep: after 0.2 s of idling the first timed launch ran at ramping clocks (41 ms against 15)
    # (until the sampler has delivered its first rows, i.e. nvidia-smi is up and looping; 3 s at most)
    t_ready = time.perf_counter() + 3.0
    j = 0
    while time.perf_counter() < t_ready and (len(sampler.rows) < 3 or j < P):
        bp, bl, bu, bx = dws[j % NB]
        with torch.cuda.stream(streams[j % P]):
            d_xs[j % P].copy_(bx)
            solvers[j % P].solve(bp, bl, bu, d_xs[j % P], d_lams[j % P])
        torch.cuda.synchronize()
        j += 1
    if world > 1:
        dist.barrier()
    sampler.rows.clear()
    launches0 = sum(sv.launch_count() for sv in solvers)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    span = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    if P == 1:
        for s in range(args.steps):
            flush.zero_()                      # L2 flush between timed iterations (outside the step's event pair)
            bp, bl, bu, bx = dws[s % NB]
            ev[s][0].record()
            d_x.copy_(bx)
            kev[s][0].record()
            obj, status, iters, _ = solver.solve(bp, bl, bu, d_x, d_lam)
            kev[s][1].record()
            if per_step:
                gather.submit(sharding.pack_results(obj, status, iters))
            else:
                results[s] = sharding.pack_results(obj, status, iters)
            ev[s][1].record()
    else:
        # step s runs on handle s % P and stream s % P: the launches of consecutive steps overlap on the device (a persistent CTA
        # of step s + 1 starts on every SM whose seven instances of step s are done), every step is still one batch solved by
        # one launch.  No flush kernel (it would serialise the streams): the steps cycle through NB different input batches and
        # P scratch arenas, a working set several times the L2 (see config.l2).
        main = torch.cuda.current_stream()
        span[0].record(main)
        for st in streams:
            st.wait_event(span[0])
        for s in range(args.steps):
            bp, bl, bu, bx = dws[s % NB]
            with torch.cuda.stream(streams[s % P]):
                ev[s][0].record()
                d_xs[s % P].copy_(bx)
                kev[s][0].record()
                obj, status, iters, _ = solvers[s % P].solve(bp, bl, bu, d_xs[s % P], d_lams[s % P])
                kev[s][1].record()
                results[s] = sharding.pack_results(obj, status, iters)
                ev[s][1].record()
        for s in range(max(0, args.steps - P), args.steps):
            main.wait_event(ev[s][1])
    tail = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    tail[0].record()
    if per_step:
        gathered = gather.wait()           # the timed region ends when every gather has landed
        n_gathered = sum(g.shape[0] for g in gathered)
    else:
        gathered = sharding.gather_results(results.view(args.steps * B, 3), world)
        n_gathered = gathered.shape[0]
    tail[1].record()
    span[1].record()
    torch.cuda.synchronize()
    assert n_gathered == args.steps * world * B
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sum(sv.launch_count() for sv in solvers) - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if P == 1:
        t_steps = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
        t_dev = t_steps + tail[0].elapsed_time(tail[1]) * 1e-3
        t_kernel = sum(a.elapsed_time(b) for a, b in kev) * 1e-3
    else:
        # overlapping launches: the device time of the job is the span from the first step's start to the end of the gather;
        # a launch's share of the device is that span / launches (its own start-to-end time on its stream includes the wait for
        # the SMs the previous launch still holds and is reported separately)
        t_dev = span[0].elapsed_time(span[1]) * 1e-3
        t_steps = span[0].elapsed_time(tail[0]) * 1e-3
        t_kernel = t_steps
    t_max = max_over_ranks(t_dev, dev, world)
    if per_step:
        res_all = torch.cat([sharding.pack_results(obj, status, iters)] * args.steps).view(args.steps, B, 3)   # last step only
    else:
        res_all = results
    status_all, iters_all = res_all[:, :, 1].cpu().numpy(), res_all[:, :, 2].cpu().numpy()     # (steps, B) of this rank
    status_h, iters_h = status_all.reshape(-1), iters_all.reshape(-1)
    conv = int((status_h == 0).sum())
    total_iters = int(iters_h.sum())
    step_ms = [a.elapsed_time(b) for a, b in kev]
    # per-rank record: kernel time per launch, own steps without the final gather, iteration statistics
    mine = torch.tensor([1e3 * t_kernel / args.steps, 1e3 * t_steps / args.steps, float(iters_h.mean()), float(iters_h.max()),
                         float(conv)], dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu().numpy()
    else:
        allr = mine.cpu().numpy()[None]

    # ---- single-solve latency (the second half of BASELINE.json's metric): batch of ONE instance, device resident, p50
    lat = []
    if rank == 0:
        x1 = d_x0[:1].clone()
        for j in range(35):
            x1.copy_(d_x0[:1])
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            solver.solve(d_p[:1], d_lbg[:1], d_ubg[:1], x1, d_lam[:1])
            b.record()
            torch.cuda.synchronize()
            if j >= 3:
                lat.append(a.elapsed_time(b))
    if world > 1:
        dist.barrier()

    # ---- end to end through the host-pointer C-ABI calls with pinned host buffers
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    import ctypes as C
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    h_obj = torch.zeros(B, dtype=torch.float64).pin_memory()
    h_st, h_it = torch.zeros(B, dtype=torch.int32).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()
    h_x = torch.zeros(B, n, dtype=torch.float64).pin_memory()
    k_e2e = max(2, min(args.steps, 5))
    # (1) the per-tick call of the host class: tick records up, x / obj / status / iterations back (cold start: no multipliers)
    h_tk = [pin(wj["ticks"]) for _, wj, _ in ws[:5]]
    order = [(j + 1) % len(h_tk) for j in range(k_e2e - 1)] + [0]   # the last timed step solves batch 0 into h_x (parity report)
    h_xs = [torch.zeros(B, n, dtype=torch.float64).pin_memory() for _ in order[:-1]] + [h_x]

    def ticks_step(hx, htk, slot=0, outs=None):
        o, st_, it_ = outs or (h_obj, h_st, h_it)
        rc = solvers[slot].lib.cmpc_solve_ticks_host(solvers[slot].handle, B, vp(htk), 0, vp(hx), None, vp(o), vp(st_), vp(it_))
        assert rc == 0, rc
    # the end-to-end path keeps more calls in flight than the device-resident loop (every call ends with a host synchronisation:
    # a handle idles between the return of one call and the arrival of the next): 2 P handles / host threads
    E = 1 if P == 1 else 2 * P
    while len(solvers) < E:
        solvers.append(pkg.BatchedCentroidalMPC(cfg))
    for slot in range(E):
        ticks_step(h_x, h_tk[0], slot)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if P == 1:
        t0 = time.perf_counter()
        for j in range(k_e2e):
            ticks_step(h_xs[j], h_tk[order[j]])
        torch.cuda.synchronize()
        e2e_value = world * B * k_e2e / max_over_ranks(time.perf_counter() - t0, dev, world)
    else:
        # P host threads, one handle each (the C call releases the interpreter lock): step j is made by thread j % P.  Every step
        # is still one blocking cmpc_solve_ticks_host call: records up, populate + solve, results down, synchronised.
        k_e2e = E * max(2, min(args.steps, 12) // E)
        h_xs = [torch.zeros(B, n, dtype=torch.float64).pin_memory() for _ in range(k_e2e)]
        outs = [(torch.zeros(B, dtype=torch.float64).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory(),
                 torch.zeros(B, dtype=torch.int32).pin_memory()) for _ in range(E)]
        gate = threading.Barrier(E + 1)

        def worker(slot):
            torch.cuda.set_device(local)   # the current device is per host thread: without this the library's device guard would
            gate.wait()                    # restore "device 0" after every call and create a context there on ranks > 0
            for j in range(slot, k_e2e, E):
                ticks_step(h_xs[j], h_tk[j % len(h_tk)], slot, outs[slot])
        th = [threading.Thread(target=worker, args=(slot,)) for slot in range(E)]
        for t_ in th:
            t_.start()
        t0 = time.perf_counter()
        gate.wait()
        for t_ in th:
            t_.join()
        torch.cuda.synchronize()
        e2e_value = world * B * k_e2e / max_over_ranks(time.perf_counter() - t0, dev, world)
    ts_ = solver.lib.cmpc_tick_stride(N)
    h2d = 8 * B * ts_
    d2h = 8 * B * (n + 1) + 8 * B
    # (2) the same through cmpc_solve_host: the full formal input (p, lbg, ubg, x0) up, x and lam_g back
    h_p, h_lbg, h_ubg, h_x0 = (pin(w[k]) for k in ("p", "lbg", "ubg", "x0"))
    h_lam = torch.zeros(B, m, dtype=torch.float64).pin_memory()
    h_xf = [h_x0.clone().pin_memory() for _ in range(3)]

    def formal_step(hx, slot=0, outs=None, hl=None):
        o, st_, it_ = outs or (h_obj, h_st, h_it)
        sv = solvers[slot]
        rc = sv.lib.cmpc_solve_host(sv.handle, B, vp(h_p), vp(h_lbg), vp(h_ubg), vp(hx), vp(hl if hl is not None else h_lam), vp(o),
                                    vp(st_), vp(it_), 0)
        assert rc == 0, rc
    k_formal = 2
    if P == 1:
        formal_step(h_xf[0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for j in (1, 2):
            formal_step(h_xf[j])
        torch.cuda.synchronize()
        e2e_formal = world * B * 2 / max_over_ranks(time.perf_counter() - t0, dev, world)
    else:
        # same threads-and-handles arrangement as the tick-record path: E blocking calls in flight
        k_formal = 2 * E
        h_xf = [h_x0.clone().pin_memory() for _ in range(k_formal)]
        h_lams = [torch.zeros(B, m, dtype=torch.float64).pin_memory() for _ in range(E)]
        for slot in range(E):
            formal_step(h_xf[slot], slot, outs[slot], h_lams[slot])
            h_xf[slot].copy_(h_x0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        gate2 = threading.Barrier(E + 1)

        def fworker(slot):
            torch.cuda.set_device(local)
            gate2.wait()
            for j in range(slot, k_formal, E):
                formal_step(h_xf[j], slot, outs[slot], h_lams[slot])
        th = [threading.Thread(target=fworker, args=(slot,)) for slot in range(E)]
        for t_ in th:
            t_.start()
        t0 = time.perf_counter()
        gate2.wait()
        for t_ in th:
            t_.join()
        torch.cuda.synchronize()
        e2e_formal = world * B * k_formal / max_over_ranks(time.perf_counter() - t0, dev, world)
    # leave batch 0's results of the ticks path in the host buffers for the parity report
    ticks_step(h_x, h_tk[0])
    torch.cuda.synchronize()
    geometry = solver.geometry()
    peak_tf, peak_src = FP64_PEAK_FALLBACK_TFLOPS, "fallback: vendor 37 TFLOP/s FP64 (no measured FP64 peak in MEASURED_PEAKS.json)"
    if rank == 0:
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak = float(mp.get("hbm_gbs", 6551.0))
            if "fp64_tflops" in mp:
                peak_tf, peak_src = float(mp["fp64_tflops"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm_peak = 6650.0
        try:
            if peak_src.startswith("fallback"):
                peak_tf = solver.measure_fp64_peak()
                peak_src = "measured live: cmpc_measure_fp64_peak (8 independent DFMA chains/thread, 148x8 CTAs x 256 threads); ncu peak_sustained 9472 DFMA/clk = 37.2"
        except Exception:
            pass
    # one handle, one stream, flush between launches: the duration of a launch that has the device to itself
    iso = []
    for j in range(3):
        flush.zero_()
        bp, bl, bu, bx = dws[j % NB]
        d_x.copy_(bx)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        solver.solve(bp, bl, bu, d_x, d_lam)
        b.record()
        torch.cuda.synchronize()
        iso.append(a.elapsed_time(b))
    for sv in solvers:
        sv.close()

    # ---- the other north-star configurations, in the same run (every rank takes part)
    extras = {}
    if not args.no_extras:
        extras["config3_ergocub_65536_sharded"] = run_config3(pkg, wl, dev, local, rank, world, flush)
        extras["config4_closed_loop_4096x100"] = run_config4(pkg, dev, local, rank, world)
        if rank == 0:
            try:
                extras["config1_single_solve_host_operator"] = host_operator_latency()
            except Exception as e:  # the host operator library is optional for the headline number
                extras["config1_single_solve_host_operator"] = {"error": repr(e)}

    if rank == 0:
        value = world * B * args.steps / t_max
        kernel_s = t_kernel / args.steps
        flop_exec = total_iters / args.steps * N * FLOP_EXEC_PER_ITER_KNOT[args.mu_strategy]      # per launch (mean over the steps)
        flop_canon = total_iters / args.steps * N * FLOP_CANON_PER_ITER_KNOT
        achieved = flop_exec / kernel_s / 1e12
        alg_bytes = 8.0 * B * (npar + 2 * m + 2 * n + 2 * m)
        line = {
            "metric": "batched centroidal-MPC solves/sec", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, B, N, args.tol),
            "solve_stats": {"mu_strategy": args.mu_strategy, "converged": f"{conv}/{B * args.steps}",
                            "mean_iterations": total_iters / (B * args.steps), "max_iterations": int(iters_h.max()),
                            "solver_grid": geometry, "result_gather": args.gather, "different_batches": NB,
                            "pipeline": P,
                            "pipeline_note": "steps alternate between `pipeline` handles (own stream, own scratch arena): one batch = "
                                             "one launch as before, consecutive launches overlap on the device (the straggler tail of "
                                             "a single-wave batch no longer idles the SMs); --pipeline 1 = one handle, one stream",
                            "kernel_ms_of_every_step_rank0": [round(v, 3) for v in step_ms],
                            "kernel_ms_isolated_launch": [round(v, 3) for v in iso]},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": k_e2e, "host_threads": E,
                    "api": "cmpc_solve_ticks_host (pinned host buffers): tick records up, populate + solve on the device, x / obj / status / iterations back; one blocking call per step, `host_threads` threads with a handle each",
                    "formal_input": {"value": e2e_formal, "h2d_bytes_per_step": 8 * B * (npar + 2 * m + n),
                                     "d2h_bytes_per_step": 8 * B * (n + m + 1) + 8 * B, "steps": k_formal,
                                     "api": "cmpc_solve_host: p, lbg, ubg, x0 up, x / lam_g / obj / status / iterations back"}},
            "gpu_launches": int(launches),
            "latency": {"p50_single_solve_ms": float(np.median(lat)) if lat else None,
                        "p95_single_solve_ms": float(np.percentile(lat, 95)) if lat else None, "samples": len(lat),
                        "iterations": int(iters_all[0, 0]),
                        "e2e_host_operator_ms": (extras.get("config1_single_solve_host_operator") or {}).get("p50_ms"),
                        "note": "batch of one instance, device resident, cold start; batches of up to 4 instances per SM run on independent single-team CTAs (one per SM: a team of 256 threads; two: 128 threads without a register cap; four: 128 threads) unless a geometry is forced; e2e_host_operator_ms = CentroidalMPC::advance() of the drop-in class"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": DRAM_TRAFFIC_PER_LAUNCH.get((args.workload, B, args.mu_strategy)),
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu, profiles/); null where this "
                                         "exact (workload, batch, strategy) has not been captured",
                         "peak_source": peak_src, "kernel": "cmpc_solve_team_kernel",
                         "kernel_ms_per_launch": 1e3 * kernel_s,
                         "kernel_ms_note": "pipeline 1: CUDA events around every launch on its stream; pipeline > 1: launches overlap, "
                                           "the figure is the device span of the timed steps / launches (a launch's share of the "
                                           "device); solve_stats.kernel_ms_isolated_launch = a launch alone on the device",
                         "flop_executed_per_launch": flop_exec,
                         "flop_per_iteration_and_knot": FLOP_EXEC_PER_ITER_KNOT[args.mu_strategy],
                         "flop_canonical_dense_per_launch": flop_canon,
                         "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / kernel_s / 1e9,
                                 "peak_gbs": hbm_peak, "frac": alg_bytes / kernel_s / 1e9 / hbm_peak}},
            "per_rank": {"kernel_ms_per_launch": [float(v) for v in allr[:, 0]], "own_ms_per_step": [float(v) for v in allr[:, 1]],
                         "mean_iterations": [float(v) for v in allr[:, 2]], "max_iterations": [int(v) for v in allr[:, 3]],
                         "converged": [int(v) for v in allr[:, 4]]},
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall,
        }
        if extras:
            line["north_star_configs"] = extras
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = min(B, max(64, 64 * cores))   # about 10-30 core-seconds of CPU work
            rate, dt, ok, its = cpu_oracle_rate(ocfg, w, sample, cores, args.tol)
            line["cpu_baseline"] = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                                    "sample": f"first {sample} instances of the workload, {dt:.1f} s, {ok}/{sample} "
                                              f"converged, mean {its:.1f} iterations (oracle = IPOPT restatement)"}
            line["parity"] = parity_report(h_x.numpy(), h_obj.numpy(), h_st.numpy(), solver.L, w["p"], ocfg.get(
                "corners", [[(0.08, 0.01, 0), (0.08, -0.01, 0), (-0.08, -0.01, 0), (-0.08, 0.01, 0)]] * 2)[0])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
