"""Closed-loop batched rollouts (BASELINE.json config 4): every tick solves the MPC on the GPU, integrates the centroidal
dynamics under the knot-0 corner forces (RK4 at the whole-body rate, forces held: what the reference's WholeBodyQPBlock does at
src/centroidal-mpc-walking/src/WholeBodyQPBlock.cpp:1083-1090, 1150-1158 and feeds back at :1259-1262), warm-starts the next
solve with the shifted solution (cmpc_shift_warmstart) and replaces planned by actual footsteps (the role of
updateContactPhaseList, src/CentroidalMPCBlock.cpp:32-110).  Push perturbations act on the plant and are reported to the MPC
as external wrench at knot 0 (CentroidalMPCBlock.cpp:407; the reference ignores wrenches below 0.7 m/s^2,
WholeBodyQPBlock.cpp:1018).

The solve, the shift and the plant are CUDA kernels of libcmpc_b200.so; this driver only moves pointers and patches a few
entries of p / lbg / ubg with torch (plumbing).  Instances are independent: with torch.distributed initialised every rank
rolls out its own shard and the statistics are gathered once at the end (sharding.gather_results)."""
from __future__ import annotations

import numpy as np
import torch

from . import BatchedCentroidalMPC
from .layout import Layout
from .workloads import walk_batch

PUSH_THRESHOLD = 0.7  # WholeBodyQPBlock.cpp:1018


def closed_loop_rollout(solver: BatchedCentroidalMPC, B: int, ticks: int, seed: int = 0, dT: float = 0.1, wbc_dt: float = 0.002,
                        push_prob: float = 1.0, push_range=(1.0, 3.0), step_adjust: bool = True, yaw_range: float = 0.0,
                        time_device: bool = False):
    """returns dict of per-instance numpy arrays: converged ticks, iterations, max CoM tracking error, min CoM height;
    time_device = True adds "device_ms": CUDA-event time of the kernels of every tick (shift + solve + plant)"""
    N, dev = solver.N, solver.device
    L = Layout(N)
    rng = np.random.default_rng(seed)
    phase0 = rng.integers(0, 16, size=B)
    push_tick = np.where(rng.uniform(size=B) < push_prob, rng.integers(5, max(6, ticks - 5), size=B), -1)
    push_len = rng.integers(1, 3, size=B)                       # 0.1 - 0.2 s
    ang = rng.uniform(0, 2 * np.pi, size=B)
    mag = rng.uniform(*push_range, size=B)
    push_vec = np.stack([mag * np.cos(ang), mag * np.sin(ang), np.zeros(B)], axis=1)
    substeps = int(round(dT / wbc_dt))

    tens = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    state = None                                                # (B, 9) com, dcom, h of the plant
    foot = None                                                 # (B, 2, 3) actual foot positions
    d_x = d_lam = None
    conv = np.zeros(B, dtype=np.int64)
    iters_sum = np.zeros(B, dtype=np.int64)
    err_max = np.zeros(B)
    zmin = np.full(B, np.inf)
    g0 = L.p_glob()
    events = []
    for t in range(ticks):
        w = walk_batch(N=N, dT=dT, B=B, seed=seed, phase=phase0 + t, step_adjust=step_adjust, yaw_range=yaw_range)
        p, lbg, ubg = tens(w["p"]), tens(w["lbg"]), tens(w["ubg"])
        active = (push_tick >= 0) & (t >= push_tick) & (t < push_tick + push_len)
        ext = np.where(active[:, None], push_vec, 0.0)
        ext_mpc = np.where(np.linalg.norm(ext, axis=1, keepdims=True) >= PUSH_THRESHOLD, ext, 0.0)
        if state is None:
            state = p[:, g0:g0 + 9].clone()
            foot = torch.stack([p[:, L.p_cur(c):L.p_cur(c) + 3] for c in range(2)], dim=1).clone()
            d_x = tens(w["x0"])
            d_lam = torch.zeros(B, L.m, dtype=torch.float64, device=dev)
            warm = False
        else:
            if time_device:
                ev_shift = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev_shift[0].record()
            solver.shift_warmstart(d_x, d_lam)
            if time_device:
                ev_shift[1].record()
                events.append(ev_shift)
            warm = True
        # feedback: plant state, actual footsteps instead of planned ones, external force at knot 0
        p[:, g0:g0 + 9] = state
        lbg[:, 0:9] = state
        ubg[:, 0:9] = state
        p[:, L.p_extf(0):L.p_extf(0) + 3] = tens(ext_mpc)
        same, stance0 = tens(w["same_contact"]), tens(w["stance0"])
        for c in range(2):
            nom = p[:, L.p_nom(c, 0):L.p_nom(c, 0) + 3 * (N + 1)].view(B, N + 1, 3)
            delta = (foot[:, c] - nom[:, 0]) * stance0[c].unsqueeze(1)          # stance: the contact is where the foot landed
            nom += delta.unsqueeze(1) * same[c].unsqueeze(2)
            p[:, L.p_cur(c):L.p_cur(c) + 3] = foot[:, c]
            lbg[:, 9 + 3 * c:12 + 3 * c] = foot[:, c]
            ubg[:, 9 + 3 * c:12 + 3 * c] = foot[:, c]
        if time_device:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        obj, status, iters, _ = solver.solve(p, lbg, ubg, d_x, d_lam, warm_duals=warm)
        # plant: RK4 under the knot-0 forces (+ the push), then the next tick's foot positions from the MPC's own plan
        ext6 = torch.zeros(B, 6, dtype=torch.float64, device=dev)
        ext6[:, :3] = tens(ext)
        solver.rollout_plant(d_x, p, state, wbc_dt, substeps, ext=ext6)
        if time_device:
            ev[1].record()
            events.append(ev)
        for c in range(2):
            foot[:, c] = d_x[:, L.x_pos(c, 1):L.x_pos(c, 1) + 3]
        st = status.cpu().numpy()
        conv += st == 0
        iters_sum += iters.cpu().numpy()
        com = state[:, 0:3].cpu().numpy()
        ref = w["comref"][:, 1]                                   # the reference of the knot the plant has just reached
        err_max = np.maximum(err_max, np.linalg.norm(com[:, :2] - ref[:, :2], axis=1))
        zmin = np.minimum(zmin, com[:, 2])
    out = dict(converged_ticks=conv, iterations=iters_sum, com_err_max=err_max, com_z_min=zmin, push_tick=push_tick, ticks=ticks)
    if time_device:
        torch.cuda.synchronize()
        out["device_ms"] = float(sum(a.elapsed_time(b) for a, b in events))
    return out
