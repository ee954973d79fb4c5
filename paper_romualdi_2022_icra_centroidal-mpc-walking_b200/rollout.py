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


class WalkSchedule:
    """The walk schedule of workloads.walk_batch as a device-resident table: the formal inputs (p, lbg, ubg, x0) and the
    contact bookkeeping of an instance depend on its phase only (yaw 0, no state noise), are periodic in the phase with
    period P = 2 (double support + single support) knots from the second cycle on, up to a translation of every
    x coordinate by one stride per cycle.  Built once from walk_batch on the host (phases 0 .. 2 P - 1: start-up cycle and
    first steady cycle; the x-coordinate mask = entries that move by exactly one stride between phase ph and ph + P),
    evaluated per tick with two gathers on the device: no host-side generation and no H2D copy of the inputs per tick
    (the role of the planner + updateContactPhaseList feeding setContactPhaseList, CentroidalMPCBlock.cpp:596-609)."""

    def __init__(self, N: int, dT: float, device, step_adjust: bool = True, **kw):
        probe = walk_batch(N=N, dT=dT, B=1, phase=0, step_adjust=step_adjust, **kw)
        self.N, self.dT, self.device = N, dT, device
        ds = int(round(kw.get("ds_time", 0.3) / dT)); ss = int(round(kw.get("ss_time", 0.5) / dT))
        self.P = P = 2 * (ds + ss)
        self.stride = 2.0 * kw.get("step_length", 0.1)
        a = walk_batch(N=N, dT=dT, B=2 * P, phase=np.arange(2 * P), step_adjust=step_adjust, **kw)
        b = walk_batch(N=N, dT=dT, B=P, phase=np.arange(2 * P, 3 * P), step_adjust=step_adjust, **kw)
        to = lambda x, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(x)).to(device=device, dtype=dt)  # noqa: E731
        self.tab, self.mask = {}, {}
        for key in ("p", "lbg", "ubg", "x0", "comref"):
            t = a[key].reshape(2 * P, -1)
            d = b[key].reshape(P, -1) - t[P:]
            m = np.abs(d - self.stride) < 1e-12                     # x coordinates: one stride further every cycle
            assert np.all(m | (np.nan_to_num(d, nan=0.0, posinf=0.0, neginf=0.0) == 0.0)), key
            assert np.all(m == m[0]), key                          # the same entries for every phase
            self.tab[key], self.mask[key] = to(t), to(m[0].astype(np.float64))
        self.same = to(np.moveaxis(a["same_contact"], 1, 0), torch.bool)      # (2P, NC, N+1)
        self.stance0 = to(np.moveaxis(a["stance0"], 1, 0), torch.bool)        # (2P, NC)
        del probe

    def __call__(self, phase: torch.Tensor):
        """phase: (B,) int64 device tensor -> dict like walk_batch (device tensors; comref (B, N+1, 3))"""
        P = self.P
        steady = phase >= 2 * P
        idx = torch.where(steady, P + phase % P, phase)
        shift = torch.where(steady, (torch.div(phase, P, rounding_mode="floor") - 1).double() * self.stride,
                            torch.zeros_like(phase, dtype=torch.float64))
        out = {k: self.tab[k][idx] + shift[:, None] * self.mask[k][None, :] for k in self.tab}
        out["comref"] = out["comref"].view(-1, self.N + 1, 3)
        out["same_contact"] = self.same[idx].permute(1, 0, 2)      # (NC, B, N+1) as in walk_batch
        out["stance0"] = self.stance0[idx].permute(1, 0)           # (NC, B)
        return out


def rollout_tables(B: int, ticks: int, seed: int, N: int, dT: float, push_prob=1.0, push_range=(1.0, 3.0), yaw_range=0.0,
                   step_length=0.1, foot_y=0.08, ds_time=0.3, ss_time=0.5):
    """Host-side set-up of a batch of rollouts (numpy, seeded): the rollout records (phase of tick 0, push schedule), the
    footstep tables of the synthetic planner and the initial plant states.  Layouts: cmpc_rollout_layout (include/cmpc_b200.h)."""
    from . import load_library
    import ctypes as C
    rs, ms, ss_ = C.c_int(), C.c_int(), C.c_int()
    load_library().cmpc_rollout_layout(C.byref(rs), C.byref(ms), C.byref(ss_))
    RS, MAXS, ST = rs.value, ms.value, ss_.value
    rng = np.random.default_rng(seed)
    phase0 = rng.integers(0, 16, size=B)
    push_tick = np.where(rng.uniform(size=B) < push_prob, rng.integers(5, max(6, ticks - 5), size=B), -1)
    push_len = rng.integers(1, 3, size=B)                       # 0.1 - 0.2 s
    ang = rng.uniform(0, 2 * np.pi, size=B)
    mag = rng.uniform(*push_range, size=B)
    roll = np.zeros((B, RS))
    roll[:, 0] = phase0
    roll[:, 1], roll[:, 2] = push_tick, push_len
    roll[:, 3], roll[:, 4] = mag * np.cos(ang), mag * np.sin(ang)
    roll[:, 9] = np.inf                                         # running minimum of the CoM height
    roll[:, 10] = -np.inf                                       # running maximum of the ZMP excess
    idx = np.arange(MAXS)
    steps = np.zeros((B, 2, MAXS, ST))
    steps[:, 0, :, 0] = 2 * step_length * idx
    steps[:, 1, :, 0] = np.where(idx == 0, 0.0, step_length * (2 * idx - 1))
    steps[:, 0, :, 1], steps[:, 1, :, 1] = foot_y, -foot_y
    if yaw_range > 0:
        steps[:, :, :, 3] = rng.uniform(-yaw_range, yaw_range, size=(B, 2, MAXS))
    w0 = walk_batch(N=N, dT=dT, B=B, phase=phase0, step_length=step_length, foot_y=foot_y, ds_time=ds_time, ss_time=ss_time)
    g0 = Layout(N).p_glob()
    state = w0["p"][:, g0:g0 + 9].copy()
    assert int(phase0.max()) + ticks + N < 16 * (MAXS - 2), "the footstep table is too short for this rollout"
    return roll, steps, state


def closed_loop_rollout_device(solver: BatchedCentroidalMPC, B: int, ticks: int, seed: int = 0, dT: float = 0.1, wbc_dt: float = 0.002,
                               push_prob: float = 1.0, push_range=(1.0, 3.0), step_adjust: bool = True, yaw_range: float = 0.0,
                               time_device: bool = False, use_graph: bool = True, ds_time=0.3, ss_time=0.5, step_length=0.1):
    """The closed loop with every per-tick step on the device: cmpc_rollout_tick (planner + updateContactPhaseList + state
    feedback -> tick records), cmpc_populate (records -> p, lbg, ubg), cmpc_shift_warmstart, cmpc_solve_batched,
    cmpc_rollout_plant, cmpc_rollout_feedback (statistics, landed footsteps).  Six kernels per tick, no torch op and no host
    synchronisation inside the loop; ticks 1 .. T - 1 are ONE captured CUDA graph replayed T - 1 times (use_graph).  Footstep
    yaw is part of the footstep tables.  Returns the same statistics as closed_loop_rollout plus zmp_excess_max."""
    from . import WalkParams
    N, dev = solver.N, solver.device
    L = Layout(N)
    roll_h, steps_h, state_h = rollout_tables(B, ticks, seed, N, dT, push_prob, push_range, yaw_range, step_length=step_length,
                                              ds_time=ds_time, ss_time=ss_time)
    tens = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    roll, steps, state = tens(roll_h), tens(steps_h), tens(state_h)
    wp = WalkParams(int(round(ds_time / dT)), int(round(ss_time / dT)), step_length, 0.7, PUSH_THRESHOLD, 0.08, 0.03)
    mk = lambda n, dt=torch.float64: torch.zeros(B, n, dtype=dt, device=dev)  # noqa: E731
    tk = mk(solver.lib.cmpc_tick_stride(N))
    p, lbg, ubg, d_x, d_lam, ext6 = mk(L.np), mk(L.m), mk(L.m), mk(L.n), mk(L.m), mk(6)
    out = (torch.zeros(B, dtype=torch.float64, device=dev), torch.zeros(B, dtype=torch.int32, device=dev),
           torch.zeros(B, dtype=torch.int32, device=dev))
    substeps = int(round(dT / wbc_dt))

    def tick(first: bool):
        solver.rollout_tick(wp, 0 if first else -1, roll, state, steps, tk, ext6, step_adjust)
        solver.populate_into(tk, p, lbg, ubg, d_x if first else None)
        if not first:
            solver.shift_warmstart(d_x, d_lam)
        solver.solve(p, lbg, ubg, d_x, d_lam, warm_duals=not first, out=out)
        solver.rollout_plant(d_x, p, state, wbc_dt, substeps, ext=ext6)
        solver.rollout_feedback(wp, 0 if first else -1, d_x, p, state, out[1], out[2], roll, steps)

    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    stream = torch.cuda.Stream(device=dev) if use_graph else torch.cuda.current_stream(dev)
    stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(stream):
        ev[0].record()
        tick(True)
        if ticks > 1:
            if use_graph:
                # capture_begin / capture_end directly: the torch.cuda.graph() context manager synchronises, runs the garbage
                # collector and empties the allocator's cache on entry -- 0.3 - 0.9 s of host time in a process that holds tens
                # of GB of cached blocks, spent between the first tick and the first replay with the device idle
                # (profiles/r2_notes.md section 10).  Nothing is allocated through torch inside the captured tick.
                graph = torch.cuda.CUDAGraph()
                graph.capture_begin()
                try:
                    tick(False)
                finally:
                    graph.capture_end()
                # capturing does not execute: the captured tick is launched T - 1 times
                for _ in range(ticks - 1):
                    graph.replay()
            else:
                for _ in range(ticks - 1):
                    tick(False)
        ev[1].record()
    torch.cuda.current_stream(dev).wait_stream(stream)
    r = roll.cpu().numpy()
    res = dict(converged_ticks=r[:, 6].astype(np.int64), iterations=r[:, 7].astype(np.int64), com_err_max=r[:, 8], com_z_min=r[:, 9],
               zmp_excess_max=r[:, 10], push_tick=roll_h[:, 1].astype(np.int64), ticks=ticks, footsteps=steps.cpu().numpy())
    if time_device:
        torch.cuda.synchronize()
        res["device_ms"] = float(ev[0].elapsed_time(ev[1]))
    return res


def closed_loop_rollout(solver: BatchedCentroidalMPC, B: int, ticks: int, seed: int = 0, dT: float = 0.1, wbc_dt: float = 0.002,
                        push_prob: float = 1.0, push_range=(1.0, 3.0), step_adjust: bool = True, yaw_range: float = 0.0,
                        time_device: bool = False, host_schedule: bool = False):
    """returns dict of per-instance numpy arrays: converged ticks, iterations, max CoM tracking error, min CoM height;
    time_device = True adds "device_ms": CUDA-event time of the kernels of every tick (shift + solve + plant).
    The contact schedule of every tick comes from the device-resident WalkSchedule table (no host work, no H2D copy and no
    host synchronisation inside the loop); host_schedule = True (or yaw_range > 0: random footstep yaw is not tabulated)
    regenerates it with workloads.walk_batch on the host every tick, as the first version of this driver did."""
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
    host_schedule = host_schedule or yaw_range > 0

    tens = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    sched = None if host_schedule else WalkSchedule(N, dT, dev, step_adjust=step_adjust)
    d_phase0, d_push_tick, d_push_len, d_push_vec = tens(phase0), tens(push_tick), tens(push_len), tens(push_vec)
    push_big = d_push_vec.norm(dim=1, keepdim=True) >= PUSH_THRESHOLD
    state = None                                                # (B, 9) com, dcom, h of the plant
    foot = None                                                 # (B, 2, 3) actual foot positions
    d_x = d_lam = None
    conv = torch.zeros(B, dtype=torch.int64, device=dev)
    iters_sum = torch.zeros(B, dtype=torch.int64, device=dev)
    err_max = torch.zeros(B, dtype=torch.float64, device=dev)
    zmin = torch.full((B,), float("inf"), dtype=torch.float64, device=dev)
    g0 = L.p_glob()
    events = []
    for t in range(ticks):
        if host_schedule:
            w = walk_batch(N=N, dT=dT, B=B, seed=seed, phase=phase0 + t, step_adjust=step_adjust, yaw_range=yaw_range)
            w = {k: (tens(v) if isinstance(v, np.ndarray) else v) for k, v in w.items()}
        else:
            w = sched(d_phase0 + t)
        p, lbg, ubg = w["p"], w["lbg"], w["ubg"]
        active = (d_push_tick >= 0) & (t >= d_push_tick) & (t < d_push_tick + d_push_len)
        ext = torch.where(active[:, None], d_push_vec, torch.zeros_like(d_push_vec))
        ext_mpc = torch.where(push_big, ext, torch.zeros_like(ext))
        if state is None:
            state = p[:, g0:g0 + 9].clone()
            foot = torch.stack([p[:, L.p_cur(c):L.p_cur(c) + 3] for c in range(2)], dim=1).clone()
            d_x = w["x0"].clone()
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
        p[:, L.p_extf(0):L.p_extf(0) + 3] = ext_mpc
        same, stance0 = w["same_contact"], w["stance0"]
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
        ext6[:, :3] = ext
        solver.rollout_plant(d_x, p, state, wbc_dt, substeps, ext=ext6)
        if time_device:
            ev[1].record()
            events.append(ev)
        for c in range(2):
            foot[:, c] = d_x[:, L.x_pos(c, 1):L.x_pos(c, 1) + 3]
        conv += (status == 0)
        iters_sum += iters
        ref = w["comref"][:, 1]                                   # the reference of the knot the plant has just reached
        err_max = torch.maximum(err_max, (state[:, 0:2] - ref[:, 0:2]).norm(dim=1))
        zmin = torch.minimum(zmin, state[:, 2])
    out = dict(converged_ticks=conv.cpu().numpy(), iterations=iters_sum.cpu().numpy(), com_err_max=err_max.cpu().numpy(),
               com_z_min=zmin.cpu().numpy(), push_tick=push_tick, ticks=ticks)
    if time_device:
        torch.cuda.synchronize()
        out["device_ms"] = float(sum(a.elapsed_time(b) for a, b in events))
    return out
