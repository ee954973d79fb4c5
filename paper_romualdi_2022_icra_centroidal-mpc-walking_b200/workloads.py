"""Synthetic, seeded inputs for the centroidal-MPC solve (numpy only; no reference data, no network).

The reference gets (references, contact list, state) from its MANN planner and the robot
(/root/reference/src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:525-609); here they are synthesised as in
SURVEY.md 8(d): walk schedule W(N, dT) = double support 0.3 s -> right swing 0.5 s -> double support 0.3 s -> left swing
0.5 s ..., step length 0.1 m, feet at y = +-0.08 m, CoM reference at z = 0.7 m (CentroidalMPCBlock.cpp:534), zero
angular-momentum reference.  Everything is produced directly in the solver's formal input (p, lbg, ubg, x0) in CasADi
order (layout.py); the population rules are those of SURVEY.md 8(a) a-7.
"""
from __future__ import annotations

import numpy as np

from .layout import Layout, NC, NJ

INF = 1e20  # |bound| >= 1e19 is "no bound" (IPOPT nlp_upper_bound_inf)
GRAVITY = 9.80665

# step-adjustment boxes of the reference ini files, contact frame
# (config/robots/*/centroidal_mpc.ini: bounding_box_upper_limit / bounding_box_lower_limit)
BOX_UPPER = np.array([[0.01, 0.05, 0.0], [0.01, 0.00, 0.0]])
BOX_LOWER = np.array([[-0.01, -0.00, 0.0], [-0.01, -0.05, 0.0]])


def _rotz(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.zeros(yaw.shape + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 1, 0], R[..., 1, 1], R[..., 2, 2] = c, -s, s, c, 1.0
    return R


def walk_batch(N=12, dT=0.1, B=1, seed=0, phase=None, yaw_range=0.0, state_noise=0.0, step_adjust=True,
               step_length=0.1, foot_y=0.08, com_height=0.7, ds_time=0.3, ss_time=0.5, push=None, ticks=False):
    """Batch of B MPC instances on the walk schedule.  Returns dict(p, lbg, ubg, x0) of float64 arrays.

    phase:       None -> random phase offset in [0, period) knots per instance (seeded); int -> that offset for all.
    yaw_range:   footstep yaw ~ U(-yaw_range, yaw_range) per footstep.
    state_noise: scale s of com0 += U(-.03,.03)^3 s, dcom0 += U(-.2,.2)^3 s, h0 += U(-.05,.05)^3 s.
    step_adjust: False -> every step box is zero width (lower = upper = 0): footsteps cannot move.
    push:        optional (B,3) external force per unit mass applied at knot 0 (setState puts the wrench in column 0).
    ticks:       True -> also "ticks": the same instances as compact tick records (state, references, contact windows).
    """
    L = Layout(N)
    rng = np.random.default_rng(seed)
    ds, ss = int(round(ds_time / dT)), int(round(ss_time / dT))
    P = 2 * (ds + ss)
    if phase is None:
        ph = rng.integers(0, P, size=B)
    elif np.ndim(phase) == 0:
        ph = np.full(B, int(phase))
    else:
        ph = np.asarray(phase, dtype=np.int64).reshape(B)
    ell = ph[:, None] + np.arange(N + 1)[None, :]               # global knot index (B, N+1)
    q, cyc = ell % P, ell // P

    swing = np.zeros((NC, B, N + 1), dtype=bool)
    swing[1] = (q >= ds) & (q < ds + ss)                         # right foot swings first
    swing[0] = q >= P - ss
    # index of the footstep the foot stands on (stance) / stood on at lift-off (first swing knot) / goes to (swing)
    step = np.zeros((NC, B, N + 1), dtype=np.int64)
    step[1] = cyc + (q >= ds + ss) + ((q > ds) & (q < ds + ss))
    step[0] = cyc + (q > P - ss)
    nsteps = int(step.max()) + 2
    # footstep tables
    idx = np.arange(nsteps)
    fx = np.zeros((NC, nsteps))
    fx[1] = np.where(idx == 0, 0.0, step_length * (2 * idx - 1))
    fx[0] = 2 * step_length * idx
    fy = np.array([foot_y, -foot_y])
    yaw = rng.uniform(-yaw_range, yaw_range, size=(NC, B, nsteps)) if yaw_range > 0 else np.zeros((NC, B, nsteps))

    p = np.zeros((B, L.np))
    lbg = np.zeros((B, L.m))
    ubg = np.zeros((B, L.m))
    x0 = np.zeros((B, L.n))
    bi = np.arange(B)

    for c in range(NC):
        nom = np.zeros((B, N + 1, 3))
        nom[..., 0] = fx[c][step[c]]
        nom[..., 1] = fy[c]
        en = (~swing[c]).astype(np.float64)
        yk = np.take_along_axis(yaw[c], step[c], axis=1)         # (B, N+1)
        R = _rotz(np.where(swing[c], 0.0, yk))                   # identity on swing knots
        # current position: stance -> the contact; mid swing -> linear interpolation lift-off -> landing
        prog = np.where(c == 1, (q[:, 0] - ds) / ss, (q[:, 0] - (P - ss)) / ss)
        prev = np.maximum(step[c][:, 0] - 1, 0)
        first_swing_knot = swing[c][:, 0] & (prog <= 0)
        midswing = swing[c][:, 0] & ~first_swing_knot
        cur = nom[:, 0, :].copy()
        cur[midswing, 0] = (fx[c][prev] + prog * (fx[c][step[c][:, 0]] - fx[c][prev]))[midswing]
        for k in range(N):
            p[:, L.p_rot(c, k):L.p_rot(c, k) + 9] = R[:, k].transpose(0, 2, 1).reshape(B, 9)   # column major
            p[:, L.p_en(c, k)] = en[:, k]
        for k in range(N + 1):
            p[:, L.p_nom(c, k):L.p_nom(c, k) + 3] = nom[:, k]
        p[:, L.p_cur(c):L.p_cur(c) + 3] = cur
        lbg[:, 9 + 3 * c:12 + 3 * c] = cur
        ubg[:, 9 + 3 * c:12 + 3 * c] = cur
        # step-adjustment box rows of knot k act on pos_{k+1}
        initial_contact = (~swing[c][:, :1]) & (step[c] == step[c][:, :1]) & ~swing[c]          # (B, N+1)
        for k in range(N):
            lo = np.zeros((B, 3))
            up = np.zeros((B, 3))
            if step_adjust:
                free = swing[c][:, k]
                adj = ~free & ~initial_contact[:, k]
                lo[adj], up[adj] = BOX_LOWER[c], BOX_UPPER[c]
                lo[free], up[free] = -INF, INF
            p[:, L.p_upper(c, k):L.p_upper(c, k) + 3] = up
            p[:, L.p_lower(c, k):L.p_lower(c, k) + 3] = lo
            lbg[:, L.g_box(c, k):L.g_box(c, k) + 3] = lo
            ubg[:, L.g_box(c, k):L.g_box(c, k) + 3] = up
            for j in range(NJ):
                lbg[:, L.g_fric(c, j, k):L.g_fric(c, j, k) + 4] = -INF
        # cold start: positions on the nominal ones, weight shared by all corners
        for k in range(N + 1):
            x0[:, L.x_pos(c, k):L.x_pos(c, k) + 3] = nom[:, k]
        for k in range(N):
            for j in range(NJ):
                x0[:, L.x_frc(c, j, k) + 2] = GRAVITY / (NC * NJ)

    # references and state
    v = step_length / ((ds + ss) * dT)
    comref = np.zeros((B, N + 1, 3))
    comref[..., 0] = np.maximum(0.0, step_length * ell / (ds + ss) - step_length / 2)
    comref[..., 2] = com_height
    com0 = comref[:, 0].copy()
    dcom0 = np.zeros((B, 3))
    dcom0[:, 0] = np.where(ell[:, 0] * step_length / (ds + ss) > step_length / 2, v, 0.0)
    h0 = np.zeros((B, 3))
    if state_noise:
        com0 += rng.uniform(-0.03, 0.03, size=(B, 3)) * state_noise
        dcom0 += rng.uniform(-0.2, 0.2, size=(B, 3)) * state_noise
        h0 += rng.uniform(-0.05, 0.05, size=(B, 3)) * state_noise
    g0 = L.p_glob()
    p[:, g0:g0 + 3], p[:, g0 + 3:g0 + 6], p[:, g0 + 6:g0 + 9] = com0, dcom0, h0
    lbg[:, 0:3], lbg[:, 3:6], lbg[:, 6:9] = com0, dcom0, h0
    ubg[:, 0:9] = lbg[:, 0:9]
    for k in range(N + 1):
        p[:, L.p_comref(k):L.p_comref(k) + 3] = comref[:, k]
        x0[:, L.x_com(k):L.x_com(k) + 3] = comref[:, k]
    if push is not None:
        p[:, L.p_extf(0):L.p_extf(0) + 3] = np.asarray(push, dtype=np.float64).reshape(B, 3)
    # knots whose nominal position refers to the contact the foot stands on at knot 0 (incl. the lift-off knot), and the
    # stance flag at knot 0: what a closed loop needs to replace planned by actual footsteps (rollout.py)
    same = np.stack([(step[c] == step[c][:, :1]) & ~swing[c][:, :1] for c in range(NC)])
    out = dict(p=p, lbg=lbg, ubg=ubg, x0=x0, N=N, dT=dT, phase=ph, same_contact=same, stance0=~swing[:, :, 0], comref=comref)
    if ticks:
        # the same instances as TICK RECORDS (include/cmpc_b200.h): state, wrench, references and the contact windows of the walk
        # as contact lists (host.walk_contact_lists): what a host controller hands to cmpc_solve_ticks_host / cmpc_populate
        ts = tick_stride(N)
        tk = np.zeros((B, ts))
        tk[:, 0:3], tk[:, 3:6], tk[:, 6:9] = com0, dcom0, h0
        if push is not None:
            tk[:, 9:12] = np.asarray(push, dtype=np.float64).reshape(B, 3)
        tk[:, 15] = 1.0 if step_adjust else 0.0
        tk[:, 17:17 + 3 * (N + 1)] = comref.reshape(B, -1)
        dTns = float(round(dT * 1e9))
        sidx = np.arange(nsteps)
        on = np.zeros((NC, nsteps))
        off = np.zeros((NC, nsteps))
        on[0], off[0] = sidx * P, sidx * P + P - ss
        on[1], off[1] = (sidx - 1) * P + ds + ss, sidx * P + ds
        on[:, 0] = -np.inf                                            # the first footstep was always there
        for c in range(NC):
            base = 17 + 6 * (N + 1) + 85 * c
            s0 = (on[c][None, :] <= ph[:, None]).sum(axis=1) - 1     # the last footstep that started at or before knot 0
            cnt = np.minimum(4, nsteps - s0)
            tk[:, base] = cnt
            for j in range(4):
                sj = np.minimum(s0 + j, nsteps - 1)
                ok = j < cnt
                rec = np.zeros((B, 14))
                rec[:, 0] = np.where(np.isinf(on[c][sj]), -1e18, (on[c][sj] - ph) * dTns)
                rec[:, 1] = (off[c][sj] - ph) * dTns
                rec[:, 2], rec[:, 3] = fx[c][sj], fy[c]
                yj = yaw[c][bi, sj]
                cy, sy = np.cos(yj), np.sin(yj)
                rec[:, 5], rec[:, 6], rec[:, 8], rec[:, 9], rec[:, 13] = cy, sy, -sy, cy, 1.0
                tk[:, base + 1 + 14 * j:base + 15 + 14 * j] = np.where(ok[:, None], rec, 0.0)
        out["ticks"] = tk
    return out


def tick_stride(N: int) -> int:
    """doubles per tick record (cmpc_tick_stride of include/cmpc_b200.h)"""
    return 6 * N + 194


def scenario_s0(dcom0=(0.0, 0.0, 0.0)):
    """Known-answer scenario S0 of SURVEY.md 8(d): N = 12, start of a walk (phase 0), right foot swings at knots 3..7."""
    w = walk_batch(N=12, dT=0.1, B=1, seed=0, phase=0, step_adjust=True)
    L = Layout(12)
    p, lbg, ubg, x0 = (w[k][0].copy() for k in ("p", "lbg", "ubg", "x0"))
    for k in range(13):
        ref = (0.05 * min(k, 8) / 8, 0.0, 0.7)
        p[L.p_comref(k):L.p_comref(k) + 3] = ref
        x0[L.x_com(k):L.x_com(k) + 3] = ref
    # S0 keeps the left foot on the ground for the whole horizon (W would lift it at knot 11)
    p[L.p_en(0, 11)] = 1.0
    p[L.p_nom(0, 12):L.p_nom(0, 12) + 3] = (0.0, 0.08, 0.0)
    x0[L.x_pos(0, 12):L.x_pos(0, 12) + 3] = (0.0, 0.08, 0.0)
    for arr in (lbg, ubg):
        arr[L.g_box(0, 11):L.g_box(0, 11) + 3] = 0.0
    p[L.p_upper(0, 11):L.p_upper(0, 11) + 3] = 0.0
    p[L.p_lower(0, 11):L.p_lower(0, 11) + 3] = 0.0
    g0 = L.p_glob()
    p[g0:g0 + 9] = (0, 0, 0.7) + tuple(dcom0) + (0, 0, 0)
    lbg[0:9] = p[g0:g0 + 9]
    ubg[0:9] = p[g0:g0 + 9]
    return dict(p=p, lbg=lbg, ubg=ubg, x0=x0, N=12, dT=0.1)
