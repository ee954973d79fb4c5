// cmpc_populate.cuh -- input population of the MPC on the device and the device side of the closed loop.
//
// What BLF's CentroidalMPC::setState / setReferenceTrajectory / setContactPhaseList do on the host before every solve
// (call sites /root/reference/src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:407, :579, :609): turn (state, references,
// contact phase list) into the solver's formal input (p, lbg, ubg, x0).  Here the host ships a compact TICK RECORD per
// instance -- the state, the references and the few contacts of each foot that the horizon can see (2.3 KB instead of the
// 24.7 KB of p / lbg / ubg / x0 at N = 15) -- and cmpc_populate_kernel expands it next to the solver.  The rules are those of
// host/CentroidalMPC.cpp (Impl::fillInputs, the host restatement the tests pin against): the two must agree bit for bit
// (tests/test_gpu_populate.py).
//
// Tick record (doubles, stride tick_stride(N) = 6 N + 194):
//   [0..8]   com, dcom, angular momentum (setState)            [9..14] external wrench: force, torque (column 0 of p only)
//   [15]     flags: 1 = step adjustment enabled (0: every step box has zero width)      [16] reserved (rollout: phase)
//   [17 .. 17 + 3 (N + 1))  CoM reference, then 3 (N + 1) angular-momentum reference (setReferenceTrajectory)
//   per foot c (0 left_foot, 1 right_foot) at tick_contacts(N, c): [0] number of contacts n <= TK_MAXC, then n records of
//       t_on, t_off  activation / deactivation time in NANOSECONDS relative to the instance's current time (exact integers held
//                    in doubles; "never" = +-1e18), position[3], rotation[9] (column major)            (TKC = 14 doubles)
//   in time order.  The window must hold: the last contact activated at or before the current time, every contact that is
//   still active at (current time - dT) or later and starts inside the horizon, and the first contact that starts after it.
#pragma once

#include "cmpc_core.cuh"

namespace cmpc {

constexpr int TK_MAXC = 6, TKC = 14;
constexpr int TK_STATE = 0, TK_WRENCH = 9, TK_FLAGS = 15, TK_AUX = 16, TK_REF = 17;
constexpr int TK_FOOT = 1 + TK_MAXC * TKC;   // 85 doubles per foot
CMPC_HD int tick_contacts(int N, int c) { return TK_REF + 6 * (N + 1) + c * TK_FOOT; }
CMPC_HD int tick_stride(int N) { return TK_REF + 6 * (N + 1) + NC * TK_FOOT + 1; }   // 6 N + 194, even: 16-byte rows
constexpr double TK_INF = 1e20;   // |bound| >= 1e19 is "no bound" (IPOPT nlp_upper_bound_inf)

#if defined(__CUDACC__)
// queries of BLF Contacts::ContactList on the window of one foot (times relative to the current time, nanoseconds)
struct TickFoot {
    const double* rec;   // n records of TKC doubles
    int n;
    __device__ int active(double t) const     // getActiveContact: on <= t < off
    {
        for (int i = 0; i < n; ++i)
            if (t >= rec[i * TKC] && t < rec[i * TKC + 1]) return i;
        return -1;
    }
    __device__ int present(double t) const    // getPresentContact: the last contact with on <= t
    {
        int best = -1;
        for (int i = 0; i < n && rec[i * TKC] <= t; ++i) best = i;
        return best;
    }
    __device__ int next(double t) const       // getNextContact: the first contact with on > t
    {
        for (int i = 0; i < n; ++i)
            if (rec[i * TKC] > t) return i;
        return -1;
    }
    __device__ const double* pos(int i) const { return rec + i * TKC + 2; }
    __device__ const double* rot(int i) const { return rec + i * TKC + 5; }
};

// One CTA per instance (grid-stride).  Threads split the knots of the two feet; nothing is read back from global memory, so the
// writes of an instance are independent and the kernel is bound by its stores (algorithmic bytes: 8 (n_p + 2 m + n) written,
// 8 tick_stride read per instance).  d_x0 may be null (warm start: the initial guess is the shifted previous solution).
__global__ void cmpc_populate_kernel(const __grid_constant__ Config cfg, int batch, const double* __restrict__ ticks,
                                     double* __restrict__ p_all, double* __restrict__ lbg_all, double* __restrict__ ubg_all,
                                     double* __restrict__ x0_all)
{
    extern __shared__ double tk[];   // the tick record of the instance
    const int N = cfg.N, n = dim_x(N), np = dim_p(N), m = dim_g(N), ts = tick_stride(N);
    const double dTns = (double)llrint(cfg.dT * 1e9);
    for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        const double* src = ticks + (size_t)inst * ts;
        for (int i = threadIdx.x; i < ts; i += blockDim.x) tk[i] = src[i];
        __syncthreads();
        double* p = p_all + (size_t)inst * np;
        double* lb = lbg_all + (size_t)inst * m;
        double* ub = ubg_all + (size_t)inst * m;
        double* x0 = x0_all ? x0_all + (size_t)inst * n : nullptr;
        const bool adjust = tk[TK_FLAGS] != 0.0;
        // ---- rows / entries that do not depend on the contacts
        for (int i = threadIdx.x; i < 15 * N; i += blockDim.x) { lb[15 + i] = 0.0; ub[15 + i] = 0.0; }   // dynamics rows
        for (int i = threadIdx.x; i < 6 * (N + 1); i += blockDim.x) p[p_comref(N, 0) + i] = tk[TK_REF + i];   // comRef | hRef
        for (int i = threadIdx.x; i < 6 * N; i += blockDim.x) {   // external wrench: column 0 only
            const int blk = i / (3 * N), r = i - blk * 3 * N;
            p[p_extf(N, 0) + i] = r < 3 ? tk[TK_WRENCH + 3 * blk + r] : 0.0;
        }
        if (threadIdx.x < 9) {
            const double v = tk[TK_STATE + threadIdx.x];
            p[p_glob(N) + threadIdx.x] = v; lb[threadIdx.x] = v; ub[threadIdx.x] = v;
        }
        if (x0) {
            for (int i = threadIdx.x; i < 3 * (N + 1); i += blockDim.x) {
                x0[x_com(N, 0) + i] = tk[TK_REF + i];   // CoM on its reference
                x0[x_dcom(N, 0) + i] = 0.0;
                x0[x_h(N, 0) + i] = 0.0;
            }
        }
        // ---- per foot and knot: item = (c, k), k = 0 .. N
        for (int it = threadIdx.x; it < NC * (N + 1); it += blockDim.x) {
            const int c = it / (N + 1), k = it - c * (N + 1);
            const TickFoot F{tk + tick_contacts(N, c) + 1, (int)tk[tick_contacts(N, c)]};
            const double t = k * dTns;
            const int a = F.active(t), a0 = F.active(0.0), before = F.active(-dTns);
            // nominal position: the active contact; first swing knot: where the foot still is; in the air: the landing position
            double nom[3] = {0.0, 0.0, 0.0};
            int from = a;
            if (from < 0) {
                from = k >= 1 ? F.active(t - dTns) : before;
                if (from < 0) from = F.next(t);
                if (from < 0) from = F.present(t);
            }
            if (from >= 0) { const double* q = F.pos(from); nom[0] = q[0]; nom[1] = q[1]; nom[2] = q[2]; }
            for (int r = 0; r < 3; ++r) {
                p[p_nom(N, c, k) + r] = nom[r];
                if (x0) x0[x_pos(N, c, k) + r] = nom[r];
            }
            if (k == 0) {
                // current position of the foot: the contact it stands on, or between lift-off and landing in mid swing
                double cur[3] = {nom[0], nom[1], nom[2]};
                if (a < 0 && before < 0) {
                    const int pr = F.present(0.0), nx = F.next(0.0);
                    if (pr >= 0 && nx >= 0) {
                        const double off = F.rec[pr * TKC + 1], on = F.rec[nx * TKC];
                        const double span = __ddiv_rn(on - off, 1e9);
                        const double prog = span > 0 ? __ddiv_rn(__ddiv_rn(0.0 - off, 1e9), span) : 1.0;
                        for (int r = 0; r < 3; ++r)
                            cur[r] = __dadd_rn(F.pos(pr)[r], __dmul_rn(F.pos(nx)[r] - F.pos(pr)[r], prog));   // no contraction: = host
                    }
                }
                for (int r = 0; r < 3; ++r) { p[19 * N + 3 + p_cbase(N, c) + r] = cur[r]; lb[9 + 3 * c + r] = cur[r]; ub[9 + 3 * c + r] = cur[r]; }
            }
            if (k < N) {
                const double* R = a >= 0 ? F.rot(a) : nullptr;
                for (int r = 0; r < 9; ++r) p[p_rot(N, c, k) + r] = R ? R[r] : ((r & 3) == 0 ? 1.0 : 0.0);   // identity on swing knots
                p[p_en(N, c, k)] = a >= 0 ? 1.0 : 0.0;
                // step-adjustment box rows of knot k: swing -> free; the contact the foot stands on now -> zero width; a contact
                // of the future -> the configured bounding box
                for (int r = 0; r < 3; ++r) {
                    double lo = 0.0, up = 0.0;
                    if (adjust) {
                        if (a < 0) { lo = -TK_INF; up = TK_INF; }
                        else if (!(a0 >= 0 && a == a0)) { lo = cfg.box_lo[c][r]; up = cfg.box_up[c][r]; }
                    }
                    p[p_cbase(N, c) + 9 * N + 3 * k + r] = up;
                    p[p_cbase(N, c) + 12 * N + 3 * k + r] = lo;
                    lb[g_box(N, c, k) + r] = lo; ub[g_box(N, c, k) + r] = up;
                }
                for (int r = 0; r < 16; ++r) { lb[g_fric(N, c, 0, k) + r] = -TK_INF; ub[g_fric(N, c, 0, k) + r] = 0.0; }   // A R' f <= 0
                if (x0) {
                    for (int r = 0; r < 3; ++r) x0[x_vel(N, c, k) + r] = 0.0;
                    for (int j = 0; j < NJ; ++j) {   // cold start: the weight shared by the corners
                        x0[x_frc(N, c, j, k)] = 0.0; x0[x_frc(N, c, j, k) + 1] = 0.0; x0[x_frc(N, c, j, k) + 2] = -GRAV_Z / (NC * NJ);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ closed loop, device side
// Rollout state of one instance (doubles, stride ROLL_STRIDE): plant state, push schedule, statistics.  The contact window of
// its tick record is the walk schedule of workloads.walk_batch as contact lists (host.walk_contact_lists), regenerated on the
// device from the phase: the synthetic "planner" of the closed loop.
constexpr int RL_PHASE0 = 0;     // global knot index of tick 0
constexpr int RL_PUSH = 1;       // tick of the push (-1: none), length in ticks, force (3)
constexpr int RL_CONV = 6;       // converged ticks
constexpr int RL_ITERS = 7;      // iterations
constexpr int RL_ERR = 8;        // max horizontal CoM tracking error
constexpr int RL_ZMIN = 9;       // min CoM height
constexpr int RL_ZMP = 10;       // max distance by which a foot's local ZMP left the clamp box before clamping
constexpr int RL_TICK = 11;      // MPC ticks done so far (the kernels take it as the tick index when called with tick < 0:
                                 // a captured CUDA graph of one tick replays without changing kernel arguments)
constexpr int ROLL_STRIDE = 12;
// per instance and foot: positions (and yaw) of the footsteps of the walk, the MPC's adjusted landing positions written in
constexpr int RL_MAXSTEPS = 64;  // footsteps per foot of a rollout table
constexpr int RL_STEP = 4;       // x, y, z, yaw

struct WalkParams {   // the synthetic planner of workloads.walk_batch
    int ds, ss;              // knots of double / single support
    double step_length, com_height, push_threshold;
    double zmp_half_length, zmp_half_width;   // clamp of the local ZMP: 0.08 / 0.03 in the reference (WholeBodyQPBlock.cpp:837-838)
};

// contact s of foot c: activation / deactivation in GLOBAL knots (host.walk_contact_lists); s = 0 was always there
__device__ __forceinline__ void walk_contact_times(const WalkParams& W, int c, int s, double& on, double& off)
{
    const int P = 2 * (W.ds + W.ss);
    if (c == 0) { on = s > 0 ? (double)(s * P) : -1e9; off = (double)(s * P + P - W.ss); }
    else { on = s > 0 ? (double)((s - 1) * P + W.ds + W.ss) : -1e9; off = (double)(s * P + W.ds); }
}

// Writes the tick record of global knot ell = phase0 + tick for every rollout: state = plant state, external force = the push
// when it is above the threshold (the reference ignores smaller wrenches, WholeBodyQPBlock.cpp:1018), references of the
// synthetic planner, and the contact window from the footstep table (planner steps with the MPC's own landings written in:
// the role of updateContactPhaseList, CentroidalMPCBlock.cpp:32-110).  One thread per rollout.
__global__ void cmpc_rollout_tick_kernel(const __grid_constant__ Config cfg, WalkParams W, int batch, int tick,
                                         const double* __restrict__ roll, const double* __restrict__ state,
                                         const double* __restrict__ steps, double* __restrict__ ticks,
                                         double* __restrict__ ext6, int step_adjust)
{
    const int N = cfg.N, ts = tick_stride(N), P = 2 * (W.ds + W.ss);
    const double dTns = (double)llrint(cfg.dT * 1e9);
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += gridDim.x * blockDim.x) {
        const double* r = roll + (size_t)b * ROLL_STRIDE;
        double* tk = ticks + (size_t)b * ts;
        if (tick < 0) tick = (int)r[RL_TICK];
        const int ell = (int)r[RL_PHASE0] + tick;
        for (int i = 0; i < 9; ++i) tk[TK_STATE + i] = state[(size_t)b * 9 + i];
        const int pt = (int)r[RL_PUSH], pl = (int)r[RL_PUSH + 1];
        const bool active = pt >= 0 && tick >= pt && tick < pt + pl;
        const double f0 = r[RL_PUSH + 2], f1 = r[RL_PUSH + 3], f2 = r[RL_PUSH + 4];
        const bool big = sqrt(f0 * f0 + f1 * f1 + f2 * f2) >= W.push_threshold;
        tk[TK_WRENCH] = active && big ? f0 : 0.0; tk[TK_WRENCH + 1] = active && big ? f1 : 0.0; tk[TK_WRENCH + 2] = active && big ? f2 : 0.0;
        tk[TK_WRENCH + 3] = tk[TK_WRENCH + 4] = tk[TK_WRENCH + 5] = 0.0;
        double* e = ext6 + (size_t)b * 6;   // what acts on the plant (also below the threshold)
        e[0] = active ? f0 : 0.0; e[1] = active ? f1 : 0.0; e[2] = active ? f2 : 0.0; e[3] = e[4] = e[5] = 0.0;
        tk[TK_FLAGS] = step_adjust ? 1.0 : 0.0;
        tk[TK_AUX] = (double)ell;
        for (int k = 0; k <= N; ++k) {   // CoM reference of the planner: constant speed after half a step, fixed height
            const double cx = W.step_length * (double)(ell + k) / (double)(W.ds + W.ss) - W.step_length / 2;
            tk[TK_REF + 3 * k] = cx > 0.0 ? cx : 0.0; tk[TK_REF + 3 * k + 1] = 0.0; tk[TK_REF + 3 * k + 2] = W.com_height;
            tk[TK_REF + 3 * (N + 1) + 3 * k] = tk[TK_REF + 3 * (N + 1) + 3 * k + 1] = tk[TK_REF + 3 * (N + 1) + 3 * k + 2] = 0.0;
        }
        for (int c = 0; c < NC; ++c) {
            double* out = tk + tick_contacts(N, c);
            // the last footstep that started at or before ell
            int s0 = 0;
            for (int s = 1; s < RL_MAXSTEPS; ++s) {
                double on, off;
                walk_contact_times(W, c, s, on, off);
                if (on <= (double)ell) s0 = s; else break;
            }
            int cnt = 0;
            for (int s = s0; s < RL_MAXSTEPS && cnt < 4; ++s, ++cnt) {
                double on, off;
                walk_contact_times(W, c, s, on, off);
                const double* st = steps + (((size_t)b * NC + c) * RL_MAXSTEPS + s) * RL_STEP;
                double* q = out + 1 + cnt * TKC;
                q[0] = on <= -1e8 ? -1e18 : (on - ell) * dTns;
                q[1] = (off - ell) * dTns;
                q[2] = st[0]; q[3] = st[1]; q[4] = st[2];
                const double cy = cos(st[3]), sy = sin(st[3]);
                q[5] = cy; q[6] = sy; q[7] = 0.0; q[8] = -sy; q[9] = cy; q[10] = 0.0; q[11] = 0.0; q[12] = 0.0; q[13] = 1.0;
            }
            out[0] = (double)cnt;
            (void)P;
        }
    }
}

// Desired ZMP from the knot-0 corner forces (computeDesiredZMP, WholeBodyQPBlock.cpp:805-873): per contact the local ZMP
// (-tau_y / f_z, tau_x / f_z) of the corner forces about the sole origin, clamped to the foot (half length / half width),
// mapped to the inertial frame and averaged with the normal forces as weights.  Returns false when no contact carries force.
// excess: how far the worst local ZMP was outside the clamp box before clamping (<= 0: inside).
__device__ __forceinline__ bool desired_zmp(const Config& cfg, const double* x, const double* p, double hl, double hw, double* zmp,
                                            double* excess)
{
    const int N = cfg.N;
    double zx = 0.0, zy = 0.0, fz_tot = 0.0, ex = -1e300;
    for (int c = 0; c < NC; ++c) {
        const double en = p[p_en(N, c, 0)];
        const double* R = p + p_rot(N, c, 0);
        double Fz = 0.0, T0 = 0.0, T1 = 0.0;
        for (int j = 0; j < NJ; ++j) {
            const double* cr = cfg.corner[c][j];
            const double* f = x + x_frc(N, c, j, 0);
            double fl[3];   // R' f: the corner force in the contact frame
            for (int a = 0; a < 3; ++a) fl[a] = en * (R[3 * a] * f[0] + R[3 * a + 1] * f[1] + R[3 * a + 2] * f[2]);
            Fz += en * f[2];
            T0 += cr[1] * fl[2] - cr[2] * fl[1];
            T1 += cr[2] * fl[0] - cr[0] * fl[2];
        }
        if (Fz <= 0.001) continue;
        double lx = -T1 / Fz, ly = T0 / Fz;
        ex = fmax(ex, fmax(fabs(lx) - hl, fabs(ly) - hw));
        lx = fmin(hl, fmax(-hl, lx)); ly = fmin(hw, fmax(-hw, ly));
        const double* pos = x + x_pos(N, c, 0);
        zx += Fz * (R[0] * lx + R[3] * ly + pos[0]);
        zy += Fz * (R[1] * lx + R[4] * ly + pos[1]);
        fz_tot += Fz;
    }
    if (excess) *excess = ex;
    if (fz_tot < 0.001) { zmp[0] = zmp[1] = 0.0; return false; }
    zmp[0] = zx / fz_tot; zmp[1] = zy / fz_tot;
    return true;
}

__global__ void cmpc_zmp_kernel(const __grid_constant__ Config cfg, int batch, const double* __restrict__ x_all,
                                const double* __restrict__ p_all, double hl, double hw, double* __restrict__ zmp, int* __restrict__ valid)
{
    const int N = cfg.N, n = dim_x(N), np = dim_p(N);
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += gridDim.x * blockDim.x) {
        const bool ok = desired_zmp(cfg, x_all + (size_t)b * n, p_all + (size_t)b * np, hl, hw, zmp + 2 * (size_t)b, nullptr);
        if (valid) valid[b] = ok ? 1 : 0;
    }
}

// After the solve and the plant step of a tick: statistics, and the footstep table takes the landing position the MPC chose
// for a foot that touches down at the next tick (from then on the contact is where the foot landed, with the planner's
// timing: updateContactPhaseList).
__global__ void cmpc_rollout_feedback_kernel(const __grid_constant__ Config cfg, WalkParams W, int batch, int tick,
                                             const double* __restrict__ x_all, const double* __restrict__ p_all,
                                             const double* __restrict__ plant, const int* __restrict__ status,
                                             const int* __restrict__ iters, double* __restrict__ roll, double* __restrict__ steps)
{
    const int N = cfg.N, n = dim_x(N), np = dim_p(N);
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += gridDim.x * blockDim.x) {
        double* r = roll + (size_t)b * ROLL_STRIDE;
        const double* x = x_all + (size_t)b * n;
        const double* p = p_all + (size_t)b * np;
        const double* s9 = plant + (size_t)b * 9;
        if (tick < 0) tick = (int)r[RL_TICK];
        r[RL_TICK] = (double)(tick + 1);
        const int ell = (int)r[RL_PHASE0] + tick;
        r[RL_CONV] += (status[b] == 0 || status[b] == 5) ? 1.0 : 0.0;
        r[RL_ITERS] += (double)iters[b];
        // tracking error against the reference of the knot the plant has just reached
        const double ex = s9[0] - p[p_comref(N, 1)], ey = s9[1] - p[p_comref(N, 1) + 1];
        r[RL_ERR] = fmax(r[RL_ERR], sqrt(ex * ex + ey * ey));
        r[RL_ZMIN] = fmin(r[RL_ZMIN], s9[2]);
        double zmp[2], excess;
        if (desired_zmp(cfg, x, p, W.zmp_half_length, W.zmp_half_width, zmp, &excess)) r[RL_ZMP] = fmax(r[RL_ZMP], excess);
        // a foot that is in the air at knot 0 and down at knot 1 lands now: its footstep is where the MPC put it
        for (int c = 0; c < NC; ++c) {
            if (p[p_en(N, c, 0)] < 0.5 && p[p_en(N, c, 1)] > 0.5) {
                int s1 = 0;   // the footstep that is active at global knot ell + 1
                for (int s = 1; s < RL_MAXSTEPS; ++s) {
                    double on, off;
                    walk_contact_times(W, c, s, on, off);
                    if (on <= (double)(ell + 1)) s1 = s; else break;
                }
                double* st = steps + (((size_t)b * NC + c) * RL_MAXSTEPS + s1) * RL_STEP;
                const double* land = x + x_pos(N, c, N >= 2 ? 2 : 1);   // pos_{landing + 1}: the row the step box constrains
                st[0] = land[0]; st[1] = land[1]; st[2] = land[2];
            }
        }
    }
}

// Reference resampling (the Math::LinearSpline frequency adapters of CentroidalMPCBlock.cpp:201-260, 525-577): n_in planner
// samples of the CoM and of the angular momentum at times t_in (increasing, shared by the batch) -> the N + 1 knots of the MPC
// at times t_out (ordered), written straight into the tick records.  The reference divides the angular momentum by the robot
// mass (:525-529) and overrides the CoM height with a constant (:531-535, 0.7 m); com_height < 0 keeps the planner's height.
// Outside the input range the end points are held.  One thread per (instance, knot).
__global__ void cmpc_resample_kernel(int N, int batch, int n_in, const double* __restrict__ t_in, const double* __restrict__ com_in,
                                     const double* __restrict__ h_in, const double* __restrict__ t_out, double inv_mass,
                                     double com_height, double* __restrict__ ticks)
{
    const int ts = tick_stride(N);
    const long long total = (long long)batch * (N + 1);
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(it / (N + 1)), k = (int)(it - (long long)b * (N + 1));
        const double t = t_out[k];
        const double* ci = com_in + (size_t)b * n_in * 3;
        const double* hi = h_in + (size_t)b * n_in * 3;
        double* tk = ticks + (size_t)b * ts + TK_REF;
        int seg = 0;
        double a = 0.0;
        if (t <= t_in[0]) { seg = 0; a = 0.0; }
        else if (t >= t_in[n_in - 1]) { seg = n_in - 2 < 0 ? 0 : n_in - 2; a = n_in >= 2 ? 1.0 : 0.0; }
        else {
            int lo = 0, hi2 = n_in - 1;   // t_in[lo] < t <= t_in[hi2]
            while (hi2 - lo > 1) { const int mid = (lo + hi2) >> 1; if (t_in[mid] < t) lo = mid; else hi2 = mid; }
            seg = lo;
            a = (t - t_in[seg]) / (t_in[seg + 1] - t_in[seg]);
        }
        const int s1 = n_in >= 2 ? seg + 1 : seg;
        for (int r = 0; r < 3; ++r) {
            double c = ci[3 * seg + r] * (1.0 - a) + ci[3 * s1 + r] * a;
            if (r == 2 && com_height >= 0.0) c = com_height;
            tk[3 * k + r] = c;
            tk[3 * (N + 1) + 3 * k + r] = (hi[3 * seg + r] * (1.0 - a) + hi[3 * s1 + r] * a) * inv_mass;
        }
    }
}
#endif  // __CUDACC__

}  // namespace cmpc
