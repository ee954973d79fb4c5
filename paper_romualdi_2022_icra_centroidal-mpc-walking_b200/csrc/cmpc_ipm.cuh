// cmpc_ipm.cuh -- the interior-point solve of one centroidal-MPC instance by one team of NT threads, STAGE MAJOR.
//
// The algorithm (IPOPT's filter line-search interior point, Waechter & Biegler 2006, the
// restatement of what BLF CentroidalMPC::advance() delegates to CasADi + IPOPT at
// /root/reference/src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:615), re-organised for the GPU:
//   * every per-instance vector lives knot by knot (z_k = [s_k | u_k], the 15 equality rows that define s_k, the 38 path rows
//     of knot k), padded so that an item (knot, role) is a shift and a mask away from the thread index: no divisions, no
//     CasADi-order index arithmetic inside the iterations (the CasADi order exists only at entry and exit);
//   * per-knot constants (rotations, friction rows A R', corner offsets R r_j, references, ...) are tabulated once per solve;
//   * the element-wise work of an iteration is fused into a few passes over (knot, role) items: one KKT pass (gradient, dual /
//     primal / complementarity residuals; the barrier-parameter loop needs no second pass because max |s z - mu| follows from
//     max and min of the products), one barrier pass (Sigma, barrier gradients, the small blocks of the Riccati sweep), one
//     step pass (eliminated rows, fraction to the boundary, directional derivative), one evaluation pass per trial point;
//   * the adjoint recursion for the dynamics multipliers is a 16-step shuffle chain on one warp.
// The linear algebra (Riccati sweeps on 3 x 3 tiles) is in cmpc_warp.cuh.
#pragma once

#include "cmpc_warp.cuh"
#if defined(CMPC_TRACE) && !defined(__CUDA_ARCH__)
#include <cstdio>
#endif

namespace cmpc {

#if defined(__CUDA_ARCH__)
#define CMPC_UNROLL2 _Pragma("unroll 2")
#define CMPC_UNROLL4 _Pragma("unroll 4")
#else
#define CMPC_UNROLL2
#define CMPC_UNROLL4
#endif

// Item loops of the element-wise passes, ROLE MAJOR: the lanes of a warp take consecutive roles (a variable, a row, ...) and
// the warps of the team split the knots, so that the decoding of a role (divisions, corner / axis indices, table offsets) is
// invariant in the inner loop over the knots and hoisted out of it, the knot index is uniform over the warp, and the loads of
// consecutive knots are independent of each other (the passes are bound by their instruction count: profiles/r1_notes.md)
#define CMPC_ROLES(nroles) for (int r = lane & 31; r < (nroles); r += 32)
#ifndef CMPC_HOIST_LOADS
#define CMPC_HOIST_LOADS 1
#endif
#ifndef CMPC_KNOTS_UNROLL
#define CMPC_KNOTS_UNROLL 1
#endif
#if defined(__CUDA_ARCH__) && CMPC_KNOTS_UNROLL > 1
#define CMPC_KNOTS_PRAGMA _Pragma("unroll 2")
#else
#define CMPC_KNOTS_PRAGMA
#endif
#define CMPC_KNOTS(kfirst, klast) CMPC_KNOTS_PRAGMA for (int k = (kfirst) + (lane >> 5); k <= (klast); k += (NT >= 32 ? NT / 32 : 1))

// the per-variable entry functions: inlined into the role-major loops (role decoding hoisted) or shared out-of-line copies
#ifndef CMPC_INLINE_ENTRIES
#define CMPC_INLINE_ENTRIES 1
#endif
#if CMPC_INLINE_ENTRIES
#define CMPC_ENTRY CMPC_HD
#else
#define CMPC_ENTRY CMPC_FN
#endif

constexpr int PC_MAX_ITER = 50;  // predictor-corrector iterations after which an instance is handed to the monotone path
constexpr int PS = 40;  // stride of a knot in the path-row arrays: friction 0..31 (16 c + 4 j + r) | step box 32..37 (32 + 3 c + q) | pad
// per-knot table of constants
constexpr int TS = 100;
constexpr int T_EN = 0, T_OMD = 2, T_VM = 4, T_AR = 8, T_R = 32, T_RR = 50, T_NOM = 74, T_NOM1 = 80, T_CREF = 86, T_HREF = 89,
              T_EXTF = 92, T_EXTT = 95, T_OM2 = 98;
constexpr int AWS = 32;  // stride of the per-knot values of the off-diagonal non-zeros of A (15 x 2)

struct WorkS {
    double *z, *dz, *zt, *gr, *res;                                                   // (N + 1) * ZS
    double *lam, *dlam, *ceq, *lamn, *vco, *beq;                                      // (N + 1) * ES
    double *gp, *sl, *slt, *zl, *zu, *lo, *up, *yp, *ypn, *sig, *tt, *dsl, *dzl, *dzu, *dyp, *ccl, *ccu;  // N * PS
    double *tab;                                                                      // (N + 1) * TS
    double *sd, *aw, *dfc;                                                            // N * SD_STRIDE, N * AWS, N * 8
    double *small;                                                                    // (N + 1) * SMALL_STRIDE
    double *ric;                                                                      // N * WRIC_STRIDE
};
CMPC_HD int works_vector_doubles(int N)   // everything but the factor blocks
{
    return 5 * (N + 1) * ZS + 6 * (N + 1) * ES + 17 * N * PS + (N + 1) * TS + N * (SD_STRIDE + AWS + 8) + (N + 1) * SMALL_STRIDE;
}
CMPC_HD int works_doubles(int N) { return works_vector_doubles(N) + N * WRIC_STRIDE; }
CMPC_HD void works_carve(double* base, int N, WorkS& w)
{
    double* c = base;
    const int nz = (N + 1) * ZS, ne = (N + 1) * ES, np = N * PS;
    w.z = c; c += nz; w.dz = c; c += nz; w.zt = c; c += nz; w.gr = c; c += nz; w.res = c; c += nz;
    w.lam = c; c += ne; w.dlam = c; c += ne; w.ceq = c; c += ne; w.lamn = c; c += ne; w.vco = c; c += ne; w.beq = c; c += ne;
    w.gp = c; c += np; w.sl = c; c += np; w.slt = c; c += np; w.zl = c; c += np; w.zu = c; c += np; w.lo = c; c += np;
    w.up = c; c += np; w.yp = c; c += np; w.ypn = c; c += np; w.sig = c; c += np; w.tt = c; c += np; w.dsl = c; c += np;
    w.dzl = c; c += np; w.dzu = c; c += np; w.dyp = c; c += np; w.ccl = c; c += np; w.ccu = c; c += np;
    w.tab = c; c += (N + 1) * TS;
    w.sd = c; c += N * SD_STRIDE; w.aw = c; c += N * AWS; w.dfc = c; c += N * 8;
    w.small = c; c += (N + 1) * SMALL_STRIDE;
    w.ric = c;
}

struct KktStats {
    double dual, viol, pmax, pmin;    // max |grad_x L|, max violation, max / min of the products slack * multiplier
    double sum_y, sum_z, nb, nrows;   // for the scaling factors s_d, s_c
};
struct StepStats { double rho, amax, az, dphi, bad; };
struct EvalStats { double f, theta, bar; };  // objective, l1 infeasibility, barrier sum (phi_mu = f + mu * bar)
// affine-scaling (predictor) step: 1 / step length to the boundary (primal, dual), the coefficients of the mean
// complementarity after the step  (s00 + ap s10 + ad s01 + ap ad s11) / nb


// shared memory of one team: the Riccati block first (P at offset 0: 16-byte aligned async copies), then everything the
// interior-point loop would otherwise keep on the thread stacks.  Local memory is poison here: the shared-memory carve-out
// leaves almost no L1, so every stack access is an L2 round trip (ncu: as many local loads as global loads, 83 % missing L1).
struct alignas(16) ISmem {
    WSmem sw;
    WorkS w;                 // array pointers of the instance's scratch block
    SweepIO io;
    KktStats ks;
    StepStats ss;
    EvalStats es;
    double filt_t[MAX_FILTER], filt_p[MAX_FILTER];
    int inst;                // instance taken from the work queue
};

// CasADi row of path row l of knot k
CMPC_HD int path_row(int N, int k, int l)
{
    if (l < 32) return g_fric(N, l >> 4, (l >> 2) & 3, k) + (l & 3);
    const int q = l - 32;
    return g_box(N, q / 3, k) + q % 3;
}
// value  a' v  of path row l of knot k for variables taken from zk (knot k) / zk1 (knot k + 1); sub = 1 subtracts the nominal
// position (row value), sub = 0 gives the jacobian row applied to a step
CMPC_HD double path_dot(const double* tab, const double* zk, const double* zk1, int l, int sub)
{
    if (l < 32) {
        const int c = l >> 4, j = (l >> 2) & 3, r = l & 3;
        const double* ar = tab + T_AR + 12 * c + 3 * r;
        const double* f = zk + NS + 6 + 12 * c + 3 * j;
        return ar[0] * f[0] + ar[1] * f[1] + ar[2] * f[2];
    }
    const int q = l - 32, c = q / 3, qq = q - 3 * c;
    const double* rc = tab + T_R + 9 * c + 3 * qq;  // column qq of R_c
    const double* pos = zk1 + 9 + 3 * c;
    const double* nom = tab + T_NOM1 + 3 * c;
    double v = 0.0;
    for (int a = 0; a < 3; ++a) v += rc[a] * (pos[a] - (sub ? nom[a] : 0.0));
    return v;
}

// gradient of the objective in variable v (0..44) of knot k at the point zsrc
CMPC_ENTRY double grad_entry(const Config& cfg, const WorkS& w, const double* zsrc, int k, int v)
{
    const int N = cfg.N;
    const double* t = w.tab + k * TS;
    const double* zk = zsrc + k * ZS;
    if (v < NS) {
        if (v < 2) return 2.0 * cfg.w_com[v] * (zk[v] - t[T_CREF + v]);
        if (v == 2) return 2.0 * t[T_OM2] * (zk[2] - t[T_CREF + 2]);
        if (v < 6) return 0.0;
        if (v < 9) return 2.0 * cfg.w_h * (zk[v] - t[T_HREF + v - 6]);
        return 2.0 * cfg.w_pos * (zk[v] - t[T_NOM + v - 9]);
    }
    const int u = v - NS;
    if (u < 6) return 0.0;
    const int f = u - 6, c = f / 12, a = f % 3;
    const double en = t[T_EN + c];
    const double* fc = zk + NS + 6 + 12 * c + a;  // corner j at fc[3 j]
    const double own = zk[v], sum = fc[0] + fc[3] + fc[6] + fc[9], mean = en / NJ * sum, d = own - mean;
    double g = 2.0 * cfg.w_sym * (d - (en / NJ) * (sum - NJ * mean));
    if (k + 1 < N) g -= 2.0 * cfg.w_rate[a] * (zsrc[(k + 1) * ZS + v] - own);
    if (k > 0) g += 2.0 * cfg.w_rate[a] * (own - zsrc[(k - 1) * ZS + v]);
    return g;
}

// (J' y)[variable v of knot k] for multipliers lam (equality rows, stage major) and yp (path rows); needs sd / aw at the iterate
CMPC_ENTRY double jty_entry(const Config& cfg, const WorkS& w, const double* lam, const double* yp, int k, int v)
{
    const int N = cfg.N;
    if (v < NS) {
        const int i = v;
        double r = lam[k * ES + i];
        if (k < N) {
            const double* y1 = lam + (k + 1) * ES;
            int rr[2];
            acol_rows(i, rr);
            r -= y1[i] + w.aw[k * AWS + 2 * i] * y1[rr[0]] + w.aw[k * AWS + 2 * i + 1] * y1[rr[1]];
        }
        if (k > 0 && i >= 9) {
            const int c = (i - 9) / 3, a = (i - 9) % 3;
            const double* R = w.tab + (k - 1) * TS + T_R + 9 * c;
            const double* yb = yp + (k - 1) * PS + 32 + 3 * c;
            r += R[a] * yb[0] + R[3 + a] * yb[1] + R[6 + a] * yb[2];
        }
        return r;
    }
    const int u = v - NS;
    const double* t = w.tab + k * TS;
    const double* y1 = lam + (k + 1) * ES;
    if (u < 6) return -t[T_OMD + u / 3] * y1[9 + u];
    const int f = u - 6, c = f / 12, j = (f % 12) / 3, a = f % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    const double* rho = w.sd + k * SD_STRIDE + SD_RHO + 3 * (4 * c + j);
    double r = -cfg.dT * t[T_EN + c] * (y1[3 + a] + y1[6 + a1] * rho[a2] - y1[6 + a2] * rho[a1]);
    const double* ar = t + T_AR + 12 * c + a;  // row q at ar[3 q]
    const double* yf = yp + k * PS + 16 * c + 4 * j;
    r += ar[0] * yf[0] + ar[3] * yf[1] + ar[6] * yf[2] + ar[9] * yf[3];
    return r;
}

// (W dz)[variable v of knot k]: hessian of the lagrangian (nlp_hess_l, tmp.c:58926) + delta_w I, times the step
CMPC_ENTRY double hess_dz_entry(const Config& cfg, const WorkS& w, double dw, int k, int v)
{
    const int N = cfg.N;
    const double* t = w.tab + k * TS;
    const double* dzk = w.dz + k * ZS;
    if (v < NS) {
        const int i = v;
        double q = i < 2 ? 2.0 * cfg.w_com[i] : i == 2 ? 2.0 * t[T_OM2] : i < 6 ? 0.0 : i < 9 ? 2.0 * cfg.w_h : 2.0 * cfg.w_pos;
        double r = (q + dw) * dzk[i];
        if (k < N && (i < 3 || i >= 9)) {  // bilinear block between forces and com / pos: dT en [lam_h]x
            const double* lamh = w.lam + (k + 1) * ES + 6;
            const int a = i < 3 ? i : (i - 9) % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
            const double* dF = w.dfc + k * 8;
            double acc;
            if (i < 3) acc = t[T_EN] * (dF[a1] * lamh[a2] - dF[a2] * lamh[a1]) + t[T_EN + 1] * (dF[3 + a1] * lamh[a2] - dF[3 + a2] * lamh[a1]);
            else { const int c = (i - 9) / 3; acc = -t[T_EN + c] * (dF[3 * c + a1] * lamh[a2] - dF[3 * c + a2] * lamh[a1]); }
            r += cfg.dT * acc;
        }
        return r;
    }
    const int u = v - NS;
    if (u < 6) return t[T_VM + u / 3] != 0.0 ? dzk[v] : dw * dzk[v];
    const int f = u - 6, c = f / 12, a = f % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    const double a4 = t[T_EN + c] / NJ, own = dzk[v];
    const double* dc = dzk + NS + 6 + 12 * c + a;
    const double sum = dc[0] + dc[3] + dc[6] + dc[9];
    double r = 2.0 * cfg.w_sym * (own + (NJ * a4 * a4 - 2.0 * a4) * sum) + dw * own;
    if (k > 0) r += 2.0 * cfg.w_rate[a] * (own - w.dz[(k - 1) * ZS + v]);
    if (k + 1 < N) r += 2.0 * cfg.w_rate[a] * (own - w.dz[(k + 1) * ZS + v]);
    const double* lamh = w.lam + (k + 1) * ES + 6;
    const double e1 = dzk[a1] - dzk[9 + 3 * c + a1], e2 = dzk[a2] - dzk[9 + 3 * c + a2];
    r += cfg.dT * t[T_EN + c] * (lamh[a1] * e2 - lamh[a2] * e1);
    return r;
}

// ------------------------------------------------------------------------------------------------ evaluation of a point
// stage data, equality residuals, path row values, objective, theta (l1 infeasibility) and the barrier function at (zsrc, slsrc)
template <int NT, int G, class Cta>
CMPC_FN void eval_point(Team T, Cta& cta, const Config& cfg, ISmem& sm, const double* zsrc, const double* slsrc)
{
    cta_align<G>(T);
    const WorkS& w = sm.w;
    const int N = cfg.N;
    const double dT = cfg.dT;
    CMPC_LANES
        CMPC_ROLES(32) CMPC_KNOTS(0, N - 1) {
            const double* t = w.tab + k * TS;
            const double* zk = zsrc + k * ZS;
            double* d = w.sd + k * SD_STRIDE;
            if (r < 24) {
                const int c = r >= 12 ? 1 : 0, a = (r - 12 * c) % 3;
                d[SD_RHO + r] = t[T_RR + r] + zk[9 + 3 * c + a] - zk[a];
            } else if (r < 30) {
                const int q = r - 24, c = q / 3, a = q - 3 * c;
                const double* fc = zk + NS + 6 + 12 * c + a;
                d[SD_FC + q] = fc[0] + fc[3] + fc[6] + fc[9];
            } else {
                d[SD_EN + r - 30] = t[T_EN + r - 30];
                d[SD_VM + r - 30] = t[T_VM + r - 30];
            }
        }
    CMPC_LANES_END
    double acc[3] = {0.0, 0.0, 0.0};  // f, theta, barrier terms (without the factor mu)
    CMPC_LANES
        CMPC_ROLES(63) CMPC_KNOTS(0, N) {
            const double* t = w.tab + k * TS;
            const double* zk = zsrc + k * ZS;
            if (r < NS) {
                // equality row block k: the rows that define s_k (k = 0: initial condition, k >= 1: dynamics of knot k - 1)
                double c;
                if (k == 0) c = zk[r];
                else {
                    const double* zp = zsrc + (k - 1) * ZS;
                    const double* tp = w.tab + (k - 1) * TS;
                    const double* d = w.sd + (k - 1) * SD_STRIDE;
                    c = zk[r] - zp[r];
                    if (r < 3) c -= dT * zp[3 + r];
                    else if (r < 6) {
                        const int a = r - 3;
                        c -= dT * ((a == 2 ? GRAV_Z : 0.0) + tp[T_EXTF + a] + tp[T_EN] * d[SD_FC + a] + tp[T_EN + 1] * d[SD_FC + 3 + a]);
                    } else if (r < 9) {
                        const int a = r - 6, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                        double tq = tp[T_EXTT + a];
                        for (int cc = 0; cc < NC; ++cc) {
                            double s = 0.0;
                            for (int j = 0; j < NJ; ++j) {
                                const double* rho = d + SD_RHO + 3 * (4 * cc + j);
                                const double* fo = zp + NS + 6 + 12 * cc + 3 * j;
                                s += rho[a1] * fo[a2] - rho[a2] * fo[a1];
                            }
                            tq += tp[T_EN + cc] * s;
                        }
                        c -= dT * tq;
                    } else {
                        const int cc = (r - 9) / 3;
                        c -= tp[T_OMD + cc] * zp[NS + r - 9];
                    }
                }
                c -= w.beq[k * ES + r];
                w.ceq[k * ES + r] = c;
                acc[1] += fabs(c);
                // stage cost of s_k[r]
                if (r < 2) { const double e = zk[r] - t[T_CREF + r]; acc[0] += cfg.w_com[r] * e * e; }
                else if (r == 2) { const double e = zk[2] - t[T_CREF + 2]; acc[0] += t[T_OM2] * e * e; }
                else if (r >= 6 && r < 9) { const double e = zk[r] - t[T_HREF + r - 6]; acc[0] += cfg.w_h * e * e; }
                else if (r >= 9) { const double e = zk[r] - t[T_NOM + r - 9]; acc[0] += cfg.w_pos * e * e; }
            } else if (k < N && r >= 16 && r < 16 + 38) {
                const int l = r - 16, pr = k * PS + l;
                const double sl = w.lo[pr], su = w.up[pr], s = slsrc[pr];   // requested together with the operands of the row value
                const double g = path_dot(t, zk, zsrc + (k + 1) * ZS, l, 1);
                w.gp[pr] = g;
                const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
                if (hl || hu) {
                    if (sl == su) acc[1] += fabs(g - sl);
                    else {
                        acc[1] += fabs(g - s);
                        if (hl) acc[2] -= log(s - sl);
                        if (hu) acc[2] -= log(su - s);
                        if (hl && !hu) acc[2] += KAPPA_D * (s - sl);
                        if (hu && !hl) acc[2] += KAPPA_D * (su - s);
                    }
                }
            } else if (k < N && r >= 54 && r < 60) {
                // force cost of (contact c, axis a): symmetry + rate of change
                const int q = r - 54, c = q / 3, a = q - 3 * c;
                const double en = t[T_EN + c];
                const double* fc = zk + NS + 6 + 12 * c + a;
                const double mean = en / NJ * (fc[0] + fc[3] + fc[6] + fc[9]);
                for (int j = 0; j < NJ; ++j) {
                    const double dd = fc[3 * j] - mean;
                    acc[0] += cfg.w_sym * dd * dd;
                    if (k + 1 < N) { const double dn = fc[3 * j + ZS] - fc[3 * j]; acc[0] += cfg.w_rate[a] * dn * dn; }
                }
            } else if (k < N && r >= 60 && r < 63) {
                const int a = r - 60;
                const double* d = w.sd + k * SD_STRIDE;
                w.sd[k * SD_STRIDE + SD_FALL + a] = t[T_EN] * d[SD_FC + a] + t[T_EN + 1] * d[SD_FC + 3 + a];
            }
        }
    CMPC_LANES_END_NOSYNC
    cta.template reduce3<0, 0, 3>(acc, acc, acc);
    CMPC_LANES
        if (lane == 0) { sm.es.f = acc[0]; sm.es.theta = acc[1]; sm.es.bar = acc[2]; }
    CMPC_LANES_END
}

// scaled optimality error E_mu (Waechter-Biegler eq. 5, 6) from the statistics of one KKT pass
CMPC_HD double kkt_E(const KktStats& s, double mu, double* compl_out)
{
    const double cmp = s.nb > 0.0 ? fmax(fabs(s.pmax - mu), fabs(s.pmin - mu)) : 0.0;
    const double sd = fmax(S_MAX, (s.sum_y + s.sum_z) / fmax(1.0, s.nrows + s.nb)) / S_MAX;
    const double sc = fmax(S_MAX, s.sum_z / fmax(1.0, s.nb)) / S_MAX;
    if (compl_out) *compl_out = cmp;
    return fmax(s.dual / sd, fmax(s.viol, cmp / sc));
}

// one pass: gradient of f (stored), dual residual, violation, complementarity products, multiplier sums
template <int NT, int G, class Cta>
CMPC_FN void kkt_pass(Team T, Cta& cta, const Config& cfg, ISmem& sm)
{
    cta_align<G>(T);
    const WorkS& w = sm.w;
    const int N = cfg.N;
    // the values of the off-diagonal non-zeros of A at the iterate (needed by jty_entry and by the adjoint recursion)
    CMPC_LANES
        for (int it = lane; it < N * NS; it += NT) {   // 15 roles would leave half of every warp idle on each of its knots: flat
            const int k = it / NS, i = it - NS * k;
            double v2[2];
            acol_vals(i, w.sd + k * SD_STRIDE, cfg.dT, v2);
            w.aw[k * AWS + 2 * i] = v2[0]; w.aw[k * AWS + 2 * i + 1] = v2[1];
        }
    CMPC_LANES_END
    double vmax[3] = {0.0, 0.0, -HUGE_VAL};   // dual, viol, pmax
    double vmin[1] = {HUGE_VAL};              // pmin
    double vsum[4] = {0.0, 0.0, 0.0, 0.0};    // |y|, z, bounds, rows
    CMPC_LANES
        CMPC_ROLES(48 + 38) CMPC_KNOTS(0, N) {
            if (r < NS + NU) {
                if (k == N && r >= NS) continue;
                const double g = grad_entry(cfg, w, w.z, k, r);
                const double jt = jty_entry(cfg, w, w.lam, w.yp, k, r);   // its loads before the store below (a store pins the loads behind it)
                w.gr[k * ZS + r] = g;
                vmax[0] = fmax(vmax[0], fabs(g + jt));
                if (r < NS) {  // the equality row that defines s_k[r]
                    const int e = k * ES + r;
                    vmax[1] = fmax(vmax[1], fabs(w.ceq[e]));
                    vsum[0] += fabs(w.lam[e]);
                    vsum[3] += 1.0;
                }
            } else if (k < N && r >= 48) {
                const int pr = k * PS + r - 48;
                // (all global operands of the row requested before the first branch: one round trip, see recover_pass)
                const double sl = w.lo[pr], su = w.up[pr], y = w.yp[pr], g = w.gp[pr], s = w.sl[pr], zlv = w.zl[pr], zuv = w.zu[pr];
                const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
                if (!hl && !hu) continue;
                vsum[0] += fabs(y);
                vsum[3] += 1.0;
                if (sl == su) { vmax[1] = fmax(vmax[1], fabs(g - sl)); continue; }
                vmax[1] = fmax(vmax[1], fabs(g - s));
                double dsl = -y;
                if (hl) { const double z = zlv, pd = (s - sl) * z; dsl -= z; vsum[1] += z; vsum[2] += 1.0; vmax[2] = fmax(vmax[2], pd); vmin[0] = fmin(vmin[0], pd); }
                if (hu) { const double z = zuv, pd = (su - s) * z; dsl += z; vsum[1] += z; vsum[2] += 1.0; vmax[2] = fmax(vmax[2], pd); vmin[0] = fmin(vmin[0], pd); }
                vmax[0] = fmax(vmax[0], fabs(dsl));
            }
        }
    CMPC_LANES_END_NOSYNC
    cta.template reduce3<3, 1, 4>(vmax, vmin, vsum);
    CMPC_LANES
        if (lane == 0) {
            KktStats& s = sm.ks;
            s.dual = vmax[0]; s.viol = vmax[1]; s.pmax = vmax[2]; s.pmin = vmin[0];
            s.sum_y = vsum[0]; s.sum_z = vsum[1]; s.nb = vsum[2]; s.nrows = vsum[3];
        }
    CMPC_LANES_END
}

// Sigma and the barrier gradient terms of the path rows, then the small blocks of the Riccati sweep for every knot
template <int NT, int G>
CMPC_FN void barrier_pass(Team T, const Config& cfg, const WorkS& w, double mu, double dw, double dc)
{
    cta_align<G>(T);
    const int N = cfg.N;
    CMPC_LANES
        {
            const double* __restrict__ lo = w.lo; const double* __restrict__ up = w.up; const double* __restrict__ ypp = w.yp;
            const double* __restrict__ gpp = w.gp; const double* __restrict__ slp = w.sl; const double* __restrict__ zlp = w.zl;
            const double* __restrict__ zup = w.zu; double* __restrict__ sigp = w.sig; double* __restrict__ ttp = w.tt;
            CMPC_UNROLL2
            for (int pr = lane; pr < N * PS; pr += NT) {  // the two pad rows of every knot have no bounds: sig = tt = 0
                const double sl = lo[pr], su = up[pr], y = ypp[pr], g = gpp[pr], s = slp[pr], zl = zlp[pr], zu = zup[pr];
                const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
                double sg = 0.0, t = 0.0;
                if (hl || hu) {
                    if (sl == su) { sg = 1.0 / dc; t = y + (g - sl) / dc; }
                    else {
                        sg = dw;
                        if (hl) { const double rd = 1.0 / (s - sl); sg += zl * rd; t -= mu * rd; }
                        if (hu) { const double rd = 1.0 / (su - s); sg += zu * rd; t += mu * rd; }
                        if (hl && !hu) t += KAPPA_D * mu;
                        if (hu && !hl) t -= KAPPA_D * mu;
                        t += sg * (g - s);
                    }
                }
                sigp[pr] = sg; ttp[pr] = t;
            }
        }
    CMPC_LANES_END
    CMPC_LANES
        CMPC_ROLES(123) CMPC_KNOTS(0, N) {
            const double* t = w.tab + k * TS;
            double val = 0.0;
            if (r < 48) {  // friction barrier block of corner cj, packed entry e
                if (k < N) {
                    const int cj = r / 6, e = r - 6 * cj, c = cj >> 2;
                    const int a = e < 3 ? 0 : (e < 5 ? 1 : 2), b = e < 3 ? e : (e < 5 ? e - 2 : 2);
                    const double* ar = t + T_AR + 12 * c;
                    const double* sg = w.sig + k * PS + 4 * cj;
                    for (int q = 0; q < NF; ++q) val += sg[q] * ar[3 * q + a] * ar[3 * q + b];
                }
            } else if (r < 60) {  // step-box barrier block on pos_c of knot k (rows of knot k - 1)
                if (k > 0) {
                    const int t2 = r - 48, c = t2 / 6, e = t2 - 6 * c;
                    const int a = e < 3 ? 0 : (e < 5 ? 1 : 2), b = e < 3 ? e : (e < 5 ? e - 2 : 2);
                    const double* R = w.tab + (k - 1) * TS + T_R + 9 * c;
                    const double* sg = w.sig + (k - 1) * PS + 32 + 3 * c;
                    for (int q = 0; q < 3; ++q) val += sg[q] * R[3 * q + a] * R[3 * q + b];
                }
            } else if (r < 60 + NS) {
                const int i = r - 60;
                val = w.gr[k * ZS + i];
                if (k > 0 && i >= 9) {
                    const int c = (i - 9) / 3, a = (i - 9) % 3;
                    const double* R = w.tab + (k - 1) * TS + T_R + 9 * c;
                    const double* tb = w.tt + (k - 1) * PS + 32 + 3 * c;
                    for (int q = 0; q < 3; ++q) val += R[3 * q + a] * tb[q];
                }
            } else if (r < 60 + NS + NU) {
                if (k < N) {
                    const int u = r - 60 - NS;
                    val = w.gr[k * ZS + NS + u];
                    if (u >= 6) {
                        const int f = u - 6, c = f / 12, j = (f % 12) / 3, a = f % 3;
                        const double* ar = t + T_AR + 12 * c + a;
                        const double* tf = w.tt + k * PS + 16 * c + 4 * j;
                        for (int q = 0; q < NF; ++q) val += ar[3 * q] * tf[q];
                    }
                }
            } else if (r < 60 + NS + NU + NS) {
                if (k < N) val = -w.ceq[(k + 1) * ES + r - 60 - NS - NU];
            } else {
                if (k < N) val = w.lam[(k + 1) * ES + 6 + r - 60 - NS - NU - NS];
            }
            w.small[(size_t)k * SMALL_STRIDE + r] = val;  // SmallBlk is laid out in item order
        }
    CMPC_LANES_END
}

// eliminated rows of the Newton system from dz (W-B eq. 13), then the multipliers of the equality rows by the adjoint
// recursion  lambda+_k = A_k' lambda+_{k+1} - [grad f + W dz + sum_box a_i (y_i + dy_i)]_{s_k}  (exact for the given dz)
template <int NT, int G>
CMPC_FN void recover_pass(Team T, const Config& cfg, const WorkS& w, double mu, double dw, double dc, bool pc)
{
    cta_align<G>(T);
    const int N = cfg.N;
    auto row = [&](int k, int l) {   // friction rows role major, step-box rows and the force sums flat: see affine_pass
            {
                const int pr = k * PS + l;
#if defined(__CUDA_ARCH__) && CMPC_HOIST_LOADS
                // every global operand of the row is requested before the first branch on any of them: one round trip per row
                // instead of bounds -> (row value, slack) -> multipliers
                const double sl = w.lo[pr], su = w.up[pr], gpv = w.gp[pr], sv = w.sl[pr], ypv = w.yp[pr], zlv = w.zl[pr], zuv = w.zu[pr],
                             sgv = w.sig[pr], cclv = pc ? w.ccl[pr] : 0.0, ccuv = pc ? w.ccu[pr] : 0.0;
                const double jd = path_dot(w.tab + k * TS, w.dz + k * ZS, w.dz + (k + 1) * ZS, l, 0);
                const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
                double ds = 0.0, dzl = 0.0, dzu = 0.0, dy = 0.0;
                if (hl || hu) {
                    if (sl == su) dy = (jd + (gpv - sl)) / dc;
                    else {
                        ds = jd + (gpv - sv);
                        double rs = -ypv;
                        const double ml = mu - cclv, mup = mu + ccuv;
                        if (hl) { const double rd = 1.0 / (sv - sl); rs -= ml * rd; dzl = ml * rd - zlv - zlv * rd * ds; }
                        if (hu) { const double rd = 1.0 / (su - sv); rs += mup * rd; dzu = mup * rd - zuv + zuv * rd * ds; }
                        if (hl && !hu) rs += KAPPA_D * mu;
                        if (hu && !hl) rs -= KAPPA_D * mu;
                        dy = sgv * ds + rs;
                    }
                }
                w.dsl[pr] = ds; w.dzl[pr] = dzl; w.dzu[pr] = dzu; w.dyp[pr] = dy;
                w.ypn[pr] = ypv + dy;
#else
                const double sl = w.lo[pr], su = w.up[pr];
                const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
                double ds = 0.0, dzl = 0.0, dzu = 0.0, dy = 0.0;
                if (hl || hu) {
                    const double jd = path_dot(w.tab + k * TS, w.dz + k * ZS, w.dz + (k + 1) * ZS, l, 0);
                    if (sl == su) dy = (jd + (w.gp[pr] - sl)) / dc;
                    else {
                        const double s = w.sl[pr];
                        ds = jd + (w.gp[pr] - s);  // (g - s) first: jd can be below ulp(g)
                        double rs = -w.yp[pr];
                        // complementarity targets: mu, or mu -+ ds_aff dz_aff (Mehrotra's second-order corrector)
                        const double ml = pc ? mu - w.ccl[pr] : mu, mup = pc ? mu + w.ccu[pr] : mu;
                        if (hl) { const double rd = 1.0 / (s - sl), zl = w.zl[pr]; rs -= ml * rd; dzl = ml * rd - zl - zl * rd * ds; }
                        if (hu) { const double rd = 1.0 / (su - s), zu = w.zu[pr]; rs += mup * rd; dzu = mup * rd - zu + zu * rd * ds; }
                        if (hl && !hu) rs += KAPPA_D * mu;
                        if (hu && !hl) rs -= KAPPA_D * mu;
                        dy = w.sig[pr] * ds + rs;
                    }
                }
                w.dsl[pr] = ds; w.dzl[pr] = dzl; w.dzu[pr] = dzu; w.dyp[pr] = dy;
                w.ypn[pr] = w.yp[pr] + dy;
#endif
            }
    };
    CMPC_LANES
        CMPC_ROLES(32) CMPC_KNOTS(0, N - 1) row(k, r);
        for (int it = lane; it < N * 6; it += NT) { const int k = it / 6; row(k, 32 + it - 6 * k); }
        for (int it = lane; it < N * 6; it += NT) {   // step of the total force of contact c (for the bilinear hessian terms)
            const int k = it / 6, q = it - 6 * k, c = q / 3, a = q - 3 * c;
            const double* dc4 = w.dz + k * ZS + NS + 6 + 12 * c + a;
            w.dfc[k * 8 + q] = dc4[0] + dc4[3] + dc4[6] + dc4[9];
        }
    CMPC_LANES_END
    CMPC_LANES
        for (int it = lane; it < (N + 1) * NS; it += NT) {   // flat (15 roles: see kkt_pass)
            const int k = it / NS, i = it - NS * k;
            double v = w.gr[k * ZS + i] + hess_dz_entry(cfg, w, dw, k, i);
            if (k > 0 && i >= 9) {
                const int c = (i - 9) / 3, a = (i - 9) % 3;
                const double* R = w.tab + (k - 1) * TS + T_R + 9 * c;
                const double* yb = w.ypn + (k - 1) * PS + 32 + 3 * c;
                v += R[a] * yb[0] + R[3 + a] * yb[1] + R[6 + a] * yb[2];
            }
            w.vco[k * ES + i] = v;
        }
    CMPC_LANES_END
    // adjoint recursion on warp 0: lane i holds lambda+_k[i]; A' couples it with two other lanes
    {
        LaneVal lamp;
        CMPC_WARP0
            const double v = lane < NS ? -w.vco[N * ES + lane] : 0.0;
            lamp.at(lane) = v;
            if (lane < NS) { w.lamn[N * ES + lane] = v; w.dlam[N * ES + lane] = v - w.lam[N * ES + lane]; }
        CMPC_WARP0_END
#if defined(__CUDA_ARCH__)
        // the operands of knot k - 1 (global memory) are fetched while knot k runs: the chain of a step is two shuffles and three
        // multiply-adds instead of an L2 / DRAM round trip (the other warps of the team wait for this chain at the barrier)
        if (T.lane < 32 && T.on) {
            const int lane = T.lane;
            const bool act = lane < NS;
            int rr[2] = {0, 0};
            if (act) acol_rows(lane, rr);
            const double* vcop = w.vco + lane;
            const double* awp = w.aw + 2 * lane;
            const double* lamq = w.lam + lane;
            double vn = act ? vcop[(N - 1) * ES] : 0.0, a0n = act ? awp[(N - 1) * AWS] : 0.0, a1n = act ? awp[(N - 1) * AWS + 1] : 0.0;
            double ln = act ? lamq[(N - 1) * ES] : 0.0;
            CMPC_ROLLED
            for (int k = N - 1; k >= 0; --k) {
                const double g0 = __shfl_sync(0xffffffffu, lamp.r, rr[0]), g1 = __shfl_sync(0xffffffffu, lamp.r, rr[1]);
                const double vc = vn, a0 = a0n, a1 = a1n, lc = ln;
                if (k > 0 && act) { vn = vcop[(k - 1) * ES]; a0n = awp[(k - 1) * AWS]; a1n = awp[(k - 1) * AWS + 1]; ln = lamq[(k - 1) * ES]; }
                if (act) {
                    const double v = fma(a1, g1, fma(a0, g0, lamp.r - vc));
                    lamp.r = v;
                    w.lamn[k * ES + lane] = v;
                    w.dlam[k * ES + lane] = v - lc;
                }
            }
        }
#else
        CMPC_IF_WARP0
        {
            CMPC_ROLLED
            for (int k = N - 1; k >= 0; --k) {
                lamp.snapshot();
                CMPC_WARP0
                    int rr[2] = {0, 0};
                    if (lane < NS) acol_rows(lane, rr);
                    const double a0 = lamp.gather(lane, rr[0]), a1 = lamp.gather(lane, rr[1]);
                    if (lane < NS) {
                        const double v = -w.vco[k * ES + lane] + lamp.at(lane) + w.aw[k * AWS + 2 * lane] * a0 + w.aw[k * AWS + 2 * lane + 1] * a1;
                        lamp.at(lane) = v;
                        w.lamn[k * ES + lane] = v;
                        w.dlam[k * ES + lane] = v - w.lam[k * ES + lane];
                    }
                CMPC_WARP0_END
            }
        }
#endif
    }
    team_sync<NT, G>(T);
}

// predictor of the Mehrotra mode.  From the affine-scaling step dz (complementarity target 0): the slack / bound-multiplier
// parts of the step, their products (the second-order terms), the step lengths to the boundary and the mean complementarity
// after the step -> the new barrier parameter (returned in mu).  The corrector only changes the right hand side of the
// Newton system: w.res receives  J_path' dt  with  dt = change of the barrier gradient terms for the complementarity targets
// mu -+ ds_aff dz_aff , the right hand side of one refinement sweep (refine_backward + riccati_forward).
template <int NT, int G, class Cta>
CMPC_FN void affine_pass(Team T, Cta& cta, const Config& cfg, ISmem& sm, double mu_min, double& mu)
{
    cta_align<G>(T);
    const WorkS& w = sm.w;
    const int N = cfg.N;
    double vmax[2] = {1.0, 1.0};                  // 1 / alpha_aff (primal), 1 / alpha_aff (dual)
    double vsum[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // s00, s10, s01, s11, number of bounds
    // The 32 friction rows of a knot fill a warp (role major: lane = row, warps split the knots); the 6 step-box rows would
    // leave 26 lanes idle on a second trip through the knots: they are taken flat (item = knot x row) in one trip of the team.
    // A trip costs a memory round trip plus its instructions whatever the number of active lanes.
    auto row = [&](int k, int r) {
            const int pr = k * PS + r;
            const double sl = w.lo[pr], su = w.up[pr], s = w.sl[pr], gpv = w.gp[pr], zlv = w.zl[pr], zuv = w.zu[pr];
            const double jd = path_dot(w.tab + k * TS, w.dz + k * ZS, w.dz + (k + 1) * ZS, r, 0);
            const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
            double cl = 0.0, cu = 0.0, rl = 0.0, ru = 0.0;
            if ((hl || hu) && !(sl == su)) {
                const double ds = jd + (gpv - s);
                if (hl) {
                    const double d = s - sl, z = zlv;
                    rl = 1.0 / d;
                    const double q = ds * rl, dzv = -z - z * q;
                    vmax[0] = fmax(vmax[0], -q); vmax[1] = fmax(vmax[1], 1.0 + q);
                    vsum[0] += d * z; vsum[1] += ds * z; vsum[2] += d * dzv; vsum[3] += ds * dzv; vsum[4] += 1.0;
                    cl = ds * dzv;
                }
                if (hu) {
                    const double d = su - s, z = zuv;
                    ru = 1.0 / d;
                    const double q = ds * ru, dzv = -z + z * q;
                    vmax[0] = fmax(vmax[0], q); vmax[1] = fmax(vmax[1], 1.0 - q);
                    vsum[0] += d * z; vsum[1] -= ds * z; vsum[2] += d * dzv; vsum[3] -= ds * dzv; vsum[4] += 1.0;
                    cu = ds * dzv;
                }
            }
            w.ccl[pr] = cl; w.ccu[pr] = cu;
            w.dzl[pr] = rl; w.dzu[pr] = ru;   // scratch until the recover pass: 1 / distance to the bound
    };
    CMPC_LANES
        CMPC_ROLES(32) CMPC_KNOTS(0, N - 1) row(k, r);
        for (int it = lane; it < N * 6; it += NT) { const int k = it / 6; row(k, 32 + it - 6 * k); }
    CMPC_LANES_END_NOSYNC
    cta.template reduce3<2, 0, 5>(vmax, vmax, vsum);
    if (T.on && vsum[4] > 0.0 && vsum[0] > 0.0) {
        // Mehrotra's rule: sigma = (mu_aff / mu_cur)^3, mu = sigma mu_cur
        const double ap = 1.0 / vmax[0], ad = 1.0 / vmax[1];
        const double aff = vsum[0] + ap * vsum[1] + ad * vsum[2] + ap * ad * vsum[3];
        const double sg = fmin(1.0, fmax(0.0, aff / vsum[0]));
        mu = fmax(mu_min, sg * sg * sg * vsum[0] / vsum[4]);
    }
    const double mun = mu;
    CMPC_LANES
        for (int pr = lane; pr < N * PS; pr += NT) {   // flat: nothing here depends on the role (the two pad rows of a knot hold zeros)
            const double rl = w.dzl[pr], ru = w.dzu[pr], cu = w.ccu[pr], cl = w.ccl[pr];
            double dt = ru * (mun + cu) - rl * (mun - cl);
            if (rl != 0.0 && ru == 0.0) dt += KAPPA_D * mun;
            if (ru != 0.0 && rl == 0.0) dt -= KAPPA_D * mun;
            w.dyp[pr] = dt;
        }
    CMPC_LANES_END
    CMPC_LANES
        // J_path' dt touches 30 of the 45 variables of a knot (24 corner forces, 6 foot positions): exactly one warp of roles;
        // the other 15 entries are zeros, written flat (two trips of the team instead of a second, half-empty role trip per knot)
        CMPC_ROLES(30) CMPC_KNOTS(0, N) {
            double v = 0.0;
            int e;   // entry of the knot's variable block
            if (r >= 24) {
                const int q = r - 24, c = q / 3, a = q - 3 * c;
                e = 9 + q;
                if (k > 0) {
                    const double* R = w.tab + (k - 1) * TS + T_R + 9 * c;
                    const double* tb = w.dyp + (k - 1) * PS + 32 + 3 * c;
                    v = R[a] * tb[0] + R[3 + a] * tb[1] + R[6 + a] * tb[2];
                }
            } else {
                if (k == N) continue;
                const int f = r, c = f / 12, j = (f % 12) / 3, a = f % 3;
                e = NS + 6 + f;
                const double* ar = w.tab + k * TS + T_AR + 12 * c + a;  // row q at ar[3 q]
                const double* tf = w.dyp + k * PS + 16 * c + 4 * j;
                v = ar[0] * tf[0] + ar[3] * tf[1] + ar[6] * tf[2] + ar[9] * tf[3];
            }
            w.res[k * ZS + e] = v;
        }
        for (int it = lane; it < (N + 1) * 15; it += NT) {   // s rows 0 .. 8 and the contact velocities
            const int k = it / 15, q = it - 15 * k;
            if (k == N && q >= 9) continue;
            w.res[k * ZS + (q < 9 ? q : NS + q - 9)] = 0.0;
        }
    CMPC_LANES_END
}

// residual of the linearised stationarity (right hand side of the refinement) and, for the line search, the fraction to the
// boundary (eq. 15) and the directional derivative of the barrier function
template <int NT, int G, class Cta>
CMPC_FN void step_pass(Team T, Cta& cta, const Config& cfg, ISmem& sm, double mu, double dw, double tau)
{
    cta_align<G>(T);
    const WorkS& w = sm.w;
    const int N = cfg.N;
    double vmax[2] = {0.0, 0.0};    // rho, non-finite flag
    double vmin[2] = {1.0, 1.0};    // alpha_max (primal), alpha_z
    double vsum[1] = {0.0};         // dphi
    CMPC_LANES
        CMPC_ROLES(48 + 38) CMPC_KNOTS(0, N) {
            if (r < NS + NU) {
                if (k == N && r >= NS) continue;
                const double d = w.dz[k * ZS + r], g = w.gr[k * ZS + r];
                double v = g + jty_entry(cfg, w, w.lamn, w.ypn, k, r) + hess_dz_entry(cfg, w, dw, k, r);
                if (r >= NS && r < NS + 6 && w.tab[k * TS + T_VM + (r - NS) / 3] != 0.0) v = 0.0;  // variable held fixed
                w.res[k * ZS + r] = v;
                vmax[0] = fmax(vmax[0], fabs(v));
                vsum[0] += g * d;
                if (!(fabs(d) < HUGE_VAL)) vmax[1] = 1.0;
            } else if (k < N && r >= 48) {
                const int pr = k * PS + r - 48;
                const double sl = w.lo[pr], su = w.up[pr], s = w.sl[pr], ds = w.dsl[pr], dzlv = w.dzl[pr], dzuv = w.dzu[pr],
                             zlv = w.zl[pr], zuv = w.zu[pr];
                const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
                if ((!hl && !hu) || sl == su) continue;
                if (hl) {
                    const double dd = s - sl, dz = dzlv;
                    vsum[0] -= mu * ds / dd;
                    if (ds < 0) vmin[0] = fmin(vmin[0], -tau * dd / ds);
                    if (dz < 0) vmin[1] = fmin(vmin[1], -tau * zlv / dz);
                }
                if (hu) {
                    const double dd = su - s, dz = dzuv;
                    vsum[0] += mu * ds / dd;
                    if (ds > 0) vmin[0] = fmin(vmin[0], tau * dd / ds);
                    if (dz < 0) vmin[1] = fmin(vmin[1], -tau * zuv / dz);
                }
                if (hl && !hu) vsum[0] += KAPPA_D * mu * ds;
                if (hu && !hl) vsum[0] -= KAPPA_D * mu * ds;
            }
        }
    CMPC_LANES_END_NOSYNC
    cta.template reduce3<2, 2, 1>(vmax, vmin, vsum);
    CMPC_LANES
        if (lane == 0) {
            StepStats& s = sm.ss;
            s.rho = vmax[0]; s.bad = vmax[1]; s.amax = vmin[0]; s.az = vmin[1]; s.dphi = vsum[0];
        }
    CMPC_LANES_END
}

// ------------------------------------------------------------------------------------------------ the solver
// Persistent driver of one team: pulls instances from the work queue and solves them one after the other.  With G > 1
// teams per CTA the teams run in lock-step: every sub-round (set-up, factorisation retry, refinement, re-evaluation for a
// new barrier parameter, line-search trial, write-back) is executed by all teams when ANY team needs it (vote_any), the
// teams that do not need it skip the work inside the phases (T.on) but never a barrier.
// Batched arrays in the reference's CasADi order, instance major; x: in = initial guess, out = solution; lam: multipliers of g
// (out; in when warm_duals).
// Work distribution: the FIRST instance of a team is static, first_inst = team * (number of CTAs) + CTA, so that a batch
// smaller than the number of resident teams is spread over all SMs (one team per SM first, then two, ...) instead of
// filling the seven teams of the first CTAs; further instances come from the atomic queue, starting at queue_base = number
// of teams of the grid.
template <int NT, int G, class Cta>
CMPC_HD void ipm_run(Team T, Cta& cta, const Config& cfg, double* scratch, ISmem& ism, const unsigned short* cmap, int batch,
                     const double* p_all, const double* lbg_all, const double* ubg_all, double* x_all, double* lam_all,
                     double* obj_all, int* status_all, int* iters_all, int warm_duals, unsigned int* counter,
                     int first_inst = 0, int queue_base = 1, double* ric_region = nullptr)
{
    const int N = cfg.N, n = dim_x(N), np = dim_p(N), m = dim_g(N);
    WSmem& sm = ism.sw;
    WorkS& w = ism.w;
    SweepIO& io = ism.io;
    T.on = true;
    CMPC_LANES
        if (lane == 0) {
            sweep_barriers_init(sm);
            works_carve(scratch, N, w);
            if (ric_region) w.ric = ric_region;   // the factor blocks of all teams live apart from the iterate vectors (L2 window)
            io.sd = w.sd; io.small = w.small; io.ric = w.ric; io.ceq = w.ceq; io.dz = w.dz; io.res = w.res; io.cmap = cmap;
        }
    CMPC_LANES_END
    double* filt_t = ism.filt_t;
    double* filt_p = ism.filt_p;
    const double refine_tol = fmin(REFINE_TOL, 0.1 * cfg.tol);  // residual of the linear system that triggers a refinement sweep

    // state of the instance the team is working on (thread-private copies, uniform over the team)
    bool alive = true;   // the queue still has work for this team
    int inst = -1;       // instance index, -1 = none
    Instance in{nullptr, nullptr, nullptr};
    double* x_io = nullptr;
    double* lam_io = nullptr;
    bool warm = false;
    double mu = cfg.mu_init, tau = TAU_MIN, f = 0, theta0 = 0, bar0 = 0, theta_max = 0, theta_min = 0, dw_last = 0, E0 = 0;
    int nfilt = 0, it = 0, it_base = 0;
    // IPOPT's gradient-based NLP scaling (nlp_scaling_method, its default): the solver works on  sf * f  with
    // sf = min(1, nlp_scaling_max_gradient / |grad f(x0)|_inf).  Instead of scaling the weights in every hot loop the driver
    // runs the UNSCALED problem with the equivalent parameters: for the scaled problem's (mu, z, lambda, phi, delta_w,
    // delta_c) the unscaled run uses (mu / sf, z / sf, lambda / sf, phi / sf, delta_w / sf, delta_c * sf) -- the Newton
    // system, the step lengths and the iterates are identical -- and every test IPOPT poses to the scaled problem (E_mu,
    // filter, switching condition, mu update) is evaluated on the re-scaled statistics.  Rows of g are never scaled: every
    // entry of the constraint jacobian is bounded by 1, dT, dT |rho| or dT |F| (far below 100 for any physical input).
    double sf = 1.0, isf = 1.0, mu_min = fmin(cfg.tol, 1e-4) / (KAPPA_EPS + 1.0);
    int nacc = 0;            // consecutive iterates that pass IPOPT's acceptable-level test
    bool pc = cfg.pc != 0;   // Mehrotra predictor-corrector barrier update (false: IPOPT's monotone update)
    bool redo = false;       // the predictor-corrector run of the instance failed: solve it again on the monotone path
    bool first = true;       // the next instance of the team is its static one

    for (;;) {
        // ---- work queue: a team without an instance takes the next one
        const bool want = alive && inst < 0;
        T.on = want;
        CMPC_LANES
            if (lane == 0) {
                if (first) ism.inst = first_inst;
                else {
#if defined(__CUDA_ARCH__)
                    ism.inst = queue_base + (int)atomicAdd(counter, 1u);
#else
                    ism.inst = queue_base + (int)((*counter)++);
#endif
                }
            }
        CMPC_LANES_END
        if (want) first = false;
        bool fresh = false;
        if (want) {
            inst = ism.inst;
            if (inst >= batch) { alive = false; inst = -1; }
            else { fresh = true; pc = cfg.pc != 0; it_base = 0; }
        }
        if (redo) { fresh = true; redo = false; }
        if (!vote_any<G>(T, alive)) break;
        int fin = -1;  // >= 0: the instance is finished with this status at the end of the round

        // ---- set-up of a fresh instance: table of constants, rows (classified, bounds relaxed), iterate; validation
        if (vote_any<G>(T, fresh)) {
            T.on = fresh;
            if (fresh) {
                in.p = p_all + (size_t)inst * np; in.lbg = lbg_all + (size_t)inst * m; in.ubg = ubg_all + (size_t)inst * m;
                x_io = x_all + (size_t)inst * n;
                lam_io = lam_all ? lam_all + (size_t)inst * m : nullptr;
                warm = warm_duals && lam_io;
            }
            double bad[1] = {0.0};
    CMPC_LANES
        for (int it = lane; it < (N + 1) * 128; it += NT) {
            const int k = it >> 7, r = it & 127;
            if (r >= TS) continue;
            const double* p = in.p;
            const bool kn = k < N;
            double v = 0.0;
            if (r < 2) v = kn ? p[p_en(N, r, k)] : 0.0;
            else if (r < 4) v = kn ? (1.0 - p[p_en(N, r - 2, k)]) * cfg.dT : 0.0;
            else if (r < 6) v = kn ? (((1.0 - p[p_en(N, r - 4, k)]) * cfg.dT == 0.0) ? 1.0 : 0.0) : 0.0;
            else if (r < 8) v = 0.0;
            else if (r < 32) { const int q = r - 8, c = q / 12, rr = (q % 12) / 3, a = q % 3; v = kn ? fric_coef(cfg, p + p_rot(N, c, k), rr, a) : 0.0; }
            else if (r < 50) { const int q = r - 32, c = q / 9; v = kn ? p[p_rot(N, c, k) + q % 9] : 0.0; }
            else if (r < 74) {
                const int q = r - 50, c = q / 12, j = (q % 12) / 3, a = q % 3;
                if (kn) { const double* R = p + p_rot(N, c, k); const double* cr = cfg.corner[c][j]; v = R[a] * cr[0] + R[3 + a] * cr[1] + R[6 + a] * cr[2]; }
            }
            else if (r < 80) { const int q = r - 74; v = p[p_nom(N, q / 3, k) + q % 3]; }
            else if (r < 86) { const int q = r - 80; v = kn ? p[p_nom(N, q / 3, k + 1) + q % 3] : 0.0; }
            else if (r < 89) v = p[p_comref(N, k) + r - 86];
            else if (r < 92) v = p[p_href(N, k) + r - 89];
            else if (r < 95) v = kn ? p[p_extf(N, k) + r - 92] : 0.0;
            else if (r < 98) v = kn ? p[p_extt(N, k) + r - 95] : 0.0;
            else if (r == 98) { const double om = com_z_omega(cfg.w_com[2], k); v = om * om; }
            w.tab[k * TS + r] = v;
        }
        for (int it = lane; it < (N + 1) * 64; it += NT) {
            const int k = it >> 6, r = it & 63;
            if (r < NS + NU) {
                double v = 0.0;
                if (!(k == N && r >= NS)) {
                    v = x_io[r < NS ? x_of_s(N, k, r) : x_of_u(N, k, r - NS)];
                    if (!(fabs(v) < HUGE_VAL)) bad[0] = 1.0;
                }
                w.z[k * ZS + r] = v;
                w.dz[k * ZS + r] = 0.0;
            } else if (r < ZS) {
                w.z[k * ZS + r] = 0.0; w.dz[k * ZS + r] = 0.0; w.zt[k * ZS + r] = 0.0; w.gr[k * ZS + r] = 0.0; w.res[k * ZS + r] = 0.0;
            } else if (r < ZS + ES) {
                const int i = r - ZS, e = k * ES + i;
                double b = 0.0, y = 0.0;
                if (i < NS) {
                    const int row = g_of_s(N, k, i);
                    b = in.lbg[row];
                    if (!(in.lbg[row] == in.ubg[row]) || !(fabs(b) < cfg.inf_bound)) bad[0] = 1.0;
                    if (warm) y = lam_io[row];
                }
                w.beq[e] = b; w.lam[e] = y; w.dlam[e] = 0.0; w.ceq[e] = 0.0; w.lamn[e] = 0.0; w.vco[e] = 0.0;
            }
        }
        for (int it = lane; it < N * 64; it += NT) {
            const int k = it >> 6, l = it & 63;
            if (l >= PS) continue;
            const int pr = k * PS + l;
            double lo = -HUGE_VAL, up = HUGE_VAL, y = 0.0;
            if (l < 38) {
                const int row = path_row(N, k, l);
                const double lb = in.lbg[row], ub = in.ubg[row];
                const bool hl = finite_lo(cfg, lb), hu = finite_up(cfg, ub);
                if (!(lb == lb) || !(ub == ub) || (hl && hu && lb > ub)) bad[0] = 1.0;
                if (hl && hu && lb == ub) { lo = lb; up = lb; }
                else {
                    if (hl) lo = lb - cfg.bound_relax * fmax(1.0, fabs(lb));
                    if (hu) up = ub + cfg.bound_relax * fmax(1.0, fabs(ub));
                }
                if (warm && (hl || hu)) y = lam_io[row];
            }
            w.lo[pr] = lo; w.up[pr] = up; w.yp[pr] = y;
            w.sl[pr] = 0.0; w.zl[pr] = 0.0; w.zu[pr] = 0.0; w.gp[pr] = 0.0; w.sig[pr] = 0.0; w.tt[pr] = 0.0;
            w.dsl[pr] = 0.0; w.dzl[pr] = 0.0; w.dzu[pr] = 0.0; w.dyp[pr] = 0.0; w.ypn[pr] = 0.0; w.slt[pr] = 0.0;
            w.ccl[pr] = 0.0; w.ccu[pr] = 0.0;
        }
    CMPC_LANES_END_NOSYNC
            cta.template maxv<1>(bad);
            team_sync<NT, G>(T);
            if (fresh && bad[0] != 0.0) { fin = 4; fresh = false; f = 0; it = 0; E0 = 0; }
            T.on = fresh;
            // objective scaling from the gradient at the initial point (IPOPT GradientScaling, nlp_scaling_min_value 1e-8)
            {
                double gmax[1] = {0.0};
                if (cfg.scal_max_grad > 0.0) {
    CMPC_LANES
        CMPC_ROLES(NS + NU) CMPC_KNOTS(0, N) {
            if (k == N && r >= NS) continue;
            gmax[0] = fmax(gmax[0], fabs(grad_entry(cfg, w, w.z, k, r)));
        }
    CMPC_LANES_END_NOSYNC
                    cta.template maxv<1>(gmax);
                    team_sync<NT, G>(T);
                }
                if (fresh) {
                    sf = (cfg.scal_max_grad > 0.0 && gmax[0] > cfg.scal_max_grad) ? fmax(cfg.scal_max_grad / gmax[0], 1e-8) : 1.0;
                    isf = 1.0 / sf;
                    // IPOPT: mu stops at min(tol, compl_inf_tol posed to the scaled problem) / (barrier_tol_factor + 1)
                    mu_min = fmin(cfg.tol, 1e-4 * sf) / (KAPPA_EPS + 1.0);
                }
            }
            // initial point: slacks pushed inside their bounds (bound_push / bound_frac), bound multipliers 1
            if (fresh) { mu = warm ? cfg.mu_warm : cfg.mu_init; tau = fmax(TAU_MIN, 1.0 - mu); nfilt = 0; dw_last = 0.0; it = 0; nacc = 0; }
            eval_point<NT, G>(T, cta, cfg, ism, w.z, w.sl);  // path row values (the slacks are not set yet)
    CMPC_LANES
        for (int it = lane; it < N * 64; it += NT) {
            const int k = it >> 6, l = it & 63;
            if (l >= 38) continue;
            const int pr = k * PS + l;
            const double sl = w.lo[pr], su = w.up[pr];
            const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
            if ((!hl && !hu) || sl == su) continue;
            double s = w.gp[pr];
            const double k1 = cfg.bound_push;
            if (hl && hu) {
                const double pl = fmin(k1 * fmax(1.0, fabs(sl)), k1 * (su - sl));
                const double pu = fmin(k1 * fmax(1.0, fabs(su)), k1 * (su - sl));
                s = fmin(fmax(s, sl + pl), su - pu);
            } else if (hl) s = fmax(s, sl + k1 * fmax(1.0, fabs(sl)));
            else s = fmin(s, su - k1 * fmax(1.0, fabs(su)));
            w.sl[pr] = s;
            double zl = hl ? isf : 0.0, zu = hu ? isf : 0.0;   // bound_mult_init_val 1 of the scaled problem
            if (warm) {
                const double yv = w.yp[pr];
                if (hl) zl = fmax(yv < 0 ? -yv : 0.0, cfg.mu_warm * isf / (s - sl));
                if (hu) zu = fmax(yv > 0 ? yv : 0.0, cfg.mu_warm * isf / (su - s));
            }
            w.zl[pr] = zl; w.zu[pr] = zu;
        }
    CMPC_LANES_END
            eval_point<NT, G>(T, cta, cfg, ism, w.z, w.sl);
            if (fresh) {
                f = ism.es.f; theta0 = ism.es.theta; bar0 = ism.es.bar;
                theta_max = 1e4 * fmax(1.0, theta0); theta_min = 1e-4 * fmax(1.0, theta0);
            }
        }

        // ---- one interior-point iteration of every team that holds an instance
        bool act = alive && inst >= 0 && fin < 0;
        CMPC_TIC
        T.on = act;
        kkt_pass<NT, G>(T, cta, cfg, ism);
        CMPC_TOC(1)
        KktStats ks = ism.ks;
        const double dual_u = ks.dual;   // dual infeasibility of the unscaled problem (dual_inf_tol applies to it)
        ks.dual *= sf; ks.pmax *= sf; ks.pmin *= sf; ks.sum_y *= sf; ks.sum_z *= sf;   // statistics of the scaled problem
        if (act) {
            double cmp0;
            E0 = kkt_E(ks, 0.0, &cmp0);
            const double cmp_u = cmp0 * isf;
            if (E0 <= cfg.tol && dual_u <= 1.0 && ks.viol <= 1e-4 && cmp_u <= 1e-4) { fin = 0; act = false; }
            else {
                // acceptable_tol / acceptable_dual_inf_tol 1e10 / acceptable_constr_viol_tol 1e-2 / acceptable_compl_inf_tol 1e-2
                if (cfg.acc_tol > 0.0 && E0 <= cfg.acc_tol && dual_u <= 1e10 && ks.viol <= 1e-2 && cmp_u <= 1e-2) ++nacc;
                else nacc = 0;
                if (cfg.acc_tol > 0.0 && nacc >= cfg.acc_iter) { fin = 5; act = false; }
                else if (it == cfg.max_iter || (pc && it == PC_MAX_ITER)) { fin = 1; act = false; }
            }
        }
        if (act && !pc) {
            // barrier update (eq. 7), filter reset
            while (kkt_E(ks, mu, nullptr) <= KAPPA_EPS * mu && mu > mu_min) {
                mu = fmax(mu_min, fmin(KAPPA_MU * mu, mu * sqrt(mu)));
                tau = fmax(TAU_MIN, 1.0 - mu);
                nfilt = 0;
            }
        }
        // ---- search direction with inertia correction (alg. IC): Cholesky failure inside the Riccati sweep <=> wrong inertia
        const double dc = fmax(DC_BAR * sqrt(sqrt(mu)), DC_FLOOR) * sf;   // delta_c of the scaled problem, times sf
        double dw = 0.0;                                                  // delta_w of the scaled problem, divided by sf
        int tries = 0;
        bool needf = act;
        while (vote_any<G>(T, needf)) {
            T.on = needf;
            barrier_pass<NT, G>(T, cfg, w, pc ? 0.0 : mu * isf, dw, dc);  // predictor-corrector: affine-scaling step first
            CMPC_TOC(2)
            const int rc = riccati_backward<NT, G>(T, cfg, io, sm, dw);
            CMPC_TOC(3)
            if (needf) {
                if (rc == 0) needf = false;
                else {
                    if (dw == 0.0) dw = (dw_last == 0.0 ? DW_FIRST : fmax(DW_MIN, KW_MINUS * dw_last)) * isf;
                    else dw *= (dw_last == 0.0 ? KW_PLUS_FIRST : KW_PLUS);
                    if (dw * sf > DW_MAX || ++tries > 60) { needf = false; fin = 3; act = false; }
                }
            }
        }
        if (act && dw > 0.0) dw_last = dw * sf;
        T.on = act;
        riccati_forward<NT, G>(T, cfg, io, sm, false);
        CMPC_TOC(4)
        // ---- Mehrotra's rule: barrier parameter from the affine-scaling step, then the corrector: same matrix, new right hand
        //      side (complementarity targets mu -+ ds_aff dz_aff) = one refinement sweep whose result is accumulated into dz
        {
            const bool pcact = act && pc;
            if (vote_any<G>(T, pcact)) {
                T.on = pcact;
                double mu_u = mu * isf;
                affine_pass<NT, G>(T, cta, cfg, ism, mu_min * isf, mu_u);
                if (pcact) { mu = mu_u * sf; tau = fmax(TAU_MIN, 1.0 - mu); nfilt = 0; }
                CMPC_TOC(5)
                refine_backward<NT, G>(T, cfg, io, sm);
                CMPC_TOC(15)
                riccati_forward<NT, G>(T, cfg, io, sm, true);
                CMPC_TOC(4)
                T.on = act;
            }
        }
        recover_pass<NT, G>(T, cfg, w, mu * isf, dw, dc, pc);
        CMPC_TOC(5)
        step_pass<NT, G>(T, cta, cfg, ism, mu * isf, dw, tau);
        StepStats ss = ism.ss;
        CMPC_TOC(6)
        // ---- iterative refinement on the stationarity residual of the Newton system (the eliminated rows hold exactly)
        double rho_prev = HUGE_VAL;
        int rf = 0;
        for (;;) {
            const bool needr = act && rf < MAX_REFINE && ss.rho * sf > refine_tol && !(ss.rho > 0.5 * rho_prev);
            if (!vote_any<G>(T, needr)) break;
            T.on = needr;
            if (needr) rho_prev = ss.rho;
            refine_backward<NT, G>(T, cfg, io, sm);
            riccati_forward<NT, G>(T, cfg, io, sm, true);
            recover_pass<NT, G>(T, cfg, w, mu * isf, dw, dc, pc);
            step_pass<NT, G>(T, cta, cfg, ism, mu * isf, dw, tau);
            if (needr) { ss = ism.ss; ++rf; }
        }
        CMPC_TOC(7)
        if (act && ss.bad != 0.0) { fin = 3; act = false; }
        const double amax = ss.amax, az = ss.az, dphi = ss.dphi * sf;
        // ---- filter line search (alg. A) on the scaled problem: phi = sf f + mu (barrier sum)
        const double theta = theta0, phi = sf * f + mu * bar0;  // barrier function of the current point for the current mu
        double amin;
        if (dphi < 0) {
            amin = fmin(GAMMA_THETA, GAMMA_PHI * theta / (-dphi));
            if (theta <= theta_min) amin = fmin(amin, DELTA_SW * pow(theta, S_THETA) / pow(-dphi, S_PHI));
        } else amin = GAMMA_THETA;
        amin *= GAMMA_ALPHA;
        double alpha = amax, ft = f, th_t = theta, ph_t = phi, bar_t = bar0;
        int accepted = 0, armijo = 0;
        bool pend = act;
        while (vote_any<G>(T, pend)) {
            T.on = pend;
            CMPC_LANES
                {
                    double* __restrict__ zt = w.zt; const double* __restrict__ z = w.z; const double* __restrict__ dz = w.dz;
                    double* __restrict__ slt = w.slt; const double* __restrict__ sl = w.sl; const double* __restrict__ dsl = w.dsl;
                    CMPC_UNROLL4
                    for (int i = lane; i < (N + 1) * ZS; i += NT) zt[i] = z[i] + alpha * dz[i];
                    CMPC_UNROLL4
                    for (int i = lane; i < N * PS; i += NT) slt[i] = sl[i] + alpha * dsl[i];
                }
            CMPC_LANES_END
            eval_point<NT, G>(T, cta, cfg, ism, w.zt, w.slt);
            if (pend) {
                ft = ism.es.f; th_t = ism.es.theta; bar_t = ism.es.bar; ph_t = sf * ft + mu * bar_t;
                bool ok = (fabs(ph_t) < HUGE_VAL) && (fabs(th_t) < HUGE_VAL) && th_t <= theta_max;
                for (int q = 0; ok && q < nfilt; ++q)
                    if (th_t >= filt_t[q] && ph_t >= filt_p[q]) ok = false;
                if (ok) {
                    const bool sw = dphi < 0 && theta <= theta_min && alpha * pow(-dphi, S_PHI) > DELTA_SW * pow(theta, S_THETA);
                    const double slack = 10.0 * 2.2e-16 * fabs(phi);
                    if (sw) {
                        if (ph_t - phi - slack <= ETA_PHI * alpha * dphi) { accepted = 1; armijo = 1; }
                    } else if (th_t <= (1.0 - GAMMA_THETA) * theta || ph_t - slack <= phi - GAMMA_PHI * theta) {
                        accepted = 1; armijo = 0;
                    }
                }
                if (accepted) pend = false;
                else {
                    alpha *= 0.5;
                    if (alpha < 1e-16 || !(alpha >= amin)) pend = false;
                }
            }
        }
        CMPC_TOC(8)
        if (act && !accepted) { fin = 2; act = false; }  // IPOPT would start its restoration phase here (not restated)
        // ---- accept the trial point (ceq, gp, sd already hold its values): filter, buffer swap, multipliers
        T.on = act;
        CMPC_LANES
            if (lane == 0) {
                if (!armijo && nfilt < MAX_FILTER) { filt_t[nfilt] = (1.0 - GAMMA_THETA) * theta; filt_p[nfilt] = phi - GAMMA_PHI * theta; }
                double* tmp = w.z; w.z = w.zt; w.zt = tmp; tmp = w.sl; w.sl = w.slt; w.slt = tmp;
            }
        CMPC_LANES_END
        if (act) {
            if (!armijo && nfilt < MAX_FILTER) nfilt++;
            f = ft; theta0 = th_t; bar0 = bar_t;
        }
        CMPC_LANES
            {
                // every load of an item is issued before anything depends on it: one memory round trip per item
                double* __restrict__ yp = w.yp; double* __restrict__ zlp = w.zl; double* __restrict__ zup = w.zu;
                const double* __restrict__ lo = w.lo; const double* __restrict__ up = w.up; const double* __restrict__ slp = w.sl;
                const double* __restrict__ dyp = w.dyp; const double* __restrict__ dzl = w.dzl; const double* __restrict__ dzu = w.dzu;
                const double mu_u = mu * isf;
                CMPC_ROLLED
                for (int i = lane; i < N * PS; i += NT) {
                    const double sl = lo[i], su = up[i], s = slp[i], y = yp[i], dy = dyp[i], zl0 = zlp[i], dl = dzl[i], zu0 = zup[i], du = dzu[i];
                    const bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
                    const bool ineq = (hl || hu) && !(sl == su);
                    yp[i] = y + alpha * dy;
                    const double rl = 1.0 / (s - sl), ru = 1.0 / (su - s);  // one division per bound (footprint: an IEEE division is ~35 instructions)
                    const double zl1 = fmax(fmin(zl0 + az * dl, (KAPPA_SIGMA * mu_u) * rl), (mu_u / KAPPA_SIGMA) * rl);
                    const double zu1 = fmax(fmin(zu0 + az * du, (KAPPA_SIGMA * mu_u) * ru), (mu_u / KAPPA_SIGMA) * ru);
                    if (ineq && hl) zlp[i] = zl1;
                    if (ineq && hu) zup[i] = zu1;
                }
            }
            for (int i = lane; i < (N + 1) * ES; i += NT) w.lam[i] += alpha * w.dlam[i];
        CMPC_LANES_END
        CMPC_TOC(9)
#if defined(CMPC_TRACE) && !defined(__CUDA_ARCH__)
        if (act) printf("it %3d f %.10e E0 %.2e (d %.2e v %.2e) mu %.2e dw %.1e alpha %.3e az %.3e rho %.1e rf %d pc %d\n", it, f, E0, ks.dual, ks.viol, mu, dw, alpha, az, ss.rho, rf, (int)pc);
#endif
        if (act) ++it;
        // IPOPT stops at a point that passes the acceptable-level test when the algorithm cannot continue from it
        if ((fin == 2 || fin == 3) && nacc > 0) fin = 5;
        // an instance the predictor-corrector path cannot finish starts again from its initial point on the monotone path
        // (an exhausted ipopt_max_iteration is final, as in IPOPT: only the predictor-corrector's own iteration cap hands over)
        if (pc && (fin == 2 || fin == 3 || (fin == 1 && it < cfg.max_iter))) { pc = false; redo = true; it_base += it; fin = -1; }

        // ---- write-back of the instances that finished in this round (solution to the CasADi order), then back to the queue
        const bool done = fin >= 0;
        if (vote_any<G>(T, done)) {
            T.on = done && fin != 4;  // rejected input: x is left untouched
    CMPC_LANES
        for (int i2 = lane; i2 < (N + 1) * 64; i2 += NT) {
            const int k = i2 >> 6, r = i2 & 63;
            if (r < NS + NU) {
                if (k == N && r >= NS) continue;
                x_io[r < NS ? x_of_s(N, k, r) : x_of_u(N, k, r - NS)] = w.z[k * ZS + r];
            } else if (lam_io && r >= ZS && r < ZS + NS) {
                lam_io[g_of_s(N, k, r - ZS)] = w.lam[k * ES + r - ZS];
            }
        }
        if (lam_io)
            for (int i2 = lane; i2 < N * 64; i2 += NT) {
                const int k = i2 >> 6, l = i2 & 63;
                if (l < 38) lam_io[path_row(N, k, l)] = w.yp[k * PS + l];
            }
    CMPC_LANES_END
            if (done) {
                if (T.lane == 0) {
                    if (obj_all) obj_all[inst] = f;
                    if (status_all) status_all[inst] = fin;
                    if (iters_all) iters_all[inst] = it_base + it;
                }
                inst = -1;
            }
        }
    }
}

}  // namespace cmpc
