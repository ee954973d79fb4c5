// cmpc_generic.cuh -- TEST-ONLY: the first-generation, one-thread-CTA formulation of the solve (generic Riccati sweeps on
// dense blocks, CasADi-order vectors, monotone barrier update).  It is an independent second implementation of the same
// mathematics as the shipped kernels (csrc/cmpc_ipm.cuh, csrc/cmpc_warp.cuh) and is compiled ONLY into tests/hostsim
// (tests/test_hostsim.py, marker "not gpu"); the product library never contains it.
#pragma once

#include "../../paper_romualdi_2022_icra_centroidal-mpc-walking_b200/csrc/cmpc_core.cuh"

namespace cmpc {

// per-slot scratch in global memory (lives in L2 while the CTA works on the instance)
struct Work {
    double *x, *dx, *xt, *grad;                                     // n
    double *y, *dy, *g, *yn;                                        // m
    double *s, *ds, *st, *zL, *zU, *dzL, *dzU, *sL, *sU, *sig, *tt; // 38 N   (path rows, index 38 k + l)
    double *sd;                                                     // N * SD_STRIDE  stage data
    double *ric;                                                    // N * RIC_STRIDE Riccati factors
};

constexpr int LDL = NU + 1;          // padded row length of L in shared memory
constexpr int LDY = NXI + 1;         // 40 columns: 39 of H_ux + 1 of h_u
constexpr int LDP = NXI + 1;
constexpr int RIC_L = 0, RIC_Y = NU * NU, RIC_DINV = RIC_Y + NU * LDY, RIC_Z = RIC_DINV + NU,
              RIC_STRIDE = RIC_Z + NU + 4;  // per knot: L 30x30 | Y 30x40 | 1/diag(L) | z of the refinement sweep

CMPC_HD int work_doubles(int N)
{
    return 4 * dim_x(N) + 4 * dim_g(N) + 11 * INEQ_PER_KNOT * N + SD_STRIDE * N + RIC_STRIDE * N;
}
CMPC_HD void work_carve(double* base, int N, Work& w)
{
    const int n = dim_x(N), m = dim_g(N), q = INEQ_PER_KNOT * N;
    double* c = base;
    w.x = c; c += n; w.dx = c; c += n; w.xt = c; c += n; w.grad = c; c += n;
    w.y = c; c += m; w.dy = c; c += m; w.g = c; c += m; w.yn = c; c += m;
    w.s = c; c += q; w.ds = c; c += q; w.st = c; c += q; w.zL = c; c += q; w.zU = c; c += q;
    w.dzL = c; c += q; w.dzU = c; c += q; w.sL = c; c += q; w.sU = c; c += q; w.sig = c; c += q; w.tt = c; c += q;
    w.sd = c; c += SD_STRIDE * N;
    w.ric = c;
}

// shared-memory block of one CTA (doubles)
struct Smem {
    double P[NXI * LDP];     // cost-to-go hessian
    double GY[NXI * NU > NU * LDY ? NXI * NU : NU * LDY];  // G = P Bbar (39 x 30), later Y = L^-1 [H_ux | h_u] (30 x 40)
    double PA[NXI * NS];     // P[:, s] A
    double HL[NU * LDL];     // H_uu, then its Cholesky factor L
    double Hus[NU * NS];     // H_us before the solve
    double pv[NXI];          // cost-to-go gradient
    double wv[NXI];          // P bbar + p
    double hu[NU];
    double dxi[NXI];         // forward sweep state
    double du[NU];
    double tv[NU];
    double dinv[NU];
    double qv[NS], rv[NU], bv[NS];
    double Mf[NC * NJ * 6];  // friction barrier blocks (3x3 symmetric: xx xy xz yy yz zz)
    double Mb[NC * 6];       // step-box barrier blocks on pos_c of this knot
    double lamh[3];
    double red[64];          // reduction scratch
    int flag;
};

struct Result {
    int status;  // 0 converged, 1 max_iter, 2 line-search failure, 3 numerical failure, 4 bad input
    int iters;
    double obj, kkt;
};

// (A_k^T v)[i] for a 15-vector v living at rows g_of_s(k+1, .) of array yv
CMPC_HD double AT_y(const Config& cfg, const double* d, const double* yv, int k, int i)
{
    const int N = cfg.N;
    double v = yv[g_of_s(N, k + 1, i)];
    if (i < 3) {  // com: + dT (v_h x F_all)_i
        const double* vh = yv + g_h(N, k);
        int a1 = (i + 1) % 3, a2 = (i + 2) % 3;
        v += cfg.dT * (vh[a1] * d[SD_FALL + a2] - vh[a2] * d[SD_FALL + a1]);
    } else if (i < 6) {
        v += cfg.dT * yv[g_com(N, k) + i - 3];
    } else if (i >= 9) {
        int c = (i - 9) / 3, a = (i - 9) % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
        const double* vh = yv + g_h(N, k);
        const double* F = d + SD_FC + 3 * c;
        v -= cfg.dT * d[SD_EN + c] * (vh[a1] * F[a2] - vh[a2] * F[a1]);
    }
    return v;
}
// (B_k^T v)[i], same v
CMPC_HD double BT_y(const Config& cfg, const double* d, const double* yv, int k, int i)
{
    const int N = cfg.N;
    if (i < 6) {
        int c = i / 3, a = i % 3;
        return (1.0 - d[SD_EN + c]) * cfg.dT * yv[g_pos(N, c, k) + a];
    }
    int f = i - 6, c = f / 12, j = (f % 12) / 3, a = f % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    const double* vh = yv + g_h(N, k);
    const double* rho = d + SD_RHO + 3 * (4 * c + j);
    return cfg.dT * d[SD_EN + c] * (yv[g_dcom(N, k) + a] + vh[a1] * rho[a2] - vh[a2] * rho[a1]);
}

// r_x = grad f + J^T yv for variable (k, i), i in 0..44 (s then u); needs grad and sd at the current x
CMPC_HD double lag_grad_entry(const Config& cfg, const Instance& in, const Work& w, const double* yv, int k, int i)
{
    const int N = cfg.N;
    if (i < NS) {
        double r = w.grad[x_of_s(N, k, i)] + yv[g_of_s(N, k, i)];
        if (k < N) r -= AT_y(cfg, w.sd + k * SD_STRIDE, yv, k, i);
        if (k > 0 && i >= 9) {
            int c = (i - 9) / 3, a = (i - 9) % 3;
            const double* R = in.p + p_rot(N, c, k - 1);
            for (int q = 0; q < 3; ++q) r += R[3 * q + a] * yv[g_box(N, c, k - 1) + q];
        }
        return r;
    }
    int u = i - NS;
    double r = w.grad[x_of_u(N, k, u)] - BT_y(cfg, w.sd + k * SD_STRIDE, yv, k, u);
    if (u >= 6) {
        int f = u - 6, c = f / 12, j = (f % 12) / 3, a = f % 3;
        const double* R = in.p + p_rot(N, c, k);
        for (int q = 0; q < NF; ++q) r += fric_coef(cfg, R, q, a) * yv[g_fric(N, c, j, k) + q];
    }
    return r;
}

struct Errs { double dual, viol, compl_, E; };

// scaled optimality error E_mu (Waechter-Biegler eq. 5, 6); uses w.g, w.grad, w.sd at the current x
template <class Cta>
CMPC_FN Errs kkt_error(Cta& cta, const Config& cfg, const Instance& in, const Work& w, double mu)
{
    const int N = cfg.N;
    double vm[3] = {0, 0, 0};     // max: dual, viol, compl
    double vs[4] = {0, 0, 0, 0};  // sum: |y|, |z|, number of bounds, number of rows with a multiplier
    for (int it = cta.tid; it < (N + 1) * (NS + NU); it += cta.nt) {
        int k = it / (NS + NU), i = it % (NS + NU);
        if (k == N && i >= NS) continue;
        vm[0] = fmax(vm[0], fabs(lag_grad_entry(cfg, in, w, w.y, k, i)));
    }
    for (int it = cta.tid; it < (N + 1) * NS; it += cta.nt) {  // initial-condition + dynamics rows
        int row = g_of_s(N, it / NS, it % NS);
        vm[1] = fmax(vm[1], fabs(w.g[row] - in.lbg[row]));
        vs[0] += fabs(w.y[row]);
        vs[3] += 1;
    }
    for (int pr = cta.tid; pr < N * INEQ_PER_KNOT; pr += cta.nt) {
        int row = g_of_ineq(N, pr / INEQ_PER_KNOT, pr % INEQ_PER_KNOT);
        double sl = w.sL[pr], su = w.sU[pr];
        bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
        if (!hl && !hu) continue;
        vs[0] += fabs(w.y[row]);
        vs[3] += 1;
        if (sl == su) { vm[1] = fmax(vm[1], fabs(w.g[row] - sl)); continue; }
        vm[1] = fmax(vm[1], fabs(w.g[row] - w.s[pr]));
        double dsl = -w.y[row];
        if (hl) { dsl -= w.zL[pr]; vs[1] += w.zL[pr]; vs[2] += 1; vm[2] = fmax(vm[2], fabs((w.s[pr] - sl) * w.zL[pr] - mu)); }
        if (hu) { dsl += w.zU[pr]; vs[1] += w.zU[pr]; vs[2] += 1; vm[2] = fmax(vm[2], fabs((su - w.s[pr]) * w.zU[pr] - mu)); }
        vm[0] = fmax(vm[0], fabs(dsl));
    }
    cta.template maxv<3>(vm);
    cta.template sumv<4>(vs);
    Errs e;
    e.dual = vm[0]; e.viol = vm[1]; e.compl_ = vm[2];
    double sd = fmax(S_MAX, (vs[0] + vs[1]) / fmax(1.0, vs[3] + vs[2])) / S_MAX;
    double sc = fmax(S_MAX, vs[1] / fmax(1.0, vs[2])) / S_MAX;
    e.E = fmax(e.dual / sd, fmax(e.viol, e.compl_ / sc));
    return e;
}

// theta = l1 norm of (c(x); d(x) - s) for constraint values gv and slacks sv
template <class Cta>
CMPC_FN void theta_phi(Cta& cta, const Config& cfg, const Instance& in, const Work& w, const double* gv,
                       const double* sv, double f, double mu, double& theta, double& phi)
{
    const int N = cfg.N;
    double v[2] = {0, 0};
    for (int it = cta.tid; it < (N + 1) * NS; it += cta.nt) {
        int row = g_of_s(N, it / NS, it % NS);
        v[0] += fabs(gv[row] - in.lbg[row]);
    }
    for (int pr = cta.tid; pr < N * INEQ_PER_KNOT; pr += cta.nt) {
        int row = g_of_ineq(N, pr / INEQ_PER_KNOT, pr % INEQ_PER_KNOT);
        double sl = w.sL[pr], su = w.sU[pr];
        bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
        if (!hl && !hu) continue;
        if (sl == su) { v[0] += fabs(gv[row] - sl); continue; }
        double s = sv[pr];
        v[0] += fabs(gv[row] - s);
        if (hl) v[1] -= mu * log(s - sl);
        if (hu) v[1] -= mu * log(su - s);
        if (hl && !hu) v[1] += KAPPA_D * mu * (s - sl);
        if (hu && !hl) v[1] += KAPPA_D * mu * (su - s);
    }
    cta.template sumv<2>(v);
    theta = v[0];
    phi = f + v[1];
}

// B_bar^T applied to a 39-vector g given through an accessor  g(i)
template <class Acc>
CMPC_HD double BbarT(const Config& cfg, const double* d, int u, Acc g)
{
    if (u < 6) {
        int c = u / 3, a = u % 3;
        return (1.0 - d[SD_EN + c]) * cfg.dT * g(9 + 3 * c + a);
    }
    int f = u - 6, c = f / 12, j = (f % 12) / 3, a = f % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    const double* rho = d + SD_RHO + 3 * (4 * c + j);
    // sum_b skew(rho, b, a) g_h[b] = (g_h x rho)_a
    return cfg.dT * d[SD_EN + c] * (g(3 + a) + g(6 + a1) * rho[a2] - g(6 + a2) * rho[a1]) + g(NS + f);
}
// A^T applied to a 15-vector v(i)
template <class Acc>
CMPC_HD double AT15(const Config& cfg, const double* d, int i, Acc v)
{
    double r = v(i);
    if (i < 3) {
        int a1 = (i + 1) % 3, a2 = (i + 2) % 3;
        r += cfg.dT * (v(6 + a1) * d[SD_FALL + a2] - v(6 + a2) * d[SD_FALL + a1]);
    } else if (i < 6) {
        r += cfg.dT * v(i - 3);
    } else if (i >= 9) {
        int c = (i - 9) / 3, a = (i - 9) % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
        const double* F = d + SD_FC + 3 * c;
        r -= cfg.dT * d[SD_EN + c] * (v(6 + a1) * F[a2] - v(6 + a2) * F[a1]);
    }
    return r;
}

// per-stage small blocks into shared memory: lamh, Mf, Mb, qv, rv, bv   (k <= N; for k == N only Mb and qv)
template <class Cta>
CMPC_FN void stage_small(Cta& cta, const Config& cfg, const Instance& in, const Work& w, Smem& sm, int k)
{
    const int N = cfg.N;
    for (int it = cta.tid; it < 8 * 6 + 2 * 6 + NS + NU + NS + 3; it += cta.nt) {
        if (it < 48) {  // friction barrier block of corner (c, j), packed entry e
            if (k == N) continue;
            int cj = it / 6, e = it % 6, c = cj / 4, j = cj % 4;
            int a = e < 3 ? 0 : (e < 5 ? 1 : 2), b = e < 3 ? e : (e < 5 ? e - 2 : 2);
            const double* R = in.p + p_rot(N, c, k);
            double acc = 0;
            for (int r = 0; r < NF; ++r)
                acc += w.sig[INEQ_PER_KNOT * k + 6 + 16 * c + 4 * j + r] * fric_coef(cfg, R, r, a) * fric_coef(cfg, R, r, b);
            sm.Mf[it] = acc;
        } else if (it < 60) {  // step-box barrier block on pos_c of knot k (rows of knot k-1)
            int t = it - 48, c = t / 6, e = t % 6;
            int a = e < 3 ? 0 : (e < 5 ? 1 : 2), b = e < 3 ? e : (e < 5 ? e - 2 : 2);
            double acc = 0;
            if (k > 0) {
                const double* R = in.p + p_rot(N, c, k - 1);
                for (int q = 0; q < 3; ++q) acc += w.sig[INEQ_PER_KNOT * (k - 1) + 3 * c + q] * R[3 * q + a] * R[3 * q + b];
            }
            sm.Mb[t] = acc;
        } else if (it < 60 + NS) {
            int i = it - 60;
            double q = w.grad[x_of_s(N, k, i)];
            if (k > 0 && i >= 9) {
                int c = (i - 9) / 3, a = (i - 9) % 3;
                const double* R = in.p + p_rot(N, c, k - 1);
                for (int r = 0; r < 3; ++r) q += R[3 * r + a] * w.tt[INEQ_PER_KNOT * (k - 1) + 3 * c + r];
            }
            sm.qv[i] = q;
        } else if (it < 60 + NS + NU) {
            if (k == N) continue;
            int u = it - 60 - NS;
            double r = w.grad[x_of_u(N, k, u)];
            if (u >= 6) {
                int f = u - 6, c = f / 12, j = (f % 12) / 3, a = f % 3;
                const double* R = in.p + p_rot(N, c, k);
                for (int q = 0; q < NF; ++q) r += fric_coef(cfg, R, q, a) * w.tt[INEQ_PER_KNOT * k + 6 + 16 * c + 4 * j + q];
            }
            sm.rv[u] = r;
        } else if (it < 60 + NS + NU + NS) {
            if (k == N) continue;
            int i = it - 60 - NS - NU, row = g_of_s(N, k + 1, i);
            sm.bv[i] = -(w.g[row] - in.lbg[row]);
        } else {
            if (k == N) continue;
            int a = it - 60 - NS - NU - NS;
            sm.lamh[a] = w.y[g_h(N, k) + a];
        }
    }
    cta.sync();
}

CMPC_HD double Qbar(const Config& cfg, const Smem& sm, int k, double dw, int i, int j)
{
    if (i < NS && j < NS) {
        double v = (i == j) ? cost_diag_s(cfg, k, i) + dw : 0.0;
        if (i >= 9 && j >= 9 && (i - 9) / 3 == (j - 9) / 3) v += sm.Mb[6 * ((i - 9) / 3) + sym3((i - 9) % 3, (j - 9) % 3)];
        return v;
    }
    if (i == j && k >= 1) return 2.0 * cfg.w_rate[(i - NS) % 3];
    return 0.0;
}
CMPC_HD double Rblk(const Config& cfg, const Smem& sm, const double* d, int k, double dw, int u, int v)
{
    if (u < 6 || v < 6) {
        if (u != v) return 0.0;
        return d[SD_VM + u / 3] != 0.0 ? 1.0 : dw;
    }
    int f1 = u - 6, f2 = v - 6, c = f1 / 12;
    if (f2 / 12 != c) return 0.0;
    int j1 = (f1 % 12) / 3, a1 = f1 % 3, j2 = (f2 % 12) / 3, a2 = f2 % 3;
    double a4 = d[SD_EN + c] / NJ, r = 0.0;
    if (a1 == a2)
        r += j1 == j2 ? 2.0 * cfg.w_sym * (1.0 - 2.0 * a4 + NJ * a4 * a4) : 2.0 * cfg.w_sym * (NJ * a4 * a4 - 2.0 * a4);
    if (j1 == j2) r += sm.Mf[6 * (4 * c + j1) + sym3(a1, a2)];
    if (u == v) r += dw + (k >= 1 ? 2.0 * cfg.w_rate[a1] : 0.0);
    return r;
}
// W[f_cj a, s_j]: bilinear hessian of lam_h' g_h (nlp_hess_l, tmp.c:58926)
CMPC_HD double Sblk(const Config& cfg, const Smem& sm, const double* d, int u, int j)
{
    if (u < 6) return 0.0;
    int f = u - 6, c = f / 12, a = f % 3;
    if (j < 3) return cfg.dT * d[SD_EN + c] * skew(sm.lamh, a, j);
    if (j >= 9 && (j - 9) / 3 == c) return -cfg.dT * d[SD_EN + c] * skew(sm.lamh, a, (j - 9) % 3);
    return 0.0;
}

// ------------------------------------------------------------------------------------------------ Riccati
// backward sweep; returns 0 or 1 (some H_uu not positive definite -> caller regularises)
template <class Cta>
CMPC_FN int riccati_backward(Cta& cta, const Config& cfg, const Instance& in, const Work& w, Smem& sm, double dw)
{
    const int N = cfg.N;
    // terminal cost-to-go
    stage_small(cta, cfg, in, w, sm, N);
    for (int it = cta.tid; it < NXI * NXI; it += cta.nt) {
        int i = it / NXI, j = it % NXI;
        sm.P[i * LDP + j] = (i < NS && j < NS) ? Qbar(cfg, sm, N, dw, i, j) : 0.0;
    }
    for (int i = cta.tid; i < NXI; i += cta.nt) sm.pv[i] = i < NS ? sm.qv[i] : 0.0;
    cta.sync();
    for (int k = N - 1; k >= 0; --k) {
        const double* d = w.sd + k * SD_STRIDE;
        double* ric = w.ric + (size_t)k * RIC_STRIDE;
        stage_small(cta, cfg, in, w, sm, k);
        // S1: G = P Bbar (39x30), PA = P[:, s] A (39x15), wv = P bbar + p
        for (int it = cta.tid; it < NXI * (NU + NS + 1); it += cta.nt) {
            int i = it / (NU + NS + 1), col = it % (NU + NS + 1);
            const double* Pi = sm.P + i * LDP;
            if (col < NU) {
                int u = col;
                double v;
                if (u < 6) {
                    int c = u / 3, a = u % 3;
                    v = (1.0 - d[SD_EN + c]) * cfg.dT * Pi[9 + 3 * c + a];
                } else {
                    int f = u - 6, c = f / 12, j = (f % 12) / 3, a = f % 3;
                    const double* rho = d + SD_RHO + 3 * (4 * c + j);
                    double hs = 0;
                    for (int b = 0; b < 3; ++b) hs += skew(rho, b, a) * Pi[6 + b];
                    v = cfg.dT * d[SD_EN + c] * (Pi[3 + a] + hs) + Pi[NS + f];
                }
                sm.GY[i * NU + u] = v;
            } else if (col < NU + NS) {
                int j = col - NU;
                double v = Pi[j];
                if (j < 3) {
                    for (int b = 0; b < 3; ++b) v += cfg.dT * skew(d + SD_FALL, b, j) * Pi[6 + b];
                } else if (j < 6) {
                    v += cfg.dT * Pi[j - 3];
                } else if (j >= 9) {
                    int c = (j - 9) / 3, a = (j - 9) % 3;
                    double hs = 0;
                    for (int b = 0; b < 3; ++b) hs += skew(d + SD_FC + 3 * c, b, a) * Pi[6 + b];
                    v -= cfg.dT * d[SD_EN + c] * hs;
                }
                sm.PA[i * NS + j] = v;
            } else {
                double v = sm.pv[i];
                for (int j = 0; j < NS; ++j) v += Pi[j] * sm.bv[j];
                sm.wv[i] = v;
            }
        }
        cta.sync();
        // S2: H_uu (lower), H_us, h_u
        for (int it = cta.tid; it < NU * (NU + NS + 1); it += cta.nt) {
            int u = it / (NU + NS + 1), col = it % (NU + NS + 1);
            if (col < NU) {
                int v = col;
                if (v > u) continue;
                sm.HL[u * LDL + v] = Rblk(cfg, sm, d, k, dw, u, v)
                                     + BbarT(cfg, d, u, [&](int i) { return sm.GY[i * NU + v]; });
            } else if (col < NU + NS) {
                int j = col - NU;
                sm.Hus[u * NS + j] = Sblk(cfg, sm, d, u, j) + BbarT(cfg, d, u, [&](int i) { return sm.PA[i * NS + j]; });
            } else {
                sm.hu[u] = sm.rv[u] + BbarT(cfg, d, u, [&](int i) { return sm.wv[i]; });
            }
        }
        if (cta.tid == 0) sm.flag = 0;
        cta.sync();
        // S3: Cholesky H_uu = L L^T on warp 0 (lane = row)
        if (cta.warp == 0) {
            for (int j = 0; j < NU; ++j) {
                for (int i = j + cta.lane; i < NU; i += cta.wsize) {
                    double v = sm.HL[i * LDL + j];
                    if (i == j) sm.dinv[j] = v;  // original diagonal, for the relative pivot test
                    for (int q = 0; q < j; ++q) v -= sm.HL[i * LDL + q] * sm.HL[j * LDL + q];
                    sm.HL[i * LDL + j] = v;
                }
                cta.syncwarp();
                double dj = sm.HL[j * LDL + j], orig = sm.dinv[j];
                bool ok = dj > 1e-11 * fabs(orig) && dj > 0.0 && dj < HUGE_VAL;
                double di = ok ? 1.0 / sqrt(dj) : 1.0;
                cta.syncwarp();
                if (!ok && cta.lane == 0) sm.flag = 1;
                for (int i = j + cta.lane; i < NU; i += cta.wsize) sm.HL[i * LDL + j] *= di;
                if (cta.lane == 0) sm.dinv[j] = di;
                cta.syncwarp();
            }
        }
        cta.sync();
        if (sm.flag) { cta.sync(); return 1; }
        // S4: Y = L^-1 [H_us | H_uphi | h_u], one column per thread, in place in GY (30 x 40)
        for (int it = cta.tid; it < NU * LDY; it += cta.nt) {
            int u = it / LDY, c = it % LDY;
            double v;
            if (c < NS) v = sm.Hus[u * NS + c];
            else if (c < NXI) v = (k >= 1 && u == 6 + (c - NS)) ? -2.0 * cfg.w_rate[(c - NS) % 3] : 0.0;
            else v = sm.hu[u];
            sm.GY[it] = v;
        }
        cta.sync();
        for (int c = cta.tid; c < LDY; c += cta.nt) {
            if (k == 0 && c >= NS && c < NXI) continue;  // no previous forces at knot 0: columns stay zero
            for (int u = 0; u < NU; ++u) {
                double v = sm.GY[u * LDY + c];
                for (int q = 0; q < u; ++q) v -= sm.HL[u * LDL + q] * sm.GY[q * LDY + c];
                sm.GY[u * LDY + c] = v * sm.dinv[u];
            }
        }
        cta.sync();
        // S5: P <- Qbar + Abar' P Abar - Y'Y ; p <- qbar + Abar' w - Y' z     (lower triangle, then mirrored)
        for (int it = cta.tid; it < NXI * (NXI + 1); it += cta.nt) {
            int i = it / (NXI + 1), j = it % (NXI + 1);
            if (j < NXI) {
                if (j > i) continue;
                double v = Qbar(cfg, sm, k, dw, i, j);
                if (i < NS && j < NS) v += AT15(cfg, d, i, [&](int r) { return sm.PA[r * NS + j]; });
                for (int u = 0; u < NU; ++u) v -= sm.GY[u * LDY + i] * sm.GY[u * LDY + j];
                sm.P[i * LDP + j] = v;
            } else {
                double v = i < NS ? sm.qv[i] + AT15(cfg, d, i, [&](int r) { return sm.wv[r]; }) : 0.0;
                for (int u = 0; u < NU; ++u) v -= sm.GY[u * LDY + i] * sm.GY[u * LDY + NXI];
                sm.pv[i] = v;
            }
        }
        cta.sync();
        for (int it = cta.tid; it < NXI * NXI; it += cta.nt) {
            int i = it / NXI, j = it % NXI;
            if (j > i) sm.P[i * LDP + j] = sm.P[j * LDP + i];
        }
        // S6: factors to global memory for the forward sweep
        for (int it = cta.tid; it < NU * NU; it += cta.nt) ric[RIC_L + it] = sm.HL[(it / NU) * LDL + it % NU];
        for (int it = cta.tid; it < NU * LDY; it += cta.nt) ric[RIC_Y + it] = sm.GY[it];
        for (int it = cta.tid; it < NU; it += cta.nt) ric[RIC_DINV + it] = sm.dinv[it];
        cta.sync();
    }
    cta.sync();
    return 0;
}

// forward sweep: dx (all variables) and dy of the initial-condition / dynamics rows
// refine = true: correction sweep of the iterative refinement (zero constraint residuals, z from the refinement
// backward sweep, result ACCUMULATED into dx)
template <class Cta>
CMPC_FN void riccati_forward(Cta& cta, const Config& cfg, const Instance& in, const Work& w, Smem& sm, bool refine)
{
    const int N = cfg.N;
    for (int i = cta.tid; i < NXI; i += cta.nt) {
        double v = 0.0;
        if (i < NS && !refine) {
            v = -(w.g[i] - in.lbg[i]);
            w.dx[x_of_s(N, 0, i)] = v;
        }
        sm.dxi[i] = v;
    }
    cta.sync();
    for (int k = 0; k < N; ++k) {
        const double* ric = w.ric + (size_t)k * RIC_STRIDE;
        const double* d = w.sd + k * SD_STRIDE;
        for (int it = cta.tid; it < NU * NU; it += cta.nt) sm.HL[(it / NU) * LDL + it % NU] = ric[RIC_L + it];
        for (int it = cta.tid; it < NU * LDY; it += cta.nt) sm.GY[it] = ric[RIC_Y + it];
        for (int it = cta.tid; it < NU; it += cta.nt) sm.dinv[it] = ric[RIC_DINV + it];
        for (int i = cta.tid; i < NS; i += cta.nt) {
            int row = g_of_s(N, k + 1, i);
            sm.bv[i] = refine ? 0.0 : -(w.g[row] - in.lbg[row]);
        }
        cta.sync();
        for (int u = cta.tid; u < NU; u += cta.nt) {
            double v = refine ? ric[RIC_Z + u] : sm.GY[u * LDY + NXI];
            for (int j = 0; j < NXI; ++j) v += sm.GY[u * LDY + j] * sm.dxi[j];
            sm.tv[u] = v;
        }
        cta.sync();
        if (cta.warp == 0) {  // du = -L^-T t
            for (int i = NU - 1; i >= 0; --i) {
                double xi = sm.tv[i] * sm.dinv[i];
                cta.syncwarp();
                for (int j = cta.lane; j < i; j += cta.wsize) sm.tv[j] -= sm.HL[i * LDL + j] * xi;
                if (cta.lane == 0) sm.du[i] = -xi;
                cta.syncwarp();
            }
        }
        cta.sync();
        for (int u = cta.tid; u < NU; u += cta.nt) {
            int xi = x_of_u(N, k, u);
            w.dx[xi] = refine ? w.dx[xi] + sm.du[u] : sm.du[u];
        }
        // dxi_{k+1} = Abar dxi + Bbar du + bbar   (into qv/rv scratch first: dxi is still being read)
        for (int i = cta.tid; i < NXI; i += cta.nt) {
            double v;
            if (i >= NS) v = sm.du[6 + i - NS];
            else {
                v = sm.dxi[i] + sm.bv[i];
                if (i < 3) v += cfg.dT * sm.dxi[3 + i];
                else if (i < 6) {
                    int a = i - 3;
                    for (int c = 0; c < NC; ++c) {
                        double sfc = 0;
                        for (int j = 0; j < NJ; ++j) sfc += sm.du[6 + 12 * c + 3 * j + a];
                        v += cfg.dT * d[SD_EN + c] * sfc;
                    }
                } else if (i < 9) {
                    int a = i - 6, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                    // dT [F_all]x dcom
                    double t = d[SD_FALL + a1] * sm.dxi[a2] - d[SD_FALL + a2] * sm.dxi[a1];
                    for (int c = 0; c < NC; ++c) {
                        const double* F = d + SD_FC + 3 * c;
                        double tc = -(F[a1] * sm.dxi[9 + 3 * c + a2] - F[a2] * sm.dxi[9 + 3 * c + a1]);
                        for (int j = 0; j < NJ; ++j) {
                            const double* rho = d + SD_RHO + 3 * (4 * c + j);
                            const double* df = sm.du + 6 + 12 * c + 3 * j;
                            tc += rho[a1] * df[a2] - rho[a2] * df[a1];
                        }
                        t += d[SD_EN + c] * tc;
                    }
                    v += cfg.dT * t;
                } else {
                    int c = (i - 9) / 3, a = (i - 9) % 3;
                    v += (1.0 - d[SD_EN + c]) * cfg.dT * sm.du[3 * c + a];
                }
            }
            sm.wv[i] = v;
        }
        cta.sync();
        for (int i = cta.tid; i < NXI; i += cta.nt) {
            sm.dxi[i] = sm.wv[i];
            if (i < NS) {
                int xi = x_of_s(N, k + 1, i);
                w.dx[xi] = refine ? w.dx[xi] + sm.wv[i] : sm.wv[i];
            }
        }
        cta.sync();
    }
    cta.sync();
}


// ds, dy, dzL, dzU of the path rows from dx (eliminated block of the Newton system, W-B eq. 13)
template <class Cta>
CMPC_FN void recover_path(Cta& cta, const Config& cfg, const Instance& in, const Work& w, double mu, double dc)
{
    const int N = cfg.N;
    for (int pr = cta.tid; pr < N * INEQ_PER_KNOT; pr += cta.nt) {
        int k = pr / INEQ_PER_KNOT, l = pr % INEQ_PER_KNOT, row = g_of_ineq(N, k, l);
        double sl = w.sL[pr], su = w.sU[pr];
        bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
        w.ds[pr] = 0; w.dzL[pr] = 0; w.dzU[pr] = 0;
        if (!hl && !hu) { w.dy[row] = 0; continue; }
        int xv0 = path_var(cfg, k, l);
        double jd = 0;
        for (int a = 0; a < 3; ++a) jd += path_coef(cfg, in, k, l, a) * w.dx[xv0 + a];
        if (sl == su) { w.dy[row] = (jd + (w.g[row] - sl)) / dc; continue; }
        double s = w.s[pr], ds = jd + (w.g[row] - s), rs = -w.y[row];  // (g - s) first: jd can be below ulp(g)
        w.ds[pr] = ds;
        if (hl) { double dd = s - sl; rs -= mu / dd; w.dzL[pr] = mu / dd - w.zL[pr] - w.zL[pr] / dd * ds; }
        if (hu) { double dd = su - s; rs += mu / dd; w.dzU[pr] = mu / dd - w.zU[pr] + w.zU[pr] / dd * ds; }
        if (hl && !hu) rs += KAPPA_D * mu;
        if (hu && !hl) rs -= KAPPA_D * mu;
        w.dy[row] = w.sig[pr] * ds + rs;
    }
    cta.sync();
}

// (W dx)[variable (k, i)] without the barrier terms: hessian of the lagrangian (nlp_hess_l) times the step
CMPC_HD double hess_dx_entry(const Config& cfg, const Work& w, double dw, int k, int i)
{
    const int N = cfg.N;
    const double* d = w.sd + (k < N ? k : 0) * SD_STRIDE;
    if (i < NS) {
        double v = (cost_diag_s(cfg, k, i) + dw) * w.dx[x_of_s(N, k, i)];
        if (k < N && (i < 3 || i >= 9)) {
            const double* lam = w.y + g_h(N, k);
            int a = i < 3 ? i : (i - 9) % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
            double acc = 0;
            for (int c = (i < 3 ? 0 : (i - 9) / 3); c < (i < 3 ? NC : (i - 9) / 3 + 1); ++c) {
                double dF[3] = {0, 0, 0};
                for (int j = 0; j < NJ; ++j)
                    for (int b = 0; b < 3; ++b) dF[b] += w.dx[x_frc(N, c, j, k) + b];
                acc += d[SD_EN + c] * (dF[a1] * lam[a2] - dF[a2] * lam[a1]);
            }
            v += (i < 3 ? cfg.dT : -cfg.dT) * acc;
        }
        return v;
    }
    int u = i - NS;
    if (u < 6) return d[SD_VM + u / 3] != 0.0 ? w.dx[x_of_u(N, k, u)] : dw * w.dx[x_of_u(N, k, u)];
    int f = u - 6, c = f / 12, j = (f % 12) / 3, a = f % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    double a4 = d[SD_EN + c] / NJ, sum = 0, own = w.dx[x_frc(N, c, j, k) + a];
    for (int j2 = 0; j2 < NJ; ++j2) sum += w.dx[x_frc(N, c, j2, k) + a];
    // symmetry block: 2 w_s (I + (4 a4^2 - 2 a4) 1 1')
    double v = 2.0 * cfg.w_sym * (own + (NJ * a4 * a4 - 2.0 * a4) * sum) + dw * own;
    if (k > 0) v += 2.0 * cfg.w_rate[a] * (own - w.dx[x_frc(N, c, j, k - 1) + a]);
    if (k + 1 < N) v += 2.0 * cfg.w_rate[a] * (own - w.dx[x_frc(N, c, j, k + 1) + a]);
    // bilinear block: dT en [lam_h]x (dcom - dpos_c)
    const double* lam = w.y + g_h(N, k);
    double e1 = w.dx[x_com(N, k) + a1] - w.dx[x_pos(N, c, k) + a1], e2 = w.dx[x_com(N, k) + a2] - w.dx[x_pos(N, c, k) + a2];
    // ([lam]x e)_a = lam[a1] e[a2] - lam[a2] e[a1]
    v += cfg.dT * d[SD_EN + c] * (lam[a1] * e2 - lam[a2] * e1);
    return v;
}

// residual of the linearised stationarity  rho = grad f + W dx + J'(y + dy)  -> w.xt (x order); returns max |rho|
template <class Cta>
CMPC_FN double lin_residual(Cta& cta, const Config& cfg, const Instance& in, const Work& w, double dw)
{
    const int N = cfg.N, m = dim_g(N);
    for (int r = cta.tid; r < m; r += cta.nt) w.yn[r] = w.y[r] + w.dy[r];
    cta.sync();
    double mx = 0;
    for (int it = cta.tid; it < (N + 1) * (NS + NU); it += cta.nt) {
        int k = it / (NS + NU), i = it % (NS + NU);
        if (k == N && i >= NS) continue;
        double v = lag_grad_entry(cfg, in, w, w.yn, k, i) + hess_dx_entry(cfg, w, dw, k, i);
        if (i >= NS && i < NS + 6 && w.sd[k * SD_STRIDE + SD_VM + (i - NS) / 3] != 0.0) v = 0.0;  // variable held fixed
        w.xt[i < NS ? x_of_s(N, k, i) : x_of_u(N, k, i - NS)] = v;
        mx = fmax(mx, fabs(v));
    }
    return cta.max(mx);
}

// backward vector sweep of the refinement: cost-to-go gradient for the right hand side rho (in w.xt) with the stored factors
template <class Cta>
CMPC_FN void refine_backward(Cta& cta, const Config& cfg, const Work& w, Smem& sm)
{
    const int N = cfg.N;
    for (int i = cta.tid; i < NXI; i += cta.nt) sm.pv[i] = i < NS ? w.xt[x_of_s(N, N, i)] : 0.0;
    cta.sync();
    for (int k = N - 1; k >= 0; --k) {
        const double* d = w.sd + k * SD_STRIDE;
        double* ric = w.ric + (size_t)k * RIC_STRIDE;
        for (int it = cta.tid; it < NU * NU; it += cta.nt) sm.HL[(it / NU) * LDL + it % NU] = ric[RIC_L + it];
        for (int it = cta.tid; it < NU * LDY; it += cta.nt) sm.GY[it] = ric[RIC_Y + it];
        for (int it = cta.tid; it < NU; it += cta.nt) sm.dinv[it] = ric[RIC_DINV + it];
        for (int u = cta.tid; u < NU; u += cta.nt)
            sm.hu[u] = w.xt[x_of_u(N, k, u)] + BbarT(cfg, d, u, [&](int i) { return sm.pv[i]; });
        cta.sync();
        if (cta.warp == 0) {  // z = L^-1 h_u
            for (int j = 0; j < NU; ++j) {
                double zj = sm.hu[j] * sm.dinv[j];
                cta.syncwarp();
                for (int i = j + 1 + cta.lane; i < NU; i += cta.wsize) sm.hu[i] -= sm.HL[i * LDL + j] * zj;
                if (cta.lane == 0) sm.tv[j] = zj;
                cta.syncwarp();
            }
        }
        cta.sync();
        for (int u = cta.tid; u < NU; u += cta.nt) ric[RIC_Z + u] = sm.tv[u];
        for (int i = cta.tid; i < NXI; i += cta.nt) {
            double v = i < NS ? w.xt[x_of_s(N, k, i)] + AT15(cfg, d, i, [&](int r) { return sm.pv[r]; }) : 0.0;
            for (int u = 0; u < NU; ++u) v -= sm.GY[u * LDY + i] * sm.tv[u];
            sm.wv[i] = v;
        }
        cta.sync();
        for (int i = cta.tid; i < NXI; i += cta.nt) sm.pv[i] = sm.wv[i];
        cta.sync();
    }
}

// multipliers of the initial-condition / dynamics rows from the stationarity of the linearised Lagrangian in s_k:
//   lambda+_k = A_k' lambda+_{k+1} - [ grad f + (W dx) + sum_box a_i (y_i + dy_i) ]_{s_k},   k = N .. 0
// (adjoint recursion on the step just computed: exact for the given dx, independent of the accuracy of the cost-to-go).
// Needs dx and the dy of the path rows; leaves dy = lambda+ - y for those rows.
template <class Cta>
CMPC_FN void costate_backward(Cta& cta, const Config& cfg, const Instance& in, const Work& w, double dw)
{
    const int N = cfg.N;
    for (int k = N; k >= 0; --k) {
        const double* d = w.sd + (k < N ? k : 0) * SD_STRIDE;
        for (int i = cta.tid; i < NS; i += cta.nt) {
            double v = w.grad[x_of_s(N, k, i)] + (cost_diag_s(cfg, k, i) + dw) * w.dx[x_of_s(N, k, i)];
            if (k > 0 && i >= 9) {
                int c = (i - 9) / 3, a = (i - 9) % 3;
                const double* R = in.p + p_rot(N, c, k - 1);
                for (int q = 0; q < 3; ++q) {
                    int row = g_box(N, c, k - 1) + q;
                    v += R[3 * q + a] * (w.y[row] + w.dy[row]);
                }
            }
            if (k < N && (i < 3 || i >= 9)) {  // (S_k' du_k): bilinear hessian block between forces and com / pos
                const double* lam = w.y + g_h(N, k);
                int a = i < 3 ? i : (i - 9) % 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                double acc = 0;
                for (int c = (i < 3 ? 0 : (i - 9) / 3); c < (i < 3 ? NC : (i - 9) / 3 + 1); ++c) {
                    double dF[3] = {0, 0, 0};
                    for (int j = 0; j < NJ; ++j)
                        for (int b = 0; b < 3; ++b) dF[b] += w.dx[x_frc(N, c, j, k) + b];
                    acc += d[SD_EN + c] * (dF[a1] * lam[a2] - dF[a2] * lam[a1]);
                }
                v += (i < 3 ? cfg.dT : -cfg.dT) * acc;
            }
            double lamp = -v;
            if (k < N) lamp += AT_y(cfg, d, w.dy, k, i);  // w.dy holds lambda+_{k+1} (absolute) at this point
            w.dy[g_of_s(N, k, i)] = lamp;
        }
        cta.sync();
    }
    for (int it = cta.tid; it < (N + 1) * NS; it += cta.nt) {
        int row = g_of_s(N, it / NS, it % NS);
        w.dy[row] -= w.y[row];
    }
    cta.sync();
}

// ------------------------------------------------------------------------------------------------ the solver
// x_io: in = initial guess, out = solution (CasADi order).  lam_io: multipliers of g (out; in when warm_duals).
// linear-algebra policy of the generic (multi-warp CTA) sweeps above; cmpc_warp.cuh has the warp-per-instance one
struct LinCta {
    Smem& sm;
    template <class Cta> CMPC_HD int backward(Cta& cta, const Config& cfg, const Instance& in, const Work& w, double dw)
    { return riccati_backward(cta, cfg, in, w, sm, dw); }
    template <class Cta> CMPC_HD void forward(Cta& cta, const Config& cfg, const Instance& in, const Work& w, bool refine)
    { riccati_forward(cta, cfg, in, w, sm, refine); }
    template <class Cta> CMPC_HD void refine_back(Cta& cta, const Config& cfg, const Work& w)
    { refine_backward(cta, cfg, w, sm); }
};

template <class Cta, class Lin>
CMPC_HD Result ipm_solve(Cta& cta, const Config& cfg, const Instance& in, const Work& w, Lin& lin, double* x_io,
                         double* lam_io, int warm_duals)
{
    const int N = cfg.N, n = dim_x(N), m = dim_g(N), npr = N * INEQ_PER_KNOT;
    Result res;
    res.status = 1; res.iters = 0; res.obj = 0; res.kkt = 0;

    // ---- rows: classify, relax bounds (bound_relax_factor), validate
    double bad = 0;
    for (int it = cta.tid; it < (N + 1) * NS; it += cta.nt) {
        int row = g_of_s(N, it / NS, it % NS);
        if (!(in.lbg[row] == in.ubg[row]) || !(fabs(in.lbg[row]) < cfg.inf_bound)) bad = 1;
    }
    for (int pr = cta.tid; pr < npr; pr += cta.nt) {
        int row = g_of_ineq(N, pr / INEQ_PER_KNOT, pr % INEQ_PER_KNOT);
        double lb = in.lbg[row], ub = in.ubg[row];
        bool hl = finite_lo(cfg, lb), hu = finite_up(cfg, ub);
        if (!(lb == lb) || !(ub == ub) || (hl && hu && lb > ub)) bad = 1;
        if (hl && hu && lb == ub) { w.sL[pr] = lb; w.sU[pr] = lb; }
        else {
            w.sL[pr] = hl ? lb - cfg.bound_relax * fmax(1.0, fabs(lb)) : -HUGE_VAL;
            w.sU[pr] = hu ? ub + cfg.bound_relax * fmax(1.0, fabs(ub)) : HUGE_VAL;
        }
    }
    for (int i = cta.tid; i < n; i += cta.nt) { double v = x_io[i]; w.x[i] = v; if (!(fabs(v) < HUGE_VAL)) bad = 1; }
    bad = cta.max(bad);
    if (bad != 0.0) { res.status = 4; return res; }

    // ---- initial point
    stage_data(cta, cfg, in, w.x, w.sd);
    eval_g(cta, cfg, in, w.x, w.sd, w.g);
    double f = eval_f(cta, cfg, in, w.x, w.grad);
    for (int r = cta.tid; r < m; r += cta.nt) w.y[r] = (warm_duals && lam_io) ? lam_io[r] : 0.0;
    cta.sync();
    for (int pr = cta.tid; pr < npr; pr += cta.nt) {
        int row = g_of_ineq(N, pr / INEQ_PER_KNOT, pr % INEQ_PER_KNOT);
        double sl = w.sL[pr], su = w.sU[pr];
        bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
        w.zL[pr] = 0; w.zU[pr] = 0; w.s[pr] = 0;
        if (!hl && !hu) { w.y[row] = 0; continue; }
        if (sl == su) continue;
        double s = w.g[row], k1 = cfg.bound_push;
        if (hl && hu) {
            double pl = fmin(k1 * fmax(1.0, fabs(sl)), k1 * (su - sl));
            double pu = fmin(k1 * fmax(1.0, fabs(su)), k1 * (su - sl));
            s = fmin(fmax(s, sl + pl), su - pu);
        } else if (hl) s = fmax(s, sl + k1 * fmax(1.0, fabs(sl)));
        else s = fmin(s, su - k1 * fmax(1.0, fabs(su)));
        w.s[pr] = s;
        if (hl) w.zL[pr] = 1.0;
        if (hu) w.zU[pr] = 1.0;
        if (warm_duals) {
            double yv = w.y[row];
            if (hl) w.zL[pr] = fmax(yv < 0 ? -yv : 0.0, cfg.mu_init / (s - sl));
            if (hu) w.zU[pr] = fmax(yv > 0 ? yv : 0.0, cfg.mu_init / (su - s));
        }
    }
    cta.sync();

    double mu = cfg.mu_init, tau = fmax(TAU_MIN, 1.0 - mu);
    const double mu_min = fmin(cfg.tol, 1e-4) / (KAPPA_EPS + 1.0);  // IPOPT: min(tol, compl_inf_tol) / (barrier_tol_factor + 1)
    double theta0, phi0;
    theta_phi(cta, cfg, in, w, w.g, w.s, f, mu, theta0, phi0);
    const double theta_max = 1e4 * fmax(1.0, theta0), theta_min = 1e-4 * fmax(1.0, theta0);
    double filt_t[MAX_FILTER], filt_p[MAX_FILTER];
    int nfilt = 0;
    double dw_last = 0.0;
    Errs e0;
    int it = 0, status = 1;

    for (it = 0; it <= cfg.max_iter; ++it) {
        e0 = kkt_error(cta, cfg, in, w, 0.0);
#ifdef CMPC_HOST_TRACE
        if (getenv("CMPC_TRACE")) fprintf(stderr, "it %3d f %.10e E0 %.2e (d %.2e v %.2e c %.2e) mu %.1e dw %.1e nfilt %d\n", it, f, e0.E, e0.dual, e0.viol, e0.compl_, mu, dw_last, nfilt);
#endif
        if (e0.E <= cfg.tol && e0.dual <= 1.0 && e0.viol <= 1e-4 && e0.compl_ <= 1e-4) { status = 0; break; }
        if (it == cfg.max_iter) { status = 1; break; }
        // barrier update (eq. 7), filter reset
        for (;;) {
            Errs em = kkt_error(cta, cfg, in, w, mu);
            if (em.E <= KAPPA_EPS * mu && mu > mu_min) {
                CMPC_STAT(4);
                mu = fmax(mu_min, fmin(KAPPA_MU * mu, mu * sqrt(mu)));
                tau = fmax(TAU_MIN, 1.0 - mu);
                nfilt = 0;
            } else break;
        }
        CMPC_STAT(3);
        // ---- search direction with inertia correction (alg. IC): Cholesky failure inside the Riccati sweep <=> wrong inertia
        const double dc = fmax(DC_BAR * sqrt(sqrt(mu)), DC_FLOOR);
        double dw = 0.0;
        int rc = 1, tries = 0;
        for (;;) {
            for (int pr = cta.tid; pr < npr; pr += cta.nt) {
                int row = g_of_ineq(N, pr / INEQ_PER_KNOT, pr % INEQ_PER_KNOT);
                double sl = w.sL[pr], su = w.sU[pr];
                bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
                double sg = 0, t = 0;
                if (hl || hu) {
                    if (sl == su) { sg = 1.0 / dc; t = w.y[row] + (w.g[row] - sl) / dc; }
                    else {
                        double s = w.s[pr];
                        sg = dw;
                        if (hl) { double dd = s - sl; sg += w.zL[pr] / dd; t -= mu / dd; }
                        if (hu) { double dd = su - s; sg += w.zU[pr] / dd; t += mu / dd; }
                        if (hl && !hu) t += KAPPA_D * mu;
                        if (hu && !hl) t -= KAPPA_D * mu;
                        t += sg * (w.g[row] - s);
                    }
                }
                w.sig[pr] = sg; w.tt[pr] = t;
            }
            cta.sync();
            CMPC_STAT(0);
            rc = lin.backward(cta, cfg, in, w, dw);
            if (rc == 0) break;
            if (dw == 0.0) dw = dw_last == 0.0 ? DW_FIRST : fmax(DW_MIN, KW_MINUS * dw_last);
            else dw *= (dw_last == 0.0 ? KW_PLUS_FIRST : KW_PLUS);
            if (dw > DW_MAX || ++tries > 60) break;
        }
        if (rc != 0) { status = 3; break; }
        if (dw > 0.0) dw_last = dw;
        lin.forward(cta, cfg, in, w, false);
        recover_path(cta, cfg, in, w, mu, dc);
        costate_backward(cta, cfg, in, w, dw);
        // ---- iterative refinement on the stationarity residual of the Newton system (the eliminated rows hold exactly)
        double rho_prev = HUGE_VAL;
        for (int rf = 0; rf < MAX_REFINE; ++rf) {
            double rho = lin_residual(cta, cfg, in, w, dw);
            if (!(rho > REFINE_TOL) || rho > 0.5 * rho_prev) break;
            rho_prev = rho;
            CMPC_STAT(1);
            lin.refine_back(cta, cfg, w);
            lin.forward(cta, cfg, in, w, true);
            recover_path(cta, cfg, in, w, mu, dc);
            costate_backward(cta, cfg, in, w, dw);
        }
        // ---- fraction to the boundary (eq. 15) and directional derivative of the barrier function
        double vmin[2] = {1.0, 1.0};  // alpha_max (primal), alpha_z
        double dphi = 0.0, nanflag = 0.0;
        for (int pr = cta.tid; pr < npr; pr += cta.nt) {
            double sl = w.sL[pr], su = w.sU[pr];
            bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
            if ((!hl && !hu) || sl == su) continue;
            double s = w.s[pr], ds = w.ds[pr];
            if (hl) {
                double dd = s - sl, dz = w.dzL[pr];
                dphi -= mu * ds / dd;
                if (ds < 0) vmin[0] = fmin(vmin[0], -tau * dd / ds);
                if (dz < 0) vmin[1] = fmin(vmin[1], -tau * w.zL[pr] / dz);
            }
            if (hu) {
                double dd = su - s, dz = w.dzU[pr];
                dphi += mu * ds / dd;
                if (ds > 0) vmin[0] = fmin(vmin[0], tau * dd / ds);
                if (dz < 0) vmin[1] = fmin(vmin[1], -tau * w.zU[pr] / dz);
            }
            if (hl && !hu) dphi += KAPPA_D * mu * ds;
            if (hu && !hl) dphi -= KAPPA_D * mu * ds;
        }
        for (int i = cta.tid; i < n; i += cta.nt) {
            double d = w.dx[i];
            dphi += w.grad[i] * d;
            if (!(fabs(d) < HUGE_VAL)) nanflag = 1.0;
        }
        cta.template minv<2>(vmin);
        dphi = cta.sum(dphi);
        nanflag = cta.max(nanflag);
        if (nanflag != 0.0) { status = 3; break; }
        const double amax = vmin[0], az = vmin[1];
        // ---- filter line search (alg. A)
        double theta, phi;
        theta_phi(cta, cfg, in, w, w.g, w.s, f, mu, theta, phi);
        double amin;
        if (dphi < 0) {
            amin = fmin(GAMMA_THETA, GAMMA_PHI * theta / (-dphi));
            if (theta <= theta_min) amin = fmin(amin, DELTA_SW * pow(theta, S_THETA) / pow(-dphi, S_PHI));
        } else amin = GAMMA_THETA;
        amin *= GAMMA_ALPHA;
        double alpha = amax, ft = f;
        int accepted = 0, armijo = 0;
        while (alpha >= amin || alpha == amax) {
            for (int i = cta.tid; i < n; i += cta.nt) w.xt[i] = w.x[i] + alpha * w.dx[i];
            for (int pr = cta.tid; pr < npr; pr += cta.nt) w.st[pr] = w.s[pr] + alpha * w.ds[pr];
            cta.sync();
            CMPC_STAT(2);
            stage_data(cta, cfg, in, w.xt, w.sd);
            eval_g(cta, cfg, in, w.xt, w.sd, w.g);
            ft = eval_f(cta, cfg, in, w.xt, (double*)nullptr);
            double th_t, ph_t;
            theta_phi(cta, cfg, in, w, w.g, w.st, ft, mu, th_t, ph_t);
            bool ok = (fabs(ph_t) < HUGE_VAL) && (fabs(th_t) < HUGE_VAL) && th_t <= theta_max;
            for (int q = 0; ok && q < nfilt; ++q)
                if (th_t >= filt_t[q] && ph_t >= filt_p[q]) ok = false;
            if (ok) {
                bool sw = dphi < 0 && theta <= theta_min && alpha * pow(-dphi, S_PHI) > DELTA_SW * pow(theta, S_THETA);
                double slack = 10.0 * 2.2e-16 * fabs(phi);
                if (sw) {
                    if (ph_t - phi - slack <= ETA_PHI * alpha * dphi) { accepted = 1; armijo = 1; }
                } else if (th_t <= (1.0 - GAMMA_THETA) * theta || ph_t - slack <= phi - GAMMA_PHI * theta) {
                    accepted = 1; armijo = 0;
                }
            }
            if (accepted) break;
            alpha *= 0.5;
            if (alpha < 1e-16) break;
        }
        if (!accepted) { status = 2; break; }  // IPOPT would start its restoration phase here (not restated)
        if (!armijo && nfilt < MAX_FILTER) {
            filt_t[nfilt] = (1.0 - GAMMA_THETA) * theta;
            filt_p[nfilt] = phi - GAMMA_PHI * theta;
            nfilt++;
        }
        // ---- accept the trial point (w.g and w.sd already hold its values)
        for (int i = cta.tid; i < n; i += cta.nt) w.x[i] = w.xt[i];
        for (int pr = cta.tid; pr < npr; pr += cta.nt) {
            double sl = w.sL[pr], su = w.sU[pr];
            bool hl = sl > -HUGE_VAL, hu = su < HUGE_VAL;
            if ((!hl && !hu) || sl == su) continue;
            double s = w.st[pr];
            w.s[pr] = s;
            if (hl) {
                double z = w.zL[pr] + az * w.dzL[pr], dd = s - sl;
                w.zL[pr] = fmax(fmin(z, KAPPA_SIGMA * mu / dd), mu / (KAPPA_SIGMA * dd));
            }
            if (hu) {
                double z = w.zU[pr] + az * w.dzU[pr], dd = su - s;
                w.zU[pr] = fmax(fmin(z, KAPPA_SIGMA * mu / dd), mu / (KAPPA_SIGMA * dd));
            }
        }
        for (int r = cta.tid; r < m; r += cta.nt) w.y[r] += alpha * w.dy[r];
        cta.sync();
        f = eval_f(cta, cfg, in, w.x, w.grad);
    }
    cta.sync();
    for (int i = cta.tid; i < n; i += cta.nt) x_io[i] = w.x[i];
    if (lam_io)
        for (int r = cta.tid; r < m; r += cta.nt) lam_io[r] = w.y[r];
    res.status = status; res.iters = it; res.obj = f; res.kkt = e0.E;
    return res;
}

}  // namespace cmpc
