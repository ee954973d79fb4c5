// cmpc_sparse.cuh -- entry-wise jacobian / hessian of the NLP in the reference's CasADi CSC order.
//
// Parity surface for nlp_jac_fg (tmp.c:71962, pattern casadi_s5 tmp.c:67) and nlp_hess_l (tmp.c:58926, pattern casadi_s4
// tmp.c:66, full symmetric).  The solver itself never materialises these arrays (it works on stage blocks,
// cmpc_core.cuh); they exist so that the hand-derived derivatives can be compared entry for entry with the reference's
// generated code.  Every structural nonzero has a fixed "emission index" e; jac_entry / hess_entry return its
// (row, col, value).  The host sorts the (row, col) pairs once per horizon into CSC and uploads emission -> nz slot.
#pragma once

#include "cmpc_core.cuh"

namespace cmpc {

constexpr int JAC_PER_KNOT = 243;
constexpr int HESS_PER_KNOT = 348;  // interior knot; knot N-1 has 300, knot N has 12

CMPC_HD void skew_rc(int q, int& r, int& c)
{
    // structural entries of a 3x3 skew matrix in row-major order
    const int rr[6] = {0, 0, 1, 1, 2, 2}, cc[6] = {1, 2, 0, 2, 0, 1};
    r = rr[q]; c = cc[q];
}

// e in [0, 15 + 243 N).  x, p may be null (pattern only; value is then meaningless).
CMPC_HD double jac_entry(const Config& cfg, const double* x, const double* p, int e, int& row, int& col)
{
    const int N = cfg.N;
    const bool val = x != nullptr;
    if (e < NS) { row = e; col = x_of_s(N, 0, e); return 1.0; }
    e -= NS;
    const int k = e / JAC_PER_KNOT;
    e %= JAC_PER_KNOT;
    if (e < 9) {
        int a = e / 3, t = e % 3;
        row = g_com(N, k) + a;
        col = t == 0 ? x_com(N, k + 1) + a : t == 1 ? x_com(N, k) + a : x_dcom(N, k) + a;
        return t == 0 ? 1.0 : t == 1 ? -1.0 : -cfg.dT;
    }
    if (e < 15) { int a = (e - 9) / 2, t = (e - 9) % 2; row = g_dcom(N, k) + a; col = (t == 0 ? x_dcom(N, k + 1) : x_dcom(N, k)) + a; return t == 0 ? 1.0 : -1.0; }
    if (e < 21) { int a = (e - 15) / 2, t = (e - 15) % 2; row = g_h(N, k) + a; col = (t == 0 ? x_h(N, k + 1) : x_h(N, k)) + a; return t == 0 ? 1.0 : -1.0; }
    if (e < 45) {
        int idx = e - 21, cj = idx / 3, a = idx % 3, c = cj / 4, j = cj % 4;
        row = g_dcom(N, k) + a; col = x_frc(N, c, j, k) + a;
        return val ? -cfg.dT * p[p_en(N, c, k)] : 0.0;
    }
    if (e < 93) {  // h rows <- forces: -dT en [rho_cj]x
        int idx = e - 45, cj = idx / 6, q = idx % 6, c = cj / 4, j = cj % 4, r, cc;
        skew_rc(q, r, cc);
        row = g_h(N, k) + r; col = x_frc(N, c, j, k) + cc;
        if (!val) return 0.0;
        const double* R = p + p_rot(N, c, k);
        const double* cr = cfg.corner[c][j];
        double rho[3];
        for (int a = 0; a < 3; ++a)
            rho[a] = R[a] * cr[0] + R[3 + a] * cr[1] + R[6 + a] * cr[2] + x[x_pos(N, c, k) + a] - x[x_com(N, k) + a];
        return -cfg.dT * p[p_en(N, c, k)] * skew(rho, r, cc);
    }
    if (e < 105) {  // h rows <- pos_c: +dT en [F_c]x
        int idx = e - 93, c = idx / 6, q = idx % 6, r, cc;
        skew_rc(q, r, cc);
        row = g_h(N, k) + r; col = x_pos(N, c, k) + cc;
        if (!val) return 0.0;
        double F[3] = {0, 0, 0};
        for (int j = 0; j < NJ; ++j)
            for (int a = 0; a < 3; ++a) F[a] += x[x_frc(N, c, j, k) + a];
        return cfg.dT * p[p_en(N, c, k)] * skew(F, r, cc);
    }
    if (e < 111) {  // h rows <- com: -dT [sum_c en_c F_c]x
        int q = e - 105, r, cc;
        skew_rc(q, r, cc);
        row = g_h(N, k) + r; col = x_com(N, k) + cc;
        if (!val) return 0.0;
        double F[3] = {0, 0, 0};
        for (int c = 0; c < NC; ++c)
            for (int j = 0; j < NJ; ++j)
                for (int a = 0; a < 3; ++a) F[a] += p[p_en(N, c, k)] * x[x_frc(N, c, j, k) + a];
        return -cfg.dT * skew(F, r, cc);
    }
    if (e < 129) {
        int idx = e - 111, c = idx / 9, a = (idx % 9) / 3, t = idx % 3;
        row = g_pos(N, c, k) + a;
        col = (t == 0 ? x_pos(N, c, k + 1) : t == 1 ? x_pos(N, c, k) : x_vel(N, c, k)) + a;
        if (t == 0) return 1.0;
        if (t == 1) return -1.0;
        return val ? -(1.0 - p[p_en(N, c, k)]) * cfg.dT : 0.0;
    }
    if (e < 147) {
        int idx = e - 129, c = idx / 9, r = (idx % 9) / 3, a = idx % 3;
        row = g_box(N, c, k) + r; col = x_pos(N, c, k + 1) + a;
        return val ? p[p_rot(N, c, k) + 3 * r + a] : 0.0;
    }
    {
        int idx = e - 147, c = idx / 48, j = (idx % 48) / 12, r = (idx % 12) / 3, a = idx % 3;
        row = g_fric(N, c, j, k) + r; col = x_frc(N, c, j, k) + a;
        return val ? fric_coef(cfg, p + p_rot(N, c, k), r, a) : 0.0;
    }
}

CMPC_HD int hess_emissions(int N) { return (N - 1) * HESS_PER_KNOT + 300 + 12; }

// e in [0, 348 N - 36).  p, lam_g may be null (pattern only).
CMPC_HD double hess_entry(const Config& cfg, const double* p, double lam_f, const double* lam_g, int e, int& row, int& col)
{
    const int N = cfg.N;
    const bool val = p != nullptr;
    int k = e / HESS_PER_KNOT;
    if (k >= N - 1) {  // knots N-1 (300 entries) and N (12 entries)
        int r = e - (N - 1) * HESS_PER_KNOT;
        if (r < 300) { k = N - 1; e = r; } else { k = N; e = r - 300; }
    } else e %= HESS_PER_KNOT;
    if (e < 12) {
        int i = e < 3 ? e : (e < 6 ? e + 3 : e + 3);  // stage-state index: com 0..2, h 6..8, pos 9..14
        row = col = x_of_s(N, k, i);
        return lam_f * cost_diag_s(cfg, k, i);
    }
    if (e < 108) {
        int idx = e - 12, cja = idx / 4, t = idx % 4, c = cja / 12, j = (cja % 12) / 3, a = cja % 3;
        row = x_frc(N, c, j, k) + a;
        double a4 = val ? p[p_en(N, c, k)] / NJ : 0.0;
        if (t == 0) {
            col = row;
            int nrate = N >= 2 ? ((k == 0 || k == N - 1) ? 1 : 2) : 0;
            return lam_f * (2.0 * cfg.w_sym * (1.0 - 2.0 * a4 + NJ * a4 * a4) + 2.0 * cfg.w_rate[a] * nrate);
        }
        col = x_frc(N, c, (j + t) % NJ, k) + a;
        return lam_f * 2.0 * cfg.w_sym * (NJ * a4 * a4 - 2.0 * a4);
    }
    if (e < 300) {  // bilinear blocks of lam_h' g_h
        int idx = e - 108, cj = idx / 24, blk = (idx % 24) / 6, q = idx % 6, c = cj / 4, j = cj % 4, r, cc;
        skew_rc(q, r, cc);
        int xf = x_frc(N, c, j, k), xp = x_pos(N, c, k), xc = x_com(N, k);
        double sgn;
        if (blk == 0) { row = xp + r; col = xf + cc; sgn = 1.0; }
        else if (blk == 1) { row = xf + r; col = xp + cc; sgn = -1.0; }
        else if (blk == 2) { row = xc + r; col = xf + cc; sgn = -1.0; }
        else { row = xf + r; col = xc + cc; sgn = 1.0; }
        if (!val) return 0.0;
        return sgn * cfg.dT * p[p_en(N, c, k)] * skew(lam_g + g_h(N, k), r, cc);
    }
    {
        int idx = e - 300, cja = idx / 2, t = idx % 2, c = cja / 12, j = (cja % 12) / 3, a = cja % 3;
        int x0 = x_frc(N, c, j, k) + a, x1 = x_frc(N, c, j, k + 1) + a;
        row = t == 0 ? x0 : x1; col = t == 0 ? x1 : x0;
        return -lam_f * 2.0 * cfg.w_rate[a];
    }
}

}  // namespace cmpc
