// cmpc_core.cuh -- definitions shared by every kernel of the centroidal-MPC solve: the solver configuration, the stage data
// of a knot, IPOPT's constants, and the CTA-wide evaluation of f / grad f / g (parity surface kernels, cmpc_kernels.cu).
//
// What the library replaces (reference = /root/reference):
//   * the CasADi-generated NLP functions nlp_fg / nlp_jac_fg / nlp_hess_l
//     (src/centroidal-mpc-walking/config/robots/ergoCubGazeboV1/tmp.c:12430, :71962, :58926) -> hand-derived stage blocks;
//   * the IPOPT (+MUMPS/MA97) solve that BLF CentroidalMPC::advance() triggers (call site
//     src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:615) -> primal-dual interior point whose KKT system is
//     factorised by a Riccati recursion over the knots (augmented state: 15 physical + 24 previous forces):
//     cmpc_ipm.cuh (interior point, stage major) and cmpc_warp.cuh (sweeps on 3 x 3 tiles).
// The first-generation generic formulation of the solve (one thread, dense blocks) lives under tests/hostsim as an
// independent second implementation; it is not part of the product.
#pragma once

#include <math.h>

#include "cmpc_layout.cuh"

// large routines are real functions on the device (they are called from several places of the solver)
#if defined(__CUDACC__)
#define CMPC_FN __host__ __device__ __noinline__
#else
#define CMPC_FN inline
#endif

namespace cmpc {
#ifdef CMPC_HOST_STATS
static long g_stat[8];  // backward sweeps, refinements, line-search trials, iterations, mu updates
#define CMPC_STAT(i) (++g_stat[i])
#else
#define CMPC_STAT(i) ((void)0)
#endif

// ------------------------------------------------------------------------------------------------ data
struct Config {
    int N;
    double dT;
    double w_com[3], w_h, w_pos, w_sym, w_rate[3];
    double corner[NC][NJ][3];
    double fricA[NF][3];
    // solver options (IPOPT names in comments)
    double tol;          // tol
    int max_iter;        // max_iter
    double mu_init;      // mu_init
    double bound_relax;  // bound_relax_factor
    double bound_push;   // bound_push = bound_frac
    double inf_bound;    // nlp_upper_bound_inf
    int pc;              // 1: Mehrotra predictor-corrector barrier update, 0: IPOPT's monotone update (mu_strategy)
    double mu_warm;      // barrier parameter of a solve that starts from given multipliers (warm_duals): floor of the bound
                         // multipliers mu_warm / slack and first mu of the monotone update
    double scal_max_grad;  // nlp_scaling_max_gradient (IPOPT default 100; 0: no scaling of the objective)
    double acc_tol;        // acceptable_tol (IPOPT default 1e-6; 0: no acceptable-level termination)
    int acc_iter;          // acceptable_iter (IPOPT default 15)
    double box_lo[NC][3], box_up[NC][3];   // CONTACT_c bounding_box_lower_limit / upper_limit (step adjustment, contact frame)
};

struct Instance {  // CasADi order, read only
    const double* p;
    const double* lbg;
    const double* ubg;
};

constexpr int SD_STRIDE = 40;  // en[2] Fc[6] Fall[3] rho[24] vmask[2] ...
constexpr int SD_EN = 0, SD_FC = 2, SD_FALL = 8, SD_RHO = 11, SD_VM = 35;
// IPOPT constants (Waechter & Biegler 2006; same values as oracle/cmpc_oracle_ipm.c)
constexpr double KAPPA_EPS = 10.0, KAPPA_MU = 0.2, THETA_MU = 1.5, TAU_MIN = 0.99, S_MAX = 100.0;
constexpr double KAPPA_SIGMA = 1e10, KAPPA_D = 1e-5, GAMMA_THETA = 1e-5, GAMMA_PHI = 1e-8, DELTA_SW = 1.0;
constexpr double S_THETA = 1.1, S_PHI = 2.3, ETA_PHI = 1e-8, GAMMA_ALPHA = 0.05;
constexpr double DW_FIRST = 1e-4, DW_MIN = 1e-20, DW_MAX = 1e40, KW_PLUS_FIRST = 100.0, KW_PLUS = 8.0, KW_MINUS = 1.0 / 3.0;
constexpr double DC_BAR = 1e-8, DC_FLOOR = 0.0;
constexpr int MAX_FILTER = 16;
constexpr int MAX_REFINE = 2;          // iterative refinement steps of the Newton system (IPOPT: max_refinement_steps 10)
#ifndef CMPC_REFINE_TOL
#define CMPC_REFINE_TOL 1e-9   // a factor 10 below the default KKT tolerance (1e-10: same iterations, 4 % slower)
#endif
constexpr double REFINE_TOL = CMPC_REFINE_TOL;   // absolute stationarity residual of the linear system that triggers a step
constexpr double GRAV_Z = -9.80665;  // tmp.c:3916

// ------------------------------------------------------------------------------------------------ small helpers
CMPC_HD void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
// entry (r, c) of the skew matrix [a]x
CMPC_HD double skew(const double* a, int r, int c)
{
    if (r == c) return 0.0;
    int k = 3 - r - c;
    double v = a[k];
    // (0,1) -a2  (0,2) +a1  (1,0) +a2  (1,2) -a0  (2,0) -a1  (2,1) +a0
    return ((c - r + 3) % 3 == 1) ? -v : v;
}
CMPC_HD double com_z_omega(double w, int k) { return (w - 0.5 * w) * exp(-(double)k) + 0.5 * w; }  // tmp.c:407-484
// symmetric 3x3 packed (xx xy xz yy yz zz)
CMPC_HD int sym3(int a, int b)
{
    if (a > b) { int t = a; a = b; b = t; }
    return a == 0 ? b : (a == 1 ? 2 + b : 5);
}
CMPC_HD bool finite_lo(const Config& c, double v) { return v > -c.inf_bound; }
CMPC_HD bool finite_up(const Config& c, double v) { return v < c.inf_bound; }

// (A R^T)[r][a] for contact c at knot k: friction row r of a corner, world-frame force component a
CMPC_HD double fric_coef(const Config& cfg, const double* R, int r, int a)
{
    return cfg.fricA[r][0] * R[a] + cfg.fricA[r][1] * R[3 + a] + cfg.fricA[r][2] * R[6 + a];
}
// coefficient of path row l of knot k on its 3 variables (box: pos_{c,k+1}; friction: f_{cjk})
CMPC_HD double path_coef(const Config& cfg, const Instance& in, int k, int l, int a)
{
    const int N = cfg.N;
    if (l < 6) return in.p[p_rot(N, l / 3, k) + 3 * (l % 3) + a];  // column (l%3) of R, component a
    int f = l - 6;
    return fric_coef(cfg, in.p + p_rot(N, f / 16, k), f % 4, a);
}
CMPC_HD int path_var(const Config& cfg, int k, int l)  // x index of the first of the 3 variables of the row
{
    const int N = cfg.N;
    if (l < 6) return x_pos(N, l / 3, k + 1);
    int f = l - 6;
    return x_frc(N, f / 16, (f % 16) / 4, k);
}

// ------------------------------------------------------------------------------------------------ evaluation
// stage data of knot k from the point xv: en, F_c, F_all, rho_cj, vel mask
template <class Cta>
CMPC_FN void stage_data(Cta& cta, const Config& cfg, const Instance& in, const double* xv, double* sd)
{
    const int N = cfg.N;
    for (int it = cta.tid; it < N * 10; it += cta.nt) {
        int k = it / 10, w = it % 10;
        double* d = sd + k * SD_STRIDE;
        if (w < 8) {  // rho of corner (c, j)
            int c = w / 4, j = w % 4;
            const double* R = in.p + p_rot(N, c, k);
            const double* r = cfg.corner[c][j];
            for (int a = 0; a < 3; ++a)
                d[SD_RHO + 3 * w + a] = R[a] * r[0] + R[3 + a] * r[1] + R[6 + a] * r[2] + xv[x_pos(N, c, k) + a]
                                        - xv[x_com(N, k) + a];
        } else {
            int c = w - 8;
            double en = in.p[p_en(N, c, k)];
            d[SD_EN + c] = en;
            d[SD_VM + c] = ((1.0 - en) * cfg.dT == 0.0) ? 1.0 : 0.0;
            for (int a = 0; a < 3; ++a) {
                double s = 0;
                for (int j = 0; j < NJ; ++j) s += xv[x_frc(N, c, j, k) + a];
                d[SD_FC + 3 * c + a] = s;
            }
        }
    }
    cta.sync();
    for (int it = cta.tid; it < N * 3; it += cta.nt) {
        int k = it / 3, a = it % 3;
        double* d = sd + k * SD_STRIDE;
        d[SD_FALL + a] = d[SD_EN] * d[SD_FC + a] + d[SD_EN + 1] * d[SD_FC + 3 + a];
    }
    cta.sync();
}

// constraint row values g(xv) (all m rows); needs stage_data(xv) in sd
template <class Cta>
CMPC_FN void eval_g(Cta& cta, const Config& cfg, const Instance& in, const double* xv, const double* sd, double* g)
{
    const int N = cfg.N;
    for (int i = cta.tid; i < NS; i += cta.nt) g[i] = xv[x_of_s(N, 0, i)];
    for (int it = cta.tid; it < N * ROWS_PER_KNOT; it += cta.nt) {
        int k = it / ROWS_PER_KNOT, l = it % ROWS_PER_KNOT;
        const double* d = sd + k * SD_STRIDE;
        if (l < NS) {
            int row = g_of_s(N, k + 1, l);
            double v = xv[x_of_s(N, k + 1, l)] - xv[x_of_s(N, k, l)];
            if (l < 3) v -= cfg.dT * xv[x_dcom(N, k) + l];
            else if (l < 6) {
                int a = l - 3;
                v -= cfg.dT * ((a == 2 ? GRAV_Z : 0.0) + in.p[p_extf(N, k) + a] + d[SD_FALL + a]);
            } else if (l < 9) {
                int a = l - 6, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                double t = in.p[p_extt(N, k) + a];
                for (int c = 0; c < NC; ++c) {
                    double en = d[SD_EN + c], acc = 0;
                    for (int j = 0; j < NJ; ++j) {
                        const double* rho = d + SD_RHO + 3 * (4 * c + j);
                        const double* f = xv + x_frc(N, c, j, k);
                        acc += rho[a1] * f[a2] - rho[a2] * f[a1];
                    }
                    t += en * acc;
                }
                v -= cfg.dT * t;
            } else {
                int c = (l - 9) / 3, a = (l - 9) % 3;
                v -= (1.0 - d[SD_EN + c]) * cfg.dT * xv[x_vel(N, c, k) + a];
            }
            g[row] = v;
        } else {
            int li = l - NS;
            int row = g_of_ineq(N, k, li), xv0 = path_var(cfg, k, li);
            double v = 0;
            for (int a = 0; a < 3; ++a) {
                double xa = xv[xv0 + a];
                if (li < 6) xa -= in.p[p_nom(N, li / 3, k + 1) + a];
                v += path_coef(cfg, in, k, li, a) * xa;
            }
            g[row] = v;
        }
    }
    cta.sync();
}

// objective f(xv) (all-reduced) and optionally its gradient
template <class Cta>
CMPC_FN double eval_f(Cta& cta, const Config& cfg, const Instance& in, const double* xv, double* grad)
{
    const int N = cfg.N;
    double acc = 0;
    // state terms: item = (k, i) i in 0..14 (dcom has no cost)
    for (int it = cta.tid; it < (N + 1) * NS; it += cta.nt) {
        int k = it / NS, i = it % NS, xi = x_of_s(N, k, i);
        double gr = 0;
        if (i < 3) {
            double e = xv[xi] - in.p[p_comref(N, k) + i];
            if (i < 2) { acc += cfg.w_com[i] * e * e; gr = 2.0 * cfg.w_com[i] * e; }
            else { double om = com_z_omega(cfg.w_com[2], k); acc += (om * e) * (om * e); gr = 2.0 * om * om * e; }
        } else if (i >= 6 && i < 9) {
            double e = xv[xi] - in.p[p_href(N, k) + i - 6];
            acc += cfg.w_h * e * e; gr = 2.0 * cfg.w_h * e;
        } else if (i >= 9) {
            int c = (i - 9) / 3, a = (i - 9) % 3;
            double e = xv[xi] - in.p[p_nom(N, c, k) + a];
            acc += cfg.w_pos * e * e; gr = 2.0 * cfg.w_pos * e;
        }
        if (grad) grad[xi] = gr;
    }
    // force terms: item = (k, c, a)
    for (int it = cta.tid; it < N * 6; it += cta.nt) {
        int k = it / 6, c = (it % 6) / 3, a = it % 3;
        double en = in.p[p_en(N, c, k)];
        double fj[NJ], sum = 0;
        for (int j = 0; j < NJ; ++j) { fj[j] = xv[x_frc(N, c, j, k) + a]; sum += fj[j]; }
        double mean = en / NJ * sum;
        for (int j = 0; j < NJ; ++j) {
            double d = fj[j] - mean;
            acc += cfg.w_sym * d * d;
            double gr = 2.0 * cfg.w_sym * (d - (en / NJ) * (sum - NJ * mean));
            if (k + 1 < N) {
                double dn = xv[x_frc(N, c, j, k + 1) + a] - fj[j];
                acc += cfg.w_rate[a] * dn * dn;
                gr -= 2.0 * cfg.w_rate[a] * dn;
            }
            if (k > 0) gr += 2.0 * cfg.w_rate[a] * (fj[j] - xv[x_frc(N, c, j, k - 1) + a]);
            if (grad) grad[x_frc(N, c, j, k) + a] = gr;
        }
        if (grad) grad[x_vel(N, c, k) + a] = 0.0;
    }
    double f = cta.sum(acc);
    return f;
}

// ------------------------------------------------------------------------------------------------ stage blocks
CMPC_HD double cost_diag_s(const Config& cfg, int k, int i)
{
    if (i < 2) return 2.0 * cfg.w_com[i];
    if (i == 2) { double om = com_z_omega(cfg.w_com[2], k); return 2.0 * om * om; }
    if (i < 6) return 0.0;
    if (i < 9) return 2.0 * cfg.w_h;
    return 2.0 * cfg.w_pos;
}

}  // namespace cmpc
