/*
 * cmpc_oracle_ipm.c -- CPU ORACLE (test infrastructure only, see cmpc_oracle.h).
 *
 * Restatement of the solve behind BLF CentroidalMPC::advance() (call site
 * /root/reference/src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:615), which the reference delegates to
 * CasADi nlpsol -> IPOPT 3.13.4 (+MUMPS / MA97), an UN-VENDORED dependency (dockerfiles/Dockerfile:49,
 * config/robots/<robot>/centroidal_mpc.ini:1).  IPOPT's published algorithm (Waechter & Biegler, "On the implementation of
 * an interior-point filter line-search algorithm for large-scale nonlinear programming", Math. Prog. 106(1), 2006;
 * section / equation numbers below refer to that paper) is restated with IPOPT's default option values:
 *
 *   * slack form  c(x) = 0, d(x) - s = 0, s_L <= s <= s_U ; rows with lbg == ubg are equalities           (sec. 3.4)
 *   * bound_relax_factor 1e-8, bound_push = bound_frac = 0.01, bound_mult_init_val 1                        (sec. 3.5, 3.6)
 *   * mu_init 0.1, monotone Fiacco-McCormick update kappa_mu 0.2, theta_mu 1.5, kappa_eps 10               (eq. 7)
 *   * tau = max(0.99, 1 - mu) fraction to the boundary for s and for z                                       (eq. 8, 15)
 *   * primal-dual Newton step from the condensed augmented system                                             (eq. 13)
 *   * inertia correction delta_w: 1e-4 first, x8 / x100 growth, /3 decay, delta_c = 1e-8 mu^0.25             (alg. IC)
 *     - without an LDL' the inertia is not available: the inertia-free curvature test of Chiang & Zavala
 *       (IPOPT option neg_curv_test_tol) triggers the correction instead
 *   * filter line search with switching condition / Armijo, gamma_theta 1e-5, gamma_phi 1e-8, delta 1,
 *     s_theta 1.1, s_phi 2.3, eta_phi 1e-8, filter reset at every mu update                                  (alg. A, sec. 2.3)
 *   * kappa_sigma 1e10 safeguard of the bound multipliers                                                     (eq. 16)
 *   * kappa_d 1e-5 damping of one-sided slacks                                                                 (sec. 3.7)
 *   * termination on the scaled optimality error E_0 <= tol with s_max 100                                    (eq. 5, 6)
 *   * nlp_scaling_method gradient-based with nlp_scaling_max_gradient 100 (IPOPT's default): objective and constraint rows
 *     scaled from the derivatives at the initial point; tol applies to the scaled problem, dual_inf_tol 1 /
 *     constr_viol_tol 1e-4 / compl_inf_tol 1e-4 to the unscaled one (IpIpoptCalculatedQuantities / OptimalityErrorConvergenceCheck)
 *   * acceptable-point termination: acceptable_tol 1e-6 for acceptable_iter 15 consecutive iterations, or an acceptable
 *     current point when the line search / the regularisation gives up ("Solved To Acceptable Level", status 5)
 * NOT restated: restoration phase, second-order correction, watchdog (documented in DESIGN.md).  The linear algebra is deliberately generic (banded LU with partial pivoting of
 * the stage-ordered condensed KKT matrix) so that it is independent of the Riccati recursion of the CUDA path.
 *
 * SOLVER-LEVEL PARITY UNPINNED (no IPOPT/CasADi in this container, no reference goldens); see cmpc_oracle.h.
 */
#define _GNU_SOURCE
#include "cmpc_oracle.h"

#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ options */
void cmpc_oracle_ipm_default_opts(cmpc_oracle_ipm_opts* o)
{
    o->tol = 1e-8;
    o->max_iter = 200;
    o->mu_init = 0.1;
    o->bound_relax = 1e-8;
    o->bound_push = 0.01;
    o->inf_bound = 1e19;
    o->warm_duals = 0;
    o->verbose = 0;
    o->mehrotra = 0;
    o->nlp_scaling_max_gradient = 100.0;
    o->acceptable_tol = 1e-6;
    o->acceptable_iter = 15;
}

/* IPOPT constants (defaults of 3.13/3.14) */
#define KAPPA_EPS 10.0
#define KAPPA_MU 0.2
#define THETA_MU 1.5
#define TAU_MIN 0.99
#define S_MAX 100.0
#define KAPPA_SIGMA 1e10
#define KAPPA_D 1e-5
#define GAMMA_THETA 1e-5
#define GAMMA_PHI 1e-8
#define DELTA_SW 1.0
#define S_THETA 1.1
#define S_PHI 2.3
#define ETA_PHI 1e-8
#define GAMMA_ALPHA 0.05
#define DW_FIRST 1e-4
#define DW_MIN 1e-20
#define DW_MAX 1e40
#define KW_PLUS_FIRST 100.0
#define KW_PLUS 8.0
#define KW_MINUS (1.0 / 3.0)
#define DC_BAR 1e-8
#define KAPPA_C 0.25
#define NEG_CURV_TOL 1e-10
#define MAX_FILTER 64
#define PC_MAX_ITER 50 /* predictor-corrector iterations after which an instance is handed to the monotone path */

/* ------------------------------------------------------------------ banded LU (LAPACK dgbtf2 / dgbtrs layout) */
typedef struct band {
    int n, kl, ku, ld;
    double* ab;
    int* ipiv;
} band;

static void band_alloc(band* B, int n, int kl, int ku)
{
    B->n = n; B->kl = kl; B->ku = ku; B->ld = 2 * kl + ku + 1;
    B->ab = malloc(sizeof(double) * (size_t)B->ld * n);
    B->ipiv = malloc(sizeof(int) * n);
}
static void band_free(band* B) { free(B->ab); free(B->ipiv); }
static inline void band_zero(band* B) { memset(B->ab, 0, sizeof(double) * (size_t)B->ld * B->n); }
static inline void band_add(band* B, int i, int j, double v) { B->ab[(size_t)j * B->ld + B->kl + B->ku + i - j] += v; }

/* returns 0 on success, j+1 if pivot j is exactly zero (or below tiny) */
static int band_factor(band* B)
{
    const int n = B->n, kl = B->kl, ku = B->ku, ld = B->ld, kv = kl + ku;
    double* ab = B->ab;
    int ju = 0, info = 0;
    for (int j = 0; j < n; ++j) {
        int km = kl < n - 1 - j ? kl : n - 1 - j;
        double* col = ab + (size_t)j * ld;
        int jp = 0;
        double best = fabs(col[kv]);
        for (int i = 1; i <= km; ++i)
            if (fabs(col[kv + i]) > best) { best = fabs(col[kv + i]); jp = i; }
        B->ipiv[j] = jp + j;
        if (best > 1e-300) {
            int t = j + ku + jp; if (t > n - 1) t = n - 1; if (t > ju) ju = t;
            if (jp != 0)
                for (int jj = j; jj <= ju; ++jj) {
                    double* c2 = ab + (size_t)jj * ld;
                    double tmp = c2[kv + jp - (jj - j)];
                    c2[kv + jp - (jj - j)] = c2[kv - (jj - j)];
                    c2[kv - (jj - j)] = tmp;
                }
            double inv = 1.0 / col[kv];
            for (int i = 1; i <= km; ++i) col[kv + i] *= inv;
            for (int jj = j + 1; jj <= ju; ++jj) {
                double* c2 = ab + (size_t)jj * ld;
                double u = c2[kv - (jj - j)];
                if (u != 0.0)
                    for (int i = 1; i <= km; ++i) c2[kv + i - (jj - j)] -= col[kv + i] * u;
            }
        } else if (!info)
            info = j + 1;
    }
    return info;
}

static void band_solve(const band* B, double* b)
{
    const int n = B->n, kl = B->kl, ku = B->ku, ld = B->ld, kv = kl + ku;
    const double* ab = B->ab;
    for (int j = 0; j < n; ++j) {
        int km = kl < n - 1 - j ? kl : n - 1 - j;
        int p = B->ipiv[j];
        if (p != j) { double t = b[j]; b[j] = b[p]; b[p] = t; }
        const double* col = ab + (size_t)j * ld;
        double bj = b[j];
        if (bj != 0.0)
            for (int i = 1; i <= km; ++i) b[j + i] -= bj * col[kv + i];
    }
    for (int j = n - 1; j >= 0; --j) {
        const double* col = ab + (size_t)j * ld;
        b[j] /= col[kv];
        double bj = b[j];
        int lo = j - kv; if (lo < 0) lo = 0;
        for (int i = lo; i < j; ++i) b[i] -= bj * col[kv - (j - i)];
    }
}

/* ------------------------------------------------------------------ stage keys (for the banded ordering) */
/* variable / row -> ordering key in units of "thirds of a stage"; layouts: SURVEY.md 8(a) a-1, a-4 */
static int x_key(int N, int i)
{
    int nb = 9 * (N + 1);
    if (i < nb) return 3 * ((i % (3 * (N + 1))) / 3);              /* com, dcom, h of knot k -> state k */
    int r = (i - nb) % (18 * N + 3);
    if (r < 3 * (N + 1)) return 3 * (r / 3);                         /* pos_k  -> state k   */
    r -= 3 * (N + 1);
    return 3 * ((r % (3 * N)) / 3) + 1;                              /* vel_k, force_jk -> control k */
}
static int g_key(int N, int r)
{
    if (r < 15) return -1;                                           /* initial conditions, before s_0 */
    r -= 15;
    if (r < 15 * N) return 3 * ((r % (3 * N)) / 3) + 2;              /* dynamics k: after u_k          */
    r -= 15 * N;
    r %= 19 * N;
    if (r < 3 * N) return 3 * (r / 3) + 2;                           /* box row of knot k (on s_{k+1})  */
    r -= 3 * N;
    return 3 * (r / 16) + 1;                                         /* friction rows of knot k         */
}
static int is_box_row(int N, int r)
{
    if (r < 15 + 15 * N) return 0;
    r = (r - 15 - 15 * N) % (19 * N);
    return r < 3 * N;
}

/* ------------------------------------------------------------------ solver workspace */
typedef struct ipm_ws {
    int n, m, N;
    const cmpc_oracle_nlp_fn* nlp;
    /* CSR view of J: for each row, list of (col, nz index) */
    int* rptr; int* rcol; int* rnz;
    /* row classes */
    int* rtype;     /* 0 eq, 1 ineq, 2 free */
    int* rslot;     /* eq: index in KKT unknowns (n + e); ineq: slack index */
    int mE, mI;
    double* dflag;  /* per eq slot: 1 if the row gets the dual regularisation delta_c */
    double* target; /* per eq slot: right hand side b */
    int* irow;      /* per slack: g row */
    double *sL, *sU; int *hasL, *hasU;
    /* iterate */
    double *x, *s, *y, *zL, *zU;            /* y: all m rows */
    double *dx, *ds, *dy, *dzL, *dzU;
    double *grad, *g, *jnz, *hnz;
    double *xt, *st, *gt;                   /* trial */
    double *rhs, *sigma, *rs, *rd;
    double *muL, *muU;                      /* complementarity target of every bound (mu, or mu -+ the Mehrotra corrector) */
    /* gradient-based NLP scaling (IPOPT nlp_scaling_method): f_s = df f, g_s = Dc g */
    double df; double* dcs; double* ytmp;
    /* KKT ordering */
    int nK; int* perm; /* unknown u (0..nK-1: x then eq slots) -> position */
    band K;
} ipm_ws;

static int cmp_keyidx(const void* a, const void* b)
{
    const int* A = a; const int* B = b;
    if (A[0] != B[0]) return A[0] - B[0];
    return A[1] - B[1];
}

static double vmaxabs(const double* v, int n) { double m = 0; for (int i = 0; i < n; ++i) if (fabs(v[i]) > m) m = fabs(v[i]); return m; }

/* r_x = grad + J' y (all rows) */
static void lag_grad(const ipm_ws* W, const double* y, double* out)
{
    const cmpc_oracle_nlp_fn* F = W->nlp;
    for (int c = 0; c < W->n; ++c) {
        double a = W->grad[c];
        for (int q = F->jc[c]; q < F->jc[c + 1]; ++q) a += W->jnz[q] * y[F->jr[q]];
        out[c] = a;
    }
}

typedef struct errs { double dual, viol, compl_, E; double dual_u, viol_u, compl_u; /* unscaled */ } errs;

/* the NLP functions of the SCALED problem */
static void s_jac_fg(const ipm_ws* W, const double* x, const double* p, double* f, double* grad, double* g, double* jnz)
{
    const cmpc_oracle_nlp_fn* F = W->nlp;
    F->jac_fg(F->ctx, x, p, f, grad, g, jnz);
    *f *= W->df;
    for (int c = 0; c < W->n; ++c) grad[c] *= W->df;
    for (int r = 0; r < W->m; ++r) g[r] *= W->dcs[r];
    for (int q = 0; q < F->nnz_j; ++q) jnz[q] *= W->dcs[F->jr[q]];
}
static void s_fg(const ipm_ws* W, const double* x, const double* p, double* f, double* g)
{
    const cmpc_oracle_nlp_fn* F = W->nlp;
    F->fg(F->ctx, x, p, f, g);
    *f *= W->df;
    for (int r = 0; r < W->m; ++r) g[r] *= W->dcs[r];
}
static void s_hess(const ipm_ws* W, const double* x, const double* p, const double* y, double* hnz)
{
    const cmpc_oracle_nlp_fn* F = W->nlp;
    for (int r = 0; r < W->m; ++r) W->ytmp[r] = W->dcs[r] * y[r];
    F->hess(F->ctx, x, p, W->df, W->ytmp, hnz);
}

static errs kkt_error(const ipm_ws* W, double mu, double* scratch)
{
    errs e = {0, 0, 0, 0, 0, 0, 0};
    lag_grad(W, W->y, scratch);
    e.dual = vmaxabs(scratch, W->n);
    double sumy = 0, sumz = 0; int nb = 0;
    for (int r = 0; r < W->m; ++r) {
        if (W->rtype[r] == 2) continue;
        sumy += fabs(W->y[r]);
        if (W->rtype[r] == 0) {
            double c = W->g[r] - W->target[W->rslot[r]];
            if (fabs(c) > e.viol) e.viol = fabs(c);
            if (fabs(c) / W->dcs[r] > e.viol_u) e.viol_u = fabs(c) / W->dcs[r];
        } else {
            int i = W->rslot[r];
            double d = W->g[r] - W->s[i];
            if (fabs(d) > e.viol) e.viol = fabs(d);
            if (fabs(d) / W->dcs[r] > e.viol_u) e.viol_u = fabs(d) / W->dcs[r];
            double ds_ = -W->y[r];
            if (W->hasL[i]) { ds_ -= W->zL[i]; sumz += W->zL[i]; nb++; double c = (W->s[i] - W->sL[i]) * W->zL[i] - mu; if (fabs(c) > e.compl_) e.compl_ = fabs(c); }
            if (W->hasU[i]) { ds_ += W->zU[i]; sumz += W->zU[i]; nb++; double c = (W->sU[i] - W->s[i]) * W->zU[i] - mu; if (fabs(c) > e.compl_) e.compl_ = fabs(c); }
            if (fabs(ds_) > e.dual) e.dual = fabs(ds_);
        }
    }
    int mact = W->mE + W->mI;
    double sd = fmax(S_MAX, (sumy + sumz) / fmax(1, mact + nb)) / S_MAX;
    double sc = fmax(S_MAX, sumz / fmax(1, nb)) / S_MAX;
    e.E = fmax(e.dual / sd, fmax(e.viol, e.compl_ / sc));
    e.dual_u = e.dual / W->df; e.compl_u = e.compl_ / W->df;
    return e;
}

static double barrier_obj(const ipm_ws* W, double f, const double* s, double mu)
{
    double phi = f;
    for (int i = 0; i < W->mI; ++i) {
        if (W->hasL[i]) phi -= mu * log(s[i] - W->sL[i]);
        if (W->hasU[i]) phi -= mu * log(W->sU[i] - s[i]);
        if (W->hasL[i] && !W->hasU[i]) phi += KAPPA_D * mu * (s[i] - W->sL[i]);
        if (W->hasU[i] && !W->hasL[i]) phi += KAPPA_D * mu * (W->sU[i] - s[i]);
    }
    return phi;
}

static double infeas_l1(const ipm_ws* W, const double* g, const double* s)
{
    double th = 0;
    for (int r = 0; r < W->m; ++r) {
        if (W->rtype[r] == 0) th += fabs(g[r] - W->target[W->rslot[r]]);
        else if (W->rtype[r] == 1) th += fabs(g[r] - s[W->rslot[r]]);
    }
    return th;
}

/* assemble and solve the condensed system; returns 0 ok, 1 singular, 2 wrong curvature */
static int solve_kkt(ipm_ws* W, double mu, double dw, double dc, double lam_f)
{
    const cmpc_oracle_nlp_fn* F = W->nlp;
    const int n = W->n;
    band* K = &W->K;
    band_zero(K);
    double* rhs = W->rhs;
    /* r_x */
    lag_grad(W, W->y, rhs);
    for (int c = 0; c < n; ++c) rhs[c] = -rhs[c];
    /* W + dw I */
    int* used = calloc(n, sizeof(int));
    for (int c = 0; c < n; ++c) {
        for (int q = F->hc[c]; q < F->hc[c + 1]; ++q) {
            band_add(K, W->perm[F->hr[q]], W->perm[c], W->hnz[q]);
            if (W->hnz[q] != 0.0) used[c] = 1;
        }
        band_add(K, W->perm[c], W->perm[c], dw);
    }
    /* constraint rows */
    for (int r = 0; r < W->m; ++r) {
        if (W->rtype[r] == 2) continue;
        if (W->rtype[r] == 0) {
            int e = W->rslot[r], pe = W->perm[n + e];
            for (int q = W->rptr[r]; q < W->rptr[r + 1]; ++q) {
                double v = W->jnz[W->rnz[q]];
                if (v != 0.0) used[W->rcol[q]] = 1;
                band_add(K, pe, W->perm[W->rcol[q]], v);
                band_add(K, W->perm[W->rcol[q]], pe, v);
            }
            band_add(K, pe, pe, -dc * W->dflag[e]);
            rhs[n + e] = -(W->g[r] - W->target[e]);
        } else {
            int i = W->rslot[r];
            double sg = dw, rs = -W->y[r];
            if (W->hasL[i]) { double d = W->s[i] - W->sL[i]; sg += W->zL[i] / d; rs -= W->muL[i] / d; }
            if (W->hasU[i]) { double d = W->sU[i] - W->s[i]; sg += W->zU[i] / d; rs += W->muU[i] / d; }
            if (W->hasL[i] && !W->hasU[i]) rs += KAPPA_D * mu;
            if (W->hasU[i] && !W->hasL[i]) rs -= KAPPA_D * mu;
            double rd = W->g[r] - W->s[i];
            W->sigma[i] = sg; W->rs[i] = rs; W->rd[i] = rd;
            double t = sg * rd + rs;
            for (int q = W->rptr[r]; q < W->rptr[r + 1]; ++q) {
                double vq = W->jnz[W->rnz[q]];
                if (vq == 0.0) continue;
                used[W->rcol[q]] = 1;
                rhs[W->rcol[q]] -= vq * t;
                for (int q2 = W->rptr[r]; q2 < W->rptr[r + 1]; ++q2)
                    band_add(K, W->perm[W->rcol[q]], W->perm[W->rcol[q2]], sg * vq * W->jnz[W->rnz[q2]]);
            }
        }
    }
    /* variables that appear nowhere (contact velocities of a foot in stance: coefficient (1-en) dT = 0 and no cost) are
     * held fixed instead of being left to the inertia correction */
    for (int c = 0; c < n; ++c)
        if (!used[c]) { band_add(K, W->perm[c], W->perm[c], 1.0); rhs[c] = 0.0; }
    free(used);
    (void)lam_f;
    /* permute rhs */
    double* b = malloc(sizeof(double) * W->nK);
    for (int u = 0; u < W->nK; ++u) b[W->perm[u]] = rhs[u];
    if (band_factor(K)) { free(b); return 1; }
    band_solve(K, b);
    for (int u = 0; u < W->nK; ++u) rhs[u] = b[W->perm[u]];
    free(b);
    for (int u = 0; u < W->nK; ++u) if (!isfinite(rhs[u])) return 1;
    memcpy(W->dx, rhs, sizeof(double) * n);
    /* recover the rest */
    double curv = 0.0, dd = 0.0, yc = 0.0;
    for (int r = 0; r < W->m; ++r) {
        if (W->rtype[r] == 2) { W->dy[r] = 0; continue; }
        if (W->rtype[r] == 0) {
            W->dy[r] = rhs[n + W->rslot[r]];
            yc += (W->y[r] + W->dy[r]) * (W->g[r] - W->target[W->rslot[r]]);
        } else {
            int i = W->rslot[r];
            double jd = 0;
            for (int q = W->rptr[r]; q < W->rptr[r + 1]; ++q) jd += W->jnz[W->rnz[q]] * W->dx[W->rcol[q]];
            double ds = jd + W->rd[i];
            W->ds[i] = ds;
            W->dy[r] = W->sigma[i] * ds + W->rs[i];
            W->dzL[i] = W->dzU[i] = 0;
            if (W->hasL[i]) { double d = W->s[i] - W->sL[i]; W->dzL[i] = W->muL[i] / d - W->zL[i] - W->zL[i] / d * ds; }
            if (W->hasU[i]) { double d = W->sU[i] - W->s[i]; W->dzU[i] = W->muU[i] / d - W->zU[i] + W->zU[i] / d * ds; }
            curv += W->sigma[i] * ds * ds; dd += ds * ds;
            yc += (W->y[r] + W->dy[r]) * W->rd[i];
        }
    }
    for (int c = 0; c < n; ++c) {
        double a = dw * W->dx[c];
        for (int q = F->hc[c]; q < F->hc[c + 1]; ++q) a += W->hnz[q] * W->dx[F->hr[q]];
        curv += a * W->dx[c];
        dd += W->dx[c] * W->dx[c];
    }
    if (curv + fmax(-yc, 0.0) < NEG_CURV_TOL * dd) return 2;
    return 0;
}

static void ws_free(ipm_ws* W)
{
    free(W->rptr); free(W->rcol); free(W->rnz); free(W->rtype); free(W->rslot); free(W->dflag); free(W->target);
    free(W->irow); free(W->sL); free(W->sU); free(W->hasL); free(W->hasU);
    free(W->x); free(W->s); free(W->y); free(W->zL); free(W->zU);
    free(W->dx); free(W->ds); free(W->dy); free(W->dzL); free(W->dzU);
    free(W->grad); free(W->g); free(W->jnz); free(W->hnz); free(W->xt); free(W->st); free(W->gt);
    free(W->rhs); free(W->sigma); free(W->rs); free(W->rd); free(W->perm); free(W->muL); free(W->muU);
    free(W->dcs); free(W->ytmp);
    band_free(&W->K);
}

#define ALLOCD(k) calloc((size_t)((k) > 0 ? (k) : 1), sizeof(double))
#define ALLOCI(k) calloc((size_t)((k) > 0 ? (k) : 1), sizeof(int))

static int ipm_core(const cmpc_oracle_nlp_fn* F, int N, const cmpc_oracle_ipm_opts* opts, int mehrotra, const double* p,
                    const double* lbg, const double* ubg, double* x, double* lam_g, cmpc_oracle_ipm_stats* stats)
{
    ipm_ws Wk; ipm_ws* W = &Wk; memset(W, 0, sizeof *W);
    const int n = F->n, m = F->m;
    W->n = n; W->m = m; W->N = N; W->nlp = F;
    memset(stats, 0, sizeof *stats);

    /* CSR of J pattern */
    W->rptr = ALLOCI(m + 1); W->rcol = ALLOCI(F->nnz_j); W->rnz = ALLOCI(F->nnz_j);
    for (int q = 0; q < F->nnz_j; ++q) W->rptr[F->jr[q] + 1]++;
    for (int r = 0; r < m; ++r) W->rptr[r + 1] += W->rptr[r];
    {
        int* fill = ALLOCI(m);
        for (int c = 0; c < n; ++c)
            for (int q = F->jc[c]; q < F->jc[c + 1]; ++q) {
                int r = F->jr[q], pos = W->rptr[r] + fill[r]++;
                W->rcol[pos] = c; W->rnz[pos] = q;
            }
        free(fill);
    }
    /* gradient-based scaling from the derivatives at the initial point (IPOPT GradientScaling::DetermineScalingParametersImpl) */
    W->df = 1.0; W->dcs = ALLOCD(m); W->ytmp = ALLOCD(m);
    W->x = ALLOCD(n); W->grad = ALLOCD(n); W->g = ALLOCD(m); W->jnz = ALLOCD(F->nnz_j);
    for (int r = 0; r < m; ++r) W->dcs[r] = 1.0;
    memcpy(W->x, x, sizeof(double) * n);
    stats->obj_scaling = stats->min_g_scaling = 1.0;
    if (opts->nlp_scaling_max_gradient > 0.0) {
        const double gmax = opts->nlp_scaling_max_gradient, smin = 1e-8;
        double f0;
        F->jac_fg(F->ctx, W->x, p, &f0, W->grad, W->g, W->jnz);
        double a = vmaxabs(W->grad, n);
        if (a > gmax) W->df = fmax(gmax / a, smin);
        double* rmax = ALLOCD(m);
        for (int q = 0; q < F->nnz_j; ++q) if (fabs(W->jnz[q]) > rmax[F->jr[q]]) rmax[F->jr[q]] = fabs(W->jnz[q]);
        for (int r = 0; r < m; ++r) {
            if (rmax[r] > gmax) W->dcs[r] = fmax(gmax / rmax[r], smin);
            if (W->dcs[r] < stats->min_g_scaling) stats->min_g_scaling = W->dcs[r];
        }
        free(rmax);
        stats->obj_scaling = W->df;
    }
    double* lbs = ALLOCD(m); double* ubs = ALLOCD(m);
    for (int r = 0; r < m; ++r) {
        lbs[r] = lbg[r] > -opts->inf_bound ? lbg[r] * W->dcs[r] : lbg[r];
        ubs[r] = ubg[r] < opts->inf_bound ? ubg[r] * W->dcs[r] : ubg[r];
    }
    /* classify rows (from here on lbg / ubg are the bounds of the scaled rows) */
    lbg = lbs; ubg = ubs;
    W->rtype = ALLOCI(m); W->rslot = ALLOCI(m);
    int mE = 0, mI = 0, bad = 0;
    for (int r = 0; r < m; ++r) {
        int hl = lbg[r] > -opts->inf_bound, hu = ubg[r] < opts->inf_bound;
        if (hl && hu && lbg[r] == ubg[r]) { W->rtype[r] = 0; W->rslot[r] = mE++; }
        else if (hl || hu) { W->rtype[r] = 1; W->rslot[r] = mI++; if (hl && hu && lbg[r] > ubg[r]) bad = 1; }
        else W->rtype[r] = 2;
        if (!(lbg[r] == lbg[r]) || !(ubg[r] == ubg[r])) bad = 1;
    }
    W->mE = mE; W->mI = mI;
    W->dflag = ALLOCD(mE); W->target = ALLOCD(mE); W->irow = ALLOCI(mI);
    W->sL = ALLOCD(mI); W->sU = ALLOCD(mI); W->hasL = ALLOCI(mI); W->hasU = ALLOCI(mI);
    for (int r = 0; r < m; ++r) {
        if (W->rtype[r] == 0) { W->target[W->rslot[r]] = lbg[r]; W->dflag[W->rslot[r]] = is_box_row(N, r) ? 1.0 : 0.0; }
        else if (W->rtype[r] == 1) {
            int i = W->rslot[r];
            W->irow[i] = r;
            W->hasL[i] = lbg[r] > -opts->inf_bound; W->hasU[i] = ubg[r] < opts->inf_bound;
            W->sL[i] = W->hasL[i] ? lbg[r] - opts->bound_relax * fmax(1.0, fabs(lbg[r])) : -INFINITY;
            W->sU[i] = W->hasU[i] ? ubg[r] + opts->bound_relax * fmax(1.0, fabs(ubg[r])) : INFINITY;
        }
    }
    W->s = ALLOCD(mI); W->y = ALLOCD(m); W->zL = ALLOCD(mI); W->zU = ALLOCD(mI);
    W->dx = ALLOCD(n); W->ds = ALLOCD(mI); W->dy = ALLOCD(m); W->dzL = ALLOCD(mI); W->dzU = ALLOCD(mI);
    W->hnz = ALLOCD(F->nnz_h);
    W->xt = ALLOCD(n); W->st = ALLOCD(mI); W->gt = ALLOCD(m);
    W->nK = n + mE;
    W->muL = ALLOCD(mI); W->muU = ALLOCD(mI);
    W->rhs = ALLOCD(W->nK); W->sigma = ALLOCD(mI); W->rs = ALLOCD(mI); W->rd = ALLOCD(mI);
    /* stage ordering + bandwidth */
    W->perm = ALLOCI(W->nK);
    {
        int* ki = malloc(sizeof(int) * 2 * W->nK);
        for (int c = 0; c < n; ++c) { ki[2 * c] = x_key(N, c); ki[2 * c + 1] = c; }
        for (int r = 0; r < m; ++r)
            if (W->rtype[r] == 0) { int u = n + W->rslot[r]; ki[2 * u] = g_key(N, r); ki[2 * u + 1] = u; }
        qsort(ki, W->nK, 2 * sizeof(int), cmp_keyidx);
        for (int pos = 0; pos < W->nK; ++pos) W->perm[ki[2 * pos + 1]] = pos;
        free(ki);
        int bw = 0;
        for (int c = 0; c < n; ++c)
            for (int q = F->hc[c]; q < F->hc[c + 1]; ++q) { int d = abs(W->perm[F->hr[q]] - W->perm[c]); if (d > bw) bw = d; }
        for (int r = 0; r < m; ++r) {
            if (W->rtype[r] == 2) continue;
            for (int q = W->rptr[r]; q < W->rptr[r + 1]; ++q) {
                if (W->rtype[r] == 0) { int d = abs(W->perm[n + W->rslot[r]] - W->perm[W->rcol[q]]); if (d > bw) bw = d; }
                else for (int q2 = W->rptr[r]; q2 < W->rptr[r + 1]; ++q2) { int d = abs(W->perm[W->rcol[q]] - W->perm[W->rcol[q2]]); if (d > bw) bw = d; }
            }
        }
        band_alloc(&W->K, W->nK, bw, bw);
    }
    if (bad) { stats->status = 4; ws_free(W); free(lbs); free(ubs); return 4; }

    /* ---------------- initial point */
    double f;
    s_jac_fg(W, W->x, p, &f, W->grad, W->g, W->jnz);
    for (int i = 0; i < mI; ++i) {
        double s = W->g[W->irow[i]];
        double k1 = opts->bound_push, k2 = opts->bound_push;
        if (W->hasL[i] && W->hasU[i]) {
            double pl = fmin(k1 * fmax(1.0, fabs(W->sL[i])), k2 * (W->sU[i] - W->sL[i]));
            double pu = fmin(k1 * fmax(1.0, fabs(W->sU[i])), k2 * (W->sU[i] - W->sL[i]));
            s = fmin(fmax(s, W->sL[i] + pl), W->sU[i] - pu);
        } else if (W->hasL[i]) s = fmax(s, W->sL[i] + k1 * fmax(1.0, fabs(W->sL[i])));
        else s = fmin(s, W->sU[i] - k1 * fmax(1.0, fabs(W->sU[i])));
        W->s[i] = s;
        W->zL[i] = W->hasL[i] ? 1.0 : 0.0;
        W->zU[i] = W->hasU[i] ? 1.0 : 0.0;
    }
    if (opts->warm_duals && lam_g) {
        for (int r = 0; r < m; ++r) W->y[r] = W->rtype[r] == 2 ? 0.0 : lam_g[r] * W->df / W->dcs[r];   /* lam_g: unscaled problem */
        for (int i = 0; i < mI; ++i) {
            double yv = W->y[W->irow[i]];
            if (W->hasL[i]) W->zL[i] = fmax(yv < 0 ? -yv : 0.0, opts->mu_init / (W->s[i] - W->sL[i]));
            if (W->hasU[i]) W->zU[i] = fmax(yv > 0 ? yv : 0.0, opts->mu_init / (W->sU[i] - W->s[i]));
        }
    }
    double mu = opts->mu_init, tau = fmax(TAU_MIN, 1.0 - mu);
    /* IPOPT: the barrier parameter stops at min(tol, compl_inf_tol) / (barrier_tol_factor + 1), compl_inf_tol = 1e-4 "posed to
     * the scaled problem" (MonotoneMuUpdate::CalcNewMuAndTau: apply_obj_scaling(compl_inf_tol)) */
    const double mu_min = fmin(opts->tol, 1e-4 * W->df) / (KAPPA_EPS + 1.0);
    double theta0 = infeas_l1(W, W->g, W->s);
    const double theta_max = 1e4 * fmax(1.0, theta0), theta_min = 1e-4 * fmax(1.0, theta0);
    double filt_t[MAX_FILTER], filt_p[MAX_FILTER]; int nfilt = 0;
    double dw_last = 0.0;
    int status = 1, it = 0;
    errs e0 = {0, 0, 0, 0, 0, 0, 0};
    double* scratch = ALLOCD(n);
    int nacc = 0;   /* consecutive acceptable iterates */

    for (it = 0; it <= opts->max_iter; ++it) {
        e0 = kkt_error(W, 0.0, scratch);
        if (opts->verbose)
            fprintf(stderr, "it %3d f %.10e  E0 %.2e (d %.2e v %.2e c %.2e) mu %.1e dw %.1e\n", it, f, e0.E, e0.dual,
                    e0.viol, e0.compl_, mu, dw_last);
        if (e0.E <= opts->tol && e0.dual_u <= 1.0 && e0.viol_u <= 1e-4 && e0.compl_u <= 1e-4) { status = 0; break; }
        if (opts->acceptable_tol > 0.0 && e0.E <= opts->acceptable_tol && e0.dual_u <= 1e10 && e0.viol_u <= 1e-2 && e0.compl_u <= 1e-2) {
            if (++nacc >= opts->acceptable_iter) { status = 5; break; }
        } else nacc = 0;
        if (it == opts->max_iter || (mehrotra && it == PC_MAX_ITER)) { status = 1; break; }
        /* barrier parameter update (eq. 7); in predictor-corrector mode mu follows from the affine step below */
        while (!mehrotra) {
            errs em = kkt_error(W, mu, scratch);
            if (em.E <= KAPPA_EPS * mu && mu > mu_min) {
                mu = fmax(mu_min, fmin(KAPPA_MU * mu, pow(mu, THETA_MU)));
                tau = fmax(TAU_MIN, 1.0 - mu);
                nfilt = 0;
            } else break;
        }
        /* search direction with inertia correction (alg. IC) */
        s_hess(W, W->x, p, W->y, W->hnz);
        double dc = DC_BAR * pow(mu, KAPPA_C);
        double dw = 0.0; int rc, tries = 0;
        for (int i = 0; i < mI; ++i) W->muL[i] = W->muU[i] = mehrotra ? 0.0 : mu;   /* predictor: affine-scaling step */
        for (;;) {
            rc = solve_kkt(W, mehrotra ? 0.0 : mu, dw, dc, 1.0);
            if (rc == 0) break;
            if (dw == 0.0) dw = dw_last == 0.0 ? DW_FIRST : fmax(DW_MIN, KW_MINUS * dw_last);
            else dw *= (dw_last == 0.0 ? KW_PLUS_FIRST : KW_PLUS);
            if (dw > DW_MAX || ++tries > 60) break;
        }
        if (rc != 0) { status = 3; break; }
        if (dw > 0.0) { dw_last = dw; stats->n_reg++; }
        if (mehrotra) {
            /* Mehrotra's rule: step lengths to the boundary of the affine step, mu_aff = mean complementarity after it,
             * sigma = (mu_aff / mu_cur)^3, new barrier parameter sigma mu_cur; corrector = second solve with the SAME matrix
             * and the complementarity targets  mu - ds_aff dz_aff  (second-order term of (s + ds)(z + dz) = mu) */
            double ap = 1.0, ad = 1.0, s00 = 0, s10 = 0, s01 = 0, s11 = 0; int nb = 0;
            for (int i = 0; i < mI; ++i) {
                if (W->hasL[i]) {
                    double d = W->s[i] - W->sL[i];
                    if (W->ds[i] < 0) ap = fmin(ap, -d / W->ds[i]);
                    if (W->dzL[i] < 0) ad = fmin(ad, -W->zL[i] / W->dzL[i]);
                    s00 += d * W->zL[i]; s10 += W->ds[i] * W->zL[i]; s01 += d * W->dzL[i]; s11 += W->ds[i] * W->dzL[i]; nb++;
                }
                if (W->hasU[i]) {
                    double d = W->sU[i] - W->s[i];
                    if (W->ds[i] > 0) ap = fmin(ap, d / W->ds[i]);
                    if (W->dzU[i] < 0) ad = fmin(ad, -W->zU[i] / W->dzU[i]);
                    s00 += d * W->zU[i]; s10 -= W->ds[i] * W->zU[i]; s01 += d * W->dzU[i]; s11 -= W->ds[i] * W->dzU[i]; nb++;
                }
            }
            if (nb > 0 && s00 > 0.0) {
                double aff = s00 + ap * s10 + ad * s01 + ap * ad * s11;
                double sg = fmin(1.0, fmax(0.0, aff / s00));
                mu = fmax(mu_min, sg * sg * sg * s00 / nb);
            }
            tau = fmax(TAU_MIN, 1.0 - mu);
            nfilt = 0;
            for (int i = 0; i < mI; ++i) {
                W->muL[i] = mu - W->ds[i] * W->dzL[i];
                W->muU[i] = mu + W->ds[i] * W->dzU[i];
            }
            rc = solve_kkt(W, mu, dw, dc, 1.0);
            if (rc == 1) { status = 3; break; }
        }
        /* fraction to the boundary (eq. 15) */
        double amax = 1.0, az = 1.0;
        for (int i = 0; i < mI; ++i) {
            if (W->hasL[i]) {
                if (W->ds[i] < 0) amax = fmin(amax, -tau * (W->s[i] - W->sL[i]) / W->ds[i]);
                if (W->dzL[i] < 0) az = fmin(az, -tau * W->zL[i] / W->dzL[i]);
            }
            if (W->hasU[i]) {
                if (W->ds[i] > 0) amax = fmin(amax, tau * (W->sU[i] - W->s[i]) / W->ds[i]);
                if (W->dzU[i] < 0) az = fmin(az, -tau * W->zU[i] / W->dzU[i]);
            }
        }
        /* filter line search (alg. A, steps A-5) */
        double theta = infeas_l1(W, W->g, W->s);
        double phi = barrier_obj(W, f, W->s, mu);
        double dphi = 0.0;
        for (int c = 0; c < n; ++c) dphi += W->grad[c] * W->dx[c];
        for (int i = 0; i < mI; ++i) {
            if (W->hasL[i]) dphi -= mu * W->ds[i] / (W->s[i] - W->sL[i]);
            if (W->hasU[i]) dphi += mu * W->ds[i] / (W->sU[i] - W->s[i]);
            if (W->hasL[i] && !W->hasU[i]) dphi += KAPPA_D * mu * W->ds[i];
            if (W->hasU[i] && !W->hasL[i]) dphi -= KAPPA_D * mu * W->ds[i];
        }
        double amin;
        if (dphi < 0) {
            amin = fmin(GAMMA_THETA, GAMMA_PHI * theta / (-dphi));
            if (theta <= theta_min) amin = fmin(amin, DELTA_SW * pow(theta, S_THETA) / pow(-dphi, S_PHI));
        } else amin = GAMMA_THETA;
        amin *= GAMMA_ALPHA;
        double alpha = amax, ft = f; int accepted = 0, armijo_type = 0;
        while (alpha >= amin || alpha == amax) {
            stats->n_ls_trials++;
            for (int c = 0; c < n; ++c) W->xt[c] = W->x[c] + alpha * W->dx[c];
            for (int i = 0; i < mI; ++i) W->st[i] = W->s[i] + alpha * W->ds[i];
            s_fg(W, W->xt, p, &ft, W->gt);
            double th_t = infeas_l1(W, W->gt, W->st);
            double ph_t = barrier_obj(W, ft, W->st, mu);
            int ok = isfinite(ph_t) && isfinite(th_t) && th_t <= theta_max;
            for (int q = 0; ok && q < nfilt; ++q)
                if (th_t >= filt_t[q] && ph_t >= filt_p[q]) ok = 0;
            if (ok) {
                int sw = dphi < 0 && theta <= theta_min
                         && alpha * pow(-dphi, S_PHI) > DELTA_SW * pow(theta, S_THETA);
                /* machine-precision slack on phi as in IPOPT's Compare_le */
                double slack = 10.0 * 2.2e-16 * fabs(phi);
                if (sw) {
                    if (ph_t - phi - slack <= ETA_PHI * alpha * dphi) { accepted = 1; armijo_type = 1; }
                } else if (th_t <= (1.0 - GAMMA_THETA) * theta || ph_t - slack <= phi - GAMMA_PHI * theta) {
                    accepted = 1; armijo_type = 0;
                }
            }
            if (accepted) break;
            alpha *= 0.5;
            if (alpha < 1e-16) break;
        }
        if (!accepted) {
            /* IPOPT would enter the restoration phase here; not restated.  If the point is already feasible to the
             * tolerance and nearly optimal we stop, otherwise the failure is reported. */
            status = 2;
            break;
        }
        if (!armijo_type && nfilt < MAX_FILTER) {
            filt_t[nfilt] = (1.0 - GAMMA_THETA) * theta;
            filt_p[nfilt] = phi - GAMMA_PHI * theta;
            nfilt++;
        }
        /* accept */
        for (int c = 0; c < n; ++c) W->x[c] = W->xt[c];
        for (int i = 0; i < mI; ++i) {
            W->s[i] = W->st[i];
            if (W->hasL[i]) {
                double z = W->zL[i] + az * W->dzL[i], d = W->s[i] - W->sL[i];
                W->zL[i] = fmax(fmin(z, KAPPA_SIGMA * mu / d), mu / (KAPPA_SIGMA * d));
            }
            if (W->hasU[i]) {
                double z = W->zU[i] + az * W->dzU[i], d = W->sU[i] - W->s[i];
                W->zU[i] = fmax(fmin(z, KAPPA_SIGMA * mu / d), mu / (KAPPA_SIGMA * d));
            }
        }
        for (int r = 0; r < m; ++r) W->y[r] += alpha * W->dy[r];
        s_jac_fg(W, W->x, p, &f, W->grad, W->g, W->jnz);
    }
    /* IPOPT: when the algorithm cannot continue at a point that passes the acceptable-level test it stops there
     * (STOP_AT_ACCEPTABLE_POINT) instead of reporting the failure */
    if ((status == 2 || status == 3) && nacc > 0) status = 5;
    free(scratch);
    memcpy(x, W->x, sizeof(double) * n);
    if (lam_g) for (int r = 0; r < m; ++r) lam_g[r] = W->y[r] * W->dcs[r] / W->df;   /* multipliers of the unscaled problem */
    stats->status = status;
    stats->iters = it;
    stats->obj = f / W->df;
    stats->kkt_error = e0.E; stats->dual_inf = e0.dual_u; stats->constr_viol = e0.viol_u; stats->compl_inf = e0.compl_u;
    ws_free(W);
    free(lbs); free(ubs);
    return status;
}

int cmpc_oracle_ipm_solve_fn(const cmpc_oracle_nlp_fn* F, int N, const cmpc_oracle_ipm_opts* opts, const double* p,
                             const double* lbg, const double* ubg, double* x, double* lam_g,
                             cmpc_oracle_ipm_stats* stats)
{
    cmpc_oracle_ipm_opts defo;
    if (!opts) { cmpc_oracle_ipm_default_opts(&defo); opts = &defo; }
    if (!opts->mehrotra) return ipm_core(F, N, opts, 0, p, lbg, ubg, x, lam_g, stats);
    /* predictor-corrector mode; an instance it cannot finish (line-search / numerical failure, iteration limit) is solved
     * again from the same initial point on the monotone path */
    double* x0 = malloc(sizeof(double) * F->n);
    double* l0 = lam_g ? malloc(sizeof(double) * F->m) : NULL;
    memcpy(x0, x, sizeof(double) * F->n);
    if (lam_g) memcpy(l0, lam_g, sizeof(double) * F->m);
    int st = ipm_core(F, N, opts, 1, p, lbg, ubg, x, lam_g, stats);
    if (st == 2 || st == 3 || (st == 1 && stats->iters < opts->max_iter)) {   /* an exhausted max_iter is final */
        cmpc_oracle_ipm_stats s1 = *stats;
        memcpy(x, x0, sizeof(double) * F->n);
        if (lam_g) memcpy(lam_g, l0, sizeof(double) * F->m);
        st = ipm_core(F, N, opts, 0, p, lbg, ubg, x, lam_g, stats);
        stats->iters += s1.iters; stats->n_reg += s1.n_reg; stats->n_ls_trials += s1.n_ls_trials; stats->n_fallback = 1;
    }
    free(x0); free(l0);
    return st;
}

/* ------------------------------------------------------------------ adapters */
typedef struct restated_ctx { const cmpc_oracle_cfg* cfg; } restated_ctx;
static void r_fg(void* c, const double* x, const double* p, double* f, double* g) { cmpc_oracle_fg(((restated_ctx*)c)->cfg, x, p, f, g); }
static void r_jac(void* c, const double* x, const double* p, double* f, double* gr, double* g, double* j) { cmpc_oracle_jac_fg(((restated_ctx*)c)->cfg, x, p, f, gr, g, j); }
static void r_hess(void* c, const double* x, const double* p, double lf, const double* lg, double* h) { cmpc_oracle_hess_l(((restated_ctx*)c)->cfg, x, p, lf, lg, h); }

int cmpc_oracle_ipm_solve(const cmpc_oracle_cfg* cfg, const cmpc_oracle_ipm_opts* opts, const double* p,
                          const double* lbg, const double* ubg, double* x, double* lam_g, cmpc_oracle_ipm_stats* stats)
{
    const int N = cfg->N, n = cmpc_oracle_nx(N);
    int nj = cmpc_oracle_nnz_jac(N), nh = cmpc_oracle_nnz_hess(N);
    int* jc = malloc(sizeof(int) * (n + 1)); int* jr = malloc(sizeof(int) * nj);
    int* hc = malloc(sizeof(int) * (n + 1)); int* hr = malloc(sizeof(int) * nh);
    cmpc_oracle_jac_sparsity(N, jc, jr);
    cmpc_oracle_hess_sparsity(N, hc, hr);
    restated_ctx ctx = {cfg};
    cmpc_oracle_nlp_fn F = {n, cmpc_oracle_ng(N), nj, nh, jc, jr, hc, hr, r_fg, r_jac, r_hess, &ctx, 0};
    int rc = cmpc_oracle_ipm_solve_fn(&F, N, opts, p, lbg, ubg, x, lam_g, stats);
    free(jc); free(jr); free(hc); free(hr);
    return rc;
}

/* the same solver on a CasADi-generated shared object (oracle/_ref/libref_*.so, N = 12); CasADi C ABI tmp.c:12352 */
typedef int (*casadi_fn)(const double**, double**, long long*, double*, int);
typedef const long long* (*casadi_sp)(long long);
typedef struct casadi_ctx { casadi_fn fg, jac, hess; } casadi_ctx;
static void c_fg(void* c, const double* x, const double* p, double* f, double* g)
{
    const double* arg[2] = {x, p}; double ftmp; double* res[2] = {f ? f : &ftmp, g};
    ((casadi_ctx*)c)->fg(arg, res, NULL, NULL, 0);
}
static void c_jac(void* c, const double* x, const double* p, double* f, double* gr, double* g, double* j)
{
    const double* arg[2] = {x, p}; double* res[4] = {f, gr, g, j};
    ((casadi_ctx*)c)->jac(arg, res, NULL, NULL, 0);
}
static void c_hess(void* c, const double* x, const double* p, double lf, const double* lg, double* h)
{
    const double* arg[4] = {x, p, &lf, lg}; double* res[1] = {h};
    ((casadi_ctx*)c)->hess(arg, res, NULL, NULL, 0);
}
static void sp_to_int(const long long* sp, int** colind, int** row, int* nnz)
{
    int ncol = (int)sp[1];
    *colind = malloc(sizeof(int) * (ncol + 1));
    for (int i = 0; i <= ncol; ++i) (*colind)[i] = (int)sp[2 + i];
    *nnz = (*colind)[ncol];
    *row = malloc(sizeof(int) * *nnz);
    for (int i = 0; i < *nnz; ++i) (*row)[i] = (int)sp[2 + ncol + 1 + i];
}
int cmpc_oracle_ipm_solve_casadi(const char* so_path, const cmpc_oracle_ipm_opts* opts, const double* p,
                                 const double* lbg, const double* ubg, double* x, double* lam_g,
                                 cmpc_oracle_ipm_stats* stats)
{
    void* h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { memset(stats, 0, sizeof *stats); stats->status = 4; return 4; }
    casadi_ctx ctx = {(casadi_fn)dlsym(h, "nlp_fg"), (casadi_fn)dlsym(h, "nlp_jac_fg"), (casadi_fn)dlsym(h, "nlp_hess_l")};
    casadi_sp jsp = (casadi_sp)dlsym(h, "nlp_jac_fg_sparsity_out"), hsp = (casadi_sp)dlsym(h, "nlp_hess_l_sparsity_out");
    int *jc, *jr, *hc, *hr, nj, nh;
    sp_to_int(jsp(3), &jc, &jr, &nj);
    sp_to_int(hsp(0), &hc, &hr, &nh);
    const int N = 12;
    cmpc_oracle_nlp_fn F = {cmpc_oracle_nx(N), cmpc_oracle_ng(N), nj, nh, jc, jr, hc, hr, c_fg, c_jac, c_hess, &ctx, 0};
    int rc = cmpc_oracle_ipm_solve_fn(&F, N, opts, p, lbg, ubg, x, lam_g, stats);
    free(jc); free(jr); free(hc); free(hr);
    dlclose(h);
    return rc;
}

/* ------------------------------------------------------------------ batch (pthread pool) */
typedef struct batch_job {
    const cmpc_oracle_cfg* cfg; const cmpc_oracle_ipm_opts* opts;
    int batch; const double *p, *lbg, *ubg; double *x, *lam; cmpc_oracle_ipm_stats* stats;
    int next; pthread_mutex_t lock;
} batch_job;

static void* batch_worker(void* arg)
{
    batch_job* J = arg;
    const int N = J->cfg->N, n = cmpc_oracle_nx(N), np = cmpc_oracle_np(N), m = cmpc_oracle_ng(N);
    for (;;) {
        pthread_mutex_lock(&J->lock);
        int i = J->next++;
        pthread_mutex_unlock(&J->lock);
        if (i >= J->batch) break;
        cmpc_oracle_ipm_solve(J->cfg, J->opts, J->p + (size_t)i * np, J->lbg + (size_t)i * m, J->ubg + (size_t)i * m,
                              J->x + (size_t)i * n, J->lam ? J->lam + (size_t)i * m : NULL, &J->stats[i]);
    }
    return NULL;
}

int cmpc_oracle_ipm_solve_batch(const cmpc_oracle_cfg* cfg, const cmpc_oracle_ipm_opts* opts, int batch, int threads,
                                const double* p, const double* lbg, const double* ubg, double* x, double* lam_g,
                                cmpc_oracle_ipm_stats* stats)
{
    /* make sure the sparsity caches exist before the workers race for them */
    { int n = cmpc_oracle_nx(cfg->N); int* a = malloc(sizeof(int) * (n + 1)); int* b = malloc(sizeof(int) * cmpc_oracle_nnz_hess(cfg->N));
      cmpc_oracle_jac_sparsity(cfg->N, a, b); cmpc_oracle_hess_sparsity(cfg->N, a, b); free(a); free(b); }
    batch_job J = {cfg, opts, batch, p, lbg, ubg, x, lam_g, stats, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads < 1) threads = 1;
    if (threads > batch) threads = batch;
    pthread_t* th = malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, batch_worker, &J);
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th);
    int worst = 0;
    for (int i = 0; i < batch; ++i) if (stats[i].status > worst) worst = stats[i].status;
    return worst;
}
