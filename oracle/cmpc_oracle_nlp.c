/*
 * cmpc_oracle_nlp.c -- CPU ORACLE (test infrastructure only, see cmpc_oracle.h).
 *
 * Plain-C restatement, for arbitrary horizon N, of the NLP functions the reference ships as CasADi-generated C:
 *   /root/reference/src/centroidal-mpc-walking/config/robots/ergoCubGazeboV1/tmp.c
 *     nlp / nlp_fg   :69, :12430      -> cmpc_oracle_fg
 *     nlp_jac_fg     :71962           -> cmpc_oracle_jac_fg   (CSC pattern = casadi_s5, tmp.c:67)
 *     nlp_hess_l     :58926           -> cmpc_oracle_hess_l   (CSC pattern = casadi_s4, tmp.c:66, full symmetric)
 * Layout of x, p, g: SURVEY.md section 8(a) rows a-1, a-2, a-4 (decoded from the generated code by probing).
 * Pinned by tests/test_oracle_nlp.py against the compiled reference file (oracle/_ref) and the golden vectors.
 */
#include "cmpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NC CMPC_O_NC
#define NJ CMPC_O_NJ
#define NF CMPC_O_NF

int cmpc_oracle_nx(int N) { return 45 * N + 15; }
int cmpc_oracle_np(int N) { return 50 * N + 27; }
int cmpc_oracle_ng(int N) { return 53 * N + 15; }
int cmpc_oracle_nnz_jac(int N) { return 243 * N + 15; }
int cmpc_oracle_nnz_hess(int N) { return 348 * N - 36; }

/* ---- index helpers (CasADi order) ------------------------------------------------------------------- */
static inline int x_com(int N, int k) { (void)N; return 3 * k; }
static inline int x_dcom(int N, int k) { return 3 * (N + 1) + 3 * k; }
static inline int x_h(int N, int k) { return 6 * (N + 1) + 3 * k; }
static inline int x_cbase(int N, int c) { return 9 * (N + 1) + c * (18 * N + 3); }
static inline int x_pos(int N, int c, int k) { return x_cbase(N, c) + 3 * k; }
static inline int x_vel(int N, int c, int k) { return x_cbase(N, c) + 3 * (N + 1) + 3 * k; }
static inline int x_frc(int N, int c, int j, int k) { return x_cbase(N, c) + 6 * N + 3 + 3 * N * j + 3 * k; }

static inline int p_cbase(int N, int c) { return c * (19 * N + 6); }
static inline int p_rot(int N, int c, int k) { return p_cbase(N, c) + 9 * k; }          /* column-major 3x3 */
static inline int p_en(int N, int c, int k) { return p_cbase(N, c) + 15 * N + k; }
static inline int p_nom(int N, int c, int k) { return p_cbase(N, c) + 16 * N + 3 * k; }
static inline int p_glob(int N) { return 38 * N + 12; }
static inline int p_comref(int N, int k) { return p_glob(N) + 9 + 3 * k; }
static inline int p_href(int N, int k) { return p_glob(N) + 9 + 3 * (N + 1) + 3 * k; }
static inline int p_extf(int N, int k) { return p_glob(N) + 9 + 6 * (N + 1) + 3 * k; }
static inline int p_extt(int N, int k) { return p_glob(N) + 9 + 6 * (N + 1) + 3 * N + 3 * k; }

static inline int g_init(int i) { return i; }                                  /* 0..8 com,dcom,h ; 9..14 pos */
static inline int g_com(int N, int k) { (void)N; return 15 + 3 * k; }
static inline int g_dcom(int N, int k) { return 15 + 3 * N + 3 * k; }
static inline int g_h(int N, int k) { return 15 + 6 * N + 3 * k; }
static inline int g_pos(int N, int c, int k) { return 15 + 9 * N + 3 * N * c + 3 * k; }
static inline int g_box(int N, int c, int k) { return 15 + 15 * N + c * (3 * N + 4 * NF * N) + 3 * k; }
static inline int g_fric(int N, int c, int j, int k)
{
    return 15 + 15 * N + c * (3 * N + 4 * NF * N) + 3 * N + (NJ * NF) * k + NF * j;
}

static const double GRAV_Z = -9.80665; /* tmp.c:3916 */

/* BLF Math::LinearizedFrictionCone (slices = 1): the cone boundary is sampled at angles i*pi/2 and each segment
 * between consecutive samples gives one half plane; reproduces the literals of tmp.c:8599,8602 bit for bit. */
void cmpc_oracle_friction_matrix(double mu, double A[NF][3])
{
    const int slices = 1;
    const int nrows = 4 * slices;
    const double seg = (M_PI / 2.0) / slices;
    for (int i = 0; i < nrows; ++i) {
        /* the cone boundary is a closed polygon: the segment after the last sample returns to sample 0
         * (this is what makes the (3,0) entry 0.9999999999999998 and not ...96, probed on tmp.c) */
        double a0 = seg * i, a1 = seg * (i + 1), a1w = seg * ((i + 1) % nrows);
        double x0 = cos(a0), y0 = sin(a0), x1 = cos(a1w), y1 = sin(a1w);
        double slope = (y1 - y0) / (x1 - x0);
        double offset = y0 - slope * x0;
        double sgn = (a0 > M_PI || a1 > M_PI) ? -1.0 : 1.0;
        A[i][0] = -sgn * slope;
        A[i][1] = sgn;
        A[i][2] = -sgn * offset * mu;
    }
}

/* omega_k of the CoM height cost (tmp.c:407-484): (w - w/2) exp(-k) + w/2, applied inside the square */
static inline double com_z_omega(double w, int k) { return (w - w / 2.0) * exp(-(double)k) + w / 2.0; }

static inline void rot_apply(const double* R, const double* v, double* out) /* out = R v, R column-major */
{
    for (int r = 0; r < 3; ++r) out[r] = R[r] * v[0] + R[3 + r] * v[1] + R[6 + r] * v[2];
}
static inline void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* ---- objective --------------------------------------------------------------------------------------- */
static double eval_f(const cmpc_oracle_cfg* cfg, const double* x, const double* p, double* grad)
{
    const int N = cfg->N;
    double f = 0.0;
    if (grad) memset(grad, 0, sizeof(double) * cmpc_oracle_nx(N));
    for (int k = 0; k <= N; ++k) {
        for (int a = 0; a < 3; ++a) {
            double e = x[x_h(N, k) + a] - p[p_href(N, k) + a];
            f += cfg->w_h * e * e;
            if (grad) grad[x_h(N, k) + a] += 2.0 * cfg->w_h * e;
        }
        for (int a = 0; a < 2; ++a) {
            double e = x[x_com(N, k) + a] - p[p_comref(N, k) + a];
            f += cfg->w_com[a] * e * e;
            if (grad) grad[x_com(N, k) + a] += 2.0 * cfg->w_com[a] * e;
        }
        {
            double om = com_z_omega(cfg->w_com[2], k);
            double e = x[x_com(N, k) + 2] - p[p_comref(N, k) + 2];
            f += (om * e) * (om * e);
            if (grad) grad[x_com(N, k) + 2] += 2.0 * om * om * e;
        }
        for (int c = 0; c < NC; ++c)
            for (int a = 0; a < 3; ++a) {
                double e = p[p_nom(N, c, k) + a] - x[x_pos(N, c, k) + a];
                f += cfg->w_pos * e * e;
                if (grad) grad[x_pos(N, c, k) + a] += -2.0 * cfg->w_pos * e;
            }
    }
    for (int c = 0; c < NC; ++c) {
        for (int k = 0; k < N; ++k) {
            double en = p[p_en(N, c, k)];
            for (int a = 0; a < 3; ++a) {
                double mean = 0.0;
                for (int j = 0; j < NJ; ++j) mean += x[x_frc(N, c, j, k) + a];
                mean *= en / NJ;
                double dsum = 0.0;
                for (int j = 0; j < NJ; ++j) {
                    double d = x[x_frc(N, c, j, k) + a] - mean;
                    f += cfg->w_sym * d * d;
                    dsum += d;
                }
                if (grad)
                    for (int j = 0; j < NJ; ++j) {
                        double d = x[x_frc(N, c, j, k) + a] - mean;
                        grad[x_frc(N, c, j, k) + a] += 2.0 * cfg->w_sym * (d - (en / NJ) * dsum);
                    }
            }
        }
        for (int j = 0; j < NJ; ++j)
            for (int k = 0; k + 1 < N; ++k)
                for (int a = 0; a < 3; ++a) {
                    double d = x[x_frc(N, c, j, k + 1) + a] - x[x_frc(N, c, j, k) + a];
                    f += cfg->w_rate[a] * d * d;
                    if (grad) {
                        grad[x_frc(N, c, j, k + 1) + a] += 2.0 * cfg->w_rate[a] * d;
                        grad[x_frc(N, c, j, k) + a] -= 2.0 * cfg->w_rate[a] * d;
                    }
                }
    }
    return f;
}

/* ---- constraints ------------------------------------------------------------------------------------- */
static void eval_g(const cmpc_oracle_cfg* cfg, const double* x, const double* p, double* g)
{
    const int N = cfg->N;
    const double dT = cfg->dT;
    double A[NF][3];
    cmpc_oracle_friction_matrix(cfg->mu, A);
    for (int a = 0; a < 3; ++a) {
        g[g_init(0) + a] = x[x_com(N, 0) + a];
        g[g_init(3) + a] = x[x_dcom(N, 0) + a];
        g[g_init(6) + a] = x[x_h(N, 0) + a];
        g[g_init(9) + a] = x[x_pos(N, 0, 0) + a];
        g[g_init(12) + a] = x[x_pos(N, 1, 0) + a];
    }
    for (int k = 0; k < N; ++k) {
        double fsum[3] = {0, 0, 0}, tsum[3] = {0, 0, 0};
        for (int c = 0; c < NC; ++c) {
            double en = p[p_en(N, c, k)];
            const double* R = p + p_rot(N, c, k);
            for (int j = 0; j < NJ; ++j) {
                const double* fr = x + x_frc(N, c, j, k);
                double Rr[3], rho[3], tq[3];
                rot_apply(R, cfg->corners[c][j], Rr);
                for (int a = 0; a < 3; ++a) rho[a] = Rr[a] + x[x_pos(N, c, k) + a] - x[x_com(N, k) + a];
                cross3(rho, fr, tq);
                for (int a = 0; a < 3; ++a) {
                    fsum[a] += en * fr[a];
                    tsum[a] += en * tq[a];
                }
            }
        }
        for (int a = 0; a < 3; ++a) {
            double grav = (a == 2) ? GRAV_Z : 0.0;
            g[g_com(N, k) + a] = x[x_com(N, k + 1) + a] - x[x_com(N, k) + a] - dT * x[x_dcom(N, k) + a];
            g[g_dcom(N, k) + a] = x[x_dcom(N, k + 1) + a] - x[x_dcom(N, k) + a]
                                  - dT * (grav + p[p_extf(N, k) + a] + fsum[a]);
            g[g_h(N, k) + a] = x[x_h(N, k + 1) + a] - x[x_h(N, k) + a] - dT * (p[p_extt(N, k) + a] + tsum[a]);
        }
        for (int c = 0; c < NC; ++c) {
            double en = p[p_en(N, c, k)];
            const double* R = p + p_rot(N, c, k);
            for (int a = 0; a < 3; ++a)
                g[g_pos(N, c, k) + a] = x[x_pos(N, c, k + 1) + a] - x[x_pos(N, c, k) + a]
                                        - (1.0 - en) * dT * x[x_vel(N, c, k) + a];
            /* box: R_k^T (pos_{k+1} - nominal_{k+1}) */
            double d[3];
            for (int a = 0; a < 3; ++a) d[a] = x[x_pos(N, c, k + 1) + a] - p[p_nom(N, c, k + 1) + a];
            for (int r = 0; r < 3; ++r) g[g_box(N, c, k) + r] = R[3 * r] * d[0] + R[3 * r + 1] * d[1] + R[3 * r + 2] * d[2];
            /* friction: A R_k^T f */
            for (int j = 0; j < NJ; ++j) {
                const double* fr = x + x_frc(N, c, j, k);
                double fl[3];
                for (int r = 0; r < 3; ++r) fl[r] = R[3 * r] * fr[0] + R[3 * r + 1] * fr[1] + R[3 * r + 2] * fr[2];
                for (int r = 0; r < NF; ++r)
                    g[g_fric(N, c, j, k) + r] = A[r][0] * fl[0] + A[r][1] * fl[1] + A[r][2] * fl[2];
            }
        }
    }
}

void cmpc_oracle_fg(const cmpc_oracle_cfg* cfg, const double* x, const double* p, double* f, double* g)
{
    if (f) *f = eval_f(cfg, x, p, NULL);
    if (g) eval_g(cfg, x, p, g);
}

/* ---- sparse emission machinery ------------------------------------------------------------------------
 * Each derivative routine visits its structural nonzeros in a fixed order through emit(); the CSC position of
 * the e-th emitted entry is computed once per N (sort by column, then row) and cached. */
typedef struct emitter {
    int mode;      /* 0: record pattern, 1: write values */
    int count;
    int* rows;     /* mode 0 */
    int* cols;     /* mode 0 */
    const int* slot; /* mode 1: emission index -> nz index */
    double* nz;    /* mode 1 */
} emitter;

static inline void emit(emitter* E, int r, int c, double v)
{
    if (E->mode == 0) {
        E->rows[E->count] = r;
        E->cols[E->count] = c;
    } else {
        E->nz[E->slot[E->count]] = v;
    }
    E->count++;
}

/* 6 structural off-diagonal entries of s*[a]x placed at (r0, c0) */
static inline void emit_skew(emitter* E, int r0, int c0, const double* a, double s)
{
    emit(E, r0 + 0, c0 + 1, -s * a[2]);
    emit(E, r0 + 0, c0 + 2, s * a[1]);
    emit(E, r0 + 1, c0 + 0, s * a[2]);
    emit(E, r0 + 1, c0 + 2, -s * a[0]);
    emit(E, r0 + 2, c0 + 0, -s * a[1]);
    emit(E, r0 + 2, c0 + 1, s * a[0]);
}

static void jac_visit(const cmpc_oracle_cfg* cfg, const double* x, const double* p, emitter* E)
{
    const int N = cfg->N;
    const double dT = cfg->dT;
    static const double zero3[3] = {0, 0, 0};
    double A[NF][3];
    cmpc_oracle_friction_matrix(cfg->mu, A);
    for (int a = 0; a < 3; ++a) {
        emit(E, g_init(0) + a, x_com(N, 0) + a, 1.0);
        emit(E, g_init(3) + a, x_dcom(N, 0) + a, 1.0);
        emit(E, g_init(6) + a, x_h(N, 0) + a, 1.0);
        emit(E, g_init(9) + a, x_pos(N, 0, 0) + a, 1.0);
        emit(E, g_init(12) + a, x_pos(N, 1, 0) + a, 1.0);
    }
    for (int k = 0; k < N; ++k) {
        for (int a = 0; a < 3; ++a) {
            emit(E, g_com(N, k) + a, x_com(N, k + 1) + a, 1.0);
            emit(E, g_com(N, k) + a, x_com(N, k) + a, -1.0);
            emit(E, g_com(N, k) + a, x_dcom(N, k) + a, -dT);
            emit(E, g_dcom(N, k) + a, x_dcom(N, k + 1) + a, 1.0);
            emit(E, g_dcom(N, k) + a, x_dcom(N, k) + a, -1.0);
            emit(E, g_h(N, k) + a, x_h(N, k + 1) + a, 1.0);
            emit(E, g_h(N, k) + a, x_h(N, k) + a, -1.0);
        }
        double Fall[3] = {0, 0, 0};
        for (int c = 0; c < NC; ++c) {
            double en = x ? p[p_en(N, c, k)] : 0.0;
            const double* R = x ? p + p_rot(N, c, k) : NULL;
            double Fc[3] = {0, 0, 0};
            for (int j = 0; j < NJ; ++j) {
                const double* fr = x ? x + x_frc(N, c, j, k) : zero3;
                double rho[3] = {0, 0, 0};
                if (x) {
                    double Rr[3];
                    rot_apply(R, cfg->corners[c][j], Rr);
                    for (int a = 0; a < 3; ++a) rho[a] = Rr[a] + x[x_pos(N, c, k) + a] - x[x_com(N, k) + a];
                }
                for (int a = 0; a < 3; ++a) {
                    emit(E, g_dcom(N, k) + a, x_frc(N, c, j, k) + a, -dT * en);
                    Fc[a] += fr[a];
                }
                /* d(rho x f)/df = [rho]x  ->  g_h gets -dT en [rho]x */
                emit_skew(E, g_h(N, k), x_frc(N, c, j, k), rho, -dT * en);
            }
            /* d(rho x F)/dpos = -[F]x -> g_h gets +dT en [Fc]x */
            emit_skew(E, g_h(N, k), x_pos(N, c, k), Fc, dT * en);
            for (int a = 0; a < 3; ++a) Fall[a] += en * Fc[a];
            for (int a = 0; a < 3; ++a) {
                emit(E, g_pos(N, c, k) + a, x_pos(N, c, k + 1) + a, 1.0);
                emit(E, g_pos(N, c, k) + a, x_pos(N, c, k) + a, -1.0);
                emit(E, g_pos(N, c, k) + a, x_vel(N, c, k) + a, -(1.0 - en) * dT);
            }
            for (int r = 0; r < 3; ++r)
                for (int a = 0; a < 3; ++a) emit(E, g_box(N, c, k) + r, x_pos(N, c, k + 1) + a, R ? R[3 * r + a] : 0.0);
            for (int j = 0; j < NJ; ++j)
                for (int r = 0; r < NF; ++r)
                    for (int a = 0; a < 3; ++a) {
                        double v = 0.0;
                        if (R) v = A[r][0] * R[a] + A[r][1] * R[3 + a] + A[r][2] * R[6 + a]; /* (A R^T)[r][a] */
                        emit(E, g_fric(N, c, j, k) + r, x_frc(N, c, j, k) + a, v);
                    }
        }
        /* d(-com x F)/dcom = [F]x -> g_h gets -dT [sum_c en Fc]x */
        emit_skew(E, g_h(N, k), x_com(N, k), Fall, -dT);
    }
}

static void hess_visit(const cmpc_oracle_cfg* cfg, const double* p, double lam_f, const double* lam_g, emitter* E)
{
    const int N = cfg->N;
    const double dT = cfg->dT;
    static const double zero3[3] = {0, 0, 0};
    for (int k = 0; k <= N; ++k) {
        double om = com_z_omega(cfg->w_com[2], k);
        emit(E, x_com(N, k) + 0, x_com(N, k) + 0, lam_f * 2.0 * cfg->w_com[0]);
        emit(E, x_com(N, k) + 1, x_com(N, k) + 1, lam_f * 2.0 * cfg->w_com[1]);
        emit(E, x_com(N, k) + 2, x_com(N, k) + 2, lam_f * 2.0 * om * om);
        for (int a = 0; a < 3; ++a) emit(E, x_h(N, k) + a, x_h(N, k) + a, lam_f * 2.0 * cfg->w_h);
        for (int c = 0; c < NC; ++c)
            for (int a = 0; a < 3; ++a) emit(E, x_pos(N, c, k) + a, x_pos(N, c, k) + a, lam_f * 2.0 * cfg->w_pos);
    }
    for (int c = 0; c < NC; ++c)
        for (int k = 0; k < N; ++k) {
            double en = p ? p[p_en(N, c, k)] : 0.0;
            double a4 = en / NJ;
            double dsym = 2.0 * cfg->w_sym * (1.0 - 2.0 * a4 + NJ * a4 * a4);
            double osym = 2.0 * cfg->w_sym * (NJ * a4 * a4 - 2.0 * a4);
            const double* lam = lam_g ? lam_g + g_h(N, k) : zero3;
            int nrate = (N >= 2) ? ((k == 0 || k == N - 1) ? 1 : 2) : 0;
            for (int j = 0; j < NJ; ++j) {
                for (int a = 0; a < 3; ++a) {
                    int idx = x_frc(N, c, j, k) + a;
                    emit(E, idx, idx, lam_f * (dsym + 2.0 * cfg->w_rate[a] * nrate));
                    for (int j2 = 0; j2 < NJ; ++j2)
                        if (j2 != j) emit(E, idx, x_frc(N, c, j2, k) + a, lam_f * osym);
                    if (k + 1 < N) {
                        emit(E, idx, x_frc(N, c, j, k + 1) + a, -lam_f * 2.0 * cfg->w_rate[a]);
                        emit(E, x_frc(N, c, j, k + 1) + a, idx, -lam_f * 2.0 * cfg->w_rate[a]);
                    }
                }
                /* lam' g_h contributes -dT en lam.(rho x f):  d2/(drho df) = -[lam]x (rows rho, cols f) */
                emit_skew(E, x_pos(N, c, k), x_frc(N, c, j, k), lam, dT * en);   /* rows pos,  cols f  */
                emit_skew(E, x_frc(N, c, j, k), x_pos(N, c, k), lam, -dT * en);  /* transpose            */
                emit_skew(E, x_com(N, k), x_frc(N, c, j, k), lam, -dT * en);     /* drho/dcom = -I       */
                emit_skew(E, x_frc(N, c, j, k), x_com(N, k), lam, dT * en);
            }
        }
}

/* pattern cache ---------------------------------------------------------------------------------------- */
typedef struct pattern {
    int N, nnz, ncol;
    int* colind;
    int* row;
    int* slot;
} pattern;

#define MAX_CACHED 16
static pattern jac_cache[MAX_CACHED], hess_cache[MAX_CACHED];
static int n_jac_cache = 0, n_hess_cache = 0;
#include <pthread.h>
static pthread_mutex_t cache_lock = PTHREAD_MUTEX_INITIALIZER;

typedef struct triple { int r, c, e; } triple;
static int cmp_triple(const void* a, const void* b)
{
    const triple *A = a, *B = b;
    if (A->c != B->c) return A->c - B->c;
    if (A->r != B->r) return A->r - B->r;
    return A->e - B->e;
}

static void build_pattern(pattern* P, int N, int which)
{
    cmpc_oracle_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.N = N;
    int nnz = which == 0 ? cmpc_oracle_nnz_jac(N) : cmpc_oracle_nnz_hess(N);
    int cap = nnz + 64;
    emitter E = {0, 0, malloc(sizeof(int) * cap), malloc(sizeof(int) * cap), NULL, NULL};
    if (which == 0)
        jac_visit(&cfg, NULL, NULL, &E);
    else
        hess_visit(&cfg, NULL, 0.0, NULL, &E);
    triple* T = malloc(sizeof(triple) * E.count);
    for (int e = 0; e < E.count; ++e) { T[e].r = E.rows[e]; T[e].c = E.cols[e]; T[e].e = e; }
    qsort(T, E.count, sizeof(triple), cmp_triple);
    P->N = N;
    P->nnz = E.count;
    P->ncol = cmpc_oracle_nx(N);
    P->colind = calloc(P->ncol + 1, sizeof(int));
    P->row = malloc(sizeof(int) * E.count);
    P->slot = malloc(sizeof(int) * E.count);
    for (int i = 0; i < E.count; ++i) {
        P->row[i] = T[i].r;
        P->slot[T[i].e] = i;
        P->colind[T[i].c + 1]++;
    }
    for (int c = 0; c < P->ncol; ++c) P->colind[c + 1] += P->colind[c];
    free(T); free(E.rows); free(E.cols);
}

static const pattern* get_pattern(int N, int which)
{
    pattern* cache = which == 0 ? jac_cache : hess_cache;
    int* n = which == 0 ? &n_jac_cache : &n_hess_cache;
    pthread_mutex_lock(&cache_lock);
    for (int i = 0; i < *n; ++i)
        if (cache[i].N == N) { pthread_mutex_unlock(&cache_lock); return &cache[i]; }
    if (*n == MAX_CACHED) { pthread_mutex_unlock(&cache_lock); return NULL; }
    build_pattern(&cache[*n], N, which);
    const pattern* P = &cache[(*n)++];
    pthread_mutex_unlock(&cache_lock);
    return P;
}

void cmpc_oracle_jac_sparsity(int N, int* colind, int* row)
{
    const pattern* P = get_pattern(N, 0);
    memcpy(colind, P->colind, sizeof(int) * (P->ncol + 1));
    memcpy(row, P->row, sizeof(int) * P->nnz);
}
void cmpc_oracle_hess_sparsity(int N, int* colind, int* row)
{
    const pattern* P = get_pattern(N, 1);
    memcpy(colind, P->colind, sizeof(int) * (P->ncol + 1));
    memcpy(row, P->row, sizeof(int) * P->nnz);
}

void cmpc_oracle_jac_fg(const cmpc_oracle_cfg* cfg, const double* x, const double* p, double* f, double* grad,
                        double* g, double* jac_nz)
{
    if (f || grad) {
        double v = eval_f(cfg, x, p, grad);
        if (f) *f = v;
    }
    if (g) eval_g(cfg, x, p, g);
    if (jac_nz) {
        const pattern* P = get_pattern(cfg->N, 0);
        emitter E = {1, 0, NULL, NULL, P->slot, jac_nz};
        jac_visit(cfg, x, p, &E);
    }
}

void cmpc_oracle_hess_l(const cmpc_oracle_cfg* cfg, const double* x, const double* p, double lam_f,
                        const double* lam_g, double* hess_nz)
{
    (void)x; /* the hessian of the lagrangian does not depend on x */
    const pattern* P = get_pattern(cfg->N, 1);
    emitter E = {1, 0, NULL, NULL, P->slot, hess_nz};
    hess_visit(cfg, p, lam_f, lam_g, &E);
}
