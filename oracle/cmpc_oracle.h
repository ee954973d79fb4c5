/*
 * cmpc_oracle.h -- CPU ORACLE for the centroidal-MPC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product (libcmpc_b200.so) never links, loads or falls back to it.
 *
 * What it restates (reference = /root/reference, GiulioRomualdi/paper_romualdi_2022_icra_centroidal-mpc-walking):
 *   * the NLP  nlp:(x[45N+15], p[50N+27]) -> (f, g[53N+15])  that BLF CentroidalMPC hands to CasADi, as shipped in
 *     generated form in  src/centroidal-mpc-walking/config/robots/ergoCubGazeboV1/tmp.c  (nlp :69, nlp_fg :12430,
 *     nlp_hess_l :58926, nlp_jac_fg :71962; sparsity casadi_s4/s5 :66-67), for ARBITRARY horizon N.
 *     PINNED: checked entry-for-entry against the compiled reference file (oracle/_ref) at N=12, both weight sets,
 *     and against the known-answer values of SURVEY.md section 4 (tests/test_oracle_nlp.py).
 *   * the solve that CentroidalMPC::advance() (call site CentroidalMPCBlock.cpp:615) delegates to IPOPT
 *     (un-vendored third-party dependency, IPOPT 3.13.4 + MUMPS, dockerfiles/Dockerfile:49): a primal-dual
 *     interior-point method following the published algorithm (Waechter & Biegler, Math. Prog. 106, 2006) with
 *     IPOPT's default option values.  SOLVER-LEVEL PARITY IS UNPINNED: neither IPOPT nor CasADi can be run in this
 *     container and the reference has no tests/golden vectors for advance(); the solve is pinned only through
 *     KKT residuals evaluated with the reference's own generated functions (oracle/_ref).
 */
#ifndef CMPC_ORACLE_H
#define CMPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define CMPC_O_NC 2 /* contacts (number_of_maximum_contacts 2 in every reference ini) */
#define CMPC_O_NJ 4 /* corners per contact (number_of_corners 4 in every reference ini) */
#define CMPC_O_NF 4 /* friction rows per corner = 4 * number_of_slices, slices = 1 */

typedef struct cmpc_oracle_cfg {
    int N;                               /* horizon knots  (time_horizon / sampling_time)            */
    double dT;                           /* sampling_time [s]                                         */
    double mu;                           /* static_friction_coefficient                               */
    double w_com[3];                     /* com_weight                                                */
    double w_h;                          /* angular_momentum_weight                                   */
    double w_pos;                        /* contact_position_weight                                   */
    double w_sym;                        /* contact_force_symmetry_weight                             */
    double w_rate[3];                    /* force_rate_of_change_weight                               */
    double corners[CMPC_O_NC][CMPC_O_NJ][3]; /* CONTACT_c / corner_j, contact frame               */
} cmpc_oracle_cfg;

/* dimension formulas */
int cmpc_oracle_nx(int N);      /* 45N+15  */
int cmpc_oracle_np(int N);      /* 50N+27  */
int cmpc_oracle_ng(int N);      /* 53N+15  */
int cmpc_oracle_nnz_jac(int N); /* 243N+15 */
int cmpc_oracle_nnz_hess(int N);/* 348N-36 */

/* BLF Math::LinearizedFrictionCone restatement, slices = 1: A[4][3], rows A f <= 0 (tmp.c:8599,8602) */
void cmpc_oracle_friction_matrix(double mu, double A[CMPC_O_NF][3]);

/* CasADi compressed-column sparsity: colind[n+1], row[nnz] (same arrays as casadi_s5 / casadi_s4 in tmp.c:66-67) */
void cmpc_oracle_jac_sparsity(int N, int* colind, int* row);
void cmpc_oracle_hess_sparsity(int N, int* colind, int* row);

/* nlp_fg (tmp.c:12430) */
void cmpc_oracle_fg(const cmpc_oracle_cfg* cfg, const double* x, const double* p, double* f, double* g);
/* nlp_jac_fg (tmp.c:71962): f, grad_f[n], g[m], jac nonzeros in CSC order */
void cmpc_oracle_jac_fg(const cmpc_oracle_cfg* cfg, const double* x, const double* p, double* f, double* grad,
                        double* g, double* jac_nz);
/* nlp_hess_l (tmp.c:58926): full symmetric hessian of lam_f*f + lam_g'g in CSC order */
void cmpc_oracle_hess_l(const cmpc_oracle_cfg* cfg, const double* x, const double* p, double lam_f,
                        const double* lam_g, double* hess_nz);

/* ------------------------------------------------------------------ interior point (IPOPT restatement) */
typedef struct cmpc_oracle_ipm_opts {
    double tol;              /* ipopt_tolerance (IPOPT 'tol'), default 1e-8                      */
    int max_iter;            /* ipopt_max_iteration, default 3000 in IPOPT; we default to 200   */
    double mu_init;          /* 0.1                                                               */
    double bound_relax;      /* bound_relax_factor 1e-8                                           */
    double bound_push;       /* bound_push = bound_frac = 0.01                                    */
    double inf_bound;        /* |b| >= inf_bound means no bound (nlp_*_bound_inf 1e19)            */
    int warm_duals;          /* 1: use lam_g on entry as initial multipliers                      */
    int verbose;
    int mehrotra;            /* 0: IPOPT's default monotone barrier update (the restatement of the reference's solve);
                              * 1: Mehrotra predictor-corrector barrier update (mu from the affine-scaling step, second-order
                              *    corrector with the same factorisation), the mu strategy of the CUDA path's default mode;
                              *    instances it cannot finish are re-solved on the monotone path                       */
    double nlp_scaling_max_gradient; /* IPOPT nlp_scaling_method gradient-based (its default): f is scaled by
                              * min(1, max_gradient / |grad f(x0)|_inf), row i of g by min(1, max_gradient / |grad g_i(x0)|_inf),
                              * floor nlp_scaling_min_value 1e-8; tol applies to the SCALED problem, dual_inf_tol /
                              * constr_viol_tol / compl_inf_tol to the unscaled one.  Default 100; 0 = no scaling           */
    double acceptable_tol;   /* IPOPT acceptable_tol 1e-6 (with acceptable_dual_inf_tol 1e10, acceptable_constr_viol_tol 1e-2,
                              * acceptable_compl_inf_tol 1e-2): acceptable_iter consecutive acceptable iterates end the solve
                              * with status 5 ("Solved To Acceptable Level"); a line-search / numerical failure at an
                              * acceptable point also reports 5.  0 = off                                                  */
    int acceptable_iter;     /* IPOPT acceptable_iter 15                                                                   */
} cmpc_oracle_ipm_opts;

void cmpc_oracle_ipm_default_opts(cmpc_oracle_ipm_opts* o);

typedef struct cmpc_oracle_ipm_stats {
    int status;      /* 0 converged, 1 max_iter, 2 line-search failure, 3 numerical failure, 4 bad input,
                      * 5 solved to acceptable level                                                       */
    int iters;
    double obj;
    double kkt_error;   /* final scaled optimality error E_0                */
    double dual_inf, constr_viol, compl_inf; /* unscaled components          */
    int n_reg;          /* iterations that needed delta_w > 0                */
    int n_ls_trials;    /* total backtracking trials                         */
    int n_fallback;     /* 1: the predictor-corrector run failed, monotone re-solve */
    double obj_scaling;     /* d_f of the gradient-based scaling (1 = none)              */
    double min_g_scaling;   /* smallest row scaling d_c (1 = no row was scaled)          */
} cmpc_oracle_ipm_stats;

/* callbacks so that the same IPM can run on the restated NLP or on the compiled reference functions */
typedef struct cmpc_oracle_nlp_fn {
    int n, m, nnz_j, nnz_h;
    const int *jc, *jr; /* jac CSC */
    const int *hc, *hr; /* hess CSC (full symmetric) */
    void (*fg)(void* ctx, const double* x, const double* p, double* f, double* g);
    void (*jac_fg)(void* ctx, const double* x, const double* p, double* f, double* grad, double* g, double* jnz);
    void (*hess)(void* ctx, const double* x, const double* p, double lam_f, const double* lam_g, double* hnz);
    void* ctx;
    int stage_of_x_ready; /* internal */
} cmpc_oracle_nlp_fn;

/* Solve  min f(x,p)  s.t. lbg <= g(x,p) <= ubg.  x: in = initial guess, out = solution.  lam_g: multipliers of g
 * (IPOPT/CasADi sign convention: lam_g >= 0 at an active upper bound).  Uses the restated NLP of `cfg`. */
int cmpc_oracle_ipm_solve(const cmpc_oracle_cfg* cfg, const cmpc_oracle_ipm_opts* opts, const double* p,
                          const double* lbg, const double* ubg, double* x, double* lam_g,
                          cmpc_oracle_ipm_stats* stats);

/* Same solver on user supplied callbacks (used by the tests to run it on oracle/_ref's functions). */
int cmpc_oracle_ipm_solve_fn(const cmpc_oracle_nlp_fn* nlp, int N, const cmpc_oracle_ipm_opts* opts,
                             const double* p, const double* lbg, const double* ubg, double* x, double* lam_g,
                             cmpc_oracle_ipm_stats* stats);

/* Same solver on a CasADi-generated shared object exporting nlp_fg / nlp_jac_fg / nlp_hess_l (oracle/_ref, N = 12). */
int cmpc_oracle_ipm_solve_casadi(const char* so_path, const cmpc_oracle_ipm_opts* opts, const double* p,
                                 const double* lbg, const double* ubg, double* x, double* lam_g,
                                 cmpc_oracle_ipm_stats* stats);

/* batch driver with a pthread pool (cpu_baseline leg of bench.py): instance i uses p+i*np etc. */
int cmpc_oracle_ipm_solve_batch(const cmpc_oracle_cfg* cfg, const cmpc_oracle_ipm_opts* opts, int batch,
                                int threads, const double* p, const double* lbg, const double* ubg, double* x,
                                double* lam_g, cmpc_oracle_ipm_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
