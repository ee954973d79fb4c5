/*
 * cmpc_b200.h -- C ABI of libcmpc_b200.so: the batched centroidal-MPC solve on NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary for ONE path of GiulioRomualdi/paper_romualdi_2022_icra_centroidal-mpc-walking: the per-tick
 * nonlinear MPC solve that BipedalLocomotion::ReducedModelControllers::CentroidalMPC::advance() (member
 * m_controller, src/centroidal-mpc-walking/include/CentroidalMPCWalking/CentroidalMPCBlock.h:72; call site
 * src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:615) hands to CasADi + IPOPT/MUMPS.  The reference has no FFI
 * for this path (it links BLF/CasADi/IPOPT as C++ libraries); what an integrator binds instead is the CasADi C ABI of
 * the generated NLP (config/robots/ergoCubGazeboV1/tmp.c:12352: int nlp_*(const double** arg, double** res, ...)) plus
 * IPOPT's solve.  The entry points below take exactly those arrays, in exactly the reference's order
 * ("CasADi order"):
 *      x[45N+15]   decision vector      (tmp.c:62  casadi_s0)
 *      p[50N+27]   parameter vector     (tmp.c:63  casadi_s1)
 *      g[53N+15]   constraint rows, their bounds lbg/ubg and multipliers lam_g   (tmp.c:65  casadi_s3)
 * so that a buffer filled for the reference's solver can be handed over unchanged.  Batched arrays are instance major
 * (instance i at base + i * dim).  Plain pointers and sizes only; no C++/torch types.  All functions return 0 on
 * success or a negative CMPC_E_* code; they never throw and never fall back to a CPU path.
 *
 * Thread safety: calls on different handles are independent; a handle must not be used from two threads at once.
 * Stream semantics: a handle owns ONE work queue and ONE scratch arena, so it has at most one solve in flight: solves
 * enqueued on different streams are serialised by the library (an event recorded behind every solve; a solve on another
 * stream than the previous one waits for it on the device).  Use one handle per stream to overlap solves.  Every entry
 * point runs on the handle's device and restores the caller's current device before it returns.
 */
#ifndef CMPC_B200_H
#define CMPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CMPC_NUM_CONTACTS 2 /* number_of_maximum_contacts (every reference ini: 2; 0 = left_foot, 1 = right_foot) */
#define CMPC_NUM_CORNERS 4  /* number_of_corners */

enum {
    CMPC_OK = 0,
    CMPC_E_INVALID = -1,     /* bad argument (null pointer, batch < 0, unsupported number_of_slices, ...) */
    CMPC_E_CUDA = -2,        /* CUDA runtime error (cmpc_last_cuda_error gives the code)                   */
    CMPC_E_NO_DEVICE = -3,   /* no CUDA device: this library has no CPU path                              */
    CMPC_E_ALLOC = -4
};

/* per-instance solver status written to d_status / status */
enum {
    CMPC_STATUS_CONVERGED = 0,    /* scaled KKT error <= ipopt_tolerance (IPOPT "Optimal Solution Found")          */
    CMPC_STATUS_MAX_ITER = 1,     /* ipopt_max_iteration reached                                                     */
    CMPC_STATUS_LINE_SEARCH = 2,  /* filter line search failed at a point that is not acceptable (IPOPT would enter  */
                                  /* its restoration phase, which is not restated)                                   */
    CMPC_STATUS_NUMERICAL = 3,    /* regularisation exhausted / non-finite step                                      */
    CMPC_STATUS_BAD_INPUT = 4,    /* NaN, lbg > ubg, or an initial-condition / dynamics row with lbg != ubg          */
    CMPC_STATUS_ACCEPTABLE = 5    /* IPOPT "Solved To Acceptable Level": acceptable_iter consecutive iterates within  */
                                  /* acceptable_tol, or the algorithm could not continue from such a point; CasADi   */
                                  /* reports it as success, so does the host operator                               */
};

/* Mirrors the keys BLF CentroidalMPC::initialize reads from centroidal_mpc.ini
 * (src/centroidal-mpc-walking/config/robots/<robot>/centroidal_mpc.ini). */
typedef struct cmpc_config {
    int horizon;                            /* N = time_horizon / sampling_time  (or controller_horizon)            */
    double sampling_time;                   /* sampling_time / controller_sampling_time [s]                         */
    int number_of_slices;                   /* friction-cone slices per quadrant; only 1 is supported               */
    double static_friction_coefficient;
    double com_weight[3];
    double contact_position_weight;
    double force_rate_of_change_weight[3];
    double angular_momentum_weight;
    double contact_force_symmetry_weight;
    double corners[CMPC_NUM_CONTACTS][CMPC_NUM_CORNERS][3]; /* CONTACT_c / corner_j, contact frame               */
    double ipopt_tolerance;                 /* ipopt_tolerance, default 1e-8                                        */
    int ipopt_max_iteration;                /* ipopt_max_iteration, default 200                                     */
    double mu_init;                         /* IPOPT mu_init, default 0.1                                           */
    double bound_relax_factor;              /* IPOPT bound_relax_factor, default 1e-8                               */
    double bound_push;                      /* IPOPT bound_push = bound_frac, default 0.01                          */
    double infinity;                        /* |bound| >= infinity means no bound, default 1e19                     */
    int device;                             /* CUDA device ordinal                                                  */
    int threads_per_instance;               /* team size: 32, 64, 96 or 128 threads per instance, 0 = default (96)  */
    int ctas_per_sm;                        /* resident CTAs per SM used to size the persistent grid, 0 = occupancy */
    int teams_per_cta;                      /* teams walking in lock-step through one CTA: 1, 3 or 7; 0 = default    */
    int lockstep_groups;                    /* independent lock-step groups the teams of a CTA form: 1 .. teams_per_cta, */
                                            /* 0 = default (3: groups of 3 + 2 + 2 teams)                              */
    int mu_strategy;                        /* barrier-parameter update of the interior-point solve:                     */
                                            /* CMPC_MU_DEFAULT (0) = CMPC_MU_MEHROTRA; CMPC_MU_MONOTONE = IPOPT's default */
                                            /* Fiacco-McCormick update with IPOPT's constants, NLP scaling, termination   */
                                            /* and acceptable-level tests; NOT restated: restoration phase, second-order  */
                                            /* correction, watchdog (a solve IPOPT would rescue through them ends with    */
                                            /* CMPC_STATUS_LINE_SEARCH here);                                             */
                                            /* CMPC_MU_MEHROTRA = predictor-corrector (mu from the affine-scaling step,   */
                                            /* second-order corrector on the same factorisation, ~0.63 x the iterations;  */
                                            /* same termination test; an instance it cannot finish is re-solved monotone) */
    double warm_start_mu_init;              /* solves that start from given multipliers (warm_duals != 0): floor mu / slack of   */
                                            /* the bound multipliers and first barrier parameter of the monotone update;      */
                                            /* 0 = default 0.01 (closed loop at tol 1e-4: 4.4 instead of 5.4 iterations per   */
                                            /* tick with IPOPT's cold-start value 0.1; smaller values lengthen the slowest    */
                                            /* solves of a batch)                                                            */
    double nlp_scaling_max_gradient;        /* IPOPT nlp_scaling_max_gradient (gradient-based scaling, IPOPT's default method): the  */
                                            /* solve runs on  min(1, max_gradient / |grad f(x0)|_inf) * f ; ipopt_tolerance applies */
                                            /* to that scaled problem (dual_inf_tol 1, constr_viol_tol / compl_inf_tol 1e-4 to the   */
                                            /* unscaled one).  0 = IPOPT's default 100, negative = no scaling.  Rows of g are never  */
                                            /* scaled: their gradients are bounded far below 100              */
    double acceptable_tol;                  /* IPOPT acceptable_tol: 0 = IPOPT's default 1e-6, negative = off                        */
    int acceptable_iter;                    /* IPOPT acceptable_iter: 0 = IPOPT's default 15                                         */
} cmpc_config;
#define CMPC_MU_DEFAULT 0
#define CMPC_MU_MONOTONE 1
#define CMPC_MU_MEHROTRA 2

typedef struct cmpc_handle_s* cmpc_handle;

/* fills every field with the defaults above and the ergoCubGazeboV1_1 weights/corners; horizon 12, dT 0.1 */
int cmpc_default_config(cmpc_config* cfg);

/* dimension formulas of the NLP for horizon N; any pointer may be null */
int cmpc_dims(int horizon, int* n_x, int* n_p, int* n_g, int* nnz_jac, int* nnz_hess);

/* BLF Math::LinearizedFrictionCone, slices = 1: A[4*3] row major, rows A f <= 0 (literals at tmp.c:8599,8602) */
int cmpc_friction_matrix(double static_friction_coefficient, int number_of_slices, double* A);

/* CasADi compressed-column patterns of jac_g (casadi_s5, tmp.c:67) and hess_l (casadi_s4, tmp.c:66, full symmetric) */
int cmpc_jac_sparsity(int horizon, int* colind, int* row);
int cmpc_hess_sparsity(int horizon, int* colind, int* row);

int cmpc_create(const cmpc_config* cfg, cmpc_handle* out);
int cmpc_destroy(cmpc_handle h);

/* ---- the hot path: what CentroidalMPC::advance() delegates to nlpsol/IPOPT (CentroidalMPCBlock.cpp:615) ------------
 * Solves `batch` independent instances  min f(x, p)  s.t.  lbg <= g(x, p) <= ubg.
 * DEVICE pointers.  d_x: in = initial guess (warm start), out = solution.  d_lam_g: out = multipliers of g
 * (CasADi/IPOPT sign: >= 0 at an active upper bound); read as initial multipliers when warm_duals != 0.
 * d_obj[batch], d_status[batch], d_iters[batch] may be null.  d_iters counts every interior-point iteration spent on
 * the instance (an instance handed from the predictor-corrector to the monotone path reports the sum).  `stream` is a
 * cudaStream_t (null = default stream); the call is asynchronous with respect to the host.  One kernel launch per call:
 * batches of up to 4 instances per SM run on independent single-team CTAs, larger ones on persistent CTAs of seven
 * teams walking in lock-step. */
int cmpc_solve_batched(cmpc_handle h, int batch, const double* d_p, const double* d_lbg, const double* d_ubg,
                       double* d_x, double* d_lam_g, double* d_obj, int* d_status, int* d_iters, int warm_duals,
                       void* stream);

/* Same solve with HOST pointers: copies p/lbg/ubg/x (and lam_g when warm_duals) to the device, solves, copies
 * x/lam_g/obj/status/iters back and synchronises.  This is the call a CentroidalMPC host object makes per tick. */
int cmpc_solve_host(cmpc_handle h, int batch, const double* p, const double* lbg, const double* ubg, double* x,
                    double* lam_g, double* obj, int* status, int* iters, int warm_duals);

/* Warm-start shift between ticks (BLF is_warm_start_enabled): every knot-indexed block of x and lam_g moves one knot
 * towards the present, the last knot is repeated.  DEVICE pointers, in place.  d_lam_g may be null. */
int cmpc_shift_warmstart(cmpc_handle h, int batch, double* d_x, double* d_lam_g, void* stream);

/* ---- the NLP functions themselves (parity surface for nlp_fg / nlp_jac_fg / nlp_hess_l of tmp.c) ------------------
 * DEVICE pointers, instance major; outputs may be null.  jac/hess nonzeros in the CasADi CSC order above. */
int cmpc_eval_fg(cmpc_handle h, int batch, const double* d_x, const double* d_p, double* d_f, double* d_g, void* stream);
int cmpc_eval_jac_fg(cmpc_handle h, int batch, const double* d_x, const double* d_p, double* d_f, double* d_grad_f,
                     double* d_g, double* d_jac_nz, void* stream);
int cmpc_eval_hess_l(cmpc_handle h, int batch, const double* d_x, const double* d_p, double lam_f,
                     const double* d_lam_g, double* d_hess_nz, void* stream);

/* ---- closed loop plant (BLF ContinuousDynamicalSystem::CentroidalDynamics + RK4 as used at
 * src/centroidal-mpc-walking/src/WholeBodyQPBlock.cpp:1083-1090, 1150-1158): integrates (com, dcom, h) of every
 * instance for `substeps` RK4 steps of `dt` under the knot-0 corner forces of d_x and the knot-0 contact data of d_p,
 * plus an external wrench d_ext[batch][6] (per unit mass; null = none).  d_state[batch][9] in/out. DEVICE pointers. */
int cmpc_rollout_plant(cmpc_handle h, int batch, const double* d_x, const double* d_p, const double* d_ext,
                       double* d_state, double dt, int substeps, void* stream);

/* measured FP64 FMA throughput of the handle's device [TFLOP/s]: 8 independent DFMA chains per thread, no memory traffic.
 * The denominator of the solver kernel's FP64 roofline (MEASURED_PEAKS.json has no FP64 figure). */
int cmpc_measure_fp64_peak(cmpc_handle h, double* tflops);

/* debugging aid: accumulated clock64() cycles of thread 0 per solve phase (all zeros unless built with -DCMPC_PROFILE) */
int cmpc_debug_profile(long long* out16);

/* number of kernels this handle has launched so far, and the last CUDA error code seen by any call */
long long cmpc_launch_count(cmpc_handle h);
int cmpc_last_cuda_error(cmpc_handle h);
const char* cmpc_error_string(int code);

/* grid geometry chosen for the solver kernel (persistent CTAs): for reporting */
int cmpc_solver_geometry(cmpc_handle h, int* grid_ctas, int* threads, int* smem_bytes, int* ctas_per_sm, int* sm_count);
/* teams per CTA and the number of independent lock-step groups they form: for reporting */
int cmpc_solver_lockstep(cmpc_handle h, int* teams_per_cta, int* lockstep_groups);

#ifdef __cplusplus
}
#endif
#endif
