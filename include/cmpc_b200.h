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
 * stream than the previous one waits for it on the device; calls made while the stream is captured into a CUDA graph are
 * ordered by the graph's own stream).  Use one handle per stream to overlap solves.  Every entry
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
    int threads_per_instance;               /* team size: 32, 64, 96, 128 (192, 256: one team per CTA), 0 = default */
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
    double bounding_box_upper_limit[CMPC_NUM_CONTACTS][3]; /* CONTACT_c / bounding_box_upper_limit, _lower_limit: the step-adjustment */
    double bounding_box_lower_limit[CMPC_NUM_CONTACTS][3]; /* box of a future contact, contact frame (used by cmpc_populate only)     */
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
 * batches of up to 4 instances per SM run on independent single-team CTAs (one instance per SM: a team of 256 threads),
 * larger ones on persistent CTAs of seven teams of 96 threads walking in lock-step.  A handle
 * serialises its solves (one work queue, one scratch arena); launches of DIFFERENT handles on different streams overlap
 * on the device: the straggler tail of a single-wave batch then no longer idles the SMs (bench.py --pipeline). */
int cmpc_solve_batched(cmpc_handle h, int batch, const double* d_p, const double* d_lbg, const double* d_ubg,
                       double* d_x, double* d_lam_g, double* d_obj, int* d_status, int* d_iters, int warm_duals,
                       void* stream);

/* Same solve with HOST pointers: copies p/lbg/ubg/x (and lam_g when warm_duals) to the device, solves, copies
 * x/lam_g/obj/status/iters back and synchronises.  This is the call a CentroidalMPC host object makes per tick.
 * The host-pointer entry points (this one and cmpc_solve_ticks_host) run on a private non-blocking stream of the handle and
 * return when that stream has drained: calls on DIFFERENT handles from different host threads overlap on the device (the
 * copies of one with the solve of the other, and the straggler tail of one batch with the start of the next: two handles
 * fed alternately deliver ~10 % more solves per second than one, bench.py "pipeline").  Pinned host buffers are needed
 * for the copies to be asynchronous. */
int cmpc_solve_host(cmpc_handle h, int batch, const double* p, const double* lbg, const double* ubg, double* x,
                    double* lam_g, double* obj, int* status, int* iters, int warm_duals);

/* Warm-start shift between ticks (BLF is_warm_start_enabled): every knot-indexed block of x and lam_g moves one knot
 * towards the present, the last knot is repeated.  DEVICE pointers, in place.  d_lam_g may be null. */
int cmpc_shift_warmstart(cmpc_handle h, int batch, double* d_x, double* d_lam_g, void* stream);

/* ---- input population on the device: what setState / setReferenceTrajectory / setContactPhaseList do before a solve ------
 * (CentroidalMPCBlock.cpp:407, :579, :609).  A TICK RECORD per instance holds the state, the external wrench, the references
 * and, per foot, the contacts the horizon can see (times in nanoseconds relative to the instance's current time):
 *      [0..8] com, dcom, angular momentum   [9..14] external force, torque   [15] 1 = step adjustment enabled   [16] reserved
 *      [17 ..) CoM reference 3 (N + 1), angular-momentum reference 3 (N + 1)
 *      per foot c at 17 + 6 (N + 1) + 85 c: [0] number of contacts n <= 6, then n x {t_on, t_off, position[3], rotation[9] column
 *      major} in time order (window: the last contact activated at or before now, every later one up to the first that starts
 *      after the horizon).  cmpc_tick_stride(N) = 6 N + 194 doubles.
 * cmpc_populate expands records into the solver's formal input (same rules as the host operator; DEVICE pointers, d_x0 may be
 * null).  cmpc_solve_ticks_host is the per-tick call of a host controller: HOST pointers; uploads the records (2.3 KB per
 * instance instead of 24.7 KB of p / lbg / ubg / x0), populates, solves, downloads x (and lam_g when not null), obj, status,
 * iters, and keeps the solution resident on the device.  warm_mode: 0 = cold start from the populated x0; 1 = warm start from
 * the solution the previous call on this handle left on the device (same batch size), shifted by one knot; 2 = warm start from
 * the x / lam_g passed in (the previous solution, NOT shifted: the shift runs on the device). */
int cmpc_tick_stride(int horizon);
int cmpc_populate(cmpc_handle h, int batch, const double* d_ticks, double* d_p, double* d_lbg, double* d_ubg, double* d_x0,
                  void* stream);
int cmpc_solve_ticks_host(cmpc_handle h, int batch, const double* ticks, int warm_mode, double* x, double* lam_g, double* obj,
                          int* status, int* iters);

/* Reference resampling on the device (the Math::LinearSpline frequency adapters of CentroidalMPCBlock.cpp:201-260, 525-577):
 * n_in planner samples of CoM / angular momentum per instance (d_com_in, d_h_in: batch x n_in x 3) at the shared times d_t_in
 * (increasing) -> the N + 1 reference knots at d_t_out (ordered), written into the tick records.  The angular momentum is
 * divided by robot_mass (:525-529); com_height >= 0 overrides the CoM height (:531-535, 0.7 in the reference). */
int cmpc_resample_references(cmpc_handle h, int batch, int n_in, const double* d_t_in, const double* d_com_in,
                             const double* d_h_in, const double* d_t_out, double robot_mass, double com_height, double* d_ticks,
                             void* stream);

/* Desired ZMP of the knot-0 corner forces (computeDesiredZMP, WholeBodyQPBlock.cpp:805-873; the reference clamps the local ZMP
 * to half_length 0.08 / half_width 0.03, :837-838).  d_zmp[batch][2]; d_valid[batch] (may be null) = 0 where no contact
 * carries force.  DEVICE pointers. */
int cmpc_desired_zmp(cmpc_handle h, int batch, const double* d_x, const double* d_p, double half_length, double half_width,
                     double* d_zmp, int* d_valid, void* stream);

/* ---- closed loop, device side (BASELINE config 4): a synthetic planner (the walk schedule of SURVEY.md 8(d)) keeps one
 * footstep table per rollout; cmpc_rollout_tick writes the tick records of MPC tick `tick` (plant state, push, references,
 * contact windows with the MPC's own landings: updateContactPhaseList, CentroidalMPCBlock.cpp:32-110) and the external wrench
 * acting on the plant; cmpc_rollout_feedback accumulates the statistics and writes the landing position of a foot that touches
 * down into the table.  d_roll[batch][roll_stride]: phase0, push tick / length / force[3], converged ticks, iterations, max
 * CoM error, min CoM height, max ZMP excess, ticks done; d_steps[batch][2][max_steps][4] = x, y, z, yaw.  DEVICE pointers.
 * tick < 0: the tick index is read from the rollout records (cmpc_rollout_feedback counts it up), so that the kernels of one
 * tick can be captured into a CUDA graph once and replayed. */
typedef struct cmpc_walk_params {
    int ds_knots, ss_knots;                 /* double / single support duration in MPC knots                      */
    double step_length, com_height;
    double push_threshold;                  /* external forces below it are not reported to the MPC (0.7)           */
    double zmp_half_length, zmp_half_width; /* clamp of the local ZMP (0.08, 0.03)                                  */
} cmpc_walk_params;
int cmpc_rollout_layout(int* roll_stride, int* max_steps, int* step_stride);
int cmpc_rollout_tick(cmpc_handle h, const cmpc_walk_params* w, int batch, int tick, const double* d_roll, const double* d_state,
                      const double* d_steps, double* d_ticks, double* d_ext6, int step_adjust, void* stream);
int cmpc_rollout_feedback(cmpc_handle h, const cmpc_walk_params* w, int batch, int tick, const double* d_x, const double* d_p,
                          const double* d_state, const int* d_status, const int* d_iters, double* d_roll, double* d_steps,
                          void* stream);

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
