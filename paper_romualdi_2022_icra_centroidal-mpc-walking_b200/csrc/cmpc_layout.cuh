// cmpc_layout.cuh -- index layout of x / p / g at the C-ABI boundary ("CasADi order") and stage-local indices.
//
// Layout decoded from the reference's generated NLP code:
//   /root/reference/src/centroidal-mpc-walking/config/robots/ergoCubGazeboV1/tmp.c:62-67 (casadi_s0..s5),
//   nlp:(x[45N+15], p[50N+27]) -> (f, g[53N+15])  (tmp.c:69).
#pragma once

#if defined(__CUDACC__)
#define CMPC_HD __host__ __device__ __forceinline__
#else
#define CMPC_HD inline
#endif

namespace cmpc {

constexpr int NC = 2;   // contacts: 0 left_foot, 1 right_foot
constexpr int NJ = 4;   // corners per contact
constexpr int NF = 4;   // friction half planes per corner (number_of_slices = 1)
constexpr int NS = 15;  // stage state   s_k = (com, dcom, h, pos_L, pos_R)
constexpr int NU = 30;  // stage control u_k = (vel_L, vel_R, f_L0..f_L3, f_R0..f_R3)
constexpr int NPHI = 24;            // previous-knot forces carried as extra state (force-rate cost)
constexpr int NXI = NS + NPHI;      // 39: augmented state of the Riccati recursion
constexpr int ROWS_PER_KNOT = 53;   // 15 dynamics + 6 step box + 32 friction rows per knot
constexpr int INEQ_PER_KNOT = 38;

CMPC_HD int dim_x(int N) { return 45 * N + 15; }
CMPC_HD int dim_p(int N) { return 50 * N + 27; }
CMPC_HD int dim_g(int N) { return 53 * N + 15; }
CMPC_HD int nnz_jac(int N) { return 243 * N + 15; }
CMPC_HD int nnz_hess(int N) { return 348 * N - 36; }

// ---- x
CMPC_HD int x_com(int N, int k) { (void)N; return 3 * k; }
CMPC_HD int x_dcom(int N, int k) { return 3 * (N + 1) + 3 * k; }
CMPC_HD int x_h(int N, int k) { return 6 * (N + 1) + 3 * k; }
CMPC_HD int x_cbase(int N, int c) { return 9 * (N + 1) + c * (18 * N + 3); }
CMPC_HD int x_pos(int N, int c, int k) { return x_cbase(N, c) + 3 * k; }
CMPC_HD int x_vel(int N, int c, int k) { return x_cbase(N, c) + 3 * (N + 1) + 3 * k; }
CMPC_HD int x_frc(int N, int c, int j, int k) { return x_cbase(N, c) + 6 * N + 3 + 3 * N * j + 3 * k; }
// stage-local index (0..14) of s_k  -> x index
CMPC_HD int x_of_s(int N, int k, int i)
{
    return i < 3 ? x_com(N, k) + i : i < 6 ? x_dcom(N, k) + i - 3 : i < 9 ? x_h(N, k) + i - 6
         : i < 12 ? x_pos(N, 0, k) + i - 9 : x_pos(N, 1, k) + i - 12;
}
// stage-local index (0..29) of u_k -> x index
CMPC_HD int x_of_u(int N, int k, int i)
{
    if (i < 3) return x_vel(N, 0, k) + i;
    if (i < 6) return x_vel(N, 1, k) + i - 3;
    int f = i - 6, c = f / 12, j = (f % 12) / 3, a = f % 3;
    return x_frc(N, c, j, k) + a;
}

// ---- p
CMPC_HD int p_cbase(int N, int c) { return c * (19 * N + 6); }
CMPC_HD int p_rot(int N, int c, int k) { return p_cbase(N, c) + 9 * k; }  // vec(R_k) column major: R(r,col) = [3*col + r]
CMPC_HD int p_en(int N, int c, int k) { return p_cbase(N, c) + 15 * N + k; }
CMPC_HD int p_nom(int N, int c, int k) { return p_cbase(N, c) + 16 * N + 3 * k; }
CMPC_HD int p_glob(int N) { return 38 * N + 12; }
CMPC_HD int p_comref(int N, int k) { return p_glob(N) + 9 + 3 * k; }
CMPC_HD int p_href(int N, int k) { return p_glob(N) + 9 + 3 * (N + 1) + 3 * k; }
CMPC_HD int p_extf(int N, int k) { return p_glob(N) + 9 + 6 * (N + 1) + 3 * k; }
CMPC_HD int p_extt(int N, int k) { return p_glob(N) + 9 + 6 * (N + 1) + 3 * N + 3 * k; }

// ---- g
CMPC_HD int g_com(int N, int k) { (void)N; return 15 + 3 * k; }
CMPC_HD int g_dcom(int N, int k) { return 15 + 3 * N + 3 * k; }
CMPC_HD int g_h(int N, int k) { return 15 + 6 * N + 3 * k; }
CMPC_HD int g_pos(int N, int c, int k) { return 15 + 9 * N + 3 * N * c + 3 * k; }
CMPC_HD int g_box(int N, int c, int k) { return 15 + 15 * N + c * 19 * N + 3 * k; }
CMPC_HD int g_fric(int N, int c, int j, int k) { return 15 + 15 * N + c * 19 * N + 3 * N + 16 * k + 4 * j; }
// multiplier / residual row of the constraint that DEFINES component i of s_k:
//   k = 0: initial-condition rows 0..14;  k >= 1: dynamics rows of knot k-1
CMPC_HD int g_of_s(int N, int k, int i)
{
    if (k == 0) return i;
    int kk = k - 1;
    return i < 3 ? g_com(N, kk) + i : i < 6 ? g_dcom(N, kk) + i - 3 : i < 9 ? g_h(N, kk) + i - 6
         : i < 12 ? g_pos(N, 0, kk) + i - 9 : g_pos(N, 1, kk) + i - 12;
}
// path-inequality row l (0..37) of knot k: 0..5 step box (c = l/3, axis), 6..37 friction (c, j, r)
CMPC_HD int g_of_ineq(int N, int k, int l)
{
    if (l < 6) return g_box(N, l / 3, k) + l % 3;
    int f = l - 6;
    return g_fric(N, f / 16, (f % 16) / 4, k) + f % 4;
}

}  // namespace cmpc
