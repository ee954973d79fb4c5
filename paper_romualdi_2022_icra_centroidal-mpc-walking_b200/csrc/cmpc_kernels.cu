// cmpc_kernels.cu -- sm_100a kernels and the C ABI (include/cmpc_b200.h) of the batched centroidal-MPC solve.
//
// Kernel inventory (see DESIGN.md for the roofline of each):
//   cmpc_solve_team_kernel  persistent CTAs, one MPC instance per team of 32/64/128 threads at a time (atomic work queue):
//                         the whole interior-point solve incl. the per-knot Riccati factorisation in shared memory
//   cmpc_shift_kernel     warm-start shift of x / lam_g by one knot (HBM bound, coalesced, staged in shared memory)
//   cmpc_eval_kernel      f, grad f, g of the NLP (parity surface for nlp_fg / nlp_jac_fg)
//   cmpc_jac_kernel / cmpc_hess_kernel   entry-wise CSC jacobian / hessian (parity surface)
//   cmpc_plant_kernel     RK4 centroidal dynamics of the closed-loop plant
// No CPU fallback exists: every entry point needs a CUDA device and says so when there is none.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/cmpc_b200.h"
#include "cmpc_sparse.cuh"
#include "cmpc_ipm.cuh"
#include "cmpc_populate.cuh"

namespace cmpc {

#ifndef CMPC_REDUCE_INTERLEAVED
#define CMPC_REDUCE_INTERLEAVED 1
#endif
// ------------------------------------------------------------------------------------------------ device CTA context
struct DevCta {
    int tid, nt, warp, lane, wsize;
    int bar = 0;  // hardware barrier of the team (0 = the whole CTA)
    double* red;  // >= 32 doubles of shared memory
    __device__ __forceinline__ void bsync() { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nt) : "memory"); }
    __device__ __forceinline__ void sync() { bsync(); }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }

    template <int K, class Op>
    __device__ __forceinline__ void allreduce(double* v, Op op)
    {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = v[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x = op(x, __shfl_xor_sync(0xffffffffu, x, off));
            v[k] = x;
        }
        const int nw = nt >> 5;
        if (lane == 0)
            for (int k = 0; k < K; ++k) red[warp * K + k] = v[k];
        bsync();
        for (int k = 0; k < K; ++k) {
            double x = red[k];
            for (int w2 = 1; w2 < nw; ++w2) x = op(x, red[w2 * K + k]);
            v[k] = x;
        }
        bsync();
    }
    template <int K> __device__ __forceinline__ void sumv(double* v) { allreduce<K>(v, [](double a, double b) { return a + b; }); }
    template <int K> __device__ __forceinline__ void maxv(double* v) { allreduce<K>(v, [](double a, double b) { return fmax(a, b); }); }
    template <int K> __device__ __forceinline__ void minv(double* v) { allreduce<K>(v, [](double a, double b) { return fmin(a, b); }); }
    __device__ __forceinline__ double sum(double x) { sumv<1>(&x); return x; }
    __device__ __forceinline__ double max(double x) { maxv<1>(&x); return x; }
    // KM maxima, KN minima and KS sums in ONE exchange through shared memory (KM + KN + KS <= 8, at most 8 warps)
    template <int KM, int KN, int KS>
    __device__ __forceinline__ void reduce3(double* vmax, double* vmin, double* vsum)
    {
        constexpr int K = KM + KN + KS;
        const int nw = nt >> 5;
#if CMPC_REDUCE_INTERLEAVED
        // the K butterflies side by side: 5 rounds of K independent shuffles instead of K chains of 5 dependent ones
        double x[K];
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = k < KM ? vmax[k] : (k < KM + KN ? vmin[k - KM] : vsum[k - KM - KN]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double y[K];
#pragma unroll
            for (int k = 0; k < K; ++k) y[k] = __shfl_xor_sync(0xffffffffu, x[k], off);
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = k < KM ? fmax(x[k], y[k]) : (k < KM + KN ? fmin(x[k], y[k]) : x[k] + y[k]);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) red[warp * K + k] = x[k];
        }
#else
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int op = k < KM ? 0 : (k < KM + KN ? 1 : 2);
            double* src = k < KM ? vmax + k : (k < KM + KN ? vmin + (k - KM) : vsum + (k - KM - KN));
            const double x = warp_reduce_op(*src, op);
            if (lane == 0) red[warp * K + k] = x;
        }
#endif
        bsync();
#if CMPC_REDUCE_INTERLEAVED
        {   // the K statistics of all warps: independent loads first, then K short combines
            double z[K];
#pragma unroll
            for (int k = 0; k < K; ++k) z[k] = red[k];
#pragma unroll 1
            for (int w2 = 1; w2 < nw; ++w2) {
                double y[K];
#pragma unroll
                for (int k = 0; k < K; ++k) y[k] = red[w2 * K + k];
#pragma unroll
                for (int k = 0; k < K; ++k) z[k] = k < KM ? fmax(z[k], y[k]) : (k < KM + KN ? fmin(z[k], y[k]) : z[k] + y[k]);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < KM) vmax[k] = z[k];
                else if (k < KM + KN) vmin[k - KM] = z[k];
                else vsum[k - KM - KN] = z[k];
            }
        }
#else
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int op = k < KM ? 0 : (k < KM + KN ? 1 : 2);
            double* dst = k < KM ? vmax + k : (k < KM + KN ? vmin + (k - KM) : vsum + (k - KM - KN));
            *dst = combine_warps(red + k, K, nw, op);
        }
#endif
        bsync();
    }
    // statistic k of all warps (one copy of the code: the passes are bound by their instruction footprint)
    static __device__ __noinline__ double combine_warps(const double* red, int K, int nw, int op)
    {
        double x = red[0];
#pragma unroll 1
        for (int w2 = 1; w2 < nw; ++w2) x = combine_op(x, red[w2 * K], op);
        return x;
    }
    static __device__ __forceinline__ double combine_op(double x, double y, int op)
    {
        return op == 0 ? (x > y ? x : y) : (op == 1 ? (x < y ? x : y) : x + y);
    }
    // butterfly over the warp with a run-time operation: one copy of the code for every statistic of every pass
    static __device__ __noinline__ double warp_reduce_op(double x, int op)
    {
#pragma unroll 1
        for (int off = 16; off > 0; off >>= 1) x = combine_op(x, __shfl_xor_sync(0xffffffffu, x, off), op);
        return x;
    }
};

__device__ __forceinline__ DevCta make_cta(double* red)
{
    DevCta c;
    c.tid = threadIdx.x; c.nt = blockDim.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31; c.wsize = 32;
    c.red = red;
    return c;
}

// one warp = one "CTA" of the solver: barriers are __syncwarp, reductions are shuffles (no shared memory, no block barrier)
struct DevWarp {
    int tid, nt, warp, lane, wsize;
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    template <int K, class Op>
    __device__ __forceinline__ void allreduce(double* v, Op op)
    {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = v[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x = op(x, __shfl_xor_sync(0xffffffffu, x, off));
            v[k] = x;
        }
    }
    template <int K> __device__ __forceinline__ void sumv(double* v) { allreduce<K>(v, [](double a, double b) { return a + b; }); }
    template <int K> __device__ __forceinline__ void maxv(double* v) { allreduce<K>(v, [](double a, double b) { return fmax(a, b); }); }
    template <int K> __device__ __forceinline__ void minv(double* v) { allreduce<K>(v, [](double a, double b) { return fmin(a, b); }); }
    __device__ __forceinline__ double sum(double x) { sumv<1>(&x); return x; }
    __device__ __forceinline__ double max(double x) { maxv<1>(&x); return x; }
    template <int KM, int KN, int KS>
    __device__ __forceinline__ void reduce3(double* vmax, double* vmin, double* vsum)
    {
        if (KM > 0) maxv<KM == 0 ? 1 : KM>(vmax);
        if (KN > 0) minv<KN == 0 ? 1 : KN>(vmin);
        if (KS > 0) sumv<KS == 0 ? 1 : KS>(vsum);
    }
};

// ------------------------------------------------------------------------------------------------ kernels
// THE hot path: persistent CTAs of NT threads (a "team": 1, 2 or 4 warps), one MPC instance per team at a time (atomic work
// queue).  The whole interior-point solve of the instance runs inside the team: Riccati factorisation on 3 x 3 tiles in
// shared memory (cmpc_warp.cuh), iterate vectors in a per-team scratch block that stays in L2.
template <int NT, int G> struct TeamCta { using type = DevCta; };
template <> struct TeamCta<32, 1> { using type = DevWarp; };

// G teams of NT threads per CTA, walking through the phases of their solves in lock-step (cmpc_ipm.cuh: ipm_run)
// Register budget: registers are allocated to groups of 4 warps, so the 21 warps of the seven-team CTA occupy 24 warp slots
// and the budget is 65536 / 768 = 85 -> 80 registers per thread (what ptxas derives from __launch_bounds__(672, 1); a kernel
// built with __maxnreg__(96) does not launch).  An eighth team of 96 threads would be free in registers.
template <int NT, int G, int CTAS>
__global__ void __launch_bounds__(NT * G, CTAS)
cmpc_solve_team_kernel(const __grid_constant__ Config cfg, int batch, const double* __restrict__ p,
                       const double* __restrict__ lbg, const double* __restrict__ ubg, double* x, double* lam, double* obj,
                       int* status, int* iters, int warm_duals, double* work, size_t work_stride, unsigned int* counter,
                       const unsigned short* __restrict__ cmap, int ngroups, double* ric, size_t ric_stride)
{
    extern __shared__ __align__(16) double smem_raw[];
    const int team = threadIdx.x / NT;
    Team T;
    T.lane = threadIdx.x - team * NT;
    T.id = team;
    T.on = true;
    {   // lock-step groups: team t belongs to group t * ngroups / G (4 + 3 teams for G = 7, ngroups = 2); barriers 8 + group
        const int g = team * ngroups / G;
        int cnt = 0;
        for (int t = 0; t < G; ++t) cnt += (t * ngroups / G == g) ? 1 : 0;
        T.gbar = 8 + g;
        T.gcount = cnt * NT;
    }
    ISmem& sm = reinterpret_cast<ISmem*>(smem_raw)[team];
    typename TeamCta<NT, G>::type cta;
    cta.tid = T.lane; cta.nt = NT; cta.warp = T.lane >> 5; cta.lane = T.lane & 31; cta.wsize = 32;
    if constexpr (!(NT == 32 && G == 1)) { cta.red = sm.sw.red; cta.bar = G > 1 ? team + 1 : 0; }
    const size_t slot = (size_t)blockIdx.x * G + team;
    double* base = work + slot * work_stride;
    ipm_run<NT, G>(T, cta, cfg, base, sm, cmap, batch, p, lbg, ubg, x, lam, obj, status, iters, warm_duals, counter,
                   team * (int)gridDim.x + (int)blockIdx.x, G * (int)gridDim.x, ric_stride ? ric + slot * ric_stride : nullptr);
}

// warm-start shift: one CTA per instance, the vector is staged in shared memory so that loads and stores are coalesced
CMPC_HD int shift_src_x(int N, int i)
{
    const int nb = 9 * (N + 1);
    if (i < nb) { int blk = i / (3 * (N + 1)), r = i % (3 * (N + 1)), k = r / 3; return k < N ? i + 3 : i; (void)blk; }
    int r = (i - nb) % (18 * N + 3);
    if (r < 3 * (N + 1)) return r / 3 < N ? i + 3 : i;
    r -= 3 * (N + 1);
    return (r % (3 * N)) / 3 < N - 1 ? i + 3 : i;  // vel, forces: N columns
}
CMPC_HD int shift_src_g(int N, int r)
{
    if (r < NS) return r;  // initial-condition multipliers are recomputed by the next solve
    int q = r - NS;
    if (q < 15 * N) return (q % (3 * N)) / 3 < N - 1 ? r + 3 : r;
    q = (q - 15 * N) % (19 * N);
    if (q < 3 * N) return q / 3 < N - 1 ? r + 3 : r;
    q -= 3 * N;
    return q / 16 < N - 1 ? r + 16 : r;
}
__global__ void cmpc_shift_kernel(int N, int batch, double* x, double* lam)
{
    extern __shared__ double buf[];
    const int n = dim_x(N), m = dim_g(N);
    for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        double* xi = x + (size_t)inst * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = xi[i];
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) xi[i] = buf[shift_src_x(N, i)];
        __syncthreads();
        if (lam) {
            double* li = lam + (size_t)inst * m;
            for (int i = threadIdx.x; i < m; i += blockDim.x) buf[i] = li[i];
            __syncthreads();
            for (int i = threadIdx.x; i < m; i += blockDim.x) li[i] = buf[shift_src_g(N, i)];
            __syncthreads();
        }
    }
}

// f, grad f, g: one CTA per instance with the solver's own evaluation routines
__global__ void cmpc_eval_kernel(Config cfg, int batch, const double* __restrict__ x, const double* __restrict__ p,
                                 double* f, double* grad, double* g, double* gscratch)
{
    extern __shared__ double sd[];  // N * SD_STRIDE stage data + 64 reduction scratch
    const int N = cfg.N, n = dim_x(N), np = dim_p(N), m = dim_g(N);
    DevCta cta = make_cta(sd + N * SD_STRIDE);
    for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        Instance in{p + (size_t)inst * np, nullptr, nullptr};
        const double* xi = x + (size_t)inst * n;
        stage_data(cta, cfg, in, xi, sd);
        double* gi = g ? g + (size_t)inst * m : gscratch + (size_t)blockIdx.x * m;
        eval_g(cta, cfg, in, xi, sd, gi);
        double fv = eval_f(cta, cfg, in, xi, grad ? grad + (size_t)inst * n : (double*)nullptr);
        if (grad) {  // dcom carries no cost and is not written by eval_f
            for (int it = threadIdx.x; it < 3 * (N + 1); it += blockDim.x) grad[(size_t)inst * n + x_dcom(N, 0) + it] = 0.0;
        }
        if (f && threadIdx.x == 0) f[inst] = fv;
        __syncthreads();
    }
}

__global__ void cmpc_jac_kernel(Config cfg, int batch, const double* __restrict__ x, const double* __restrict__ p,
                                const int* __restrict__ slot, double* jnz)
{
    const int N = cfg.N, n = dim_x(N), np = dim_p(N), nnz = nnz_jac(N);
    const long long total = (long long)batch * nnz;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int inst = (int)(t / nnz), e = (int)(t % nnz), row, col;
        double v = jac_entry(cfg, x + (size_t)inst * n, p + (size_t)inst * np, e, row, col);
        jnz[(size_t)inst * nnz + slot[e]] = v;
    }
}

__global__ void cmpc_hess_kernel(Config cfg, int batch, const double* __restrict__ p, double lam_f,
                                 const double* __restrict__ lam_g, const int* __restrict__ slot, double* hnz)
{
    const int N = cfg.N, np = dim_p(N), m = dim_g(N), nnz = nnz_hess(N);
    const long long total = (long long)batch * nnz;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int inst = (int)(t / nnz), e = (int)(t % nnz), row, col;
        double v = hess_entry(cfg, p + (size_t)inst * np, lam_f, lam_g + (size_t)inst * m, e, row, col);
        hnz[(size_t)inst * nnz + slot[e]] = v;
    }
}

// closed-loop plant: RK4 of  d(com) = dcom, d(dcom) = g + ext_f + sum en f, d(h) = ext_t + sum en (R r_j + pos - com) x f
// (BLF CentroidalDynamics with unit mass, WholeBodyQPBlock.cpp:1039-1041, 1083-1090, 1150-1158); forces held constant
__global__ void cmpc_plant_kernel(Config cfg, int batch, const double* __restrict__ x, const double* __restrict__ p,
                                  const double* __restrict__ ext, double* state, double dt, int substeps)
{
    const int N = cfg.N, n = dim_x(N), np = dim_p(N);
    for (int inst = blockIdx.x * blockDim.x + threadIdx.x; inst < batch; inst += gridDim.x * blockDim.x) {
        const double* xi = x + (size_t)inst * n;
        const double* pi = p + (size_t)inst * np;
        double F[3] = {0, 0, GRAV_Z}, T0[3] = {0, 0, 0};  // dh = T0 - com x Fc  with  Fc = sum of contact forces
        double Fc[3] = {0, 0, 0};
        if (ext) {
            for (int a = 0; a < 3; ++a) { F[a] += ext[(size_t)inst * 6 + a]; T0[a] += ext[(size_t)inst * 6 + 3 + a]; }
        }
        for (int c = 0; c < NC; ++c) {
            double en = pi[p_en(N, c, 0)];
            const double* R = pi + p_rot(N, c, 0);
            for (int j = 0; j < NJ; ++j) {
                const double* cr = cfg.corner[c][j];
                double arm[3], f[3], t[3];
                for (int a = 0; a < 3; ++a) {
                    arm[a] = R[a] * cr[0] + R[3 + a] * cr[1] + R[6 + a] * cr[2] + xi[x_pos(N, c, 0) + a];
                    f[a] = en * xi[x_frc(N, c, j, 0) + a];
                }
                cross3(arm, f, t);
                for (int a = 0; a < 3; ++a) { Fc[a] += f[a]; T0[a] += t[a]; }
            }
        }
        for (int a = 0; a < 3; ++a) F[a] += Fc[a];
        double* s = state + (size_t)inst * 9;
        double com[3] = {s[0], s[1], s[2]}, dcom[3] = {s[3], s[4], s[5]}, h[3] = {s[6], s[7], s[8]};
        for (int it = 0; it < substeps; ++it) {
            // the vector field is affine in (com, dcom): RK4 written out
            double k1c[3], k2c[3], k3c[3], k4c[3], k1h[3], k2h[3], k3h[3], k4h[3], tmp[3], cx[3];
            for (int a = 0; a < 3; ++a) k1c[a] = dcom[a];
            cross3(com, Fc, cx); for (int a = 0; a < 3; ++a) k1h[a] = T0[a] - cx[a];
            for (int a = 0; a < 3; ++a) { k2c[a] = dcom[a] + 0.5 * dt * F[a]; tmp[a] = com[a] + 0.5 * dt * k1c[a]; }
            cross3(tmp, Fc, cx); for (int a = 0; a < 3; ++a) k2h[a] = T0[a] - cx[a];
            for (int a = 0; a < 3; ++a) { k3c[a] = dcom[a] + 0.5 * dt * F[a]; tmp[a] = com[a] + 0.5 * dt * k2c[a]; }
            cross3(tmp, Fc, cx); for (int a = 0; a < 3; ++a) k3h[a] = T0[a] - cx[a];
            for (int a = 0; a < 3; ++a) { k4c[a] = dcom[a] + dt * F[a]; tmp[a] = com[a] + dt * k3c[a]; }
            cross3(tmp, Fc, cx); for (int a = 0; a < 3; ++a) k4h[a] = T0[a] - cx[a];
            for (int a = 0; a < 3; ++a) {
                com[a] += dt / 6.0 * (k1c[a] + 2.0 * k2c[a] + 2.0 * k3c[a] + k4c[a]);
                h[a] += dt / 6.0 * (k1h[a] + 2.0 * k2h[a] + 2.0 * k3h[a] + k4h[a]);
                dcom[a] += dt * F[a];
            }
        }
        for (int a = 0; a < 3; ++a) { s[a] = com[a]; s[3 + a] = dcom[a]; s[6 + a] = h[a]; }
    }
}

// FP64 FMA throughput probe: 8 independent DFMA chains per thread, no memory traffic (roofline denominator of bench.py)
__global__ void cmpc_dfma_probe_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 123.456) out[0] = x0;  // keeps the chains alive, never true in practice
}

// ------------------------------------------------------------------------------------------------ host side
static void friction_matrix_host(double mu, double A[NF][3])
{
    // BLF Math::LinearizedFrictionCone with one slice per quadrant: the unit circle is sampled every pi/2, each closed
    // segment gives the half plane  -s*slope*fx + s*fy - s*offset*mu*fz <= 0  (s = -1 below the x axis)
    const double kPi = 3.14159265358979323846;
    const int nrows = NF;
    const double seg = kPi / 2.0;
    for (int i = 0; i < nrows; ++i) {
        double a0 = seg * i, a1 = seg * (i + 1), a1w = seg * ((i + 1) % nrows);
        double x0 = cos(a0), y0 = sin(a0), x1 = cos(a1w), y1 = sin(a1w);
        double slope = (y1 - y0) / (x1 - x0), offset = y0 - slope * x0;
        double s = (a0 > kPi || a1 > kPi) ? -1.0 : 1.0;
        A[i][0] = -s * slope; A[i][1] = s; A[i][2] = -s * offset * mu;
    }
}

struct Csc {
    std::vector<int> colind, row, slot;
};
static void build_csc(int N, bool hess, Csc& out)
{
    Config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.N = N;
    const int ne = hess ? hess_emissions(N) : NS + JAC_PER_KNOT * N;
    const int ncol = dim_x(N);
    struct T { int r, c, e; };
    std::vector<T> t(ne);
    for (int e = 0; e < ne; ++e) {
        int r, c;
        if (hess) hess_entry(cfg, nullptr, 0.0, nullptr, e, r, c);
        else jac_entry(cfg, nullptr, nullptr, e, r, c);
        t[e] = {r, c, e};
    }
    std::sort(t.begin(), t.end(), [](const T& a, const T& b) { return a.c != b.c ? a.c < b.c : (a.r != b.r ? a.r < b.r : a.e < b.e); });
    out.colind.assign(ncol + 1, 0);
    out.row.resize(ne);
    out.slot.resize(ne);
    for (int i = 0; i < ne; ++i) { out.row[i] = t[i].r; out.slot[t[i].e] = i; out.colind[t[i].c + 1]++; }
    for (int c = 0; c < ncol; ++c) out.colind[c + 1] += out.colind[c];
}

}  // namespace cmpc

using namespace cmpc;

#ifndef CMPC_DEFAULT_TEAM
#define CMPC_DEFAULT_TEAM 96
#endif
#ifndef CMPC_CTAS_PER_SM
#define CMPC_CTAS_PER_SM 7
#endif
#ifndef CMPC_L2_WINDOW
#define CMPC_L2_WINDOW 0   // 1: persisting-L2 access window over the iterate vectors of the resident teams.  Measured (B200: 79 MB of
                           // set-aside, 128 MB window): 74.0 k -> 71.8 k solves/s with the full set-aside, unchanged with half of it or
                           // with 40 % (profiles/r2_notes.md): off
#endif
#ifndef CMPC_SPLIT_RIC
#define CMPC_SPLIT_RIC 1   // factor blocks of all teams in their own region of the arena (needed by the window; a hair faster: 90.0 k
                           // against 89.9 k solves/s over three alternating runs each)
#endif
#ifndef CMPC_L2_MISS_NORMAL
#define CMPC_L2_MISS_NORMAL 0
#endif
#ifndef CMPC_L2_FRACTION
#define CMPC_L2_FRACTION 1.0
#endif
#ifndef CMPC_DEFAULT_GROUPS
#define CMPC_DEFAULT_GROUPS 3     // independent lock-step groups per CTA (3 + 2 + 2 teams)
#endif
#ifndef CMPC_DEFAULT_LOCKSTEP
#define CMPC_DEFAULT_LOCKSTEP 7   // teams per CTA of the default (team 96) configuration; 1 = independent teams
#endif
static_assert(sizeof(ISmem) % 16 == 0, "teams of a CTA are laid out back to back in shared memory");
// (team size, teams per CTA) -> kernel
#define CMPC_FOR_EACH_KERNEL(X) X(32, 1, CMPC_CTAS_PER_SM) X(64, 1, CMPC_CTAS_PER_SM) X(96, 1, CMPC_CTAS_PER_SM) \
    X(128, 1, CMPC_CTAS_PER_SM) X(96, 7, 1) X(96, 3, 2) X(64, 7, 1) X(128, 7, 1) X(192, 1, 1) X(256, 1, 1)
static const void* team_kernel(int nt, int g)
{
#define X(NT, G, C) if (nt == NT && g == G) return (const void*)cmpc_solve_team_kernel<NT, G, C>;
    CMPC_FOR_EACH_KERNEL(X)
#undef X
    return nullptr;
}

struct cmpc_handle_s {
    Config cfg;
    cmpc_config user;
    int groups = 1;  // independent lock-step groups inside a CTA
    bool latency_path = false;  // default geometry: batches of at most one instance per SM go to the single-team kernel
    int device = 0, sm_count = 0, lockstep = 1, threads = CMPC_DEFAULT_TEAM, ctas_per_sm = 0, grid = 0, smem = 0;
    int shift_smem = 0, eval_smem = 0;  // dynamic shared memory of the auxiliary kernels
    size_t work_stride = 0, work_slots = 0, ric_stride = 0;
    double* d_work = nullptr;   // [slots x work_stride] iterate vectors of every team, then [slots x ric_stride] factor blocks (d_ric)
    double* d_ric = nullptr;
    size_t l2_window_bytes = 0; // bytes of the vector region covered by the persisting-L2 access window (0: none)
    float l2_hit_ratio = 0.f;
    unsigned int* d_counter = nullptr;
    int *d_jslot = nullptr, *d_hslot = nullptr;
    unsigned short* d_cmap = nullptr;
    double* d_gscratch = nullptr;
    // staging buffers of cmpc_solve_host
    int host_cap = 0;
    double *d_p = nullptr, *d_lbg = nullptr, *d_ubg = nullptr, *d_x = nullptr, *d_lam = nullptr, *d_obj = nullptr;
    double* d_ticks = nullptr;
    int resident_batch = 0;  // batch size of the solution the last host call left in d_x / d_lam (warm_mode 1)
    int *d_status = nullptr, *d_iters = nullptr;
    long long launches = 0;
    int last_cuda = 0;
    // a handle has ONE work queue and ONE scratch arena: solves on it are serialised on the device.  `done` is recorded
    // behind every solve; a solve enqueued on another stream than the previous one waits for it first.
    cudaEvent_t done = nullptr;
    cudaStream_t last_stream = nullptr;
    // the host-pointer entry points (cmpc_solve_host, cmpc_solve_ticks_host) run on a private non-blocking stream of the handle:
    // two handles driven from two host threads overlap their copies and solves instead of meeting on the legacy default stream
    cudaStream_t hstream = nullptr;
    bool in_flight = false;
};

// every entry point runs on the handle's device and leaves the caller's current device as it found it
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
        else if (err == cudaSuccess) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define CMPC_ON_DEVICE(h) DeviceGuard guard_((h)->device); CK(guard_.err)

#define CK(call)                                           \
    do {                                                   \
        cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) { h->last_cuda = (int)e_; return CMPC_E_CUDA; } \
    } while (0)

extern "C" {

int cmpc_default_config(cmpc_config* c)
{
    if (!c) return CMPC_E_INVALID;
    memset(c, 0, sizeof *c);
    c->horizon = 12; c->sampling_time = 0.1; c->number_of_slices = 1; c->static_friction_coefficient = 0.33;
    c->com_weight[0] = 10; c->com_weight[1] = 10; c->com_weight[2] = 200;
    c->contact_position_weight = 2e3;
    for (int a = 0; a < 3; ++a) c->force_rate_of_change_weight[a] = 10;
    c->angular_momentum_weight = 1e2; c->contact_force_symmetry_weight = 10;
    const double cr[4][3] = {{0.08, 0.01, 0}, {0.08, -0.01, 0}, {-0.08, -0.01, 0}, {-0.08, 0.01, 0}};
    for (int k = 0; k < 2; ++k) memcpy(c->corners[k], cr, sizeof cr);
    c->ipopt_tolerance = 1e-8; c->ipopt_max_iteration = 200; c->mu_init = 0.1; c->bound_relax_factor = 1e-8;
    c->bound_push = 0.01; c->infinity = 1e19; c->device = 0; c->threads_per_instance = 0; c->ctas_per_sm = 0;
    c->mu_strategy = CMPC_MU_DEFAULT;
    c->warm_start_mu_init = 0.01;
    c->nlp_scaling_max_gradient = 100.0; c->acceptable_tol = 1e-6; c->acceptable_iter = 15;
    const double up[2][3] = {{0.01, 0.05, 0.0}, {0.01, 0.0, 0.0}}, lo[2][3] = {{-0.01, 0.0, 0.0}, {-0.01, -0.05, 0.0}};
    memcpy(c->bounding_box_upper_limit, up, sizeof up);
    memcpy(c->bounding_box_lower_limit, lo, sizeof lo);
    return CMPC_OK;
}

int cmpc_dims(int N, int* n, int* np, int* m, int* nj, int* nh)
{
    if (N < 1) return CMPC_E_INVALID;
    if (n) *n = dim_x(N);
    if (np) *np = dim_p(N);
    if (m) *m = dim_g(N);
    if (nj) *nj = nnz_jac(N);
    if (nh) *nh = nnz_hess(N);
    return CMPC_OK;
}

int cmpc_friction_matrix(double mu, int slices, double* A)
{
    if (!A || slices != 1) return CMPC_E_INVALID;
    double M[NF][3];
    friction_matrix_host(mu, M);
    memcpy(A, M, sizeof M);
    return CMPC_OK;
}

int cmpc_jac_sparsity(int N, int* colind, int* row)
{
    if (N < 1 || !colind || !row) return CMPC_E_INVALID;
    Csc c; build_csc(N, false, c);
    memcpy(colind, c.colind.data(), sizeof(int) * c.colind.size());
    memcpy(row, c.row.data(), sizeof(int) * c.row.size());
    return CMPC_OK;
}
int cmpc_hess_sparsity(int N, int* colind, int* row)
{
    if (N < 1 || !colind || !row) return CMPC_E_INVALID;
    Csc c; build_csc(N, true, c);
    memcpy(colind, c.colind.data(), sizeof(int) * c.colind.size());
    memcpy(row, c.row.data(), sizeof(int) * c.row.size());
    return CMPC_OK;
}

int cmpc_create(const cmpc_config* u, cmpc_handle* out)
{
    if (!u || !out) return CMPC_E_INVALID;
    *out = nullptr;
    if (u->horizon < 2 || u->horizon > 256 || !(u->sampling_time > 0) || u->number_of_slices != 1) return CMPC_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return CMPC_E_NO_DEVICE;
    if (u->device < 0 || u->device >= ndev) return CMPC_E_INVALID;
    cmpc_handle h = new (std::nothrow) cmpc_handle_s;
    if (!h) return CMPC_E_ALLOC;
    h->user = *u;
    Config& c = h->cfg;
    memset(&c, 0, sizeof c);
    c.N = u->horizon; c.dT = u->sampling_time;
    for (int a = 0; a < 3; ++a) { c.w_com[a] = u->com_weight[a]; c.w_rate[a] = u->force_rate_of_change_weight[a]; }
    c.w_h = u->angular_momentum_weight; c.w_pos = u->contact_position_weight; c.w_sym = u->contact_force_symmetry_weight;
    memcpy(c.corner, u->corners, sizeof c.corner);
    friction_matrix_host(u->static_friction_coefficient, c.fricA);
    c.tol = u->ipopt_tolerance > 0 ? u->ipopt_tolerance : 1e-8;
    c.max_iter = u->ipopt_max_iteration > 0 ? u->ipopt_max_iteration : 200;
    c.mu_init = u->mu_init > 0 ? u->mu_init : 0.1;
    c.mu_warm = u->warm_start_mu_init > 0 ? u->warm_start_mu_init : 0.01;
    c.bound_relax = u->bound_relax_factor >= 0 ? u->bound_relax_factor : 1e-8;
    c.bound_push = u->bound_push > 0 ? u->bound_push : 0.01;
    c.inf_bound = u->infinity > 0 ? u->infinity : 1e19;
    if (u->mu_strategy < 0 || u->mu_strategy > CMPC_MU_MEHROTRA) { delete h; return CMPC_E_INVALID; }
    c.pc = u->mu_strategy == CMPC_MU_MONOTONE ? 0 : 1;
    c.scal_max_grad = u->nlp_scaling_max_gradient == 0.0 ? 100.0 : (u->nlp_scaling_max_gradient > 0.0 ? u->nlp_scaling_max_gradient : 0.0);
    c.acc_tol = u->acceptable_tol == 0.0 ? 1e-6 : (u->acceptable_tol > 0.0 ? u->acceptable_tol : 0.0);
    c.acc_iter = u->acceptable_iter > 0 ? u->acceptable_iter : 15;
    memcpy(c.box_lo, u->bounding_box_lower_limit, sizeof c.box_lo);
    memcpy(c.box_up, u->bounding_box_upper_limit, sizeof c.box_up);
    h->device = u->device;
    DeviceGuard guard(h->device);
    cudaError_t e = guard.err;
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, h->device);
    if (e != cudaSuccess) { delete h; return CMPC_E_CUDA; }
    h->sm_count = prop.multiProcessorCount;
    // threads_per_instance: team size (0 = default); teams_per_cta: teams walking in lock-step through one CTA (0 = default)
    h->threads = u->threads_per_instance == 0 ? CMPC_DEFAULT_TEAM : u->threads_per_instance;
    h->lockstep = u->teams_per_cta == 0 ? (u->threads_per_instance == 0 ? CMPC_DEFAULT_LOCKSTEP : 1) : u->teams_per_cta;
    h->groups = u->lockstep_groups > 0 ? u->lockstep_groups : CMPC_DEFAULT_GROUPS;
    if (h->groups < 1) h->groups = 1;
    if (h->groups > h->lockstep) h->groups = h->lockstep;
    const void* kfn = team_kernel(h->threads, h->lockstep);
    if (!kfn) { delete h; return CMPC_E_INVALID; }
    // small batches (at most four instances per SM) are latency bound: one team of 128 threads per CTA (default geometry only)
    h->latency_path = u->threads_per_instance == 0 && u->teams_per_cta == 0;
    h->smem = (int)sizeof(ISmem) * h->lockstep;
    int occ = 0;
    e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, h->threads * h->lockstep, h->smem);
    if (e != cudaSuccess || occ < 1) { h->last_cuda = (int)e; delete h; return CMPC_E_CUDA; }
    h->ctas_per_sm = (u->ctas_per_sm > 0 && u->ctas_per_sm < occ) ? u->ctas_per_sm : occ;
    // Function attributes are per device, not per handle, so they are set HERE, once, to values that depend on nothing but
    // the kernel: every handle writes the same numbers and no solve call touches them.
    //  * carve-out of the single-team kernels: what their resident CTAs need (2 x / 4 x (ISmem + 1 KB)), the rest of the
    //    256 KB stays L1 behind the scratch vectors and the spills (a lone team is 10 % faster than with the carve-out at its
    //    maximum, profiles/r1_notes.md)
    //  * the auxiliary kernels take 8 (53 N + 15) / 8 (40 N + 64) bytes of dynamic shared memory: opt in above 48 KB
    {
        auto pct = [](int ctas) { int v = (int)((100.0 * ctas * (sizeof(ISmem) + 1024)) / (228.0 * 1024.0)) + 1; return v > 100 ? 100 : v; };
        if (h->ctas_per_sm < occ) e = cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, pct(h->ctas_per_sm * h->lockstep));
        if (e == cudaSuccess && h->latency_path) {
            e = cudaFuncSetAttribute((const void*)cmpc_solve_team_kernel<128, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ISmem));
            if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)cmpc_solve_team_kernel<128, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ISmem));
            if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)cmpc_solve_team_kernel<256, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ISmem));
            if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)cmpc_solve_team_kernel<256, 1, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct(1));
            if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)cmpc_solve_team_kernel<128, 1, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct(2));
            if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)cmpc_solve_team_kernel<128, 1, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct(4));
        }
        h->shift_smem = (int)(sizeof(double) * dim_g(c.N));
        h->eval_smem = (int)(sizeof(double) * (c.N * SD_STRIDE + 64));
        if (e == cudaSuccess && h->shift_smem > 48 * 1024) e = cudaFuncSetAttribute((const void*)cmpc_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->shift_smem);
        if (e == cudaSuccess && h->eval_smem > 48 * 1024) e = cudaFuncSetAttribute((const void*)cmpc_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->eval_smem);
        if (e != cudaSuccess) { h->last_cuda = (int)e; delete h; return CMPC_E_CUDA; }
    }
    h->grid = h->sm_count * h->ctas_per_sm;  // persistent grid: a multiple of the SM count
#if CMPC_SPLIT_RIC
    h->work_stride = ((size_t)works_vector_doubles(c.N) + 15) & ~(size_t)15;
    h->ric_stride = ((size_t)c.N * WRIC_STRIDE + 15) & ~(size_t)15;
#else
    h->work_stride = ((size_t)works_doubles(c.N) + 15) & ~(size_t)15;   // factor blocks behind the vectors of their team
    h->ric_stride = 0;
#endif
    h->work_slots = std::max((size_t)h->grid * h->lockstep, h->latency_path ? (size_t)4 * h->sm_count : (size_t)0);
    Csc jc, hc;
    build_csc(c.N, false, jc);
    build_csc(c.N, true, hc);
    if (cudaMalloc(&h->d_work, sizeof(double) * (h->work_stride + h->ric_stride) * h->work_slots) != cudaSuccess ||
        cudaMalloc(&h->d_counter, sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&h->d_jslot, sizeof(int) * jc.slot.size()) != cudaSuccess ||
        cudaMalloc(&h->d_hslot, sizeof(int) * hc.slot.size()) != cudaSuccess ||
        cudaMalloc(&h->d_cmap, sizeof(unsigned short) * CF_DINV) != cudaSuccess ||
        cudaMalloc(&h->d_gscratch, sizeof(double) * dim_g(c.N) * h->sm_count * 4) != cudaSuccess) {
        cmpc_destroy(h);
        return CMPC_E_ALLOC;
    }
    std::vector<unsigned short> cm(CF_DINV);
    build_cmap(cm.data());
    h->d_ric = h->d_work + h->work_stride * h->work_slots;
    e = cudaMemset(h->d_work, 0, sizeof(double) * (h->work_stride + h->ric_stride) * h->work_slots);
#if CMPC_L2_WINDOW
    // The iterate vectors of all resident teams (163 KB each at N = 15: more than the L2 for 1036 teams) are read and written by
    // every element-wise pass of every iteration; the factor blocks stream through once per sweep.  A persisting-L2 access window
    // over the vector region (hit ratio = what the set-aside can hold) keeps that fraction of the vectors resident instead of
    // letting the whole set thrash.  Device-wide limit, set once; the window is a launch attribute of the solver kernel only.
    if (e == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        const size_t vec_bytes = sizeof(double) * h->work_stride * h->work_slots;
        const size_t persist = std::min((size_t)(CMPC_L2_FRACTION * prop.persistingL2CacheMaxSize), vec_bytes);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
            h->l2_window_bytes = std::min(vec_bytes, (size_t)prop.accessPolicyMaxWindowSize);
            h->l2_hit_ratio = (float)std::min(1.0, (double)persist / (double)h->l2_window_bytes);
        } else cudaGetLastError();
    }
#endif
    if (e == cudaSuccess) e = cudaMemcpy(h->d_cmap, cm.data(), sizeof(unsigned short) * CF_DINV, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_jslot, jc.slot.data(), sizeof(int) * jc.slot.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_hslot, hc.slot.data(), sizeof(int) * hc.slot.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { cmpc_destroy(h); return CMPC_E_CUDA; }
    *out = h;
    return CMPC_OK;
}

int cmpc_destroy(cmpc_handle h)
{
    if (!h) return CMPC_E_INVALID;
    {
        DeviceGuard guard(h->device);
        if (h->in_flight && h->done) cudaEventSynchronize(h->done);  // the scratch arena must outlive the last solve
        if (h->done) cudaEventDestroy(h->done);
        if (h->hstream) cudaStreamDestroy(h->hstream);
        cudaFree(h->d_work); cudaFree(h->d_counter); cudaFree(h->d_jslot); cudaFree(h->d_hslot); cudaFree(h->d_gscratch); cudaFree(h->d_cmap);
        cudaFree(h->d_p); cudaFree(h->d_lbg); cudaFree(h->d_ubg); cudaFree(h->d_x); cudaFree(h->d_lam); cudaFree(h->d_obj);
        cudaFree(h->d_status); cudaFree(h->d_iters); cudaFree(h->d_ticks);
    }
    delete h;
    return CMPC_OK;
}

int cmpc_solve_batched(cmpc_handle h, int batch, const double* d_p, const double* d_lbg, const double* d_ubg,
                       double* d_x, double* d_lam_g, double* d_obj, int* d_status, int* d_iters, int warm_duals,
                       void* stream)
{
    if (!h || batch < 0 || !d_p || !d_lbg || !d_ubg || !d_x) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CMPC_ON_DEVICE(h);
    // one work queue and one scratch arena per handle: a solve on another stream than the previous one is ordered behind it.
    // While the stream is being captured into a CUDA graph the event is left alone: a graph replays in stream order.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CK(cudaStreamIsCapturing(st, &cap));
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    if (!capturing && h->in_flight && st != h->last_stream) CK(cudaStreamWaitEvent(st, h->done, 0));
    CK(cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned int), st));
#if CMPC_L2_WINDOW
    const bool window = h->l2_window_bytes > 0 && !capturing;
    if (window) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        av.accessPolicyWindow.base_ptr = h->d_work;
        av.accessPolicyWindow.num_bytes = h->l2_window_bytes;
        av.accessPolicyWindow.hitRatio = h->l2_hit_ratio;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = CMPC_L2_MISS_NORMAL ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
#endif
    // small batches are latency bound: independent single-team CTAs of 128 threads (no lock-step, no 80-register cap) beat the
    // seven-team CTAs while at most four of them share an SM (profiles/r1_notes.md): up to 2 per SM the kernel compiled
    // without a register cap (254 registers, two CTAs run as fast as one), up to 4 per SM the one compiled for 4 CTAs (128)
    if (h->latency_path && batch <= 4 * h->sm_count) {
        const int per_sm = (batch + h->sm_count - 1) / h->sm_count;
        if (per_sm <= 1)   // one instance per SM: a team of 256 threads (8 warps share the tensor-core tiles, the element-wise passes
                           // and the split dot products of the sweeps: 5.5 -> 4.3 ms for a cold iCub3 solve)
            cmpc_solve_team_kernel<256, 1, 1><<<batch, 256, (int)sizeof(ISmem), st>>>(h->cfg, batch, d_p, d_lbg, d_ubg, d_x, d_lam_g, d_obj,
                                                                                     d_status, d_iters, warm_duals, h->d_work,
                                                                                     h->work_stride, h->d_counter, h->d_cmap, 1, h->d_ric, h->ric_stride);
        else if (per_sm <= 2)
            cmpc_solve_team_kernel<128, 1, 1><<<batch, 128, (int)sizeof(ISmem), st>>>(h->cfg, batch, d_p, d_lbg, d_ubg, d_x, d_lam_g, d_obj,
                                                                                     d_status, d_iters, warm_duals, h->d_work,
                                                                                     h->work_stride, h->d_counter, h->d_cmap, 1, h->d_ric, h->ric_stride);
        else
            cmpc_solve_team_kernel<128, 1, 4><<<batch, 128, (int)sizeof(ISmem), st>>>(h->cfg, batch, d_p, d_lbg, d_ubg, d_x, d_lam_g, d_obj,
                                                                                     d_status, d_iters, warm_duals, h->d_work,
                                                                                     h->work_stride, h->d_counter, h->d_cmap, 1, h->d_ric, h->ric_stride);
    } else {
        int grid = std::min(batch, h->grid);  // a small batch is spread over the SMs (static first instance: team * grid + CTA)
#define X(NT, G, C)                                                                                                      \
    if (h->threads == NT && h->lockstep == G)                                                                            \
        cmpc_solve_team_kernel<NT, G, C><<<grid, NT * G, h->smem, st>>>(h->cfg, batch, d_p, d_lbg, d_ubg, d_x, d_lam_g, d_obj, \
                                                                        d_status, d_iters, warm_duals, h->d_work,         \
                                                                        h->work_stride, h->d_counter, h->d_cmap, h->groups, h->d_ric, h->ric_stride);
        CMPC_FOR_EACH_KERNEL(X)
#undef X
    }
    h->launches++;
    CK(cudaGetLastError());
#if CMPC_L2_WINDOW
    if (window) {   // the window belongs to this launch, not to whatever the caller enqueues on the stream next
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        av.accessPolicyWindow.num_bytes = 0;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
#endif
    if (!capturing) {
        CK(cudaEventRecord(h->done, st));
        h->last_stream = st;
        h->in_flight = true;
    }
    return CMPC_OK;
}

static int ensure_host_staging(cmpc_handle h, int batch)
{
    if (batch <= h->host_cap) return CMPC_OK;
    const int N = h->cfg.N;
    if (h->in_flight) cudaEventSynchronize(h->done);  // a solve in flight may still read the old buffers
    cudaFree(h->d_p); cudaFree(h->d_lbg); cudaFree(h->d_ubg); cudaFree(h->d_x); cudaFree(h->d_lam); cudaFree(h->d_obj);
    cudaFree(h->d_status); cudaFree(h->d_iters); cudaFree(h->d_ticks);
    h->d_p = h->d_lbg = h->d_ubg = h->d_x = h->d_lam = h->d_obj = h->d_ticks = nullptr;
    h->d_status = h->d_iters = nullptr;
    h->host_cap = 0;
    h->resident_batch = 0;
    size_t b = (size_t)batch;
    if (cudaMalloc(&h->d_p, 8 * b * dim_p(N)) != cudaSuccess || cudaMalloc(&h->d_lbg, 8 * b * dim_g(N)) != cudaSuccess ||
        cudaMalloc(&h->d_ubg, 8 * b * dim_g(N)) != cudaSuccess || cudaMalloc(&h->d_x, 8 * b * dim_x(N)) != cudaSuccess ||
        cudaMalloc(&h->d_lam, 8 * b * dim_g(N)) != cudaSuccess || cudaMalloc(&h->d_obj, 8 * b) != cudaSuccess ||
        cudaMalloc(&h->d_status, 4 * b) != cudaSuccess || cudaMalloc(&h->d_iters, 4 * b) != cudaSuccess ||
        cudaMalloc(&h->d_ticks, 8 * b * tick_stride(N)) != cudaSuccess) {
        cudaFree(h->d_p); cudaFree(h->d_lbg); cudaFree(h->d_ubg); cudaFree(h->d_x); cudaFree(h->d_lam); cudaFree(h->d_obj);
        cudaFree(h->d_status); cudaFree(h->d_iters); cudaFree(h->d_ticks);
        h->d_p = h->d_lbg = h->d_ubg = h->d_x = h->d_lam = h->d_obj = h->d_ticks = nullptr;
        h->d_status = h->d_iters = nullptr;
        return CMPC_E_ALLOC;
    }
    h->host_cap = batch;
    return CMPC_OK;
}

int cmpc_solve_host(cmpc_handle h, int batch, const double* p, const double* lbg, const double* ubg, double* x,
                    double* lam_g, double* obj, int* status, int* iters, int warm_duals)
{
    if (!h || batch < 0 || !p || !lbg || !ubg || !x) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    int rc = ensure_host_staging(h, batch);
    if (rc) return rc;
    const int N = h->cfg.N;
    size_t b = (size_t)batch;
    CK(cudaMemcpyAsync(h->d_p, p, 8 * b * dim_p(N), cudaMemcpyHostToDevice, h->hstream));
    CK(cudaMemcpyAsync(h->d_lbg, lbg, 8 * b * dim_g(N), cudaMemcpyHostToDevice, h->hstream));
    CK(cudaMemcpyAsync(h->d_ubg, ubg, 8 * b * dim_g(N), cudaMemcpyHostToDevice, h->hstream));
    CK(cudaMemcpyAsync(h->d_x, x, 8 * b * dim_x(N), cudaMemcpyHostToDevice, h->hstream));
    if (warm_duals && lam_g) CK(cudaMemcpyAsync(h->d_lam, lam_g, 8 * b * dim_g(N), cudaMemcpyHostToDevice, h->hstream));
    rc = cmpc_solve_batched(h, batch, h->d_p, h->d_lbg, h->d_ubg, h->d_x, h->d_lam, h->d_obj, h->d_status, h->d_iters,
                            warm_duals && lam_g, h->hstream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(x, h->d_x, 8 * b * dim_x(N), cudaMemcpyDeviceToHost, h->hstream));
    if (lam_g) CK(cudaMemcpyAsync(lam_g, h->d_lam, 8 * b * dim_g(N), cudaMemcpyDeviceToHost, h->hstream));
    if (obj) CK(cudaMemcpyAsync(obj, h->d_obj, 8 * b, cudaMemcpyDeviceToHost, h->hstream));
    if (status) CK(cudaMemcpyAsync(status, h->d_status, 4 * b, cudaMemcpyDeviceToHost, h->hstream));
    if (iters) CK(cudaMemcpyAsync(iters, h->d_iters, 4 * b, cudaMemcpyDeviceToHost, h->hstream));
    CK(cudaStreamSynchronize(h->hstream));
    h->resident_batch = batch;
    return CMPC_OK;
}

int cmpc_tick_stride(int horizon) { return horizon < 1 ? CMPC_E_INVALID : tick_stride(horizon); }

int cmpc_populate(cmpc_handle h, int batch, const double* d_ticks, double* d_p, double* d_lbg, double* d_ubg, double* d_x0,
                  void* stream)
{
    if (!h || batch < 0 || !d_ticks || !d_p || !d_lbg || !d_ubg) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    const int grid = std::min(batch, h->sm_count * 8);
    cmpc_populate_kernel<<<grid, 64, sizeof(double) * tick_stride(h->cfg.N), (cudaStream_t)stream>>>(h->cfg, batch, d_ticks, d_p, d_lbg,
                                                                                                    d_ubg, d_x0);
    h->launches++;
    CK(cudaGetLastError());
    return CMPC_OK;
}

int cmpc_solve_ticks_host(cmpc_handle h, int batch, const double* ticks, int warm_mode, double* x, double* lam_g, double* obj,
                          int* status, int* iters)
{
    if (!h || batch < 0 || !ticks || !x || warm_mode < 0 || warm_mode > 2) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    if (warm_mode == 1 && h->resident_batch != batch) return CMPC_E_INVALID;   // nothing (or something else) is resident
    CMPC_ON_DEVICE(h);
    int rc = ensure_host_staging(h, batch);
    if (rc) return rc;
    const int N = h->cfg.N;
    const size_t b = (size_t)batch;
    CK(cudaMemcpyAsync(h->d_ticks, ticks, 8 * b * tick_stride(N), cudaMemcpyHostToDevice, h->hstream));
    bool warm_duals = false;
    if (warm_mode == 2) {
        CK(cudaMemcpyAsync(h->d_x, x, 8 * b * dim_x(N), cudaMemcpyHostToDevice, h->hstream));
        if (lam_g) CK(cudaMemcpyAsync(h->d_lam, lam_g, 8 * b * dim_g(N), cudaMemcpyHostToDevice, h->hstream));
        warm_duals = lam_g != nullptr;
    } else if (warm_mode == 1) warm_duals = true;
    rc = cmpc_populate(h, batch, h->d_ticks, h->d_p, h->d_lbg, h->d_ubg, warm_mode == 0 ? h->d_x : nullptr, h->hstream);
    if (rc) return rc;
    if (warm_mode != 0) {
        rc = cmpc_shift_warmstart(h, batch, h->d_x, warm_duals ? h->d_lam : nullptr, h->hstream);
        if (rc) return rc;
    }
    rc = cmpc_solve_batched(h, batch, h->d_p, h->d_lbg, h->d_ubg, h->d_x, h->d_lam, h->d_obj, h->d_status, h->d_iters,
                            warm_duals ? 1 : 0, h->hstream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(x, h->d_x, 8 * b * dim_x(N), cudaMemcpyDeviceToHost, h->hstream));
    if (lam_g) CK(cudaMemcpyAsync(lam_g, h->d_lam, 8 * b * dim_g(N), cudaMemcpyDeviceToHost, h->hstream));
    if (obj) CK(cudaMemcpyAsync(obj, h->d_obj, 8 * b, cudaMemcpyDeviceToHost, h->hstream));
    if (status) CK(cudaMemcpyAsync(status, h->d_status, 4 * b, cudaMemcpyDeviceToHost, h->hstream));
    if (iters) CK(cudaMemcpyAsync(iters, h->d_iters, 4 * b, cudaMemcpyDeviceToHost, h->hstream));
    CK(cudaStreamSynchronize(h->hstream));
    h->resident_batch = batch;
    return CMPC_OK;
}

int cmpc_resample_references(cmpc_handle h, int batch, int n_in, const double* d_t_in, const double* d_com_in,
                             const double* d_h_in, const double* d_t_out, double robot_mass, double com_height, double* d_ticks,
                             void* stream)
{
    if (!h || batch < 0 || n_in < 1 || !d_t_in || !d_com_in || !d_h_in || !d_t_out || !d_ticks || !(robot_mass > 0)) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    const int N = h->cfg.N;
    const long long total = (long long)batch * (N + 1);
    const int grid = (int)std::min<long long>((total + 127) / 128, (long long)h->sm_count * 8);
    cmpc_resample_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(N, batch, n_in, d_t_in, d_com_in, d_h_in, d_t_out, 1.0 / robot_mass,
                                                                 com_height, d_ticks);
    h->launches++;
    CK(cudaGetLastError());
    return CMPC_OK;
}

int cmpc_desired_zmp(cmpc_handle h, int batch, const double* d_x, const double* d_p, double half_length, double half_width,
                     double* d_zmp, int* d_valid, void* stream)
{
    if (!h || batch < 0 || !d_x || !d_p || !d_zmp) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    const int grid = std::min((batch + 127) / 128, h->sm_count * 8);
    cmpc_zmp_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(h->cfg, batch, d_x, d_p, half_length, half_width, d_zmp, d_valid);
    h->launches++;
    CK(cudaGetLastError());
    return CMPC_OK;
}

int cmpc_rollout_layout(int* roll_stride, int* max_steps, int* step_stride)
{
    if (roll_stride) *roll_stride = ROLL_STRIDE;
    if (max_steps) *max_steps = RL_MAXSTEPS;
    if (step_stride) *step_stride = RL_STEP;
    return CMPC_OK;
}

static bool walk_params(const cmpc_walk_params* w, WalkParams& W)
{
    if (!w || w->ds_knots < 0 || w->ss_knots < 1) return false;
    W.ds = w->ds_knots; W.ss = w->ss_knots; W.step_length = w->step_length; W.com_height = w->com_height;
    W.push_threshold = w->push_threshold; W.zmp_half_length = w->zmp_half_length; W.zmp_half_width = w->zmp_half_width;
    return true;
}

int cmpc_rollout_tick(cmpc_handle h, const cmpc_walk_params* w, int batch, int tick, const double* d_roll, const double* d_state,
                      const double* d_steps, double* d_ticks, double* d_ext6, int step_adjust, void* stream)
{
    WalkParams W;
    if (!h || batch < 0 || !d_roll || !d_state || !d_steps || !d_ticks || !d_ext6 || !walk_params(w, W)) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    const int grid = std::min((batch + 63) / 64, h->sm_count * 8);
    cmpc_rollout_tick_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>(h->cfg, W, batch, tick, d_roll, d_state, d_steps, d_ticks, d_ext6,
                                                                    step_adjust);
    h->launches++;
    CK(cudaGetLastError());
    return CMPC_OK;
}

int cmpc_rollout_feedback(cmpc_handle h, const cmpc_walk_params* w, int batch, int tick, const double* d_x, const double* d_p,
                          const double* d_state, const int* d_status, const int* d_iters, double* d_roll, double* d_steps,
                          void* stream)
{
    WalkParams W;
    if (!h || batch < 0 || !d_x || !d_p || !d_state || !d_status || !d_iters || !d_roll || !d_steps || !walk_params(w, W))
        return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    const int grid = std::min((batch + 63) / 64, h->sm_count * 8);
    cmpc_rollout_feedback_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>(h->cfg, W, batch, tick, d_x, d_p, d_state, d_status, d_iters,
                                                                        d_roll, d_steps);
    h->launches++;
    CK(cudaGetLastError());
    return CMPC_OK;
}

int cmpc_shift_warmstart(cmpc_handle h, int batch, double* d_x, double* d_lam_g, void* stream)
{
    if (!h || batch < 0 || !d_x) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    const int N = h->cfg.N;
    int grid = std::min(batch, h->sm_count * 8);
    cmpc_shift_kernel<<<grid, 256, h->shift_smem, (cudaStream_t)stream>>>(N, batch, d_x, d_lam_g);
    h->launches++;
    CK(cudaGetLastError());
    return CMPC_OK;
}

int cmpc_eval_fg(cmpc_handle h, int batch, const double* d_x, const double* d_p, double* d_f, double* d_g, void* stream)
{
    return cmpc_eval_jac_fg(h, batch, d_x, d_p, d_f, nullptr, d_g, nullptr, stream);
}

int cmpc_eval_jac_fg(cmpc_handle h, int batch, const double* d_x, const double* d_p, double* d_f, double* d_grad,
                     double* d_g, double* d_jac, void* stream)
{
    if (!h || batch < 0 || !d_x || !d_p) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->cfg.N;
    if (d_f || d_grad || d_g) {
        int grid = std::min(batch, h->sm_count * 4);
        cmpc_eval_kernel<<<grid, 128, h->eval_smem, st>>>(h->cfg, batch, d_x, d_p, d_f, d_grad, d_g, h->d_gscratch);
        h->launches++;
        CK(cudaGetLastError());
    }
    if (d_jac) {
        long long total = (long long)batch * nnz_jac(N);
        int grid = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
        cmpc_jac_kernel<<<grid, 256, 0, st>>>(h->cfg, batch, d_x, d_p, h->d_jslot, d_jac);
        h->launches++;
        CK(cudaGetLastError());
    }
    return CMPC_OK;
}

int cmpc_eval_hess_l(cmpc_handle h, int batch, const double* d_x, const double* d_p, double lam_f, const double* d_lam_g,
                     double* d_hess, void* stream)
{
    (void)d_x;  // the hessian of the lagrangian does not depend on x
    if (!h || batch < 0 || !d_p || !d_lam_g || !d_hess) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    const int N = h->cfg.N;
    long long total = (long long)batch * nnz_hess(N);
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
    cmpc_hess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h->cfg, batch, d_p, lam_f, d_lam_g, h->d_hslot, d_hess);
    h->launches++;
    CK(cudaGetLastError());
    return CMPC_OK;
}

int cmpc_rollout_plant(cmpc_handle h, int batch, const double* d_x, const double* d_p, const double* d_ext,
                       double* d_state, double dt, int substeps, void* stream)
{
    if (!h || batch < 0 || !d_x || !d_p || !d_state || substeps < 0) return CMPC_E_INVALID;
    if (batch == 0) return CMPC_OK;
    CMPC_ON_DEVICE(h);
    int grid = std::min((batch + 127) / 128, h->sm_count * 8);
    cmpc_plant_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(h->cfg, batch, d_x, d_p, d_ext, d_state, dt, substeps);
    h->launches++;
    CK(cudaGetLastError());
    return CMPC_OK;
}

int cmpc_measure_fp64_peak(cmpc_handle h, double* tflops)
{
    if (!h || !tflops) return CMPC_E_INVALID;
    CMPC_ON_DEVICE(h);
    const int iters = 1 << 16, threads = 256, blocks = h->sm_count * 8;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; e == cudaSuccess && rep < 4; ++rep) {  // first repetition is the warm-up
        e = cudaEventRecord(e0, 0);
        cmpc_dfma_probe_kernel<<<blocks, threads>>>(h->d_gscratch, iters, 0.999999, 1e-9);
        if (e == cudaSuccess) e = cudaEventRecord(e1, 0);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) break;
        const double tf = 2.0 * 8.0 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    CK(e);
    *tflops = best;
    return CMPC_OK;
}

// cycle counters of the solve phases (zeros unless the library was built with -DCMPC_PROFILE); resets them
int cmpc_debug_profile(long long* out16)
{
    if (!out16) return CMPC_E_INVALID;
    memset(out16, 0, 16 * sizeof(long long));
#if defined(CMPC_PROFILE)
    long long z[16] = {0};
    cudaMemcpyFromSymbol(out16, cmpc::g_prof, sizeof z);
    cudaMemcpyToSymbol(cmpc::g_prof, z, sizeof z);
#endif
    return CMPC_OK;
}

long long cmpc_launch_count(cmpc_handle h) { return h ? h->launches : 0; }
int cmpc_last_cuda_error(cmpc_handle h) { return h ? h->last_cuda : 0; }

const char* cmpc_error_string(int code)
{
    switch (code) {
        case CMPC_OK: return "ok";
        case CMPC_E_INVALID: return "invalid argument";
        case CMPC_E_CUDA: return "CUDA runtime error";
        case CMPC_E_NO_DEVICE: return "no CUDA device (libcmpc_b200 has no CPU path)";
        case CMPC_E_ALLOC: return "allocation failed";
        default: return "unknown error";
    }
}

int cmpc_solver_geometry(cmpc_handle h, int* grid, int* threads, int* smem, int* ctas_per_sm, int* sm_count)
{
    if (!h) return CMPC_E_INVALID;
    if (grid) *grid = h->grid;
    if (threads) *threads = h->threads;
    if (smem) *smem = h->smem;
    if (ctas_per_sm) *ctas_per_sm = h->ctas_per_sm;
    if (sm_count) *sm_count = h->sm_count;
    return CMPC_OK;
}

int cmpc_solver_lockstep(cmpc_handle h, int* teams_per_cta, int* lockstep_groups)
{
    if (!h) return CMPC_E_INVALID;
    if (teams_per_cta) *teams_per_cta = h->lockstep;
    if (lockstep_groups) *lockstep_groups = h->groups;
    return CMPC_OK;
}

}  // extern "C"
