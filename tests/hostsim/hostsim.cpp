// tests/hostsim/hostsim.cpp -- TEST-ONLY host build of the CUDA solver source (csrc/cmpc_ipm.cuh, csrc/cmpc_warp.cuh with emulated
// lanes) next to the first-generation generic formulation (cmpc_generic.cuh, an independent second implementation).
//
// There is no GPU in the development container, so the solver's math is debugged by compiling the very same
// __host__ __device__ source with a one-thread "CTA".  This library is built and loaded ONLY by
// tests/test_hostsim.py (marker: not gpu); the product library never contains or falls back to it.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../paper_romualdi_2022_icra_centroidal-mpc-walking_b200/csrc/cmpc_ipm.cuh"
#include "cmpc_generic.cuh"

namespace {
struct HostCta {
    int tid = 0, nt = 1, warp = 0, lane = 0, wsize = 1;
    void sync() {}
    void syncwarp() {}
    double sum(double v) { return v; }
    double max(double v) { return v; }
    template <int K> void sumv(double*) {}
    template <int K> void maxv(double*) {}
    template <int K> void minv(double*) {}
    template <int KM, int KN, int KS> void reduce3(double*, double*, double*) {}
};
}  // namespace

extern "C" int hostsim_solve(const cmpc::Config* cfg, const double* p, const double* lbg, const double* ubg, double* x,
                             double* lam, int warm_duals, int* iters, double* obj, double* kkt)
{
    HostCta cta;
    std::vector<double> buf(cmpc::work_doubles(cfg->N), 0.0);
    cmpc::Work w;
    cmpc::work_carve(buf.data(), cfg->N, w);
    static cmpc::Smem sm;
    cmpc::Instance in{p, lbg, ubg};
    cmpc::LinCta lin{sm};
    cmpc::Result r = cmpc::ipm_solve(cta, *cfg, in, w, lin, x, lam, warm_duals);
    *iters = r.iters; *obj = r.obj; *kkt = r.kkt;
    return r.status;
}

// the team sweeps (csrc/cmpc_warp.cuh) with NT emulated threads
template <int NT>
static int solve_team(const cmpc::Config* cfg, const double* p, const double* lbg, const double* ubg, double* x, double* lam,
                      int warm_duals, int* iters, double* obj, double* kkt)
{
    HostCta cta;
    std::vector<double> buf(cmpc::works_doubles(cfg->N), 0.0);
    static cmpc::ISmem sm;
    static unsigned short cmap[cmpc::CF_DINV];
    cmpc::build_cmap(cmap);
    unsigned int counter = 0;
    int st = -1;
    cmpc::Team T{0, 0, true};
    cmpc::ipm_run<NT, 1>(T, cta, *cfg, buf.data(), sm, cmap, 1, p, lbg, ubg, x, lam, obj, &st, iters, warm_duals, &counter);
    *kkt = 0.0;
    return st;
}
extern "C" int hostsim_solve_team32(const cmpc::Config* cfg, const double* p, const double* lbg, const double* ubg, double* x,
                                    double* lam, int warm_duals, int* iters, double* obj, double* kkt)
{ return solve_team<32>(cfg, p, lbg, ubg, x, lam, warm_duals, iters, obj, kkt); }
extern "C" int hostsim_solve_team96(const cmpc::Config* cfg, const double* p, const double* lbg, const double* ubg, double* x,
                                    double* lam, int warm_duals, int* iters, double* obj, double* kkt)
{ return solve_team<96>(cfg, p, lbg, ubg, x, lam, warm_duals, iters, obj, kkt); }
extern "C" int hostsim_solve_team128(const cmpc::Config* cfg, const double* p, const double* lbg, const double* ubg, double* x,
                                     double* lam, int warm_duals, int* iters, double* obj, double* kkt)
{ return solve_team<128>(cfg, p, lbg, ubg, x, lam, warm_duals, iters, obj, kkt); }

// debug: run `max_iter` iterations and export the internal state and the last search direction
extern "C" int hostsim_dump(const cmpc::Config* cfg, const double* p, const double* lbg, const double* ubg, double* x,
                            double* lam, double* s, double* zL, double* zU, double* sL, double* sU, double* dx, double* dy)
{
    HostCta cta;
    std::vector<double> buf(cmpc::work_doubles(cfg->N), 0.0);
    cmpc::Work w;
    cmpc::work_carve(buf.data(), cfg->N, w);
    static cmpc::Smem sm;
    cmpc::Instance in{p, lbg, ubg};
    cmpc::LinCta lin{sm};
    cmpc::Result r = cmpc::ipm_solve(cta, *cfg, in, w, lin, x, lam, 0);
    const int q = 38 * cfg->N;
    std::memcpy(s, w.s, 8 * q); std::memcpy(zL, w.zL, 8 * q); std::memcpy(zU, w.zU, 8 * q);
    std::memcpy(sL, w.sL, 8 * q); std::memcpy(sU, w.sU, 8 * q);
    std::memcpy(dx, w.dx, 8 * cmpc::dim_x(cfg->N)); std::memcpy(dy, w.dy, 8 * cmpc::dim_g(cfg->N));
    return r.status;
}

#ifdef CMPC_HOST_STATS
extern "C" void hostsim_stats(long* out) { for (int i = 0; i < 8; ++i) { out[i] = cmpc::g_stat[i]; cmpc::g_stat[i] = 0; } }
#endif
extern "C" int hostsim_config_size() { return (int)sizeof(cmpc::Config); }
