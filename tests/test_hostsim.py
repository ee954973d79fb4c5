"""CPU-side check of the CUDA solver's mathematics (marker: not gpu).

tests/hostsim/hostsim.cpp compiles the very same __host__ __device__ solver source (csrc/cmpc_core.cuh, csrc/cmpc_warp.cuh)
with g++: the generic sweeps with a one-thread CTA and the warp-per-instance sweeps with 32 emulated lanes.  Both must
land on the oracle's optimum.  TEST-ONLY: the product library never contains this build.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, pkg
from oracle.oracle import make_cfg

HS = os.path.join(ROOT, "tests", "hostsim")


class DevConfig(C.Structure):
    """struct cmpc::Config of csrc/cmpc_core.cuh"""
    _fields_ = [("N", C.c_int), ("dT", C.c_double), ("w_com", C.c_double * 3), ("w_h", C.c_double), ("w_pos", C.c_double),
                ("w_sym", C.c_double), ("w_rate", C.c_double * 3), ("corner", C.c_double * 24), ("fricA", C.c_double * 12),
                ("tol", C.c_double), ("max_iter", C.c_int), ("mu_init", C.c_double), ("bound_relax", C.c_double),
                ("bound_push", C.c_double), ("inf_bound", C.c_double), ("pc", C.c_int), ("mu_warm", C.c_double),
                ("scal_max_grad", C.c_double), ("acc_tol", C.c_double), ("acc_iter", C.c_int),
                ("box_lo", C.c_double * 6), ("box_up", C.c_double * 6)]


@pytest.fixture(scope="module")
def hostsim():
    src = os.path.join(HS, "hostsim.cpp")
    lib = os.path.join(HS, "libhostsim.so")
    deps = [src, os.path.join(HS, "cmpc_generic.cuh")] + [os.path.join(ROOT, pkg().__name__, "csrc", f) for f in ("cmpc_core.cuh", "cmpc_warp.cuh", "cmpc_ipm.cuh", "cmpc_layout.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", lib], check=True)
    L = C.CDLL(lib)
    assert L.hostsim_config_size() == C.sizeof(DevConfig)
    return L


def dev_config(N, tol=1e-8, pc=0, scal=100.0, acc_tol=1e-6, acc_iter=15, **kw):
    o = make_cfg(N=N, **kw)   # oracle config: same fields (weights, corners, friction matrix)
    c = DevConfig()
    c.N, c.dT = N, o.dT
    c.w_com[:] = list(o.w_com)
    c.w_h, c.w_pos, c.w_sym = o.w_h, o.w_pos, o.w_sym
    c.w_rate[:] = list(o.w_rate)
    c.corner[:] = list(o.corners)
    from oracle.oracle import Oracle
    c.fricA[:] = np.asarray(Oracle().friction_matrix(o.mu), dtype=np.float64).reshape(-1).tolist()
    c.tol, c.max_iter, c.mu_init, c.bound_relax, c.bound_push, c.inf_bound = tol, 200, 0.1, 1e-8, 0.01, 1e19
    c.pc = pc
    c.mu_warm = 0.01
    c.scal_max_grad, c.acc_tol, c.acc_iter = scal, acc_tol, acc_iter
    return c, o


def run(L, fn, c, w, b):
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    x = np.array(w["x0"][b], dtype=np.float64, copy=True)
    lam = np.zeros(w["lbg"].shape[1])
    it, obj, kkt = C.c_int(), C.c_double(), C.c_double()
    p, lb, ub = (np.ascontiguousarray(w[k][b]) for k in ("p", "lbg", "ubg"))
    st = getattr(L, fn)(C.byref(c), vp(p), vp(lb), vp(ub), vp(x), vp(lam), 0, C.byref(it), C.byref(obj), C.byref(kkt))
    return st, it.value, obj.value, x, lam


@pytest.mark.parametrize("fn", ["hostsim_solve", "hostsim_solve_team32", "hostsim_solve_team96", "hostsim_solve_team128"])
@pytest.mark.parametrize("N,kw,wkw", [
    (12, dict(), dict(state_noise=1.0, yaw_range=0.2)),
    (15, dict(w_com=(1.0, 1.0, 200.0), w_pos=200.0, w_sym=0.0,
              corners=[[(0.08, 0.03, 0), (0.08, -0.03, 0), (-0.08, -0.03, 0), (-0.08, 0.03, 0)]] * 2),
     dict(state_noise=1.0, step_adjust=False)),
])
@pytest.mark.parametrize("pc", [0, 1], ids=["monotone", "mehrotra"])
def test_hostsim_matches_oracle(hostsim, oracle, workloads, fn, N, kw, wkw, pc):
    if fn == "hostsim_solve" and pc:
        pytest.skip("the generic one-thread solver of cmpc_core.cuh has the monotone update only")
    c, o = dev_config(N, pc=pc, **kw)
    w = workloads.walk_batch(N=N, B=3, seed=5, **wkw)
    xo, lo, st = oracle.solve_batch(o, w["p"], w["lbg"], w["ubg"], w["x0"], threads=3, opts=oracle.default_opts(mehrotra=pc))
    for b in range(3):
        status, it, obj, x, lam = run(hostsim, fn, c, w, b)
        assert status == 0 and st[b].status == 0
        assert abs(obj - st[b].obj) <= 1e-6 * max(1.0, abs(st[b].obj)), (obj, st[b].obj)
        assert np.max(np.abs(x - xo[b])) <= 1e-5 * max(1.0, np.max(np.abs(xo[b])))
        assert abs(it - st[b].iters) <= 3, (it, st[b].iters)


@pytest.mark.parametrize("pc", [0, 1], ids=["monotone", "mehrotra"])
@pytest.mark.parametrize("tol", [1e-8, 1e-4])
def test_hostsim_objective_scaling_matches_oracle(hostsim, oracle, workloads, pc, tol):
    """IPOPT's gradient-based NLP scaling: from x0 = 0 (CasADi's initial guess when BLF's warm start is off) the gradient of
    the objective is ~5.6e4, so the solve runs on 1.8e-3 f.  The kernel source gets there by re-scaling the parameters of the
    unscaled run (cmpc_ipm.cuh), the oracle by scaling the functions: same iterates, same iteration counts; and the scaled
    run stops earlier / elsewhere than the unscaled one, as IPOPT would."""
    c, o = dev_config(12, tol=tol, pc=pc, w_pos=2000.0)
    w = workloads.walk_batch(N=12, B=3, seed=5, state_noise=2.0, yaw_range=0.3)
    w["x0"] = np.zeros_like(w["x0"])
    xo, lo, st = oracle.solve_batch(o, w["p"], w["lbg"], w["ubg"], w["x0"], threads=3, opts=oracle.default_opts(mehrotra=pc, tol=tol))
    xu, lu, su = oracle.solve_batch(o, w["p"], w["lbg"], w["ubg"], w["x0"], threads=3,
                                    opts=oracle.default_opts(mehrotra=pc, tol=tol, nlp_scaling_max_gradient=0.0))
    for b in range(3):
        assert st[b].obj_scaling < 2e-3 and st[b].min_g_scaling == 1.0     # no row of g is ever scaled
        status, it, obj, x, lam = run(hostsim, "hostsim_solve_team128", c, w, b)
        assert status == 0 and st[b].status == 0
        assert abs(it - st[b].iters) <= 1, (it, st[b].iters)
        assert abs(obj - st[b].obj) <= 1e-6 * max(1.0, abs(st[b].obj)), (obj, st[b].obj)
        assert np.max(np.abs(x - xo[b])) <= 1e-5 * max(1.0, np.max(np.abs(xo[b])))
        assert np.max(np.abs(lam - lo[b])) <= 1e-4 * max(1.0, np.max(np.abs(lo[b])))    # multipliers of the UNSCALED problem
    assert [s.iters for s in st] != [s.iters for s in su]


@pytest.mark.parametrize("pc", [0, 1], ids=["monotone", "mehrotra"])
def test_hostsim_acceptable_level_termination(hostsim, oracle, workloads, pc):
    """IPOPT's acceptable-level test (acceptable_tol for acceptable_iter consecutive iterates): with a tolerance nobody can
    meet (1e-14) and acceptable_tol 1e-6 x 3 both implementations stop with status 5 after the same number of iterations"""
    c, o = dev_config(12, tol=1e-14, pc=pc, acc_tol=1e-6, acc_iter=3, w_pos=2000.0)
    w = workloads.walk_batch(N=12, B=3, seed=6, state_noise=1.0, yaw_range=0.2)
    xo, lo, st = oracle.solve_batch(o, w["p"], w["lbg"], w["ubg"], w["x0"], threads=3,
                                    opts=oracle.default_opts(mehrotra=pc, tol=1e-14, acceptable_tol=1e-6, acceptable_iter=3))
    for b in range(3):
        status, it, obj, x, lam = run(hostsim, "hostsim_solve_team32", c, w, b)
        assert status == 5 and st[b].status == 5, (status, st[b].status)
        if not pc:   # with mu below 1e-12 the predictor-corrector iterates are rounding noise: only the monotone counts are comparable
            assert abs(it - st[b].iters) <= 1, (it, st[b].iters)
        assert abs(obj - st[b].obj) <= 1e-6 * max(1.0, abs(st[b].obj))
