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
                ("bound_push", C.c_double), ("inf_bound", C.c_double), ("pc", C.c_int), ("mu_warm", C.c_double)]


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


def dev_config(N, tol=1e-8, pc=0, **kw):
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
    return c, o


def run(L, fn, c, w, b):
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    x = np.array(w["x0"][b], dtype=np.float64, copy=True)
    lam = np.zeros(w["lbg"].shape[1])
    it, obj, kkt = C.c_int(), C.c_double(), C.c_double()
    p, lb, ub = (np.ascontiguousarray(w[k][b]) for k in ("p", "lbg", "ubg"))
    st = getattr(L, fn)(C.byref(c), vp(p), vp(lb), vp(ub), vp(x), vp(lam), 0, C.byref(it), C.byref(obj), C.byref(kkt))
    return st, it.value, obj.value, x, lam


@pytest.mark.parametrize("fn", ["hostsim_solve", "hostsim_solve_team32", "hostsim_solve_team128"])
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
