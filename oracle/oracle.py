"""ctypes front end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product package never does (it fails loudly when its CUDA library is missing instead).

Two libraries are wrapped:
  * oracle/libcmpc_oracle.so  -- plain-C restatement of the NLP functions and of the interior-point solve
                                 (oracle/cmpc_oracle_nlp.c, oracle/cmpc_oracle_ipm.c), any horizon N;
  * oracle/_ref/libref_{tmp,jit}.so -- the reference's own CasADi-generated code (N = 12), compiled by
                                 oracle/Makefile from /root/reference/.../ergoCubGazeboV1/{tmp.c,jit_tmpComMiH.c}.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(quiet: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class Cfg(C.Structure):
    _fields_ = [("N", C.c_int), ("dT", C.c_double), ("mu", C.c_double), ("w_com", C.c_double * 3),
                ("w_h", C.c_double), ("w_pos", C.c_double), ("w_sym", C.c_double), ("w_rate", C.c_double * 3),
                ("corners", C.c_double * 24)]


class IpmOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int), ("mu_init", C.c_double), ("bound_relax", C.c_double),
                ("bound_push", C.c_double), ("inf_bound", C.c_double), ("warm_duals", C.c_int),
                ("verbose", C.c_int), ("mehrotra", C.c_int), ("nlp_scaling_max_gradient", C.c_double),
                ("acceptable_tol", C.c_double), ("acceptable_iter", C.c_int)]


class IpmStats(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("obj", C.c_double), ("kkt_error", C.c_double),
                ("dual_inf", C.c_double), ("constr_viol", C.c_double), ("compl_inf", C.c_double),
                ("n_reg", C.c_int), ("n_ls_trials", C.c_int), ("n_fallback", C.c_int), ("obj_scaling", C.c_double),
                ("min_g_scaling", C.c_double)]


def make_cfg(N=12, dT=0.1, mu=0.33, w_com=(10.0, 10.0, 200.0), w_h=100.0, w_pos=200.0, w_sym=10.0,
             w_rate=(10.0, 10.0, 10.0), corners=None) -> Cfg:
    """Defaults = the constants baked into the reference's tmp.c."""
    if corners is None:
        one = [(0.08, 0.01, 0.0), (0.08, -0.01, 0.0), (-0.08, -0.01, 0.0), (-0.08, 0.01, 0.0)]
        corners = [one, one]
    cfg = Cfg()
    cfg.N, cfg.dT, cfg.mu = N, dT, mu
    cfg.w_com[:] = w_com
    cfg.w_h, cfg.w_pos, cfg.w_sym = w_h, w_pos, w_sym
    cfg.w_rate[:] = w_rate
    cfg.corners[:] = np.asarray(corners, dtype=np.float64).reshape(-1).tolist()
    return cfg


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class Oracle:
    """libcmpc_oracle.so (restatement)."""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "libcmpc_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        for name in ("nx", "np", "ng", "nnz_jac", "nnz_hess"):
            fn = getattr(L, "cmpc_oracle_" + name)
            fn.argtypes, fn.restype = [C.c_int], C.c_int
        L.cmpc_oracle_friction_matrix.argtypes = [C.c_double, _dp]
        L.cmpc_oracle_jac_sparsity.argtypes = [C.c_int, _ip, _ip]
        L.cmpc_oracle_hess_sparsity.argtypes = [C.c_int, _ip, _ip]
        L.cmpc_oracle_fg.argtypes = [C.POINTER(Cfg), _dp, _dp, _dp, _dp]
        L.cmpc_oracle_jac_fg.argtypes = [C.POINTER(Cfg), _dp, _dp, _dp, _dp, _dp, _dp]
        L.cmpc_oracle_hess_l.argtypes = [C.POINTER(Cfg), _dp, _dp, C.c_double, _dp, _dp]
        if hasattr(L, "cmpc_oracle_ipm_solve"):
            L.cmpc_oracle_ipm_default_opts.argtypes = [C.POINTER(IpmOpts)]
            L.cmpc_oracle_ipm_solve.argtypes = [C.POINTER(Cfg), C.POINTER(IpmOpts), _dp, _dp, _dp, _dp, _dp,
                                                C.POINTER(IpmStats)]
            L.cmpc_oracle_ipm_solve.restype = C.c_int
            L.cmpc_oracle_ipm_solve_batch.argtypes = [C.POINTER(Cfg), C.POINTER(IpmOpts), C.c_int, C.c_int, _dp,
                                                      _dp, _dp, _dp, _dp, C.POINTER(IpmStats)]
            L.cmpc_oracle_ipm_solve_batch.restype = C.c_int

    # dims
    def dims(self, N):
        L = self.lib
        return dict(n=L.cmpc_oracle_nx(N), np=L.cmpc_oracle_np(N), m=L.cmpc_oracle_ng(N),
                    nnz_j=L.cmpc_oracle_nnz_jac(N), nnz_h=L.cmpc_oracle_nnz_hess(N))

    def friction_matrix(self, mu):
        A = np.zeros((4, 3))
        self.lib.cmpc_oracle_friction_matrix(mu, A.ctypes.data_as(_dp))
        return A

    def jac_sparsity(self, N):
        d = self.dims(N)
        colind = np.zeros(d["n"] + 1, dtype=np.int32)
        row = np.zeros(d["nnz_j"], dtype=np.int32)
        self.lib.cmpc_oracle_jac_sparsity(N, colind.ctypes.data_as(_ip), row.ctypes.data_as(_ip))
        return colind, row

    def hess_sparsity(self, N):
        d = self.dims(N)
        colind = np.zeros(d["n"] + 1, dtype=np.int32)
        row = np.zeros(d["nnz_h"], dtype=np.int32)
        self.lib.cmpc_oracle_hess_sparsity(N, colind.ctypes.data_as(_ip), row.ctypes.data_as(_ip))
        return colind, row

    def fg(self, cfg, x, p):
        d = self.dims(cfg.N)
        x, xp = _arr(x)
        p, pp = _arr(p)
        f = C.c_double()
        g = np.zeros(d["m"])
        self.lib.cmpc_oracle_fg(C.byref(cfg), xp, pp, C.cast(C.byref(f), _dp), g.ctypes.data_as(_dp))
        return f.value, g

    def jac_fg(self, cfg, x, p):
        d = self.dims(cfg.N)
        x, xp = _arr(x)
        p, pp = _arr(p)
        f = C.c_double()
        grad, g, jnz = np.zeros(d["n"]), np.zeros(d["m"]), np.zeros(d["nnz_j"])
        self.lib.cmpc_oracle_jac_fg(C.byref(cfg), xp, pp, C.cast(C.byref(f), _dp), grad.ctypes.data_as(_dp),
                                    g.ctypes.data_as(_dp), jnz.ctypes.data_as(_dp))
        return f.value, grad, g, jnz

    def hess_l(self, cfg, x, p, lam_f, lam_g):
        d = self.dims(cfg.N)
        x, xp = _arr(x)
        p, pp = _arr(p)
        lam_g, lp = _arr(lam_g)
        hnz = np.zeros(d["nnz_h"])
        self.lib.cmpc_oracle_hess_l(C.byref(cfg), xp, pp, lam_f, lp, hnz.ctypes.data_as(_dp))
        return hnz

    # solver
    def default_opts(self, **kw) -> IpmOpts:
        o = IpmOpts()
        self.lib.cmpc_oracle_ipm_default_opts(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def solve(self, cfg, p, lbg, ubg, x0, lam_g0=None, opts=None):
        d = self.dims(cfg.N)
        opts = opts or self.default_opts()
        p, pp = _arr(p)
        lbg, lp = _arr(lbg)
        ubg, up = _arr(ubg)
        x = np.array(x0, dtype=np.float64, copy=True)
        lam = np.zeros(d["m"]) if lam_g0 is None else np.array(lam_g0, dtype=np.float64, copy=True)
        st = IpmStats()
        self.lib.cmpc_oracle_ipm_solve(C.byref(cfg), C.byref(opts), pp, lp, up, x.ctypes.data_as(_dp),
                                       lam.ctypes.data_as(_dp), C.byref(st))
        return x, lam, st

    def solve_batch(self, cfg, p, lbg, ubg, x0, lam_g0=None, opts=None, threads=1):
        opts = opts or self.default_opts()
        p, pp = _arr(p)
        lbg, lp = _arr(lbg)
        ubg, up = _arr(ubg)
        B = p.shape[0]
        x = np.array(x0, dtype=np.float64, copy=True)
        lam = np.zeros_like(lbg) if lam_g0 is None else np.array(lam_g0, dtype=np.float64, copy=True)
        stats = (IpmStats * B)()
        self.lib.cmpc_oracle_ipm_solve_batch(C.byref(cfg), C.byref(opts), B, threads, pp, lp, up,
                                             x.ctypes.data_as(_dp), lam.ctypes.data_as(_dp), stats)
        return x, lam, stats


class RefNLP:
    """The reference's CasADi-generated functions (N = 12), compiled into oracle/_ref by oracle/Makefile.

    which = "tmp" (weights com (10,10,200), h 100, pos 200, sym 10, rate 10) or "jit" (com (10,100,200), sym 100).
    CasADi C ABI: int fn(const double** arg, double** res, long long* iw, double* w, int mem)  (tmp.c:12352).
    """
    N = 12

    def __init__(self, which: str = "tmp"):
        path = os.path.join(HERE, "_ref", f"libref_{which}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = L = C.CDLL(path)
        self.n, self.np_, self.m = 555, 627, 651
        for name in ("nlp_fg", "nlp_jac_fg", "nlp_hess_l"):
            fn = getattr(L, name)
            fn.argtypes = [C.POINTER(_dp), C.POINTER(_dp), C.POINTER(C.c_longlong), _dp, C.c_int]
            fn.restype = C.c_int
            sp = getattr(L, name + "_sparsity_out")
            sp.argtypes, sp.restype = [C.c_longlong], C.POINTER(C.c_longlong)
        self.jc, self.jr = self._sparsity("nlp_jac_fg", 3)
        self.hc, self.hr = self._sparsity("nlp_hess_l", 0)

    def _sparsity(self, fn, i):
        sp = getattr(self.lib, fn + "_sparsity_out")(i)
        nrow, ncol = sp[0], sp[1]
        colind = np.array([sp[2 + j] for j in range(ncol + 1)], dtype=np.int32)
        nnz = int(colind[-1])
        row = np.array([sp[2 + ncol + 1 + j] for j in range(nnz)], dtype=np.int32)
        return colind, row

    def _call(self, name, args, outs):
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in args]
        argv = (_dp * len(keep))(*[a.ctypes.data_as(_dp) for a in keep])
        resv = (_dp * len(outs))(*[o.ctypes.data_as(_dp) for o in outs])
        rc = getattr(self.lib, name)(argv, resv, None, None, 0)
        assert rc == 0

    def fg(self, x, p):
        f, g = np.zeros(1), np.zeros(self.m)
        self._call("nlp_fg", [x, p], [f, g])
        return f[0], g

    def jac_fg(self, x, p):
        f, grad, g, jnz = np.zeros(1), np.zeros(self.n), np.zeros(self.m), np.zeros(int(self.jc[-1]))
        self._call("nlp_jac_fg", [x, p], [f, grad, g, jnz])
        return f[0], grad, g, jnz

    def hess_l(self, x, p, lam_f, lam_g):
        hnz = np.zeros(int(self.hc[-1]))
        self._call("nlp_hess_l", [x, p, np.array([lam_f]), lam_g], [hnz])
        return hnz
