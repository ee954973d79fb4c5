"""B200-native batched centroidal-MPC solve: Python binding of the C ABI (include/cmpc_b200.h).

The product is libcmpc_b200.so (hand-written sm_100a kernels).  This module only moves pointers: torch is used for
device memory and streams.  There is NO CPU path: constructing a solver without the CUDA library or without a GPU
raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .layout import Layout  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMPC_B200_LIB", os.path.join(HERE, "libcmpc_b200.so"))  # override: A/B builds under profiles/_build

STATUS_NAMES = {0: "converged", 1: "max_iter", 2: "line_search", 3: "numerical", 4: "bad_input", 5: "acceptable"}


class CmpcConfig(C.Structure):
    """struct cmpc_config of include/cmpc_b200.h"""
    _fields_ = [("horizon", C.c_int), ("sampling_time", C.c_double), ("number_of_slices", C.c_int),
                ("static_friction_coefficient", C.c_double), ("com_weight", C.c_double * 3),
                ("contact_position_weight", C.c_double), ("force_rate_of_change_weight", C.c_double * 3),
                ("angular_momentum_weight", C.c_double), ("contact_force_symmetry_weight", C.c_double),
                ("corners", C.c_double * 24), ("ipopt_tolerance", C.c_double), ("ipopt_max_iteration", C.c_int),
                ("mu_init", C.c_double), ("bound_relax_factor", C.c_double), ("bound_push", C.c_double),
                ("infinity", C.c_double), ("device", C.c_int), ("threads_per_instance", C.c_int),
                ("ctas_per_sm", C.c_int), ("teams_per_cta", C.c_int), ("lockstep_groups", C.c_int),
                ("mu_strategy", C.c_int), ("warm_start_mu_init", C.c_double), ("nlp_scaling_max_gradient", C.c_double),
                ("acceptable_tol", C.c_double), ("acceptable_iter", C.c_int),
                ("bounding_box_upper_limit", C.c_double * 6), ("bounding_box_lower_limit", C.c_double * 6)]


class WalkParams(C.Structure):
    """struct cmpc_walk_params of include/cmpc_b200.h: the synthetic planner of the closed loop"""
    _fields_ = [("ds_knots", C.c_int), ("ss_knots", C.c_int), ("step_length", C.c_double), ("com_height", C.c_double),
                ("push_threshold", C.c_double), ("zmp_half_length", C.c_double), ("zmp_half_width", C.c_double)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None

# cmpc_config.mu_strategy (include/cmpc_b200.h)
MU_DEFAULT, MU_MONOTONE, MU_MEHROTRA = 0, 1, 2


def load_library() -> C.CDLL:
    """dlopen libcmpc_b200.so (built in-tree by build.py / __graft_entry__.build()); fails loudly when missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (nvcc, sm_100a). "
                           "This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.cmpc_default_config.argtypes = [C.POINTER(CmpcConfig)]
    L.cmpc_dims.argtypes = [i, _ip, _ip, _ip, _ip, _ip]
    L.cmpc_friction_matrix.argtypes = [d, i, _dp]
    L.cmpc_jac_sparsity.argtypes = [i, _ip, _ip]
    L.cmpc_hess_sparsity.argtypes = [i, _ip, _ip]
    L.cmpc_create.argtypes = [C.POINTER(CmpcConfig), C.POINTER(vp)]
    L.cmpc_destroy.argtypes = [vp]
    L.cmpc_solve_batched.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, i, vp]
    L.cmpc_solve_host.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, i]
    L.cmpc_shift_warmstart.argtypes = [vp, i, vp, vp, vp]
    L.cmpc_eval_fg.argtypes = [vp, i, vp, vp, vp, vp, vp]
    L.cmpc_eval_jac_fg.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp]
    L.cmpc_eval_hess_l.argtypes = [vp, i, vp, vp, d, vp, vp, vp]
    L.cmpc_rollout_plant.argtypes = [vp, i, vp, vp, vp, vp, d, i, vp]
    L.cmpc_tick_stride.argtypes = [i]
    L.cmpc_populate.argtypes = [vp, i, vp, vp, vp, vp, vp, vp]
    L.cmpc_solve_ticks_host.argtypes = [vp, i, vp, i, vp, vp, vp, vp, vp]
    L.cmpc_resample_references.argtypes = [vp, i, i, vp, vp, vp, vp, d, d, vp, vp]
    L.cmpc_desired_zmp.argtypes = [vp, i, vp, vp, d, d, vp, vp, vp]
    L.cmpc_rollout_layout.argtypes = [_ip, _ip, _ip]
    L.cmpc_rollout_tick.argtypes = [vp, C.POINTER(WalkParams), i, i, vp, vp, vp, vp, vp, i, vp]
    L.cmpc_rollout_feedback.argtypes = [vp, C.POINTER(WalkParams), i, i, vp, vp, vp, vp, vp, vp, vp, vp]
    L.cmpc_measure_fp64_peak.argtypes = [vp, _dp]
    L.cmpc_launch_count.argtypes = [vp]
    L.cmpc_launch_count.restype = C.c_longlong
    L.cmpc_last_cuda_error.argtypes = [vp]
    L.cmpc_error_string.argtypes = [i]
    L.cmpc_error_string.restype = C.c_char_p
    L.cmpc_solver_geometry.argtypes = [vp, _ip, _ip, _ip, _ip, _ip]
    L.cmpc_solver_lockstep.argtypes = [vp, _ip, _ip]
    _lib = L
    return L


def default_config(**overrides) -> CmpcConfig:
    cfg = CmpcConfig()
    load_library().cmpc_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if k in ("com_weight", "force_rate_of_change_weight"):
            getattr(cfg, k)[:] = list(v)
        elif k in ("bounding_box_upper_limit", "bounding_box_lower_limit"):
            getattr(cfg, k)[:] = np.asarray(v, dtype=np.float64).reshape(-1).tolist()
        elif k == "corners":
            cfg.corners[:] = np.asarray(v, dtype=np.float64).reshape(-1).tolist()
        else:
            setattr(cfg, k, v)
    return cfg


# the two robots of BASELINE.json's configs (values of config/robots/<robot>/centroidal_mpc.ini)
def icub3_config(**kw) -> CmpcConfig:
    """iCubGazeboV3/centroidal_mpc.ini: N = 15, dT = 0.1, com (1,1,200), contact position 2e2, no symmetry term."""
    one = [(0.08, 0.03, 0.0), (0.08, -0.03, 0.0), (-0.08, -0.03, 0.0), (-0.08, 0.03, 0.0)]
    base = dict(horizon=15, sampling_time=0.1, com_weight=(1.0, 1.0, 200.0), contact_position_weight=2e2,
                force_rate_of_change_weight=(10.0, 10.0, 10.0), angular_momentum_weight=1e2,
                contact_force_symmetry_weight=0.0, corners=[one, one])
    base.update(kw)
    return default_config(**base)


def ergocub_config(**kw) -> CmpcConfig:
    """ergoCubGazeboV1_1/centroidal_mpc.ini: N = 12, dT = 0.1, com (10,10,200), contact position 2e3, symmetry 10."""
    base = dict(horizon=12, sampling_time=0.1)
    base.update(kw)
    return default_config(**base)


def _check(rc: int, what: str, handle=None):
    if rc != 0:
        L = load_library()
        extra = f" (cuda error {L.cmpc_last_cuda_error(handle)})" if handle and rc == -2 else ""
        raise RuntimeError(f"{what} failed: {L.cmpc_error_string(rc).decode()}{extra}")


class BatchedCentroidalMPC:
    """Thin owner of a cmpc_handle.  Device-pointer entry points take torch CUDA float64 tensors (contiguous)."""

    def __init__(self, cfg: CmpcConfig | None = None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the centroidal-MPC solve has no CPU path in this package")
        self.lib = load_library()
        self.cfg = cfg or default_config()
        self.N = self.cfg.horizon
        self.L = Layout(self.N)
        self.handle = C.c_void_p()
        _check(self.lib.cmpc_create(C.byref(self.cfg), C.byref(self.handle)), "cmpc_create")
        self.device = torch.device("cuda", self.cfg.device)

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.cmpc_destroy(self.handle)
            self.handle = C.c_void_p()

    __del__ = close

    # ---- helpers
    @staticmethod
    def _ptr(t):
        if t is None:
            return None
        import torch
        assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float64, torch.int32), (t.device, t.dtype)
        return C.c_void_p(t.data_ptr())

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def geometry(self):
        v = [C.c_int() for _ in range(5)]
        self.lib.cmpc_solver_geometry(self.handle, *[C.byref(a) for a in v])
        t, g = C.c_int(), C.c_int()
        self.lib.cmpc_solver_lockstep(self.handle, C.byref(t), C.byref(g))
        return dict(grid=v[0].value, threads=v[1].value, smem=v[2].value, ctas_per_sm=v[3].value, sm_count=v[4].value,
                    teams_per_cta=t.value, lockstep_groups=g.value)

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        _check(self.lib.cmpc_measure_fp64_peak(self.handle, C.byref(v)), "cmpc_measure_fp64_peak", self.handle)
        return v.value

    def launch_count(self) -> int:
        return int(self.lib.cmpc_launch_count(self.handle))

    # ---- hot path
    def solve(self, p, lbg, ubg, x, lam_g=None, warm_duals=False, out=None):
        """In place on x (and lam_g).  Returns (obj, status, iters, lam_g) device tensors.  Asynchronous.  out = (obj, status,
        iters) reuses caller-owned result tensors (CUDA-graph capture needs static buffers)."""
        import torch
        B = p.shape[0]
        if out is not None:
            obj, status, iters = out
        else:
            obj = torch.empty(B, dtype=torch.float64, device=self.device)
            status = torch.empty(B, dtype=torch.int32, device=self.device)
            iters = torch.empty(B, dtype=torch.int32, device=self.device)
        if lam_g is None:
            lam_g = torch.zeros(B, self.L.m, dtype=torch.float64, device=self.device)
        _check(self.lib.cmpc_solve_batched(self.handle, B, self._ptr(p), self._ptr(lbg), self._ptr(ubg), self._ptr(x),
                                           self._ptr(lam_g), self._ptr(obj), self._ptr(status), self._ptr(iters),
                                           int(bool(warm_duals)), self._stream()), "cmpc_solve_batched", self.handle)
        return obj, status, iters, lam_g

    def solve_host(self, p, lbg, ubg, x0, lam_g0=None):
        """numpy in / numpy out through cmpc_solve_host (H2D + solve + D2H + sync inside the library)."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        lbg = np.ascontiguousarray(lbg, dtype=np.float64)
        ubg = np.ascontiguousarray(ubg, dtype=np.float64)
        x = np.array(x0, dtype=np.float64, copy=True, order="C")
        B = p.shape[0] if p.ndim == 2 else 1
        lam = np.zeros((B, self.L.m)) if lam_g0 is None else np.array(lam_g0, dtype=np.float64, copy=True, order="C")
        obj = np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros(B, dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _check(self.lib.cmpc_solve_host(self.handle, B, vp(p), vp(lbg), vp(ubg), vp(x), vp(lam), vp(obj), vp(status),
                                        vp(iters), int(lam_g0 is not None)), "cmpc_solve_host", self.handle)
        return x, lam, obj, status, iters

    def solve_ticks_host(self, ticks, warm_mode=0, x_prev=None, lam_prev=None, want_lam=True):
        """numpy tick records in / numpy solution out through cmpc_solve_ticks_host (the per-tick call of a host controller):
        uploads the records, populates (p, lbg, ubg, x0) on the device, solves, downloads.  warm_mode 1: warm start from the
        solution the previous call left on the device; 2: from x_prev / lam_prev (unshifted)."""
        ticks = np.ascontiguousarray(ticks, dtype=np.float64)
        B = ticks.shape[0]
        assert ticks.shape[1] == self.lib.cmpc_tick_stride(self.N)
        x = np.zeros((B, self.L.n)) if x_prev is None else np.array(x_prev, dtype=np.float64, copy=True, order="C")
        lam = None
        if want_lam or lam_prev is not None:
            lam = np.zeros((B, self.L.m)) if lam_prev is None else np.array(lam_prev, dtype=np.float64, copy=True, order="C")
        obj, status, iters = np.zeros(B), np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _check(self.lib.cmpc_solve_ticks_host(self.handle, B, vp(ticks), int(warm_mode), vp(x), vp(lam), vp(obj), vp(status),
                                              vp(iters)), "cmpc_solve_ticks_host", self.handle)
        return x, lam, obj, status, iters

    def populate(self, ticks, want_x0=True):
        """tick records (device tensor) -> (p, lbg, ubg, x0) device tensors (cmpc_populate)"""
        import torch
        B = ticks.shape[0]
        mk = lambda n: torch.empty(B, n, dtype=torch.float64, device=self.device)  # noqa: E731
        p, lbg, ubg = mk(self.L.np), mk(self.L.m), mk(self.L.m)
        x0 = mk(self.L.n) if want_x0 else None
        self.populate_into(ticks, p, lbg, ubg, x0)
        return p, lbg, ubg, x0

    def populate_into(self, ticks, p, lbg, ubg, x0=None):
        _check(self.lib.cmpc_populate(self.handle, ticks.shape[0], self._ptr(ticks), self._ptr(p), self._ptr(lbg), self._ptr(ubg),
                                      self._ptr(x0), self._stream()), "cmpc_populate", self.handle)

    def resample_references(self, ticks, t_in, com_in, h_in, t_out, robot_mass=1.0, com_height=-1.0):
        _check(self.lib.cmpc_resample_references(self.handle, ticks.shape[0], t_in.shape[0], self._ptr(t_in), self._ptr(com_in),
                                                 self._ptr(h_in), self._ptr(t_out), float(robot_mass), float(com_height),
                                                 self._ptr(ticks), self._stream()), "cmpc_resample_references", self.handle)

    def desired_zmp(self, x, p, half_length=0.08, half_width=0.03):
        import torch
        B = x.shape[0]
        zmp = torch.empty(B, 2, dtype=torch.float64, device=self.device)
        valid = torch.empty(B, dtype=torch.int32, device=self.device)
        _check(self.lib.cmpc_desired_zmp(self.handle, B, self._ptr(x), self._ptr(p), float(half_length), float(half_width),
                                         self._ptr(zmp), self._ptr(valid), self._stream()), "cmpc_desired_zmp", self.handle)
        return zmp, valid

    def rollout_tick(self, wp, tick, roll, state, steps, ticks, ext6, step_adjust=True):
        _check(self.lib.cmpc_rollout_tick(self.handle, C.byref(wp), roll.shape[0], int(tick), self._ptr(roll), self._ptr(state),
                                          self._ptr(steps), self._ptr(ticks), self._ptr(ext6), int(bool(step_adjust)), self._stream()),
               "cmpc_rollout_tick", self.handle)

    def rollout_feedback(self, wp, tick, x, p, state, status, iters, roll, steps):
        _check(self.lib.cmpc_rollout_feedback(self.handle, C.byref(wp), roll.shape[0], int(tick), self._ptr(x), self._ptr(p),
                                              self._ptr(state), self._ptr(status), self._ptr(iters), self._ptr(roll),
                                              self._ptr(steps), self._stream()), "cmpc_rollout_feedback", self.handle)

    def shift_warmstart(self, x, lam_g=None):
        _check(self.lib.cmpc_shift_warmstart(self.handle, x.shape[0], self._ptr(x), self._ptr(lam_g), self._stream()),
               "cmpc_shift_warmstart", self.handle)

    # ---- NLP functions (parity surface)
    def eval_jac_fg(self, x, p, want_jac=True):
        import torch
        B = x.shape[0]
        nj = 243 * self.N + 15
        f = torch.empty(B, dtype=torch.float64, device=self.device)
        grad = torch.empty(B, self.L.n, dtype=torch.float64, device=self.device)
        g = torch.empty(B, self.L.m, dtype=torch.float64, device=self.device)
        jnz = torch.empty(B, nj, dtype=torch.float64, device=self.device) if want_jac else None
        _check(self.lib.cmpc_eval_jac_fg(self.handle, B, self._ptr(x), self._ptr(p), self._ptr(f), self._ptr(grad),
                                         self._ptr(g), self._ptr(jnz), self._stream()), "cmpc_eval_jac_fg", self.handle)
        return f, grad, g, jnz

    def eval_hess_l(self, x, p, lam_f, lam_g):
        import torch
        B = p.shape[0]
        nh = 348 * self.N - 36
        hnz = torch.empty(B, nh, dtype=torch.float64, device=self.device)
        _check(self.lib.cmpc_eval_hess_l(self.handle, B, self._ptr(x), self._ptr(p), float(lam_f), self._ptr(lam_g),
                                         self._ptr(hnz), self._stream()), "cmpc_eval_hess_l", self.handle)
        return hnz

    def rollout_plant(self, x, p, state, dt, substeps, ext=None):
        _check(self.lib.cmpc_rollout_plant(self.handle, x.shape[0], self._ptr(x), self._ptr(p), self._ptr(ext),
                                           self._ptr(state), float(dt), int(substeps), self._stream()),
               "cmpc_rollout_plant", self.handle)


def sparsity(N: int):
    """(jac colind, jac row, hess colind, hess row) in the reference's CasADi CSC order; pure host code."""
    L = load_library()
    n, nj, nh = 45 * N + 15, 243 * N + 15, 348 * N - 36
    jc, jr = np.zeros(n + 1, np.int32), np.zeros(nj, np.int32)
    hc, hr = np.zeros(n + 1, np.int32), np.zeros(nh, np.int32)
    _check(L.cmpc_jac_sparsity(N, jc.ctypes.data_as(_ip), jr.ctypes.data_as(_ip)), "cmpc_jac_sparsity")
    _check(L.cmpc_hess_sparsity(N, hc.ctypes.data_as(_ip), hr.ctypes.data_as(_ip)), "cmpc_hess_sparsity")
    return jc, jr, hc, hr
