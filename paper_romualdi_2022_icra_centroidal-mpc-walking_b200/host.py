"""ctypes binding of libcmpc_host.so: the C++ host operator BipedalLocomotion::ReducedModelControllers::CentroidalMPC
(host/BipedalLocomotion/CentroidalMPC.h) through its flat C entry points (host/capi.cpp).

The object behind a `CentroidalMPCHost` is the very class a C++ integrator links in place of BLF's CentroidalMPC; Python only
drives it for tests and examples.  advance() needs a GPU (it calls cmpc_solve_host of libcmpc_b200.so); initialisation and
solver_inputs() are host-only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import CmpcConfig, load_library
from .layout import Layout

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, "libcmpc_host.so")
_dp = C.POINTER(C.c_double)
_host = None


def load_host_library() -> C.CDLL:
    global _host
    if _host is not None:
        return _host
    if not os.path.exists(HOST_LIB_PATH):
        raise RuntimeError(f"{HOST_LIB_PATH} is missing: run `python __graft_entry__.py`")
    load_library()  # libcmpc_b200.so first (the host library links against it)
    L = C.CDLL(HOST_LIB_PATH)
    vp, i, d, s = C.c_void_p, C.c_int, C.c_double, C.c_char_p
    L.cmpch_create.argtypes, L.cmpch_create.restype = [s, s], vp
    L.cmpch_destroy.argtypes = [vp]
    L.cmpch_horizon.argtypes = [vp]
    L.cmpch_sampling_time.argtypes, L.cmpch_sampling_time.restype = [vp], d
    L.cmpch_current_time.argtypes, L.cmpch_current_time.restype = [vp], d
    L.cmpch_config.argtypes = [vp, C.POINTER(CmpcConfig)]
    L.cmpch_last_error.argtypes, L.cmpch_last_error.restype = [vp], s
    L.cmpch_set_state.argtypes = [vp, _dp, _dp, _dp, _dp]
    L.cmpch_set_reference.argtypes = [vp, i, _dp, _dp]
    L.cmpch_set_contact_list.argtypes = [vp, s, i, _dp, _dp, _dp, _dp]
    L.cmpch_commit_contacts.argtypes = [vp, d]
    L.cmpch_get_inputs.argtypes = [vp, _dp, _dp, _dp, _dp]
    L.cmpch_get_tick.argtypes = [vp, _dp]
    L.cmpch_advance.argtypes = [vp]
    L.cmpch_advance_batch.argtypes = [C.POINTER(vp), i]
    L.cmpch_is_output_valid.argtypes = [vp]
    L.cmpch_get_contact_output.argtypes = [vp, s, _dp, _dp, _dp]
    L.cmpch_get_next_planned_contact.argtypes = [vp, s, _dp, _dp]
    L.cmpch_get_trajectories.argtypes = [vp, _dp, _dp, _dp]
    L.cmpch_get_stats.argtypes = [vp, C.POINTER(i), C.POINTER(i), _dp]
    L.cmpch_get_output_contact_list.argtypes = [vp, s, i, _dp, _dp, _dp]
    L.cmpch_desired_zmp.argtypes = [vp, _dp]
    L.cmpch_commit_contacts_merged.argtypes = [vp, d, i]
    L.cmpch_resample_linear.argtypes = [i, _dp, _dp, i, _dp, _dp]
    _host = L
    return L


def _p(a):
    return a.ctypes.data_as(_dp)


class CentroidalMPCHost:
    """One C++ CentroidalMPC object initialised from an ini file (group_path like "TRAJECTORY_ADJUSTMENT/CENTROIDAL_MPC")."""

    def __init__(self, ini_path: str, group_path: str = ""):
        self.lib = load_host_library()
        self.h = self.lib.cmpch_create(ini_path.encode(), group_path.encode())
        if not self.h:
            raise RuntimeError(f"CentroidalMPC::initialize failed for {ini_path} [{group_path}]")
        self.N = self.lib.cmpch_horizon(self.h)
        self.dT = self.lib.cmpch_sampling_time(self.h)
        self.L = Layout(self.N)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmpch_destroy(self.h)
            self.h = None

    __del__ = close

    def config(self) -> CmpcConfig:
        c = CmpcConfig()
        assert self.lib.cmpch_config(self.h, C.byref(c)) == 0
        return c

    def current_time(self) -> float:
        return self.lib.cmpch_current_time(self.h)

    def last_error(self) -> str:
        return self.lib.cmpch_last_error(self.h).decode()

    def set_state(self, com, dcom, h, wrench=None) -> bool:
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (com, dcom, h)]
        w = None if wrench is None else np.ascontiguousarray(wrench, dtype=np.float64)
        return self.lib.cmpch_set_state(self.h, _p(a[0]), _p(a[1]), _p(a[2]), None if w is None else _p(w)) == 0

    def set_reference_trajectory(self, com, h) -> bool:
        com = np.ascontiguousarray(com, dtype=np.float64).reshape(-1, 3)
        h = np.ascontiguousarray(h, dtype=np.float64).reshape(-1, 3)
        return self.lib.cmpch_set_reference(self.h, min(len(com), len(h)), _p(com), _p(h)) == 0

    def set_contact_phase_list(self, lists: dict, force_sample_time: float = 0.0) -> bool:
        """lists: name -> list of (t_on, t_off, (x, y, z), yaw)"""
        for name, contacts in lists.items():
            n = len(contacts)
            t_on = np.array([c[0] for c in contacts], dtype=np.float64)
            t_off = np.array([c[1] for c in contacts], dtype=np.float64)
            pos = np.array([c[2] for c in contacts], dtype=np.float64).reshape(n, 3)
            yaw = np.array([c[3] if len(c) > 3 else 0.0 for c in contacts], dtype=np.float64)
            if self.lib.cmpch_set_contact_list(self.h, name.encode(), n, _p(t_on), _p(t_off), _p(pos), _p(yaw)) != 0:
                return False
        return self.lib.cmpch_commit_contacts(self.h, float(force_sample_time)) == 0

    def set_planner_contact_lists(self, lists: dict, force_sample_time: float = 0.0, first_run: bool = False) -> int:
        """per-tick sequence of the reference's block: planner lists merged with the MPC's own output (updateContactPhaseList)"""
        for name, contacts in lists.items():
            n = len(contacts)
            t_on = np.array([c[0] for c in contacts], dtype=np.float64)
            t_off = np.array([c[1] for c in contacts], dtype=np.float64)
            pos = np.array([c[2] for c in contacts], dtype=np.float64).reshape(n, 3)
            yaw = np.array([c[3] if len(c) > 3 else 0.0 for c in contacts], dtype=np.float64)
            if self.lib.cmpch_set_contact_list(self.h, name.encode(), n, _p(t_on), _p(t_off), _p(pos), _p(yaw)) != 0:
                return -3
        return self.lib.cmpch_commit_contacts_merged(self.h, float(force_sample_time), int(first_run))

    def desired_zmp(self):
        z = np.zeros(2)
        return z if self.lib.cmpch_desired_zmp(self.h, _p(z)) == 0 else None

    def solver_inputs(self):
        p, lbg, ubg, x0 = np.zeros(self.L.np), np.zeros(self.L.m), np.zeros(self.L.m), np.zeros(self.L.n)
        if self.lib.cmpch_get_inputs(self.h, _p(p), _p(lbg), _p(ubg), _p(x0)) != 0:
            raise RuntimeError(self.last_error())
        return p, lbg, ubg, x0

    def tick_record(self):
        """the compact tick record advance() uploads (include/cmpc_b200.h): 6 N + 194 doubles"""
        t = np.zeros(6 * self.N + 194)
        if self.lib.cmpch_get_tick(self.h, _p(t)) != 0:
            raise RuntimeError(self.last_error())
        return t

    def advance(self) -> bool:
        return self.lib.cmpch_advance(self.h) == 0

    @staticmethod
    def advance_batch(hosts) -> bool:
        arr = (C.c_void_p * len(hosts))(*[h.h for h in hosts])
        return hosts[0].lib.cmpch_advance_batch(arr, len(hosts)) == 0

    def is_output_valid(self) -> bool:
        return bool(self.lib.cmpch_is_output_valid(self.h))

    def contact_output(self, name):
        pos, rot, f = np.zeros(3), np.zeros(9), np.zeros(12)
        if self.lib.cmpch_get_contact_output(self.h, name.encode(), _p(pos), _p(rot), _p(f)) != 0:
            raise KeyError(name)
        return pos, rot.reshape(3, 3).T, f.reshape(4, 3)

    def next_planned_contact(self, name):
        pos, t = np.zeros(3), C.c_double()
        rc = self.lib.cmpch_get_next_planned_contact(self.h, name.encode(), _p(pos), C.byref(t))
        return None if rc != 0 else (pos, t.value)

    def trajectories(self):
        com, dcom, h = (np.zeros((self.N + 1, 3)) for _ in range(3))
        self.lib.cmpch_get_trajectories(self.h, _p(com), _p(dcom), _p(h))
        return com, dcom, h

    def stats(self):
        st, it, obj = C.c_int(), C.c_int(), C.c_double()
        self.lib.cmpch_get_stats(self.h, C.byref(st), C.byref(it), C.byref(obj))
        return st.value, it.value, obj.value

    def output_contact_list(self, name, cap=64):
        t_on, t_off, pos = np.zeros(cap), np.zeros(cap), np.zeros((cap, 3))
        n = self.lib.cmpch_get_output_contact_list(self.h, name.encode(), cap, _p(t_on), _p(t_off), _p(pos))
        return [(t_on[i], t_off[i], pos[i].copy()) for i in range(max(n, 0))]


def resample_linear(t_in, p_in, t_out):
    """CentroidalMPCWalking::resampleLinear (the reference's LinearSpline frequency adapter)"""
    L = load_host_library()
    t_in = np.ascontiguousarray(t_in, dtype=np.float64)
    p_in = np.ascontiguousarray(p_in, dtype=np.float64).reshape(-1, 3)
    t_out = np.ascontiguousarray(t_out, dtype=np.float64)
    out = np.zeros((len(t_out), 3))
    if L.cmpch_resample_linear(len(t_in), _p(t_in), _p(p_in), len(t_out), _p(t_out), _p(out)) != 0:
        raise ValueError("resampleLinear: inconsistent input")
    return out


def walk_contact_lists(phase: int, dT=0.1, n_steps=12, step_length=0.1, foot_y=0.08, ds_time=0.3, ss_time=0.5, t_past=-100.0):
    """The walk schedule W of workloads.walk_batch as BLF-style contact lists, times relative to the instance's knot 0
    (= global knot `phase`).  Right foot swings first."""
    ds, ss = int(round(ds_time / dT)), int(round(ss_time / dT))
    P = 2 * (ds + ss)
    left, right = [], []
    for s in range(n_steps):
        on, off = s * P, s * P + P - ss
        left.append(((on - phase) * dT if s > 0 else t_past, (off - phase) * dT, (2 * step_length * s, foot_y, 0.0), 0.0))
        on, off = (s - 1) * P + ds + ss, s * P + ds
        x = 0.0 if s == 0 else step_length * (2 * s - 1)
        right.append(((on - phase) * dT if s > 0 else t_past, (off - phase) * dT, (x, -foot_y, 0.0), 0.0))
    keep = lambda lst: [c for c in lst if c[1] > -1e-9]  # noqa: E731  contacts that ended before knot 0 - dT are history
    return {"left_foot": left, "right_foot": right}
