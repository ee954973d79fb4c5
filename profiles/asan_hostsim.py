"""AddressSanitizer + UBSan over the solver source itself (CPU, no GPU needed): tests/hostsim compiles csrc/cmpc_ipm.cuh and
csrc/cmpc_warp.cuh -- the code the kernels run -- for the host with 32 / 128 emulated lanes; the per-instance scratch arena is a
heap block of exactly works_doubles(N) doubles and the shared-memory image a static ISmem, so any out-of-range index of the
stage-major vectors, the factor blocks or the shared-memory arrays is reported.  (compute-sanitizer is closed on the GPU pool of
this round, see profiles/r2_notes.md.)
usage, from the repo root:
  g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -x c++ tests/hostsim/hostsim.cpp -o /tmp/libhostsim_asan.so
  LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 python profiles/asan_hostsim.py"""
import ctypes as C, sys, os, importlib, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import test_hostsim as T
from oracle import oracle as om
L = C.CDLL('/tmp/libhostsim_asan.so')
assert L.hostsim_config_size() == C.sizeof(T.DevConfig)
wl = importlib.import_module('paper_romualdi_2022_icra_centroidal-mpc-walking_b200.workloads')
for N, kw, wkw in [(12, {}, dict(state_noise=1.0, yaw_range=0.2)), (15, dict(w_com=(1.0,1.0,200.0), w_pos=200.0, w_sym=0.0), dict(state_noise=1.0, step_adjust=False)), (2, {}, dict(state_noise=0.5))]:
    for pc in (0, 1):
        c, o = T.dev_config(N, pc=pc, **kw)
        w = wl.walk_batch(N=N, B=2, seed=5, **wkw)
        for fn in ('hostsim_solve_team32', 'hostsim_solve_team128'):
            for b in range(2):
                st, it, obj, x, lam = T.run(L, fn, c, w, b)
                print(N, pc, fn, b, 'status', st, 'it', it)
                assert st == 0
print('asan/ubsan run complete')
