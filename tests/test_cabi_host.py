"""CPU-side checks of the product library: it loads, exports every symbol declared in include/cmpc_b200.h, its
host-only helpers agree with the oracle, and it refuses to work without a GPU instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, pkg


@pytest.fixture(scope="module")
def lib():
    b = pkg("build")
    b.build()
    return pkg().load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "cmpc_b200.h")).read()
    names = set(re.findall(r"\b(cmpc_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n


def test_sparsity_and_friction_match_oracle(lib, oracle):
    P = pkg()
    for N in (2, 12, 15, 22, 50):
        jc, jr, hc, hr = P.sparsity(N)
        ojc, ojr = oracle.jac_sparsity(N)
        ohc, ohr = oracle.hess_sparsity(N)
        assert np.array_equal(jc, ojc) and np.array_equal(jr, ojr) and np.array_equal(hc, ohc) and np.array_equal(hr, ohr)
    A = np.zeros(12)
    assert lib.cmpc_friction_matrix(0.33, 1, A.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert np.array_equal(A.reshape(4, 3), oracle.friction_matrix(0.33))
    assert lib.cmpc_friction_matrix(0.33, 2, A.ctypes.data_as(C.POINTER(C.c_double))) == -1   # only one slice supported


def test_dims(lib):
    v = [C.c_int() for _ in range(5)]
    assert lib.cmpc_dims(15, *[C.byref(a) for a in v]) == 0
    assert [a.value for a in v] == [690, 777, 810, 3660, 5184]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    P = pkg()
    cfg = P.default_config()
    h = C.c_void_p()
    assert lib.cmpc_create(C.byref(cfg), C.byref(h)) == -3          # CMPC_E_NO_DEVICE
    with pytest.raises(RuntimeError):
        P.BatchedCentroidalMPC(cfg)


def test_product_never_touches_the_oracle():
    """no include / import / dlopen of anything under oracle/ in the product package (comments may cite it)"""
    pat = re.compile(r'#include\s*[<"][^>"]*oracle|from\s+oracle|import\s+oracle|libcmpc_oracle|oracle/_ref|libref_')
    pdir = os.path.join(ROOT, "paper_romualdi_2022_icra_centroidal-mpc-walking_b200")
    for dirpath, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), (dirpath, f)
