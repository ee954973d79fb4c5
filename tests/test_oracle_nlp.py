"""Pin the CPU oracle's NLP restatement (oracle/cmpc_oracle_nlp.c) against the reference's generated code:
   * committed golden vectors produced from oracle/_ref (tests/golden/make_golden.py),
   * the known-answer numbers of SURVEY.md section 4 (reference tmp.c / jit_tmpComMiH.c),
   * the compiled reference itself when oracle/_ref is present.
Tolerances: the two sides sum the same terms in different orders -> relative 1e-13."""
import os

import numpy as np
import pytest

from conftest import have_ref
from oracle.oracle import RefNLP, make_cfg

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "nlp_kat.npz"))
VARIANTS = {"tmp": {}, "jit": dict(w_com=(10.0, 100.0, 200.0), w_sym=100.0)}

# SURVEY.md section 4 table (f, |g|, g[87], g[231], |grad f|, |J|_F, |H|_F)
KAT = {
    "tmp": (6.932149043740876e+04, 4.655521758023571e+01, -5.207176360868204e-03, 2.169360698869973e+00,
            4.927067035334469e+04, 4.179579374421389e+01, 1.123111411068608e+05),
    "jit": (8.440790874750898e+04, 4.655521758023571e+01, -5.207176360868204e-03, 2.169360698869973e+00,
            4.935573090696579e+04, 4.179579374421389e+01, 1.125093631803559e+05),
}


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1.0, np.max(np.abs(b)))


def test_dims(oracle):
    for N, n, npar, m, nj, nh in ((10, 465, 527, 545, 2445, 3444), (12, 555, 627, 651, 2931, 4140),
                                  (15, 690, 777, 810, 3660, 5184), (50, 2265, 2527, 2665, 12165, 17364)):
        d = oracle.dims(N)
        assert (d["n"], d["np"], d["m"], d["nnz_j"], d["nnz_h"]) == (n, npar, m, nj, nh)
        jc, jr = oracle.jac_sparsity(N)
        hc, hr = oracle.hess_sparsity(N)
        assert jc[-1] == nj and hc[-1] == nh           # no duplicate structural entries
        assert np.all(np.diff(jc) >= 0) and np.all(np.diff(hc) >= 0)


def test_friction_matrix_bit_exact(oracle):
    # literal values probed from the reference's jacobian (tmp.c:8599,8602), mu = 0.33, one slice
    A = oracle.friction_matrix(0.33)
    ref = np.array([[1.0000000000000002, 1.0, -0.33000000000000007], [-0.9999999999999999, 1.0, -0.32999999999999996],
                    [-1.0000000000000004, -1.0, -0.33000000000000007], [0.9999999999999998, -1.0, -0.32999999999999996]])
    assert np.array_equal(A, ref)


@pytest.mark.parametrize("variant", ["tmp", "jit"])
def test_sparsity_matches_reference(oracle, variant):
    jc, jr = oracle.jac_sparsity(12)
    hc, hr = oracle.hess_sparsity(12)
    assert np.array_equal(jc, GOLD[f"{variant}_jc"]) and np.array_equal(jr, GOLD[f"{variant}_jr"])
    assert np.array_equal(hc, GOLD[f"{variant}_hc"]) and np.array_equal(hr, GOLD[f"{variant}_hr"])


@pytest.mark.parametrize("variant", ["tmp", "jit"])
def test_golden_vectors(oracle, variant):
    cfg = make_cfg(**VARIANTS[variant])
    for i in range(int(GOLD["ncases"])):
        x, p, lam, lf = (GOLD[f"{variant}_{i}_{k}"] for k in ("x", "p", "lam", "lamf"))
        f, grad, g, jnz = oracle.jac_fg(cfg, x, p)
        f2, g2 = oracle.fg(cfg, x, p)
        h = oracle.hess_l(cfg, x, p, float(lf), lam)
        assert abs(f - GOLD[f"{variant}_{i}_f"]) <= 1e-13 * abs(f)
        assert f2 == f and np.array_equal(g, g2)
        assert rel(g, GOLD[f"{variant}_{i}_g"]) < 1e-13
        assert rel(grad, GOLD[f"{variant}_{i}_grad"]) < 1e-13
        assert rel(jnz, GOLD[f"{variant}_{i}_jnz"]) < 1e-13
        assert rel(h, GOLD[f"{variant}_{i}_hnz"]) < 1e-13


@pytest.mark.parametrize("variant", ["tmp", "jit"])
def test_known_answers_survey(oracle, variant):
    cfg = make_cfg(**VARIANTS[variant])
    x = np.sin(0.1 * np.arange(555)); p = np.cos(0.05 * np.arange(627)); lam = np.sin(0.3 * np.arange(651))
    f, grad, g, jnz = oracle.jac_fg(cfg, x, p)
    h = oracle.hess_l(cfg, x, p, 1.0, lam)
    got = (f, np.linalg.norm(g), g[87], g[231], np.linalg.norm(grad), np.linalg.norm(jnz), np.linalg.norm(h))
    for a, b in zip(got, KAT[variant]):
        assert abs(a - b) <= 1e-12 * max(1.0, abs(b))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("variant", ["tmp", "jit"])
def test_against_compiled_reference(oracle, variant):
    R = RefNLP(variant)
    cfg = make_cfg(**VARIANTS[variant])
    rng = np.random.default_rng(123)
    for _ in range(10):
        x = rng.normal(size=555); p = rng.normal(size=627); lam = 100 * rng.normal(size=651)
        f, grad, g, jnz = R.jac_fg(x, p)
        h = R.hess_l(x, p, 1.0, lam)
        f2, grad2, g2, jnz2 = oracle.jac_fg(cfg, x, p)
        h2 = oracle.hess_l(cfg, x, p, 1.0, lam)
        assert abs(f - f2) <= 1e-13 * abs(f)
        assert rel(g2, g) < 1e-13 and rel(grad2, grad) < 1e-13 and rel(jnz2, jnz) < 1e-13 and rel(h2, h) < 1e-13


def test_general_horizon_finite_differences(oracle):
    """N != 12 has no compiled reference: check the restated derivatives against central differences."""
    for N in (5, 15):
        cfg = make_cfg(N=N)
        d = oracle.dims(N)
        rng = np.random.default_rng(N)
        x = rng.normal(size=d["n"]); p = rng.normal(size=d["np"]); lam = rng.normal(size=d["m"])
        f, grad, g, jnz = oracle.jac_fg(cfg, x, p)
        jc, jr = oracle.jac_sparsity(N)
        hc, hr = oracle.hess_sparsity(N)
        h = oracle.hess_l(cfg, x, p, 1.0, lam)
        J = np.zeros((d["m"], d["n"])); H = np.zeros((d["n"], d["n"]))
        for c in range(d["n"]):
            J[jr[jc[c]:jc[c + 1]], c] = jnz[jc[c]:jc[c + 1]]
            H[hr[hc[c]:hc[c + 1]], c] = h[hc[c]:hc[c + 1]]
        assert np.allclose(H, H.T, atol=1e-12)
        eps = 1e-6
        for c in rng.choice(d["n"], size=40, replace=False):
            e = np.zeros(d["n"]); e[c] = eps
            fp, gradp, gp, jp = oracle.jac_fg(cfg, x + e, p)
            fm, gradm, gm, jm = oracle.jac_fg(cfg, x - e, p)
            assert abs((fp - fm) / (2 * eps) - grad[c]) < 4e-15 * abs(f) / eps + 1e-6 * abs(grad[c])  # roundoff of f / eps
            assert np.max(np.abs((gp - gm) / (2 * eps) - J[:, c])) < 1e-6
            Jp = np.zeros_like(J); Jm = np.zeros_like(J)
            for cc in range(d["n"]):
                Jp[jr[jc[cc]:jc[cc + 1]], cc] = jp[jc[cc]:jc[cc + 1]]
                Jm[jr[jc[cc]:jc[cc + 1]], cc] = jm[jc[cc]:jc[cc + 1]]
            hcol = (gradp - gradm) / (2 * eps) + ((Jp - Jm) / (2 * eps)).T @ lam
            assert np.max(np.abs(hcol - H[:, c])) < 1e-4 * max(1.0, np.max(np.abs(H[:, c])))


@pytest.mark.parametrize("mehrotra", [0, 1])
def test_oracle_ipm_converges_at_the_ini_tolerances(oracle, workloads, mehrotra):
    """ipopt_tolerance 1e-2 (ergoCubSN000/SN001 ini) and 1e-4 (ergoCubGazeboV1_1 ini): the barrier parameter must go down to
    min(tol, compl_inf_tol) / 11, otherwise compl_inf_tol = 1e-4 can never be met"""
    from oracle.oracle import make_cfg
    w = workloads.walk_batch(N=12, B=6, seed=2, state_noise=2.0, yaw_range=0.3)
    tight = oracle.solve_batch(make_cfg(w_pos=2000.0), w["p"], w["lbg"], w["ubg"], w["x0"], threads=3)[2]
    for tol in (1e-2, 1e-4):
        x, lam, st = oracle.solve_batch(make_cfg(w_pos=2000.0), w["p"], w["lbg"], w["ubg"], w["x0"], threads=3,
                                        opts=oracle.default_opts(tol=tol, mehrotra=mehrotra))
        for b in range(6):
            assert st[b].status == 0 and st[b].iters < tight[b].iters
            assert abs(st[b].obj - tight[b].obj) <= 1e-3 * abs(tight[b].obj)
