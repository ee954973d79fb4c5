"""Generate tests/golden/nlp_kat.npz from the REFERENCE's own generated code (oracle/_ref, built by oracle/Makefile
from /root/reference/src/centroidal-mpc-walking/config/robots/ergoCubGazeboV1/{tmp.c,jit_tmpComMiH.c}).

Run in the dev container (where /root/reference exists):  python tests/golden/make_golden.py
The fixture travels with the repo so that the CPU tests can pin the restated oracle on a machine without the reference.
Contents per variant v in {tmp, jit}: for seeds 0..3 random (x, p, lam_g) -> f, g, grad_f, jac_nz, hess_nz of the
reference; plus the deterministic known-answer inputs of SURVEY.md section 4 (x_i = sin(.1 i), p_i = cos(.05 i), ...).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefNLP  # noqa: E402

out = {}
for v in ("tmp", "jit"):
    R = RefNLP(v)
    out[f"{v}_jc"], out[f"{v}_jr"], out[f"{v}_hc"], out[f"{v}_hr"] = R.jc, R.jr, R.hc, R.hr
    cases = []
    x = np.sin(0.1 * np.arange(555)); p = np.cos(0.05 * np.arange(627)); lam = np.sin(0.3 * np.arange(651))
    cases.append((x, p, lam, 1.0))
    for seed in range(4):
        rng = np.random.default_rng(seed)
        x = rng.normal(size=555); p = rng.normal(size=627); lam = 10 * rng.normal(size=651)
        # realistic parameter pieces: isEnabled in {0,1}, orientation = rotations about z
        for c in range(2):
            base = c * (19 * 12 + 6)
            p[base + 180: base + 192] = rng.integers(0, 2, size=12)
            for k in range(12):
                a = rng.uniform(-0.5, 0.5)
                p[base + 9 * k: base + 9 * k + 9] = [np.cos(a), np.sin(a), 0, -np.sin(a), np.cos(a), 0, 0, 0, 1]
        cases.append((x, p, lam, 0.7 + 0.1 * seed))
    for i, (x, p, lam, lf) in enumerate(cases):
        f, grad, g, jnz = R.jac_fg(x, p)
        f2, g2 = R.fg(x, p)
        assert f == f2 and np.array_equal(g, g2)
        h = R.hess_l(x, p, lf, lam)
        out[f"{v}_{i}_x"], out[f"{v}_{i}_p"], out[f"{v}_{i}_lam"], out[f"{v}_{i}_lamf"] = x, p, lam, np.array(lf)
        out[f"{v}_{i}_f"], out[f"{v}_{i}_g"], out[f"{v}_{i}_grad"] = np.array(f), g, grad
        out[f"{v}_{i}_jnz"], out[f"{v}_{i}_hnz"] = jnz, h
out["ncases"] = np.array(len(cases))
np.savez_compressed(os.path.join(os.path.dirname(__file__), "nlp_kat.npz"), **out)
print("wrote nlp_kat.npz", len(out), "arrays")
