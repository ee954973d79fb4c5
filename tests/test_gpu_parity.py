"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path through the C ABI against the CPU oracle.

Tolerances (BASELINE.json north_star): objective within 1e-6 relative; forces / CoM trajectory / footsteps within 1e-5
scaled inf-norm (max |delta| / max(1, max |ref|) per group); NLP functions (f, g, grad, jac, hess) within 1e-12 relative
of the oracle / golden vectors (same arithmetic in a different summation order).
"""
import os

import numpy as np
import pytest

from conftest import have_ref, pkg

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle.oracle import make_cfg  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "nlp_kat.npz"))
ICUB_ORACLE = dict(N=15, w_com=(1.0, 1.0, 200.0), w_pos=200.0, w_sym=0.0,
                   corners=[[(0.08, 0.03, 0), (0.08, -0.03, 0), (-0.08, -0.03, 0), (-0.08, 0.03, 0)]] * 2)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1.0, np.max(np.abs(b)))


def groups(L):
    N = L.N
    g = {"com": np.arange(0, 3 * (N + 1)), "dcom": np.arange(L.x_dcom(0), L.x_dcom(0) + 3 * (N + 1)),
         "h": np.arange(L.x_h(0), L.x_h(0) + 3 * (N + 1))}
    pos, frc = [], []
    for c in range(2):
        pos += list(range(L.x_pos(c, 0), L.x_pos(c, 0) + 3 * (N + 1)))
        frc += list(range(L.x_frc(c, 0, 0), L.x_frc(c, 0, 0) + 12 * N))
    g["footsteps"] = np.array(pos)
    g["forces"] = np.array(frc)
    return g


ICUB_CORNERS = np.array(ICUB_ORACLE["corners"][0], dtype=np.float64)
ERGOCUB_CORNERS = np.array([(0.08, 0.01, 0), (0.08, -0.01, 0), (-0.08, -0.01, 0), (-0.08, 0.01, 0)], dtype=np.float64)


def foot_wrenches(L, x, p, corners):
    """(N, 2, 6): resultant force and torque about the foot origin of the corner forces of every contact at every knot --
    what the dynamics see of the forces, and at knot 0 what the reference's whole-body layer consumes
    (WholeBodyQPBlock.cpp:1084, 1121: the corner forces are summed into one contact wrench per foot)"""
    N = L.N
    out = np.zeros((N, 2, 6))
    for k in range(N):
        for c in range(2):
            R = p[L.p_rot(c, k):L.p_rot(c, k) + 9].reshape(3, 3).T
            en = p[L.p_en(c, k)]
            for j in range(4):
                f = en * x[L.x_frc(c, j, k):L.x_frc(c, j, k) + 3]
                out[k, c, :3] += f
                out[k, c, 3:] += np.cross(R @ corners[j], f)
    return out


def compare_solutions(L, x, xo, obj, obj_o, forces=True, p=None, corners=None):
    """objective 1e-6 relative, every variable group 1e-5 scaled inf-norm.  forces=False (iCub3: no symmetry weight, the
    split of a foot's wrench over its corners is not unique): the per-foot resultant wrenches of every knot instead"""
    assert abs(obj - obj_o) <= 1e-6 * max(1.0, abs(obj_o)), (obj, obj_o)
    for name, idx in groups(L).items():
        if name == "forces" and not forces:
            continue
        err = np.max(np.abs(x[idx] - xo[idx])) / max(1.0, np.max(np.abs(xo[idx])))
        assert err <= 1e-5, (name, err)
    if not forces:
        wa, wb = foot_wrenches(L, x, p, corners), foot_wrenches(L, xo, p, corners)
        err = np.max(np.abs(wa - wb)) / max(1.0, np.max(np.abs(wb)))
        assert err <= 1e-5, ("per-foot wrench", err)


STRATEGIES = ["mehrotra", "monotone"]   # cmpc_config.mu_strategy; the oracle runs the same barrier update (opts.mehrotra)


def strategy_kw(name):
    P = pkg()
    return dict(mu_strategy=P.MU_MEHROTRA if name == "mehrotra" else P.MU_MONOTONE)


def oracle_opts(oracle, name, **kw):
    return oracle.default_opts(mehrotra=1 if name == "mehrotra" else 0, **kw)


@pytest.fixture(scope="module", params=[(g, s) for s in STRATEGIES for g in ("auto", "lockstep7")], ids=lambda p: f"{p[0]}-{p[1]}")
def solver12(request):
    """default geometry (batches of at most one instance per SM take the single-team latency kernel) and the 7-team lock-step
    kernel forced for every batch size, each with both barrier updates (predictor-corrector = default, monotone = IPOPT's
    path): all four must pass every parity test"""
    P = pkg()
    geom, strat = request.param
    kw = dict(teams_per_cta=7) if geom == "lockstep7" else {}
    s = P.BatchedCentroidalMPC(P.ergocub_config(contact_position_weight=200.0, **strategy_kw(strat), **kw))  # weights of the reference's tmp.c
    s.strategy = strat
    geo = s.geometry()
    assert geo["teams_per_cta"] == 7 and geo["threads"] == 96, geo
    yield s
    s.close()


@pytest.mark.parametrize("variant,kw", [("tmp", {}), ("jit", dict(com_weight=(10.0, 100.0, 200.0), contact_force_symmetry_weight=100.0))])
def test_nlp_functions_match_golden_vectors(variant, kw):
    """f, g, grad f, jac (CSC), hess (CSC) of the CUDA kernels vs the vectors generated from the reference's tmp.c"""
    P = pkg()
    s = P.BatchedCentroidalMPC(P.ergocub_config(contact_position_weight=200.0, **kw))
    nc = int(GOLD["ncases"])
    x = dev(np.stack([GOLD[f"{variant}_{i}_x"] for i in range(nc)]))
    p = dev(np.stack([GOLD[f"{variant}_{i}_p"] for i in range(nc)]))
    f, grad, g, jnz = s.eval_jac_fg(x, p)
    torch.cuda.synchronize()
    for i in range(nc):
        assert abs(f[i].item() - GOLD[f"{variant}_{i}_f"]) <= 1e-12 * abs(GOLD[f"{variant}_{i}_f"])
        assert rel(g[i].cpu().numpy(), GOLD[f"{variant}_{i}_g"]) < 1e-12
        assert rel(grad[i].cpu().numpy(), GOLD[f"{variant}_{i}_grad"]) < 1e-12
        assert rel(jnz[i].cpu().numpy(), GOLD[f"{variant}_{i}_jnz"]) < 1e-12
        lam = dev(GOLD[f"{variant}_{i}_lam"][None])
        h = s.eval_hess_l(x[i:i + 1], p[i:i + 1], float(GOLD[f"{variant}_{i}_lamf"]), lam)
        assert rel(h[0].cpu().numpy(), GOLD[f"{variant}_{i}_hnz"]) < 1e-12
    s.close()


@pytest.mark.parametrize("N", [5, 15, 30])
def test_nlp_functions_match_oracle_other_horizons(oracle, N):
    P = pkg()
    s = P.BatchedCentroidalMPC(P.ergocub_config(horizon=N, contact_position_weight=200.0))
    cfg = make_cfg(N=N)
    d = oracle.dims(N)
    rng = np.random.default_rng(N)
    B = 4
    x, p, lam = rng.normal(size=(B, d["n"])), rng.normal(size=(B, d["np"])), 10 * rng.normal(size=(B, d["m"]))
    f, grad, g, jnz = s.eval_jac_fg(dev(x), dev(p))
    h = s.eval_hess_l(dev(x), dev(p), 0.9, dev(lam))
    torch.cuda.synchronize()
    for b in range(B):
        fo, grado, go, jo = oracle.jac_fg(cfg, x[b], p[b])
        ho = oracle.hess_l(cfg, x[b], p[b], 0.9, lam[b])
        assert abs(f[b].item() - fo) <= 1e-12 * abs(fo)
        assert rel(g[b].cpu().numpy(), go) < 1e-12 and rel(grad[b].cpu().numpy(), grado) < 1e-12
        assert rel(jnz[b].cpu().numpy(), jo) < 1e-12 and rel(h[b].cpu().numpy(), ho) < 1e-12
    s.close()


def test_solve_scenario_s0_known_answers(solver12, workloads):
    """SURVEY.md 8(d) scenario S0 (three pushes): optimum values of the survey's independent prototype"""
    kat = {(0.0, 0.0, 0.0): 3.208349928306e+01, (0.0, -0.3, 0.0): 1.5201340882e+02, (0.4, 0.0, 0.0): 1.4626640626e+02}
    for dcom0, fstar in kat.items():
        s = workloads.scenario_s0(dcom0)
        x, lam, obj, status, iters = solver12.solve_host(s["p"][None], s["lbg"][None], s["ubg"][None], s["x0"][None])
        assert status[0] == 0
        assert abs(obj[0] - fstar) <= 1e-6 * fstar, (obj[0], fstar)
    L = solver12.L
    assert abs(x[0][L.x_pos(1, 12)] - 0.11) < 1e-6      # pushed forward: landing at the edge of the step box


def test_solve_batch_matches_oracle_ergocub(solver12, oracle, workloads):
    w = workloads.walk_batch(N=12, B=96, seed=11, state_noise=2.0, yaw_range=0.3)
    x, lam, obj, status, iters = solver12.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    xo, lo, st = oracle.solve_batch(make_cfg(), w["p"], w["lbg"], w["ubg"], w["x0"], threads=os.cpu_count() or 4,
                                    opts=oracle_opts(oracle, solver12.strategy))
    assert (status == 0).all(), np.bincount(status)
    same = 0
    for b in range(96):
        assert st[b].status == 0
        compare_solutions(solver12.L, x[b], xo[b], obj[b], st[b].obj)
        same += 1
    it_o = np.array([s.iters for s in st])
    assert same == 96 and np.mean(np.abs(iters - it_o) <= 1) > 0.8   # same local optimum, (nearly) the same path


@pytest.mark.parametrize("strategy", STRATEGIES)
def test_solve_batch_matches_oracle_icub3_no_step_adjustment(oracle, workloads, strategy):
    P = pkg()
    s = P.BatchedCentroidalMPC(P.icub3_config(**strategy_kw(strategy)))
    w = workloads.walk_batch(N=15, B=48, seed=5, state_noise=1.0, step_adjust=False)
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    xo, lo, st = oracle.solve_batch(make_cfg(**ICUB_ORACLE), w["p"], w["lbg"], w["ubg"], w["x0"], threads=os.cpu_count() or 4,
                                    opts=oracle_opts(oracle, strategy))
    assert (status == 0).all(), np.bincount(status)
    # contact_force_symmetry_weight is 0 in the iCub3 ini: the split of a foot's wrench over its corners is not unique, and
    # rounding-level differences between the two implementations of the corrector solve (refinement sweep here, second
    # banded solve in the oracle) move along that flat set.  The monotone path tracks the oracle's iterates closely
    # enough for the corner forces to agree too; for the predictor-corrector what IS unique is asserted: the per-foot resultant
    # force and torque of every knot (what the dynamics and the reference's whole-body layer see) and the trajectories.
    for b in range(48):
        compare_solutions(s.L, x[b], xo[b], obj[b], st[b].obj, forces=(strategy == "monotone"), p=w["p"][b], corners=ICUB_CORNERS)
    s.close()


def reference_nlp():
    """the reference's own CasADi-generated functions (tmp.c compiled in place into oracle/_ref by oracle/Makefile; the
    prebuilt library travels to the GPU box with the snapshot).  A box without it fails loudly: the KKT certificate below is
    only worth something when it comes from the reference's code."""
    from oracle.oracle import RefNLP
    assert have_ref(), "oracle/_ref/libref_tmp.so is missing: run `make -C oracle ref` where /root/reference exists"
    return RefNLP("tmp")


def kkt_certificate(fn, jc, jr, x, lam, p, lbg, ubg, obj, n=555, m=651):
    """KKT conditions of (x, lam) evaluated with the functions `fn` (jac_fg): scaled stationarity, feasibility, objective,
    multiplier signs"""
    f, grad, g, jnz = fn(x, p)
    r = grad.copy()
    for c in range(n):
        sl = slice(jc[c], jc[c + 1])
        r[c] += np.dot(jnz[sl], lam[jr[sl]])
    scale = max(100.0, np.sum(np.abs(lam)) / m) / 100.0
    assert np.max(np.abs(r)) / scale < 1e-7, np.max(np.abs(r))
    viol = np.maximum(np.maximum(lbg - g, g - ubg), 0.0)
    assert viol.max() < 4e-8                               # bounds are relaxed by 1e-8 (bound_relax_factor)
    assert abs(f - obj) <= 1e-10 * abs(f)
    ineq = ubg > lbg
    assert np.all(np.abs(lam[ineq & (g < ubg - 1e-3) & (g > lbg + 1e-3)]) < 1e-4)   # mu / slack
    assert np.all(lam[ineq & (ubg > 1e19)] <= 1e-12) and np.all(lam[ineq & (lbg < -1e19)] >= -1e-12)


def test_kkt_conditions_through_reference_functions(solver12, oracle, workloads):
    """optimality of the GPU solution certified with the REFERENCE's own generated code (nlp_jac_fg of tmp.c, compiled into
    oracle/_ref/libref_tmp.so: N = 12, contact_position_weight 200) and, second, with the oracle's restatement of it"""
    ref = reference_nlp()
    w = workloads.walk_batch(N=12, B=16, seed=3, state_noise=1.5, yaw_range=0.2)
    x, lam, obj, status, iters = solver12.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    assert (status == 0).all()
    for b in range(16):
        kkt_certificate(ref.jac_fg, ref.jc, ref.jr, x[b], lam[b], w["p"][b], w["lbg"][b], w["ubg"][b], obj[b])
    cfg = make_cfg()
    jc, jr = oracle.jac_sparsity(12)
    for b in range(16):
        f, grad, g, jnz = oracle.jac_fg(cfg, x[b], w["p"][b])
        r = grad.copy()
        for c in range(555):
            sl = slice(jc[c], jc[c + 1])
            r[c] += np.dot(jnz[sl], lam[b][jr[sl]])
        scale = max(100.0, np.sum(np.abs(lam[b])) / 651) / 100.0
        assert np.max(np.abs(r)) / scale < 1e-7
        lb, ub = w["lbg"][b], w["ubg"][b]
        viol = np.maximum(np.maximum(lb - g, g - ub), 0.0)
        assert viol.max() < 2e-8 * 2                      # bounds are relaxed by 1e-8 (bound_relax_factor)
        assert abs(f - obj[b]) <= 1e-10 * abs(f)
        # sign of the multipliers: >= 0 only at upper bounds, <= 0 only at lower bounds (inequality rows)
        ineq = ub > lb
        assert np.all(np.abs(lam[b][ineq & (g < ub - 1e-3) & (g > lb + 1e-3)]) < 1e-4)   # mu / slack
        assert np.all(lam[b][ineq & (ub > 1e19)] <= 1e-12) and np.all(lam[b][ineq & (lb < -1e19)] >= -1e-12)


def test_full_size_batch_properties(workloads):
    """BASELINE config 2 at full size: every instance converges; the result does not depend on the position of an
    instance in the batch (determinism, bit exact); a converged point is a fixed point of a warm restart."""
    P = pkg()
    s = P.BatchedCentroidalMPC(P.icub3_config())
    w = workloads.walk_batch(N=15, B=1024, seed=0, state_noise=1.0, step_adjust=False)
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    assert (status == 0).all(), np.bincount(status)
    perm = np.random.default_rng(0).permutation(1024)
    x2, lam2, obj2, status2, iters2 = s.solve_host(w["p"][perm], w["lbg"][perm], w["ubg"][perm], w["x0"][perm])
    assert np.array_equal(x2, x[perm]) and np.array_equal(obj2, obj[perm]) and np.array_equal(iters2, iters[perm])
    x3, lam3, obj3, status3, iters3 = s.solve_host(w["p"][:64], w["lbg"][:64], w["ubg"][:64], x[:64])
    assert (status3 == 0).all()
    assert np.max(np.abs(obj3 - obj[:64]) / np.maximum(1.0, np.abs(obj[:64]))) < 1e-7
    s.close()


def test_bad_input_is_reported_not_fatal(solver12, workloads):
    w = workloads.walk_batch(N=12, B=4, seed=1)
    w["lbg"][1, 20] = 1.0     # a dynamics row with lbg != ubg
    w["x0"][2, 5] = np.nan
    x, lam, obj, status, iters = solver12.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    assert status[0] == 0 and status[3] == 0 and status[1] == 4 and status[2] == 4


def test_warmstart_shift(solver12):
    L = solver12.L
    N = 12
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, L.n)); lam = rng.normal(size=(3, L.m))
    dx, dl = dev(x), dev(lam)
    solver12.shift_warmstart(dx, dl)
    torch.cuda.synchronize()
    xs, ls = dx.cpu().numpy(), dl.cpu().numpy()
    for b in range(3):
        for blk, cols in ((0, N + 1), (L.x_dcom(0), N + 1), (L.x_h(0), N + 1), (L.x_pos(0, 0), N + 1), (L.x_vel(0, 0), N),
                          (L.x_frc(0, 2, 0), N), (L.x_pos(1, 0), N + 1), (L.x_frc(1, 3, 0), N)):
            a = x[b, blk:blk + 3 * cols].reshape(cols, 3)
            exp = np.vstack([a[1:], a[-1:]])
            assert np.array_equal(xs[b, blk:blk + 3 * cols].reshape(cols, 3), exp)
        assert np.array_equal(ls[b, :15], lam[b, :15])
        a = lam[b, L.g_h(0):L.g_h(0) + 3 * N].reshape(N, 3)
        assert np.array_equal(ls[b, L.g_h(0):L.g_h(0) + 3 * N].reshape(N, 3), np.vstack([a[1:], a[-1:]]))
        a = lam[b, L.g_fric(1, 0, 0):L.g_fric(1, 0, 0) + 16 * N].reshape(N, 16)
        assert np.array_equal(ls[b, L.g_fric(1, 0, 0):L.g_fric(1, 0, 0) + 16 * N].reshape(N, 16), np.vstack([a[1:], a[-1:]]))


def test_plant_rk4(solver12, workloads):
    """closed-loop plant against a numpy RK4 of the same centroidal dynamics"""
    L = solver12.L
    w = workloads.walk_batch(N=12, B=5, seed=2, yaw_range=0.3)
    rng = np.random.default_rng(1)
    x = w["x0"] + 0.1 * rng.normal(size=w["x0"].shape)
    state = np.hstack([w["p"][:, L.p_glob():L.p_glob() + 9]]) + 0.01 * rng.normal(size=(5, 9))
    ext = 0.5 * rng.normal(size=(5, 6))
    ds = dev(state)
    solver12.rollout_plant(dev(x), dev(w["p"]), ds, 0.002, 50, ext=dev(ext))
    torch.cuda.synchronize()
    out = ds.cpu().numpy()
    corners = np.array([(0.08, 0.01, 0), (0.08, -0.01, 0), (-0.08, -0.01, 0), (-0.08, 0.01, 0)])
    for b in range(5):
        F = np.array([0, 0, -9.80665]) + ext[b, :3]
        arms, forces = [], []
        for c in range(2):
            en = w["p"][b, L.p_en(c, 0)]
            R = w["p"][b, L.p_rot(c, 0):L.p_rot(c, 0) + 9].reshape(3, 3).T
            for j in range(4):
                arms.append(R @ corners[j] + x[b, L.x_pos(c, 0):L.x_pos(c, 0) + 3])
                forces.append(en * x[b, L.x_frc(c, j, 0):L.x_frc(c, j, 0) + 3])
        arms, forces = np.array(arms), np.array(forces)
        F = F + forces.sum(0)

        def fdot(s):
            com = s[:3]
            return np.hstack([s[3:6], F, ext[b, 3:] + np.cross(arms - com, forces).sum(0)])
        s = state[b].copy()
        for _ in range(50):
            k1 = fdot(s); k2 = fdot(s + 0.001 * k1); k3 = fdot(s + 0.001 * k2); k4 = fdot(s + 0.002 * k3)
            s = s + 0.002 / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        assert np.max(np.abs(out[b] - s)) < 1e-12


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("N", [2, 22, 50])
def test_solve_other_horizons_match_oracle(oracle, workloads, N, strategy):
    """reference horizons (ergoCubSN001: 22 knots) and the ends of BASELINE's horizon sweep: same optimum as the oracle"""
    P = pkg()
    s = P.BatchedCentroidalMPC(P.ergocub_config(horizon=N, contact_position_weight=200.0, **strategy_kw(strategy)))
    w = workloads.walk_batch(N=N, B=6, seed=N, state_noise=1.0, yaw_range=0.2)
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    xo, lo, st = oracle.solve_batch(make_cfg(N=N), w["p"], w["lbg"], w["ubg"], w["x0"], threads=3, opts=oracle_opts(oracle, strategy))
    L = pkg("layout").Layout(N)
    for b in range(6):
        assert status[b] == 0 and st[b].status == 0, (b, status[b], st[b].status)
        compare_solutions(L, x[b], xo[b], obj[b], st[b].obj)
    s.close()


def test_empty_batch_and_team_sizes(workloads, oracle):
    """batch of zero instances is a no-op; every team size gives the same solutions"""
    P = pkg()
    w = workloads.walk_batch(N=12, B=5, seed=3, state_noise=1.0)
    ref = None
    for team in (32, 64, 96, 128):
        s = P.BatchedCentroidalMPC(P.ergocub_config(contact_position_weight=200.0, threads_per_instance=team))
        e = torch.empty(0, 1, dtype=torch.float64, device="cuda")
        assert s.lib.cmpc_solve_batched(s.handle, 0, None, None, None, None, None, None, None, None, 0, None) in (0, -1)
        x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
        assert (status == 0).all()
        if ref is None:
            ref = (x, obj)
        else:
            assert np.max(np.abs(x - ref[0])) < 1e-6 and np.max(np.abs(obj - ref[1]) / np.abs(ref[1])) < 1e-9
        s.close()


def test_lockstep_groups_agree(workloads):
    """the teams of a CTA walk in 1, 2, 3 or 7 independent lock-step groups (named barriers): same solutions, every instance
    of a batch that spans several CTAs and leaves some teams without work"""
    P = pkg()
    w = workloads.walk_batch(N=12, B=45, seed=11, state_noise=1.5, yaw_range=0.2)
    ref = None
    for groups in (1, 2, 3, 7):
        s = P.BatchedCentroidalMPC(P.ergocub_config(contact_position_weight=200.0, lockstep_groups=groups, teams_per_cta=7))
        geo = s.geometry()
        assert geo["teams_per_cta"] == 7 and geo["lockstep_groups"] == groups, geo
        x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
        assert (status == 0).all(), status
        if ref is None:
            ref = (x, obj, iters)
        else:
            assert np.max(np.abs(x - ref[0])) < 1e-9 and np.array_equal(iters, ref[2])
        s.close()


@pytest.mark.parametrize("robot", ["ergocub", "icub3"])
def test_predictor_corrector_lands_on_the_monotone_optimum(oracle, workloads, robot):
    """The default barrier update (Mehrotra predictor-corrector) against the oracle's restatement of IPOPT's DEFAULT path
    (monotone update): objective within 1e-6, CoM / momentum trajectories and footsteps within 1e-5, and with fewer
    iterations.  Contact forces: within 1e-5 where the optimum is unique (ergoCub, symmetry weight > 0); the iCub3 ini has
    contact_force_symmetry_weight 0, which leaves the split of a foot's wrench over its four corners undetermined (constant
    offsets in the null space of the wrench map cost nothing): there the fraction of instances with equal corner forces is
    reported and the resultant quantities (the trajectories) are checked."""
    P = pkg()
    if robot == "ergocub":
        s = P.BatchedCentroidalMPC(P.ergocub_config(contact_position_weight=200.0))
        w = workloads.walk_batch(N=12, B=64, seed=21, state_noise=2.0, yaw_range=0.3)
        ocfg = make_cfg()
    else:
        s = P.BatchedCentroidalMPC(P.icub3_config())
        w = workloads.walk_batch(N=15, B=64, seed=22, state_noise=1.0, step_adjust=False)
        ocfg = make_cfg(**ICUB_ORACLE)
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    xo, lo, st = oracle.solve_batch(ocfg, w["p"], w["lbg"], w["ubg"], w["x0"], threads=os.cpu_count() or 4)  # monotone
    assert (status == 0).all(), np.bincount(status)
    g = groups(s.L)
    same_forces = 0
    for b in range(64):
        assert st[b].status == 0
        assert abs(obj[b] - st[b].obj) <= 1e-6 * max(1.0, abs(st[b].obj)), (b, obj[b], st[b].obj)
        for name in ("com", "dcom", "h", "footsteps"):
            idx = g[name]
            assert np.max(np.abs(x[b][idx] - xo[b][idx])) / max(1.0, np.max(np.abs(xo[b][idx]))) <= 1e-5, (b, name)
        idx = g["forces"]
        same_forces += np.max(np.abs(x[b][idx] - xo[b][idx])) / max(1.0, np.max(np.abs(xo[b][idx]))) <= 1e-5
        # what IS unique, and what the reference consumes: the per-foot resultant force / torque (every knot, incl. knot 0)
        corners = ERGOCUB_CORNERS if robot == "ergocub" else ICUB_CORNERS
        wa, wb = foot_wrenches(s.L, x[b], w["p"][b], corners), foot_wrenches(s.L, xo[b], w["p"][b], corners)
        assert np.max(np.abs(wa - wb)) / max(1.0, np.max(np.abs(wb))) <= 1e-5, (b, "per-foot wrench")
    it_o = np.mean([q.iters for q in st])
    print(f"{robot}: corner forces equal to the monotone optimum in {same_forces}/64 instances; iterations {iters.mean():.1f} vs {it_o:.1f}")
    if robot == "ergocub":
        assert same_forces == 64
    assert iters.mean() < 0.8 * it_o
    s.close()


def test_full_size_ergocub_shard_properties(oracle, workloads):
    """BASELINE config 3, the shard of one GPU (65536 / 8 instances, step adjustment and friction cones active, state noise
    x 2, yaw +-0.3): every instance converges in a bounded number of iterations on both barrier updates, the two land on
    the same optimum (objective 1e-6, all variable groups 1e-5: unique optimum with the symmetry weight), and a sample
    satisfies the KKT conditions evaluated with the ORACLE's functions."""
    P = pkg()
    B = 8192
    w = workloads.walk_batch(N=12, B=B, seed=3, state_noise=2.0, yaw_range=0.3, step_adjust=True)
    res = {}
    for strat in STRATEGIES:
        s = P.BatchedCentroidalMPC(P.ergocub_config(**strategy_kw(strat)))
        res[strat] = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
        L = s.L
        s.close()
        x, lam, obj, status, iters = res[strat]
        assert (status == 0).all(), (strat, np.bincount(status), np.where(status != 0)[0][:10])
        assert iters.max() <= (25 if strat == "mehrotra" else 40), (strat, iters.max())
    xm, lm, om_, _, itm = res["mehrotra"]
    xq, lq, oq, _, itq = res["monotone"]
    assert np.max(np.abs(om_ - oq) / np.maximum(1.0, np.abs(oq))) <= 1e-6
    for name, idx in groups(L).items():
        err = np.max(np.abs(xm[:, idx] - xq[:, idx]), axis=1) / np.maximum(1.0, np.max(np.abs(xq[:, idx]), axis=1))
        assert err.max() <= 1e-5, (name, err.max(), int(err.argmax()))
    assert itm.mean() < 0.75 * itq.mean()
    # the same shard with the weight set baked into the reference's tmp.c (contact_position_weight 200): every 16th instance
    # certified with the reference's own generated functions
    ref = reference_nlp()
    s = P.BatchedCentroidalMPC(P.ergocub_config(contact_position_weight=200.0))
    sub = slice(0, B, 16)
    xr, lr, or_, str_, itr = s.solve_host(w["p"][sub], w["lbg"][sub], w["ubg"][sub], w["x0"][sub])
    s.close()
    assert (str_ == 0).all(), np.bincount(str_)
    for i, b in enumerate(range(0, B, 16)):
        kkt_certificate(ref.jac_fg, ref.jc, ref.jr, xr[i], lr[i], w["p"][b], w["lbg"][b], w["ubg"][b], or_[i])
    cfg = make_cfg(w_pos=2000.0)
    jc, jr = oracle.jac_sparsity(12)
    for b in range(0, B, 512):
        f, grad, g, jnz = oracle.jac_fg(cfg, xm[b], w["p"][b])
        r = grad.copy()
        for c in range(555):
            sl = slice(jc[c], jc[c + 1])
            r[c] += np.dot(jnz[sl], lm[b][jr[sl]])
        scale = max(100.0, np.sum(np.abs(lm[b])) / 651) / 100.0
        assert np.max(np.abs(r)) / scale < 1e-7, (b, np.max(np.abs(r)))
        assert np.maximum(np.maximum(w["lbg"][b] - g, g - w["ubg"][b]), 0.0).max() < 4e-8
        assert abs(f - om_[b]) <= 1e-10 * abs(f)


def test_predictor_corrector_hands_hard_instances_to_the_monotone_path(oracle, workloads):
    """4 x the nominal state noise: one instance of this batch defeats the predictor-corrector (line-search failure after six
    iterations); it is solved again from its initial point on the monotone path inside the same launch -- same result and
    (nearly) the same iteration total as the oracle, which does the same."""
    P = pkg()
    s = P.BatchedCentroidalMPC(P.ergocub_config())
    w = workloads.walk_batch(N=12, B=96, seed=4, state_noise=4.0, yaw_range=0.3, step_adjust=True)
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    xo, lo, st = oracle.solve_batch(make_cfg(w_pos=2000.0), w["p"], w["lbg"], w["ubg"], w["x0"], threads=os.cpu_count() or 4,
                                    opts=oracle.default_opts(mehrotra=1))
    s.close()
    assert (status == 0).all(), np.bincount(status)
    fell_back = [b for b in range(96) if st[b].n_fallback]
    assert fell_back, "the workload no longer contains an instance that needs the monotone path"
    for b in range(96):
        assert st[b].status == 0
        assert abs(obj[b] - st[b].obj) <= 1e-6 * max(1.0, abs(st[b].obj)), (b, obj[b], st[b].obj)
    for b in fell_back:
        assert iters[b] > 25 and abs(int(iters[b]) - st[b].iters) <= 5, (b, iters[b], st[b].iters)


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("tol", [1e-2, 1e-4])
def test_ini_tolerances(oracle, workloads, tol, strategy):
    """ipopt_tolerance of the reference's ergoCub ini files: 1e-4 (ergoCubGazeboV1 / V1_1) and 1e-2 (ergoCubSN000 / SN001,
    where IPOPT's constr_viol_tol / compl_inf_tol 1e-4 decide and the barrier parameter goes down to min(tol, 1e-4) / 11):
    every instance converges, along the same path as the oracle (same iteration count within one, objective within 1e-6)
    and to within the tolerance of the tight solution."""
    P = pkg()
    w = workloads.walk_batch(N=12, B=64, seed=9, state_noise=2.0, yaw_range=0.3)
    s = P.BatchedCentroidalMPC(P.ergocub_config(ipopt_tolerance=tol, **strategy_kw(strategy)))
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    s.close()
    xo, lo, st = oracle.solve_batch(make_cfg(w_pos=2000.0), w["p"], w["lbg"], w["ubg"], w["x0"], threads=os.cpu_count() or 4,
                                    opts=oracle_opts(oracle, strategy, tol=tol))
    xt, lt, stt = oracle.solve_batch(make_cfg(w_pos=2000.0), w["p"], w["lbg"], w["ubg"], w["x0"], threads=os.cpu_count() or 4)
    assert (status == 0).all(), np.bincount(status)
    for b in range(64):
        assert st[b].status == 0
        assert abs(int(iters[b]) - st[b].iters) <= 1, (b, iters[b], st[b].iters)
        assert abs(obj[b] - st[b].obj) <= 1e-6 * max(1.0, abs(st[b].obj)), (b, obj[b], st[b].obj)
        assert abs(obj[b] - stt[b].obj) <= 1e-3 * max(1.0, abs(stt[b].obj)), (b, obj[b], stt[b].obj)


def test_small_batch_kernels_agree_with_the_lockstep_kernel(workloads):
    """default geometry: batches of one / up to 2 / up to 4 instances per SM run on independent single-team CTAs (a team of
    256 threads / of 128 without a register cap / the one compiled for four CTAs per SM), larger ones on the seven-team
    lock-step CTAs.  Same solutions from all four (different team sizes: rounding-level differences only)."""
    P = pkg()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    w = workloads.walk_batch(N=12, B=4 * sms, seed=17, state_noise=1.5, yaw_range=0.2)
    ref = P.BatchedCentroidalMPC(P.ergocub_config(teams_per_cta=7))
    auto = P.BatchedCentroidalMPC(P.ergocub_config())
    xr, lr, objr, str_, itr = ref.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    assert (str_ == 0).all()
    for B in (1, sms, sms + 3, 2 * sms, 2 * sms + 1, 3 * sms + 5, 4 * sms):
        n0 = auto.launch_count()
        x, lam, obj, status, iters = auto.solve_host(w["p"][:B], w["lbg"][:B], w["ubg"][:B], w["x0"][:B])
        assert auto.launch_count() == n0 + 1
        assert (status == 0).all(), (B, np.bincount(status))
        assert np.max(np.abs(x - xr[:B])) < 1e-6 and np.max(np.abs(obj - objr[:B]) / np.abs(objr[:B])) < 1e-9, B
        assert np.max(np.abs(iters - itr[:B])) <= 1
    ref.close()
    auto.close()


@pytest.mark.parametrize("strategy", STRATEGIES)
def test_iteration_limit_is_final(workloads, strategy):
    """ipopt_max_iteration (the reference's ini files carry a commented `ipopt_max_iteration 14`): an instance that runs into
    the limit reports CMPC_STATUS_MAX_ITER with exactly that many iterations -- the predictor-corrector does not start a
    monotone re-solve on top of an exhausted budget -- and its x is the last iterate."""
    P = pkg()
    w = workloads.walk_batch(N=12, B=8, seed=2, state_noise=1.0)
    s = P.BatchedCentroidalMPC(P.ergocub_config(ipopt_max_iteration=5, **strategy_kw(strategy)))
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    s.close()
    assert (status == 1).all() and (iters == 5).all(), (status, iters)
    assert np.isfinite(x).all() and np.any(x != w["x0"])


def test_long_horizon_auxiliary_kernels(oracle):
    """horizons above 115 / 152 knots need more than 48 KB of dynamic shared memory in the shift / evaluation kernels (opted in
    by cmpc_create): N = 160 evaluates f, g, grad f like the oracle and shifts like numpy"""
    P = pkg()
    N = 160
    s = P.BatchedCentroidalMPC(P.ergocub_config(horizon=N, contact_position_weight=200.0))
    cfg = make_cfg(N=N)
    d = oracle.dims(N)
    rng = np.random.default_rng(N)
    x, p = rng.normal(size=(2, d["n"])), rng.normal(size=(2, d["np"]))
    f, grad, g, _ = s.eval_jac_fg(dev(x), dev(p), want_jac=False)
    torch.cuda.synchronize()
    for b in range(2):
        fo, grado, go, jo = oracle.jac_fg(cfg, x[b], p[b])
        assert abs(f[b].item() - fo) <= 1e-12 * abs(fo)
        assert rel(g[b].cpu().numpy(), go) < 1e-12 and rel(grad[b].cpu().numpy(), grado) < 1e-12
    L = s.L
    lam = rng.normal(size=(2, L.m))
    dx, dl = dev(x), dev(lam)
    s.shift_warmstart(dx, dl)
    torch.cuda.synchronize()
    xs, ls = dx.cpu().numpy(), dl.cpu().numpy()
    for b in range(2):
        for blk, cols in ((0, N + 1), (L.x_h(0), N + 1), (L.x_pos(1, 0), N + 1), (L.x_vel(0, 0), N), (L.x_frc(1, 3, 0), N)):
            a = x[b, blk:blk + 3 * cols].reshape(cols, 3)
            assert np.array_equal(xs[b, blk:blk + 3 * cols].reshape(cols, 3), np.vstack([a[1:], a[-1:]]))
        a = lam[b, L.g_fric(0, 0, 0):L.g_fric(0, 0, 0) + 16 * N].reshape(N, 16)
        assert np.array_equal(ls[b, L.g_fric(0, 0, 0):L.g_fric(0, 0, 0) + 16 * N].reshape(N, 16), np.vstack([a[1:], a[-1:]]))
    s.close()


def test_two_streams_on_one_handle_are_ordered(workloads):
    """a handle owns one work queue and one scratch arena: two solves enqueued on different streams without any
    synchronisation by the caller are serialised by the library and give the results of two separate calls"""
    P = pkg()
    s = P.BatchedCentroidalMPC(P.ergocub_config(teams_per_cta=7))
    w1 = workloads.walk_batch(N=12, B=300, seed=41, state_noise=1.5, yaw_range=0.2)
    w2 = workloads.walk_batch(N=12, B=300, seed=42, state_noise=1.5, yaw_range=0.2)
    ref = [s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"]) for w in (w1, w2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    dx, out = [], []
    torch.cuda.synchronize()
    for w, st in zip((w1, w2), streams):
        with torch.cuda.stream(st):
            x = dev(w["x0"])
            out.append(s.solve(dev(w["p"]), dev(w["lbg"]), dev(w["ubg"]), x))
            dx.append(x)
    torch.cuda.synchronize()
    for i in range(2):
        assert (out[i][1].cpu().numpy() == 0).all()
        assert np.array_equal(dx[i].cpu().numpy(), ref[i][0]) and np.array_equal(out[i][2].cpu().numpy(), ref[i][4])
    s.close()


def test_two_handles_two_host_threads_overlap_and_agree(workloads):
    """the host-pointer entry points run on a private stream of their handle: two handles driven from two host threads (the
    pipelined mode of bench.py) give bit for bit the results of the same calls made one after the other"""
    import threading
    P = pkg()
    hs = [P.BatchedCentroidalMPC(P.icub3_config()) for _ in range(2)]
    ws = [workloads.walk_batch(N=15, B=1100, seed=60 + j, state_noise=1.0, step_adjust=False) for j in range(4)]
    ref = [hs[0].solve_host(w["p"], w["lbg"], w["ubg"], w["x0"]) for w in ws]
    got = [None] * 4

    def work(slot):
        for j in range(slot, 4, 2):
            w = ws[j]
            got[j] = hs[slot].solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for j in range(4):
        assert (got[j][3] == 0).all()
        assert np.array_equal(got[j][0], ref[j][0]) and np.array_equal(got[j][4], ref[j][4])
    for h in hs:
        h.close()
