"""GPU tests of the callers on either side of the solve that moved to the device:
input population from compact tick records (cmpc_populate, cmpc_solve_ticks_host), reference resampling
(cmpc_resample_references), desired ZMP (cmpc_desired_zmp).  Checked against the host restatements (C++ operator) and numpy."""
import os

import numpy as np
import pytest

from conftest import ROOT, pkg

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
DATA = os.path.join(ROOT, "tests", "data")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.parametrize("robot,N,kw", [("ergocub", 12, dict(yaw_range=0.3, step_adjust=True)),
                                        ("icub3", 15, dict(step_adjust=False)),
                                        ("ergocub", 22, dict(yaw_range=0.2, step_adjust=True))])
def test_populate_matches_the_synthetic_workloads(workloads, robot, N, kw):
    """cmpc_populate(tick records) == the (p, lbg, ubg, x0) that workloads.walk_batch writes directly, for every phase of the
    walk (random phases, footstep yaw, pushes)"""
    P = pkg()
    B = 256
    push = np.random.default_rng(1).normal(size=(B, 3))
    w = workloads.walk_batch(N=N, B=B, seed=7, state_noise=1.0, push=push, ticks=True, **kw)
    cfg = (P.ergocub_config if robot == "ergocub" else P.icub3_config)(horizon=N)
    s = P.BatchedCentroidalMPC(cfg)
    p, lbg, ubg, x0 = s.populate(dev(w["ticks"]))
    torch.cuda.synchronize()
    for name, a in (("p", p), ("lbg", lbg), ("ubg", ubg), ("x0", x0)):
        a, b = a.cpu().numpy(), w[name]
        bad = np.nonzero(~np.isclose(a, b, rtol=0, atol=1e-12))
        assert bad[0].size == 0, (name, bad[0][:5], bad[1][:5], a[bad][:5], b[bad][:5])
    assert len(set(w["phase"].tolist())) >= 12          # the batch covers the walk cycle
    s.close()


@pytest.mark.parametrize("phase", [0, 2, 3, 4, 7, 8, 10, 11, 13, 15])
def test_populate_is_bit_identical_to_the_host_operator(workloads, phase):
    """the record the C++ operator uploads, expanded on the device, against the C++ operator's own host restatement of the
    population rules (Impl::fillInputs): bit for bit, including rotated contacts, a mid-swing foot and an external wrench"""
    P, H = pkg(), pkg("host")
    N = 12
    w = workloads.walk_batch(N=N, B=1, seed=3, phase=phase, state_noise=1.0)
    L = pkg("layout").Layout(N)
    p = w["p"][0]
    m = H.CentroidalMPCHost(os.path.join(DATA, "ergocub", "centroidal_mpc.ini"), "")
    g0 = L.p_glob()
    assert m.set_state(p[g0:g0 + 3], p[g0 + 3:g0 + 6], p[g0 + 6:g0 + 9], wrench=[1.0, -2.0, 0.5, 0.1, 0.2, 0.3])
    assert m.set_reference_trajectory(p[L.p_comref(0):L.p_comref(0) + 3 * (N + 1)], 0.01 * np.arange(3 * (N + 1)).reshape(-1, 3))
    lists = H.walk_contact_lists(phase)
    lists = {k: [(c[0] + 0.013, c[1] + 0.013, c[2], 0.1 * (i % 3) - 0.1) for i, c in enumerate(v)] for k, v in lists.items()}   # yaw, off-grid times
    assert m.set_contact_phase_list(lists)
    hp, hl, hu, hx = m.solver_inputs()
    s = P.BatchedCentroidalMPC(m.config())
    dp, dl, du, dx = s.populate(dev(m.tick_record()[None]))
    torch.cuda.synchronize()
    for name, a, b in (("p", dp, hp), ("lbg", dl, hl), ("ubg", du, hu), ("x0", dx, hx)):
        a = a.cpu().numpy()[0]
        assert np.array_equal(a, b), (name, np.nonzero(a != b)[0][:8])
    s.close()


def test_solve_ticks_host_modes(workloads):
    """cmpc_solve_ticks_host: cold start == cmpc_solve_host on the expanded input; warm start from the resident solution ==
    warm start from an uploaded previous solution == cmpc_solve_host with the host-shifted solution"""
    P = pkg()
    ph = np.random.default_rng(9).integers(0, 16, size=64)
    w = workloads.walk_batch(N=12, B=64, seed=9, state_noise=1.0, ticks=True, phase=ph)
    s = P.BatchedCentroidalMPC(P.ergocub_config())
    xr, lr, objr, str_, itr = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    x0_, l0, obj0, st0, it0 = s.solve_ticks_host(w["ticks"], warm_mode=0)
    assert (st0 == 0).all() and np.array_equal(it0, itr)
    assert np.max(np.abs(x0_ - xr)) < 1e-9 and np.max(np.abs(obj0 - objr) / np.abs(objr)) < 1e-12
    # next tick: same instances one knot later
    w1 = workloads.walk_batch(N=12, B=64, seed=9, state_noise=1.0, ticks=True, phase=ph + 1)
    xa, la, obja, sta, ita = s.solve_ticks_host(w1["ticks"], warm_mode=1)                       # resident solution of the last call
    xb, lb, objb, stb, itb = s.solve_ticks_host(w1["ticks"], warm_mode=2, x_prev=x0_, lam_prev=l0)   # uploaded, shifted on the device
    assert (sta == 0).all() and np.array_equal(ita, itb) and np.array_equal(xa, xb)
    dx, dl = dev(x0_), dev(l0)
    s.shift_warmstart(dx, dl)
    torch.cuda.synchronize()
    xc, lc, objc, stc, itc = s.solve_host(w1["p"], w1["lbg"], w1["ubg"], dx.cpu().numpy(), lam_g0=dl.cpu().numpy())
    assert np.array_equal(itc, ita) and np.max(np.abs(xc - xa)) < 1e-9
    assert ita.mean() <= it0.mean()                                                             # the warm start pays
    with pytest.raises(RuntimeError):
        s.solve_ticks_host(w1["ticks"][:10], warm_mode=1)                                       # nothing of that size is resident
    s.close()


def test_resample_references_matches_numpy(workloads):
    """the LinearSpline frequency adapters of CentroidalMPCBlock.cpp:201-260, 525-577 on the device: 50 Hz planner samples
    (slow-down factor 1.5) -> N + 1 MPC knots, angular momentum divided by the robot mass, CoM height overridden"""
    P = pkg()
    N, B, n_in = 12, 33, 60
    rng = np.random.default_rng(0)
    t_in = np.arange(n_in) * 0.02 * 1.5
    com, ang = rng.normal(size=(B, n_in, 3)), rng.normal(size=(B, n_in, 3))
    t_out = 0.05 + np.arange(N + 1) * 0.1          # starts inside, ends beyond the last planner sample (1.77 s): end point held
    s = P.BatchedCentroidalMPC(P.ergocub_config())
    ts = workloads.tick_stride(N)
    ticks = torch.zeros(B, ts, dtype=torch.float64, device="cuda")
    s.resample_references(ticks, dev(t_in), dev(com), dev(ang), dev(t_out), robot_mass=56.0, com_height=0.7)
    torch.cuda.synchronize()
    t = ticks.cpu().numpy()
    ref_com = np.stack([np.stack([np.interp(t_out, t_in, com[b, :, a]) for a in range(3)], axis=1) for b in range(B)])
    ref_ang = np.stack([np.stack([np.interp(t_out, t_in, ang[b, :, a]) for a in range(3)], axis=1) for b in range(B)]) / 56.0
    ref_com[:, :, 2] = 0.7
    assert np.allclose(t[:, 17:17 + 3 * (N + 1)].reshape(B, N + 1, 3), ref_com, rtol=0, atol=1e-14)
    assert np.allclose(t[:, 17 + 3 * (N + 1):17 + 6 * (N + 1)].reshape(B, N + 1, 3), ref_ang, rtol=0, atol=1e-14)
    assert np.all(t[:, :17] == 0) and np.all(t[:, 17 + 6 * (N + 1):] == 0)      # nothing else is touched
    s.resample_references(ticks, dev(t_in), dev(com), dev(ang), dev(t_out), robot_mass=1.0, com_height=-1.0)   # height kept
    torch.cuda.synchronize()
    assert np.allclose(ticks.cpu().numpy()[:, 17 + 2:17 + 3 * (N + 1):3], np.stack([np.interp(t_out, t_in, com[b, :, 2]) for b in range(B)]), atol=1e-14)
    # same numbers as the host adapter of the C++ operator
    H = pkg("host")
    assert np.allclose(H.resample_linear(t_in, com[0], t_out)[:, :2], ref_com[0][:, :2], atol=1e-14)
    s.close()


def zmp_numpy(L, x, p, corners, hl, hw):
    """computeDesiredZMP (WholeBodyQPBlock.cpp:805-873) restated with numpy"""
    num, den = np.zeros(2), 0.0
    for c in range(2):
        R = p[L.p_rot(c, 0):L.p_rot(c, 0) + 9].reshape(3, 3).T
        en = p[L.p_en(c, 0)]
        pos = x[L.x_pos(c, 0):L.x_pos(c, 0) + 3]
        F, T = np.zeros(3), np.zeros(3)
        for j in range(4):
            f = en * x[L.x_frc(c, j, 0):L.x_frc(c, j, 0) + 3]
            F += f
            T += np.cross(corners[j], R.T @ f)
        if F[2] <= 0.001:
            continue
        local = np.array([np.clip(-T[1] / F[2], -hl, hl), np.clip(T[0] / F[2], -hw, hw), 0.0])
        world = R @ local + pos
        num += F[2] * world[:2]
        den += F[2]
    return num / den if den >= 0.001 else None


def test_desired_zmp_matches_numpy_and_the_host_operator(workloads):
    P, H = pkg(), pkg("host")
    N = 12
    L = pkg("layout").Layout(N)
    w = workloads.walk_batch(N=N, B=200, seed=5, state_noise=2.0, yaw_range=0.3)
    s = P.BatchedCentroidalMPC(P.ergocub_config())
    x, lam, obj, status, iters = s.solve_host(w["p"], w["lbg"], w["ubg"], w["x0"])
    corners = np.array([(0.08, 0.01, 0), (0.08, -0.01, 0), (-0.08, -0.01, 0), (-0.08, 0.01, 0)])
    for hl, hw in ((0.08, 0.03), (0.05, 0.005)):           # the reference's clamp, and one that bites
        zmp, valid = s.desired_zmp(dev(x), dev(w["p"]), hl, hw)
        torch.cuda.synchronize()
        zmp, valid = zmp.cpu().numpy(), valid.cpu().numpy()
        for b in range(200):
            ref = zmp_numpy(L, x[b], w["p"][b], corners, hl, hw)
            assert (ref is not None) == bool(valid[b])
            assert np.allclose(zmp[b], ref, rtol=0, atol=1e-12), (b, zmp[b], ref)
    # the ZMP of a walking robot lies between the feet
    assert np.all(np.abs(zmp[:, 1]) < 0.12)
    # the C++ operator's computeDesiredZMP on the output of advance()
    m = H.CentroidalMPCHost(os.path.join(DATA, "ergocub", "centroidal_mpc.ini"), "")
    w1 = workloads.walk_batch(N=N, B=1, seed=11, phase=2, state_noise=1.0)
    p = w1["p"][0]
    g0 = L.p_glob()
    m.set_state(p[g0:g0 + 3], p[g0 + 3:g0 + 6], p[g0 + 6:g0 + 9])
    m.set_reference_trajectory(p[L.p_comref(0):L.p_comref(0) + 39], p[L.p_href(0):L.p_href(0) + 39])
    m.set_contact_phase_list(H.walk_contact_lists(2))
    assert m.advance(), m.last_error()
    zh = m.desired_zmp()
    x1, _, _, st1, _ = s.solve_host(w1["p"], w1["lbg"], w1["ubg"], w1["x0"])      # tol 1e-8 here, 1e-4 in the ini
    zd, _ = s.desired_zmp(dev(x1), dev(w1["p"]))
    assert np.allclose(zd.cpu().numpy()[0], zh, atol=2e-3)
    s.close()
