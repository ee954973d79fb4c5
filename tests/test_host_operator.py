"""Host-side operator (C++ CentroidalMPC drop-in, libcmpc_host.so): configuration, input population, error behaviour.
CPU only; the solve itself is covered by tests/test_gpu_host_operator.py (marker gpu)."""
import os

import numpy as np
import pytest

from conftest import ROOT, pkg

DATA = os.path.join(ROOT, "tests", "data")


@pytest.fixture(scope="module")
def H():
    import importlib
    b = importlib.import_module(pkg().__name__ + ".build")
    b.build()
    return pkg("host")


def test_ini_dialect_2023_through_includes(H):
    m = H.CentroidalMPCHost(os.path.join(DATA, "ergocub", "centroidal_mpc_walking.ini"), "TRAJECTORY_ADJUSTMENT/CENTROIDAL_MPC")
    c = m.config()
    assert (m.N, c.horizon) == (12, 12) and abs(m.dT - 0.1) < 1e-15
    assert list(c.com_weight) == [10.0, 10.0, 200.0] and c.contact_position_weight == 2e3
    assert c.contact_force_symmetry_weight == 10.0 and c.angular_momentum_weight == 100.0
    assert c.ipopt_tolerance == 1e-4 and c.static_friction_coefficient == 0.33
    corners = np.array(c.corners[:]).reshape(2, 4, 3)
    # the reference files write `corner_3 (-0.08 0.01, 0.0)` without the first comma: still three numbers
    assert np.allclose(corners[0], [[0.08, 0.01, 0], [0.08, -0.01, 0], [-0.08, -0.01, 0], [-0.08, 0.01, 0]])
    assert np.allclose(corners[0], corners[1])


def test_ini_dialect_2022(H):
    m = H.CentroidalMPCHost(os.path.join(DATA, "icub3", "centroidal_mpc_walking.ini"), "CENTROIDAL_MPC")
    c = m.config()
    assert (m.N, c.horizon) == (15, 15) and abs(m.dT - 0.1) < 1e-15
    assert list(c.com_weight) == [1.0, 1.0, 200.0] and c.contact_position_weight == 2e2
    assert c.contact_force_symmetry_weight == 0.0          # the 2022 NLP has no symmetry term
    assert c.ipopt_tolerance == 1e-8                        # IPOPT default
    assert np.allclose(np.array(c.corners[:]).reshape(2, 4, 3)[1, 3], [-0.08, 0.03, 0.0])


def test_missing_group_and_bad_values_fail(H, tmp_path):
    with pytest.raises(RuntimeError):
        H.CentroidalMPCHost(os.path.join(DATA, "ergocub", "centroidal_mpc_walking.ini"), "NOT_A_GROUP")
    bad = tmp_path / "bad.ini"
    bad.write_text(open(os.path.join(DATA, "icub3", "centroidal_mpc.ini")).read().replace("number_of_slices                1", "number_of_slices 2"))
    with pytest.raises(RuntimeError):
        H.CentroidalMPCHost(str(bad))
    sqp = tmp_path / "sqp.ini"
    sqp.write_text(open(os.path.join(DATA, "ergocub", "centroidal_mpc.ini")).read().replace('"ipopt"', '"sqpmethod"'))
    with pytest.raises(RuntimeError):
        H.CentroidalMPCHost(str(sqp))


def test_calls_before_inputs_fail(H):
    m = H.CentroidalMPCHost(os.path.join(DATA, "ergocub", "centroidal_mpc.ini"))
    with pytest.raises(RuntimeError):
        m.solver_inputs()
    assert not m.is_output_valid()
    assert not m.set_reference_trajectory(np.zeros((5, 3)), np.zeros((5, 3)))   # needs N + 1 = 13 samples
    assert "13" in m.last_error()
    assert not m.set_contact_phase_list({"left_foot": [(0.0, 1.0, (0, 0.08, 0), 0.0)]})   # right_foot missing


@pytest.mark.parametrize("robot,group,N,step_adjust", [("ergocub", "", 12, True), ("icub3", "", 15, True)])
@pytest.mark.parametrize("phase", [0, 2, 3, 4, 7, 8, 10, 11, 13, 15])
def test_input_population_matches_the_synthetic_workloads(H, workloads, robot, group, N, step_adjust, phase):
    """the C++ operator fed with the walk schedule as contact lists must produce the (p, lbg, ubg, x0) that
    workloads.walk_batch writes directly (same population rules)"""
    w = workloads.walk_batch(N=N, B=1, seed=3, phase=phase, state_noise=1.0, step_adjust=step_adjust)
    L = pkg("layout").Layout(N)
    p = w["p"][0]
    m = H.CentroidalMPCHost(os.path.join(DATA, robot, "centroidal_mpc.ini"), group)
    g0 = L.p_glob()
    assert m.set_state(p[g0:g0 + 3], p[g0 + 3:g0 + 6], p[g0 + 6:g0 + 9])
    assert m.set_reference_trajectory(p[L.p_comref(0):L.p_comref(0) + 3 * (N + 1)], p[L.p_href(0):L.p_href(0) + 3 * (N + 1)])
    assert m.set_contact_phase_list(H.walk_contact_lists(phase))
    pp, lb, ub, x0 = m.solver_inputs()
    for name, a, b in (("p", pp, p), ("lbg", lb, w["lbg"][0]), ("ubg", ub, w["ubg"][0]), ("x0", x0, w["x0"][0])):
        bad = np.nonzero(~np.isclose(a, b, rtol=0, atol=1e-12))[0]
        assert bad.size == 0, (name, bad[:10], a[bad[:10]], b[bad[:10]])


def test_rotation_and_external_wrench_population(H):
    m = H.CentroidalMPCHost(os.path.join(DATA, "ergocub", "centroidal_mpc.ini"))
    L = m.L
    assert m.set_state([0, 0, 0.7], [0, 0, 0], [0, 0, 0], wrench=[1.0, -2.0, 0.5, 0.1, 0.2, 0.3])
    assert m.set_reference_trajectory(np.tile([0, 0, 0.7], (13, 1)), np.zeros((13, 3)))
    yaw = 0.3
    assert m.set_contact_phase_list({"left_foot": [(-1.0, 10.0, (0.0, 0.08, 0.0), yaw)],
                                     "right_foot": [(-1.0, 0.35, (0.0, -0.08, 0.0), 0.0), (0.85, 10.0, (0.1, -0.08, 0.0), -0.2)]},
                                    force_sample_time=0.1)
    p, lb, ub, x0 = m.solver_inputs()
    R = p[L.p_rot(0, 5):L.p_rot(0, 5) + 9].reshape(3, 3).T     # column major
    assert np.allclose(R, [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    # external wrench in column 0 only
    assert np.allclose(p[L.p_extf(0):L.p_extf(0) + 3], [1.0, -2.0, 0.5]) and np.allclose(p[L.p_extt(0):L.p_extt(0) + 3], [0.1, 0.2, 0.3])
    assert np.all(p[L.p_extf(1):L.p_extf(1) + 3 * 11] == 0)
    # forceSampleTime moved 0.35 -> 0.3 and 0.85 -> 0.8: right foot enabled at knots 0..2 and 8..11
    en = p[L.p_en(1, 0):L.p_en(1, 0) + 12]
    assert en.tolist() == [1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1]
    # step box: zero width on the current contact, free in swing, configured box on the future contact
    assert np.all(lb[L.g_box(1, 0):L.g_box(1, 0) + 9] == 0) and np.all(ub[L.g_box(1, 0):L.g_box(1, 0) + 9] == 0)
    assert np.all(lb[L.g_box(1, 3):L.g_box(1, 3) + 15] <= -1e19) and np.all(ub[L.g_box(1, 3):L.g_box(1, 3) + 15] >= 1e19)
    assert np.allclose(lb[L.g_box(1, 8):L.g_box(1, 8) + 3], [-0.01, -0.05, 0.0]) and np.allclose(ub[L.g_box(1, 8):L.g_box(1, 8) + 3], [0.01, 0.0, 0.0])
    # friction rows: one sided
    assert np.all(lb[L.g_fric(0, 0, 0):L.g_fric(0, 0, 0) + 16 * 12] <= -1e19) and np.all(ub[L.g_fric(0, 0, 0):L.g_fric(0, 0, 0) + 16 * 12] == 0)


def test_resample_linear_matches_numpy(H):
    rng = np.random.default_rng(0)
    t_in = np.arange(0, 60) * 0.02 * 1.5            # 50 Hz planner samples with a slow-down factor
    p_in = rng.normal(size=(60, 3))
    t_out = np.arange(0, 13) * 0.1
    out = H.resample_linear(t_in, p_in, t_out)
    ref = np.stack([np.interp(t_out, t_in, p_in[:, a]) for a in range(3)], axis=1)
    assert np.allclose(out, ref, atol=1e-12)
    assert np.allclose(H.resample_linear(t_in, p_in, [-1.0, 100.0]), [p_in[0], p_in[-1]])   # end points are held
    with pytest.raises(ValueError):
        H.resample_linear([0.0, 0.0], p_in[:2], t_out)                                       # times must increase


@pytest.mark.parametrize("phase", [0, 3, 4, 7, 8, 11, 15, 21])
def test_tick_record_matches_the_synthetic_workloads(H, workloads, phase):
    """the compact tick record the C++ operator uploads (state, references, contact windows; expanded on the device by
    cmpc_populate) against the record workloads.walk_batch(ticks=True) writes for the same instance"""
    N = 12
    w = workloads.walk_batch(N=N, B=1, seed=3, phase=phase, state_noise=1.0, ticks=True)
    L = pkg("layout").Layout(N)
    p, tk = w["p"][0], w["ticks"][0]
    m = H.CentroidalMPCHost(os.path.join(DATA, "ergocub", "centroidal_mpc.ini"), "")
    g0 = L.p_glob()
    assert m.set_state(p[g0:g0 + 3], p[g0 + 3:g0 + 6], p[g0 + 6:g0 + 9], wrench=[0.5, -1.0, 0.25, 0.01, 0.02, 0.03])
    assert m.set_reference_trajectory(p[L.p_comref(0):L.p_comref(0) + 3 * (N + 1)], p[L.p_href(0):L.p_href(0) + 3 * (N + 1)])
    assert m.set_contact_phase_list(H.walk_contact_lists(phase))
    t = m.tick_record()
    assert t.shape == tk.shape == (workloads.tick_stride(N),)
    assert np.array_equal(t[0:9], tk[0:9]) and np.allclose(t[9:15], [0.5, -1.0, 0.25, 0.01, 0.02, 0.03]) and t[15] == 1.0
    assert np.array_equal(t[17:17 + 6 * (N + 1)], tk[17:17 + 6 * (N + 1)])
    horizon_ns = N * 1e8
    for c in range(2):
        base = 17 + 6 * (N + 1) + 85 * c
        n_cpp, n_np = int(t[base]), int(tk[base])
        assert 1 <= n_cpp <= 6
        rec = lambda a, j: a[base + 1 + 14 * j:base + 15 + 14 * j]  # noqa: E731
        # same window start (the contact the foot stands / last stood on); the C++ window ends with the first contact that
        # starts after the horizon, the numpy one always holds four
        for j in range(min(n_cpp, n_np)):
            a, b = rec(t, j), rec(tk, j)
            assert a[0] == b[0] or (a[0] < -1e10 and b[0] < -1e10), (c, j, a[0], b[0])    # "always there": any time far in the past
            assert a[1] == b[1] and np.allclose(a[2:], b[2:], atol=1e-15), (c, j)
        assert rec(t, n_cpp - 1)[0] > horizon_ns or n_cpp >= n_np
