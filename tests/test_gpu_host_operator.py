"""GPU tests of the C++ host operator (drop-in CentroidalMPC): advance() through the C ABI against the CPU oracle, the warm-started
tick-to-tick loop, advanceBatch()."""
import os

import numpy as np
import pytest

from conftest import ROOT, pkg
from oracle.oracle import make_cfg

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
DATA = os.path.join(ROOT, "tests", "data")


def setup(H, workloads, phase, seed=11, noise=1.0, ini=None):
    w = workloads.walk_batch(N=12, B=1, seed=seed, phase=phase, state_noise=noise)
    if ini is not None:
        m = H.CentroidalMPCHost(ini, "")
    else:
        m = H.CentroidalMPCHost(os.path.join(DATA, "ergocub", "centroidal_mpc_walking.ini"), "TRAJECTORY_ADJUSTMENT/CENTROIDAL_MPC")
    L, p = m.L, w["p"][0]
    g0 = L.p_glob()
    assert m.set_state(p[g0:g0 + 3], p[g0 + 3:g0 + 6], p[g0 + 6:g0 + 9])
    assert m.set_reference_trajectory(p[L.p_comref(0):L.p_comref(0) + 39], p[L.p_href(0):L.p_href(0) + 39])
    assert m.set_contact_phase_list(H.walk_contact_lists(phase))
    return m, w


@pytest.mark.parametrize("tol", [1e-4, 1e-8], ids=["ini-tolerance", "tight"])
def test_advance_matches_oracle(oracle, workloads, tol, tmp_path):
    """advance() of the drop-in class against the oracle at the ini's own ipopt_tolerance (1e-4: both sides stop early, loose
    bounds) and at 1e-8 (the boundary bar: objective 1e-6 relative, trajectories / forces / footsteps 1e-5)"""
    H = pkg("host")
    ini = None
    if tol != 1e-4:
        src = open(os.path.join(DATA, "ergocub", "centroidal_mpc.ini")).read()
        assert "ipopt_tolerance                 1e-4" in src
        ini = str(tmp_path / "centroidal_mpc.ini")
        open(ini, "w").write(src.replace("ipopt_tolerance                 1e-4", f"ipopt_tolerance                 {tol:g}"))
    m, w = setup(H, workloads, phase=5, ini=ini)
    assert abs(m.config().ipopt_tolerance - tol) < 1e-20
    assert m.advance(), m.last_error()
    assert m.is_output_valid()
    st, it, obj = m.stats()
    o = make_cfg(N=12, w_pos=2000.0)
    O = oracle
    xo, lo, so = O.solve_batch(o, w["p"], w["lbg"], w["ubg"], w["x0"], threads=1, opts=O.default_opts(tol=tol))
    assert st == 0 and so[0].status == 0
    tight = tol <= 1e-8
    assert abs(obj - so[0].obj) <= (1e-6 if tight else 1e-4) * max(1.0, abs(so[0].obj))
    L = m.L
    com, dcom, h = m.trajectories()
    assert np.max(np.abs(com.reshape(-1) - xo[0][:39])) < (1e-5 if tight else 1e-3)
    assert np.max(np.abs(dcom.reshape(-1) - xo[0][39:78])) < (1e-5 if tight else 1e-2)
    assert np.max(np.abs(h.reshape(-1) - xo[0][78:117])) < (1e-5 if tight else 1e-2)
    # knot-0 corner forces: right foot is in swing at phase 5 -> zero; left foot carries the weight
    posl, Rl, fl = m.contact_output("left_foot")
    posr, Rr, fr = m.contact_output("right_foot")
    assert np.all(fr == 0.0)
    fo = np.array([xo[0][L.x_frc(0, j, 0):L.x_frc(0, j, 0) + 3] for j in range(4)])
    assert abs(fl[:, 2].sum() - fo[:, 2].sum()) < (1e-5 * 9.81 if tight else 1e-2)
    if tight:
        assert np.max(np.abs(fl - fo)) <= 1e-5 * max(1.0, np.max(np.abs(fo)))     # ergoCub: symmetry weight > 0, unique split
    # the right foot lands inside the horizon: its adjusted position is reported and edited into the phase list
    nxt = m.next_planned_contact("right_foot")
    assert nxt is not None
    land = int(round(nxt[1] / 0.1))
    assert np.allclose(nxt[0], xo[0][L.x_pos(1, land + 1):L.x_pos(1, land + 1) + 3], atol=1e-5 if tight else 1e-3)
    lst = m.output_contact_list("right_foot")
    assert any(abs(c[0] - nxt[1]) < 1e-12 and np.allclose(c[2], nxt[0]) for c in lst)
    assert abs(m.current_time() - 0.1) < 1e-15                          # one sampling time per advance()


def test_warm_started_ticks_and_batch(workloads):
    H = pkg("host")
    ms = []
    for b, phase in enumerate((0, 3, 9)):
        m, w = setup(H, workloads, phase=phase, seed=20 + b, noise=0.5)
        ms.append(m)
    assert H.CentroidalMPCHost.advance_batch(ms), ms[0].last_error()
    its0 = [m.stats()[1] for m in ms]
    # next tick: feed the MPC's own prediction back as the state (the reference integrates the model)
    for m, phase in zip(ms, (0, 3, 9)):
        com, dcom, h = m.trajectories()
        assert m.set_state(com[1], dcom[1], h[1])
        ref = np.vstack([com[1:], com[-1:]])
        ref[:, 2] = 0.7
        assert m.set_reference_trajectory(ref, np.zeros((13, 3)))
        assert m.set_contact_phase_list(H.walk_contact_lists(phase))      # absolute-time list; the controller's clock moved on
    assert H.CentroidalMPCHost.advance_batch(ms), ms[0].last_error()
    its1 = [m.stats()[1] for m in ms]
    assert all(m.stats()[0] == 0 for m in ms)
    # is_warm_start_enabled true: the shifted solution is the initial guess; the monotone barrier schedule (mu_init 0.1)
    # keeps the iteration count from collapsing, but it must not grow
    assert sum(its1) <= sum(its0), (its0, its1)
    for m in ms:
        assert abs(m.current_time() - 0.2) < 1e-15


def test_block_tick_sequence_with_merged_contact_lists(workloads):
    """the per-tick sequence of CentroidalMPCBlock::advance(): planner lists merged with the MPC's own (adjusted) current
    contact (updateContactPhaseList), solve, desired ZMP from the corner forces"""
    H = pkg("host")
    phase = 6                                        # right foot in mid swing: it lands inside the horizon
    m, w = setup(H, workloads, phase=phase, seed=31, noise=1.0)
    planner = H.walk_contact_lists(phase)
    assert m.set_planner_contact_lists(planner, force_sample_time=0.1, first_run=True) == 0
    assert m.advance(), m.last_error()
    nxt = m.next_planned_contact("right_foot")
    assert nxt is not None
    z = m.desired_zmp()
    posl, Rl, fl = m.contact_output("left_foot")
    assert z is not None and abs(z[0] - posl[0]) <= 0.08 + 1e-9 and abs(z[1] - posl[1]) <= 0.03 + 1e-9   # single support: ZMP in the left foot
    # tick until the right foot has landed: from then on the merged list must carry the MPC's adjusted pose, not the planner's
    landed_pose = None
    for tick in range(1, 8):
        com, dcom, h = m.trajectories()
        assert m.set_state(com[1], dcom[1], h[1])
        ref = np.vstack([com[1:], com[-1:]])
        ref[:, 2] = 0.7
        assert m.set_reference_trajectory(ref, np.zeros((13, 3)))
        land_plan = m.next_planned_contact("right_foot")
        if land_plan is not None:
            landed_pose = land_plan
        assert m.set_planner_contact_lists(planner, force_sample_time=0.1, first_run=False) == 0, m.last_error()
        assert m.advance(), m.last_error()
    assert landed_pose is not None
    t_now = m.current_time() - 0.1                   # time of the last solve
    cur = [c for c in m.output_contact_list("right_foot") if c[0] <= t_now + 1e-9 < c[1]]
    assert len(cur) == 1
    planned = [c for c in planner["right_foot"] if c[0] <= t_now + 1e-9 < c[1]][0]
    assert np.allclose(cur[0][2], landed_pose[0], atol=2e-3)            # the MPC's own landing position is kept ...
    assert abs(cur[0][0] - planned[0]) < 1e-9 and abs(cur[0][1] - planned[1]) < 1e-9   # ... with the planner's timing
