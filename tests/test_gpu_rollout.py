"""Closed loop on the GPU (BASELINE config 4, reduced): MPC solve + warm-start shift + RK4 plant every tick, with pushes."""
import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_closed_loop_walk_with_pushes():
    P = pkg()
    R = pkg("rollout")
    s = P.BatchedCentroidalMPC(P.ergocub_config(ipopt_tolerance=1e-6))
    out = R.closed_loop_rollout(s, B=48, ticks=40, seed=1, push_range=(1.0, 2.0))
    s.close()
    assert np.all(out["converged_ticks"] >= 39), out["converged_ticks"]     # a failed tick would stop the reference's runner
    assert np.all(out["com_z_min"] > 0.6) and np.all(out["com_err_max"] < 0.15), (out["com_z_min"].min(), out["com_err_max"].max())
    assert out["iterations"].mean() / 40 < 20


def test_device_schedule_equals_host_schedule():
    """the device-resident schedule table drives the closed loop exactly like the per-tick host generator"""
    P = pkg()
    R = pkg("rollout")
    s = P.BatchedCentroidalMPC(P.ergocub_config(ipopt_tolerance=1e-6))
    a = R.closed_loop_rollout(s, B=24, ticks=24, seed=3, push_range=(1.0, 2.0))
    b = R.closed_loop_rollout(s, B=24, ticks=24, seed=3, push_range=(1.0, 2.0), host_schedule=True)
    s.close()
    assert np.array_equal(a["converged_ticks"], b["converged_ticks"]) and np.all(a["converged_ticks"] == 24)
    assert np.max(np.abs(a["iterations"] - b["iterations"])) <= 2
    assert np.allclose(a["com_err_max"], b["com_err_max"], atol=1e-6) and np.allclose(a["com_z_min"], b["com_z_min"], atol=1e-6)
