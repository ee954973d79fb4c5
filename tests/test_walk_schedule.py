"""The device-resident walk-schedule table of the closed-loop driver (rollout.WalkSchedule) against the host generator
(workloads.walk_batch), on CPU tensors: same p / lbg / ubg / x0 / references / contact bookkeeping for any phase."""
import numpy as np
import pytest
import torch

from conftest import pkg


@pytest.mark.parametrize("N,step_adjust", [(12, True), (15, False)])
def test_walk_schedule_table_matches_host_generator(workloads, N, step_adjust):
    R = pkg("rollout")
    S = R.WalkSchedule(N, 0.1, torch.device("cpu"), step_adjust=step_adjust)
    ph = np.array([0, 1, 5, 15, 16, 17, 31, 32, 33, 47, 48, 100, 115, 116, 1000])
    ref = workloads.walk_batch(N=N, dT=0.1, B=len(ph), phase=ph, step_adjust=step_adjust)
    got = S(torch.from_numpy(ph))
    for k in ("p", "lbg", "ubg", "x0"):
        a, b = got[k].numpy(), ref[k]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin) and np.array_equal(a[~fin], b[~fin]), k     # same infinite bounds
        assert np.max(np.abs(a[fin] - b[fin])) < 1e-12, k
    assert np.max(np.abs(got["comref"].numpy() - ref["comref"])) < 1e-12
    assert np.array_equal(got["same_contact"].numpy(), ref["same_contact"])
    assert np.array_equal(got["stance0"].numpy(), ref["stance0"])
