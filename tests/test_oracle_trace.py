"""Step trace (ForwardSimulationStepTrace filled at spcs.hpp:1583-1617, :1703, :1714, :1778) of the oracle:
structural invariants that follow from the reference's control flow, on the CPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import parity  # noqa: E402
from fast_kinematic_simulator_b200 import capi, workloads as W  # noqa: E402


def _check(res, tr, allow_contacts):
    kinds = tr["kind"]
    # one control_input + control_input_step pair per controller step, one post-action configuration per microstep, one
    # resolution step per resolver iteration (spcs.hpp:1585-1587, :1617, :1703)
    assert (kinds == capi.TRACE_CONTROL_INPUT).sum() == res["n_steps"][0]
    assert (kinds == capi.TRACE_CONTROL_INPUT_STEP).sum() == res["n_steps"][0]
    assert (kinds == capi.TRACE_POST_ACTION).sum() == res["n_microsteps"][0]
    assert (kinds == capi.TRACE_RESOLUTION_STEP).sum() == res["n_resolver_iters"][0]
    # records are appended in simulation order
    order = tr["step"].astype(np.int64) * 1_000_000 + tr["microstep"].astype(np.int64) * 100 + tr["iteration"]
    assert np.all(np.diff(order[kinds >= capi.TRACE_POST_ACTION]) >= 0)
    # control_input_step = control_input / number of microsteps of that step
    for st in range(int(res["n_steps"][0])):
        u = tr[(kinds == capi.TRACE_CONTROL_INPUT) & (tr["step"] == st)]["values"][0]
        du = tr[(kinds == capi.TRACE_CONTROL_INPUT_STEP) & (tr["step"] == st)]["values"][0]
        n_micro = int(np.round(np.abs(u).max() / np.abs(du).max())) if np.abs(du).max() > 0 else 1
        assert np.allclose(du * n_micro, u, rtol=1e-12, atol=0)
    failed = bool(res["flags"][0] & capi.FLAG_RESOLVE_FAILED)
    stopped = bool(res["flags"][0] & capi.FLAG_ENDED_BY_NOCONTACT)
    assert (kinds == capi.TRACE_RETURNED_PREVIOUS).sum() == int(failed) + int(stopped)
    if failed:  # a failed resolve returns the configuration before the microstep, which is what the particle keeps (spcs.hpp:1714,1745)
        assert kinds[-1] == capi.TRACE_RETURNED_PREVIOUS
        assert np.array_equal(tr["values"][-1][: res["cfg"].shape[1]], res["cfg"][0])
    elif allow_contacts:  # the last recorded configuration is where the particle ends
        assert np.array_equal(tr["values"][-1][: res["cfg"].shape[1]], res["cfg"][0])


@pytest.mark.parametrize("name,pid", [("se2_arena", 3), ("arm_elbow", 0), ("se3_narrow_passage", 5), ("arm_free", 1)])
def test_oracle_trace_structure(name, pid):
    w = W.make(name, n_particles=8)
    o = parity.make_oracle(w)
    t = w.targets[0] if w.targets.shape[0] == 1 else w.targets[pid]
    res, tr = o.forward_simulate_traced(w.starts[pid], t, True, capi.NOISE_PHILOX, particle_id=pid)
    _check(res, tr, True)
    # the traced call simulates the same particle as the batch call
    batch = o.forward_simulate(w.starts[pid:pid + 1], t.reshape(1, -1), True, capi.NOISE_PHILOX, first_particle_id=pid)
    batch = batch[0] if isinstance(batch, tuple) else batch
    assert batch.tobytes() == res.tobytes()


def test_oracle_trace_without_contacts_returns_the_previous_configuration():
    w = W.se2_arena(4)
    o = parity.make_oracle(w)
    res, tr = o.forward_simulate_traced(w.starts[0], w.targets[0], False, capi.NOISE_PHILOX, particle_id=0)
    assert res["flags"][0] & capi.FLAG_ENDED_BY_NOCONTACT
    _check(res, tr, False)
    assert tr["kind"][-1] == capi.TRACE_RETURNED_PREVIOUS and tr["kind"][-2] == capi.TRACE_POST_ACTION
