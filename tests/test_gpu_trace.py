"""Device-side step trace (fks_forward_simulate_traced) against the oracle's trace: same records in the
same order (kind / step / microstep / iteration exact, values within 1e-9), and the traced particle ends exactly where
the batch call puts it."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import parity  # noqa: E402
from fast_kinematic_simulator_b200 import capi, workloads as W  # noqa: E402

pytestmark = pytest.mark.gpu


def _compare(name, pid, allow_contacts=True):
    w = W.make(name, n_particles=8)
    t = w.targets[0] if w.targets.shape[0] == 1 else w.targets[pid]
    o = parity.make_oracle(w)
    ref, rtr = o.forward_simulate_traced(w.starts[pid], t, allow_contacts, capi.NOISE_PHILOX, particle_id=pid)
    sim = w.make_simulator()
    got, gtr = sim.forward_simulate_robot_traced(w.starts[pid], t, allow_contacts, capi.NOISE_PHILOX, particle_id=pid)
    assert len(gtr) == len(rtr) > 0
    for f in ("kind", "step", "microstep", "iteration"):
        assert np.array_equal(gtr[f], rtr[f]), f
    err = np.abs(gtr["values"] - rtr["values"]) / np.maximum(1.0, np.abs(rtr["values"]))
    assert err.max() <= parity.RTOL
    assert got.records["flags"][0] & parity.SEMANTIC_FLAGS == ref["flags"][0] & parity.SEMANTIC_FLAGS
    # the batch path (no trace) gives the same record, byte for byte
    batch = sim.forward_simulate_robots(w.starts[pid:pid + 1], t.reshape(1, -1), allow_contacts, capi.NOISE_PHILOX, first_particle_id=pid)
    assert batch.records.tobytes() == got.records.tobytes()
    return gtr


@pytest.mark.parametrize("name,pid", [("se2_arena", 3), ("arm_elbow", 0), ("se3_narrow_passage", 5), ("arm_free", 1), ("gantry", 2)])
def test_trace_matches_oracle(name, pid):
    tr = _compare(name, pid)
    if name in ("se2_arena", "arm_elbow"):
        assert (tr["kind"] == capi.TRACE_RESOLUTION_STEP).any()
    if name == "arm_elbow":  # every particle of this workload ends in a failed resolve
        assert tr["kind"][-1] == capi.TRACE_RETURNED_PREVIOUS


def test_trace_without_contacts():
    tr = _compare("se2_arena", 0, allow_contacts=False)
    assert tr["kind"][-1] == capi.TRACE_RETURNED_PREVIOUS


def test_trace_capacity_is_respected():
    w = W.se2_arena(4)
    sim = w.make_simulator()
    with pytest.raises(capi.FksError):
        sim.forward_simulate_robot_traced(w.starts[0], w.targets[0], True, capi.NOISE_PHILOX, particle_id=0, capacity=10)
