"""The oracle's restatement of the batch consumer: split by did_contact and ComputeConfigurationDistanceTo
(the call at spcs.hpp:898) -- known answers and metric properties, on the CPU."""
import numpy as np

from fast_kinematic_simulator_b200 import capi, workloads as W
from oracle import oracle_binding as OB

import parity


def test_partition_is_stable_and_complete():
    rng = np.random.default_rng(5)
    flags = rng.integers(0, 8, 1000).astype(np.uint32)
    order, a, b = OB.end_states_partition(flags)
    contact = (flags & capi.FLAG_DID_CONTACT) != 0
    assert a == int((~contact).sum()) and b == int(contact.sum())
    assert np.array_equal(order[:a], np.flatnonzero(~contact)) and np.array_equal(order[a:], np.flatnonzero(contact))
    assert OB.end_states_partition(np.zeros(0, np.uint32))[1:] == (0, 0)


def test_se2_distance_known_answers():
    w = W.make("se2_arena", n_particles=4)
    orc = parity.make_oracle(w)
    cfg = np.array([[0.0, 0.0, 0.0], [3.0, 4.0, 0.0], [0.0, 0.0, np.pi - 0.1], [0.0, 0.0, -np.pi + 0.1]])
    d = orc.pairwise_config_distance(cfg)
    pw, rw = w.robot.pos_w, w.robot.rot_w
    assert d[0, 1] == 5.0 * pw and d[1, 0] == d[0, 1]
    assert abs(d[2, 3] - 0.2 * rw) < 1e-15 and abs(d[0, 2] - (np.pi - 0.1) * rw) < 1e-15   # the angle wraps
    assert np.all(np.diag(d) == 0.0)


def test_distance_is_a_metric_on_arm_and_se3_end_states():
    for name in ("arm_table", "se3_narrow_passage"):
        w = W.make(name, n_particles=48)
        orc = parity.make_oracle(w)
        rec = orc.forward_simulate(w.starts, w.targets, True)
        d = orc.pairwise_config_distance(rec["cfg"])
        assert np.all(np.isfinite(d)) and np.all(d >= 0.0) and np.allclose(d, d.T, rtol=0, atol=1e-7)
        assert np.all(np.diag(d) <= 2e-7)  # SE3: acos of a trace that is 1 up to round-off, sqrt(2 k eps) ~ 1e-7 rad
        assert np.all(d[:, :, None] <= d[:, None, :] + d[None, :, :] + 1e-7)  # d(i,j) <= d(i,k) + d(k,j)
