"""The oracle's actuator restatement (Robot::actuate_axis) against the REFERENCE's own
TruncatedNormalUncertainVelocityActuator: simple_uncertainty_models.hpp compiled from /root/reference into
oracle/_ref/unc_ref against a two-name stand-in of arc_utilities (oracle/shim), draws injected -- plus the committed golden
vectors generated from that binary (tests/golden/actuator_golden.json, made by tests/golden/make_golden.py).  Bit-exact: the
arithmetic is a clamp, two products, a max, a product and a sum."""
import json
import os

import numpy as np
import pytest

from oracle import oracle_binding as OB

HERE = os.path.dirname(os.path.abspath(__file__))


def test_golden_vectors():
    g = json.load(open(os.path.join(HERE, "golden", "actuator_golden.json")))
    assert len(g["cases"]) >= 20
    for case in g["cases"]:
        vl, al, pn, mn, pv = case["params"]
        quiet, noisy = OB.actuator_oracle(vl, pn, mn, case["controls"], case["draws"])
        assert np.array_equal(quiet, np.array(case["quiet"])) and np.array_equal(noisy, np.array(case["noisy"])), \
            "actuator restatement is not bit-identical to the reference"
        # what the reference hands its noise distribution (unc.hpp:61): N(0, clamp(|percent_variance|, 0, 1)) truncated to [-1, 1]
        # -- the distribution fks_philox_truncated_normal and the oracle's mt19937 sampler draw from
        assert case["distribution"] == [0.0, min(abs(pv), 1.0), -1.0, 1.0]


@pytest.mark.skipif(not os.path.exists(OB.UNC_REF), reason="oracle/_ref/unc_ref not built (reference tree absent)")
def test_against_reference_binary():
    rng = np.random.default_rng(11)
    for _ in range(40):
        vl, pn, mn = rng.uniform(-2.0, 2.0), rng.uniform(-0.5, 0.5), rng.uniform(-0.1, 0.1)   # the reference takes |.| of each
        pv = rng.choice([0.5, -0.5, 0.0, 1.7, 0.25])
        controls = np.concatenate([rng.normal(0.0, 1.5, 60), [0.0, vl, -vl, 10.0 * vl, 1e-300, -1e300]])
        draws = np.concatenate([rng.uniform(-1.0, 1.0, 60), [1.0, -1.0, 0.0, 0.5, -0.5, 1.0]])
        rq, rn, dist = OB.actuator_reference(vl, 0.3, pn, mn, pv, controls, draws)
        oq, on = OB.actuator_oracle(vl, pn, mn, controls, draws)
        assert np.array_equal(rq, oq) and np.array_equal(rn, on)
        assert dist == (0.0, min(abs(pv), 1.0), -1.0, 1.0)
