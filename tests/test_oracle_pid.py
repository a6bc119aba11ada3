"""The oracle's PID restatement against the REFERENCE's own simple_pid_controller.hpp, compiled from
/root/reference into oracle/_ref/pid_ref (the one reference file that builds stand-alone), plus the
committed golden vectors generated from it (tests/golden/pid_golden.json, made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle_binding as OB

HERE = os.path.dirname(os.path.abspath(__file__))


def oracle_pid(kp, ki, kd, ic, errs, dts):
    errs = np.ascontiguousarray(errs, dtype=np.float64)
    dts = np.ascontiguousarray(dts, dtype=np.float64)
    out = np.zeros(len(errs))
    OB.lib().oracle_pid_run(kp, ki, kd, ic, errs.ctypes.data, dts.ctypes.data, len(errs), out.ctypes.data)
    return out


def test_known_answers_from_reference_header():
    # SimplePIDController(1, 0.5, 0.1, 2): ComputeFeedbackTerm(1.0, 0.04) = 3.51, then (0.5, 0.04) = -0.725
    out = oracle_pid(1.0, 0.5, 0.1, 2.0, [1.0, 0.5], [0.04, 0.04])
    assert out[0] == pytest.approx(3.51, abs=1e-12)
    assert out[1] == pytest.approx(-0.725, abs=1e-12)


def test_golden_vectors():
    g = json.load(open(os.path.join(HERE, "golden", "pid_golden.json")))
    for case in g["cases"]:
        out = oracle_pid(*case["gains"], case["errors"], case["timesteps"])
        assert np.array_equal(out, np.array(case["outputs"])), "PID restatement is not bit-identical to the reference"


@pytest.mark.skipif(not os.path.exists(OB.PID_REF), reason="oracle/_ref/pid_ref not built (reference tree absent)")
def test_against_reference_binary():
    rng = np.random.default_rng(7)
    for _ in range(20):
        kp, ki, kd, ic = rng.uniform(-3, 3, 4)
        errs = rng.normal(0, 2, 50)
        dts = rng.uniform(0.01, 0.1, 50)
        ref = OB.pid_reference(kp, ki, kd, ic, errs, dts)
        assert np.array_equal(oracle_pid(kp, ki, kd, ic, errs, dts), ref)
