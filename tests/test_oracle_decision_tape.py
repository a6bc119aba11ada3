"""The DECISION TAPE protocol (include/fksgpu.h, fks_noise_tape.decisions), checked on the CPU.

The reference's contact solve (J.colPivHouseholderQr().solve(c), spcs.hpp:1990-1998) takes its rank decision on a pivot
that is pure round-off whenever a single distal link of the arm is in contact, so no second implementation -- another
summation order on x86, a GPU -- reproduces individual trajectories of BASELINE config 3 without being told those
decisions.  These tests replay the oracle's recorded decisions through a solver with DIFFERENT arithmetic (the CPU model
of the round-1 device QR, oracle/fks_qr_model.cpp: butterfly sums, FMA contraction, recomputed norms) and check what the
GPU parity tests rely on:
  * with the tape, every particle whose solves stayed well conditioned (estimate < 1e3) reproduces to 1e-9, and nearly
    all the others do;
  * with ill-conditioned solves injected too, every particle reproduces;
  * without the tape the same arithmetic diverges on most particles in contact (so the tape is what makes them comparable).
"""
import numpy as np

from fast_kinematic_simulator_b200 import capi, workloads as W
from oracle import oracle_binding as OB

import parity

N = 256
MODEL_ARITHMETIC = 3          # fma everywhere (what nvcc makes of the round-1 solver)
REPLAY_DECISIONS = 4          # take rank / order / flagged solutions from the tape


def _replay(w, tape, mode, starts=None):
    o2 = parity.make_oracle(w)
    OB.lib().oracle_set_qr_model(o2._h, mode)
    return o2.forward_simulate(w.starts if starts is None else starts, w.targets, True, capi.NOISE_INJECTED, tape)


def _matches(a, b):
    err = np.max(np.abs(a["cfg"] - b["cfg"]) / np.maximum(1.0, np.abs(b["cfg"])), axis=1)
    same = ((a["flags"] & 0xFFFF) == (b["flags"] & 0xFFFF)) & (a["n_microsteps"] == b["n_microsteps"]) & \
        (a["n_resolver_iters"] == b["n_resolver_iters"]) & (a["n_steps"] == b["n_steps"])
    return same & (err <= parity.RTOL)


def test_decision_records_are_consistent():
    w = W.arm_table(N)
    orc = parity.make_oracle(w)
    ref, tape, sens = OB.run_with_tape(orc, w.starts, w.targets, True)
    d = OB.decision_records(tape, 7)
    # one record per solve = per resolver iteration (no empty Jacobians on this workload)
    per_particle = np.diff(d["offsets"])
    assert np.array_equal(per_particle, ref["n_resolver_iters"])
    size = np.minimum(d["rows"], 7)
    assert np.all(d["rank"] <= size) and np.all(d["rows"] % 3 == 0) and np.all(d["rows"] > 0)
    for k in range(len(size)):  # the order lists distinct columns
        cols = [(int(d["order"][k]) >> (4 * i)) & 15 for i in range(int(size[k]))]
        assert len(set(cols)) == len(cols) and max(cols) < 7
    roundoff = (d["flags"] & OB.DECISION_ROUNDOFF_PIVOT) != 0
    kept = (d["flags"] & OB.DECISION_OVERRIDE_SOLUTION) != 0
    assert roundoff.sum() > 0.2 * len(size)             # rank <= 6 in 7 unknowns is the normal case here
    assert 0 < kept.sum() < 0.25 * roundoff.sum()       # ... and Eigen keeps the round-off pivot now and then
    assert np.all(d["rank"][kept] == size[kept]) and np.all(roundoff[kept])
    assert np.all(np.isfinite(d["solution"]))


def test_other_arithmetic_needs_the_tape_and_reproduces_with_it():
    w = W.arm_table(N)
    orc = parity.make_oracle(w)
    ref, tape, sens = OB.run_with_tape(orc, w.starts, w.targets, True)
    cond = OB.max_condition_of_last_call(orc, N)
    free = _matches(_replay(w, tape[:2], MODEL_ARITHMETIC), ref)
    taped = _matches(_replay(w, tape, MODEL_ARITHMETIC | REPLAY_DECISIONS), ref)
    print("matching without the tape %d, with it %d of %d; well-conditioned particles %d" % (free.sum(), taped.sum(), N, (cond < 1e3).sum()))
    assert free.sum() < 0.6 * N                      # the same noise, another last bit: most particles in contact diverge
    assert taped.sum() >= 0.95 * N
    assert np.all(taped[cond < 1e3])                 # no exception without the enumerated reason
    assert (cond < 1e3).sum() > 0.5 * N
    assert np.array_equal(((sens & OB.SENS_ILL_CONDITIONED) != 0), cond > 1e3)


def test_everything_reproduces_with_ill_conditioned_solves_injected():
    w = W.arm_table(N)
    orc = parity.make_oracle(w)
    OB.lib().oracle_set_decision_cond_limit(orc._h, 100.0)
    ref, tape, sens = OB.run_with_tape(orc, w.starts, w.targets, True)
    rep = _replay(w, tape, MODEL_ARITHMETIC | REPLAY_DECISIONS)
    assert _matches(rep, ref).all()
    assert not np.any(rep["flags"] & capi.FLAG_DECISION_DESYNC)
    assert not np.any(sens & OB.SENS_ILL_CONDITIONED)
    d = OB.decision_records(tape, 7)
    injected = ((d["flags"] & OB.DECISION_OVERRIDE_SOLUTION) != 0).mean()
    assert 0.1 < injected < 0.7                      # the majority of the solves is still computed, not injected


def test_se3_pivot_ties_are_on_the_tape():
    """[R | R (e x p)]: the translation columns have equal norms up to round-off -> the reference's first pivots are ties."""
    w = W.se3_narrow_passage(128)
    orc = parity.make_oracle(w)
    ref, tape, sens = OB.run_with_tape(orc, w.starts, w.targets, True)
    d = OB.decision_records(tape, 6)
    assert ((d["flags"] & OB.DECISION_PIVOT_TIE) != 0).any()
    taped = _matches(_replay(w, tape, MODEL_ARITHMETIC | REPLAY_DECISIONS), ref)
    assert taped.mean() > 0.98


def test_device_solver_model_agrees_with_the_eigen_restatement_when_well_conditioned():
    rng = np.random.default_rng(5)
    for rows, cols in ((3, 7), (9, 7), (24, 7), (63, 6), (130, 7), (45, 3)):
        A = rng.normal(size=(rows, cols))
        b = rng.normal(size=rows)
        x0, r0, o0, f0 = OB.qr_solve_info(A, b)
        for bits in (0, 3):
            x1, r1, o1, ratio = OB.qr_device_model(A, b, bits)
            assert r1 == r0 == min(rows, cols) and o1 == o0
            assert np.max(np.abs(x1 - x0)) < 1e-11 * max(1.0, np.max(np.abs(x0)))
