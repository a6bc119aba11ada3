"""Known-answer and cross-check tests of the oracle's restated primitives (the reference has no
tests, so the builder pins every primitive against an independent implementation: scipy / numpy / closed forms)."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg

from oracle import oracle_binding as OB


def qr_solve(A, b):
    A = np.asarray(A, dtype=np.float64)
    rows, cols = A.shape
    Af = np.asfortranarray(A)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(cols)
    sens = C.c_uint32(0)
    OB.lib().oracle_colpiv_qr_solve(Af.ctypes.data, bb.ctypes.data, rows, cols, x.ctypes.data, C.addressof(sens))
    return x, sens.value


def test_qr_full_rank_matches_lstsq():
    rng = np.random.default_rng(0)
    for rows, cols in [(3, 3), (9, 6), (12, 7), (30, 7), (6, 6), (21, 3)]:
        A = rng.normal(size=(rows, cols))
        b = rng.normal(size=rows)
        x, _ = qr_solve(A, b)
        ref = np.linalg.lstsq(A, b, rcond=None)[0]
        assert np.allclose(x, ref, rtol=1e-10, atol=1e-12)


def test_qr_pivot_order_matches_scipy():
    rng = np.random.default_rng(1)
    A = rng.normal(size=(12, 7)) * np.array([1, 5, 0.1, 2, 9, 0.5, 3])
    b = rng.normal(size=12)
    x, _ = qr_solve(A, b)
    Q, R, P = scipy.linalg.qr(A, pivoting=True, mode="economic")
    y = scipy.linalg.solve_triangular(R, Q.T @ b)
    ref = np.zeros(7)
    ref[P] = y
    assert np.allclose(x, ref, rtol=1e-10, atol=1e-12)


def test_qr_underdetermined_is_basic_solution():
    """3 x 6 (one SE3 contact point): Eigen returns the BASIC solution -- the non-pivot unknowns are zero."""
    rng = np.random.default_rng(2)
    A = rng.normal(size=(3, 6))
    b = rng.normal(size=3)
    x, _ = qr_solve(A, b)
    assert np.count_nonzero(x) == 3
    assert np.allclose(A @ x, b, atol=1e-12)
    # the pivots are the greedy largest-residual columns: first pivot = largest column norm
    assert x[np.argmax(np.linalg.norm(A, axis=0))] != 0.0


def test_qr_structural_zero_columns_and_rows():
    """SE2: every point contributes an all-zero z row; linked: distal joints give all-zero columns."""
    rng = np.random.default_rng(3)
    A = rng.normal(size=(9, 7))
    A[:, 4:] = 0.0
    A[2::3, :] = 0.0
    b = rng.normal(size=9)
    x, _ = qr_solve(A, b)
    assert np.all(x[4:] == 0.0)
    ref = np.linalg.lstsq(A[:, :4], b, rcond=None)[0]
    assert np.allclose(x[:4], ref, rtol=1e-10, atol=1e-12)


def test_qr_all_zero_matrix_is_not_rank_cut():
    """Eigen's rank test is `norm^2 < threshold`, and for an all-zero matrix the threshold is 0: 0 < 0 is false, no
    pivot is cut and the triangular solve divides by zero.  The restatement keeps that (the simulator then raises
    FKS_FLAG_WOULD_ASSERT_NAN, where the reference's ApplyControlInput asserts, unc.hpp:72-73)."""
    x, _ = qr_solve(np.zeros((6, 3)), np.ones(6))
    assert not np.all(np.isfinite(x))


def test_exp_twist_matches_expm():
    rng = np.random.default_rng(4)
    for _ in range(20):
        tw = rng.normal(size=6) * rng.choice([1e-3, 0.1, 1.0, 3.0])
        out = np.zeros(12)
        OB.lib().oracle_exp_twist(tw.ctypes.data, out.ctypes.data)
        v, w = tw[:3], tw[3:]
        X = np.zeros((4, 4))
        X[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        X[:3, 3] = v
        ref = scipy.linalg.expm(X)[:3, :].reshape(12)
        assert np.allclose(out, ref, atol=1e-12)
    tw = np.array([0.1, -0.2, 0.3, 0.0, 0.0, 0.0])
    out = np.zeros(12)
    OB.lib().oracle_exp_twist(tw.ctypes.data, out.ctypes.data)
    assert np.allclose(out.reshape(3, 4), np.hstack([np.eye(3), tw[:3, None]]))


def test_twist_between_inverts_exp():
    rng = np.random.default_rng(5)
    for _ in range(20):
        tw = rng.normal(size=6)
        tw[3:] *= rng.uniform(0.01, 2.5) / max(np.linalg.norm(tw[3:]), 1e-9)
        a = np.zeros(12)
        tw0 = rng.normal(size=6)
        OB.lib().oracle_exp_twist(tw0.ctypes.data, a.ctypes.data)
        e = np.zeros(12)
        OB.lib().oracle_exp_twist(tw.ctypes.data, e.ctypes.data)
        A = np.vstack([a.reshape(3, 4), [0, 0, 0, 1]])
        E = np.vstack([e.reshape(3, 4), [0, 0, 0, 1]])
        b = np.ascontiguousarray((A @ E)[:3, :].reshape(12))
        out = np.zeros(6)
        OB.lib().oracle_twist_between(a.ctypes.data, b.ctypes.data, out.ctypes.data)
        assert np.allclose(out, tw, atol=1e-9)


def test_wrap_angle():
    w = OB.lib().oracle_wrap_angle
    assert w(0.5) == 0.5
    assert w(np.pi) == np.pi            # (-pi, pi]
    assert w(-np.pi) == pytest.approx(np.pi)
    assert w(3 * np.pi / 2) == pytest.approx(-np.pi / 2)
    assert w(-7.0) == pytest.approx(-7.0 + 2 * np.pi)


def test_truncated_normal_distribution():
    out = np.zeros(20000)
    OB.lib().oracle_truncated_normal(123, 0.5, len(out), out.ctypes.data)
    assert np.all(np.abs(out) <= 1.0)
    # N(0, 0.5) truncated to [-1, 1] (standardised [-2, 2]): variance = 0.25 * (1 - 2*2*phi(2)/(2*Phi(2)-1))
    phi2, Phi2 = np.exp(-2.0) / np.sqrt(2 * np.pi), 0.9772498680518208
    var = 0.25 * (1 - 4 * phi2 / (2 * Phi2 - 1))
    assert abs(out.mean()) < 0.02 and abs(out.var() - var) < 0.01
    z = np.zeros(4)
    OB.lib().oracle_truncated_normal(1, 0.0, 4, z.ctypes.data)
    assert np.all(z == 0.0)  # sigma == 0 returns the mean


def test_philox_known_answer():
    """Random123 known-answer vectors of philox4x32-10."""
    def run(ctr, key):
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        OB.lib().oracle_philox_raw(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        return [int(v) for v in o]
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_truncated_normal_is_a_pure_function():
    f = OB.lib().oracle_philox_truncated_normal
    a = f(42, 7, 3, 2, 1, 0.5)
    assert a == f(42, 7, 3, 2, 1, 0.5) and abs(a) <= 1.0
    assert a != f(42, 8, 3, 2, 1, 0.5) and a != f(43, 7, 3, 2, 1, 0.5)
    assert f(42, 7, 3, 2, 1, 0.0) == 0.0
