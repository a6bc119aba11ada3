"""The device's contact solver (fks_debug_qr_solve = the colpiv_qr_lanes the simulate kernel calls) against the oracle's
restatement of Eigen's ColPivHouseholderQR solve (spcs.hpp:1990-1998) ON THE SAME SYSTEMS: the stacked Jacobians the
oracle met while simulating the BASELINE workloads, plus random and degenerate ones.  The device follows Eigen operation
for operation with unfused arithmetic, so the solutions must be BIT-identical -- including the round-off coin flips of
rank-deficient systems (a kept round-off pivot gives the same garbage on both sides)."""
import numpy as np
import pytest

from fast_kinematic_simulator_b200 import capi, simulator as S, workloads as W
from oracle import oracle_binding as OB

import parity

pytestmark = pytest.mark.gpu


def _captured(name, n):
    w = W.make(name, n_particles=n)
    orc = parity.make_oracle(w)
    OB.lib().oracle_debug_capture_systems(1 << 20)
    OB.run_with_tape(orc, w.starts, w.targets, True)
    systems = OB.captured_systems()
    OB.lib().oracle_debug_capture_systems(0)
    return systems


@pytest.mark.parametrize("name,n", [("arm_table", 512), ("se3_narrow_passage", 1024), ("se2_arena", 128), ("arm_selfcollision", 64),
                                    ("gantry", 128)])
def test_device_solver_returns_the_oracles_bits(name, n):
    systems = _captured(name, n)
    assert len(systems) > 500
    x, flags = S.debug_qr_solve(systems)
    ref = np.array([OB.qr_solve_info(A, b)[0] for A, b in systems])
    info = np.array([OB.qr_solve_info(A, b)[3] for A, b in systems])
    same = np.all(x.view(np.uint64) == ref.view(np.uint64), axis=1) | np.all(x == ref, axis=1)  # (-0.0 == 0.0)
    print("%s: %d systems, rows up to %d, %d with a round-off pivot (%d kept), bit-identical %d" % (
        name, len(systems), max(A.shape[0] for A, _ in systems), int((info & 1).sum()), int(((info & 2) != 0).sum()), int(same.sum())))
    assert same.all(), np.nonzero(~same)[0][:10]
    if name == "arm_table":
        assert ((info & 2) != 0).sum() > 20  # kept round-off pivots are in the sample, and reproduce


def test_device_solver_on_random_and_degenerate_systems():
    rng = np.random.default_rng(11)
    for cols in (3, 6, 7, 4):
        systems = []
        for rows in (1, 2, 3, 5, 6, 7, 8, 9, 24, 31, 32, 33, 63, 64, 65, 200, 1152):
            A = rng.normal(size=(rows, cols))
            systems.append((A, rng.normal(size=rows)))
            A2 = A.copy()
            A2[:, -1] = A2[:, 0] * 2.0  # exactly dependent columns
            systems.append((A2, rng.normal(size=rows)))
            A3 = A.copy()
            A3[:, 1] = 0.0  # a structurally zero column (a joint that does not move the point)
            systems.append((A3, rng.normal(size=rows)))
            systems.append((np.zeros((rows, cols)), rng.normal(size=rows)))
        x, flags = S.debug_qr_solve(systems)
        ref = np.array([OB.qr_solve_info(A, b)[0] for A, b in systems])
        assert np.all((x.view(np.uint64) == ref.view(np.uint64)) | (x == ref))
