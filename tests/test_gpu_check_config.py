"""CheckConfigCollision batch (spcs.hpp:1398-1416): GPU against the oracle on random configurations."""
import numpy as np
import pytest

from fast_kinematic_simulator_b200 import capi, workloads as W

import parity

pytestmark = pytest.mark.gpu


def _configs(w, n, rng):
    if w.kind == capi.ROBOT_SE2:
        return np.column_stack([rng.uniform(-4.5, 4.5, n), rng.uniform(-4.5, 4.5, n), rng.uniform(-3.1, 3.1, n)])
    if w.kind == capi.ROBOT_SE3:
        return np.stack([W._se3_config(np.array([rng.uniform(-0.6, 0.6), rng.uniform(-1.6, 1.2), rng.uniform(-0.5, 0.5)]),
                                       rng.normal(0.0, 0.6, 3)) for _ in range(n)])
    return rng.uniform(-2.8, 2.8, (n, w.robot.n_dof))


@pytest.mark.parametrize("name", ["se2_arena", "se3_narrow_passage", "arm_table"])
@pytest.mark.parametrize("inflation", [0.0, 1.5])
def test_check_config_collision_matches_oracle(name, inflation):
    w = W.make(name, n_particles=4)
    rng = np.random.default_rng(11)
    cfg = _configs(w, 2000, rng)
    sim = w.make_simulator()
    got = sim.check_config_collision(cfg, inflation)
    ref = parity.make_oracle(w).check_config_collision(cfg, inflation)
    # configurations within rounding of a threshold may differ: none is expected among 2000 random ones
    assert np.array_equal(got, ref), np.nonzero(got != ref)[0][:10]
    assert 0.02 < got.mean() < 0.999  # the sample mixes free and colliding configurations
    if name == "arm_table":
        # folded arms: self collisions only (lift the arm away from the obstacles)
        q = np.tile(np.array([0.0, 0.3, 0.0, 2.6, 0.0, 2.6, 0.0]), (64, 1)) + rng.normal(0.0, 0.05, (64, 7))
        g2, r2 = sim.check_config_collision(q, inflation), parity.make_oracle(w).check_config_collision(q, inflation)
        assert np.array_equal(g2, r2) and g2.any()


def test_check_config_collision_edge_cases():
    w = W.se2_arena(4)
    sim = w.make_simulator()
    assert sim.check_config_collision(np.zeros((0, 3))).shape == (0,)
    far = np.array([[100.0, 100.0, 0.0]])  # outside the grid: out-of-bounds points are free (spcs.hpp:943-955)
    assert not sim.check_config_collision(far)[0]
    inside = np.array([[1.0, 0.0, 0.0]])   # centre of the interior box
    assert sim.check_config_collision(inside)[0]
