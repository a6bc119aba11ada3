"""Obstacle sets for the environment-builder tests (CPU: oracle vs host builder, GPU: device builder vs oracle) and the
comparison they share.  Obstacles are (pose12, half_extents3, object_id) -- OBSTACLE_CONFIG
(simulator_environment_builder.hpp:25-49)."""
import numpy as np

from fast_kinematic_simulator_b200 import workloads as W
from fast_kinematic_simulator_b200.simulator import IDENTITY12, make_transform


def _rotation(rotvec):
    rotvec = np.asarray(rotvec, dtype=np.float64)
    angle = np.linalg.norm(rotvec)
    if angle == 0.0:
        return np.eye(3)
    a = rotvec / angle
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def se2_arena():
    w = W.se2_arena(1)
    return w.obstacles, w.resolution


def narrow_passage():
    w = W.se3_narrow_passage(1)
    return w.obstacles, w.resolution


def rotated_boxes():
    """overlapping, arbitrarily rotated boxes: later obstacles rewrite the surface cells of earlier ones"""
    rng = np.random.Generator(np.random.MT19937(7))
    obstacles = []
    for i in range(7):
        c = rng.uniform(-0.8, 0.8, 3)
        ext = rng.uniform(0.12, 0.5, 3)
        obstacles.append((make_transform(c, _rotation(rng.normal(0.0, 0.8, 3))), tuple(ext), i + 1))
    return obstacles, 0.07


def thin_plates():
    """obstacles with a single sample along an axis (index == 0 and index == count - 1 at once), with two samples,
    one with no sample at all, and a duplicate obstacle"""
    res = 0.1
    obstacles = [
        (make_transform((0.0, 0.0, 0.0)), (0.5, 0.4, 0.03), 1),    # 1 sample in z
        (make_transform((0.3, 0.1, 0.4), _rotation((0.3, -0.2, 0.5))), (0.03, 0.03, 0.45), 2),  # 1 x 1 x 18 needle
        (make_transform((-0.4, 0.5, 0.6)), (0.05, 0.05, 0.05), 3),  # 2 x 2 x 2 samples: one cell
        (make_transform((0.9, 0.9, 0.9)), (0.02, 0.3, 0.3), 4),     # no sample along x: contributes nothing
        (make_transform((0.0, 0.0, 0.0)), (0.5, 0.4, 0.03), 5),     # same as the first
        (make_transform((0.2, -0.7, 0.2), _rotation((0.0, 0.0, np.pi / 4))), (0.3, 0.3, 0.1), 6),
    ]
    return obstacles, res


def no_obstacles():
    return [], 0.5  # default 10 x 10 x 10 m grid (envb.cpp:51-65)


def solid_cube():
    """deep interior (pass-1 gradient entries survive) and grid-face cells inside nothing"""
    return [(IDENTITY12, (0.6, 0.6, 0.6), 1)], 0.1


CASES = {
    "se2_arena": se2_arena,
    "narrow_passage": narrow_passage,
    "rotated_boxes": rotated_boxes,
    "thin_plates": thin_plates,
    "no_obstacles": no_obstacles,
    "solid_cube": solid_cube,
}


def assert_same_environment(ref, built, who):
    """ref: dict from oracle_binding.build_environment; built: BuiltEnvironment (host builder or fks_env_download).
    Everything is compared exactly: cell counts, origin, occupancy, float SDF, surface-normal cells and entries."""
    assert tuple(ref["shape"]) == tuple(built.shape), who
    assert np.array_equal(ref["origin"], built.origin), who
    assert np.array_equal(ref["inverse_origin"], built.inverse_origin), who
    assert ref["resolution"] == built.resolution, who
    if built.occupancy is not None:
        assert np.array_equal(ref["occupancy"], built.occupancy), who + ": occupancy"
    bad = np.flatnonzero(ref["sdf"].reshape(-1) != built.sdf.reshape(-1))
    assert bad.size == 0, "%s: %d SDF cells differ, first %s" % (who, bad.size, bad[:5])
    assert np.array_equal(ref["normal_cell_index"], built.normal_cell_index), who + ": normal cells"
    assert np.array_equal(ref["normal_cell_start"], built.normal_cell_start), who + ": normal offsets"
    assert np.array_equal(ref["normal_entries"], built.normal_entries), who + ": normal entries"
