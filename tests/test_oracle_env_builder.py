"""Environment builder on the CPU: the oracle's literal restatement of BuildCompleteEnvironment
(simulator_environment_builder.cpp:470-476) against independent implementations -- scipy's exact Euclidean distance
transform, closed-form box geometry -- and against the product's host builder (fks_build_environment), which uses a
different distance-transform algorithm and a folded form of the 26-way surface chain."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle_binding as OB  # noqa: E402
from fast_kinematic_simulator_b200 import simulator as S, workloads as W  # noqa: E402

from env_cases import CASES, assert_same_environment  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_builder_matches_host_builder(name):
    obstacles, res = CASES[name]()
    ref = OB.build_environment(obstacles, res)
    host = S.build_complete_environment(obstacles, res)
    assert_same_environment(ref, host, "host builder")


@pytest.mark.parametrize("name", ["se2_arena", "rotated_boxes", "thin_plates"])
def test_oracle_sdf_matches_scipy_edt(name):
    from scipy import ndimage

    obstacles, res = CASES[name]()
    ref = OB.build_environment(obstacles, res)
    occ = ref["occupancy"].astype(bool)
    assert occ.any() and (~occ).any()
    to_filled = ndimage.distance_transform_edt(~occ)  # distance of free cells to the nearest filled cell (in cells)
    to_free = ndimage.distance_transform_edt(occ)
    expect = (to_filled * res - to_free * res).astype(np.float32)
    assert np.array_equal(expect, ref["sdf"])


def test_oracle_axis_aligned_box_geometry():
    """One axis-aligned box, placed off the cell boundaries by a small anchor obstacle that fixes the grid origin: occupancy
    is the box, the SDF outside is the distance to it, and a surface cell carries the normals of the LAST surface sample that
    falls into it in the reference's x/y/z loop order (two samples per cell and axis; envb.cpp:280-463, :162-187)."""
    res = 0.1
    anchor = (S.make_transform((-1.0, -1.0, -1.0)), (0.03, 0.03, 0.03), 1)  # one sample at -0.98 -> grid origin at -1.33
    box = (S.make_transform((0.033, 0.037, 0.041)), (0.5, 0.3, 0.2), 2)
    ref = OB.build_environment([anchor, box], res)
    nx, ny, nz = ref["shape"]
    occ = ref["occupancy"].copy()
    assert occ[3, 3, 3] == 1  # the anchor's single cell
    occ[3, 3, 3] = 0
    # box samples at (9.13 + i/2, 11.17 + j/2, 12.21 + k/2) cells, i < 20, j < 12, k < 8
    assert occ.sum() == 10 * 6 * 4 and occ[9:19, 11:17, 12:16].all()
    sdf = ref["sdf"]
    assert sdf[9, 13, 13] == np.float32(-res) and sdf[8, 13, 13] == np.float32(res) and sdf[6, 13, 13] == np.float32(3 * res)
    idx = ref["normal_cell_index"]
    start = ref["normal_cell_start"]
    ent = ref["normal_entries"]

    def entries_of(x, y, z):
        li = (x * ny + y) * nz + z
        k = np.searchsorted(idx, li)
        assert idx[k] == li
        return ent[start[k]:start[k + 1]]

    # maximum corner: the last sample of the cell is the corner sample (19, 11, 7) -> +X, +Y, +Z with opposite entry directions
    corner = entries_of(18, 16, 15)
    assert np.array_equal(corner[:, 4:], np.eye(3)) and np.array_equal(corner[:, :3], -np.eye(3)) and not corner[:, 3].any()
    # minimum corner: samples (1, 1, *) are interior, so the last surface sample of the cell is (1, 1, 0): the -Z face only
    assert np.array_equal(entries_of(9, 11, 12)[:, 4:], np.array([[0, 0, -1.0]]))
    edge = entries_of(18, 16, 13)
    assert np.array_equal(edge[:, 4:], np.array([[1.0, 0, 0], [0, 1.0, 0]]))
    face = entries_of(13, 13, 15)
    assert np.array_equal(face[:, 4:], np.array([[0, 0, 1.0]])) and np.array_equal(face[:, :3], np.array([[0, 0, -1.0]]))
    assert np.array_equal(entries_of(9, 13, 13)[:, 4:], np.array([[-1.0, 0, 0]]))
    # every filled cell has an entry (pass 1 at least), no free cell has one
    assert len(idx) == ref["occupancy"].sum()


def test_oracle_deep_interior_keeps_gradient_entry():
    res = 0.1
    ref = OB.build_environment([(S.IDENTITY12, [0.6, 0.6, 0.6], 1)], res)
    nx, ny, nz = ref["shape"]
    li = (9 * ny + 9) * nz + 8  # centre region: sdf < -1.5 res
    assert ref["sdf"].reshape(-1)[li] < -1.5 * res
    k = np.searchsorted(ref["normal_cell_index"], li)
    assert ref["normal_cell_index"][k] == li
    e = ref["normal_entries"][ref["normal_cell_start"][k]:ref["normal_cell_start"][k + 1]]
    assert e.shape == (1, 7) and not e[0, :4].any()
    assert abs(np.linalg.norm(e[0, 4:]) - 1.0) < 1e-12 or not e[0, 4:].any()


def test_workload_environments_match_oracle_builder():
    """The environments the parity and bench workloads run in (configs 1-3) are the ones the oracle's builder makes."""
    for w in (W.se2_arena(4), W.se3_narrow_passage(4)):
        ref = OB.build_environment(w.obstacles, w.resolution)
        assert_same_environment(ref, w.environment(), w.name)
