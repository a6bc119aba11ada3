"""Device environment builder (fks_env_build_device): the CUDA rasteriser, distance transform, SDF and
surface-normal kernels against the oracle's restatement of BuildCompleteEnvironment (simulator_environment_builder.cpp:470-476)
-- every array compared exactly -- and, at the full 512^3 size of BASELINE config 4, against the host builder plus
size-independent properties of a distance field."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle_binding as OB  # noqa: E402
from fast_kinematic_simulator_b200 import capi, simulator as S, workloads as W  # noqa: E402

from env_cases import CASES, assert_same_environment  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_builder_matches_oracle(name):
    obstacles, res = CASES[name]()
    ref = OB.build_environment(obstacles, res)
    env = S.build_complete_environment_on_device(obstacles, res)
    got = env.download()
    assert got.occupancy is not None
    assert_same_environment(ref, got, "device builder")
    assert env.build_timings_ms["total"] > 0.0


def test_device_builder_matches_oracle_on_the_arm_room():
    """BASELINE config 3 environment: 5 x 5 x 2.5 m room + table at res 0.04 (1.2 M cells, 170 k surface cells)."""
    w = W.arm_table(4)
    ref = OB.build_environment(w.obstacles, w.resolution)
    got = S.build_complete_environment_on_device(w.obstacles, w.resolution).download()
    assert_same_environment(ref, got, "device builder")


def test_uploaded_environment_downloads_unchanged():
    """fks_env_download of an environment made by fks_env_create returns what was uploaded (no occupancy)."""
    w = W.se3_narrow_passage(4)
    host = w.environment()
    back = S.GpuEnvironment(host).download()
    assert back.occupancy is None
    assert np.array_equal(host.sdf, back.sdf)
    assert np.array_equal(host.normal_cell_index, back.normal_cell_index)
    assert np.array_equal(host.normal_cell_start, back.normal_cell_start)
    assert np.array_equal(host.normal_entries, back.normal_entries)


@pytest.mark.parametrize("name,n", [("se2_arena", 128), ("se3_narrow_passage", 512), ("arm_table", 256), ("arm_elbow", 128)])
def test_simulation_in_device_built_environment_is_identical(name, n):
    """The simulate kernels read the device-built SDF / normal hash: same records, bit for bit, as with the uploaded one."""
    w = W.make(name, n_particles=n)
    a = w.make_simulator().forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    sim = w.make_simulator(build_on_device=True)
    b = sim.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    assert a.did_contact.any() or name == "arm_elbow"
    assert a.records.tobytes() == b.records.tobytes()
    cfg = np.repeat(w.starts[:1], 64, axis=0)
    assert np.array_equal(w.make_simulator().check_config_collision(cfg, 0.5), sim.check_config_collision(cfg, 0.5))


def test_full_size_512_cubed_build_matches_host_builder_and_is_a_distance_field():
    """BASELINE config 4: ~512^3 cells, 264 cuboids.  The oracle's whole-line transform is too slow here; the host builder
    (checked against the oracle on the small cases, tests/test_oracle_env_builder.py) stands in, plus properties."""
    w = W.se3_highres(n_particles=4)
    env = S.build_complete_environment_on_device(w.obstacles, w.resolution)
    got = env.download()
    host = w.environment()
    assert got.shape == host.shape and min(got.shape) >= 500
    assert np.array_equal(host.occupancy, got.occupancy)
    assert np.array_equal(host.sdf, got.sdf)
    assert np.array_equal(host.normal_cell_index, got.normal_cell_index)
    assert np.array_equal(host.normal_cell_start, got.normal_cell_start)
    assert np.array_equal(host.normal_entries, got.normal_entries)
    # properties: sign = occupancy; neighbouring cells differ by at most one cell inside a sign region, two across
    sdf, occ, res = got.sdf, got.occupancy.astype(bool), got.resolution
    assert np.array_equal(sdf < 0, occ)
    assert np.abs(sdf[occ]).min() >= np.float32(res) and np.abs(sdf[~occ]).min() >= np.float32(res)
    for axis in range(3):
        d = np.abs(np.diff(sdf, axis=axis))
        assert d.max() <= 2.0 * res * (1 + 1e-4)
    # squared distances in cells are sums of three squares: (sdf / res)^2 rounds to an integer
    # (float rounding of the stored value: relative 6e-8 on the distance, 1.2e-7 on its square)
    s2 = (sdf.reshape(-1)[:: 997].astype(np.float64) / res) ** 2
    assert (np.abs(s2 - np.round(s2)) <= 4e-7 * s2 + 1e-9).all()
    print("device build of %s: %s ms" % (got.shape, {k: round(v, 2) for k, v in env.build_timings_ms.items()}))
