"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs,
with the oracle's own truncated-normal draws injected."""
import numpy as np
import pytest

from fast_kinematic_simulator_b200 import capi, workloads as W

import parity

pytestmark = pytest.mark.gpu


def _assert_parity(rep, sens, min_insensitive_frac=0.0):
    print(parity.describe(rep, sens))
    assert len(rep["bad_insensitive"]) == 0, "insensitive particles differ: %s" % rep["bad_insensitive"][:16]
    assert rep["n_insensitive"] >= min_insensitive_frac * rep["n"]


def test_se2_arena_parity():
    w = W.se2_arena(128)
    rep, gpu, ref, sens = parity.run_parity(w, 128)
    _assert_parity(rep, sens, 0.25)
    assert gpu.did_contact.all()
    # every particle insensitive in the oracle must reproduce the 8 statistics too when nothing is enumerated
    if len(rep["bad_sensitive"]) == 0:
        assert rep["gpu_stats"] == rep["oracle_stats"]


def test_se2_no_contacts_allowed():
    w = W.se2_arena(64)
    rep, gpu, ref, sens = parity.run_parity(w, 64, allow_contacts=False)
    _assert_parity(rep, sens, 0.25)
    assert ((gpu.flags & capi.FLAG_ENDED_BY_NOCONTACT) != 0).any()


def test_se3_narrow_passage_parity():
    w = W.se3_narrow_passage(1024)
    rep, gpu, ref, sens = parity.run_parity(w, 1024)
    _assert_parity(rep, sens, 0.25)
    assert gpu.did_contact.any() and not gpu.did_contact.all()


def test_arm_table_parity():
    w = W.arm_table(512)
    rep, gpu, ref, sens = parity.run_parity(w, 512)
    _assert_parity(rep, sens)
    assert gpu.did_contact.any()


def test_arm_free_motion_parity():
    """No contact: a target above the table.  Every particle must match (nothing is near a threshold)."""
    w = W.arm_free(256)
    rep, gpu, ref, sens = parity.run_parity(w, 256)
    _assert_parity(rep, sens, 0.9)
    assert not gpu.did_contact.any()
    assert rep["n_match"] == 256


def test_arm_elbow_parity():
    """Contact on the proximal links: structurally-zero Jacobian columns, spherical-shoulder rank loss (rank 2 of 3),
    failing resolves.  Every particle must reproduce exactly."""
    w = W.arm_elbow(256)
    rep, gpu, ref, sens = parity.run_parity(w, 256)
    _assert_parity(rep, sens)
    assert rep["n_match"] == 256 and gpu.resolve_failed.any()
    assert rep["gpu_stats"] == rep["oracle_stats"]


def test_arm_selfcollision_parity():
    w = W.arm_selfcollision(128)
    rep, gpu, ref, sens = parity.run_parity(w, 128)
    _assert_parity(rep, sens)
    assert rep["gpu_stats"]["unsuccessful_self_collision_resolves"] == rep["oracle_stats"]["unsuccessful_self_collision_resolves"]


def test_failed_resolves_do_not_end_motion():
    sp = capi.default_solver_params()
    sp.failed_resolves_end_motion = 0
    sp.max_resolver_iterations = 3
    w = W.se3_narrow_passage(256)
    rep, gpu, ref, sens = parity.run_parity(w, 256, solver_params=sp)
    _assert_parity(rep, sens)


def test_empty_and_ragged_batches():
    w = W.se2_arena(8)
    sim = w.make_simulator()
    out = sim.forward_simulate_robots(np.zeros((0, 3)), w.targets, True, capi.NOISE_NONE)
    assert len(out) == 0
    # targets must be 1 or one per start (assert at spcs.hpp:790-793)
    with pytest.raises(capi.FksError):
        sim.forward_simulate_robots(w.starts[:4], np.repeat(w.targets, 3, axis=0), True, capi.NOISE_NONE)
    # injected mode without a tape is an argument error, not a crash
    with pytest.raises(capi.FksError):
        sim.forward_simulate_robots(w.starts[:4], w.targets, True, capi.NOISE_INJECTED, None)
    # one target per start == broadcasting the same target
    a = sim.forward_simulate_robots(w.starts[:5], w.targets, True, capi.NOISE_PHILOX)
    b = sim.forward_simulate_robots(w.starts[:5], np.repeat(w.targets, 5, axis=0), True, capi.NOISE_PHILOX)
    assert np.array_equal(a.records, b.records)


def test_reverse_is_forward_and_philox_is_partition_independent():
    w = W.se3_narrow_passage(96)
    sim = w.make_simulator()
    full = sim.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    rev = sim.reverse_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    assert np.array_equal(full.records, rev.records)
    # sharding the batch (as the multi-GPU path does) must not change any particle: noise is keyed by global id
    lo = sim.forward_simulate_robots(w.starts[:40], w.targets, True, capi.NOISE_PHILOX, first_particle_id=0)
    hi = sim.forward_simulate_robots(w.starts[40:], w.targets, True, capi.NOISE_PHILOX, first_particle_id=40)
    assert np.array_equal(np.concatenate([lo.records, hi.records]), full.records)


def test_statistics_accumulate_and_reset():
    w = W.se2_arena(32)
    sim = w.make_simulator()
    r = sim.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    st = sim.get_statistics()
    assert st["total_microsteps"] == int(r.n_microsteps.sum())
    assert st["total_resolver_iterations"] == int(r.n_resolver_iters.sum())
    assert st["successful_resolves"] + st["unsuccessful_resolves"] == int(r.n_steps.sum())
    sim.reset_statistics()
    assert all(v == 0 for v in sim.get_statistics().values())


def test_philox_matches_oracle_statistically():
    """Philox mode: device libm differs from glibc in the last ulp, so compare with a tolerance on a free-space run."""
    w = W.arm_free(64)
    sim = w.make_simulator()
    gpu = sim.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    orc = parity.make_oracle(w)
    ref = orc.forward_simulate(w.starts, w.targets, True, capi.NOISE_PHILOX)
    assert np.array_equal(gpu.n_microsteps, ref["n_microsteps"])
    assert np.max(np.abs(gpu.configs - ref["cfg"])) < 1e-9


def test_gantry_parity_all_joint_types():
    """PRISMATIC + FIXED + REVOLUTE (limit reached) + CONTINUOUS (wraps through pi) joints, wall contact."""
    w = W.gantry(128)
    rep, gpu, ref, sens = parity.run_parity(w, 128)
    _assert_parity(rep, sens, 0.5)
    assert gpu.did_contact.mean() > 0.9 and np.array_equal(gpu.did_contact, (ref["flags"] & 1) != 0)
    # the y axis target lies beyond the prismatic limit: the joint saturates at +0.2
    assert np.all(gpu.configs[:, 1] <= 0.2 + 1e-12) and np.all(np.abs(gpu.configs[:, 3]) <= np.pi + 1e-12)


@pytest.mark.parametrize("name,n,dist", [("se2_arena", 64, 1.5), ("se3_narrow_passage", 128, 0.5), ("arm_free", 64, 0.35), ("gantry", 64, 0.8)])
def test_shortcut_distance_parity(name, n, dist):
    """simulation_shortcut_distance > 0: ComputeConfigurationDistanceTo runs on the device after every step (spcs:898-902)."""
    sp = capi.default_solver_params()
    sp.simulation_shortcut_distance = dist
    w = W.make(name, n_particles=n)
    rep, gpu, ref, sens = parity.run_parity(w, n, solver_params=sp)
    _assert_parity(rep, sens)
    assert ((gpu.flags & capi.FLAG_ENDED_BY_SHORTCUT) != 0).any()
    assert np.array_equal(gpu.n_steps[sens == 0], ref["n_steps"][sens == 0])


def test_recovered_resolves_counter():
    sp = capi.default_solver_params()
    sp.failed_resolves_end_motion = 0
    sp.max_resolver_iterations = 1
    w = W.se3_narrow_passage(256)
    rep, gpu, ref, sens = parity.run_parity(w, 256, solver_params=sp)
    _assert_parity(rep, sens)
    assert rep["oracle_stats"]["recovered_unsuccessful_resolves"] > 0
    if len(rep["bad_sensitive"]) == 0:
        assert rep["gpu_stats"] == rep["oracle_stats"]
