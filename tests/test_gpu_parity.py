"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs,
with the oracle's own truncated-normal draws injected."""
import numpy as np
import pytest

from fast_kinematic_simulator_b200 import capi, workloads as W

import parity
from oracle import oracle_binding as OB

pytestmark = pytest.mark.gpu


def _assert_parity(rep, sens, min_insensitive_frac=0.5, min_match_frac=0.999):
    """No particle without an enumerated reason may differ, most particles must BE without such a reason, and nearly all
    particles (sensitive or not) must reproduce: a green test means the trajectories were compared, not excused."""
    print(parity.describe(rep, sens))
    assert len(rep["bad_insensitive"]) == 0, "insensitive particles differ: %s" % rep["bad_insensitive"][:16]
    assert rep["n_insensitive"] >= min_insensitive_frac * rep["n"], (rep["n_insensitive"], rep["n"])
    assert rep["n_match"] >= min_match_frac * rep["n"], (rep["n_match"], rep["n"])
    if rep["n_match"] == rep["n"]:
        assert rep["n_desync"] == 0


def test_se2_arena_parity():
    w = W.se2_arena(128)
    rep, gpu, ref, sens = parity.run_parity(w, 128)
    _assert_parity(rep, sens, 0.25)
    assert gpu.did_contact.all() and rep["n_match"] == 128
    # every particle insensitive in the oracle must reproduce the 8 statistics too when nothing is enumerated
    if len(rep["bad_sensitive"]) == 0:
        assert rep["gpu_stats"] == rep["oracle_stats"]


def test_se2_no_contacts_allowed():
    w = W.se2_arena(64)
    rep, gpu, ref, sens = parity.run_parity(w, 64, allow_contacts=False)
    _assert_parity(rep, sens, 0.25)
    assert ((gpu.flags & capi.FLAG_ENDED_BY_NOCONTACT) != 0).any()


def test_se3_narrow_passage_parity():
    """BASELINE config 2 at its full size (16 384 particles) in injection mode.  The three translation columns of the SE(3)
    Jacobian [R | R (e x p)] have EQUAL norms up to round-off, so Eigen's first pivots are ties the reference breaks by the
    last bit: the decision tape carries the order."""
    w = W.se3_narrow_passage(16384)
    rep, gpu, ref, sens = parity.run_parity(w, 16384)
    _assert_parity(rep, sens, 0.8)
    assert gpu.did_contact.any() and not gpu.did_contact.all()


def test_arm_table_parity():
    """BASELINE config 3's contact regime, 4 096 particles: one distal link on the table -> stacked Jacobians of rank <= 6
    in 7 unknowns, whose last pivot is round-off.  With the oracle's decisions injected every
    particle must reproduce unless the oracle met an ill-conditioned solve on its way (condition estimate > 1e3: last-bit
    differences of the solve's INPUT come out multiplied by that) -- those are enumerated with that reason, and most of
    them reproduce anyway."""
    w = W.arm_table(4096)
    rep, gpu, ref, sens = parity.run_parity(w, 4096)
    _assert_parity(rep, sens, min_insensitive_frac=0.5, min_match_frac=0.97)
    assert gpu.did_contact.all() and gpu.resolve_failed.any()
    for i in rep["bad_sensitive"]:
        assert int(sens[i]) & OB.SENS_ILL_CONDITIONED, (i, int(sens[i]))  # every exception has the non-rank reason
    # the recorded decisions are what the test is about: round-off pivots cut AND kept must both occur
    d = rep["decisions"]
    assert ((d["flags"] & OB.DECISION_ROUNDOFF_PIVOT) != 0).sum() > 1000 and ((d["flags"] & OB.DECISION_OVERRIDE_SOLUTION) != 0).sum() > 50


def test_arm_table_parity_with_ill_conditioned_solves_injected():
    """The same batch with the solution of every solve whose condition estimate exceeds 100 taken from the tape as well
    (a third of the solves): little is left to amplify round-off, so every particle must reproduce -- all the way through
    kinematics, collision checks, normals, Jacobians, the well-conditioned solves, step scaling and failure handling:
    every flag and counter of every particle identical, every configuration within 1e-7, all but a handful (a chain of
    hundreds of condition-100 solves can still carry 1e-16 to a few 1e-9) within the 1e-9 of the north star."""
    w = W.arm_table(4096)
    rep, gpu, ref, sens = parity.run_parity(w, 4096, decision_cond_limit=100.0)
    print(parity.describe(rep, sens))
    assert rep["discrete_ok"].all() and rep["n_desync"] == 0
    assert rep["gpu_stats"] == rep["oracle_stats"]
    assert rep["n_match"] >= 0.999 * rep["n"], rep["n_match"]
    assert rep["cfg_err"].max() < 1e-7


def test_arm_table_aggregates_free_running():
    """No tapes: Philox noise on both sides, every decision the solver's own.  Individual particles diverge (the keep-or-cut
    of a round-off pivot is a coin flip in the reference itself), the DISTRIBUTIONS must agree: the device solver follows
    Eigen operation for operation, so it keeps such pivots at the reference's rate."""
    n = 4096
    w = W.arm_table(n)
    sim = w.make_simulator()
    g = sim.forward_simulate_robots(w.starts, w.targets, True, capi.NOISE_PHILOX)
    orc = parity.make_oracle(w)
    r = orc.forward_simulate(w.starts, w.targets, True, capi.NOISE_PHILOX)
    gf, rf = g.resolve_failed.mean(), ((r["flags"] & capi.FLAG_RESOLVE_FAILED) != 0).mean()
    se_f = np.sqrt(2 * 0.25 / n)
    print("failed %.4f / %.4f, microsteps %.2f / %.2f, iterations %.2f / %.2f, steps %.3f / %.3f" % (
        gf, rf, g.n_microsteps.mean(), r["n_microsteps"].mean(), g.n_resolver_iters.mean(), r["n_resolver_iters"].mean(),
        g.n_steps.mean(), r["n_steps"].mean()))
    assert abs(gf - rf) < 4 * se_f
    for a, b in ((g.n_microsteps, r["n_microsteps"]), (g.n_resolver_iters, r["n_resolver_iters"]), (g.n_steps, r["n_steps"])):
        se = np.sqrt((a.var() + b.var()) / n)
        assert abs(a.mean() - b.mean()) < 4 * se, (a.mean(), b.mean(), se)


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
    w = W.arm_elbow(2048)
    rep, gpu, ref, sens = parity.run_parity(w, 2048)
    _assert_parity(rep, sens)
    assert rep["n_match"] == 2048 and gpu.resolve_failed.any()
    assert rep["gpu_stats"] == rep["oracle_stats"]


def test_arm_selfcollision_parity():
    w = W.arm_selfcollision(512)
    rep, gpu, ref, sens = parity.run_parity(w, 512)
    _assert_parity(rep, sens)
    assert rep["n_match"] == 512
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
    w = W.gantry(1024)
    rep, gpu, ref, sens = parity.run_parity(w, 1024)
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
