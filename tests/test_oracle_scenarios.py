"""Analytic scenarios and invariants of the oracle (SURVEY.md 4 ii-iii) + the environment builder."""
import ctypes as C

import numpy as np

from fast_kinematic_simulator_b200 import capi, workloads as W
from fast_kinematic_simulator_b200.simulator import RobotDescription, build_complete_environment, make_transform
from oracle import oracle_binding as OB

import parity


def point_robot_se2(vlim=1.0, noise=0.0):
    ax = dict(kp=1.0, velocity_limit=vlim, proportional_noise=noise, minimum_noise=noise)
    axr = dict(kp=1.0, velocity_limit=0.5, proportional_noise=noise, minimum_noise=noise)
    return RobotDescription(capi.ROBOT_SE2, [[0.0, 0.0, 0.0]], [0], [ax, ax, axr])


def wall_env(res=0.125):
    # one wall: x in [1, 1.5], big in y, centred slightly off z = 0
    return build_complete_environment([(make_transform((1.25, 0.0, 0.05)), (0.25, 2.0, 0.5), 1)], res)


def test_free_space_se2_is_the_clamped_pid_recursion():
    env = wall_env()
    rob = point_robot_se2()
    orc = OB.OracleSimulator(env.desc, rob.to_c(), capi.default_solver_params(), 25.0, 42, 1)
    start = np.array([[-1.5, 0.3, 0.2]])
    target = np.array([[-0.2, -3.0, 1.2]])
    rec = orc.forward_simulate(start, target, True, capi.NOISE_NONE)
    q = start[0].copy()
    for _ in range(25):  # kp = 1, ki = kd = 0: action = clamp(err, +-vlim); displacement = action * dt in n equal microsteps
        err = target[0] - q
        act = np.clip(err, [-1.0, -1.0, -0.5], [1.0, 1.0, 0.5])
        q = q + act * 0.04
    assert np.allclose(rec["cfg"][0], q, atol=1e-12)
    assert rec["flags"][0] == 0 and rec["n_steps"][0] == 25


def test_point_pushed_into_wall_stops_at_the_wall():
    env = wall_env()
    rob = point_robot_se2()
    orc = OB.OracleSimulator(env.desc, rob.to_c(), capi.default_solver_params(), 25.0, 42, 1)
    rec = orc.forward_simulate(np.array([[0.6, 0.0, 0.0]]), np.array([[1.4, 0.0, 0.0]]), True, capi.NOISE_NONE)
    assert rec["flags"][0] & capi.FLAG_DID_CONTACT
    x = rec["cfg"][0][0]
    assert 1.0 - 0.125 <= x <= 1.0 + 0.01, x  # within one voxel of the wall face at x = 1
    assert abs(rec["cfg"][0][1]) < 1e-9


def test_no_contact_mode_returns_last_free_configuration():
    env = wall_env()
    rob = point_robot_se2()
    orc = OB.OracleSimulator(env.desc, rob.to_c(), capi.default_solver_params(), 25.0, 42, 1)
    rec = orc.forward_simulate(np.array([[0.6, 0.0, 0.0]]), np.array([[1.4, 0.0, 0.0]]), False, capi.NOISE_NONE)
    assert rec["flags"][0] & capi.FLAG_ENDED_BY_NOCONTACT
    assert not (rec["flags"][0] & capi.FLAG_DID_CONTACT)  # the reference reports did_contact = false here (spcs.hpp:877-880,904-909)
    assert rec["cfg"][0][0] < 1.0 and rec["n_steps"][0] < 25


def test_invariants_on_the_se3_workload():
    w = W.se3_narrow_passage(64)
    orc = parity.make_oracle(w)
    rec, tape, sens = OB.run_with_tape(orc, w.starts, w.targets)
    st = orc.statistics()
    assert not np.any(rec["flags"] & capi.FLAG_WOULD_ASSERT_MICROSTEP)  # microstep motion <= one cell (spcs.hpp:1570-1575)
    assert st[8] == rec["n_microsteps"].sum() and st[9] == rec["n_resolver_iters"].sum()
    assert st[0] + st[1] == rec["n_steps"].sum()          # every step is a successful or an unsuccessful resolve
    assert st[0] == st[2] + st[3]                          # successful = free + collision
    assert len(tape[0]) == 6 * rec["n_microsteps"].sum()   # one draw per axis per microstep
    # rotation part of every final pose is still orthonormal
    R = rec["cfg"].reshape(-1, 3, 4)[:, :, :3]
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-9)
    # replaying the recorded tape reproduces the run bit for bit, with any thread count
    orc1 = parity.make_oracle(w, num_threads=1)
    rec2 = orc1.forward_simulate(w.starts, w.targets, True, capi.NOISE_INJECTED, tape)
    assert np.array_equal(rec2, rec)


def test_failed_resolve_returns_previous_configuration():
    sp = capi.default_solver_params()
    sp.max_resolver_iterations = 0  # the first resolver iteration already exceeds the budget (iters > max, spcs.hpp:1705)
    w = W.se2_arena(8)
    orc = parity.make_oracle(w, solver_params=sp)
    rec = orc.forward_simulate(w.starts, w.targets, True, capi.NOISE_NONE)
    assert np.all(rec["flags"] & capi.FLAG_ENDED_BY_FAILURE)
    assert np.all(rec["n_resolver_iters"] == 1)
    # one more microstep from the returned configuration would collide again: the returned state itself is free
    orc2 = parity.make_oracle(w, solver_params=capi.default_solver_params())
    again = orc2.forward_simulate(rec["cfg"], rec["cfg"], True, capi.NOISE_NONE)
    assert not np.any(again["flags"] & capi.FLAG_DID_CONTACT)


def test_environment_builder_shapes_and_signs():
    w = W.arm_table(1)
    env = w.environment()
    assert env.shape == (131, 131, 71)
    w2 = W.se3_narrow_passage(1)
    assert w2.environment().shape == (93, 46, 46)
    env = wall_env()
    sdf, occ = env.sdf, env.occupancy
    res = env.resolution
    assert np.all(sdf[occ == 1] <= -res + 1e-6) and np.all(sdf[occ == 0] >= res - 1e-6)
    # exact Euclidean distances: compare against brute force on a sample of cells
    filled = np.argwhere(occ == 1)
    free = np.argwhere(occ == 0)
    rng = np.random.default_rng(0)
    for idx in rng.integers(0, sdf.size, 40):
        c = np.array(np.unravel_index(idx, sdf.shape))
        d1 = np.sqrt(((filled - c) ** 2).sum(1).min()) * res
        d2 = np.sqrt(((free - c) ** 2).sum(1).min()) * res
        assert abs(sdf[tuple(c)] - np.float32(d1 - d2)) <= 1e-6
    # grid = discretised bounding box + 3 cells (envb.cpp:130-142).  The discretisation is lopsided (locations run from
    # -ext + res/2 to +ext, envb.cpp:29-31) and only the low side gets the extra half cell, so the high side keeps 2 free slices
    assert occ[:3].sum() == 0 and occ[3].sum() > 0 and occ[-2:].sum() == 0 and occ[-3].sum() > 0


def test_surface_normals_of_a_box():
    env = wall_env()
    # a cell on the -x face of the wall (x = 1.0 .. 1.125), mid-height: stored normal -x with entry direction +x
    p = np.array([1.03, 0.2, 0.05])
    raw = C.c_float()
    est = C.c_double()
    grad = np.zeros(3)
    inb = OB.lib().oracle_env_query(C.addressof(env.desc), p.ctypes.data, C.addressof(raw), C.addressof(est), grad.ctypes.data)
    assert inb == 1 and raw.value < 0 and grad[0] < 0 and abs(grad[1]) < 1e-6
    g = (env.inverse_origin.reshape(3, 4) @ np.append(p, 1.0)) / env.resolution
    li = (int(g[0]) * env.shape[1] + int(g[1])) * env.shape[2] + int(g[2])
    k = np.searchsorted(env.normal_cell_index, li)
    assert env.normal_cell_index[k] == li
    ent = env.normal_entries[env.normal_cell_start[k]:env.normal_cell_start[k + 1]]
    assert len(ent) == 1
    assert np.allclose(ent[0], [1, 0, 0, 0, -1, 0, 0])
    # estimated distance of a point 0.03 inside the face: about -0.03 (cell-centre distance +- half a cell + projection)
    assert -0.1 < est.value < 0.0


def test_linked_robot_kinematics_against_numpy():
    rob = W.arm_robot()
    d = rob.to_c()
    rng = np.random.default_rng(3)
    for _ in range(5):
        q = rng.uniform(-1.5, 1.5, 7)
        T = np.zeros(8 * 12)
        Jm = np.zeros(21)
        cfg = np.zeros(7)
        p = np.array([0.04, 0.0, 0.27])
        assert OB.lib().oracle_robot_kinematics(C.addressof(d), q.ctypes.data, 7, p.ctypes.data, T.ctypes.data, Jm.ctypes.data, cfg.ctypes.data) == 0
        tip = T.reshape(8, 3, 4)[7] @ np.array([0, 0, W.ARM_LINK_LENGTH, 1.0])
        assert np.allclose(tip, W.arm_fk_tip(q), atol=1e-12)
        # numerical Jacobian of the point
        def pos(qq):
            TT = np.zeros(8 * 12)
            OB.lib().oracle_robot_kinematics(C.addressof(d), np.ascontiguousarray(qq).ctypes.data, 7, p.ctypes.data, TT.ctypes.data,
                                             np.zeros(21).ctypes.data, np.zeros(7).ctypes.data)
            return TT.reshape(8, 3, 4)[7] @ np.append(p, 1.0)
        num = np.stack([(pos(q + 1e-6 * np.eye(7)[j]) - pos(q - 1e-6 * np.eye(7)[j])) / 2e-6 for j in range(7)], axis=1)
        assert np.allclose(Jm.reshape(3, 7), num, atol=1e-8)
