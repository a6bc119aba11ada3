"""Generates the committed golden fixtures.  Run in the build container (needs /root/reference for the PID
vectors, which come from the REFERENCE's own header through oracle/_ref/pid_ref):

    python tests/golden/make_golden.py

* pid_golden.json        -- outputs of the reference's SimplePIDController (bit-exact pin of the PID restatement)
* actuator_golden.json   -- outputs of the reference's TruncatedNormalUncertainVelocityActuator with injected draws
                            (simple_uncertainty_models.hpp through oracle/_ref/unc_ref: bit-exact pin of the actuator restatement)
* env_golden.npz         -- BuildCompleteEnvironment of two small obstacle sets (tests/env_cases.py: thin_plates, rotated_boxes)
                            by the oracle: occupancy, float SDF, surface-normal table (host and device builder are held to it)
* trace_golden.npz       -- step traces (ForwardSimulationStepTrace, flat) of single particles by the oracle, Philox noise
* forward_golden.npz     -- oracle end states for small seeded batches of every robot kind (and of BASELINE config 3's
                            contact regime, arm_table) with the recorded noise and decision tapes: the GPU tests replay the
                            tapes and compare, so a GPU box without the oracle
                            build (or a future oracle change) is still pinned to these numbers.  The reference
                            ships no fixtures of its own: these are the builder's.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fast_kinematic_simulator_b200 import capi, workloads as W  # noqa: E402
from oracle import oracle_binding as OB  # noqa: E402


def actuator():
    rng = np.random.default_rng(2025)
    cases = []
    for _ in range(24):
        params = [float(rng.uniform(-2.0, 2.0)), 0.3, float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-0.1, 0.1)),
                  float(rng.choice([0.5, -0.5, 0.0, 1.7, 0.25]))]
        controls = [float(x) for x in rng.normal(0.0, 1.5, 28)] + [0.0, params[0], -params[0], 10.0 * params[0]]
        draws = [float(x) for x in rng.uniform(-1.0, 1.0, 28)] + [1.0, -1.0, 0.0, 0.5]
        quiet, noisy, dist = OB.actuator_reference(*params, controls, draws)
        cases.append(dict(params=params, controls=controls, draws=draws, quiet=[float(x) for x in quiet],
                          noisy=[float(x) for x in noisy], distribution=list(dist)))
    json.dump(dict(source="reference simple_uncertainty_models.hpp via oracle/_ref/unc_ref (arc_utilities stand-in: oracle/shim)", cases=cases),
              open(os.path.join(HERE, "actuator_golden.json"), "w"))


def pid():
    rng = np.random.default_rng(2024)
    cases = []
    for _ in range(8):
        gains = [float(x) for x in rng.uniform(-2, 2, 4)]
        errs = [float(x) for x in rng.normal(0, 1.5, 32)]
        dts = [float(x) for x in rng.uniform(0.01, 0.1, 32)]
        out = OB.pid_reference(*gains, errs, dts)
        cases.append(dict(gains=gains, errors=errs, timesteps=dts, outputs=[float(x) for x in out]))
    json.dump(dict(source="reference simple_pid_controller.hpp via oracle/_ref/pid_ref", cases=cases),
              open(os.path.join(HERE, "pid_golden.json"), "w"))


GOLDEN_CASES = (("se2_arena", 32), ("se3_narrow_passage", 48), ("arm_elbow", 24), ("arm_selfcollision", 8), ("arm_table", 48))
# arm_table (BASELINE config 3's contact regime): its solves are rank deficient by round-off, so its
# decision tape also carries the solution of every solve with a condition estimate above 100 -- every particle reproduces
DECISION_COND_LIMIT = {"arm_table": 100.0}


def forward():
    out = {}
    for name, n in GOLDEN_CASES:
        w = W.make(name, n_particles=n)
        orc = OB.OracleSimulator(w.environment().desc, w.robot.to_c(), capi.default_solver_params(), 25.0, 42, 1)
        if name in DECISION_COND_LIMIT:
            OB.lib().oracle_set_decision_cond_limit(orc._h, DECISION_COND_LIMIT[name])
        rec, (draws, offs, dec, dec_offs), sens = OB.run_with_tape(orc, w.starts, w.targets)
        out[name + "_decisions"] = dec
        out[name + "_decision_offsets"] = dec_offs
        out[name + "_cfg"] = rec["cfg"]
        out[name + "_tail"] = np.stack([rec["flags"], rec["n_microsteps"], rec["n_resolver_iters"], rec["n_steps"]], axis=1)
        out[name + "_draws"] = draws
        out[name + "_offsets"] = offs
        out[name + "_sens"] = sens
        out[name + "_stats"] = orc.statistics()
    np.savez_compressed(os.path.join(HERE, "forward_golden.npz"), **out)


ENV_CASES = ("thin_plates", "rotated_boxes")
TRACE_CASES = (("se2_arena", 3), ("arm_elbow", 0), ("se3_narrow_passage", 5))


def environments():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from env_cases import CASES

    out = {}
    for name in ENV_CASES:
        obstacles, res = CASES[name]()
        e = OB.build_environment(obstacles, res)
        out[name + "_shape"] = np.array(e["shape"], dtype=np.int64)
        out[name + "_origin"] = e["origin"]
        out[name + "_occupancy"] = np.packbits(e["occupancy"].reshape(-1))
        out[name + "_sdf"] = e["sdf"]
        out[name + "_cells"] = e["normal_cell_index"]
        out[name + "_starts"] = e["normal_cell_start"]
        out[name + "_entries"] = e["normal_entries"]
    np.savez_compressed(os.path.join(HERE, "env_golden.npz"), **out)


def traces():
    out = {}
    for name, pid_ in TRACE_CASES:
        w = W.make(name, n_particles=8)
        orc = OB.OracleSimulator(w.environment().desc, w.robot.to_c(), capi.default_solver_params(), 25.0, 42, 1)
        t = w.targets[0] if w.targets.shape[0] == 1 else w.targets[pid_]
        res, tr = orc.forward_simulate_traced(w.starts[pid_], t, True, capi.NOISE_PHILOX, particle_id=pid_)
        out[name + "_header"] = np.stack([tr["kind"], tr["step"], tr["microstep"], tr["iteration"]], axis=1)
        out[name + "_values"] = tr["values"]
        out[name + "_cfg"] = res["cfg"]
    np.savez_compressed(os.path.join(HERE, "trace_golden.npz"), **out)


if __name__ == "__main__":
    pid()
    actuator()
    forward()
    environments()
    traces()
    print("golden fixtures written")
