"""Generates the committed golden fixtures.  Run in the build container (needs /root/reference for the PID
vectors, which come from the REFERENCE's own header through oracle/_ref/pid_ref):

    python tests/golden/make_golden.py

* pid_golden.json        -- outputs of the reference's SimplePIDController (bit-exact pin of the PID restatement)
* forward_golden.npz     -- oracle end states for small seeded batches of every robot kind with the recorded
                            noise tape: the GPU tests replay the tape and compare, so a GPU box without the oracle
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


GOLDEN_CASES = (("se2_arena", 32), ("se3_narrow_passage", 48), ("arm_elbow", 24), ("arm_selfcollision", 8))


def forward():
    out = {}
    for name, n in GOLDEN_CASES:
        w = W.make(name, n_particles=n)
        orc = OB.OracleSimulator(w.environment().desc, w.robot.to_c(), capi.default_solver_params(), 25.0, 42, 1)
        rec, (draws, offs), sens = OB.run_with_tape(orc, w.starts, w.targets)
        out[name + "_cfg"] = rec["cfg"]
        out[name + "_tail"] = np.stack([rec["flags"], rec["n_microsteps"], rec["n_resolver_iters"], rec["n_steps"]], axis=1)
        out[name + "_draws"] = draws
        out[name + "_offsets"] = offs
        out[name + "_sens"] = sens
        out[name + "_stats"] = orc.statistics()
    np.savez_compressed(os.path.join(HERE, "forward_golden.npz"), **out)


if __name__ == "__main__":
    pid()
    forward()
    print("golden fixtures written")
