"""Developer aid (not a test): the parity comparison of tests/test_gpu_parity.py on larger batches, plus aggregate
outcomes on the workload whose contact solves are round-off coin flips in the reference."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
from fast_kinematic_simulator_b200 import workloads as W
for name, n in (("arm_table", 4096), ("se3_narrow_passage", 8192), ("arm_elbow", 2048), ("se2_arena", 4096), ("arm_selfcollision", 512), ("gantry", 1024)):
    w = W.make(name, n_particles=n)
    rep, gpu, ref, sens = parity.run_parity(w, n)
    print(name, n, "insensitive", rep["n_insensitive"], "matching", rep["n_match"], "BAD insensitive", len(rep["bad_insensitive"]),
          "bad sensitive", len(rep["bad_sensitive"]), "max err insensitive %.3g" % rep["max_err_insensitive"], flush=True)
    if name == "arm_table":
        import numpy as np
        g, r = gpu.records, ref
        print("   aggregate (GPU / oracle): failed %.3f / %.3f, mean microsteps %.1f / %.1f, mean resolver iterations %.1f / %.1f, mean steps %.2f / %.2f" % (
            ((g["flags"] & 2) != 0).mean(), ((r["flags"] & 2) != 0).mean(), g["n_microsteps"].mean(), r["n_microsteps"].mean(),
            g["n_resolver_iters"].mean(), r["n_resolver_iters"].mean(), g["n_steps"].mean(), r["n_steps"].mean()))
