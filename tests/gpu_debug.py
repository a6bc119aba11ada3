"""Developer aid (not a test): run every parity workload on the GPU and print the reports."""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import parity
from fast_kinematic_simulator_b200 import capi, workloads as W

cases = [("se2_arena", 128), ("se3_narrow_passage", 512), ("arm_table", 256), ("arm_selfcollision", 64)]
if len(sys.argv) > 1:
    cases = [(sys.argv[1], int(sys.argv[2]))]
for name, n in cases:
    w = W.make(name, n_particles=n)
    t = time.time()
    rep, gpu, ref, sens = parity.run_parity(w, n)
    print("==", name, n, "%.2fs" % (time.time() - t))
    print(parity.describe(rep, sens))
    print("  gpu stats", rep["gpu_stats"])
    print("  orc stats", rep["oracle_stats"])
    bad = rep["bad_insensitive"][:5]
    for i in bad:
        print("  BAD", i, "gpu", gpu.records[i], "ref", ref[i])
