"""Developer aid (not a test): throughput of the batched CheckConfigCollision against the CPU oracle."""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import parity
from fast_kinematic_simulator_b200 import workloads as W

for name, n in (("arm_table", 1 << 20), ("se3_narrow_passage", 1 << 20)):
    w = W.make(name, n_particles=4)
    rng = np.random.default_rng(5)
    if name == "arm_table":
        cfg = rng.uniform(-2.8, 2.8, (n, 7))
    else:
        base = np.stack([W._se3_config(np.array([rng.uniform(-0.6, 0.6), rng.uniform(-1.6, 1.2), rng.uniform(-0.5, 0.5)]), rng.normal(0, 0.6, 3)) for _ in range(4096)])
        cfg = base[rng.integers(0, 4096, n)]
    sim = w.make_simulator()
    sim.check_config_collision(cfg[:1024])
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        got = sim.check_config_collision(cfg, 0.5)
        best = min(best, time.perf_counter() - t)
    m = 1 << 15
    orc = parity.make_oracle(w)
    t = time.perf_counter()
    ref = orc.check_config_collision(cfg[:m], 0.5)
    cpu = time.perf_counter() - t
    assert np.array_equal(got[:m], ref)
    print("%-20s GPU %.3e configs/s end to end (host buffers, %d configs, %.1f ms)   CPU oracle %.3e configs/s (%d threads)   ratio %.0fx   colliding %.2f" % (
        name, n / best, n, best * 1e3, m / cpu, orc.num_threads, (n / best) / (m / cpu), got.mean()), flush=True)
