"""Developer aid (not a test): device time of fks_env_build_device per phase against the host builder, for the BASELINE
environments.  Prints one JSON line per environment (profiles/r1_env_builder.md is made from them)."""
import json
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

from fast_kinematic_simulator_b200 import simulator as S, workloads as W

names = sys.argv[1:] or ["se2_arena", "se3_narrow_passage", "arm_table", "se3_highres"]
for name in names:
    w = W.make(name, n_particles=4)
    S.build_complete_environment_on_device(w.obstacles, w.resolution).close()  # warm-up (context, first-launch costs)
    best, wall = None, 1e9
    for _ in range(3):
        t = time.perf_counter()
        env = S.build_complete_environment_on_device(w.obstacles, w.resolution)
        wall = min(wall, time.perf_counter() - t)
        tm = env.build_timings_ms
        if best is None or tm["total"] < best["total"]:
            best = tm
        got = env.download() if _ == 0 else None
        env.close()
        if got is not None:
            shape, ncells, nnorm = got.shape, int(np.prod(got.shape)), got.n_normal_cells
            got.close()
    t = time.perf_counter()
    host = w.environment()
    host_s = time.perf_counter() - t
    t = time.perf_counter()
    S.GpuEnvironment(host).close()
    upload_s = time.perf_counter() - t
    # algorithmic HBM bytes of the build: occupancy 1 B written + 1 B read, the z pass writes 4 B,
    # the y and x passes read and write 4 B each, the surface pass clears and reads 8 B and reads the SDF once, the count
    # array is written and read once (1 B), the distance-field check reads the SDF once
    algorithmic = ncells * (1 + 1 + 4 + 8 + 8 + 8 + 8 + 4 + 1 + 1 + 4)
    print(json.dumps({"environment": name, "cells": list(shape), "n_cells": ncells, "surface_normal_cells": nnorm,
                      "obstacles": len(w.obstacles), "device_ms": {k: round(v, 3) for k, v in best.items()},
                      "device_wall_ms": round(wall * 1e3, 2), "host_builder_s": round(host_s, 3), "host_upload_s": round(upload_s, 3),
                      "speedup_vs_host_builder_and_upload": round((host_s + upload_s) / wall, 1),
                      "algorithmic_bytes": algorithmic, "achieved_GBps": round(algorithmic / (best["total"] * 1e-3) / 1e9, 1)}), flush=True)
