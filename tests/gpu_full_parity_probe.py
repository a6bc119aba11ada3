"""Developer aid: error distribution of the full-size injected parity run of BASELINE config 3 (tests/test_gpu_full_size.py)."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

from fast_kinematic_simulator_b200 import workloads as W
from oracle import oracle_binding as OB

import parity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
limit = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
w = W.arm_table(n)
rep, gpu, ref, sens = parity.run_parity(w, n, decision_cond_limit=limit)
print(parity.describe(rep, sens))
err = rep["cfg_err"]
print("discrete ok", int(rep["discrete_ok"].sum()), "of", n, "stats equal", rep["gpu_stats"] == rep["oracle_stats"])
for t in (1e-12, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3):
    print("cfg_err > %g: %d" % (t, int((err > t).sum())))
bad = np.flatnonzero(err > 1e-7)
print("sens of those:", [hex(int(sens[i])) for i in bad[:20]], "ill-conditioned flag:", int(((sens[bad] & OB.SENS_ILL_CONDITIONED) != 0).sum()), "of", len(bad))
