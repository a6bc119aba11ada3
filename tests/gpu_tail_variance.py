"""Developer aid (not a test): small batches are timed by their slowest particle -- run time against the largest per-particle
resolver-iteration count for several noise seeds (arm_table)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from fast_kinematic_simulator_b200 import capi, workloads as W
for n in (2368, 16384):
    w = W.make("arm_table", n_particles=n)
    for seed in (42, 43, 44, 45, 46):
        sim = w.make_simulator(seed=seed)
        dev = torch.device("cuda")
        ds = torch.from_numpy(w.starts).to(dev); dt = torch.from_numpy(w.targets).to(dev)
        dr = torch.empty(n * sim.result_stride, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream()
        best = 1e9
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            sim.forward_simulate_device(ds, dt, n, w.targets.shape[0], dr, True, capi.NOISE_PHILOX, stream=st.cuda_stream)
            e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        rec = dr.cpu().numpy().view(sim.dtype)
        work = rec["n_microsteps"].astype(np.int64) + 8 * rec["n_resolver_iters"].astype(np.int64)
        print("n=%d seed=%d  %.2f ms  max iters %d  max micro %d  max work %d mean work %.0f" % (n, seed, best, rec["n_resolver_iters"].max(), rec["n_microsteps"].max(), work.max(), work.mean()), flush=True)
        sim.close()
