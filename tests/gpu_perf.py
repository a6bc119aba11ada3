"""Developer aid (not a test): device-timed throughput of several workloads / sizes."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from fast_kinematic_simulator_b200 import capi, workloads as W


def run(name, n, free=False, reps=3):
    w = W.make(name, n_particles=n)
    if free and name.startswith("arm"):
        w.targets = (W.ARM_START + np.array([0.3, -0.3, 0.2, -0.3, 0.1, -0.2, 0.4])).reshape(1, 7)
    sim = w.make_simulator()
    dev = torch.device("cuda")
    ds = torch.from_numpy(w.starts).to(dev)
    dt = torch.from_numpy(w.targets).to(dev)
    dr = torch.empty(n * sim.result_stride, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    best = 1e9
    for i in range(reps):
        sim.reset_statistics()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        sim.forward_simulate_device(ds, dt, n, w.targets.shape[0], dr, True, capi.NOISE_PHILOX, stream=st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    s = sim.get_statistics()
    rec = dr.cpu().numpy().view(sim.dtype)
    print("%-20s n=%7d free=%d  %8.2f ms  micro %9d iters %8d pts %9d  -> %.3e microsteps/s  (failed %d, contact %d)" % (
        name, n, free, best, s["total_microsteps"], s["total_resolver_iterations"], s["total_corrected_points"],
        s["total_microsteps"] / best * 1e3, int(((rec["flags"] & 2) != 0).sum()), int((rec["flags"] & 1).sum())), flush=True)
    import ctypes as C
    if hasattr(capi.lib, "fks_debug_phase_cycles"):
        ph = (C.c_uint64 * 112)()
        capi.lib.fks_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p]
        capi.lib.fks_debug_phase_cycles(sim._h, ph)
        tot = float(sum(ph[:10])) or 1.0
        if sum(ph[:10]):
            names = ["A apply", "claim + load", "B measure", "census + park", "T transitions", "start barrier", "C collect", "swap barrier", "D solve + E estimate", "end of cycle"]
            print("    warp clocks: " + ", ".join("%s %.1f%%" % (n, 100 * v / tot) for n, v in zip(names, ph) if n != "-"), flush=True)
            print("    per call: collect %.0f clk (n=%d), solve in shared memory %.0f clk (n=%d), solve in the global store %.0f clk (n=%d), estimate %.0f clk" % (
                ph[10] / max(ph[11], 1), ph[11], ph[12] / max(ph[13], 1), ph[13], ph[14] / max(ph[15], 1), ph[15],
                ph[19] / max(ph[11], 1)), flush=True)
            nw = 148 * 32  # counters are summed over all warps
            cyc = ph[16] / nw
            print("    cycles per CTA %.0f (%.0f clk each), of which solve cycles %.0f with %.1f solver warps each" % (
                cyc, tot / nw / max(cyc, 1), ph[18] / nw, ph[17] / max(ph[18] / 32.0, 1) / 1.0), flush=True)
            print("    cycle length (warp 0 of every CTA): round cycles %.0f clk (n=%d per CTA), solve cycles %.0f clk (n=%d per CTA)" % (
                ph[20] / max(ph[21], 1), ph[21] / 148, ph[22] / max(ph[23], 1), ph[23] / 148), flush=True)
            nr_, ns_ = max(ph[21], 1), max(ph[23], 1)
            print("    slowest warp per cycle (mean over cycles): A %.0f, B %.0f, T %.0f per round cycle; swap %.0f per cycle; collect %.0f, solve %.0f, estimate %.0f per solve cycle" % (
                ph[24] / nr_, ph[25] / nr_, ph[26] / nr_, ph[27] / (nr_ + ns_), ph[28] / ns_, ph[29] / ns_, ph[30] / ns_), flush=True)
            d = ph[48:]
            if sum(d[:16]):
                print("    check_env clocks, bins of 2048: " + " ".join(str(int(v)) for v in d[0:16]), flush=True)
                print("    collect_self clocks, bins of 2048: " + " ".join(str(int(v)) for v in d[16:32]), flush=True)
                print("    A / T / swap clocks, bins of 2048: " + " ".join(str(int(v)) for v in d[40:48]) + " / " + " ".join(str(int(v)) for v in d[48:56]) + " / " + " ".join(str(int(v)) for v in d[56:64]), flush=True)
                print("    check_env mean %.0f clk, collect_self mean %.0f clk; check_env no-collision %d calls mean %.0f clk, collision %d calls mean %.0f clk; self: %d without / %d with" % (
                    d[32] / max(sum(d[:16]), 1), d[33] / max(sum(d[:16]), 1), d[34], d[36] / max(d[34], 1), d[35], d[37] / max(d[35], 1), d[38], d[39]), flush=True)
    sim.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], int(sys.argv[2]), len(sys.argv) > 3 and sys.argv[3] == "free")
    else:
        run("arm_free", 65536)
        run("arm_table", 2368, False)
        run("arm_table", 16384, False)
        run("arm_table", 65536, False)
        run("arm_elbow", 65536, False)
        run("se3_narrow_passage", 16384)
        run("se3_narrow_passage", 65536)
        run("se2_arena", 128)
        run("se2_arena", 65536)
