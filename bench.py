#!/usr/bin/env python
"""Benchmark of the batched particle forward-simulate path (BASELINE.json metric:
particle-microsteps/sec with contact resolution).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload arm_table] [--particles P] [--impl reference]

A "step" is ONE ForwardSimulateRobots call (simple_particle_contact_simulator.hpp:788) over one batch of
synthetic particles: 25 controller steps x their microsteps x contact resolution per particle.  Default
workload: BASELINE config 3, the 7-DoF linked arm with 65,536 particles PER GPU (weak scaling; at N GPUs the
job simulates N x 65,536 particles, and for N > 1 every step ends with the NCCL all-gather of end-state records
the north star names).  Noise is counter-based Philox keyed by global particle id.

value  = particle-microsteps / s, whole job, inputs resident in HBM, device-timed (CUDA events), max over ranks.
e2e    = the same metric through the C ABI with HOST buffers (pinned): H2D of starts/targets + kernel + D2H of
         the result records inside the timed region.
--impl reference times the CPU restatement of the reference (oracle/, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle-microsteps/sec (with contact resolution)"
UNIT = "particle-microsteps/s"
DEFAULT_PARTICLES = {"arm_table": 65536, "se3_narrow_passage": 16384, "se2_arena": 128, "se3_highres": 65536,
                     "arm_elbow": 65536, "arm_selfcollision": 4096}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="arm_table")
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the BASELINE config's count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline sample")
    return ap.parse_args()


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, nm in enumerate(names):
                if len(s) > 4 + k and s[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def algorithmic_work(stats, P, D, J):
    """ALGORITHMIC bytes and FP64 flops of a batch from its counters.
    bytes = 4 P per microstep + (28 P + 4 P) per resolver iteration + 56 per corrected point + (16 D + 16) per step;
    flops = (FK + 39 P) per microstep + (3 FK + 174 P + 12 D P) per iteration + (40 + 6 D^2) per corrected point."""
    M = stats["total_microsteps"]
    I = stats["total_resolver_iterations"]
    K = stats["total_corrected_points"]
    S = stats["successful_resolves"] + stats["unsuccessful_resolves"]
    fk = 130 * J + 40 * J if J else (20 if D == 3 else 0)
    nbytes = 4 * P * M + 32 * P * I + 56 * K + (16 * D + 16) * S
    flops = (fk + 39 * P) * M + (3 * fk + P * (36 + 12 * D + 21 + 30) + 36 * P + 39 * P) * I + (40 + 6 * D * D) * K
    return nbytes, flops


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path: here the oracle port (the reference itself cannot be
    compiled without Eigen/ROS/arc_utilities/sdf_tools), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from fast_kinematic_simulator_b200 import capi, workloads as W
    from oracle import oracle_binding as OB

    n_per_gpu = args.particles or DEFAULT_PARTICLES.get(args.workload, 65536)
    w = W.make(args.workload, n_particles=min(n_per_gpu, 8192))
    # all host threads (torchrun exports OMP_NUM_THREADS=1: ask for the cores explicitly)
    orc = OB.OracleSimulator(w.environment().desc, w.robot.to_c(), capi.default_solver_params(), 25.0, 42, host_threads())
    # calibrate the sample so that one step is ~4 s of CPU work
    n0 = min(128, w.n_particles)
    s0, t0 = w.subset(n0)
    t = time.perf_counter()
    orc.forward_simulate(s0, t0, True, capi.NOISE_PHILOX)
    dt = max(time.perf_counter() - t, 1e-3)
    n = int(max(n0, min(w.n_particles, n0 * 4.0 / dt)))
    starts, targets = w.subset(n)
    times, micro = [], 0
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        rec = orc.forward_simulate(starts, targets, True, capi.NOISE_PHILOX)
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
            micro = int(rec["n_microsteps"].sum())
    total = sum(times)
    value = micro * len(times) / total
    sample = "first %d of %d particles of workload %s per step, Philox noise, %d OpenMP threads" % (n, n_per_gpu, w.name, orc.num_threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.name, "particles_per_step": n, "description": w.description,
                   "note": "CPU restatement of the reference (oracle port); the reference needs Eigen/ROS/arc_utilities/sdf_tools"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.num_threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from fast_kinematic_simulator_b200 import capi, workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_local = args.particles or DEFAULT_PARTICLES.get(args.workload, 65536)
    w = W.make(args.workload, n_particles=n_local)
    if world > 1:
        # weak scaling: every rank simulates its own contiguous shard of the (world * n_local)-particle job;
        # starts are regenerated per shard with a shard-specific seed
        w = W.make(args.workload, n_particles=n_local, seed=1003 + rank) if args.workload != "se2_arena" else w
    sim = w.make_simulator(device=local_rank)
    rd = w.robot
    P, D, J = rd.points.shape[0], rd.n_dof, len(rd.joints)
    stride, rec = sim.config_stride, sim.result_stride
    n_targets = w.targets.shape[0]
    first_id = rank * n_local

    dev = torch.device("cuda", local_rank)
    d_starts = torch.from_numpy(w.starts).to(dev)
    d_targets = torch.from_numpy(w.targets).to(dev)
    d_results = torch.empty(n_local * rec, dtype=torch.uint8, device=dev)
    d_gather = torch.empty(world * n_local * rec, dtype=torch.uint8, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def device_step():
        sim.forward_simulate_device(d_starts, d_targets, n_local, n_targets, d_results, True, capi.NOISE_PHILOX,
                                    first_particle_id=first_id, stream=stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(d_gather, d_results)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    barrier()
    sim.reset_statistics()
    launches0 = sim.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the timed events)
        ev[i][0].record(stream)
        kev[i][0].record(stream)
        sim.forward_simulate_device(d_starts, d_targets, n_local, n_targets, d_results, True, capi.NOISE_PHILOX,
                                    first_particle_id=first_id, stream=stream.cuda_stream)
        kev[i][1].record(stream)
        if world > 1:
            dist.all_gather_into_tensor(d_gather, d_results)
        ev[i][1].record(stream)
    barrier()
    step_ms = sum(a.elapsed_time(b) for a, b in ev)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev)
    stats = sim.get_statistics()
    launches = sim.launch_count - launches0
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()

    # ---- end to end through the C ABI with host (pinned) buffers ---------------------------------
    h_starts = torch.from_numpy(w.starts).pin_memory()
    h_targets = torch.from_numpy(w.targets).pin_memory()
    h_results = torch.empty(n_local * rec, dtype=torch.uint8).pin_memory()
    out = h_results.numpy().view(sim.dtype)
    hs, ht = h_starts.numpy(), h_targets.numpy()
    for _ in range(max(1, args.warmup // 2)):
        sim.forward_simulate_robots(hs, ht, True, capi.NOISE_PHILOX, first_particle_id=first_id, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = sim.forward_simulate_robots(hs, ht, True, capi.NOISE_PHILOX, first_particle_id=first_id, out=out)
        if world > 1:
            pass  # the host API returns each rank's records; gathering them on the host is the caller's choice
    barrier()
    e2e_s = time.perf_counter() - t0
    micro_e2e = int(res.n_microsteps.sum())

    # ---- reduce over ranks: max time, sum of work ------------------------------------------------
    vals = torch.tensor([step_ms, kernel_ms, e2e_s], dtype=torch.float64, device=dev)
    work = torch.tensor([float(stats[k]) for k in capi.STAT_NAMES] + [float(micro_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    step_ms, kernel_ms, e2e_s = [float(x) for x in vals.tolist()]
    tot = {k: int(v) for k, v in zip(capi.STAT_NAMES, work.tolist()[:-1])}
    micro_e2e_total = int(work.tolist()[-1])

    if rank == 0:
        value = tot["total_microsteps"] / (step_ms * 1e-3)
        e2e_value = micro_e2e_total * args.steps / e2e_s
        # roofline of the dominant (only) kernel, per launch, from THIS rank's counters and kernel events
        nbytes, flops = algorithmic_work(stats, P, D, J)
        k_s = (sum(a.elapsed_time(b) for a, b in kev) * 1e-3) / args.steps
        peak, how = measured_peaks()
        achieved = nbytes / args.steps / k_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(w.name)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.name, "particles_per_gpu": n_local, "particles_total": n_local * world,
                       "description": w.description, "noise": "philox4x32-10 truncated normal, keyed by global particle id",
                       "controller_hz": 25, "forward_simulation_time_s": 1.0, "l2": "flushed between timed steps (256 MiB fill)",
                       "collective": "ncclAllGather of %d-byte end-state records" % rec if world > 1 else "none (1 GPU)",
                       "microsteps_per_step": tot["total_microsteps"] // args.steps,
                       "resolver_iterations_per_step": tot["total_resolver_iterations"] // args.steps},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * (n_local + n_targets) * stride * 8),
                    "d2h_bytes_per_step": int(world * n_local * rec), "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": how, "kernel": "simulate_kernel<%d>" % w.kind,
                         "kernel_ms_per_launch": 1e3 * k_s, "algorithmic_bytes_per_launch": nbytes // args.steps,
                         "note": "SURVEY 8(d) counts only SDF/normal gathers + particle state; the SDF is L2-resident by design, "
                                 "so the binding unit is the FP64 pipe / LSU, see roofline_fp64"},
            "kernel_info": sim.kernel_info,
        }
        # FP64 and L2-gather denominators measured on this box in the same job
        import ctypes as C
        fp = C.c_double()
        ga = C.c_double()
        gh = C.c_double()
        if capi.lib.fks_measure_fp64_peak(local_rank, C.byref(fp)) == 0:
            af = flops / args.steps / k_s
            line["roofline_fp64"] = {"achieved": af / 1e12, "peak": fp.value / 1e12, "unit": "TFLOP/s", "frac": af / fp.value,
                                     "peak_source": "dependent-free DFMA micro-benchmark in this job",
                                     "algorithmic_flops_per_launch": flops // args.steps}
        if capi.lib.fks_measure_gather_rate(local_rank, 64 << 20, C.byref(ga)) == 0 and \
                capi.lib.fks_measure_gather_rate(local_rank, 2 << 30, C.byref(gh)) == 0:
            gathers = (P * tot_local(stats, "total_microsteps") + 8 * P * tot_local(stats, "total_resolver_iterations")) / args.steps / k_s
            line["gather"] = {"achieved_gathers_per_s": gathers, "l2_peak_gathers_per_s": ga.value, "hbm_peak_gathers_per_s": gh.value,
                              "frac_of_l2_peak": gathers / ga.value}
        if not args.no_cpu_baseline and world == 1:  # reported at N = 1 only
            line["cpu_baseline"] = cpu_baseline(args, w, capi)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def tot_local(stats, key):
    return stats[key]


def cpu_baseline(args, w, capi):
    """The oracle port on this box's host cores, bounded sample of the same workload (rank 0, N = 1 shape)."""
    from oracle import oracle_binding as OB

    orc = OB.OracleSimulator(w.environment().desc, w.robot.to_c(), capi.default_solver_params(), 25.0, 42, host_threads())
    n0 = min(256, w.n_particles)
    s0, t0 = w.subset(n0)
    t = time.perf_counter()
    orc.forward_simulate(s0, t0, True, capi.NOISE_PHILOX)
    dt = max(time.perf_counter() - t, 1e-3)
    n = int(max(n0, min(w.n_particles, n0 * args.cpu_seconds / dt)))
    starts, targets = w.subset(n)
    t = time.perf_counter()
    rec = orc.forward_simulate(starts, targets, True, capi.NOISE_PHILOX)
    dt = time.perf_counter() - t
    return {"value": float(rec["n_microsteps"].sum()) / dt, "unit": UNIT, "cores": orc.num_threads, "kind": "port",
            "sample": "first %d of %d particles of %s, one call, Philox noise, %.1f s" % (n, w.n_particles, w.name, dt)}


if __name__ == "__main__":
    main()
