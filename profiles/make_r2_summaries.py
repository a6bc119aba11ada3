"""Regenerates the tracked round-2 summaries from gpurun_out/ (scratch, written by profiles/run_round2_measurements.sh):
   python profiles/make_r2_summaries.py     (needs `ncu` on PATH to read gpurun_out/prof_r2_*.ncu-rep)
 -> profiles/r2_bench.md, r2_ncu_{arm_table,se3_narrow_passage,se3_highres}.md, r2_launches_arm_table.csv, traffic.json"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
B = "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config5-particles 0"
CAPTURES = [
    ("prof_r2_arm_table.ncu-rep", "r2_ncu_arm_table.md", "arm_table",
     "Round 2 — ncu --set full, simulate_kernel<LINKED>, BASELINE config 3 (arm_table, 65 536 particles)", B),
    ("prof_r2_se3_narrow.ncu-rep", "r2_ncu_se3_narrow_passage.md", "se3_narrow_passage",
     "Round 2 — ncu --set full, simulate_kernel<SE3>, BASELINE config 2 (se3_narrow_passage, 16 384 particles)",
     B + " --workload se3_narrow_passage --particles 16384"),
    ("prof_r2_se3_highres.ncu-rep", "r2_ncu_se3_highres.md", "se3_highres",
     "Round 2 — ncu --set full, simulate_kernel<SE3>, BASELINE config 4 (se3_highres, 511^3 SDF, 65 536 particles)",
     B + " --workload se3_highres"),
]


def load(name):
    try:
        return json.loads(open(os.path.join(OUT, name)).read().strip().splitlines()[-1])
    except Exception:
        return None


def ncu_summaries():
    traffic = {}
    for rep, md, workload, title, cmd in CAPTURES:
        path = os.path.join(OUT, rep)
        if not os.path.exists(path):
            continue
        out = subprocess.run([sys.executable, os.path.join(PROF, "ncu_summary.py"), path, os.path.join(PROF, md), title, cmd],
                             capture_output=True, text=True).stdout
        for line in out.splitlines():
            if line.startswith("DRAM traffic bytes"):
                traffic[workload] = float(line.split()[-1])
    if traffic:
        traffic["source"] = ("ncu --set full, profiles/r2_ncu_*.md (dram__bytes_read.sum + dram__bytes_write.sum of one simulate_kernel "
                             "launch at the bench's particle count for that workload)")
        json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    # config 4: link-level culling of check_env on / off
    rows = []
    for c in (1, 0):
        path = os.path.join(OUT, "dram_r2_highres_cull%d.csv" % c)
        if os.path.exists(path):
            import csv
            vals = {r["Metric Name"]: (float(r["Metric Value"].replace(",", "")), r["Metric Unit"]) for r in csv.DictReader(
                l for l in open(path) if l.startswith('"'))}
            rows.append((c, vals))
    md = os.path.join(PROF, "r2_ncu_se3_highres.md")
    if rows and os.path.exists(md):
        L = ["\n## Link-level culling of `check_env` on / off (`FKS_CULL=1` default / `FKS_CULL=0`; same command, `ncu --metrics dram__bytes_read.sum,"
             "dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors.sum -k regex:simulate_kernel -s 1 -c 1`)\n",
             "| culling | kernel duration | DRAM read | DRAM written | L2 sectors |", "|---|---|---|---|---|"]
        for c, v in rows:
            def g(k):
                return "%.4g %s" % v[k] if k in v else "-"
            L.append("| %s | %s | %s | %s | %s |" % ("on" if c else "off", g("gpu__time_duration.sum"), g("dram__bytes_read.sum"),
                                                      g("dram__bytes_write.sum"), g("lts__t_sectors.sum")))
        L.append("\nCulling tests one cell per link centre before the link's points: a link whose centre-cell distance exceeds its bounding radius "
                 "+ sqrt(3) cells cannot collide (exact: the SDF is a distance field, verified at upload).  In this workload half of the particles "
                 "fly through free space, where culling removes nearly all of the 4 bytes per point and microstep that SURVEY 8(d) counts as "
                 "algorithmic; the other half is pressed against a cuboid face and gathers from the HBM-resident SDF in every resolver iteration.\n")
        open(md, "a").write("\n".join(L))
    src = os.path.join(OUT, "launches_r2.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, "r2_launches_arm_table.csv"))


def bench_md():
    main, ref = load("bench_r2.json"), load("bench_r2_reference.json")
    if not main:
        return
    L = ["# Round 2 — bench lines (B200, one GPU, fresh box, SM clock %s MHz, throttle reasons %s)\n" % (
             main["clocks"]["sm_mhz"], main["clocks"]["reasons"]),
         "`python bench.py` (default: BASELINE config 3, `arm_table`, 65 536 particles, 10 timed steps after 3 warm-ups, L2 flushed between "
         "steps) and the same command with `--workload ... [--particles ...]` for the other BASELINE configurations.  A step is one "
         "`ForwardSimulateRobots` call: 25 controller steps x microsteps x contact resolution per particle.  `value` = device-resident inputs; "
         "`e2e` = the C-ABI call with host buffers (H2D, kernel, D2H inside the timed region); CPU = the oracle port on the box's host cores over "
         "a bounded sample of the same workload.\n",
         "| workload (BASELINE config) | particles | ms / step | particle-microsteps/s (device) | e2e (host buffers) | microsteps / step | "
         "resolver iterations / step | roofline (bound: fraction) | FP64 fraction | CPU oracle port (cores) | e2e / CPU |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]

    def row(d, tag):
        c, cb, r, r64 = d["config"], d.get("cpu_baseline") or {}, d.get("roofline") or {}, d.get("roofline_fp64") or {}
        return "| %s %s | %d | %.2f | %.3e | %.3e | %d | %d | %s: %.3f | %.3f | %s | %s |" % (
            c["workload"], tag, c["particles_per_gpu"], d["ms_per_step"], d["value"], d["e2e"]["value"], c["microsteps_per_step"],
            c["resolver_iterations_per_step"], r.get("bound", "-"), r.get("frac", 0.0), r64.get("frac", 0.0),
            ("%.3e (%d)" % (cb["value"], cb["cores"])) if cb.get("value") else "-",
            ("%.0fx" % (d["e2e"]["value"] / cb["value"])) if cb.get("value") else "-")

    L.append(row(main, "(3)"))
    for f, tag in (("bench_r2_se3_16384.json", "(2)"), ("bench_r2_se2_128.json", "(1)"), ("bench_r2_se3_highres.json", "(4, 511^3 SDF)"),
                   ("bench_r2_arm_2368.json", "(small batch)"), ("bench_r2_arm_free.json", ""), ("bench_r2_arm_elbow.json", "")):
        d = load(f)
        if d:
            L.append(row(d, tag))
    c5 = main.get("config5")
    if c5:
        L.append("\nBASELINE config 5 on one GPU (`config5` of the default line): %d arm particles in %.1f ms = %.3e particle-microsteps/s "
                 "(the same kernel without the tail of a 65 536-particle call: %.0f %% more).\n" % (
                     c5["particles_total"], c5["ms_per_step"], c5["value"], 100.0 * (c5["value"] / main["value"] - 1.0)))
    if ref:
        L.append("`--impl reference` (oracle port, all host threads, bounded sample, the product library never loaded: `native_libraries` = %s): "
                 "%.3e particle-microsteps/s on %d cores (%s).\n" % (ref["config"].get("native_libraries"), ref["value"],
                                                                    ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["sample"]))
    r, rh, r64 = main["roofline"], main.get("roofline_hbm"), main.get("roofline_fp64")
    L.append("Roofline objects of the default line: `%s` achieved %.1f of %.1f %s = %.3f; HBM %.1f of %.0f GB/s = %.3f (measured DRAM traffic "
             "of one launch: %s bytes against %.3e algorithmic bytes); FP64 %.2f of %.1f TFLOP/s = %.3f.  No unit binds: the kernel is bound by "
             "the latency of one warp's dependent instruction stream and by the lock-step barrier (`r2_ncu_arm_table.md`, "
             "`r2_kernel_experiments.md`).\n" % (r["bound"], r["achieved"], r["peak"], r["unit"], r["frac"], rh["achieved"], rh["peak"], rh["frac"],
                                                 r["traffic"], r["algorithmic_bytes_per_launch"], r64["achieved"], r64["peak"], r64["frac"]))
    L.append("Full JSON of the default run:\n\n```json\n%s\n```\n" % json.dumps(main))
    for name, title in (("bench_r2_2gpu.json", "Two GPUs"), ("bench_r2_4gpu.json", "Four GPUs"), ("bench_r2_8gpu.json", "Eight GPUs")):
        d = load(name)
        if d:
            c5 = d.get("config5") or {}
            L.append("## %s (`gpurun --gpus %d`, `python -m torch.distributed.run --nproc-per-node %d ... bench.py --gpus %d`, NCCL all-gather of the "
                     "72-byte end-state records and the D2H of all records inside `e2e`)\n\nweak scaling: %.3e particle-microsteps/s over %d "
                     "particles, %.1f ms/step, e2e %.3e; config 5 (1 048 576 particles in total, strong scaling): %s ms/step = %s "
                     "particle-microsteps/s.\n\n```json\n%s\n```\n" % (
                         title, d["n_gpus"], d["n_gpus"], d["n_gpus"], d["value"], d["config"]["particles_total"], d["ms_per_step"],
                         d["e2e"]["value"], "%.1f" % c5["ms_per_step"] if c5 else "-", "%.3e" % c5["value"] if c5 else "-", json.dumps(d)))
    open(os.path.join(PROF, "r2_bench.md"), "w").write("\n".join(L))


if __name__ == "__main__":
    ncu_summaries()
    bench_md()
