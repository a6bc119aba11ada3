"""Summarises one `ncu --set full --import-source on` report of a simulate kernel into markdown:
   python profiles/ncu_summary.py gpurun_out/X.ncu-rep profiles/OUT.md "title" ["command line that was profiled"]
Headline metrics, stall reasons, and per enclosing __device__ function of fks_kernels.cu the share of stall samples and of
executed warp instructions (function boundaries are read from the source embedded in the report)."""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active"]


def summarise(rep, title, command="", launch=0):
    def ncu(*args):
        return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout

    rows = list(csv.reader(ncu("--page", "raw", "--csv").splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + launch]
    m = {k: (vals[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr}
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""

    def tobytes(k):
        v, u = m[k]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    out = ["# %s\n" % title]
    if command:
        out.append("Command (after the same command exited 0 without ncu): `%s`.  Per-launch times under ncu are serialised and cold: "
                   "use them for shares only.\n" % command)
    out.append("Kernel: `%s`\n\n| metric | value |\n|---|---|" % name)
    for k in KEYS:
        if k in m:
            out.append("| `%s` | %s %s |" % (k, m[k][0], m[k][1]))
    traffic = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    out.append("| DRAM traffic of the launch (read + write) | %.1f MB |" % (traffic / 1e6))
    r2 = list(csv.reader(ncu("--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(launch), "--launch-count", "1").splitlines()))
    h2, d2 = r2[1], r2[2:]
    ix = {h: i for i, h in enumerate(h2)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0

    ts = sum(f(r, "# Samples") for r in d2) or 1.0
    stalls = sorted(((sum(f(r, s) for r in d2), s) for s in h2 if s.startswith("stall_") and "Not Issued" not in s), reverse=True)
    out.append("\n## Warp stall reasons (sampled, all samples)\n\n| reason | share |\n|---|---|")
    for v, s in stalls[:10]:
        out.append("| %s | %.1f %% |" % (s, 100 * v / ts))
    r3 = list(csv.reader(ncu("--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch), "--launch-count", "1").splitlines()))
    cur = hdr3 = None
    inst, samp, src = collections.Counter(), collections.Counter(), {}
    for r in r3:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 3 and r[0] == "Line No":
            hdr3 = r
            continue
        if hdr3 is None or len(r) < len(hdr3) or cur is None or r[2] != "-":
            continue
        try:
            key = (cur, int(r[0]))
            inst[key] += float(r[hdr3.index("Instructions Executed")])
            samp[key] += float(r[hdr3.index("# Samples")])
            src[key] = r[1]
        except Exception:
            pass
    ti, tss = sum(inst.values()) or 1.0, sum(samp.values()) or 1.0
    funcs = []
    infile = False
    for r in csv.reader(ncu("--page", "source", "--csv", "--print-source", "cuda", "--launch-skip", str(launch), "--launch-count", "1").splitlines()):
        if len(r) >= 2 and r[0] == "File Name":
            infile = r[1].endswith("fks_kernels.cu")
            continue
        if not infile or len(r) < 2 or not r[0].isdigit():
            continue
        mm = re.match(r"^(?:static )?__(?:device|global)__ .*?(\w+)\(", r[1])
        if mm:
            funcs.append((int(r[0]), mm.group(1)))

    def fn_of(fl, line):
        if fl != "fks_kernels.cu":
            return fl
        nm = "?"
        for st, n in funcs:
            if st <= line:
                nm = n
            else:
                break
        return "simulate_kernel (main loop, barriers)" if nm == "__launch_bounds__" else nm

    fi, fs = collections.Counter(), collections.Counter()
    for k, v in inst.items():
        fi[fn_of(*k)] += v
        fs[fn_of(*k)] += samp[k]
    out.append("\n## Per function (share of stall samples / of executed warp instructions)\n\n| function | samples | instructions |\n|---|---|---|")
    for n, v in fs.most_common(18):
        out.append("| `%s` | %.1f %% | %.1f %% |" % (n, 100 * v / tss, 100 * fi[n] / ti))
    out.append("\n## Hottest source lines (share of samples / of executed instructions)\n\n| samples | instructions | line | source |\n|---|---|---|---|")
    for k, v in samp.most_common(25):
        out.append("| %.1f %% | %.1f %% | %s:%d | `%s` |" % (100 * v / tss, 100 * inst[k] / ti, k[0], k[1], src[k].strip().replace("|", "\\|")[:100]))
    return "\n".join(out) + "\n", traffic


if __name__ == "__main__":
    text, traffic = summarise(sys.argv[1], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "", int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    open(sys.argv[2], "w").write(text)
    print("DRAM traffic bytes", traffic)
