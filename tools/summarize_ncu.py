"""Summarise ncu outputs under gpurun_out/ into tracked text files under profiles/ (round tag as argument)."""
import collections
import csv
import gzip
import json
import os
import subprocess
import sys


def page(name, tag, which):
    """CSV of an ncu page: the export made on the GPU box (tools/profile_round.sh) or, failing that, ncu -i on the report."""
    for f in ("gpurun_out/%s_%s.%s.csv" % (name, tag, which), "gpurun_out/%s_%s.%s.csv.gz" % (name, tag, which)):
        if os.path.exists(f) and os.path.getsize(f) > 100:
            return (gzip.open(f, "rt") if f.endswith(".gz") else open(f)).read()
    return subprocess.run(["ncu", "-i", "gpurun_out/%s_%s.ncu-rep" % (name, tag), "--page", which, "--csv"],
                          capture_output=True, text=True).stdout


tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []

# 1. launch list (gpu__time_duration.sum, --clock-control none): share of each kernel in the bench step
rows = [r for r in csv.reader(open("gpurun_out/launches_%s.csv" % tag)) if len(r) > 10]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[r[ki][:70]][0] += 1
    agg[r[ki][:70]][1] += v
tot = sum(v[1] for v in agg.values())
out.append("# ncu launch list: python bench.py --steps 2 --warmup 3 --cpu-seconds 0 (cold-cache, serialised; compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-72s n=%4d total=%10.3f ms share=%6.2f%%" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))

# 2. key metrics of the dominant kernel from the --set full capture (148 cases, one per SM)
raw = page("prof_snorm", tag, "raw")
rr = list(csv.reader(raw.splitlines()))
hh, units, vals = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
out.append("")
out.append("# ncu --set full --clock-control none -k regex:k_snorm_batch : python bench.py --cases 148 --steps 1 --warmup 3")
m = {}
for i, n in enumerate(hh):
    if n in want:
        out.append("%-90s %-12s %s" % (n, units[i], vals[i]))
        m[n] = (units[i], vals[i])


def to_bytes(u, v):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return float(v.replace(",", "")) * f


traffic = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
ncase = 148
tj = {"k_snorm_batch": {"dram_bytes_per_case": traffic / ncase, "cases_in_capture": ncase, "source": "profiles/ncu_%s.txt" % tag}}
out.append("")
out.append("dram traffic per launch (148 cases): %.1f MB = %.2f MB per case" % (traffic / 1e6, traffic / 1e6 / ncase))

# 2b. the headline kernel: k_contac_batch on 148 complete hertz-91 contact cases (N=1, T=3, G=0)
raw = page("prof_contac", tag, "raw")
rr = list(csv.reader(raw.splitlines()))
if len(rr) >= 3:
    hh, units, vals = rr[0], rr[1], rr[2]
    out.append("")
    out.append("# ncu --set full --clock-control none -k regex:k_contac_batch : python bench.py --contact-cases 148 --steps 1 --warmup 3 (hertz-91, T=3, G=0)")
    mc = {}
    for i, n in enumerate(hh):
        if n in want:
            out.append("%-90s %-12s %s" % (n, units[i], vals[i]))
            mc[n] = (units[i], vals[i])
    tc = to_bytes(*mc["dram__bytes_read.sum"]) + to_bytes(*mc["dram__bytes_write.sum"])
    tj["k_contac_batch"] = {"dram_bytes_per_case": tc / ncase, "cases_in_capture": ncase, "source": "profiles/ncu_%s.txt" % tag}
    out.append("")
    out.append("dram traffic per launch (148 cases): %.1f MB = %.2f MB per case" % (tc / 1e6, tc / 1e6 / ncase))
json.dump(tj, open("profiles/traffic_%s.json" % tag, "w"))

# 3. the three phase kernels of the whole-GPU product (575x647), one launch each
for name, title in (("prof_large", "# ncu --set full --clock-control none -k regex:k_lg_ : python tools/large_product_only.py (575x647, 1x1 product)"),
                    ("prof_gd", "# ncu --set full --clock-control none -k regex:k_lg_contac -c 1 : python tools/gd_timing.py 2c (143x161, T=3, G=5 GDsteady on the whole GPU)")):
    raw = page(name, tag, "raw")
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        continue
    hh, units = rr[0], rr[1]
    out.append("")
    out.append(title)
    ki = hh.index("Kernel Name")
    for vals in rr[2:]:
        out.append("## " + vals[ki][:60])
        for i, nme in enumerate(hh):
            if nme in want:
                out.append("%-90s %-12s %s" % (nme, units[i], vals[i]))

# 4. where the warp-stall samples of the dominant kernel sit, by SASS opcode (source page, SASS view)
src = page("prof_snorm", tag, "source")
rs = list(csv.reader(src.splitlines()))
hi = next((i for i, r in enumerate(rs) if r and r[0] == "Address"), None)
if hi is not None:
    hdr = rs[hi]
    si, ci, ei = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    byop = collections.defaultdict(lambda: [0.0, 0.0])
    tot = 0.0
    for r in rs[hi + 1:]:
        if len(r) <= max(ci, ei):
            continue
        try:
            smp, exe = float(r[ci]), float(r[ei])
        except ValueError:
            continue
        toks = r[si].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST", "BAR")) and "." in op else "")
        byop[op][0] += smp; byop[op][1] += exe
        tot += smp
    out.append("")
    out.append("# k_snorm_batch: warp-stall samples by SASS opcode (ncu source page; total %.0f samples)" % tot)
    for op, (smp, exe) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:18]:
        out.append("%-12s %6.2f%% of samples   %12.0f warp-instructions executed" % (op, 100 * smp / max(tot, 1), exe))
else:
    out.append("# (source page not available)")
open("profiles/ncu_%s.txt" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
