"""Summarise ncu outputs under gpurun_out/ into tracked text files under profiles/ (round tag as argument)."""
import collections
import csv
import json
import subprocess
import sys

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
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_snorm_%s.ncu-rep" % tag, "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
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
json.dump({"k_snorm_batch": {"dram_bytes_per_case": traffic / ncase, "cases_in_capture": ncase,
                             "source": "profiles/ncu_%s.txt" % tag}}, open("profiles/traffic_%s.json" % tag, "w"))
out.append("")
out.append("dram traffic per launch (148 cases): %.1f MB = %.2f MB per case" % (traffic / 1e6, traffic / 1e6 / ncase))
open("profiles/ncu_%s.txt" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
