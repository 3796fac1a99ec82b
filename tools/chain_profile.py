"""Dynamic instruction profile of the SteadyGS element-step chain (warp 0 of a CTA) from an `ncu --page source --csv` export of
k_contac_batch joined with the nvdisasm line table of the same cubin (see tools/sass_line_profile.py for the two inputs):
   python tools/chain_profile.py prof.source.csv.gz sass_lines.txt <element steps in the capture>
Prints, for the source-line range of gs_walk_row (steady_solver.cuh), warp-level instructions executed per element step by opcode,
the local-memory accesses (register spills, stack arguments) per step, and the hottest source lines."""
import collections, csv, gzip, re, sys
src_csv, dis, steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 0)
kern = "k_contac_batch"
rows = list(csv.reader(gzip.open(src_csv, "rt") if src_csv.endswith(".gz") else open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; col = {n: i for i, n in enumerate(hdr)}
ncu = [r for r in rows[h + 1:] if len(r) >= len(hdr)]
lines = []; cur = None; inside = False; fn = ""
for l in open(dis):
    if l.startswith("//---") and ".text." in l:
        inside = kern in l; continue
    if not inside: continue
    if l.startswith("$") and l.rstrip().endswith(":"):
        fn = l; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m: lines.append((cur, m.group(2).strip(), fn))
assert len(ncu) == len(lines), (len(ncu), len(lines))
# the instantiation with the most executed instructions among the functions whose name contains gs_walk_row (or, for builds
# before the walk was a function of its own, the given source-line range of stdygs_dev)
per_fn = collections.Counter()
for (key, ins, f), r in zip(lines, ncu):
    if "gs_walk_row" in f or (lo and key and key[0] == "steady_solver.cuh" and lo <= key[1] <= hi and "stdygs_dev" in f):
        per_fn[f] += int(r[col["Instructions Executed"]] or 0)
best = per_fn.most_common(1)[0][0]
op = collections.Counter(); ln = collections.Counter(); tot = 0
for (key, ins, f), r in zip(lines, ncu):
    if f != best: continue
    if lo and not (key and key[0] == "steady_solver.cuh" and lo <= key[1] <= hi) and "gs_walk_row" not in f: continue
    ex = int(r[col["Instructions Executed"]] or 0)
    o = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    op[o.split(".")[0]] += ex; ln[key] += ex; tot += ex
print("function:", best.strip()[:60], "...", best.strip()[-70:])
print("warp instructions per element step: %.0f" % (tot / steps))
print("by opcode:", ", ".join("%s %.0f" % (o, c / steps) for o, c in op.most_common(24)))
print("local memory per step: LDL %.1f, STL %.1f" % (op["LDL"] / steps, op["STL"] / steps))
print("hottest lines:", ", ".join("%s:%d %.0f" % (k[0], k[1], c / steps) for k, c in ln.most_common(14) if k))
