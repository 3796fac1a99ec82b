"""Dynamic instruction profile of the SteadyGS walker warp (gs_sweeps<1, false, true>: warp 0 of a CTA) from an
`ncu --section SourceCounters --page source --csv` export of k_contac_batch joined with the nvdisasm line table of the same cubin:
   cuobjdump -xelf all libcontact_addon_b200.so ; nvdisasm -g -c contact_addon_b200.sm_100a.cubin > sass_lines.txt
   python tools/chain_profile.py prof.source.csv.gz sass_lines.txt <element steps in the capture> [function-name substring]
Prints warp-level instructions executed per element step by opcode, local-memory accesses per step, the stall-reason split of the
function's samples and the source lines with the most samples."""
import collections, csv, gzip, re, sys
src_csv, dis, steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
want = sys.argv[4] if len(sys.argv) > 4 else "gs_sweepsILi1ELb0ELb1"
kern = "k_contac_batch"
rows = list(csv.reader(gzip.open(src_csv, "rt") if src_csv.endswith(".gz") else open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; col = {n: i for i, n in enumerate(hdr)}
ncu = [r for r in rows[h + 1:] if len(r) >= len(hdr)]
stall = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
lines = []; cur = None; inside = False; fn = ""
for l in open(dis):
    if l.startswith("//---") and ".text." in l:
        inside = kern in l; continue
    if not inside: continue
    if l.startswith("$") and l.rstrip().endswith(":"):
        fn = l.strip(); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m: lines.append((cur, m.group(2).strip(), fn))
assert len(ncu) == len(lines), (len(ncu), len(lines))
op = collections.Counter(); ln_ex = collections.Counter(); ln_smp = collections.Counter(); st = collections.Counter()
tot = smp = nstat = 0
for (key, ins, f), r in zip(lines, ncu):
    if want not in f: continue
    ex = int(r[col["Instructions Executed"]] or 0); sm = int(r[col["# Samples"]] or 0)
    o = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    op[o.split(".")[0]] += ex; ln_ex[key] += ex; ln_smp[key] += sm; tot += ex; smp += sm; nstat += 1
    for n in stall:
        v = int(r[col[n]] or 0)
        if v: st[n[6:]] += v
print("function: ...%s (%d SASS instructions)" % (want, nstat))
print("warp instructions per element step: %.0f" % (tot / steps))
print("by opcode:", ", ".join("%s %.0f" % (o, c / steps) for o, c in op.most_common(26)))
print("local memory per element step: LDL %.2f, STL %.2f" % (op["LDL"] / steps, op["STL"] / steps))
ssum = sum(st.values())
print("stall reasons of the function's %d samples:" % smp, ", ".join("%s %.0f%%" % (k, 100.0 * v / ssum) for k, v in st.most_common(8)))
print("lines by samples (share of the function, instructions per step):")
for k, c in ln_smp.most_common(16):
    if k: print("   %s:%d  %.1f%%  %.1f" % (k[0], k[1], 100.0 * c / smp, ln_ex[k] / steps))
