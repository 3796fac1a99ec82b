"""Attribute the warp-stall samples of an `ncu --page source --csv` export (SASS view) to CUDA source lines, using the line table
that nvdisasm prints for the same cubin (the library is built with -lineinfo):
   cuobjdump -xelf all libcontact_addon_b200.so ; nvdisasm -g -c contact_addon_b200.sm_100a.cubin > sass_lines.txt
   python tools/sass_line_profile.py prof.source.csv[.gz] sass_lines.txt k_contac_batch [file-filter]"""
import collections, csv, gzip, re, sys
src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
flt = sys.argv[4] if len(sys.argv) > 4 else ""
rows = list(csv.reader(gzip.open(src_csv, "rt") if src_csv.endswith(".gz") else open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ncu = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
lines = []; cur = None; inside = False
for l in open(dis):
    if l.startswith("//---") and ".text." in l:
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2).strip()))
print("instructions: ncu %d, nvdisasm %d" % (len(ncu), len(lines)))
n = min(len(ncu), len(lines))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for i in range(n):
    r = ncu[i]
    smp = int(r[col["# Samples"]] or 0); ex = int(r[col["Instructions Executed"]] or 0)
    key = lines[i][0]
    a = agg[key]; a[0] += smp; a[1] += ex; tot += smp
    for h in stall:
        v = int(r[col[h]] or 0)
        if v: a[2][h[6:]] += v
print("total samples", tot)
for key, (smp, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    if key is None or (flt and flt not in key[0]):
        continue
    print("%-22s %5d  %6.2f%%  exec %12d  %s" % (key[0], key[1], 100.0 * smp / max(1, tot), ex, ", ".join("%s %d" % kv for kv in st.most_common(3))))
