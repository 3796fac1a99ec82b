"""Summarise an `ncu --page source --csv` export: stall samples by reason and by opcode, executed instructions by opcode.
usage: python tools/stall_summary.py gpurun_out/pt_warp.source.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_reason = collections.Counter(); by_op = collections.Counter(); ex_op = collections.Counter(); samp_total = 0
op_reason = collections.defaultdict(collections.Counter)
wf = collections.Counter(); wf_ideal = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]].split()
    if not src: continue
    op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
    op0 = op.split(".")[0]
    n = int(r[col["# Samples"]] or 0); samp_total += n
    by_op[op0] += n
    ex_op[op0] += int(r[col["Instructions Executed"]] or 0)
    for h in stall_cols:
        v = int(r[col[h]] or 0)
        by_reason[h] += v; op_reason[op0][h] += v
    try:
        wf[op0] += int(r[col["L1 Wavefronts Shared"]] or 0); wf_ideal[op0] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    except ValueError: pass
tot_ex = sum(ex_op.values())
print("samples", samp_total, "warp-instructions executed", tot_ex)
print("by reason:", ", ".join(f"{k[6:]} {v} ({100*v/max(1,samp_total):.0f}%)" for k, v in by_reason.most_common(9)))
print("opcode  samples%  executed%  top reasons")
for op, n in by_op.most_common(16):
    tr = ", ".join(f"{k[6:]} {v}" for k, v in op_reason[op].most_common(3))
    print(f"{op:8s} {100*n/samp_total:6.1f} {100*ex_op[op]/tot_ex:6.1f}   {tr}")
print("shared wavefronts:", {k: (wf[k], wf_ideal[k]) for k in wf if wf[k]})
