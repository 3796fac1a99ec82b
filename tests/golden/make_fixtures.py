"""Generate golden fixtures from the reference tree (/root/reference, only available in the build container).

Run once here; the outputs are committed so that nothing under tests/, smoke() or bench.py reads /root/reference
at run time.  Sources: perfc_test/norm_problm_1p.inp (IBASE=2 height table of the Manchester-benchmark right
wheel at 6.2 mm, used by the whole perf suite), perfc_test/get_times.ref_out (iteration counts),
examples/cattaneo.ref_out (element picture, statistics).
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def mbench_profile():
    lines = open(os.path.join(REF, "perfc_test/norm_problm_1p.inp")).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if "NN, XM" in l)
    hdr = lines[i0].split("%")[0].split()
    nn, xm, rm, y1, dy1 = int(hdr[0]), float(hdr[1]), float(hdr[2]), float(hdr[3]), float(hdr[4])
    vals = []
    for l in lines[i0 + 1:]:
        t = l.split("%")[0].split()
        if not t:
            break
        vals += [float(v) for v in t]
        if len(vals) >= nn:
            break
    assert len(vals) == nn
    return dict(nn=nn, xm=xm, rm=rm, y1=y1, dy1=dy1, heights=vals, pen=0.02084,
                source="perfc_test/norm_problm_1p.inp:16-71")


def get_times():
    out = {}
    for l in open(os.path.join(REF, "perfc_test/get_times.ref_out")):
        m = re.match(r"\s*(\w+)\s*:\s*ncon=\s*(\d+), ItCG=\s*(\d+)", l)
        if m:
            out[m.group(1)] = dict(ncon=int(m.group(2)), itcg=int(m.group(3)))
        m = re.match(r"\s*(\w+)\s*:\s*nslp=\s*(\d+), ItGS=\s*(\d+)", l)
        if m:
            out[m.group(1)] = dict(nslp=int(m.group(2)), itgs=int(m.group(3)))
    return out


def ref_out_stats(path):
    """Per case of a .ref_out file: the statistics row NPOT NCON NADH NSLIP INORM ITANG and the row
    FN/G FX/FSTAT/FN FY/FSTAT/FN APPROACH PMAX as printed."""
    lines = open(os.path.join(REF, path)).read().splitlines()
    cases = []
    for i, l in enumerate(lines):
        m = re.match(r"\s*FN/G\s+(FX/FSTAT/FN|SHIFT X|CREEP X)\s+(FY/FSTAT/FN|SHIFT Y|CREEP Y)\s+APPROACH", l)
        if m:       # columns 2, 3 are the relative forces (F = 0) or the resulting shifts / creepages (F = 1, 2)
            cases.append(dict(forces=lines[i + 1].split(), col2=m.group(1), col3=m.group(2)))
        if re.match(r"\s*NPOT\s+NCON\s+NADH\s+NSLIP\s+INORM\s+ITANG", l):
            cases[-1]["stats"] = [int(v) for v in lines[i + 1].split()]
    return cases


def inp_cases(path):
    """The module-3 cases of a reference .inp file as parsed by contact_b200.inp (derived data: lets the GPU box, which
    has no /root/reference, run the same sequence)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from contact_b200 import inp as INP
    return INP.parse_inp(open(os.path.join(REF, path)).read())


def cattaneo():
    txt = open(os.path.join(REF, "examples/cattaneo.ref_out")).read().splitlines()
    pics = []
    i = 0
    while i < len(txt):
        if "FORM OF THE CONTACT" in txt[i]:
            rows = []
            i += 1
            while re.match(r"\s+\d+\s+[o.*S|]", txt[i]):
                rows.append(txt[i].split()[1:])
                i += 1
            pics.append(rows[::-1])     # row iy=1 first
        i += 1
    return dict(pictures=pics, source="examples/cattaneo.ref_out")


def subsurf():
    """examples/subsurf.ref_subs (second case: pn = 1, px = 0.999 on the unit square): block 1 fully, block 2 thinned."""
    blocks, cur = [], None
    for l in open(os.path.join(REF, "examples/subsurf.ref_subs")):
        if l.startswith("%"):
            if "NX" in l:
                cur = dict(dims=None, rows=[])
                blocks.append(cur)
            continue
        vals = [float(v) for v in l.split()]
        if cur["dims"] is None:
            cur["dims"] = [int(v) for v in vals[:3]]
        else:
            cur["rows"].append(vals)
    b2 = blocks[1]["rows"]
    keep = b2[::7]                                   # every 7th point keeps the fixture small (~950 rows)
    return dict(source="examples/subsurf.ref_subs, examples/subsurf.inp:21-50",
                grid=dict(mx=1, my=1, xl=-0.4999, yl=-0.4998, dx=1.0, dy=1.0), gg=[1.0, 1.0], poiss=[0.28, 0.28],
                pn=1.0, px=0.999, columns="X Y Z UX UY UZ SIGHYD SIGVM SIGXX SIGXY SIGXZ SIGYY SIGYZ SIGZZ",
                block1=blocks[0]["rows"], block2_every7=keep)


def write_sequences():
    for name in ("spence35", "cattaneo", "carter2d"):
        json.dump(dict(source="examples/%s.inp, examples/%s.ref_out" % (name, name), cases=inp_cases("examples/%s.inp" % name),
                       ref_out=ref_out_stats("examples/%s.ref_out" % name)),
                  open(os.path.join(HERE, "%s_sequence.json" % name), "w"), indent=0)


def write_spence71():
    json.dump(dict(source="perfc_test/spence71_8281pt.inp; perfc_test/get_times.ref_out:76-79 (spence71_nosubs: ncon 3657, nout 511)",
                   cases=inp_cases("perfc_test/spence71_8281pt.inp"), golden=dict(ncon=3657, nout=511)),
              open(os.path.join(HERE, "spence71_sequence.json"), "w"), indent=0)


def write_tang_problm_c():
    """perfc_test/tang_problm_{1,2}c.inp (T=3, G=5 GDsteady with its 8-parameter record) as parsed cases."""
    cs = []
    for k in (1, 2):
        cs += inp_cases("perfc_test/tang_problm_%dc.inp" % k)
    json.dump(dict(source="perfc_test/tang_problm_1c.inp, tang_problm_2c.inp", cases=cs),
              open(os.path.join(HERE, "tang_problm_c_sequence.json"), "w"), indent=0)


def write_test_table():
    """testbank/test_table_mx11.ref_fx: 3220 Hertzian creepage cases on an 11x11 grid (NormCG + SteadyGS), Fx, Fy, Mz relative
    to mu Fn (Mz also to cp) with 6 decimals and the slip share; case order of src/test_table.f90:196-292."""
    lines = open(os.path.join(REF, "testbank/test_table_mx11.ref_fx")).read().splitlines()
    hz, rows = [], []
    for l in lines:
        m = re.match(r"iell=(\d+), a/b=\s*([\d.]+): rho=\s*([\d.Ee+-]+)\s*, cp=\s*([\d.Ee+-]+)", l)
        if m:
            hz.append(dict(iell=int(m.group(1)), aob=float(m.group(2)), rho=float(m.group(3)), cp=float(m.group(4))))
        m = re.match(r"Fx=\s*([-\d.]+), Fy=\s*([-\d.]+), Mz=\s*([-\d.]+), %Slip=\s*([\d.]+)", l)
        if m:
            rows.append([float(m.group(1)), float(m.group(2)), float(m.group(3)), float(m.group(4))])
    assert len(hz) == 5 and len(rows) == 5 * 4 * 7 * 23
    json.dump(dict(source="testbank/test_table_mx11.ref_fx (driver: src/test_table.f90)", hertz=hz, columns=["fx", "fy", "mz", "pct_slip"],
                   rows=rows), open(os.path.join(HERE, "test_table_mx11.json"), "w"))


if __name__ == "__main__":
    write_test_table()
    write_tang_problm_c()
    write_sequences()
    write_spence71()
    json.dump(subsurf(), open(os.path.join(HERE, "subsurf_ref_subs.json"), "w"))
    json.dump(mbench_profile(), open(os.path.join(HERE, "mbench_profile.json"), "w"))
    json.dump(get_times(), open(os.path.join(HERE, "get_times.json"), "w"), indent=1)
    json.dump(cattaneo(), open(os.path.join(HERE, "cattaneo_pictures.json"), "w"))
    print("fixtures written")
