"""Seeded synthetic inputs shared by the parity tests and bench.py."""
import numpy as np

STEEL = dict(gg=(82000.0, 82000.0), poiss=(0.28, 0.28))
# grid of perfc_test/spence71_8281pt.inp:12 with its quadratic undeformed distance (:15)
HERTZ91 = dict(mx=91, my=91, xl=-4.55, yl=-4.55, dx=0.1, dy=0.1, ibase=1,
               prmudf=[0.004116, 0.0, 0.004116, 0.0, 0.0, 0.0])
# examples/cattaneo.inp:37-46 (second case)
CATTANEO2 = dict(mx=19, my=19, xl=-1.26667, yl=-1.26667, dx=.13333, dy=.13333, ibase=1,
                 prmudf=[0.0100, 0.0, 0.0100, 0.0, 0.0, 0.0], gg=(200.0, 200.0), poiss=(0.42, 0.42), fn=9.1954,
                 maxgs=100, maxin=100, eps=1e-4)


def hertz91_fn(ncase, fn0=2000.0, seed=20240229):
    """FN = FN0 (1 + 0.2 u), u ~ U(-1,1); case i uses draws [4i:4i+4] (first draw = load)."""
    u = np.random.default_rng(seed).uniform(-1.0, 1.0, size=4 * ncase).reshape(ncase, 4)
    return fn0 * (1.0 + 0.2 * u[:, 0]), u


def grid_xy(mx, my, xl, yl, dx, dy):
    x = xl + 0.5 * dx + dx * np.arange(mx)
    y = yl + 0.5 * dy + dy * np.arange(my)
    X, Y = np.meshgrid(x, y)           # (my, mx), x fastest
    return X.ravel(), Y.ravel()


def quadratic_h(g):
    X, Y = grid_xy(g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"])
    b = g["prmudf"]
    return b[0] * X * X + b[1] * X * Y + b[2] * Y * Y + b[3] * X + b[4] * Y + b[5]


def disk_mask(mx, my, frac=0.4):
    iy, ix = np.mgrid[0:my, 0:mx]
    r = frac * min(mx, my)
    return (((ix - (mx - 1) / 2.0) ** 2 + (iy - (my - 1) / 2.0) ** 2) <= r * r).astype(np.int32).ravel()


def microbench_p(mx, my, ncase, seed=7):
    """p ~ N(0,1) on a disk of radius 0.4 min(mx,my), zero elsewhere (kernel microbenchmark input)."""
    rng = np.random.default_rng(seed)
    el = disk_mask(mx, my)
    p = np.zeros((ncase, 3, mx * my))
    p[:] = rng.standard_normal((ncase, 3, mx * my)) * el
    return p, np.tile(el, (ncase, 1))


# ---- reference input files as committed, parsed sequences (tests/golden/*_sequence.json, made by make_fixtures.py) ----
def sequence(name):
    import json, os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "%s_sequence.json" % name)))


def inp_text_from_cases(name):
    """The GPU box has no /root/reference: rebuild an equivalent .inp text from the committed parsed cases."""
    d = sequence(name)
    out = []
    for c in d["cases"]:
        out.append(" 3 MODULE")
        out.append(" %d%d%d%d%d%d" % (c["P"], c["B"], c["T"], c["N"], c["F"], c["S"]))
        out.append(" %d%d%d%d%d%d%d" % (c["V"], c["L"], c["D"], c["C"], c["M"], c["Z"], c["E"]))
        out.append(" %d%d%d%d%d%d%d%d" % (0, c["H"], c["G"], c["I"], c["A"], c["O"], c["W"], c["R"]))      # X = 0: no debug record
        if "solver" in c:
            s = c["solver"]
            out.append(" %d %d %d %d %r" % (s["maxgs"], s["maxin"], s["maxnr"], s["maxout"], s["eps"]))
            if c["G"] in (2, 3):
                out.append(" %r %r %d %r" % (s["omegah"], s["omegas"], s["inislp"], s["omgslp"]))
            elif c["G"] == 4:
                out.append(" %d %r" % (s["inislp"], s["omgslp"]))
            elif c["G"] == 5:                                   # FDECAY BETATH KDOWFB D_IFC D_LIN D_CNS D_SLP POW_S
                gd = s["gdsteady"]
                out.append(" %r %r %d %r %r %r %r %r" % (gd[0], gd[1], int(gd[2]), gd[3], gd[4], gd[5], gd[6], gd[7]))
        out.append(" " + " ".join(repr(v) for v in c["kin"]))
        if "fric" in c:
            out.append(" %r %r" % tuple(c["fric"]))
        if "roll" in c:
            out.append(" %r %r %r" % (c["roll"]["chi"], c["roll"]["dq"], c["roll"]["veloc"]))
        if "mater" in c:
            out.append(" %r %r %r %r" % (c["mater"]["poiss"][0], c["mater"]["poiss"][1], c["mater"]["gg"][0], c["mater"]["gg"][1]))
        if "potcon" in c:
            p = c["potcon"]
            out.append(" %d" % p["ipotcn"])
            if p["ipotcn"] < 0:
                out.append(" %d %d %r %r %r" % (p["mx"], p["my"], p["p1"], p["p2"], p["scale"]))
            else:
                out.append(" %d %d " % (p["mx"], p["my"]) + " ".join(repr(v) for v in p["prm"]))
        if "geom" in c:
            out.append(" %d %d" % (c["geom"]["ibase"], c["geom"]["iplan"]))
            prm = c["geom"]["prm"]
            if c["geom"]["ibase"] == 2:                         # NN XM RM Y1 DY1, then the NN profile heights
                out.append(" %d %r %r %r %r" % (int(prm[0]), prm[1], prm[2], prm[3], prm[4]))
                prm = prm[5:]
            out.append(" " + " ".join(repr(v) for v in prm))
        if c["S"] >= 2:
            out.append(" 0 0")
        if c["S"] >= 3:
            for b in c["subs"]:
                out.append(" %d" % b["isubs"])
                if b["isubs"] in (2, 6):
                    out.append(" %d %d %d" % tuple(b["ix"])); out.append(" %d %d %d" % tuple(b["iy"]))
                if b["isubs"] <= 3:
                    out.append(" %d %r %r" % tuple(b["zparam"]))
                else:
                    out.append(" %d" % len(b["z"])); out.append(" " + " ".join(repr(v) for v in b["z"]))
            out.append(" 0")
    out.append(" 0 MODULE")
    return "\n".join(out) + "\n", d
