"""Run the cases of a parsed .inp file through the CPU oracle (test infrastructure): the same digit semantics as
contact_b200.inp.run_inp, with the sequence state (element division, tractions, previous tractions) carried in Python."""
import numpy as np

from contact_b200 import inp as INP
from oracle import oracle as O


def run_cases(cases):
    out = []
    el = ps = None
    grid = None
    for n, c in enumerate(INP.resolve_cases(cases), 1):
        pc, k, so = c["potcon"], c["kin"], c["solver"]
        hertz = None
        if pc["ipotcn"] < 0:
            key = {-1: ("a1", "b1"), -3: ("aa", "bb")}[pc["ipotcn"]]
            hertz = {"ipotcn": pc["ipotcn"], key[0]: pc["p1"], key[1]: pc["p2"], "scale": pc["scale"]}
            g = dict(mx=pc["mx"], my=pc["my"], xl=0.0, yl=0.0, dx=1.0, dy=1.0, ibase=1, prmudf=[0.0] * 6)
            nn = 0
        else:
            assert pc["ipotcn"] in (1, 3), "tests use IPOTCN = 1, 3 or Hertzian input"
            xl, yl = pc["prm"][0], pc["prm"][1]
            if pc["ipotcn"] == 3:                          # centre of the first element given (potcon_fill)
                xl, yl = xl - 0.5 * pc["prm"][2], yl - 0.5 * pc["prm"][3]
            g = dict(mx=pc["mx"], my=pc["my"], xl=xl, yl=yl, dx=pc["prm"][2], dy=pc["prm"][3],
                     ibase=c["geom"]["ibase"], prmudf=np.array(c["geom"]["prm"], dtype=float))
            nn = int(c["geom"]["prm"][0]) if c["geom"]["ibase"] == 2 else 0
        have_prev = el is not None and grid == (pc["mx"], pc["my"])
        kw = dict(tang=c["T"], norm=c["N"], force3=c["F"], fstat=c["fric"][0], fkin=c["fric"][1], maxgs=so["maxgs"], maxin=so["maxin"],
                  maxnr=so["maxnr"], maxout=so["maxout"], eps=so["eps"], nn=nn, gausei=c["G_eff"], omegah=so.get("omegah", 0.9),
                  omegas=so.get("omegas", 0.9), chi=c["roll"]["chi"], dq=c["roll"]["dq"], hertz=hertz, cphi=k[3])
        if "gdsteady" in so:
            kw["gd"] = tuple(so["gdsteady"])
        kw["pen" if c["N"] == 0 else "fn"] = k[0]
        if c["F"] == 0:
            kw.update(cksi=k[1], ceta=k[2])
        elif c["F"] == 1:
            kw.update(fxrel=k[1], ceta=k[2])
        else:
            kw.update(fxrel=k[1], fyrel=k[2])
        if have_prev:
            if c["I"] >= 1:
                kw.update(iestim=c["I"], el_in=el, ps_in=ps)
            if c["P"] == 0:
                kw["pv_in"] = ps.copy()
            elif c["P"] == 1:
                pv = np.zeros_like(ps); pv[2] = ps[2]
                kw["pv_in"] = pv
        r = O.contac(g, c["mater"]["gg"], c["mater"]["poiss"], **kw)
        el, ps, grid = r["el"].copy(), r["ps"].copy(), (pc["mx"], pc["my"])
        r["ncon"], r["nadh"], r["nslip"] = int((el >= 1).sum()), int((el == 1).sum()), int((el == 2).sum())
        r["pmax"] = float(ps[2].max())
        out.append(r)
    return out
