"""Reader for CONTACT's .inp text format (module 3) and a driver that runs the cases through the cntc_* C-ABI.

The second drop-in boundary of the hot path: record order, control digits and the "0 = keep the
value of the previous case" rules follow /root/reference/src/m_sinput.f90:34-317 (input), :321-597 (ic_input),
:601-689 (solv_input), :693-758 (kincns_input), :762-811 (veloc_input), :815-1028 (mater_input), :1130-1297
(potcon_input), :1301-1502 (geom_input) and /root/reference/src/m_subsurf.f90:65-298 (subsurf_input).  Records are line
based, '%' starts a comment, Fortran 'd' exponents are accepted.  Inputs outside the hot-path scope (module 1, friction
laws L > 1, materials M > 0, planforms, temperature ...) raise NotImplementedError instead of being guessed.
"""
import re
import time

import numpy as np

__all__ = ["parse_inp", "run_inp", "InpError"]


class InpError(ValueError):
    pass


class _Reader:
    def __init__(self, text):
        self.lines = []
        for no, raw in enumerate(text.splitlines(), 1):
            s = raw.split("%", 1)[0].strip()
            if s:
                self.lines.append((no, s))
        self.pos = 0

    def eof(self):
        return self.pos >= len(self.lines)

    @staticmethod
    def _num(tok):
        t = tok.strip().rstrip(",")
        try:
            return int(t)
        except ValueError:
            return float(re.sub(r"[dD]", "e", t))

    def readline(self, what, nmin):
        """One record: at least nmin numbers from the next non-empty line."""
        if self.eof():
            raise InpError("unexpected end of file while reading " + what)
        no, s = self.lines[self.pos]
        self.pos += 1
        vals = []
        for tok in re.split(r"[\s,]+", s):
            if not tok:
                continue
            try:
                vals.append(self._num(tok))
            except ValueError:
                break                      # trailing description text
        if len(vals) < nmin:
            raise InpError("line %d: expected %d values for %s, got %r" % (no, nmin, what, s))
        return vals

    def read1darr(self, what, n):
        """n numbers spread over as many lines as needed."""
        out = []
        while len(out) < n:
            out += self.readline(what, 1)
        return [float(v) for v in out[:n]]


def _digits(word, n):
    """Unpack an n-digit control word, most significant digit first (ic_unpack, m_hierarch_data.f90:1128-1216)."""
    return [(int(word) // 10 ** (n - 1 - i)) % 10 for i in range(n)]


def parse_inp(text):
    """Parse module-3 cases.  Returns a list of dicts with the digits and only those inputs that the record order of the
    reference reads for these digits ("kept" inputs are absent and inherit from the previous case)."""
    rd = _Reader(text)
    cases = []
    state = dict(tang=0, mx=None, my=None, ipotcn=1)
    while not rd.eof():
        modul = int(rd.readline("module number", 1)[0])
        if modul == 0:
            break
        if modul != 3:
            raise NotImplementedError("module %d is outside the hot-path scope (only module 3)" % modul)
        c = {}
        w = int(rd.readline("control integers pbtnfs", 1)[0])
        c["P"], c["B"], c["T"], c["N"], c["F"], c["S"] = _digits(w % 1000000, 6)
        r2 = rd.readline("control integers vldcmze", 1)
        c["V"], c["L"], c["D"], c["C"], c["M"], c["Z"], c["E"] = _digits(int(r2[0]), 7)
        mater2 = int(r2[1]) % 10 if len(r2) >= 2 else 1
        w = int(rd.readline("control integers hgiaowr", 1)[0])
        c["X"], c["H"], c["G"], c["I"], c["A"], c["O"], c["W"], c["R"] = _digits(w, 8)
        ncase = len(cases) + 1
        if ncase == 1:                                      # ic_input: the first case initiates contact
            c["P"], c["I"] = 2, 0
        if c["X"] >= 1:
            rd.readline("debug output psflcin", 1)
        if c["V"] != 0 or c["L"] > 1 or c["M"] != 0 or c["H"] != 0 or c["B"] != 0 or mater2 >= 2:
            raise NotImplementedError("case %d: digits V=%d L=%d M=%d H=%d B=%d are outside the hot-path scope" %
                                      (ncase, c["V"], c["L"], c["M"], c["H"], c["B"]))
        if c["G"] != 1:
            r = rd.readline("iteration constants", 5)
            c["solver"] = dict(maxgs=int(r[0]), maxin=int(r[1]), maxnr=int(r[2]), maxout=int(r[3]), eps=float(r[4]))
            if c["G"] in (2, 3):
                r = rd.readline("relaxation parameters", 4)
                c["solver"].update(omegah=float(r[0]), omegas=float(r[1]), inislp=int(r[2]), omgslp=float(r[3]))
            elif c["G"] == 4:
                r = rd.readline("relaxation parameters", 2)
                c["solver"].update(inislp=int(r[0]), omgslp=float(r[1]))
            elif c["G"] == 5:
                r = rd.readline("parameters for gdsteady", 8)
                c["solver"].update(gdsteady=[float(v) for v in r[:8]])
        r = rd.readline("kinematic inputs", 4)
        c["kin"] = [float(v) for v in r[:4]]                 # FN|PEN, CKSI|FX, CETA|FY, CPHI
        if c["L"] != 1:
            r = rd.readline("friction parameters", 2)
            c["fric"] = (float(r[0]), float(r[1]))
        is_roll = c["T"] in (2, 3)
        if 2 <= c["C"] <= 4 or c["C"] == 9:
            if c["C"] != 2:
                raise NotImplementedError("case %d: C=%d (only piecewise-constant analytical coefficients, C=2)" % (ncase, c["C"]))
            if is_roll:
                r = rd.readline("rolling direction and step", 3)
                c["roll"] = dict(chi=float(r[0]), dq=float(r[1]), veloc=float(r[2]))
            r = rd.readline("material properties", 4)
            c["mater"] = dict(poiss=(float(r[0]), float(r[1])), gg=(float(r[2]), float(r[3])))
        if c["D"] == 2:
            ip = int(rd.readline("IPOTCN", 1)[0])
            if -5 <= ip <= -1:
                r = rd.readline("Hertzian potential contact", 5)
                c["potcon"] = dict(ipotcn=ip, mx=int(r[0]), my=int(r[1]), p1=float(r[2]), p2=float(r[3]), scale=float(r[4]))
            elif 1 <= ip <= 4:
                r = rd.readline("potential contact", 6)
                c["potcon"] = dict(ipotcn=ip, mx=int(r[0]), my=int(r[1]), prm=[float(v) for v in r[2:6]])
            else:
                raise NotImplementedError("case %d: IPOTCN=%d" % (ncase, ip))
            state.update(mx=c["potcon"]["mx"], my=c["potcon"]["my"], ipotcn=ip)
        if c["Z"] >= 2 and state["ipotcn"] > 0:
            r = rd.readline("IBASE, IPLAN", 2)
            ibase, iplan = int(r[0]), int(r[1])
            if iplan != 1:
                raise NotImplementedError("case %d: IPLAN=%d (only the unrestricted planform)" % (ncase, iplan))
            if ibase == 1:
                prm = rd.read1darr("geometry B(1:6)", 6)
            elif ibase == 2:
                r = rd.readline("geometry NN, XM, RM, Y1, DY1", 5)
                nn = int(r[0])
                prm = [float(nn)] + [float(v) for v in r[1:5]] + rd.read1darr("profile heights", nn)
            elif ibase == 3:
                prm = rd.read1darr("geometry B(1:8)", 8)
            elif ibase == 9:
                prm = rd.read1darr("undeformed distance per element", state["mx"] * state["my"])
            else:
                raise InpError("case %d: invalid IBASE=%d" % (ncase, ibase))
            c["geom"] = dict(ibase=ibase, iplan=iplan, prm=prm)
        if c["E"] == 9:
            raise NotImplementedError("case %d: E=9 (extra rigid slip per element)" % ncase)
        if c["S"] >= 2:
            rd.readline("A, O digits for subsurface stresses", 2)
        if c["S"] >= 3:
            blocks = []
            isubs = int(rd.readline("ISUBS", 1)[0])
            while isubs >= 1:
                b = dict(isubs=isubs)
                if isubs in (2, 6):
                    b["ix"] = [int(v) for v in rd.readline("IXL, INC, IXH", 3)[:3]]
                    b["iy"] = [int(v) for v in rd.readline("IYL, INC, IYH", 3)[:3]]
                elif isubs in (3, 7):
                    r = rd.readline("NX, NY", 2)
                    b["ixs"] = [int(round(v)) for v in rd.read1darr("ix numbers", int(r[0]))]
                    b["iys"] = [int(round(v)) for v in rd.read1darr("iy numbers", int(r[1]))]
                if 1 <= isubs <= 3:
                    r = rd.readline("NZ, ZL, DZ", 3)
                    b["z"] = [float(r[1]) + k * float(r[2]) for k in range(int(r[0]))]
                    b["zparam"] = [int(r[0]), float(r[1]), float(r[2])]
                elif 5 <= isubs <= 7:
                    nz = int(rd.readline("NZ", 1)[0])
                    b["z"] = rd.read1darr("z coordinates", nz)
                elif isubs == 9:
                    r = rd.readline("NX, NY, NZ", 3)
                    b["x"] = rd.read1darr("x coordinates", int(r[0]))
                    b["y"] = rd.read1darr("y coordinates", int(r[1]))
                    b["z"] = rd.read1darr("z coordinates", int(r[2]))
                else:
                    if isubs not in (1, 5):
                        raise InpError("case %d: invalid ISUBS=%d" % (ncase, isubs))
                blocks.append(b)
                isubs = int(rd.readline("ISUBS", 1)[0])
            c["subs"] = blocks
        state["tang"] = c["T"]
        cases.append(c)
    return cases


def resolve_cases(cases):
    """Apply the inheritance rules: returns one complete description per case (what the solver sees)."""
    cur = dict(solver=dict(maxgs=999, maxin=20, maxnr=25, maxout=1, eps=1e-5), fric=(0.3, 0.3), roll=dict(chi=0.0, dq=1.0, veloc=1.0),
               mater=dict(poiss=(0.28, 0.28), gg=(82000.0, 82000.0)), potcon=None, geom=None, subs=[], G=0)
    out = []
    for c in cases:
        for k in ("solver", "fric", "roll", "mater", "potcon", "geom", "subs"):
            if k in c:
                cur[k] = dict(cur[k], **c[k]) if k == "solver" else c[k]
        if c["G"] != 1:
            cur["G"] = c["G"]
        full = dict(c)
        full.update(solver=dict(cur["solver"]), fric=cur["fric"], roll=dict(cur["roll"]), mater=cur["mater"], potcon=cur["potcon"],
                    geom=cur["geom"], subs=list(cur["subs"]), G_eff=cur["G"])
        out.append(full)
    return out


def run_inp(text, ire=1, icp=1, api=None, with_fields=True, subs_cases=None, before_case=None):
    """Run all cases of an .inp text through the cntc_* interface; returns one result dict per case.
    subs_cases: case numbers whose subsurface tables are fetched (None: all cases, when with_fields).
    before_case(n): called just before cntc_calculate of case n (1-based), after all its inputs are set."""
    if api is None:
        import contact_b200 as api
    cb = api
    K = cb.CNTC
    results = []
    cb.cntc_initialize(ire, 3)
    prev_geom = None
    for n, c in enumerate(resolve_cases(parse_inp(text)), 1):
        flags = [K["ic_pvtime"], K["ic_bound"], K["ic_tang"], K["ic_norm"], K["ic_force"], K["ic_iestim"], K["ic_return"]]
        vals = [c["P"], c["B"], c["T"], c["N"], c["F"], c["I"], c["R"]]
        cb.cntc_setflags(ire, icp, flags, vals)
        so = c["solver"]
        g = c["G_eff"]
        if g in (0, 1):
            cb.cntc_setsolverflags(ire, icp, 0, [so["maxgs"], so["maxin"], so["maxnr"], so["maxout"]], [so["eps"]])
        elif g in (2, 3):
            cb.cntc_setsolverflags(ire, icp, g, [so["maxgs"], so["maxin"], so["maxnr"], so["maxout"], so.get("inislp", 0)],
                                   [so["eps"], so["omegah"], so["omegas"], so.get("omgslp", 1.0)])
        elif g == 4:
            cb.cntc_setsolverflags(ire, icp, 4, [so["maxgs"], so["maxin"], so["maxnr"], so["maxout"], so.get("inislp", 0)],
                                   [so["eps"], so.get("omgslp", 1.0)])
        else:
            gd = so["gdsteady"]
            cb.cntc_setsolverflags(ire, icp, 5, [so["maxgs"], so["maxin"], so["maxnr"], so["maxout"], int(gd[2])],
                                   [so["eps"], gd[0], gd[1], gd[3], gd[4], gd[5], gd[6], gd[7]])
        m = c["mater"]
        cb.cntc_setmaterialparameters(ire, icp, 0, [m["poiss"][0], m["poiss"][1], m["gg"][0], m["gg"][1]])
        cb.cntc_setfrictionmethod(ire, icp, 0, list(c["fric"]))
        pc = c["potcon"]
        if pc is None:
            raise InpError("case %d: no potential contact defined" % n)
        if pc is not prev_geom or "geom" in c:
            if pc["ipotcn"] < 0:
                cb.cntc_sethertzcontact(ire, icp, pc["ipotcn"], [pc["mx"], pc["my"], pc["p1"], pc["p2"], pc["scale"]])
            else:
                cb.cntc_setpotcontact(ire, icp, pc["ipotcn"], [pc["mx"], pc["my"]] + pc["prm"])
                cb.cntc_setundeformeddistc(ire, icp, c["geom"]["ibase"], c["geom"]["prm"])
            prev_geom = pc
        if c["T"] in (2, 3):
            cb.cntc_setrollingstepsize(ire, icp, c["roll"]["chi"], c["roll"]["dq"])
        k = c["kin"]
        if c["N"] == 0:
            cb.cntc_setpenetration(ire, icp, k[0])
        else:
            cb.cntc_setnormalforce(ire, icp, k[0])
        if c["F"] == 0:
            cb.cntc_setcreepages(ire, icp, k[1], k[2], k[3])
        elif c["F"] == 1:
            cb.cntc_setcreepages(ire, icp, 0.0, k[2], k[3])
            cb.cntc_settangentialforces(ire, icp, k[1], 0.0)
        else:
            cb.cntc_setcreepages(ire, icp, 0.0, 0.0, k[3])
            cb.cntc_settangentialforces(ire, icp, k[1], k[2])
        if before_case is not None:
            before_case(n)
        t0 = time.perf_counter()
        ierr = cb.cntc_calculate(ire, icp)
        res = dict(case=n, ierror=ierr, wall_s=time.perf_counter() - t0)
        if ierr >= 0:
            its = cb.lowlevel.get_iterations(ire, icp)
            its["outer_history"] = cb.lowlevel.get_outer_history(ire, icp)
            el = cb.cntc_getelementdivision(ire, icp)
            fn, tx, ty, mz = cb.cntc_getcontactforces(ire, icp)
            res.update(its=its, ncon=int((el >= 1).sum()), nadh=int((el == 1).sum()), nslip=int((el == 2).sum()),
                       pen=cb.cntc_getpenetration(ire, icp), fn=fn, fx=tx, fy=ty, mz=mz,
                       pmax=cb.cntc_getmaximumpressure(ire, icp), creep=cb.cntc_getcreepages(ire, icp))
            if with_fields:
                pn, px, py = cb.cntc_gettractions(ire, icp)
                res.update(el=el, pn=pn, px=px, py=py)
            if c["S"] >= 1 and c["subs"]:
                for ib, b in enumerate(c["subs"], 1):
                    z = b["zparam"] if b["isubs"] <= 3 else b["z"]
                    if b["isubs"] == 9:
                        cb.subs_addblock(ire, icp, ib, 9, b["x"], b["y"], z)
                    elif b["isubs"] in (2, 6):
                        cb.subs_addblock(ire, icp, ib, b["isubs"], b["ix"], b["iy"], z)
                    elif b["isubs"] in (3, 7):
                        cb.subs_addblock(ire, icp, ib, b["isubs"], b["ixs"], b["iys"], z)
                    else:
                        cb.subs_addblock(ire, icp, ib, b["isubs"], [], [], z)
                t0 = time.perf_counter()
                res["subs_ierror"] = cb.subs_calculate(ire, icp)
                res["subs_wall_s"] = time.perf_counter() - t0
                if with_fields and res["subs_ierror"] == 0 and (subs_cases is None or n in subs_cases):
                    res["subs"] = [cb.subs_getresults(ire, icp, ib, list(range(1, 22))) for ib in range(1, len(c["subs"]) + 1)]
        else:
            res["message"] = cb.lib.last_error()
        results.append(res)
    cb.cntc_finalize(ire)
    return results
