"""ctypes loader for the ORACLE (test infrastructure, not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Builds oracle/_build/libcontact_oracle.so on demand with the committed Makefile.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcontact_oracle.so")
_lib = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.co_opt_fft_size.restype = C.c_int
        _lib.co_ctx_new.restype = C.c_void_p
    return _lib


def _d(a):
    return a.ctypes.data_as(c_dbl_p)


def _i(a):
    return a.ctypes.data_as(c_int_p)


def opt_fft_size(n):
    return int(lib().co_opt_fft_size(C.c_int(n)))


def fft2_r2c(a):
    """a: (n2, n1) real, x (n1) fastest. Returns (n2, n1//2+1) complex, unscaled forward transform."""
    L = lib()
    a = np.ascontiguousarray(a, dtype=np.float64)
    n2, n1 = a.shape
    out = np.zeros((n2, n1 // 2 + 1), dtype=np.complex128)
    pc = (C.c_char * 4096)()
    L.co_fft2_r2c(C.byref(pc), C.c_int(n1), C.c_int(n2), _d(a), out.ctypes.data_as(C.c_void_p))
    L.co_plancache_clear(C.byref(pc))
    return out


def fft2_c2r(A, n1, scale=1.0):
    L = lib()
    A = np.array(A, dtype=np.complex128, order="C")
    n2 = A.shape[0]
    out = np.zeros((n2, n1), dtype=np.float64)
    pc = (C.c_char * 4096)()
    L.co_fft2_c2r(C.byref(pc), C.c_int(n1), C.c_int(n2), A.ctypes.data_as(C.c_void_p), _d(out), C.c_double(scale))
    L.co_plancache_clear(C.byref(pc))
    return out


class Inflcf(C.Structure):
    _fields_ = [("cf_mx", C.c_int), ("cf_my", C.c_int), ("cf", c_dbl_p),
                ("dx", C.c_double), ("dy", C.c_double), ("dq", C.c_double),
                ("nt_cpl", C.c_int), ("ga", C.c_double), ("ga_inv", C.c_double),
                ("use_3bl", C.c_int), ("use_flxz", C.c_int), ("flx_3bl", C.c_double), ("flx_z", C.c_double),
                ("fft_ok", C.c_int * 9), ("fft_mx", C.c_int), ("fft_my", C.c_int),
                ("fft_cf", C.c_void_p * 9), ("fft_len", C.c_long), ("n_cfft", C.c_long)]

    def block(self, ik, jk):
        """numpy view of cf(-mx:mx-1, -my:my-1, ik, jk) as array [iy+my, ix+mx]."""
        n = 4 * self.cf_mx * self.cf_my
        arr = np.ctypeslib.as_array(self.cf, shape=(9 * n,))
        b = arr[((jk - 1) * 3 + (ik - 1)) * n:((jk - 1) * 3 + ik) * n]
        return b.reshape(2 * self.cf_my, 2 * self.cf_mx)


class Mater(C.Structure):
    _fields_ = [("gg", C.c_double * 2), ("poiss", C.c_double * 2),
                ("ga", C.c_double), ("nu", C.c_double), ("ak", C.c_double)]


class Eldiv(C.Structure):
    _fields_ = [("mx", C.c_int), ("my", C.c_int), ("el", c_int_p), ("row1st", c_int_p), ("rowlst", c_int_p),
                ("ixmin", C.c_int), ("ixmax", C.c_int), ("iymin", C.c_int), ("iymax", C.c_int)]


def mater(gg=(82000.0, 82000.0), poiss=(0.28, 0.28)):
    m = Mater()
    m.gg[0], m.gg[1] = gg
    m.poiss[0], m.poiss[1] = poiss
    lib().co_combin_mater(C.byref(m))
    return m


def sgencr(m, mx, my, dx, dy, is_roll=False, chi=0.0, dq=1.0):
    """Returns (cs, cv, csv, ms) Inflcf structures (caller keeps them alive; free with inflcf_free)."""
    cs, cv, csv, ms = Inflcf(), Inflcf(), Inflcf(), Inflcf()
    lib().co_sgencr(C.byref(m), C.c_int(mx), C.c_int(my), C.c_double(dx), C.c_double(dy), C.c_int(int(is_roll)),
                    C.c_double(chi), C.c_double(dq), C.byref(cs), C.byref(cv), C.byref(csv), C.byref(ms))
    return cs, cv, csv, ms


def inflcf_free(*cs):
    for c in cs:
        lib().co_inflcf_free(C.byref(c))


class Ctx:
    def __init__(self, fullbox=False):
        self.p = C.c_void_p(lib().co_ctx_new())
        # co_ctx layout: plancache (int + pad + 24 ptrs), stats (2 long, 2 double), int fullbox
        self._set_fullbox(fullbox)

    def _set_fullbox(self, v):
        off = 8 + 8 * 24 + 32
        C.c_int.from_address(self.p.value + off).value = int(v)

    def stats(self):
        off = 8 + 8 * 24
        n_prod = C.c_long.from_address(self.p.value + off).value
        n_rowsum = C.c_long.from_address(self.p.value + off + 8).value
        ab = C.c_double.from_address(self.p.value + off + 16).value
        af = C.c_double.from_address(self.p.value + off + 24).value
        return dict(n_prod=n_prod, n_rowsum=n_rowsum, alg_bytes=ab, alg_flops=af)

    def close(self):
        if self.p:
            lib().co_ctx_free(self.p)
            self.p = None

    def __del__(self):
        self.close()


class EldivBuf:
    """Owns an element division; el is a numpy int32 array of (npot,)."""

    def __init__(self, mx, my, el=None):
        self.mx, self.my = mx, my
        self.el = np.zeros(mx * my, dtype=np.int32) if el is None else np.ascontiguousarray(el, dtype=np.int32).copy()
        self.row1st = np.zeros(my, dtype=np.int32)
        self.rowlst = np.zeros(my, dtype=np.int32)
        self.s = Eldiv(mx, my, _i(self.el), _i(self.row1st), _i(self.rowlst), mx, 1, my, 1)
        self.areas()

    def areas(self):
        lib().co_areas(C.byref(self.s))


def vecaijpj(ctx, igs, iigs, u, ikarg, p, jkarg, c, pel=None):
    """u, p: (3, npot) float64 arrays (column ik-1 = direction ik)."""
    pel = igs if pel is None else pel
    lib().co_vecaijpj(ctx.p, C.byref(igs.s), C.c_int(iigs), _d(u), C.c_int(ikarg), _d(p), C.byref(pel.s),
                      C.c_int(jkarg), C.byref(c))


def vecaijpj_direct(igs, iigs, u, ikarg, p, jkarg, c, pel=None):
    pel = igs if pel is None else pel
    lib().co_vecaijpj_direct(C.byref(igs.s), C.c_int(iigs), _d(u), C.c_int(ikarg), _d(p), C.byref(pel.s),
                             C.c_int(jkarg), C.byref(c))


def fft_makeprec(ctx, ik, c, jk, m):
    lib().co_fft_makeprec(ctx.p, C.c_int(ik), C.byref(c), C.c_int(jk), C.byref(m))


def norm_case(mx, my, xl, yl, dx, dy, gg, poiss, ibase, prmudf, ic_norm, pen=0.0, fn=0.0,
              maxgs=999, maxin=20, eps=1e-5, fullbox=False, nn=0):
    """One module-3 normal-contact case (T=0, IPOTCN=1, I=0). Returns dict."""
    npot = mx * my
    prm = np.ascontiguousarray(prmudf, dtype=np.float64)
    el = np.zeros(npot, dtype=np.int32)
    pn = np.zeros(npot, dtype=np.float64)
    pen_o, fn_o = C.c_double(), C.c_double()
    itcg, itnorm = C.c_int(), C.c_int()
    stats = np.zeros(4)
    L = lib()
    L.co_norm_case.restype = C.c_int
    ierr = L.co_norm_case(C.c_int(mx), C.c_int(my), C.c_double(xl), C.c_double(yl), C.c_double(dx), C.c_double(dy),
                          C.c_double(gg[0]), C.c_double(gg[1]), C.c_double(poiss[0]), C.c_double(poiss[1]),
                          C.c_int(ibase), C.c_int(nn), _d(prm), C.c_int(ic_norm), C.c_double(pen), C.c_double(fn),
                          C.c_int(maxgs), C.c_int(maxin), C.c_double(eps), C.c_int(int(fullbox)),
                          _i(el), _d(pn), C.byref(pen_o), C.byref(fn_o), C.byref(itcg), C.byref(itnorm), _d(stats))
    return dict(ierror=ierr, el=el, pn=pn, pen=pen_o.value, fn=fn_o.value, itcg=itcg.value, itnorm=itnorm.value,
                n_prod=int(stats[0]), n_rowsum=int(stats[1]), alg_bytes=stats[2], alg_flops=stats[3])


def norm_batch(g, gg, poiss, ic_norm, loads, maxgs=999, maxin=20, eps=1e-5, nthreads=1, fullbox=False, nn=0):
    """Batch of normal-contact cases on one grid (dict g: mx,my,xl,yl,dx,dy,ibase,prmudf), one case per thread."""
    loads = np.ascontiguousarray(loads, dtype=np.float64)
    ncase, npot = len(loads), g["mx"] * g["my"]
    prm = np.ascontiguousarray(g["prmudf"], dtype=np.float64)
    el = np.zeros((ncase, npot), dtype=np.int32)
    pn = np.zeros((ncase, npot))
    scal = np.zeros((ncase, 4))
    L = lib()
    L.co_norm_batch.restype = C.c_int
    nfail = L.co_norm_batch(C.c_int(ncase), C.c_int(nthreads), C.c_int(g["mx"]), C.c_int(g["my"]), C.c_double(g["xl"]),
                            C.c_double(g["yl"]), C.c_double(g["dx"]), C.c_double(g["dy"]), C.c_double(gg[0]),
                            C.c_double(gg[1]), C.c_double(poiss[0]), C.c_double(poiss[1]), C.c_int(g["ibase"]),
                            C.c_int(nn), _d(prm), C.c_int(ic_norm), _d(loads), C.c_int(maxgs), C.c_int(maxin),
                            C.c_double(eps), C.c_int(int(fullbox)), _i(el), _d(pn), _d(scal))
    return dict(nfail=nfail, el=el, pn=pn, pen=scal[:, 0], fn=scal[:, 1], itcg=scal[:, 2].astype(int),
                n_prod=scal[:, 3].astype(int))


def subsurf_points(mx, my, xl, yl, dx, dy, gg, poiss, ps, xs, ys, zs):
    """ISUBS=9: direct evaluation (sstres) in the points xs x ys x zs (x fastest). ps: (3, npot).
    Returns (npoint, 21) table: x, y, z, ux, uy, uz, sighyd, sigvm, sigtr, sigma1..3, sigma(3,3) column-major."""
    L = lib()
    npot = mx * my
    x = np.zeros(npot); y = np.zeros(npot)
    L.co_grid_coords(C.c_int(mx), C.c_int(my), C.c_double(xl), C.c_double(yl), C.c_double(dx), C.c_double(dy), _d(x), _d(y))
    ps = np.ascontiguousarray(ps, dtype=np.float64)
    g2 = (C.c_double * 2)(*gg); p2 = (C.c_double * 2)(*poiss)
    rows = []
    out = np.zeros(18)
    for zz in zs:
        for yy in ys:
            for xx in xs:
                xw = (C.c_double * 3)(xx, yy, zz)
                L.co_sstres_point(C.c_int(mx), C.c_int(my), C.c_double(dx), C.c_double(dy), _d(x), _d(y), g2, p2, _d(ps), xw, _d(out))
                rows.append([xx, yy, zz] + list(out))
    return np.array(rows)


def subsurf_block(mx, my, dx, dy, gg, poiss, el, ps, zs, use_fft=True):
    """ISUBS=1/5 block: all elements x depths zs. Returns (nz, npot, 18): columns 4..21 of the reference's table."""
    L = lib()
    npot = mx * my
    ps = np.ascontiguousarray(ps, dtype=np.float64)
    el = np.ascontiguousarray(el, dtype=np.int32)
    zs = np.ascontiguousarray(zs, dtype=np.float64)
    g2 = (C.c_double * 2)(*gg); p2 = (C.c_double * 2)(*poiss)
    tbl = np.zeros((len(zs), npot, 18))
    ctx = Ctx(fullbox=True)
    L.co_subsurf_block_fft(ctx.p, C.c_int(mx), C.c_int(my), C.c_double(dx), C.c_double(dy), g2, p2, _i(el), _d(ps),
                           C.c_int(len(zs)), _d(zs), C.c_int(int(use_fft)), _d(tbl))
    return tbl


class Case(C.Structure):
    _fields_ = [("mx", C.c_int), ("my", C.c_int), ("xl", C.c_double), ("yl", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
                ("gg", C.c_double * 2), ("poiss", C.c_double * 2), ("ibase", C.c_int), ("nn", C.c_int), ("prmudf", c_dbl_p),
                ("tang", C.c_int), ("norm", C.c_int), ("force3", C.c_int),
                ("pen", C.c_double), ("fn", C.c_double), ("cksi", C.c_double), ("ceta", C.c_double), ("cphi", C.c_double),
                ("fxrel", C.c_double), ("fyrel", C.c_double), ("fstat", C.c_double), ("fkin", C.c_double),
                ("maxgs", C.c_int), ("maxin", C.c_int), ("maxnr", C.c_int), ("maxout", C.c_int), ("eps", C.c_double),
                ("fullbox", C.c_int), ("chi", C.c_double), ("dq", C.c_double), ("facphi", C.c_double), ("gausei", C.c_int),
                ("omegah", C.c_double), ("omegas", C.c_double),
                ("el", c_int_p), ("ps", c_dbl_p), ("ss", c_dbl_p),
                ("pen_out", C.c_double), ("fn_out", C.c_double), ("fx_out", C.c_double), ("fy_out", C.c_double),
                ("itnorm", C.c_int), ("ittang", C.c_int), ("itcg_norm", C.c_int), ("itgs_tang", C.c_int),
                ("nr_n", C.c_int), ("nr_itcg", C.c_int * 64), ("nr_cksi", C.c_double * 64), ("nr_ceta", C.c_double * 64),
                ("nr_fx", C.c_double * 64), ("nr_fy", C.c_double * 64), ("n_prod", C.c_long),
                ("iestim", C.c_int), ("el_in", c_int_p), ("ps_in", c_dbl_p), ("pv_in", c_dbl_p),
                ("ipotcn", C.c_int), ("hz_a1", C.c_double), ("hz_b1", C.c_double), ("hz_aa", C.c_double),
                ("hz_bb", C.c_double), ("hz_scale", C.c_double), ("itout", C.c_int),
                ("gd", C.c_double * 8), ("gd_fallback", C.c_int),
                ("pan_dif", C.c_double * 16), ("pan_difid", C.c_double * 16)]


def contac(g, gg, poiss, tang=0, norm=0, force3=0, pen=0.0, fn=0.0, cksi=0.0, ceta=0.0, cphi=0.0, fxrel=0.0, fyrel=0.0,
           fstat=0.3, fkin=0.3, maxgs=999, maxin=20, maxnr=25, maxout=1, eps=1e-5, fullbox=False, nn=0, chi=0.0, dq=1.0,
           facphi=0.0, gausei=0, omegah=0.9, omegas=0.9, iestim=0, el_in=None, ps_in=None, pv_in=None, hertz=None,
           gd=(1.0, 0.05, 1, 2.0, -1.0, 1.0, 2.6, 1.0)):
    """One module-3 case (T = 0/1/3) through the oracle's contac/panprc. g: dict mx,my,xl,yl,dx,dy,ibase,prmudf."""
    npot = g["mx"] * g["my"]
    prm = np.ascontiguousarray(g["prmudf"], dtype=np.float64)
    el = np.zeros(npot, dtype=np.int32); ps = np.zeros((3, npot)); ss = np.zeros((3, npot))
    c = Case()
    c.mx, c.my, c.xl, c.yl, c.dx, c.dy = g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"]
    c.gg[0], c.gg[1], c.poiss[0], c.poiss[1] = gg[0], gg[1], poiss[0], poiss[1]
    c.ibase, c.nn, c.prmudf = g["ibase"], nn, _d(prm)
    c.tang, c.norm, c.force3 = tang, norm, force3
    c.pen, c.fn, c.cksi, c.ceta, c.cphi, c.fxrel, c.fyrel, c.fstat, c.fkin = pen, fn, cksi, ceta, cphi, fxrel, fyrel, fstat, fkin
    c.maxgs, c.maxin, c.maxnr, c.maxout, c.eps, c.fullbox = maxgs, maxin, maxnr, maxout, eps, int(fullbox)
    c.chi, c.dq, c.facphi, c.gausei = chi, dq, facphi, gausei
    c.omegah, c.omegas = omegah, omegas
    for i in range(8):
        c.gd[i] = float(gd[i])
    keep = []
    if el_in is not None and ps_in is not None:
        e_ = np.ascontiguousarray(el_in, dtype=np.int32); p_ = np.ascontiguousarray(ps_in, dtype=np.float64)
        keep += [e_, p_]
        c.iestim, c.el_in, c.ps_in = iestim, _i(e_), _d(p_)
    if pv_in is not None:
        v_ = np.ascontiguousarray(pv_in, dtype=np.float64)
        keep.append(v_)
        c.pv_in = _d(v_)
    if hertz is not None:          # dict(ipotcn, a1, b1, aa, bb, scale): grid and geometry from the Hertz solution
        c.ipotcn = hertz["ipotcn"]
        c.hz_a1, c.hz_b1, c.hz_aa, c.hz_bb = hertz.get("a1", 0.0), hertz.get("b1", 0.0), hertz.get("aa", 0.0), hertz.get("bb", 0.0)
        c.hz_scale = hertz.get("scale", 1.0)
    c.el, c.ps, c.ss = _i(el), _d(ps), _d(ss)
    L = lib()
    L.co_contac.restype = C.c_int
    ierr = L.co_contac(C.byref(c))
    n = c.nr_n
    return dict(ierror=ierr, el=el, ps=ps, ss=ss, pen=c.pen_out, fn=c.fn_out, fx=c.fx_out, fy=c.fy_out, cksi=c.cksi, ceta=c.ceta,
                grid=dict(mx=c.mx, my=c.my, xl=c.xl, yl=c.yl, dx=c.dx, dy=c.dy), hz=dict(a1=c.hz_a1, b1=c.hz_b1, aa=c.hz_aa, bb=c.hz_bb),
                itnorm=c.itnorm, ittang=c.ittang, itout=c.itout, itcg_norm=c.itcg_norm, itgs_tang=c.itgs_tang,
                nr_itcg=list(c.nr_itcg[:n]), nr_cksi=list(c.nr_cksi[:n]), nr_ceta=list(c.nr_ceta[:n]), nr_fx=list(c.nr_fx[:n]),
                nr_fy=list(c.nr_fy[:n]), n_prod=c.n_prod, gd_fallback=c.gd_fallback,
                pan_dif=list(c.pan_dif[:min(16, c.itout)]), pan_difid=list(c.pan_difid[:min(16, c.itout)]))
