"""Host-side mirror of the reference's python_intfc for the hot path (module 3).

Function names, argument order/meaning, defaults and return shapes follow /root/reference/python_intfc/*.py
(e.g. cntc_gettractions returns pn, px, py reshaped to (my, mx): python_intfc/cntc_gettractions.py:24-41).
"""
import ctypes as C

import numpy as np

from .lib import load_library

__all__ = [
    "CNTC", "cntc_getmagicnumbers", "cntc_initlibrary", "cntc_initialize", "cntc_setglobalflags", "cntc_setflags",
    "cntc_getflags", "cntc_setmetadata", "cntc_setsolverflags", "cntc_setmaterialparameters", "cntc_settimestep",
    "cntc_setreferencevelocity", "cntc_setrollingstepsize", "cntc_setfrictionmethod", "cntc_sethertzcontact",
    "cntc_setpotcontact", "cntc_setpenetration", "cntc_setnormalforce", "cntc_setundeformeddistc",
    "cntc_setcreepages", "cntc_settangentialforces", "cntc_calculate", "cntc_calculate_batch",
    "cntc_getnumelements", "cntc_getgriddiscretization", "cntc_getpotcontact", "cntc_getpenetration",
    "cntc_getcreepages", "cntc_getcontactforces", "cntc_getcontactpatchareas", "cntc_getelementdivision",
    "cntc_getmaximumpressure", "cntc_getmaximumtraction", "cntc_getfielddata", "cntc_gettractions",
    "cntc_getmicroslip", "cntc_getdisplacements", "cntc_getcalculationtime", "cntc_getparameters",
    "cntc_getreferencevelocity", "cntc_gethertzcontact", "cntc_getsensitivities", "cntc_resetcalculationtime",
    "subs_addblock", "subs_calculate",
    "subs_getblocksize", "subs_getresults", "cntc_finalize", "cntc_finalizelast",
]

# magic numbers: /root/reference/src/caddon_flags.inc:13-178 (python_intfc/cntc_getmagicnumbers.py)
CNTC = dict(
    if_units=1933, un_cntc=1934, un_spck=1935, un_si=1936, un_imper=1937,
    ic_config=1967, ic_pvtime=1970, ic_bound=1971, ic_tang=1972, ic_norm=1973, ic_force=1974, ic_frclaw=1976,
    ic_discns=1977, ic_inflcf=1978, ic_mater=1979, ic_exrhs=1980, ic_xflow=1981, ic_heat=1982, ic_iestim=1983,
    ic_output=1984, ic_flow=1985, ic_return=1986, ic_matfil=1987, ic_sens=1988, ic_ifmeth=1989, ic_ifvari=1990,
    ic_sbsout=1991, ic_sbsfil=1992, ic_npomax=1993,
    if_idebug=2000, if_licdbg=2001, if_wrtinp=2002, if_openmp=2003, if_timers=2004, if_ncase=2005,
    fld_h=1, fld_mu=2, fld_px=3, fld_py=4, fld_pn=5, fld_ux=7, fld_uy=8, fld_un=9, fld_taucrt=11, fld_uplsx=12,
    fld_uplsy=13, fld_sx=15, fld_sy=16, fld_temp1=20, fld_temp2=21, fld_wx=22, fld_wy=23,
    err_allow=-12, err_search=-25, err_ftot=-26, err_norm=-27, err_tang=-28, err_tol=-29, err_icp=-31,
    err_profil=-32, err_frclaw=-33, err_discr=-34, err_input=-39, err_other=-99, err_broydn=-26,
)

ci, cd = C.c_int, C.c_double


def _ia(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _da(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def cntc_getmagicnumbers():
    return dict(CNTC)


def cntc_initlibrary(outdir=" ", expnam=" ", idebug=1):
    """python_intfc/cntc_initlibrary.py: returns (CNTC, ifcver, ierror)."""
    dll = load_library()
    ifcver, ierror, ioutput, one = ci(-1), ci(-1), ci(0), ci(1)
    dll.cntc_initializefirst(ifcver, ierror, ioutput, b" ", outdir.encode(), expnam.encode(), one,
                             ci(len(outdir)), ci(len(expnam)))
    return cntc_getmagicnumbers(), ifcver.value, ierror.value


def cntc_initialize(ire=1, imodul=3, outdir=" ", idebug=1):
    dll = load_library()
    ifcver, ierror = ci(-1), ci(-1)
    dll.cntc_initialize(ci(ire), ci(imodul), ifcver, ierror, outdir.encode(), ci(len(outdir)))
    return ifcver.value, ierror.value


def cntc_setglobalflags(params, values):
    p, v = _ia(params), _ia(values)
    load_library().cntc_setglobalflags(ci(len(p)), _ip(p), _ip(v))


def cntc_setflags(ire, icp, params, values):
    p, v = _ia(params), _ia(values)
    load_library().cntc_setflags(ci(ire), ci(icp), ci(len(p)), _ip(p), _ip(v))


def cntc_getflags(ire, icp, params):
    p = _ia(params)
    v = np.zeros(len(p), dtype=np.int32)
    load_library().cntc_getflags(ci(ire), ci(icp), ci(len(p)), _ip(p), _ip(v))
    return v


def cntc_setmetadata(ire, icp, params, values):
    p, v = _ia(params), _da(values)
    load_library().cntc_setmetadata(ci(ire), ci(icp), ci(len(p)), _ip(p), _dp(v))


def cntc_setsolverflags(ire, icp, gdigit, iparam, rparam):
    i, r = _ia(iparam), _da(rparam)
    load_library().cntc_setsolverflags(ci(ire), ci(icp), ci(gdigit), ci(len(i)), _ip(i), ci(len(r)), _dp(r))


def cntc_setmaterialparameters(ire, icp, m_digit, rparam):
    r = _da(rparam)
    load_library().cntc_setmaterialparameters(ci(ire), ci(icp), ci(m_digit), ci(len(r)), _dp(r))


def cntc_settimestep(ire, icp, dt):
    load_library().cntc_settimestep(ci(ire), ci(icp), cd(dt))


def cntc_setreferencevelocity(ire, icp, veloc):
    load_library().cntc_setreferencevelocity(ci(ire), ci(icp), cd(veloc))


def cntc_setrollingstepsize(ire, icp, chi, dq):
    load_library().cntc_setrollingstepsize(ci(ire), ci(icp), cd(chi), cd(dq))


def cntc_setfrictionmethod(ire, icp, imeth, params):
    r = _da(params)
    load_library().cntc_setfrictionmethod(ci(ire), ci(icp), ci(imeth), ci(len(r)), _dp(r))


def cntc_sethertzcontact(ire, icp, ipotcn, params):
    r = _da(params)
    load_library().cntc_sethertzcontact(ci(ire), ci(icp), ci(ipotcn), ci(len(r)), _dp(r))


def cntc_setpotcontact(ire, icp, ipotcn, params):
    r = _da(params)
    load_library().cntc_setpotcontact(ci(ire), ci(icp), ci(ipotcn), ci(len(r)), _dp(r))


def cntc_setpenetration(ire, icp, pen):
    load_library().cntc_setpenetration(ci(ire), ci(icp), cd(pen))


def cntc_setnormalforce(ire, icp, fn):
    load_library().cntc_setnormalforce(ci(ire), ci(icp), cd(fn))


def cntc_setundeformeddistc(ire, icp, ibase, prmudf):
    r = _da(prmudf).ravel()
    load_library().cntc_setundeformeddistc(ci(ire), ci(icp), ci(ibase), ci(len(r)), _dp(r))


def cntc_setcreepages(ire, icp, vx, vy, phi):
    load_library().cntc_setcreepages(ci(ire), ci(icp), cd(vx), cd(vy), cd(phi))


def cntc_settangentialforces(ire, icp, fx, fy):
    load_library().cntc_settangentialforces(ci(ire), ci(icp), cd(fx), cd(fy))


def cntc_calculate(ire=1, icp=1, idebug=1):
    ierr = ci(-999)
    load_library().cntc_calculate(ci(ire), ci(icp), ierr)
    return ierr.value


def cntc_calculate_batch(ires, icp=1):
    """B200 extension: solve many result elements in one launch per grid class; returns the ierror array."""
    r = _ia(ires)
    ierr = np.full(len(r), -999, dtype=np.int32)
    load_library().cntc_calculate_batch(ci(len(r)), _ip(r), ci(icp), _ip(ierr))
    return ierr


def cntc_getnumelements(ire=1, icp=1):
    mx, my = ci(0), ci(0)
    load_library().cntc_getnumelements(ci(ire), ci(icp), mx, my)
    return mx.value, my.value


def cntc_getgriddiscretization(ire=1, icp=1):
    dx, dy = cd(0), cd(0)
    load_library().cntc_getgriddiscretization(ci(ire), ci(icp), dx, dy)
    return dx.value, dy.value


def cntc_getpotcontact(ire=1, icp=1):
    v = np.zeros(6)
    load_library().cntc_getpotcontact(ci(ire), ci(icp), ci(6), _dp(v))
    return int(v[0]), int(v[1]), v[2], v[3], v[4], v[5]


def cntc_getpenetration(ire=1, icp=1):
    pen = cd(0)
    load_library().cntc_getpenetration(ci(ire), ci(icp), pen)
    return pen.value


def cntc_getcreepages(ire=1, icp=1):
    a, b, c = cd(0), cd(0), cd(0)
    load_library().cntc_getcreepages(ci(ire), ci(icp), a, b, c)
    return a.value, b.value, c.value


def cntc_getcontactforces(ire=1, icp=1):
    fn, tx, ty, mz = cd(0), cd(0), cd(0), cd(0)
    load_library().cntc_getcontactforces(ci(ire), ci(icp), fn, tx, ty, mz)
    return fn.value, tx.value, ty.value, mz.value


def cntc_getcontactpatchareas(ire=1, icp=1):
    a, b, c = cd(0), cd(0), cd(0)
    load_library().cntc_getcontactpatchareas(ci(ire), ci(icp), a, b, c)
    return a.value, b.value, c.value


def cntc_getelementdivision(ire=1, icp=1):
    mx, my = cntc_getnumelements(ire, icp)
    el = np.zeros(mx * my, dtype=np.int32)
    load_library().cntc_getelementdivision(ci(ire), ci(icp), ci(mx * my), _ip(el))
    return el.reshape(my, mx)


def cntc_getmaximumpressure(ire=1, icp=1):
    v = cd(0)
    load_library().cntc_getmaximumpressure(ci(ire), ci(icp), v)
    return v.value


def cntc_getmaximumtraction(ire=1, icp=1):
    v = cd(0)
    load_library().cntc_getmaximumtraction(ci(ire), ci(icp), v)
    return v.value


def cntc_getfielddata(ire, icp, ifld):
    mx, my = cntc_getnumelements(ire, icp)
    f = np.zeros(mx * my)
    load_library().cntc_getfielddata(ci(ire), ci(icp), ci(ifld), ci(mx * my), _dp(f))
    return f.reshape(my, mx)


def cntc_gettractions(ire=1, icp=1):
    mx, my = cntc_getnumelements(ire, icp)
    n = mx * my
    pn, px, py = np.zeros(n), np.zeros(n), np.zeros(n)
    load_library().cntc_gettractions(ci(ire), ci(icp), ci(n), _dp(pn), _dp(px), _dp(py))
    return pn.reshape(my, mx), px.reshape(my, mx), py.reshape(my, mx)


def cntc_getmicroslip(ire=1, icp=1):
    mx, my = cntc_getnumelements(ire, icp)
    n = mx * my
    sx, sy = np.zeros(n), np.zeros(n)
    load_library().cntc_getmicroslip(ci(ire), ci(icp), ci(n), _dp(sx), _dp(sy))
    return sx.reshape(my, mx), sy.reshape(my, mx)


def cntc_getdisplacements(ire=1, icp=1):
    mx, my = cntc_getnumelements(ire, icp)
    n = mx * my
    un, ux, uy = np.zeros(n), np.zeros(n), np.zeros(n)
    load_library().cntc_getdisplacements(ci(ire), ci(icp), ci(n), _dp(un), _dp(ux), _dp(uy))
    return un.reshape(my, mx), ux.reshape(my, mx), uy.reshape(my, mx)


def cntc_getcalculationtime(ire=1, icp=1):
    a, b = cd(0), cd(0)
    load_library().cntc_getcalculationtime(ci(ire), ci(icp), a, b)
    return a.value, b.value


def cntc_getparameters(ire, icp, itask, lenarr=22):
    """python_intfc/cntc_getparameters.py: itask 1 kinematic constants, 2 material, 3 friction."""
    v = np.zeros(lenarr)
    load_library().cntc_getparameters(ci(ire), ci(icp), ci(itask), ci(lenarr), _dp(v))
    return v


def cntc_getreferencevelocity(ire=1, icp=1):
    v = cd(0)
    load_library().cntc_getreferencevelocity(ci(ire), ci(icp), v)
    return v.value


def cntc_gethertzcontact(ire=1, icp=1):
    """python_intfc/cntc_gethertzcontact.py: [a1, b1, aa, bb, rho, cp, scale, bneg, bpos, aob]."""
    v = np.zeros(10)
    load_library().cntc_gethertzcontact(ci(ire), ci(icp), ci(10), _dp(v))
    return v


def cntc_getsensitivities(ire=1, icp=1, lenout=4, lenin=4):
    """python_intfc/cntc_getsensitivities.py: sens[iout, iin], outputs fn, fx, fy, mz; inputs pen, cksi, ceta, cphi."""
    s = np.zeros(lenout * lenin)
    load_library().cntc_getsensitivities(ci(ire), ci(icp), ci(lenout), ci(lenin), _dp(s))
    return s.reshape(lenin, lenout).T


def cntc_resetcalculationtime(ire=1, icp=1):
    load_library().cntc_resetcalculationtime(ci(ire), ci(icp))


def subs_addblock(ire, icp, iblk, isubs, xparam, yparam, zparam):
    x, y, z = _da(xparam), _da(yparam), _da(zparam)
    load_library().subs_addblock(ci(ire), ci(icp), ci(iblk), ci(isubs), ci(len(x)), ci(len(y)), ci(len(z)),
                                 _dp(x), _dp(y), _dp(z))


def subs_calculate(ire=1, icp=1, idebug=1):
    ierr = ci(-999)
    load_library().subs_calculate(ci(ire), ci(icp), ierr)
    return ierr.value


def subs_getblocksize(ire=1, icp=1, iblk=1):
    nx, ny, nz = ci(0), ci(0), ci(0)
    load_library().subs_getblocksize(ci(ire), ci(icp), ci(iblk), nx, ny, nz)
    return nx.value, ny.value, nz.value


def subs_getresults(ire, icp, iblk, icol):
    nx, ny, nz = subs_getblocksize(ire, icp, iblk)
    cols = _ia(icol)
    n = nx * ny * nz
    tbl = np.zeros(n * len(cols))
    load_library().subs_getresults(ci(ire), ci(icp), ci(iblk), ci(n), ci(len(cols)), _ip(cols), _dp(tbl))
    return tbl.reshape(len(cols), n).T if n else tbl.reshape(0, len(cols))


def cntc_finalize(ire=1):
    load_library().cntc_finalize(ci(ire))


def cntc_finalizelast():
    load_library().cntc_finalizelast()
