"""Kernel-level entry points (cb200_*) for tests and benchmarks: batched influence product and NORM solve.

Device-buffer variants take torch CUDA tensors (PyTorch is used for device memory and streams only).
"""
import ctypes as C

import numpy as np

from .lib import load_library, last_error

ALLELM, ALLINT = -9, -8
SET_CS, SET_CV, SET_CSV, SET_MS = 0, 1, 2, 3


class CB200Error(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise CB200Error("libcontact_addon_b200 error %d: %s" % (rc, last_error()))
    return rc


def opt_fft_size(n):
    return load_library().cb200_opt_fft_size(n)


def num_launches():
    return load_library().cb200_num_launches()


def conv_prof(reset=True):
    """Cycle counters of CTA 0: dict(products, conv_cycles, kernel_cycles)."""
    import ctypes as C
    out = (C.c_ulonglong * 4)()
    _check(load_library().cb200_conv_prof(out, 1 if reset else 0))
    return dict(products=out[0], conv_cycles=out[1], kernel_cycles=out[2])


def work_counters(reset=True):
    """Work of all CTAs since the last reset: products, their nominal flops and algorithmic bytes at the transform sizes
    actually used, Gauss-Seidel row-sum units (ncon (ncon + 2 my) per sweep)."""
    out = (C.c_ulonglong * 4)()
    _check(load_library().cb200_work_counters(out, 1 if reset else 0))
    return dict(products=int(out[0]), product_flops=int(out[1]), product_bytes=int(out[2]), gs_units=int(out[3]))


SOLVER_SECTIONS = {4: "normcg set-up", 5: "first residual (product + passes)", 6: "product z = M r", 7: "pass A (z, dots)",
                   8: "pass B (v)", 9: "product q = A v", 10: "pass C (q, dots)", 11: "pass D (ps, release)",
                   12: "release bookkeeping / rescale", 13: "product d = A p - rhs", 14: "pass E (residual, entry, box)",
                   15: "snorm before normcg", 16: "snorm after normcg", 18: "waiting for twiddle tables", 19: "table loads (count)"}


def solver_prof(reset=True):
    """All 32 cycle counters of CTA 0 as {label: cycles} (development aid; see CB_T in csrc/device_core.cuh)."""
    import ctypes as C
    out = (C.c_ulonglong * 32)()
    _check(load_library().cb200_solver_prof(out, 1 if reset else 0))
    d = dict(products=out[0], conv_cycles=out[1], kernel_cycles=out[2])
    for k, name in SOLVER_SECTIONS.items():
        d[name] = out[k]
    return d


def steady_prof(reset=True):
    """Cycle counters of the SteadyGS element step: dict(steps, plstrc, reintegrate, update, calls)."""
    import ctypes as C
    out = (C.c_ulonglong * 8)()
    _check(load_library().cb200_steady_prof(out, 1 if reset else 0))
    return dict(steps=out[0], plstrc=out[1], reintegrate=out[2], update=out[3], calls=out[4], rowupdate=out[5], changes=out[6], rowchanges=out[7])


def batch_timing():
    """Wall-clock split (s) of the last cntc_calculate[_batch] call: dict(setup, coefficients, upload, kernel, output, total)."""
    import ctypes as C
    out = (C.c_double * 6)()
    _check(load_library().cb200_batch_timing(out))
    d = dict(zip(("setup", "coefficients", "upload", "kernel", "output", "total"), [round(v, 5) for v in out]))
    out2 = (C.c_double * 4)()
    _check(load_library().cb200_batch_timing_output(out2))
    d["output_split"] = dict(zip(("products", "downloads", "host", "free"), [round(v, 5) for v in out2]))
    return d


def gd_prof(reset=True):
    """Cycle counters of GDsteady (leader thread): dict(iterations, products, searchdir, ls_rows, ls_elements, step, trials, total)."""
    import ctypes as C
    out = (C.c_ulonglong * 12)()
    _check(load_library().cb200_gd_prof(out, 1 if reset else 0))
    return dict(iterations=out[0], products=out[1], searchdir=out[2], ls_rows=out[3], ls_elements=out[4], step=out[5], trials=out[6], total=out[7],
                sd_copy=out[8], sd_leader_rows=out[9], sd_wait=out[10], sd_fallbacks=out[11])


def num_sms():
    return _check(load_library().cb200_num_sms())


class CoefSet:
    """Influence coefficients for (grid, material, rolling step), cached inside the library."""

    def __init__(self, mx, my, dx, dy, gg=(82000.0, 82000.0), poiss=(0.28, 0.28), is_roll=False, chi=0.0, dq=1.0):
        self.mx, self.my, self.npot, self.dx, self.dy = mx, my, mx * my, dx, dy
        self.h = _check(load_library().cb200_coefset_create(mx, my, dx, dy, gg[0], gg[1], poiss[0], poiss[1],
                                                            int(is_roll), chi, dq))

    def plan(self):
        out = (C.c_int * 8)()
        _check(load_library().cb200_coefset_plan(self.h, out))
        return dict(zip(("Fx", "Fy", "C", "nchunk", "smem_bytes", "fits", "nsx", "nsy"), list(out)))

    def block(self, set_, ik, jk):
        """cf(-mx:mx-1, -my:my-1, ik, jk) as array [iy+my, ix+mx]."""
        out = np.zeros(4 * self.npot)
        _check(load_library().cb200_coefset_get_block(self.h, set_, ik, jk, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.reshape(2 * self.my, 2 * self.mx)

    def vecaijpj(self, p, el, iigs=ALLELM, ikarg=3, jkarg=3, set_=SET_CS, u=None):
        """HOST buffers through the C-ABI. p: (ncase, 3, npot); el: (ncase, npot) or None. Returns u (ncase,3,npot)."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        ncase = p.shape[0]
        u = np.zeros_like(p) if u is None else np.ascontiguousarray(u, dtype=np.float64)
        elp = None
        if el is not None:
            el = np.ascontiguousarray(el, dtype=np.int32)
            elp = el.ctypes.data_as(C.POINTER(C.c_int))
        _check(load_library().cb200_vecaijpj(self.h, set_, ncase, iigs, ikarg, jkarg,
                                             p.ctypes.data_as(C.POINTER(C.c_double)), elp,
                                             u.ctypes.data_as(C.POINTER(C.c_double))))
        return u

    def aijpj(self, ii, ik, p, el, jkarg=-3, set_=SET_CS):
        """gf3_AijPj: direct row sums for the 0-based elements ii of one case; p (3, npot), el (npot,). Returns (len(ii),)."""
        ii = np.ascontiguousarray(ii, dtype=np.int32); p = np.ascontiguousarray(p, dtype=np.float64)
        el = np.ascontiguousarray(el, dtype=np.int32)
        out = np.zeros(ii.size)
        _check(load_library().cb200_aijpj(self.h, set_, ik, jkarg, ii.size, ii.ctypes.data_as(C.POINTER(C.c_int)),
                                          p.ctypes.data_as(C.POINTER(C.c_double)), el.ctypes.data_as(C.POINTER(C.c_int)),
                                          out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def vecaijpj_dev(self, d_p, d_el, d_u, iigs=ALLELM, ikarg=3, jkarg=3, set_=SET_CS, stream=None):
        """DEVICE buffers (torch CUDA tensors: d_p, d_u float64 (ncase,3,npot); d_el int32 (ncase,npot) or None)."""
        ncase = d_p.shape[0]
        _check(load_library().cb200_vecaijpj_dev(self.h, set_, ncase, iigs, ikarg, jkarg, d_p.data_ptr(),
                                                 0 if d_el is None else d_el.data_ptr(), d_u.data_ptr(),
                                                 _stream_ptr(stream)))

    def snorm_batch_dev(self, d_hs, d_el, d_pn, d_un, d_scal, ic_norm, maxgs=999, maxin=20, eps=1e-5, stream=None):
        """Batched NORM solve on DEVICE tensors; d_scal (ncase,8): [pen, fn, itcg, itnorm, ncon, status, err, -]."""
        ncase = d_hs.shape[0]
        _check(load_library().cb200_snorm_batch_dev(self.h, ncase, ic_norm, maxgs, maxin, eps, d_hs.data_ptr(),
                                                    d_el.data_ptr(), d_pn.data_ptr(),
                                                    0 if d_un is None else d_un.data_ptr(), d_scal.data_ptr(),
                                                    _stream_ptr(stream)))


    def subsurf_batch(self, ps, zs, gg=(82000.0, 82000.0), poiss=(0.28, 0.28)):
        """HOST buffers: ps (ncase, 3, npot), depths zs -> table (ncase, nz, npot, 18)."""
        ps = np.ascontiguousarray(ps, dtype=np.float64)
        zs = np.ascontiguousarray(zs, dtype=np.float64)
        tbl = np.zeros((ps.shape[0], len(zs), self.npot, 18))
        dpp = C.POINTER(C.c_double)
        _check(load_library().cb200_subsurf_batch(self.h, ps.shape[0], len(zs), zs.ctypes.data_as(dpp), gg[0], gg[1],
                                                  poiss[0], poiss[1], ps.ctypes.data_as(dpp), tbl.ctypes.data_as(dpp)))
        return tbl

    def subsurf_batch_dev(self, d_ps, zs, d_table, gg=(82000.0, 82000.0), poiss=(0.28, 0.28), stream=None):
        zs = np.ascontiguousarray(zs, dtype=np.float64)
        _check(load_library().cb200_subsurf_batch_dev(self.h, d_ps.shape[0], len(zs), zs.ctypes.data_as(C.POINTER(C.c_double)),
                                                      gg[0], gg[1], poiss[0], poiss[1], d_ps.data_ptr(), d_table.data_ptr(),
                                                      _stream_ptr(stream)))

    def snorm_batch(self, hs, el, pn, un, scal, ic_norm, maxgs=999, maxin=20, eps=1e-5):
        """HOST buffers (numpy or pinned torch CPU tensors viewed as numpy), modified in place."""
        ncase = hs.shape[0]
        dpp, ipp = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _check(load_library().cb200_snorm_batch(self.h, ncase, ic_norm, maxgs, maxin, eps, hs.ctypes.data_as(dpp),
                                                el.ctypes.data_as(ipp), pn.ctypes.data_as(dpp),
                                                None if un is None else un.ctypes.data_as(dpp),
                                                scal.ctypes.data_as(dpp)))


def eldiv0(mx, my, dx, dy, gg, poiss, ibase, prmudf, ic_norm, fn, pen, h):
    """Initial element division / approach estimate of the library's eldiv0 restatement (host logic)."""
    h = np.ascontiguousarray(h, dtype=np.float64)
    prm = np.ascontiguousarray(prmudf, dtype=np.float64)
    el = np.zeros(mx * my, dtype=np.int32)
    pen_o = C.c_double(0.0)
    _check(load_library().cb200_eldiv0(mx, my, dx, dy, gg[0], gg[1], poiss[0], poiss[1], ibase,
                                       prm.ctypes.data_as(C.POINTER(C.c_double)), ic_norm, fn, pen,
                                       h.ctypes.data_as(C.POINTER(C.c_double)), el.ctypes.data_as(C.POINTER(C.c_int)),
                                       C.byref(pen_o)))
    return el, pen_o.value


def subsurf_points(mx, my, xc1, yc1, dx, dy, gg, poiss, ps, xyz):
    """ISUBS 9 direct evaluation; ps (3, npot), xyz (npoint, 3). Returns (npoint, 18)."""
    ps = np.ascontiguousarray(ps, dtype=np.float64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    tbl = np.zeros((xyz.shape[0], 18))
    dpp = C.POINTER(C.c_double)
    _check(load_library().cb200_subsurf_points(mx, my, xc1, yc1, dx, dy, gg[0], gg[1], poiss[0], poiss[1],
                                               ps.ctypes.data_as(dpp), xyz.shape[0], xyz.ctypes.data_as(dpp),
                                               tbl.ctypes.data_as(dpp)))
    return tbl


def get_iterations(ire, icp=1):
    out = (C.c_int * 8)(); log = (C.c_int * 64)()
    _check(load_library().cb200_get_iterations(ire, icp, out, 64, log))
    return dict(itnorm=out[0], itcg=out[1], ittang=out[2], itgs=out[3], ncon=out[4], nr_itcg=list(log[:out[5]]), itout=out[6],
                gd_trials=(out[7] if out[7] >= 0 else -out[7] - 1), gd_fallback=int(out[7] < 0))


def set_state(ire, icp, el, ps):
    """Test hook: the stored solution (el (npot,), ps (3, npot): x, y, n) that the next case of a sequence starts from."""
    el = np.ascontiguousarray(el, dtype=np.int32).ravel(); ps = np.ascontiguousarray(ps, dtype=np.float64).reshape(3, -1)
    _check(load_library().cb200_set_state(ire, icp, el.size, el.ctypes.data_as(C.POINTER(C.c_int)), ps.ctypes.data_as(C.POINTER(C.c_double))))


def set_devices(n, devs=None):
    """Devices cntc_calculate_batch spreads a batch over (1: the current device only); returns the number in use."""
    arr = None if devs is None else (C.c_int * len(devs))(*devs)
    r = load_library().cb200_set_devices(int(n), arr)
    if r < 0:
        _check(r)
    return r


def get_soutpt(ire, icp=1):
    """soutpt scalars (m_soutpt.f90:424-500): dict fn, fx, fy, mx, my, mz, elen, frpow, pmax"""
    o = (C.c_double * 9)()
    _check(load_library().cb200_get_soutpt(ire, icp, 9, o))
    return dict(zip(("fn", "fx", "fy", "mx", "my", "mz", "elen", "frpow", "pmax"), list(o)))


def get_deformed_distance(ire, icp, npot):
    o = np.zeros(npot)
    _check(load_library().cb200_get_deformed_distance(ire, icp, npot, o.ctypes.data_as(C.POINTER(C.c_double))))
    return o


def get_outer_history(ire, icp=1):
    """(dif, difid) of panprc's convergence test per outer iteration of the last solve"""
    d = (C.c_double * 16)(); e = (C.c_double * 16)()
    n = load_library().cb200_get_outer_history(ire, icp, 16, d, e)
    if n < 0:
        _check(n)
    return list(d[:n]), list(e[:n])


def snorm_kernel_ms():
    return load_library().cb200_snorm_kernel_ms()


def fp64_peak_tflops(reps=3):
    return load_library().cb200_fp64_peak_tflops(reps)


def _stream_ptr(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
