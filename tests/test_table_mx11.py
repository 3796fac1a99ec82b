"""testbank/test_table_mx11.ref_fx: 3220 Hertzian creepage cases on an 11x11 grid (T=3, G=0: NormCG + SteadyGS), the
reference's own threaded sweep (src/test_table.f90:105-292, one result element per thread) and its highest-precision
golden file for this path: Fx, Fy, Mz relative to mu*Fn (Mz also to cp) with 6 decimals, and the share of slip elements.

CPU: the oracle on a sample of the cases against the golden rows.  GPU: all 3220 cases through the cntc_* C-ABI
(cntc_sethertzcontact / cntc_setcreepages / cntc_calculate_batch / cntc_getcontactforces), and a sample of them against
the oracle element by element.  The fixture tests/golden/test_table_mx11.json is written by tests/golden/make_fixtures.py.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "test_table_mx11.json")))

# src/test_table.f90:33-42 -- REAL(4) literals in double arrays: ellip, vecmagn and the 11/7 scale carry single precision
F32 = lambda v: float(np.float32(v))
ELLIP = [F32(0.2), F32(0.5), 1.0, 2.0, 5.0]
AA = [4.0, 6.0, 9.0, 12.0, 20.0]
SPIN = [0.0, 0.5, 1.0, 2.0]
VECANGL = [-math.pi / 2, -math.pi / 3, -math.pi / 6, 0.0, math.pi / 6, math.pi / 3, math.pi / 2]
VECMAGN = [F32(v) for v in (0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.7, 0.9, 1.1, 1.4, 1.7, 2.0, 2.5, 3.0, 3.5, 4.0, 5.0, 6.0, 7.0,
                            8.0, 9.0, 10.0, 12.0)]
MX = 11
SCALE = float(np.float32(MX) / np.float32(max(1, MX - 4)))
G1, NU1, FSTAT, FN = 82000.0, 0.28, 0.3, 82000.0
NCASE = 5 * 4 * 7 * 23


def decompose(icase):
    """icase (1-based) -> (iell, ispin, ivang, ivmgn), 0-based; test_table.f90:214-227"""
    k = icase - 1
    return k // (23 * 7 * 4) % 5, k // (23 * 7) % 4, k // 23 % 7, k % 23


def creepages(icase, rho, cp):
    iell, ispin, ivang, ivmgn = decompose(icase)
    cmgn, cang = VECMAGN[ivmgn], VECANGL[ivang]
    return (cmgn * FSTAT * math.cos(cang) * cp[iell] / rho[iell], cmgn * FSTAT * math.sin(cang) * cp[iell] / rho[iell],
            FSTAT * SPIN[ispin] / rho[iell])


def oracle_case(iell, cksi, ceta, cphi):
    g = dict(mx=MX, my=MX, xl=0.0, yl=0.0, dx=1.0, dy=1.0, ibase=1, prmudf=[0.0] * 6)
    hz = dict(ipotcn=-3, aa=AA[iell], bb=AA[iell] / ELLIP[iell], scale=SCALE)
    return O.contac(g, [G1, G1], [NU1, NU1], tang=3, norm=1, force3=0, fn=FN, cksi=cksi, ceta=ceta, cphi=cphi, fstat=FSTAT, fkin=FSTAT,
                    maxgs=299, maxin=1, maxnr=30, maxout=1, eps=1e-6, gausei=0, hertz=hz)


def forces_from_fields(r, cp):
    """Fx, Fy, Mz as test_table prints them, from the oracle's tractions (m_soutpt.f90:424-450)."""
    gd = r["grid"]
    mx, my, dx, dy = gd["mx"], gd["my"], gd["dx"], gd["dy"]
    x = gd["xl"] + dx * (np.arange(mx) + 0.5)
    y = gd["yl"] + dy * (np.arange(my) + 0.5)
    X, Y = np.meshgrid(x, y)
    px, py, pn = (r["ps"][k].reshape(my, mx) for k in range(3))
    dxdy = dx * dy
    fn = dxdy * pn.sum()
    fx, fy = dxdy * px.sum() / (FSTAT * fn), dxdy * py.sum() / (FSTAT * fn)
    mz = (-dxdy * (px * Y).sum() + dxdy * (py * X).sum()) / (FSTAT * fn * cp)
    el = r["el"]
    return fx, fy, mz, 100.0 * (el == 2).sum() / max(1, (el >= 1).sum())


def hertz_constants_oracle():
    rho, cp = [], []
    for iell in range(5):
        r = oracle_case(iell, 0.0, 0.0, 0.0)
        assert r["ierror"] == 0
        rho.append(2.0 / (r["hz"]["a1"] + r["hz"]["b1"]))
        cp.append(math.sqrt(r["hz"]["aa"] * r["hz"]["bb"]))
    return rho, cp


def test_hertz_constants_match_the_golden_header():
    rho, cp = hertz_constants_oracle()
    for iell, h in enumerate(GOLD["hertz"]):
        assert abs(rho[iell] - h["rho"]) <= 0.5e-3 * h["rho"] + 0.5            # printed with 4 significant digits
        assert abs(cp[iell] - h["cp"]) <= 0.6e-3 * h["cp"]


@pytest.mark.parametrize("first", [1, 2])
def test_oracle_reproduces_golden_rows(first):
    """Every 61st case (53 cases per run, all ellipticities / spins / angles reached): Fx, Fy, Mz to the 6 printed decimals
    (the iteration stops at eps = 1e-6, the print rounds at 5e-7: 2e-6 allowed), slip share to the printed 0.1 %."""
    rho, cp = hertz_constants_oracle()
    worst = 0.0
    for icase in range(first, NCASE + 1, 61):
        iell = decompose(icase)[0]
        r = oracle_case(iell, *creepages(icase, rho, cp))
        assert r["ierror"] == 0, icase
        fx, fy, mz, slip = forces_from_fields(r, cp[iell])
        g = GOLD["rows"][icase - 1]
        d = max(abs(fx - g[0]), abs(fy - g[1]), abs(mz - g[2]))
        worst = max(worst, d)
        assert d < 2e-6, (icase, (fx, fy, mz), g)
        assert abs(slip - g[3]) < 0.051 + 1e-9, (icase, slip, g[3])
    assert worst < 2e-6


@pytest.mark.gpu
def test_all_3220_cases_through_the_cabi_batch():
    """The whole table through cntc_calculate_batch (result elements 1..920 per call, as a multibody code with one result
    element per wheel would), against the golden file; a sample of the cases against the oracle field by field."""
    import contact_b200 as cb
    nre = 920                                           # 3220 = 3.5 x 920: four calls
    for ire in range(1, nre + 1):
        cb.cntc_initialize(ire, 3)
        cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_norm"]], [3, 1])
        cb.cntc_setfrictionmethod(ire, 1, 0, [FSTAT, FSTAT])
        cb.cntc_setmaterialparameters(ire, 1, 0, [NU1, NU1, G1, G1])
        cb.cntc_setreferencevelocity(ire, 1, 10000.0)
        cb.cntc_setnormalforce(ire, 1, FN)
        cb.cntc_setsolverflags(ire, 1, 0, [299, 1, 30, 1], [1e-6])
    # Hertzian constants per ellipticity from the library itself (cntc_gethertzcontact), as test_table.f90:165-189
    rho, cp = [], []
    for iell in range(5):
        cb.cntc_sethertzcontact(1, 1, -3, [MX, MX, AA[iell], AA[iell] / ELLIP[iell], SCALE])
        cb.cntc_setcreepages(1, 1, 0.0, 0.0, 0.0)
        assert cb.cntc_calculate(1, 1) == 0
        h = cb.cntc_gethertzcontact(1, 1)
        rho.append(h[4]); cp.append(h[5])
    rho_o, cp_o = hertz_constants_oracle()
    assert np.allclose(rho, rho_o, rtol=1e-12) and np.allclose(cp, cp_o, rtol=1e-12)
    got = np.zeros((NCASE, 4))
    sample = {}
    for c0 in range(1, NCASE + 1, nre):
        ids = list(range(c0, min(NCASE, c0 + nre - 1) + 1))
        for k, icase in enumerate(ids):
            ire = k + 1
            iell = decompose(icase)[0]
            cb.cntc_sethertzcontact(ire, 1, -3, [MX, MX, AA[iell], AA[iell] / ELLIP[iell], SCALE])
            cb.cntc_setcreepages(ire, 1, *creepages(icase, rho, cp))
        ierr = cb.cntc_calculate_batch(list(range(1, len(ids) + 1)), 1)
        assert all(e == 0 for e in ierr), ([(ids[i], int(e)) for i, e in enumerate(ierr) if e != 0][:5], sum(1 for e in ierr if e != 0), cb.lib.last_error())
        for k, icase in enumerate(ids):
            ire = k + 1
            iell = decompose(icase)[0]
            fn, fx, fy, mz = cb.cntc_getcontactforces(ire, 1)
            dx, dy = cb.cntc_getgriddiscretization(ire, 1)
            carea, harea, sarea = cb.cntc_getcontactpatchareas(ire, 1)
            got[icase - 1] = (fx / (FSTAT * fn), fy / (FSTAT * fn), mz / (FSTAT * fn * cp[iell]),
                              100.0 * round(sarea / (dx * dy)) / max(1, round(carea / (dx * dy))))
            if icase % 161 == 7:
                sample[icase] = (cb.cntc_getelementdivision(ire, 1).ravel().copy(), [a.ravel().copy() for a in cb.cntc_gettractions(ire, 1)])
    for ire in range(1, nre + 1):
        cb.cntc_finalize(ire)
    gold = np.array(GOLD["rows"])
    d = np.abs(got[:, :3] - gold[:, :3]).max(axis=1)
    bad = np.nonzero(d >= 2e-6)[0]
    assert bad.size == 0, [(int(i) + 1, got[i].tolist(), gold[i].tolist()) for i in bad[:5]]
    assert np.abs(got[:, 3] - gold[:, 3]).max() < 0.051
    # field-level parity with the oracle on the sample: element division bit-exact, tractions 1e-9 relative
    assert len(sample) >= 15
    for icase, (el, (pn, px, py)) in sample.items():
        iell = decompose(icase)[0]
        r = oracle_case(iell, *creepages(icase, rho, cp))
        assert np.array_equal(el, r["el"]), icase
        scale = np.abs(r["ps"][2]).max()
        for a, b in ((px, r["ps"][0]), (py, r["ps"][1]), (pn, r["ps"][2])):
            assert np.abs(a - b).max() < 1e-9 * scale + 1e-6 * 1e-6 * scale, icase
