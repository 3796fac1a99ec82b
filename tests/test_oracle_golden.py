"""The CPU oracle pinned against the reference's golden files (printed precision, exact pictures and counts)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))


def test_cattaneo_case2_normal_problem():
    """examples/cattaneo.ref_out:82-84 (197 -> 177 elements, ItCG 6), :44 approach/pmax, contact picture."""
    c = cases.CATTANEO2
    r = O.norm_case(c["mx"], c["my"], c["xl"], c["yl"], c["dx"], c["dy"], c["gg"], c["poiss"], 1, c["prmudf"], 1,
                    fn=c["fn"], maxgs=c["maxgs"], maxin=c["maxin"], eps=c["eps"])
    assert r["ierror"] == 0 and r["itnorm"] == 1
    assert int((r["el"] > 0).sum()) == 177
    assert r["itcg"] == 6
    assert "%.3E" % r["pen"] == "1.998E-02"
    assert "%.3f" % r["pn"].max() == "4.393"
    pics = json.load(open(os.path.join(HERE, "golden", "cattaneo_pictures.json")))["pictures"]
    gold = np.array([[1 if ch in "*S|" else 0 for ch in row] for row in pics[0]], dtype=np.int32)
    assert np.array_equal(gold.ravel(), (r["el"] > 0).astype(np.int32))
    # eldiv0's first guess: "Norm: size of Contact, Exterior : 197 164"
    m = O.mater(gg=c["gg"], poiss=c["poiss"])
    assert abs(m.ga - 200.0) < 1e-12 and abs(m.nu - 0.42) < 1e-12 and abs(m.ak) < 1e-15


@pytest.mark.parametrize("name,mx,my,dx", [("norm_problm_1p", 71, 81, 0.1), ("norm_problm_2p", 143, 161, 0.05)])
def test_perfc_norm_problem_iteration_counts(mbench, name, mx, my, dx):
    """perfc_test/get_times.ref_out:7-8: ncon and ItCG of the NormCG perf problems."""
    gold = json.load(open(os.path.join(HERE, "golden", "get_times.json")))[name]
    r = O.norm_case(mx, my, -3.55, -6.15, dx, dx, (82000.0, 82000.0), (0.28, 0.28), 2, mbench["prmudf"], 0,
                    pen=mbench["pen"], maxgs=1000, maxin=100, eps=1e-7, nn=mbench["nn"])
    assert int((r["el"] > 0).sum()) == gold["ncon"]
    assert r["itcg"] == gold["itcg"]


def test_bbox_and_fullbox_products_agree():
    """The reference crops AllInt products to the contact bounding box (m_aijpj.f90:774-793); the GPU path always uses
    the full grid.  Same flags, pressures equal to rounding."""
    c = cases.CATTANEO2
    kw = dict(fn=c["fn"], maxgs=c["maxgs"], maxin=c["maxin"], eps=c["eps"])
    a = O.norm_case(c["mx"], c["my"], c["xl"], c["yl"], c["dx"], c["dy"], c["gg"], c["poiss"], 1, c["prmudf"], 1, **kw)
    b = O.norm_case(c["mx"], c["my"], c["xl"], c["yl"], c["dx"], c["dy"], c["gg"], c["poiss"], 1, c["prmudf"], 1,
                    fullbox=True, **kw)
    assert np.array_equal(a["el"], b["el"]) and a["itcg"] == b["itcg"]
    assert np.abs(a["pn"] - b["pn"]).max() < 1e-12 * a["pn"].max()


@pytest.mark.parametrize("mx,my", [(19, 19), (33, 27), (12, 7), (1, 9), (8, 1)])
def test_fft_product_equals_direct_sum(mx, my):
    """VecAijPj (FFT) against its twin AijPj (direct row sums): the high-precision pin of the product."""
    rng = np.random.default_rng(5)
    for gg, poiss in (((82000.0, 82000.0), (0.28, 0.28)), ((0.5, 1e5), (0.0, 0.0))):
        m = O.mater(gg=gg, poiss=poiss)
        cs, cv, csv, ms = O.sgencr(m, mx, my, 0.2, 0.15)
        el = (rng.random(mx * my) < 0.7).astype(np.int32)
        if el.sum() == 0:
            el[0] = 1
        p = rng.standard_normal((3, mx * my)) * el
        igs = O.EldivBuf(mx, my, el)
        u = np.zeros((3, mx * my)); ud = np.zeros((3, mx * my))
        O.vecaijpj(O.Ctx(), igs, -8, u, -3, p, -3, cs)
        O.vecaijpj_direct(igs, -8, ud, -3, p, -3, cs)
        assert np.abs(u - ud).max() < 1e-13 * np.abs(ud).max()
        O.inflcf_free(cs, cv, csv, ms)


def test_reference_test_fft_stencil():
    """The reference's (dead) self-check m_snorm.f90:1192-1230: 3x2 grid, stencil in block (1,1)."""
    mx, my = 3, 2
    m = O.mater()
    cs, cv, csv, ms = O.sgencr(m, mx, my, 1.0, 1.0)
    for ik in (1, 2, 3):
        for jk in (1, 2, 3):
            cs.block(ik, jk)[:] = 0.0
    b = cs.block(1, 1)                     # cf(ix, iy) at [iy+my, ix+mx]
    b[0 + my, 0 + mx] = 2.0
    b[0 + my, 1 + mx] = -1.0
    b[1 + my, 0 + mx] = 1.0
    cs.ga = 1.0; cs.ga_inv = 1.0; cs.nt_cpl = 0
    p = np.zeros((3, 6)); p[0, 1] = 1.0    # unit traction px at element (2,1)
    igs = O.EldivBuf(mx, my, np.ones(6, np.int32))
    u = np.zeros((3, 6))
    O.vecaijpj(O.Ctx(), igs, -9, u, 1, p, 1, cs)
    assert np.allclose(u[0].reshape(2, 3), [[0.0, 2.0, -1.0], [0.0, 1.0, 0.0]], atol=1e-14)
    O.inflcf_free(cs, cv, csv, ms)
