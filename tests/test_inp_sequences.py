"""The reference's own example inputs as sequences of cases: .inp reader + the CPU oracle against the printed results of
examples/spence35.ref_out (35 stages, P=0 / I=1 sequence with normal-tangential coupling and the Panagiotopoulos
process) and examples/cattaneo.ref_out (Hertzian input IPOTCN=-3, then a shift with prescribed forces).
The parsed cases are committed as tests/golden/*_sequence.json (made by tests/golden/make_fixtures.py), so nothing here
needs /root/reference at run time; when the reference tree is present the reader is also run on the original files."""
import json
import os

import numpy as np
import pytest

from contact_b200 import inp as INP
from tests import inp_oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def _fmt(v, like):
    if "E" in like:
        s = "%.3E" % v
    else:
        s = "%.*f" % (len(like.split(".")[1]) if "." in like else 0, v)
    return s[1:] if s.startswith("-") and float(s) == 0.0 else s


def _load(name):
    return json.load(open(os.path.join(HERE, "golden", "%s_sequence.json" % name)))


def check_against_ref_out(results, ref_out):
    for k, (x, g) in enumerate(zip(results, ref_out), 1):
        f = g["forces"]                                   # FN/G, FX/FSTAT/FN, FY/FSTAT/FN, APPROACH, PMAX as printed
        assert _fmt(x["pen"], f[3]) == f[3] and _fmt(x["pmax"], f[4]) == f[4], (k, x["pen"], x["pmax"], f)
        v2 = x["fx"] if g.get("col2", "FX/FSTAT/FN") == "FX/FSTAT/FN" else x["cksi"]
        v3 = x["fy"] if g.get("col3", "FY/FSTAT/FN") == "FY/FSTAT/FN" else x["ceta"]
        assert _fmt(v2, f[1]) == f[1] and _fmt(v3, f[2]) == f[2], (k, v2, v3, f)
        if "stats" in g:                                  # NPOT NCON NADH NSLIP INORM ITANG
            assert [x["ncon"], x["nadh"], x["nslip"], x["itnorm"], x["ittang"]] == g["stats"][1:6], (k, g["stats"])


def test_oracle_reproduces_spence35_ref_out():
    d = _load("spence35")
    assert len(d["cases"]) == 35 and len(d["ref_out"]) == 35
    check_against_ref_out(inp_oracle.run_cases(d["cases"]), d["ref_out"])


def test_oracle_reproduces_cattaneo_ref_out():
    d = _load("cattaneo")
    r = inp_oracle.run_cases(d["cases"])
    check_against_ref_out(r, d["ref_out"])
    assert r[0]["itcg_norm"] == 4 and r[0]["ncon"] == 177                     # cattaneo.ref_out:10
    assert abs(r[0]["hz"]["a1"] - 0.01) < 1e-8 and abs(r[0]["hz"]["b1"] - 0.01) < 1e-8      # "THE CURVATURES A1,B1 ARE"
    assert r[1]["itgs_tang"] == 59 and (r[1]["nadh"], r[1]["nslip"]) == (45, 132)


def test_oracle_reproduces_carter2d_ref_out():
    """examples/carter2d.inp: the 2-D Carter problem (55 x 1 strip, steady rolling T=3 with SteadyGS, N=1, IPOTCN=3, DQ
    overruled to DX): examples/carter2d.ref_out:70, :519, :561 -- ItCG 11, ItGS 19, C/A/S = 50/30/20, Fx = 0.6480,
    approach 6.492E-03, pmax 113.9."""
    d = _load("carter2d")
    r = inp_oracle.run_cases(d["cases"])
    check_against_ref_out(r, d["ref_out"])
    assert r[0]["itcg_norm"] == 11 and r[0]["itgs_tang"] == 19


@pytest.mark.skipif(not os.path.exists("/root/reference/examples/spence35.inp"), reason="reference tree not present")
def test_reader_on_reference_files_matches_fixtures():
    for name in ("spence35", "cattaneo", "carter2d"):
        cases = INP.parse_inp(open("/root/reference/examples/%s.inp" % name).read())
        assert json.loads(json.dumps(cases)) == _load(name)["cases"]
    # the perf-suite inputs parse too (BASELINE configs): 69-case Spence sequence with the 11-depth subsurface block
    cs = INP.parse_inp(open("/root/reference/perfc_test/spence71_8281pt.inp").read())
    assert len(cs) == 69 and cs[0]["potcon"]["mx"] == 91 and len(cs[0]["subs"][0]["z"]) == 11 and cs[1]["P"] == 0 and cs[1]["I"] == 1
    cs = INP.parse_inp(open("/root/reference/perfc_test/tang_problm_8c.inp").read())
    assert cs[0]["G"] == 5 and cs[0]["potcon"]["mx"] == 575 and len(cs[0]["geom"]["prm"]) == 5 + 271


def test_reader_rules():
    txt = """
 3 MODULE
 201100   PBTNFS
 022020   LDCMZE
 0000011  HGIAOWR
   50 20 30 1 1d-5   MAXGS ...
  0.01, 0.0, 0.0, 0.0    PEN ...
  0.3 0.3
  0.28 0.28 82000. 82000.
  1
  5 4 -1. -1. 0.4 0.5
  1 1
  0.01 0. 0.02 % split record
  0. 0. 0.
 3 MODULE
 001100
 100000
 0110011
  0.02 0 0 0
 0 MODULE
"""
    cs = INP.parse_inp(txt)
    assert len(cs) == 2 and cs[0]["solver"]["eps"] == 1e-5 and cs[0]["geom"]["prm"][2] == 0.02
    assert "potcon" not in cs[1] and "fric" not in cs[1] and "solver" not in cs[1] and cs[1]["kin"][0] == 0.02
    full = INP.resolve_cases(cs)
    assert full[1]["potcon"]["mx"] == 5 and full[1]["fric"] == (0.3, 0.3) and full[1]["solver"]["maxgs"] == 50
    with pytest.raises(NotImplementedError):
        INP.parse_inp(" 1 MODULE\n")
