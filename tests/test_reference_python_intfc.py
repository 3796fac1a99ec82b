"""Drop-in acceptance: the reference's UNMODIFIED python_intfc package (python_intfc/__init__.py:22-44 loads
`<package>/../bin/contact_addon_linux64.so` and binds every prototype in contact_addon_headers.py) drives the B200 library.

The package is never part of this repository: the test takes it from /root/reference/python_intfc when that exists (this
container) or from oracle/_ref/python_intfc (a verbatim copy made by __graft_entry__.build(), git-ignored, which travels to the
GPU box beside the built libraries), copies it to a temporary directory whose bin/contact_addon_linux64.so is a symlink to
contact_b200/lib/libcontact_addon_b200.so, and runs a driver script in a fresh interpreter.

CPU: the package imports (every prototype of contact_addon_headers.py resolves in the library), the set-up calls and getters
that need no device work.  GPU: the normal problem of examples/cattaneo.inp (second case) through the reference wrappers, against the oracle and the
golden numbers of examples/cattaneo.ref_out.
"""
import json
import os
import shutil
import subprocess
import sys
import textwrap

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CANDIDATES = ["/root/reference/python_intfc", os.path.join(ROOT, "oracle", "_ref", "python_intfc")]
LIB = os.path.join(ROOT, "contact_b200", "lib", "libcontact_addon_b200.so")


def _stage(tmp_path):
    src = next((c for c in CANDIDATES if os.path.isfile(os.path.join(c, "__init__.py"))), None)
    if src is None:
        pytest.skip("the reference's python_intfc package is not available (neither /root/reference nor oracle/_ref)")
    assert os.path.exists(LIB), "library not built: python __graft_entry__.py"
    shutil.copytree(src, tmp_path / "python_intfc")
    os.makedirs(tmp_path / "bin")
    os.symlink(LIB, tmp_path / "bin" / "contact_addon_linux64.so")
    return tmp_path


def _run(tmp_path, body):
    script = tmp_path / "driver.py"
    script.write_text("import sys, json\nsys.path.insert(0, %r)\nimport numpy as np\nimport python_intfc as cntc\n" % str(tmp_path)
                      + "\n".join(textwrap.dedent(part) for part in body))
    env = dict(os.environ)
    env.setdefault("LD_LIBRARY_PATH", "/usr/local/cuda/lib64")          # __init__.py:33 indexes it unconditionally
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, r.stdout[-2000:]
    return json.loads(line[-1][7:])


SETUP = """
CNTC, ifcver, ierror = cntc.initlibrary(' ', ' ', ' ', 1)
assert ierror == 0, ierror
ire, icp = 1, 1
ifcver2, ierror = cntc.initialize(ire, 3)
assert ierror == 0
cntc.setflags(ire, icp, [CNTC['if_units'], CNTC['ic_norm'], CNTC['ic_tang'], CNTC['ic_pvtime']], [CNTC['un_cntc'], 1, 0, 2])
cntc.setmaterialparameters(ire, icp, 0, [0.28, 0.28, 82000.0, 82000.0])
cntc.setfrictionmethod(ire, icp, 0, [0.4, 0.4])
cntc.setsolverflags(ire, icp, 0, [100, 100, 30, 1], [1e-4])
"""


def test_unmodified_python_intfc_binds_and_sets_up(tmp_path):
    """Every prototype the reference binds exists in the library; set-up calls and state getters work without a device."""
    out = _run(_stage(tmp_path), (SETUP, """
    cntc.setpotcontact(ire, icp, 1, [19, 19, -0.475, -0.475, 0.05, 0.05])
    cntc.setundeformeddistc(ire, icp, 1, [0.01, 0.0, 0.01, 0.0, 0.0, 0.0])
    cntc.setnormalforce(ire, icp, 0.4705)
    cntc.setcreepages(ire, icp, 0.001, -0.002, 0.0)
    mx, my = cntc.getnumelements(ire, icp)
    dx, dy = cntc.getgriddiscretization(ire, icp)
    # (python_intfc/cntc_getflags.py:37 passes a double pointer against its own int prototype: unusable in the reference too)
    veloc = cntc.getreferencevelocity(ire, icp)
    print("RESULT " + json.dumps(dict(ifcver=ifcver, mx=int(mx), my=int(my), dx=float(dx), dy=float(dy),
                                       veloc=float(veloc), nmagic=len(CNTC))))
    cntc.finalize(ire)
    """))
    assert out["mx"] == 19 and out["my"] == 19 and abs(out["dx"] - 0.05) < 1e-15 and abs(out["dy"] - 0.05) < 1e-15
    assert out["ifcver"] > 0 and out["nmagic"] > 50 and out["veloc"] > 0.0


@pytest.mark.gpu
def test_unmodified_python_intfc_solves_cattaneo(tmp_path):
    """The normal problem of examples/cattaneo.inp:37-46 through the reference wrappers, against the oracle and the numbers of
    cattaneo.ref_out (177 elements, approach 1.998e-2)."""
    from tests import cases
    from oracle import oracle as O
    c = cases.CATTANEO2
    out = _run(_stage(tmp_path), (SETUP, """
    c = %r
    cntc.setmaterialparameters(ire, icp, 0, [c['poiss'][0], c['poiss'][1], c['gg'][0], c['gg'][1]])
    cntc.setsolverflags(ire, icp, 0, [c['maxgs'], c['maxin'], 30, 1], [c['eps']])
    cntc.setpotcontact(ire, icp, 1, [c['mx'], c['my'], c['xl'], c['yl'], c['dx'], c['dy']])
    cntc.setundeformeddistc(ire, icp, 1, c['prmudf'])
    cntc.setnormalforce(ire, icp, c['fn'])
    ierr = cntc.calculate(ire, icp)
    assert ierr == 0, ierr
    pn, px, py = cntc.gettractions(ire, icp)
    el = cntc.getelementdivision(ire, icp)
    pen = cntc.getpenetration(ire, icp)
    pmax = cntc.getmaximumpressure(ire, icp)
    carea, harea, sarea, parea = cntc.getcontactpatchareas(ire, icp)
    fn, fx, fy, mz = cntc.getcontactforces(ire, icp)
    un, ux, uy = cntc.getdisplacements(ire, icp)
    print("RESULT " + json.dumps(dict(pn=pn.ravel().tolist(), el=np.asarray(el).ravel().astype(int).tolist(), pen=float(pen),
                                       pmax=float(pmax), carea=float(carea), fn=float(fn), shape=list(pn.shape),
                                       un=un.ravel().tolist())))
    cntc.finalize(ire)
    """ % (dict(c),)))
    ref = O.norm_case(c["mx"], c["my"], c["xl"], c["yl"], c["dx"], c["dy"], c["gg"], c["poiss"], 1, c["prmudf"], 1,
                      fn=c["fn"], maxgs=c["maxgs"], maxin=c["maxin"], eps=c["eps"])
    el = np.array(out["el"])
    assert out["shape"] == [c["my"], c["mx"]]
    assert np.array_equal(el, ref["el"]) and int((el > 0).sum()) == 177              # cattaneo.ref_out: 177 elements in contact
    pn = np.array(out["pn"])
    assert np.abs(pn - ref["pn"]).max() < 1e-9 * np.abs(ref["pn"]).max()
    assert abs(out["pen"] - 1.998e-2) < 1e-5 and abs(out["pmax"] - pn.max()) < 1e-12
    assert abs(out["fn"] - c["fn"]) < 1e-9 * c["fn"]
    assert abs(out["carea"] - 177 * c["dx"] * c["dy"]) < 1e-12
