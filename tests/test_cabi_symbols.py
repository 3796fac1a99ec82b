"""The C-ABI library loads without a GPU and exports every symbol include/contact_addon_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "contact_addon_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:cntc|subs|cb200)_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def dll():
    from contact_b200 import build
    build.build()
    import contact_b200
    return contact_b200.load_library()


def test_all_declared_symbols_exported(dll):
    names = _declared()
    assert len(names) >= 50
    raw = C.CDLL(os.path.join(ROOT, "contact_b200", "lib", "libcontact_addon_b200.so"))
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing


def test_prototypes_cover_header(dll):
    from contact_b200.lib import PROTOTYPES
    assert sorted(PROTOTYPES) == _declared()


def test_setters_and_getters_without_gpu(dll):
    """Host-side state handling (units, potential contact, flags) needs no device."""
    import contact_b200 as cb
    ire, icp = 901, 1
    ifcver, ierr = cb.cntc_initialize(ire, 3)
    assert ierr == 0
    cb.cntc_setflags(ire, icp, [cb.CNTC["if_units"], cb.CNTC["ic_tang"], cb.CNTC["ic_norm"]], [cb.CNTC["un_spck"], 0, 1])
    assert list(cb.cntc_getflags(ire, icp, [cb.CNTC["if_units"], cb.CNTC["ic_norm"]])) == [cb.CNTC["un_spck"], 1]
    cb.cntc_setpotcontact(ire, icp, 2, [10, 20, -0.005, -0.01, 0.005, 0.01])        # SI lengths [m]
    assert cb.cntc_getnumelements(ire, icp) == (10, 20)
    dx, dy = cb.cntc_getgriddiscretization(ire, icp)
    assert abs(dx - 0.001) < 1e-15 and abs(dy - 0.001) < 1e-15
    cb.cntc_setpenetration(ire, icp, 1e-5)
    assert abs(cb.cntc_getpenetration(ire, icp) - 1e-5) < 1e-18
    assert cb.cntc_getflags(ire, icp, [cb.CNTC["ic_norm"]])[0] == 0                   # setpenetration selects N=0
    cb.cntc_finalize(ire)


def test_invalid_ids_and_scope_errors(dll):
    import contact_b200 as cb
    assert cb.cntc_calculate(1000, 1) == -101           # invalid result element
    assert cb.cntc_calculate(5, -1) == cb.CNTC["err_icp"]      # module 1 (icp = -1) is out of scope


def test_no_cpu_fallback(dll):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import contact_b200 as cb
    cb.cntc_initialize(902, 3)
    cb.cntc_setpotcontact(902, 1, 1, [5, 5, -1.0, -1.0, 0.4, 0.4])
    cb.cntc_setundeformeddistc(902, 1, 1, [0.01, 0, 0.01, 0, 0, 0])
    cb.cntc_setpenetration(902, 1, 0.001)
    assert cb.cntc_calculate(902, 1) == -99
    assert "no CUDA device" in cb.lib.last_error()
    with pytest.raises(cb.lowlevel.CB200Error):
        cb.lowlevel.CoefSet(5, 5, 0.1, 0.1)
    cb.cntc_finalize(902)


def test_fortran_interface_file_is_current_and_complete():
    """include/contact_addon_b200.ifc (ISO_C_BINDING interface block, the counterpart of the reference's src/contact_addon.ifc)
    is what tools/gen_ifc.py derives from the C header, and it declares every cntc_* / subs_* entry point of the header."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ifc", os.path.join(root, "tools", "gen_ifc.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(root, "include", "contact_addon_b200.ifc")).read()
    assert text == gen.render()
    declared = set(re.findall(r"bind\(c, name='(\w+)'\)", text))
    header = {name for name, _, _ in gen.prototypes()}
    assert declared == header and len(declared) >= 66 and "cntc_calculate_batch" in declared
    assert len(re.findall(r"^end subroutine ", text, flags=re.M)) == len(declared)
