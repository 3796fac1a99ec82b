"""Small steady-rolling cases for compute-sanitizer (racecheck / memcheck of the three-role SteadyGS sweeps):
   compute-sanitizer --tool racecheck python tools/sanitize_steady.py
19x19 cattaneo grid, T=3: one case alone (walker + owners on one CTA), and a batch of three with different loads."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contact_b200 as cb
from tests import cases

c = cases.CATTANEO2


def setup(ire, fn, ck):
    cb.cntc_initialize(ire, 3)
    cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_force"], cb.CNTC["ic_iestim"]], [3, 0, 0])
    cb.cntc_setsolverflags(ire, 1, 0, [200, 100, 30, 1], [1e-5])
    cb.cntc_setmaterialparameters(ire, 1, 0, [c["poiss"][0], c["poiss"][1], c["gg"][0], c["gg"][1]])
    cb.cntc_setfrictionmethod(ire, 1, 0, [0.3, 0.3])
    cb.cntc_setpotcontact(ire, 1, 1, [c["mx"], c["my"], c["xl"], c["yl"], c["dx"], c["dy"]])
    cb.cntc_setundeformeddistc(ire, 1, 1, c["prmudf"])
    cb.cntc_setnormalforce(ire, 1, fn)
    cb.cntc_setrollingstepsize(ire, 1, 0.0, c["dx"])
    cb.cntc_setcreepages(ire, 1, ck, 0.3 * ck, 0.1 * ck)


setup(1, c["fn"], 1.5e-3)
assert cb.cntc_calculate(1, 1) == 0, cb.lib.last_error()
f1 = np.array(cb.cntc_getcontactforces(1, 1))
it = cb.lowlevel.get_iterations(1, 1)
ires = [2, 3, 4]
for k, ire in enumerate(ires):
    setup(ire, c["fn"] * (0.8 + 0.2 * k), 1.5e-3)
ierr = cb.cntc_calculate_batch(ires, 1)
assert (ierr == 0).all(), cb.lib.last_error()
f3 = np.array(cb.cntc_getcontactforces(3, 1))
assert np.array_equal(f1, f3), (f1, f3)                          # the same case alone and inside a batch
print("ok: itgs %d, ncon %d, forces %s" % (it["itgs"], it["ncon"], f1[:3]))
