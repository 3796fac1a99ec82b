"""Repro of a refused group in the test_table_mx11 sweep: python tools/repro_tt2.py first count [first count ...] (one batch each)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import contact_b200 as cb
from tests import test_table_mx11 as T
rho, cp = T.hertz_constants_oracle()
args = [int(a) for a in sys.argv[1:]]
nre = max(args[1::2])
for ire in range(1, nre + 1):
    cb.cntc_initialize(ire, 3)
    cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_norm"]], [3, 1])
    cb.cntc_setfrictionmethod(ire, 1, 0, [T.FSTAT, T.FSTAT])
    cb.cntc_setmaterialparameters(ire, 1, 0, [T.NU1, T.NU1, T.G1, T.G1])
    cb.cntc_setreferencevelocity(ire, 1, 10000.0)
    cb.cntc_setnormalforce(ire, 1, T.FN)
    cb.cntc_setsolverflags(ire, 1, 0, [299, 1, 30, 1], [1e-6])
for first, n in zip(args[0::2], args[1::2]):
    ids = list(range(first, first + n))
    for k, icase in enumerate(ids):
        iell = T.decompose(icase)[0]
        cb.cntc_sethertzcontact(k + 1, 1, -3, [T.MX, T.MX, T.AA[iell], T.AA[iell] / T.ELLIP[iell], T.SCALE])
        cb.cntc_setcreepages(k + 1, 1, *T.creepages(icase, rho, cp))
    ierr = np.array(cb.cntc_calculate_batch(list(range(1, n + 1)), 1))
    bad = np.nonzero(ierr != 0)[0]
    print("batch", first, n, "failures", bad.size, "first/last", (ids[bad[0]], ids[bad[-1]]) if bad.size else None,
          "codes", sorted(set(ierr[bad].tolist())), cb.lib.last_error() if bad.size else "")
    if bad.size:
        k = bad[0]
        print("  first bad: decomp", T.decompose(ids[k]), "ncon/nadh/nslip areas", cb.cntc_getcontactpatchareas(k + 1, 1),
              "numel", cb.cntc_getnumelements(k + 1, 1) if hasattr(cb, "cntc_getnumelements") else None)
        runs = np.split(bad, np.nonzero(np.diff(bad) > 1)[0] + 1)
        print("  runs of bad cases:", [(ids[r[0]], ids[r[-1]]) for r in runs][:20])
