import sys, math
sys.path.insert(0,'/root/repo')
import numpy as np
import contact_b200 as cb
from tests import test_table_mx11 as T
rho,cp=T.hertz_constants_oracle()
n=int(sys.argv[1]) if len(sys.argv)>1 else 40
first=int(sys.argv[2]) if len(sys.argv)>2 else 1289
for ire in range(1,n+1):
    cb.cntc_initialize(ire,3)
    import os
    cb.cntc_setflags(ire,1,[cb.CNTC["ic_tang"],cb.CNTC["ic_norm"]],[int(os.environ.get("TT_TANG","3")),1])
    cb.cntc_setfrictionmethod(ire,1,0,[T.FSTAT,T.FSTAT])
    cb.cntc_setmaterialparameters(ire,1,0,[T.NU1,T.NU1,T.G1,T.G1])
    cb.cntc_setreferencevelocity(ire,1,10000.0)
    cb.cntc_setnormalforce(ire,1,T.FN)
    cb.cntc_setsolverflags(ire,1,0,[299,1,30,1],[1e-6])
ids=list(range(first,first+n))
for k,icase in enumerate(ids):
    iell=T.decompose(icase)[0]
    cb.cntc_sethertzcontact(k+1,1,-3,[T.MX,T.MX,T.AA[iell],T.AA[iell]/T.ELLIP[iell],T.SCALE])
    cb.cntc_setcreepages(k+1,1,*T.creepages(icase,rho,cp))
ierr=cb.cntc_calculate_batch(list(range(1,n+1)),1)
print(ierr[:10], cb.lib.last_error())
