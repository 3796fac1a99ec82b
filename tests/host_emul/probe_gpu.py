import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
import contact_b200 as cb
from tests import cases
ll=cb.lowlevel
print("sms", ll.num_sms())
for (mx,my) in [(91,91),(19,19),(71,81)]:
    cs=ll.CoefSet(mx,my,0.1,0.1)
    print(mx,my,cs.plan())
    for ncase in (1,148,148*8):
        p,el=cases.microbench_p(mx,my,ncase)
        dp=torch.tensor(p,device='cuda'); de=torch.tensor(el,device='cuda',dtype=torch.int32); du=torch.zeros_like(dp)
        cs.vecaijpj_dev(dp,de,du); torch.cuda.synchronize()
        t0=torch.cuda.Event(enable_timing=True); t1=torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5): cs.vecaijpj_dev(dp,de,du)
        t1.record(); torch.cuda.synchronize()
        ms=t0.elapsed_time(t1)/5
        print("  ncase",ncase,"ms/launch",ms,"us/product/SM", ms*1e3/max(1,ncase/148) if ncase>=148 else ms*1e3)
