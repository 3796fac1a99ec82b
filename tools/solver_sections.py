"""Development tool: cycle split of the batched NORM kernel (CTA 0) over the sections of snorm / normcg (CB_T counters).
usage (GPU box): python tools/solver_sections.py [cases]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import contact_b200 as cb
from tests import cases

ll = cb.lowlevel
torch.cuda.set_device(0)
cb.load_library()
g = cases.HERTZ91
nsm = ll.num_sms()
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 8 * nsm
fns, _ = cases.hertz91_fn(ncase, fn0=bench.FN0)
cset = ll.CoefSet(g["mx"], g["my"], g["dx"], g["dy"], **cases.STEEL)
hs0, el0, pn0, scal0 = bench.initial_state(g, fns)
dev = torch.device("cuda", 0)
d_hs = torch.tensor(hs0, device=dev)
d_un = torch.zeros(ncase, g["mx"] * g["my"], dtype=torch.float64, device=dev)
for rep in range(3):
    d_el = torch.tensor(el0, device=dev); d_pn = torch.tensor(pn0, device=dev); d_scal = torch.tensor(scal0, device=dev)
    ll.solver_prof(reset=True)
    cset.snorm_batch_dev(d_hs, d_el, d_pn, d_un, d_scal, ic_norm=1, maxgs=bench.MAXGS, maxin=bench.MAXIN, eps=bench.EPS)
    torch.cuda.synchronize()
    ms = ll.snorm_kernel_ms()
    p = ll.solver_prof(reset=True)
tot = p["kernel_cycles"]
print("kernel %.3f ms, CTA 0: %d cycles, %d products, %.0f cycles per product inside the products" % (ms, tot, p["products"], p["conv_cycles"] / max(1, p["products"])))
for k, v in p.items():
    if k in ("products", "kernel_cycles", "conv_cycles") or k.startswith("table loads"):
        continue
    print("  %-40s %12d  %5.1f %%" % (k, v, 100.0 * v / tot))
print("  table loads: %d" % p["table loads (count)"])
