import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import contact_b200 as cb
mx, my, dx = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (575, 647, 0.0125)
npot = mx * my
cset = cb.lowlevel.CoefSet(mx, my, dx, dx)
rng = np.random.default_rng(7)
d_p = torch.tensor(rng.standard_normal((1, 3, npot)), device="cuda"); d_el = torch.ones((1, npot), dtype=torch.int32, device="cuda")
d_u = torch.zeros_like(d_p)
for _ in range(6):
    cset.vecaijpj_dev(d_p, d_el, d_u, iigs=cb.lowlevel.ALLINT, ikarg=3, jkarg=3)
torch.cuda.synchronize()
