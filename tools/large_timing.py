"""Timing probe for the whole-GPU path: stand-alone 1x1 products and the NORM solve on the perfc grids."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import contact_b200 as cb

def main():
    mb = json.load(open(os.path.join(ROOT, "tests", "golden", "mbench_profile.json")))
    prm = np.array([mb["nn"], mb["xm"], mb["rm"], mb["y1"], mb["dy1"]] + mb["heights"])
    for mx, my, dx in ((143, 163, 0.05), (287, 323, 0.025), (575, 647, 0.0125)):
        npot = mx * my
        cset = cb.lowlevel.CoefSet(mx, my, dx, dx)
        pl = cset.plan()
        rng = np.random.default_rng(7)
        el = np.ones((1, npot), dtype=np.int32)
        p = np.zeros((1, 3, npot)); p[0, 2] = rng.standard_normal(npot)
        d_p = torch.tensor(p, device="cuda"); d_el = torch.tensor(el, device="cuda"); d_u = torch.zeros_like(d_p)
        for _ in range(3):
            cset.vecaijpj_dev(d_p, d_el, d_u, iigs=cb.lowlevel.ALLINT, ikarg=3, jkarg=3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = 50
        e0.record()
        for _ in range(nrep):
            cset.vecaijpj_dev(d_p, d_el, d_u, iigs=cb.lowlevel.ALLINT, ikarg=3, jkarg=3)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / nrep
        S = (pl["Fx"] + 1) * 2 * pl["Fy"]
        B = npot * 17 + 16 * S
        N = 4.0 * pl["Fx"] * pl["Fy"]
        F = 2 * 2.5 * N * np.log2(N) + 6 * S
        print("%dx%d: product %.1f us -> %.0f GB/s algorithmic (B = %.2f MB), %.2f TFLOP/s nominal" % (mx, my, us, B / us * 1e-3, B / 1e6, F / us * 1e-6))
        # NORM solve through the C-ABI
        for rep in range(3):
            ire = 5
            cb.cntc_initialize(ire, 3)
            cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_iestim"]], [0, 0])
            cb.cntc_setsolverflags(ire, 1, 0, [1000, 100, 30, 1], [1e-7])
            cb.cntc_setmaterialparameters(ire, 1, 0, [0.28, 0.28, 82000.0, 82000.0])
            cb.cntc_setpotcontact(ire, 1, 1, [mx, my, -3.55, -6.15, dx, dx])
            cb.cntc_setundeformeddistc(ire, 1, 2, prm)
            cb.cntc_setpenetration(ire, 1, mb["pen"])
            t0 = time.perf_counter()
            ierr = cb.cntc_calculate(ire, 1)
            dt = time.perf_counter() - t0
            its = cb.lowlevel.get_iterations(ire, 1)
            kms = cb.lowlevel.snorm_kernel_ms()
            cb.cntc_finalize(ire)
        print("   NORM: ierr %d, ncon %d, ItCG %d, cntc_calculate %.1f ms wall, solver kernel %.2f ms" % (ierr, its["ncon"], its["itcg"], dt * 1e3, kms))

main()
