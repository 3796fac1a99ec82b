"""Timing probe: batch of mbench 71x81 steady-rolling cases (T=3, SteadyGS) through cntc_calculate_batch."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contact_b200 as cb
from tests import cases

def setup(ire, g, pen, tang, cks):
    cb.cntc_initialize(ire, 3)
    cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_force"], cb.CNTC["ic_iestim"]], [tang, 0, 0])
    cb.cntc_setsolverflags(ire, 1, 0, [1000, 100, 30, 1], [1e-7])
    cb.cntc_setmaterialparameters(ire, 1, 0, [0.28, 0.28, 82000.0, 82000.0])
    cb.cntc_setfrictionmethod(ire, 1, 0, [0.3, 0.3])
    cb.cntc_setpotcontact(ire, 1, 1, [g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"]])
    cb.cntc_setundeformeddistc(ire, 1, g["ibase"], g["prmudf"])
    cb.cntc_setpenetration(ire, 1, pen)
    if tang:
        cb.cntc_setrollingstepsize(ire, 1, 0.0, 0.1)
        cb.cntc_setcreepages(ire, 1, *cks)

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    mb = json.load(open(os.path.join(ROOT, "tests", "golden", "mbench_profile.json")))
    prm = [mb["nn"], mb["xm"], mb["rm"], mb["y1"], mb["dy1"]] + mb["heights"]
    g = dict(mx=71, my=81, xl=-3.55, yl=-6.15, dx=0.1, dy=0.1, ibase=2, prmudf=np.array(prm))
    u = np.random.default_rng(20240229).uniform(-1, 1, size=(n, 4))
    ires = list(range(1, n + 1))
    for tang in (0, 3):
        for rep in range(2):
            for i, ire in enumerate(ires):
                setup(ire, g, mb["pen"] * (1 + 0.1 * u[i, 0]), tang, (2e-3 * u[i, 1], 2e-3 * u[i, 2], 3e-4 * u[i, 3]))
            t0 = time.perf_counter()
            ierr = cb.cntc_calculate_batch(ires, 1)
            dt = time.perf_counter() - t0
            its = [cb.lowlevel.get_iterations(ire, 1) for ire in ires]
            pr = cb.lowlevel.steady_prof()
            if pr["steps"]:
                print("   cycles per element step: plstrc %.0f, re-integration %.0f, in-row update %.0f, other rows %.0f (%d steps, %d calls)" % (
                    pr["plstrc"] / pr["steps"], pr["reintegrate"] / pr["steps"], pr["update"] / pr["steps"], pr["rowupdate"] / pr["steps"], pr["steps"], pr["calls"]))
                print("   changes per step %.2f, net changes per row-block total %d" % (pr["changes"] / pr["steps"], pr["rowchanges"]))
            print("T=%d rep %d: %d cases in %.3f s = %.1f cases/s; ierr %s; mean itgs %.1f max %d, ncon %.0f" % (
                tang, rep, n, dt, n / dt, sorted(set(ierr.tolist())), np.mean([t["itgs"] for t in its]),
                max(t["itgs"] for t in its), np.mean([t["ncon"] for t in its])))

main()
