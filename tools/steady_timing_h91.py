"""Timing probe: one batch of hertz-91 contact cases (N=1, T=3, G=0) through cntc_calculate_batch with the cycle split of the
SteadyGS element step (cb200_steady_prof).  usage: python tools/steady_timing_h91.py [ncase]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contact_b200 as cb
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
ires = list(range(1, n + 1))
draws = bench.hertz91_draws(n)
bench.hertz91_setup(cb, ires)
sink = [None] * n
for rep in range(2):
    cb.lowlevel.steady_prof()
    t0 = time.perf_counter()
    ierr, kms = bench.hertz91_step(cb, ires, draws, sink)
    dt = time.perf_counter() - t0
    pr = cb.lowlevel.steady_prof()
    its = [cb.lowlevel.get_iterations(ire, 1) for ire in ires]
    print("rep %d: %d cases in %.3f s (kernel %.1f ms); mean itgs %.1f, ncon %.0f" % (rep, n, dt, kms, np.mean([t["itgs"] for t in its]), np.mean([t["ncon"] for t in its])))
    print("   cycles per element step: plstrc %.0f, re-integration %.0f, in-row update %.0f, other rows %.0f (%d steps, %d calls); changes per step %.2f" % (
        pr["plstrc"] / pr["steps"], pr["reintegrate"] / pr["steps"], pr["update"] / pr["steps"], pr["rowupdate"] / pr["steps"], pr["steps"], pr["calls"], pr["changes"] / pr["steps"]))
if "--dump" in sys.argv:                                          # per-case inputs and work (scheduler cost model)
    import json
    json.dump({"draws": [list(map(float, d)) for d in draws], "itgs": [int(t["itgs"]) for t in its], "ncon": [int(t["ncon"]) for t in its]},
              open(os.path.join(ROOT, "gpurun_out", "h91_case_work.json"), "w"))
