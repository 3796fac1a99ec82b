"""Timing probe: perfc_test/spence71_8281pt.inp (69 cases + subsurface blocks) before and after the process has allocated the
work-space pool of a 592-case contact batch.  usage: python tools/spence71_timing.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contact_b200 as cb
import bench
r = bench.spence71_leg(cb)
print("fresh process: wall %.2f s, contact %.2f s, subsurface %.2f s" % (r["wall_s"], r["contact_s"], r["subsurf_s"]))
ires = list(range(1, 593)); draws = bench.hertz91_draws(592); sink = [None] * 592
bench.hertz91_setup(cb, ires, gausei=5)
bench.hertz91_step(cb, ires, draws, sink)
r = bench.spence71_leg(cb)
print("after a 592-case batch (pool allocated): wall %.2f s, contact %.2f s, subsurface %.2f s" % (r["wall_s"], r["contact_s"], r["subsurf_s"]))
