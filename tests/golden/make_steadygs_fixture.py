"""Oracle run of perfc_test/tang_problm_2c with the default solver (T=3, G=0: SteadyGS) on the 143x161 grid -- 61 s on one core,
so its result is committed as tests/golden/steadygs_2c.json (used by tests/test_gpu_parity.py::test_steadygs_whole_gpu_tang_problm_2c).
usage: python tests/golden/make_steadygs_fixture.py"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O      # noqa: E402
from tests import cases             # noqa: E402

mb = json.load(open(os.path.join(HERE, "mbench_profile.json")))
prm = np.array([mb["nn"], mb["xm"], mb["rm"], mb["y1"], mb["dy1"]] + mb["heights"])
g = dict(mx=143, my=161, xl=-3.55, yl=-6.15, dx=0.05, dy=0.05, ibase=2, prmudf=prm)
t = time.time()
r = O.contac(g, cases.STEEL["gg"], cases.STEEL["poiss"], tang=3, norm=0, force3=0, pen=mb["pen"], cksi=0.0005, ceta=0.0, cphi=0.0003,
             fstat=0.3, fkin=0.3, maxgs=1000, maxin=100, maxnr=30, maxout=1, eps=1e-7, nn=mb["nn"], chi=0.0, dq=0.05, gausei=0)
out = dict(ierror=r["ierror"], itgs=r["itgs_tang"], ncon=int((r["el"] >= 1).sum()), nslip=int((r["el"] == 2).sum()),
           el_sha1=hashlib.sha1(r["el"].astype(np.int8).tobytes()).hexdigest(), fx=r["fx"], fy=r["fy"],
           ps_absmax=float(np.abs(r["ps"]).max()), ps_sum=[float(r["ps"][0].sum()), float(r["ps"][1].sum())],
           oracle_seconds=round(time.time() - t, 1), golden="perfc_test/get_times.ref_out:26 nslp 7735, ItGS 90 (2016 revision)")
json.dump(out, open(os.path.join(HERE, "steadygs_2c.json"), "w"), indent=1)
print(out)
