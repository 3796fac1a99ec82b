"""Oracle runs of perfc_test/tang_problm_{4,8}c.inp (T=3, G=5 GDsteady, solver record tang_problm_8c.inp:9) that take
minutes on the CPU: element counts, iterations, forces and a checksum of the element division are stored as fixtures
for the GPU tests (tests/test_gpu_parity.py).  These are ORACLE outputs (no golden file of the reference runs GDsteady
on these grids).  Usage: python tests/golden/make_gdsteady_fixtures.py [4c] [8c]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from tests import cases                 # noqa: E402

GRIDS = {"1c": (71, 81, 0.1), "2c": (143, 161, 0.05), "4c": (287, 323, 0.025), "8c": (575, 647, 0.0125)}
GD_8C = (1.0, 0.05, 1, 2.0, -1.0, 1.0, 2.6, 1.0)


def run(name):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "mbench_profile.json")))
    prm = [g["nn"], g["xm"], g["rm"], g["y1"], g["dy1"]] + g["heights"]
    mx, my, dx = GRIDS[name]
    grid = dict(mx=mx, my=my, xl=-3.55, yl=-6.15, dx=dx, dy=dx, ibase=2, prmudf=np.array(prm))
    t = time.time()
    r = O.contac(grid, cases.STEEL["gg"], cases.STEEL["poiss"], tang=3, norm=0, force3=0, pen=g["pen"], cksi=0.0005, ceta=0.0,
                 cphi=0.0003, fstat=0.3, fkin=0.3, maxgs=5000, maxin=100, maxnr=30, maxout=1, eps=1e-7, nn=g["nn"], chi=0.0,
                 dq=dx, gausei=5, gd=GD_8C)
    el = r["el"].astype(np.int8)
    return dict(grid=[mx, my, dx], ierror=int(r["ierror"]), gd_fallback=int(r["gd_fallback"]), itgs=int(r["itgs_tang"]),
                ittang=int(r["ittang"]), ncon=int((el >= 1).sum()), nadh=int((el == 1).sum()), nslip=int((el == 2).sum()),
                fx=float(r["fx"]), fy=float(r["fy"]), el_sha1=hashlib.sha1(el.tobytes()).hexdigest(),
                ps_absmax=float(np.abs(r["ps"][:2]).max()), ps_sum=[float(r["ps"][0].sum()), float(r["ps"][1].sum())],
                n_prod=int(r["n_prod"]), oracle_seconds=round(time.time() - t, 1))


if __name__ == "__main__":
    out_path = os.path.join(ROOT, "tests", "golden", "gdsteady_mbench.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for name in (sys.argv[1:] or ["4c"]):
        out[name] = run(name)
        print(name, out[name], flush=True)
        json.dump(out, open(out_path, "w"), indent=1)
