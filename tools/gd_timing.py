"""Timing probe: perfc_test/tang_problm_{1,2,4,8}c.inp (mbench grids 71x81 ... 575x647, T=3, G=5 GDsteady with the solver
record of tang_problm_8c.inp:9) through the cntc_* C-ABI; prints wall time, solver-kernel time, iterations, line-search
trials and element counts, and compares with the oracle fixtures of tests/golden/gdsteady_mbench.json when present."""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contact_b200 as cb

GRIDS = {"1c": (71, 81, 0.1), "2c": (143, 161, 0.05), "4c": (287, 323, 0.025), "8c": (575, 647, 0.0125)}
GD_8C = (1.0, 0.05, 1, 2.0, -1.0, 1.0, 2.6, 1.0)


def run(name, gausei=5, reps=2):
    mb = json.load(open(os.path.join(ROOT, "tests", "golden", "mbench_profile.json")))
    prm = [mb["nn"], mb["xm"], mb["rm"], mb["y1"], mb["dy1"]] + mb["heights"]
    mx, my, dx = GRIDS[name]
    ire = 1
    out = None
    for rep in range(reps):
        cb.cntc_initialize(ire, 3)
        cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_force"], cb.CNTC["ic_iestim"]], [3, 0, 0])
        if gausei == 5:
            cb.cntc_setsolverflags(ire, 1, 5, [5000, 100, 30, 1, int(GD_8C[2])], [1e-7, GD_8C[0], GD_8C[1]] + list(GD_8C[3:]))
        else:
            cb.cntc_setsolverflags(ire, 1, 0, [5000, 100, 30, 1], [1e-7])
        cb.cntc_setmaterialparameters(ire, 1, 0, [0.28, 0.28, 82000.0, 82000.0])
        cb.cntc_setfrictionmethod(ire, 1, 0, [0.3, 0.3])
        cb.cntc_setpotcontact(ire, 1, 1, [mx, my, -3.55, -6.15, dx, dx])
        cb.cntc_setundeformeddistc(ire, 1, 2, np.array(prm))
        cb.cntc_setpenetration(ire, 1, mb["pen"])
        cb.cntc_setrollingstepsize(ire, 1, 0.0, dx)
        cb.cntc_setcreepages(ire, 1, 0.0005, 0.0, 0.0003)
        cb.lowlevel.gd_prof(reset=True)
        t0 = time.perf_counter()
        ierr = cb.cntc_calculate(ire, 1)
        wall = time.perf_counter() - t0
        its = cb.lowlevel.get_iterations(ire, 1)
        el = cb.cntc_getelementdivision(ire, 1).ravel().astype(np.int8)
        fn, tx, ty, mz = cb.cntc_getcontactforces(ire, 1)
        out = dict(case=name, G=gausei, ierr=int(ierr), wall_s=round(wall, 4), kernel_ms=round(cb.lowlevel.snorm_kernel_ms(), 3),
                   itgs=its["itgs"], trials=its["gd_trials"], fallback=its["gd_fallback"], ncon=int((el >= 1).sum()),
                   nadh=int((el == 1).sum()), nslip=int((el == 2).sum()), fx=tx / (0.3 * fn), fy=ty / (0.3 * fn),
                   el_sha1=hashlib.sha1(el.tobytes()).hexdigest())
        pr = cb.lowlevel.gd_prof()
        if pr["total"]:
            out["cycle_share"] = {k: round(pr[k] / pr["total"], 3) for k in ("products", "searchdir", "ls_rows", "ls_elements", "step")}
            out["kcycles_per_trial"] = {k: round(pr[k] / max(1, pr["trials"]) / 1e3, 1) for k in ("ls_rows", "ls_elements")}
            out["kcycles_per_iteration"] = round(pr["total"] / max(1, pr["iterations"]) / 1e3, 1)
            out["searchdir_kcycles_per_iteration"] = {k: round(pr[k] / max(1, pr["iterations"]) / 1e3, 1) for k in ("searchdir", "sd_copy", "sd_leader_rows", "sd_wait")}
            out["sd_fallbacks"] = pr["sd_fallbacks"]
        if ierr < 0:
            out["error"] = cb.lib.last_error()
        cb.cntc_finalize(ire)
    fx_path = os.path.join(ROOT, "tests", "golden", "gdsteady_mbench.json")
    if gausei == 5 and os.path.exists(fx_path):
        ref = json.load(open(fx_path)).get(name)
        if ref:
            out["oracle"] = dict(itgs=ref["itgs"], nslip=ref["nslip"], ncon=ref["ncon"], el_equal=(ref["el_sha1"] == out["el_sha1"]),
                                 dfx=out["fx"] - ref["fx"], dfy=out["fy"] - ref["fy"], oracle_seconds=ref["oracle_seconds"])
    return out


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["1c", "2c", "4c", "8c"]):
        print(json.dumps(run(name)), flush=True)
