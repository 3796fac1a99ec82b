#!/usr/bin/env python
"""bench.py -- contact solves/s on the 8281-element (91x91) grid, with roofline and CPU baseline.

Contract (driver):  python bench.py --gpus N --steps K --warmup W   [--impl reference]
One "step" = one batch of independent synthetic contact cases "hertz-91" exactly as SURVEY.md 8(d).2 defines them: grid
and quadratic gap of perfc_test/spence71_8281pt.inp:12-15, steel, N=1 with FN = 5e4 (1 + 0.2 u0), T=3 (steady rolling) with
CKSI, CETA = 2e-3 u1,2 and CPHI = 3e-4 u3, FSTAT = FKIN = 0.3, EPS 1e-6, MAXGS 999, default solver digits (G=0: NormCG +
SteadyGS); seed 20240229, case i uses draws [4i:4i+4].  Every case is a complete contac() call: NORM + TANG (panprc).
Cases are sharded over ranks with no data-path collective (weak scaling: the batch per GPU is fixed).

  e2e    : THE headline -- the drop-in path with HOST buffers: cntc_setnormalforce / cntc_setcreepages per case,
           cntc_calculate_batch, then cntc_gettractions / cntc_getelementdivision / cntc_getcontactforces per case; wall clock
           over K steps, uploads and downloads inside
  value  : the same cases / the time of the solver kernel(s) of those calls (CUDA events of the library on its stream):
           inputs resident in HBM when the timed region starts
  roofline / roofline_fp64 : dominant kernel k_contac_batch; work counted by the kernels themselves (cb200_work_counters):
           products at the transform size actually used, Gauss-Seidel sweeps in the reference's row-sum units
  norm_only : the NORM (NormCG) slice alone, device-resident and through cb200_snorm_batch (round-1 headline, kept as a
           secondary figure), with the roofline of k_snorm_batch
  gdsteady  : the same hertz-91 cases with G=5 (GDsteady, solver record of perfc_test/tang_problm_8c.inp:9)
  cpu_baseline : the CPU oracle (restatement of the reference algorithm, "port") timed on one host core on a bounded
           sample of the same cases; fft_calibration: its own FFT product beside an MKL-family FFT (torch CPU, oneMKL)
  --impl reference : the oracle with all host threads (one case per thread, as test_table.f90 / run_perfc.pl)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import cases  # noqa: E402  (seeded synthetic inputs only; no oracle import here)

FN0 = 50000.0
SUBS_Z = [1e-6, 0.1, 0.2, 0.3, 0.5, 1.0, 1.5, 2.0, 3.0, 5.0, 9.0]      # perfc_test/spence71_8281pt.inp:17-24
EPS, MAXGS, MAXIN = 1e-6, 999, 20
METRIC, UNIT = "contact_solves_per_s_8281el", "solves/s"


def workload_config(ncase_per_gpu, n_gpus):
    return {"workload": "hertz-91 (SURVEY 8(d).2): 91x91 (8281-element) grid and quadratic gap of spence71_8281pt.inp, steel, "
                        "N=1 FN=5e4*(1+0.2u), T=3 steady rolling CKSI,CETA=2e-3u CPHI=3e-4u, FSTAT=FKIN=0.3, EPS 1e-6, MAXGS 999, "
                        "G=0 (NormCG + SteadyGS), seed 20240229; every case a complete contac() call through cntc_calculate_batch",
            "cases_per_gpu": ncase_per_gpu, "cases_total": ncase_per_gpu * n_gpus,
            "l2_policy": "inputs larger than L2: per-step state of a batch (44 grid functions per case, 2.9 MB x cases) exceeds "
                         "126 MB and is re-uploaded and rewritten every step",
            "parallelism": "independent cases sharded over %d GPU(s), no data-path collective" % n_gpus}


def hertz91_draws(n, offset=0):
    """(FN, CKSI, CETA, CPHI) of cases offset .. offset+n-1 of the seeded hertz-91 sweep."""
    fns, u = cases.hertz91_fn(offset + n, fn0=FN0)
    return [(float(fns[i]), 2e-3 * float(u[i, 1]), 2e-3 * float(u[i, 2]), 3e-4 * float(u[i, 3])) for i in range(offset, offset + n)]


GD_RECORD = (1.0, 0.05, 1, 2.0, -1.0, 1.0, 2.6, 1.0)          # perfc_test/tang_problm_8c.inp:9


def hertz91_setup(cb, ires, gausei=0):
    """Result elements for hertz-91 cases: everything but the per-case load and creepages."""
    g = cases.HERTZ91
    for ire in ires:
        cb.cntc_initialize(ire, 3)
        cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_norm"], cb.CNTC["ic_force"], cb.CNTC["ic_iestim"]], [3, 1, 0, 0])
        if gausei == 5:
            cb.cntc_setsolverflags(ire, 1, 5, [MAXGS, 100, 30, 1, int(GD_RECORD[2])],
                                   [EPS, GD_RECORD[0], GD_RECORD[1]] + list(GD_RECORD[3:]))
        else:
            cb.cntc_setsolverflags(ire, 1, 0, [MAXGS, 100, 30, 1], [EPS])
        cb.cntc_setmaterialparameters(ire, 1, 0, [0.28, 0.28, 82000.0, 82000.0])
        cb.cntc_setfrictionmethod(ire, 1, 0, [0.3, 0.3])
        cb.cntc_setpotcontact(ire, 1, 1, [g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"]])
        cb.cntc_setundeformeddistc(ire, 1, g["ibase"], g["prmudf"])
        cb.cntc_setrollingstepsize(ire, 1, 0.0, g["dx"])


def hertz91_step(cb, ires, draws, sink):
    """One e2e step through the drop-in path: per-case inputs in, one batched solve, per-case results out (host arrays)."""
    for ire, (fn, cksi, ceta, cphi) in zip(ires, draws):
        cb.cntc_setnormalforce(ire, 1, fn)
        cb.cntc_setcreepages(ire, 1, cksi, ceta, cphi)
    ierr = cb.cntc_calculate_batch(ires, 1)
    kms = cb.lowlevel.snorm_kernel_ms()
    for k, ire in enumerate(ires):
        pn, px, py = cb.cntc_gettractions(ire, 1)
        el = cb.cntc_getelementdivision(ire, 1)
        f = cb.cntc_getcontactforces(ire, 1)
        sink[k] = (float(pn.max()), int((el == 2).sum()), f[1])
    return ierr, kms


def hertz91_oracle_case(O, d, gausei=0):
    g = cases.HERTZ91
    return O.contac(g, cases.STEEL["gg"], cases.STEEL["poiss"], tang=3, norm=1, force3=0, fn=d[0], cksi=d[1], ceta=d[2], cphi=d[3],
                    fstat=0.3, fkin=0.3, maxgs=MAXGS, maxin=100, maxnr=30, maxout=1, eps=EPS, chi=0.0, dq=g["dx"], gausei=gausei,
                    gd=GD_RECORD)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def mbench_grid(mx=71, my=81, dx=0.1):
    """The module-3 'right wheel at 6.2 mm' section of the perf suite (perfc_test/norm_problm_*p.inp, tang_problm_*c.inp)."""
    mb = json.load(open(os.path.join(ROOT, "tests", "golden", "mbench_profile.json")))
    prm = np.array([mb["nn"], mb["xm"], mb["rm"], mb["y1"], mb["dy1"]] + mb["heights"])
    return dict(mx=mx, my=my, xl=-3.55, yl=-6.15, dx=dx, dy=dx, ibase=2, prmudf=prm, pen=mb["pen"], nn=mb["nn"])


def rolling_sweep_leg(cb, n, rank_offset, gausei=0):
    """sweep-4096 class: mbench 71x81 grid, steady rolling T=3, penetration PEN (1 + 0.1 u) and creepages
    (2e-3 u, 2e-3 u, 3e-4 u), seed 20240229 -- n cases through cntc_calculate_batch.  gausei 0: the default solver
    (SteadyGS); 5: GDsteady with the solver record of perfc_test/tang_problm_8c.inp:9."""
    g = mbench_grid()
    u = np.random.default_rng(20240229).uniform(-1.0, 1.0, size=(4096, 4))[rank_offset:rank_offset + n]
    ires = list(range(1, n + 1))
    for i, ire in enumerate(ires):
        cb.cntc_initialize(ire, 3)
        cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_force"], cb.CNTC["ic_iestim"]], [3, 0, 0])
        if gausei == 5:
            cb.cntc_setsolverflags(ire, 1, 5, [999, 100, 30, 1, 1], [1e-5, 1.0, 0.05, 2.0, -1.0, 1.0, 2.6, 1.0])
        else:
            cb.cntc_setsolverflags(ire, 1, 0, [999, 100, 30, 1], [1e-5])
        cb.cntc_setmaterialparameters(ire, 1, 0, [0.28, 0.28, 82000.0, 82000.0])
        cb.cntc_setfrictionmethod(ire, 1, 0, [0.3, 0.3])
        cb.cntc_setpotcontact(ire, 1, 1, [g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"]])
        cb.cntc_setundeformeddistc(ire, 1, 2, g["prmudf"])
        cb.cntc_setpenetration(ire, 1, g["pen"] * (1.0 + 0.1 * u[i, 0]))
        cb.cntc_setrollingstepsize(ire, 1, 0.0, g["dx"])
        cb.cntc_setcreepages(ire, 1, 2e-3 * u[i, 1], 2e-3 * u[i, 2], 3e-4 * u[i, 3])
    t0 = time.perf_counter()
    ierr = cb.cntc_calculate_batch(ires, 1)
    dt = time.perf_counter() - t0
    kms = cb.lowlevel.snorm_kernel_ms()
    split = cb.lowlevel.batch_timing()
    its = [cb.lowlevel.get_iterations(ire, 1) for ire in ires]
    fxy = np.array([cb.cntc_getcontactforces(ire, 1)[:3] for ire in ires])
    for ire in ires:
        cb.cntc_finalize(ire)
    out = {"cases": n, "s": dt, "cases_per_s": n / dt, "solver_kernel_ms": kms, "mean_itgs": float(np.mean([t["itgs"] for t in its])),
           "errors": int((ierr < 0).sum()), "wall_split_s": split,
           "note": "mbench 71x81, T=3 %s, eps 1e-5, host buffers through cntc_calculate_batch (one launch)" % ("GDsteady (G=5)" if gausei == 5 else "SteadyGS")}
    if gausei == 5:
        out["fallbacks_to_steadygs"] = int(sum(t["gd_fallback"] for t in its))
        out["mean_linesearch_trials"] = float(np.mean([t["gd_trials"] for t in its]))
    out["_forces"] = fxy
    return out


def sweep4096_leg(cb, rank, world, total=4096, chunk=888):
    """BASELINE config 5: the 4096-case creepage / lateral-shift sweep on the mbench 71x81 grid (T=3, GDsteady as in
    today's perfc_test/tang_problm_*c.inp), a fixed total sharded over the ranks (contiguous blocks, no data-path
    collective), each rank in chunks of at most `chunk` cases per cntc_calculate_batch call (result elements are 1..999)."""
    lo, hi = (total * rank) // world, (total * (rank + 1)) // world
    s, kms, n, err, fb, its, splits = 0.0, 0.0, 0, 0, 0, 0.0, []
    nch = max(1, -(-(hi - lo) // chunk))                   # equal chunks (a short last chunk would leave most SMs idle)
    bounds = [lo + ((hi - lo) * k) // nch for k in range(nch + 1)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        m = b - a
        r = rolling_sweep_leg(cb, m, a, gausei=5)
        s += r["s"]; kms += r["solver_kernel_ms"]; n += m; err += r["errors"]; fb += r["fallbacks_to_steadygs"]
        its += r["mean_itgs"] * m
        w = r["wall_split_s"]
        splits.append([m] + [round(w[k], 4) for k in ("setup", "coefficients", "upload", "kernel", "output", "total")] + [round(r["s"], 4)])
    return {"cases": n, "s": s, "solver_kernel_ms": kms, "errors": err, "fallbacks_to_steadygs": fb, "mean_itgd": its / max(1, n),
            "chunks": splits}


def spence71_leg(cb):
    """perfc_test/spence71_8281pt.inp itself (BASELINE config: 69 cases in sequence on the 91x91 grid, dissimilar
    materials, Panagiotopoulos process, 11-depth subsurface block per case) through the .inp reader and cntc_calculate /
    subs_calculate.  A sequence runs one case at a time (one CTA), so this is a latency figure, not a throughput one."""
    from contact_b200 import inp as INP
    from tests.cases import inp_text_from_cases
    text, d = inp_text_from_cases("spence71")
    t0 = time.perf_counter()
    res = INP.run_inp(text, ire=950, with_fields=False)
    dt = time.perf_counter() - t0
    return {"cases": len(res), "errors": sum(1 for r in res if r["ierror"] != 0), "wall_s": dt,
            "contact_s": float(sum(r["wall_s"] for r in res)), "subsurf_s": float(sum(r.get("subs_wall_s", 0.0) for r in res)),
            "first12_contact_s": float(sum(r["wall_s"] for r in res[:12])),
            "ncon_final": res[-1].get("ncon"), "nout": int(sum(r["its"]["itout"] for r in res if "its" in r)),
            "golden": "perfc_test/get_times.ref_out:76-79: ncon 3657, nout 511, 30.7 s + 22.0 s subsurface (2016 host)"}


def gdsteady_leg(cb):
    """BASELINE config 3, perfc_test/tang_problm_8c.inp: 575x647 = 372 025 elements, PEN prescribed, T=3, G=5 (GDsteady with
    the solver record of line 9: E_trl, betath 0.05, kdowfb 1, d_ifc 2, d_lin -1, d_cns 1, d_slp 2.6, pow_s 1), MAXGS 5000,
    EPS 1e-7, through the cntc_* C-ABI on the whole-GPU path; the 287x323 case (4c) beside it.  Oracle results from
    tests/golden/gdsteady_mbench.json (the oracle needs minutes on these grids)."""
    import hashlib
    fx_path = os.path.join(ROOT, "tests", "golden", "gdsteady_mbench.json")
    fx = json.load(open(fx_path)) if os.path.exists(fx_path) else {}
    out = {}
    for name, (mx, my, dx) in (("tang_problm_4c", (287, 323, 0.025)), ("tang_problm_8c", (575, 647, 0.0125))):
        g = mbench_grid(mx, my, dx)
        ire = 901
        best = None
        for rep in range(2):
            cb.cntc_initialize(ire, 3)
            cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_force"], cb.CNTC["ic_iestim"]], [3, 0, 0])
            cb.cntc_setsolverflags(ire, 1, 5, [5000, 100, 30, 1, 1], [1e-7, 1.0, 0.05, 2.0, -1.0, 1.0, 2.6, 1.0])
            cb.cntc_setmaterialparameters(ire, 1, 0, [0.28, 0.28, 82000.0, 82000.0])
            cb.cntc_setfrictionmethod(ire, 1, 0, [0.3, 0.3])
            cb.cntc_setpotcontact(ire, 1, 1, [g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"]])
            cb.cntc_setundeformeddistc(ire, 1, 2, g["prmudf"])
            cb.cntc_setpenetration(ire, 1, g["pen"])
            cb.cntc_setrollingstepsize(ire, 1, 0.0, dx)
            cb.cntc_setcreepages(ire, 1, 0.0005, 0.0, 0.0003)
            t0 = time.perf_counter()
            ierr = cb.cntc_calculate(ire, 1)
            wall = time.perf_counter() - t0
            its = cb.lowlevel.get_iterations(ire, 1)
            el = cb.cntc_getelementdivision(ire, 1).ravel().astype(np.int8)
            r = {"ierror": int(ierr), "wall_s": wall, "solver_kernel_ms": cb.lowlevel.snorm_kernel_ms(), "itgd": its["itgs"],
                 "linesearch_trials": its["gd_trials"], "fallback_to_steadygs": its["gd_fallback"], "ncon": int((el >= 1).sum()),
                 "nadh": int((el == 1).sum()), "nslip": int((el == 2).sum())}
            ref = fx.get(name[-2:])
            if ref:
                r["oracle"] = {"itgd": ref["itgs"], "nslip": ref["nslip"], "seconds_1core_here": ref["oracle_seconds"],
                               "element_division_identical": hashlib.sha1(el.tobytes()).hexdigest() == ref["el_sha1"]}
            cb.cntc_finalize(ire)
            if best is None or r["wall_s"] < best["wall_s"]:
                best = r
        out[name] = best
    out["note"] = "T=3, G=5 (GDsteady), one case on the whole GPU: FFT products in three grid-wide phases, one warp per grid row for the integration along the rolling direction, line search on device"
    return out


def steadygs_large_leg(cb, with_4c=False):
    """perfc_test/tang_problm_2c (143x161) and, on request, tang_problm_4c (287x323) with the default solver (T=3, G=0: SteadyGS) on
    the whole-GPU path: the direct form of the sweep (one O(ncon) row sum per element step, CTA 0; golden ItGS 90 / 151, 97 s /
    2210 s on the 2016 host, perfc_test/get_times.ref_out:26-27)."""
    out = {}
    for name, (mx, my, dx) in (("tang_problm_2c", (143, 161, 0.05)), ("tang_problm_4c", (287, 323, 0.025)))[:2 if with_4c else 1]:
        g = mbench_grid(mx, my, dx)
        ire = 902
        cb.cntc_initialize(ire, 3)
        cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_force"], cb.CNTC["ic_iestim"]], [3, 0, 0])
        cb.cntc_setsolverflags(ire, 1, 0, [1000, 100, 30, 1], [1e-7])
        cb.cntc_setmaterialparameters(ire, 1, 0, [0.28, 0.28, 82000.0, 82000.0])
        cb.cntc_setfrictionmethod(ire, 1, 0, [0.3, 0.3])
        cb.cntc_setpotcontact(ire, 1, 1, [g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"]])
        cb.cntc_setundeformeddistc(ire, 1, 2, g["prmudf"])
        cb.cntc_setpenetration(ire, 1, g["pen"])
        cb.cntc_setrollingstepsize(ire, 1, 0.0, dx)
        cb.cntc_setcreepages(ire, 1, 0.0005, 0.0, 0.0003)
        t0 = time.perf_counter()
        ierr = cb.cntc_calculate(ire, 1)
        wall = time.perf_counter() - t0
        its = cb.lowlevel.get_iterations(ire, 1)
        el = cb.cntc_getelementdivision(ire, 1).ravel()
        out[name] = {"ierror": int(ierr), "wall_s": wall, "itgs": its["itgs"], "ncon": int((el >= 1).sum()), "nslip": int((el == 2).sum()),
                     "golden": {"tang_problm_2c": "nslp 7735, ItGS 90, 97 s (2016 host); oracle today: ItGS 91, 61 s on one core here",
                                "tang_problm_4c": "nslp 30181, ItGS 151, 2210 s (2016 host)"}[name]}
        cb.cntc_finalize(ire)
    return out


def large_grid_leg(cb, torch):
    """575x647 grid of perfc_test/norm_problm_8p.inp / tang_problm_8c.inp: stand-alone 1x1 products (three grid-wide
    phases over the L2-resident spectrum) and the whole NORM solve in one cooperative launch."""
    ll = cb.lowlevel
    g = mbench_grid(575, 647, 0.0125)
    npot = g["mx"] * g["my"]
    cset = ll.CoefSet(g["mx"], g["my"], g["dx"], g["dy"], **cases.STEEL)
    pl = cset.plan()
    rng = np.random.default_rng(7)
    d_p = torch.tensor(rng.standard_normal((1, 3, npot)), device="cuda")
    d_el = torch.ones((1, npot), dtype=torch.int32, device="cuda")
    d_u = torch.zeros_like(d_p)
    for _ in range(3):
        cset.vecaijpj_dev(d_p, d_el, d_u, iigs=ll.ALLINT, ikarg=3, jkarg=3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nrep = 40
    e0.record()
    for _ in range(nrep):
        cset.vecaijpj_dev(d_p, d_el, d_u, iigs=ll.ALLINT, ikarg=3, jkarg=3)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / nrep
    S = (pl["Fx"] + 1) * 2 * pl["Fy"]
    B = npot * 17 + 16 * S                                   # 8 npot (nj + ni) + npot + 16 S nb
    N = 4.0 * pl["Fx"] * pl["Fy"]
    F = 2 * 2.5 * N * np.log2(N) + 6 * S
    ire = 900
    kms = []
    for _ in range(3):
        cb.cntc_initialize(ire, 3)
        cb.cntc_setflags(ire, 1, [cb.CNTC["ic_tang"], cb.CNTC["ic_iestim"]], [0, 0])
        cb.cntc_setsolverflags(ire, 1, 0, [1000, 100, 30, 1], [1e-7])
        cb.cntc_setmaterialparameters(ire, 1, 0, [0.28, 0.28, 82000.0, 82000.0])
        cb.cntc_setpotcontact(ire, 1, 1, [g["mx"], g["my"], g["xl"], g["yl"], g["dx"], g["dy"]])
        cb.cntc_setundeformeddistc(ire, 1, 2, g["prmudf"])
        cb.cntc_setpenetration(ire, 1, g["pen"])
        t0 = time.perf_counter()
        ierr = cb.cntc_calculate(ire, 1)
        wall = time.perf_counter() - t0
        its = ll.get_iterations(ire, 1)
        kms.append(ll.snorm_kernel_ms())
        cb.cntc_finalize(ire)
    return {"grid": "575x647", "product_us": us, "alg_bytes_per_product": B, "alg_GBps": B / us * 1e-3,
            "nominal_flops_per_product": F, "nominal_TFLOPs": F / us * 1e-6,
            "norm_8p": {"ierror": int(ierr), "ncon": its["ncon"], "itcg": its["itcg"], "solver_kernel_ms": float(np.min(kms)),
                        "cntc_calculate_wall_ms": wall * 1e3,
                        "golden": "perfc_test/get_times.ref_out:10 ncon 200980, ItCG 31, 17.6 s (2016 host)"}}


def initial_state(g, fns):
    """Host-side solver inputs of each case: gap h, initial element division and pressure (zero)."""
    import contact_b200  # noqa: F401
    h = cases.quadratic_h(g)
    ncase, npot = len(fns), g["mx"] * g["my"]
    hs = np.tile(h, (ncase, 1))
    from contact_b200 import lowlevel as ll
    el = np.zeros((ncase, npot), dtype=np.int32)
    pn = np.zeros((ncase, npot))
    scal = np.zeros((ncase, 8))
    scal[:, 1] = fns
    for i, fn in enumerate(fns):        # the reference's own initial estimate (eldiv0), library host logic
        el[i], scal[i, 0] = ll.eldiv0(g["mx"], g["my"], g["dx"], g["dy"], cases.STEEL["gg"], cases.STEEL["poiss"],
                                      g["ibase"], g["prmudf"] + [0.0, 0.0], 1, float(fn), 0.0, h)
    return hs, el, pn, scal


def norm_only_leg(args, torch, dist, cb, rank, world, local, steps):
    """The NORM slice alone (round-1 headline): 8 x SM-count hertz-91 normal problems per step on the device-resident NormCG path
    (k_snorm_batch), and the same through cb200_snorm_batch with pinned host buffers.  Returns this rank's numbers."""
    ll = cb.lowlevel
    from contact_b200 import scheduler
    g = cases.HERTZ91
    npot = g["mx"] * g["my"]
    nsm = ll.num_sms()
    ncase = args.cases if args.cases > 0 else 8 * nsm           # multiple of the SM count: one CTA per case, 8 waves
    fns_all, _ = cases.hertz91_fn(ncase * world, fn0=FN0)
    lo, hi = scheduler.my_range(ncase * world, rank, world)      # static block sharding of the case index
    fns = fns_all[lo:hi]
    cset = ll.CoefSet(g["mx"], g["my"], g["dx"], g["dy"], **cases.STEEL)
    hs0, el0, pn0, scal0 = initial_state(g, fns)
    dev = torch.device("cuda", local)
    d_hs = torch.tensor(hs0, device=dev)
    d_el0 = torch.tensor(el0, device=dev); d_pn0 = torch.tensor(pn0, device=dev); d_scal0 = torch.tensor(scal0, device=dev)
    d_el = torch.empty_like(d_el0); d_pn = torch.empty_like(d_pn0); d_scal = torch.empty_like(d_scal0)
    d_un = torch.empty_like(d_pn0)

    def step_device():
        d_el.copy_(d_el0); d_pn.copy_(d_pn0); d_scal.copy_(d_scal0)          # fresh initial estimate every step
        cset.snorm_batch_dev(d_hs, d_el, d_pn, d_un, d_scal, ic_norm=1, maxgs=MAXGS, maxin=MAXIN, eps=EPS)

    e2e_steps = max(1, min(steps, 5))
    p_hs = torch.tensor(hs0).pin_memory()
    p_un = torch.zeros(ncase, npot, dtype=torch.float64).pin_memory()
    sets = [(torch.tensor(el0).pin_memory(), torch.tensor(pn0).pin_memory(), torch.tensor(scal0).pin_memory()) for _ in range(e2e_steps)]
    p_el, p_pn, p_scal = sets[0]

    def reset_set(k):
        sets[k][0].copy_(torch.from_numpy(el0)); sets[k][1].zero_(); sets[k][2].copy_(torch.from_numpy(scal0))

    def step_host(k=0):
        el_k, pn_k, scal_k = sets[k]
        cset.snorm_batch(p_hs.numpy(), el_k.numpy(), pn_k.numpy(), p_un.numpy(), scal_k.numpy(), ic_norm=1,
                         maxgs=MAXGS, maxin=MAXIN, eps=EPS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    ll.work_counters(reset=True)
    kt = []
    for _ in range(min(3, steps)):
        step_device()
        kt.append(ll.snorm_kernel_ms())
    torch.cuda.synchronize()
    work = ll.work_counters(reset=True)
    nk = len(kt)
    kernel_ms = float(np.mean(kt))
    cprof = ll.conv_prof()
    table = scheduler.gather_case_results(d_scal, ncase * world)   # the only collective: final gather of per-case results
    assert table.shape[0] == ncase * world
    scal = d_scal.cpu().numpy()
    # ---- subsurface leg (ISUBS 5 block of spence71_8281pt.inp: 11 depths x 8281 points, 36 products per depth) ----
    nsub = min(ncase, nsm)
    d_ps = torch.zeros(nsub, 3, npot, dtype=torch.float64, device=dev)
    d_ps[:, 2] = d_pn[:nsub]
    d_tab = torch.empty(nsub, len(SUBS_Z), npot, 18, dtype=torch.float64, device=dev)
    cset.subsurf_batch_dev(d_ps, SUBS_Z, d_tab, **cases.STEEL)              # builds + caches the 11 x 36 transforms
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(2):
        cset.subsurf_batch_dev(d_ps, SUBS_Z, d_tab, **cases.STEEL)
    s1.record()
    torch.cuda.synchronize()
    subs_ms = s0.elapsed_time(s1) / 2
    del d_tab
    for _ in range(2):
        step_host(0)
        reset_set(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        step_host(k)
    barrier()
    e2e_s = time.perf_counter() - t0
    plan = cset.plan()
    N = 4.0 * plan["Fx"] * plan["Fy"]; S = (plan["Fx"] + 1) * 2.0 * plan["Fy"]
    return {"ncase": ncase, "steps": steps, "ms_total": ms_total, "kernel_ms": kernel_ms, "e2e_s": e2e_s, "e2e_steps": e2e_steps,
            "products_per_step": work["products"] / nk, "product_flops_per_step": work["product_flops"] / nk,
            "product_bytes_per_step": work["product_bytes"] / nk, "full_grid_flops_per_product": 2 * 2.5 * N * np.log2(N) + 6 * S,
            "mean_itcg": float(scal[:, 2].mean()),
            "cta0": {"per_product": cprof["conv_cycles"] / max(1, cprof["products"]),
                     "conv_share_of_kernel": cprof["conv_cycles"] / max(1, cprof["kernel_cycles"])},
            "h2d": int(p_hs.numel() * 8 + p_el.numel() * 4 + p_pn.numel() * 8 + p_scal.numel() * 8),
            "d2h": int(p_pn.numel() * 8 + p_el.numel() * 4 + p_un.numel() * 8 + p_scal.numel() * 8),
            "subsurf": {"cases": nsub, "depths": len(SUBS_Z), "points_per_case": npot * len(SUBS_Z),
                        "products": nsub * len(SUBS_Z) * 36, "ms": subs_ms, "cases_per_s": nsub / (subs_ms * 1e-3),
                        "note": "ISUBS=5 block of spence71_8281pt.inp on the solved pressures of this rank; "
                                "coefficient transforms cached per (grid, material, z)"}}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import contact_b200 as cb
    ll = cb.lowlevel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; contact_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cb.load_library()
    dev = torch.device("cuda", local)
    g = cases.HERTZ91
    npot = g["mx"] * g["my"]
    nsm = ll.num_sms()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- headline: complete hertz-91 contact cases (N=1, T=3, G=0) through the drop-in path ----
    ncase = args.contact_cases if args.contact_cases > 0 else 4 * nsm          # four waves of one CTA per case (dynamic queue)
    ncase = min(ncase, 999)                                                     # result elements are 1..999
    ires = list(range(1, ncase + 1))
    draws = hertz91_draws(ncase, offset=rank * ncase)
    hertz91_setup(cb, ires, gausei=0)
    sink = [None] * ncase
    for _ in range(args.warmup):
        ierr, _ = hertz91_step(cb, ires, draws, sink)
    assert (np.asarray(ierr) >= 0).all(), (ierr, cb.lib.last_error())
    barrier()
    ll.work_counters(reset=True)
    launches0 = ll.num_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    kms_sum = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ierr, kms = hertz91_step(cb, ires, draws, sink)
        kms_sum += kms
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = ll.num_launches() - launches0
    work = ll.work_counters(reset=True)
    split = ll.batch_timing()
    its = [ll.get_iterations(ire, 1) for ire in ires]
    contact_stats = {"mean_itcg": float(np.mean([t["itcg"] for t in its])), "mean_itgs": float(np.mean([t["itgs"] for t in its])),
                     "mean_ncon": float(np.mean([t["ncon"] for t in its])), "mean_nslip": float(np.mean([v[1] for v in sink])),
                     "errors": int((np.asarray(ierr) < 0).sum()), "last_call_wall_split_s": split}
    # the same cases with GDsteady (G=5), the reference's newest steady-rolling solver
    gd = None
    if not args.skip_extra:
        hertz91_setup(cb, ires, gausei=5)
        hertz91_step(cb, ires, draws, sink)
        barrier()
        tg = time.perf_counter()
        ierr5, kms5 = hertz91_step(cb, ires, draws, sink)
        barrier()
        gd_s = time.perf_counter() - tg
        it5 = [ll.get_iterations(ire, 1) for ire in ires]
        gd = {"s": gd_s, "kms": kms5, "errors": int((np.asarray(ierr5) < 0).sum()), "mean_itgd": float(np.mean([t["itgs"] for t in it5])),
              "fallbacks_to_steadygs": int(sum(t["gd_fallback"] for t in it5))}
    for ire in ires:
        cb.cntc_finalize(ire)

    # ---- secondary legs ----
    norm = norm_only_leg(args, torch, dist, cb, rank, world, local, min(args.steps, 10))
    nroll = nsm if args.cases <= 0 else min(args.cases, nsm)
    roll = rolling_sweep_leg(cb, nroll, (rank * nroll) % (4096 - nroll)) if not args.skip_extra else None
    roll_gd = rolling_sweep_leg(cb, nroll, (rank * nroll) % (4096 - nroll), gausei=5) if not args.skip_extra else None
    if roll and roll_gd:                     # the two solvers on the same cases: largest difference of the total forces
        f0, f5 = roll.pop("_forces"), roll_gd.pop("_forces")
        roll_gd["max_rel_force_diff_vs_steadygs"] = float(np.abs(f5 - f0).max() / np.abs(f0).max())
    sweep = sweep4096_leg(cb, rank, world) if (args.sweep4096 and not args.skip_extra) else None
    inlib = None
    if world == 1 and args.inlib_devices > 1 and torch.cuda.device_count() >= args.inlib_devices:
        # the batched case scheduler inside the library: one process, every cntc_calculate_batch call cut into contiguous shards over the
        # devices (one host thread each); first pass warms the per-device coefficient caches and work-space pools
        nd = ll.set_devices(args.inlib_devices)
        sweep4096_leg(cb, 0, 1, total=2 * 888)
        r = sweep4096_leg(cb, 0, 1)
        ll.set_devices(1)
        inlib = {"devices": nd, "cases_total": r["cases"], "s": r["s"], "cases_per_s": r["cases"] / r["s"], "errors": r["errors"],
                 "chunks": r["chunks"], "note": "BASELINE config 5 through ONE process: cb200_set_devices(n), result elements 1..999 => chunks of "
                                                "<= 888 cases per call, each call spread over the devices; wall clock"}
    large = large_grid_leg(cb, torch) if (rank == 0 and not args.skip_extra) else None
    sp71 = spence71_leg(cb) if (rank == 0 and not args.skip_extra) else None
    gdl = gdsteady_leg(cb) if (rank == 0 and not args.skip_extra) else None
    gsl = steadygs_large_leg(cb, args.gs_4c) if (rank == 0 and not args.skip_extra) else None

    roll_s = roll["s"] if roll else 0.0
    rollgd_s = roll_gd["s"] if roll_gd else 0.0
    t = torch.tensor([kms_sum, e2e_s, norm["ms_total"], norm["e2e_s"], norm["kernel_ms"], roll_s, rollgd_s, sweep["s"] if sweep else 0.0,
                      sweep["solver_kernel_ms"] if sweep else 0.0, gd["s"] if gd else 0.0, gd["kms"] if gd else 0.0],
                     dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmax = [float(v) for v in tmax]
    kms_sum, e2e_s = tmax[0], tmax[1]
    if roll:
        roll["cases_per_s"] = roll["cases"] * world / tmax[5]; roll["cases_total"] = roll["cases"] * world
    if roll_gd:
        roll_gd["cases_per_s"] = roll_gd["cases"] * world / tmax[6]; roll_gd["cases_total"] = roll_gd["cases"] * world

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp64_peak = ll.fp64_peak_tflops(3)
        fp64_src = "measured here with a DFMA chain kernel (cb200_fp64_peak_tflops); MEASURED_PEAKS.json has no FP64 entry"
        # k_contac_batch, this rank's launches of the timed region: work counted by the kernels themselves
        ksec = kms_sum * 1e-3
        gs_flops = 16.0 * work["gs_units"]                   # 2 ncon row sums x 2 directions x 2 flops x 2 (ncon + 2 my) columns per sweep
        c_flops = work["product_flops"] + gs_flops
        c_bytes = float(work["product_bytes"])
        traffic = norm_traffic = None
        try:                                    # dram bytes per case of the kernels from the committed ncu captures
            import glob
            tf = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")))[-1]))
            norm_traffic = tf["k_snorm_batch"]["dram_bytes_per_case"] * norm["ncase"]
            traffic = tf["k_contac_batch"]["dram_bytes_per_case"] * ncase if "k_contac_batch" in tf else None
        except Exception:
            pass
        n_ach = norm["product_bytes_per_step"] / (norm["kernel_ms"] * 1e-3) / 1e9
        n_fl = norm["product_flops_per_step"] / (norm["kernel_ms"] * 1e-3) / 1e12
        per_case_up, per_case_down = 6 * npot * 8 + npot * 4, 8 * npot * 8 + npot * 4
        out = {
            "metric": METRIC, "value": ncase * world * args.steps / ksec, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kms_sum / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(ncase, world),
            "e2e": {"value": ncase * world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": per_case_up * ncase,
                    "d2h_bytes_per_step": per_case_down * ncase,
                    "path": "cntc_setnormalforce + cntc_setcreepages per case, cntc_calculate_batch, cntc_gettractions + "
                            "cntc_getelementdivision + cntc_getcontactforces per case (host arrays); bytes = what the library moves per "
                            "case: hs, rigid slip, tractions (6 npot f64) + element division up; tractions, slip, displacements "
                            "(8 npot f64) + element division down"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_contac_batch", "achieved": c_bytes / ksec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": c_bytes / ksec / 1e9 / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "note": "algorithmic bytes of the FFT products only (17 per element of the box each product ran on); the "
                                 "kernel is bound by the FP64 latency chain of the Gauss-Seidel sweeps, not by HBM: see roofline_fp64"},
            "roofline_fp64": {"kernel": "k_contac_batch", "achieved": c_flops / ksec / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": c_flops / ksec / 1e12 / fp64_peak if fp64_peak > 0 else None, "peak_source": fp64_src,
                              "flops": {"products_at_level_size": work["product_flops"], "gauss_seidel_row_sums": gs_flops,
                                        "products": work["products"]},
                              "note": "SURVEY 8(d) units: products 2*2.5 N log2 N + 6 S at the transform size actually used; a "
                                      "Gauss-Seidel sweep = 2 ncon row sums over 2 (ncon + 2 my) columns, 4 flops each (the device keeps "
                                      "U = A dp in registers and does rank-1 updates: same count)"},
            "kernel_ms": kms_sum / args.steps, "contact": contact_stats,
            "norm_only": {"workload": "hertz-91 normal problems (T=0) only, %d per step" % norm["ncase"],
                          "value": norm["ncase"] * world * norm["steps"] / (tmax[2] * 1e-3), "unit": UNIT,
                          "e2e": norm["ncase"] * world * norm["e2e_steps"] / tmax[3], "e2e_path": "cb200_snorm_batch, pinned host buffers",
                          "h2d_bytes_per_step": norm["h2d"], "d2h_bytes_per_step": norm["d2h"],
                          "kernel_ms": tmax[4], "products_per_step": norm["products_per_step"], "mean_itcg": norm["mean_itcg"],
                          "roofline": {"bound": "hbm", "kernel": "k_snorm_batch", "achieved": n_ach, "peak": hbm_peak, "unit": "GB/s",
                                       "frac": n_ach / hbm_peak, "traffic": norm_traffic},
                          "roofline_fp64": {"achieved": n_fl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": n_fl / fp64_peak if fp64_peak > 0 else None,
                                            "note": "flops counted per product at the transform size of its contact-box level",
                                            "frac_if_all_products_were_full_grid": norm["products_per_step"] * norm["full_grid_flops_per_product"]
                                            / (norm["kernel_ms"] * 1e-3) / 1e12 / fp64_peak if fp64_peak > 0 else None},
                          "cta0_cycles": norm["cta0"]},
            "subsurf": norm["subsurf"],
        }
        if gd:
            out["gdsteady"] = {"workload": "the same hertz-91 cases with G=5 (GDsteady)", "value": ncase * world / (tmax[10] * 1e-3),
                               "e2e": ncase * world / tmax[9], "unit": UNIT, "mean_itgd": gd["mean_itgd"], "errors": gd["errors"],
                               "fallbacks_to_steadygs": gd["fallbacks_to_steadygs"]}
        if roll:
            out["rolling_sweep"] = roll
        if roll_gd:
            out["rolling_sweep_gdsteady"] = roll_gd
        if sweep:
            out["sweep4096"] = {"cases_total": 4096, "s": tmax[7], "cases_per_s": 4096 / tmax[7],
                                "solver_kernel_ms_max_rank": tmax[8], "cases_this_rank": sweep["cases"],
                                "errors_this_rank": sweep["errors"], "fallbacks_this_rank": sweep["fallbacks_to_steadygs"],
                                "mean_itgd_this_rank": sweep["mean_itgd"], "scaling": "strong",
                                "chunks_this_rank": {"columns": ["cases", "setup", "coefficients", "upload", "kernel", "output", "total", "wall"],
                                                     "rows": sweep["chunks"]},
                                "note": "BASELINE config 5: 4096 mbench 71x81 cases (PEN and creepage draws of seed 20240229), T=3 GDsteady, "
                                        "eps 1e-5, host buffers through cntc_calculate_batch in chunks of <= 888 cases; wall clock, max over ranks"}
        if large:
            large["frac_hbm"] = large["alg_GBps"] / hbm_peak
            large["frac_fp64"] = large["nominal_TFLOPs"] / fp64_peak if fp64_peak > 0 else None
            out["large_grid"] = large
        if sp71:
            out["spence71_inp"] = sp71
        if gdl:
            out["gdsteady_large"] = gdl
        if gsl:
            out["steadygs_large"] = gsl
        if inlib:
            out["sweep4096_inlib"] = inlib
        if args.cpu_seconds > 0:
            out["cpu_baseline"] = cpu_baseline(args.cpu_seconds, threads=1)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def fft_calibration():
    """The port's own mixed-radix FFT product beside an MKL-family one: torch's CPU FFT is oneMKL, the library the reference
    links (MKL DFTI, m_aijpj.f90:841-970).  One 1x1 product on the 91x91 grid = real 192x192 forward transform, pointwise
    multiply with the 97x192 coefficient spectrum, inverse transform; single thread, microseconds per product."""
    import torch
    from oracle import oracle as O
    F = 192
    rng = np.random.default_rng(7)
    a = rng.standard_normal((F, F))
    ch = rng.standard_normal((F, F // 2 + 1)) + 1j * rng.standard_normal((F, F // 2 + 1))
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    ta, tc = torch.tensor(a), torch.tensor(ch)
    for _ in range(5):
        torch.fft.irfft2(torch.fft.rfft2(ta) * tc, s=(F, F))
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        torch.fft.irfft2(torch.fft.rfft2(ta) * tc, s=(F, F))
    mkl_us = (time.perf_counter() - t0) / n * 1e6
    torch.set_num_threads(nt)
    for _ in range(2):
        O.fft2_c2r(O.fft2_r2c(a) * ch, F)
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        O.fft2_c2r(O.fft2_r2c(a) * ch, F)
    port_us = (time.perf_counter() - t0) / n * 1e6
    return {"product": "192x192 real forward FFT + 97x192 complex multiply + inverse, 1 thread", "port_us": port_us, "mkl_us": mkl_us,
            "port_over_mkl": port_us / mkl_us,
            "note": "torch CPU FFT = oneMKL; the port times include a plan-cache rebuild per call (python wrapper), so this is an upper "
                    "bound of the port's disadvantage; an MKL build of the reference would be faster than the port by at most this factor on "
                    "the product share of a solve"}


def cpu_baseline(budget_s, threads):
    """The CPU oracle on a bounded sample of the same cases (imports oracle/: allowed for this leg only)."""
    from oracle import oracle as O
    g = cases.HERTZ91
    # headline workload: complete hertz-91 contact cases (N=1, T=3, G=0) on ONE core, as many as fit the budget
    draws = hertz91_draws(64)
    t0 = time.perf_counter()
    nc, itgs = 0, []
    while nc < len(draws) and (nc < 2 or time.perf_counter() - t0 < 0.6 * budget_s):
        r = hertz91_oracle_case(O, draws[nc])
        assert r["ierror"] == 0
        itgs.append(r["itgs_tang"]); nc += 1
    dtc = time.perf_counter() - t0
    t0 = time.perf_counter()
    r5 = hertz91_oracle_case(O, draws[0], gausei=5)
    dt5 = time.perf_counter() - t0
    # NORM slice
    n = max(threads, 4)
    fns, _ = cases.hertz91_fn(4096, fn0=FN0)
    t0 = time.perf_counter()
    O.norm_batch(g, cases.STEEL["gg"], cases.STEEL["poiss"], 1, fns[:n], maxgs=MAXGS, maxin=MAXIN, eps=EPS, nthreads=threads)
    dt = time.perf_counter() - t0
    ns = int(max(n, min(4096, n * 0.3 * budget_s / max(dt, 1e-3))))
    t0 = time.perf_counter()
    r = O.norm_batch(g, cases.STEEL["gg"], cases.STEEL["poiss"], 1, fns[:ns], maxgs=MAXGS, maxin=MAXIN, eps=EPS, nthreads=threads)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    el1 = (r["pn"][0] > 0).astype(np.int32)
    ps1 = np.zeros((3, g["mx"] * g["my"])); ps1[2] = r["pn"][0]
    O.subsurf_block(g["mx"], g["my"], g["dx"], g["dy"], cases.STEEL["gg"], cases.STEEL["poiss"], el1, ps1, SUBS_Z)
    dts = time.perf_counter() - t1
    gm = mbench_grid()
    t2 = time.perf_counter()
    rr = O.contac(gm, cases.STEEL["gg"], cases.STEEL["poiss"], tang=3, norm=0, force3=0, pen=gm["pen"], cksi=0.0005, ceta=0.0,
                  cphi=0.0003, fstat=0.3, fkin=0.3, maxgs=999, maxin=100, maxnr=30, maxout=1, eps=1e-5, nn=gm["nn"], chi=0.0,
                  dq=0.1, gausei=0)
    dtr = time.perf_counter() - t2
    t2 = time.perf_counter()
    rg = O.contac(gm, cases.STEEL["gg"], cases.STEEL["poiss"], tang=3, norm=0, force3=0, pen=gm["pen"], cksi=0.0005, ceta=0.0,
                  cphi=0.0003, fstat=0.3, fkin=0.3, maxgs=999, maxin=100, maxnr=30, maxout=1, eps=1e-5, nn=gm["nn"], chi=0.0,
                  dq=0.1, gausei=5, gd=GD_RECORD)
    dtg = time.perf_counter() - t2
    from tests import inp_oracle
    from tests.cases import sequence
    t3 = time.perf_counter()
    inp_oracle.run_cases(sequence("spence71")["cases"][:12])
    dt71 = time.perf_counter() - t3
    return {"value": nc / dtc, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d complete hertz-91 contact cases (N=1, T=3, G=0: NormCG + SteadyGS, mean %.0f sweeps), the first of the seeded "
                      "sweep, %.1f s wall on one core; CPU restatement of the reference algorithm (oracle/, gcc -O3 -march=x86-64-v3 "
                      "-ffp-contract=off), own mixed-radix FFT instead of MKL" % (nc, float(np.mean(itgs)), dtc),
            "gdsteady_cases_per_s": 1.0 / dt5, "gdsteady_sample": "first case with G=5, %d iterations, %.2f s" % (r5["itgs_tang"], dt5),
            "norm_only_solves_per_s": ns / dt, "norm_only_sample": "%d hertz-91 normal problems, %.1f s" % (ns, dt),
            "norm_only_mean_itcg": float(r["itcg"].mean()),
            "subsurf_cases_per_s": 1.0 / dts, "spence71_first12_contact_s": dt71,
            "rolling_cases_per_s": 1.0 / dtr, "rolling_sample": "tang_problm_1c creepages on mbench 71x81, T=3 SteadyGS, eps 1e-5, "
                                                              "%d sweeps, %.2f s" % (rr["itgs_tang"], dtr),
            "rolling_gdsteady_cases_per_s": 1.0 / dtg,
            "rolling_gdsteady_sample": "same case, G=5 GDsteady, eps 1e-5, %d iterations, %.2f s" % (rg["itgs_tang"], dtg),
            "fft_calibration": fft_calibration()}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Fortran/MKL build is impossible here) on
    all host threads, one complete hertz-91 contact case per thread at a time (test_table.f90:196-292 pattern)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    threads = host_cores()
    from oracle import oracle as O
    per_step = threads                  # bounded sample: one case per thread and step (a case takes seconds on a core)
    draws = hertz91_draws(per_step * (args.steps + args.warmup))
    pool = ThreadPoolExecutor(max_workers=threads)          # co_contac runs outside the GIL (ctypes)

    def one(k):
        t0 = time.perf_counter()
        rs = list(pool.map(lambda d: hertz91_oracle_case(O, d)["ierror"], draws[k * per_step:(k + 1) * per_step]))
        assert all(e == 0 for e in rs)
        return time.perf_counter() - t0
    for k in range(args.warmup):
        one(k)
    ts = [one(args.warmup + k) for k in range(args.steps)]
    total = float(np.sum(ts))
    value = per_step * args.steps / total
    # the NORM slice beside it (round-1 reference figure)
    g = cases.HERTZ91
    fns, _ = cases.hertz91_fn(4096, fn0=FN0)
    t0 = time.perf_counter()
    O.norm_batch(g, cases.STEEL["gg"], cases.STEEL["poiss"], 1, fns[:8 * threads], maxgs=MAXGS, maxin=MAXIN, eps=EPS, nthreads=threads)
    norm_value = 8 * threads / (time.perf_counter() - t0)
    nsm = 148
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(4 * nsm, args.gpus),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": "%d complete hertz-91 contact cases per step (N=1, T=3, G=0), one case per thread (test_table.f90 "
                                      "pattern), CPU restatement of the reference algorithm (oracle/, gcc -O3, own FFT instead of MKL)" % per_step,
                            "norm_only_solves_per_s": norm_value},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cases", type=int, default=0, help="norm_only leg: cases per GPU per step (default 8 x SM count)")
    ap.add_argument("--contact-cases", type=int, default=0, help="headline: complete contact cases per GPU per step (default 4 x SM count, <= 999)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg (0 = skip)")
    ap.add_argument("--no-sweep4096", dest="sweep4096", action="store_false",
                    help="skip BASELINE config 5 at full size (4096 rolling cases sharded over the ranks, ~5 s on one GPU)")
    ap.add_argument("--skip-extra", action="store_true", help="skip the rolling-sweep and 575x647 legs")
    ap.add_argument("--inlib-devices", type=int, default=0, help="single process: also run the 4096-case sweep with the library's own "
                    "scheduler spreading each cntc_calculate_batch call over this many GPUs (cb200_set_devices)")
    ap.add_argument("--gs-4c", action="store_true", help="steadygs_large leg: also tang_problm_4c with G=0 (minutes)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == "__main__":
    main()
