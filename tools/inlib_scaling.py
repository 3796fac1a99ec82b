"""Strong scaling of the library's own multi-GPU scheduler: the 4096-case rolling sweep of BASELINE config 5 (mbench 71x81, T=3,
G=5, chunks of <= 888 result elements per cntc_calculate_batch call) from ONE process with cb200_set_devices(n), n = 1, 2, 4, 8 as
far as the box has devices; and the hertz-91 contact batch (T=3, G=0, 888 cases per call).  usage: python tools/inlib_scaling.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import contact_b200 as cb
import bench

ll = cb.lowlevel
ndev = torch.cuda.device_count()
out = {"devices_on_box": ndev, "sweep4096": {}, "hertz91_contact_888": {}}
for n in (1, 2, 4, 8):
    if n > ndev:
        break
    assert ll.set_devices(n) == n
    bench.sweep4096_leg(cb, 0, 1, total=2 * 888)                 # warms the per-device coefficient caches and work-space pools
    r = bench.sweep4096_leg(cb, 0, 1)
    out["sweep4096"][n] = {"cases_per_s": r["cases"] / r["s"], "s": r["s"], "errors": r["errors"]}
    ires = list(range(1, 889))
    draws = bench.hertz91_draws(888)
    bench.hertz91_setup(cb, ires)
    sink = [None] * 888
    bench.hertz91_step(cb, ires, draws, sink)
    t0 = time.perf_counter()
    ierr, _ = bench.hertz91_step(cb, ires, draws, sink)
    dt = time.perf_counter() - t0
    out["hertz91_contact_888"][n] = {"solves_per_s": 888 / dt, "s": dt, "errors": int((np.asarray(ierr) < 0).sum())}
    for ire in ires:
        cb.cntc_finalize(ire)
ll.set_devices(1)
print(json.dumps(out))
