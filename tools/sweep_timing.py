"""Per-chunk wall-clock split of the 4096-case rolling sweep (bench.py leg sweep4096) on one GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import bench  # noqa: E402
import contact_b200 as cb  # noqa: E402

torch.cuda.set_device(0)
for rep in range(2):
    r = bench.sweep4096_leg(cb, 0, 1)
    print(json.dumps({"rep": rep, "s": r["s"], "cases_per_s": r["cases"] / r["s"], "kernel_ms": r["solver_kernel_ms"], "chunks": r["chunks"]}))
