"""The batched case scheduler behind cntc_calculate across the GPUs of one box, driven by a C-only caller (tests/cabi/
multi_gpu_caller.c: gcc, the C header, the shared library -- no Python in the solving process, no torchrun).

cb200_set_devices(n, devs) makes cntc_calculate_batch cut a batch into contiguous shards, one host thread per device.  The
single-GPU test runs both shards on device 0 (the same code path: two scheduler threads, per-device engine state behind its
locks); the two-GPU test needs two devices and is skipped on a one-GPU box.  Both compare the spread batch with the
one-device batch bit for bit."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "contact_b200", "lib")
EXE = os.path.join(HERE, "cabi", "_build", "multi_gpu_caller")


def build_caller():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(HERE, "cabi", "multi_gpu_caller.c")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(LIBDIR, "libcontact_addon_b200.so"))):
        subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-L", LIBDIR, "-lcontact_addon_b200",
                               "-Wl,-rpath," + LIBDIR, "-lm", "-o", EXE])
    return EXE


def test_c_caller_builds_against_header_and_library():
    """gcc compiles the caller against include/contact_addon_b200.h and links it with the library (CPU: no run)."""
    assert os.path.exists(build_caller())


def run_caller(devs, ncase):
    r = subprocess.run([build_caller(), devs, str(ncase)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


@pytest.mark.gpu
def test_c_caller_two_shards_on_one_gpu():
    out = run_caller("0,0", 96)
    assert "2 device(s) in use" in out and "0 of 192 result arrays differ" in out, out


@pytest.mark.gpu
def test_c_caller_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    out = run_caller("0,1", 296)
    print(out)
    assert "2 device(s) in use" in out and "0 of 592 result arrays differ" in out, out
