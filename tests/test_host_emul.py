"""The product's device algorithm (contact_b200/csrc/fftconv.cuh) stepped through on the CPU against the oracle.

tests/host_emul/emul_conv.cu compiles the very same phase functions for the host and runs every thread id
sequentially, so indexing, digit-reversed coefficient layout, split/merge steps and pruning are validated without GPU.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_emul", "emul_conv.cu")
SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libemul.so")


@pytest.fixture(scope="module")
def emul():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "contact_b200", "csrc", f) for f in ("fftconv.cuh", "fftconv_warp.cuh", "fft_radix.cuh", "plan.h", "conv_sequence.inc")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call([nvcc, "-O1", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-Wno-deprecated-gpu-targets",
                               "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.emul_radix_error.restype = C.c_double
    return lib


@pytest.mark.parametrize("R", [2, 3, 4, 5, 6, 7, 8, 9, 12, 16])
def test_radix_butterflies(emul, R):
    assert emul.emul_radix_error(R, 0) < 2e-12
    assert emul.emul_radix_error(R, 1) < 2e-12


@pytest.mark.parametrize("mx,my", [(19, 19), (91, 91), (71, 81), (43, 93), (11, 11), (35, 35), (1, 1), (5, 1), (1, 7),
                                   (3, 2), (8, 8), (25, 60), (63, 64)])
def test_emulated_product_matches_direct_sum(emul, mx, my):
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    rng = np.random.default_rng(1)
    m = O.mater()
    cs, cv, csv, ms = O.sgencr(m, mx, my, 0.1, 0.13)
    npot = mx * my
    el = (rng.random(npot) < 0.6).astype(np.int32)
    p = np.zeros((3, npot)); p[2] = rng.standard_normal(npot) * el
    full = O.EldivBuf(mx, my, np.ones(npot, np.int32))
    ud = np.zeros((3, npot))
    O.vecaijpj_direct(full, -9, ud, 3, p, 3, cs, pel=full)
    blk = np.ascontiguousarray(cs.block(3, 3))
    for mask_mode, add in ((0, 0), (1, 0), (1, 1)):
        ue = np.full(npot, 7.0)
        pz = p[2].copy()
        rc = emul.emul_conv(mx, my, pz.ctypes.data_as(dp), blk.ctypes.data_as(dp), mx, my, C.c_double(cs.ga_inv),
                            el.ctypes.data_as(ip), mask_mode, add, ue.ctypes.data_as(dp), 37)
        assert rc == 0
        exp = np.full(npot, 7.0)
        sel = np.ones(npot, bool) if mask_mode == 0 else el > 0
        exp[sel] = ud[2][sel] + (7.0 if add else 0.0)
        assert np.abs(ue - exp).max() < 1e-13 * max(1.0, np.abs(ud[2]).max()) + 1e-12
        # the warp-scheduled passes (fftconv_warp.cuh, what the device runs) regroup the same butterflies: bit-identical
        for nthr in (384, 64):
            uw = np.full(npot, 7.0)
            rc = emul.emul_conv_warp(mx, my, pz.ctypes.data_as(dp), blk.ctypes.data_as(dp), mx, my, C.c_double(cs.ga_inv),
                                     el.ctypes.data_as(ip), mask_mode, add, uw.ctypes.data_as(dp), nthr)
            assert rc == 0
            assert np.array_equal(uw, ue)
    O.inflcf_free(cs, cv, csv, ms)


def test_plan_radices_cover_product_sizes(emul):
    info = (C.c_int * 32)()
    for mx, my in [(19, 19), (91, 91), (71, 81), (43, 93), (11, 11), (575, 647)]:
        assert emul.emul_plan_info(mx, my, info) == 0
        fx, fy, nsx, nsy = info[0], info[1], info[6], info[7]
        px = 1
        for i in range(nsx):
            px *= info[8 + i]
        py = 1
        for i in range(nsy):
            py *= info[16 + i]
        assert px == fx and py == 2 * fy
    assert info[5] == 0          # 575x647 does not fit one CTA's shared memory: reported, not silently wrong
