"""Oracle FFT against numpy (pocketfft) and torch-CPU (oneMKL, the reference's FFT family)."""
import numpy as np
import pytest

from oracle import oracle as O


@pytest.mark.parametrize("n1,n2", [(40, 40), (192, 192), (38, 38), (182, 182), (24, 30), (2, 2), (90, 192), (70, 22),
                                   (6, 1), (144, 162), (26, 14)])
def test_fft2_matches_numpy_and_mkl(n1, n2):
    import torch
    a = np.random.default_rng(n1 * 1000 + n2).standard_normal((n2, n1))
    A = O.fft2_r2c(a)
    R = np.fft.rfft2(a)
    T = torch.fft.rfft2(torch.from_numpy(a)).numpy()
    s = np.abs(R).max()
    assert np.abs(A - R).max() < 1e-13 * s
    assert np.abs(A - T).max() < 1e-13 * s
    b = O.fft2_c2r(R, n1, 1.0 / (n1 * n2))
    assert np.abs(b - a).max() < 1e-13


def test_opt_fft_size_values():
    # sizes chosen by opt_fft_size (m_aijpj.f90:1022-1119)
    exp = {19: 20, 91: 96, 575: 576, 647: 648, 43: 45, 93: 96, 71: 72, 81: 81, 11: 12, 35: 36, 1: 1, 2: 2, 3: 3,
           143: 144, 161: 162, 287: 288, 323: 324}
    for n, f in exp.items():
        assert O.opt_fft_size(n) == f
    for n in range(1, 700):
        f = O.opt_fft_size(n)
        assert f >= n
        m = f
        for p in (2, 3, 5, 7):
            while m % p == 0:
                m //= p
        assert m == 1
