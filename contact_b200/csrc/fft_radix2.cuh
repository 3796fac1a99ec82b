// fft_radix2.cuh -- further register-resident butterflies for the warp-resident product (fftconv2.cuh):
// radix 1 (identity), 10 = 2 x 5, 18 = 2 x 9, and input-pruned radix 6 / 10 / 18 (upper half of the inputs zero).
// Same conventions as fft_radix.cuh: forward w = exp(-2 pi i / R), INV conjugates.  Constants to 20 digits.
// Part of the replacement for the MKL DFTI calls of /root/reference/src/m_aijpj.f90:548-591.
#pragma once
#include "fft_radix.cuh"

namespace cb200 {

template <bool INV> struct Dft<1, INV> { static CB_HD void run(cd *) {} };

// w_N^n = exp(-+ 2 pi i n / N) for the composite butterflies below (n is a compile-time constant after unrolling)

template <bool INV> CB_HD cd twc_6(int n)
{
    const double g = INV ? 1.0 : -1.0;
    switch (n) {
    case 1: return make_double2(0.50000000000000011102, g * 0.86602540378443859659);
    case 2: return make_double2(-0.49999999999999977796, g * 0.86602540378443870761);
    default: return make_double2(1.0, 0.0);
    }
}
template <bool INV> CB_HD cd twc_10(int n)
{
    const double g = INV ? 1.0 : -1.0;
    switch (n) {
    case 1: return make_double2(0.80901699437494745126, g * 0.58778525229247313710);
    case 2: return make_double2(0.30901699437494745126, g * 0.95105651629515353118);
    case 3: return make_double2(-0.30901699437494734024, g * 0.95105651629515364220);
    case 4: return make_double2(-0.80901699437494734024, g * 0.58778525229247324813);
    default: return make_double2(1.0, 0.0);
    }
}
template <bool INV> CB_HD cd twc_18(int n)
{
    const double g = INV ? 1.0 : -1.0;
    switch (n) {
    case 1: return make_double2(0.93969262078590842791, g * 0.34202014332566871291);
    case 2: return make_double2(0.76604444311897801345, g * 0.64278760968653925190);
    case 3: return make_double2(0.50000000000000011102, g * 0.86602540378443859659);
    case 4: return make_double2(0.17364817766693041445, g * 0.98480775301220802032);
    case 5: return make_double2(-0.17364817766693030343, g * 0.98480775301220802032);
    case 6: return make_double2(-0.49999999999999977796, g * 0.86602540378443870761);
    case 7: return make_double2(-0.76604444311897790243, g * 0.64278760968653947394);
    case 8: return make_double2(-0.93969262078590831688, g * 0.34202014332566887944);
    default: return make_double2(1.0, 0.0);
    }
}

// Dft<2r> from two Dft<r> on the even / odd samples (decimation in time): X[k] = E[k] + w^k O[k], X[k+r] = E[k] - w^k O[k]
#define CB_DFT_2R(N, r)                                                                     \
    template <bool INV> struct Dft<N, INV> {                                                \
        static CB_HD void run(cd *x) {                                                      \
            cd e[r], o[r];                                                                  \
            _Pragma("unroll") for (int n = 0; n < r; n++) { e[n] = x[2 * n]; o[n] = x[2 * n + 1]; } \
            Dft<r, INV>::run(e); Dft<r, INV>::run(o);                                       \
            x[0] = cadd(e[0], o[0]); x[r] = csub(e[0], o[0]);                               \
            _Pragma("unroll") for (int k = 1; k < r; k++) {                                 \
                const cd t = cmul(o[k], twc_##N<INV>(k));                                   \
                x[k] = cadd(e[k], t); x[k + r] = csub(e[k], t);                             \
            }                                                                               \
        }                                                                                   \
    };
CB_DFT_2R(10, 5)
CB_DFT_2R(18, 9)
#undef CB_DFT_2R

// input-pruned Dft<2r>, x[r..2r-1] = 0 (decimation in frequency): X[2k] = Dft_r(x)[k], X[2k+1] = Dft_r(x[n] w^n)[k]
#define CB_HALFIN_2R(N, r)                                                                  \
    template <bool INV> struct DftHalfIn<N, INV> {                                          \
        static const bool ok = true;                                                        \
        static CB_HD void run(cd *x) {                                                      \
            cd e[r], o[r];                                                                  \
            e[0] = x[0]; o[0] = x[0];                                                       \
            _Pragma("unroll") for (int n = 1; n < r; n++) { e[n] = x[n]; o[n] = cmul(x[n], twc_##N<INV>(n)); } \
            Dft<r, INV>::run(e); Dft<r, INV>::run(o);                                       \
            _Pragma("unroll") for (int k = 0; k < r; k++) { x[2 * k] = e[k]; x[2 * k + 1] = o[k]; } \
        }                                                                                   \
    };
CB_HALFIN_2R(6, 3)
CB_HALFIN_2R(10, 5)
CB_HALFIN_2R(18, 9)
#undef CB_HALFIN_2R

}  // namespace cb200
