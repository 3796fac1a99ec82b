// fft_radix.cuh -- register-resident radix-R DFT butterflies in FP64 for sm_100a.
//
// Part of the B200-native replacement for the MKL DFTI calls of the reference's influence product
// (/root/reference/src/m_aijpj.f90:548-591, 841-866, 907, 945, 970).  The reference sizes its transforms with
// opt_fft_size (m_aijpj.f90:1022-1119) => lengths 2^a 3^b 5^c 7^d, so radices 2,3,4,5,7,8 (+ composite 9, 16) cover
// every product; fft_makePrec's un-optimised sizes go through the dense DFT path (prec.cu) instead.
#pragma once
#include <cuda_runtime.h>

#define CB_HD __host__ __device__ __forceinline__
#define CB_HNI __host__ __device__ __noinline__

namespace cb200 {

typedef double2 cd;

CB_HD cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
CB_HD cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
CB_HD cd cmul(cd a, cd b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
CB_HD cd cmulc(cd a, cd b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
CB_HD cd cconj(cd a) { return make_double2(a.x, -a.y); }
CB_HD cd cscale(cd a, double s) { return make_double2(a.x * s, a.y * s); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> CB_HD cd mul_mi(cd a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

// X[q'] = sum_q x[q] w^(q q'),  w = exp(-2 pi i / R) (forward) or exp(+2 pi i / R) (INV). In place, natural order.
template <int R, bool INV> struct Dft;

template <bool INV> struct Dft<2, INV> {
    static CB_HD void run(cd *x) {
        cd a = x[0], b = x[1];
        x[0] = cadd(a, b); x[1] = csub(a, b);
    }
};

template <bool INV> struct Dft<3, INV> {
    static CB_HD void run(cd *x) {
        const double s = INV ? 0.86602540378443864676 : -0.86602540378443864676;   // imag of w
        cd t = cadd(x[1], x[2]), d = csub(x[1], x[2]);
        cd h = make_double2(x[0].x - 0.5 * t.x, x[0].y - 0.5 * t.y);
        cd e = make_double2(-s * d.y, s * d.x);                                     // i*s*d
        x[0] = cadd(x[0], t); x[1] = cadd(h, e); x[2] = csub(h, e);
    }
};

template <bool INV> struct Dft<4, INV> {
    static CB_HD void run(cd *x) {
        cd s0 = cadd(x[0], x[2]), s1 = csub(x[0], x[2]);
        cd s2 = cadd(x[1], x[3]), s3 = mul_mi<INV>(csub(x[1], x[3]));
        x[0] = cadd(s0, s2); x[2] = csub(s0, s2);
        x[1] = cadd(s1, s3); x[3] = csub(s1, s3);
    }
};

template <bool INV> struct Dft<5, INV> {
    static CB_HD void run(cd *x) {
        const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
        const double s1 = INV ? 0.95105651629515357212 : -0.95105651629515357212;
        const double s2 = INV ? 0.58778525229247312917 : -0.58778525229247312917;
        cd a1 = cadd(x[1], x[4]), b1 = csub(x[1], x[4]);
        cd a2 = cadd(x[2], x[3]), b2 = csub(x[2], x[3]);
        cd p1 = make_double2(x[0].x + c1 * a1.x + c2 * a2.x, x[0].y + c1 * a1.y + c2 * a2.y);
        cd p2 = make_double2(x[0].x + c2 * a1.x + c1 * a2.x, x[0].y + c2 * a1.y + c1 * a2.y);
        cd q1 = make_double2(-(s1 * b1.y + s2 * b2.y), s1 * b1.x + s2 * b2.x);      // i (s1 b1 + s2 b2)
        cd q2 = make_double2(-(s2 * b1.y - s1 * b2.y), s2 * b1.x - s1 * b2.x);      // i (s2 b1 - s1 b2)
        x[0] = make_double2(x[0].x + a1.x + a2.x, x[0].y + a1.y + a2.y);
        x[1] = cadd(p1, q1); x[4] = csub(p1, q1);
        x[2] = cadd(p2, q2); x[3] = csub(p2, q2);
    }
};

template <bool INV> struct Dft<7, INV> {
    static CB_HD void run(cd *x) {
        const double c1 = 0.62348980185873353053, c2 = -0.22252093395631440429, c3 = -0.90096886790241912624;
        const double g = INV ? 1.0 : -1.0;
        const double s1 = g * 0.78183148246802980871, s2 = g * 0.97492791218182360702, s3 = g * 0.43388373911755812048;
        cd a1 = cadd(x[1], x[6]), b1 = csub(x[1], x[6]);
        cd a2 = cadd(x[2], x[5]), b2 = csub(x[2], x[5]);
        cd a3 = cadd(x[3], x[4]), b3 = csub(x[3], x[4]);
        // X[k] = x0 + sum_j a_j cos(2 pi jk/7) + i * sum_j b_j sin(+-2 pi jk/7)
        cd p1 = make_double2(x[0].x + c1 * a1.x + c2 * a2.x + c3 * a3.x, x[0].y + c1 * a1.y + c2 * a2.y + c3 * a3.y);
        cd p2 = make_double2(x[0].x + c2 * a1.x + c3 * a2.x + c1 * a3.x, x[0].y + c2 * a1.y + c3 * a2.y + c1 * a3.y);
        cd p3 = make_double2(x[0].x + c3 * a1.x + c1 * a2.x + c2 * a3.x, x[0].y + c3 * a1.y + c1 * a2.y + c2 * a3.y);
        // k=1: s1 b1 + s2 b2 + s3 b3 ; k=2: s2 b1 - s3 b2 - s1 b3 ; k=3: s3 b1 - s1 b2 + s2 b3
        cd r1 = make_double2(s1 * b1.x + s2 * b2.x + s3 * b3.x, s1 * b1.y + s2 * b2.y + s3 * b3.y);
        cd r2 = make_double2(s2 * b1.x - s3 * b2.x - s1 * b3.x, s2 * b1.y - s3 * b2.y - s1 * b3.y);
        cd r3 = make_double2(s3 * b1.x - s1 * b2.x + s2 * b3.x, s3 * b1.y - s1 * b2.y + s2 * b3.y);
        cd q1 = make_double2(-r1.y, r1.x), q2 = make_double2(-r2.y, r2.x), q3 = make_double2(-r3.y, r3.x);
        x[0] = make_double2(x[0].x + a1.x + a2.x + a3.x, x[0].y + a1.y + a2.y + a3.y);
        x[1] = cadd(p1, q1); x[6] = csub(p1, q1);
        x[2] = cadd(p2, q2); x[5] = csub(p2, q2);
        x[3] = cadd(p3, q3); x[4] = csub(p3, q3);
    }
};

template <bool INV> struct Dft<8, INV> {
    static CB_HD void run(cd *x) {
        const double h = 0.70710678118654752440;
        // two radix-4 on even / odd samples, then twiddles w8^k
        cd e[4] = { x[0], x[2], x[4], x[6] }, o[4] = { x[1], x[3], x[5], x[7] };
        Dft<4, INV>::run(e); Dft<4, INV>::run(o);
        // w8^1 = h(1 -+ i), w8^2 = -+i, w8^3 = h(-1 -+ i)
        cd o1 = INV ? make_double2(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y))
                    : make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
        cd o2 = mul_mi<INV>(o[2]);
        cd o3 = INV ? make_double2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y))
                    : make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
        x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
        x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
        x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
        x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
    }
};

template <bool INV> struct Dft<9, INV> {
    static CB_HD void run(cd *x) {
        // 9 = 3 x 3: columns n = 3 n1 + n2, k = k1 + 3 k2
        const double g = INV ? 1.0 : -1.0;
        const cd w1 = make_double2(0.76604444311897803520, g * 0.64278760968653932632);    // w9^1
        const cd w2 = make_double2(0.17364817766693034885, g * 0.98480775301220805937);    // w9^2
        const cd w4 = make_double2(-0.93969262078590838405, g * 0.34202014332566873304);   // w9^4
        cd a[3] = { x[0], x[3], x[6] }, b[3] = { x[1], x[4], x[7] }, c[3] = { x[2], x[5], x[8] };
        Dft<3, INV>::run(a); Dft<3, INV>::run(b); Dft<3, INV>::run(c);
        b[1] = cmul(b[1], w1); b[2] = cmul(b[2], w2);
        c[1] = cmul(c[1], w2); c[2] = cmul(c[2], w4);
#pragma unroll
        for (int k1 = 0; k1 < 3; k1++) {
            cd t[3] = { a[k1], b[k1], c[k1] };
            Dft<3, INV>::run(t);
            x[k1] = t[0]; x[k1 + 3] = t[1]; x[k1 + 6] = t[2];
        }
    }
};

template <bool INV> struct Dft<16, INV> {
    static CB_HD void run(cd *x) {
        // 16 = 4 x 4: n = 4 n1 + n2, k = k1 + 4 k2
        const double g = INV ? 1.0 : -1.0;
        const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
        cd y[4][4];
#pragma unroll
        for (int n2 = 0; n2 < 4; n2++) {
            cd t[4] = { x[n2], x[4 + n2], x[8 + n2], x[12 + n2] };
            Dft<4, INV>::run(t);
            y[n2][0] = t[0]; y[n2][1] = t[1]; y[n2][2] = t[2]; y[n2][3] = t[3];
        }
        // twiddles w16^(n2*k1)
        const cd w1 = make_double2(c1, g * s1), w2 = make_double2(h, g * h), w3 = make_double2(s1, g * c1);
        const cd w6 = make_double2(-h, g * h), w9 = make_double2(-c1, -g * s1);
        y[1][1] = cmul(y[1][1], w1); y[1][2] = cmul(y[1][2], w2); y[1][3] = cmul(y[1][3], w3);
        y[2][1] = cmul(y[2][1], w2); y[2][2] = mul_mi<INV>(y[2][2]); y[2][3] = cmul(y[2][3], w6);
        y[3][1] = cmul(y[3][1], w3); y[3][2] = cmul(y[3][2], w6); y[3][3] = cmul(y[3][3], w9);
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) {
            cd t[4] = { y[0][k1], y[1][k1], y[2][k1], y[3][k1] };
            Dft<4, INV>::run(t);
            x[k1] = t[0]; x[k1 + 4] = t[1]; x[k1 + 8] = t[2]; x[k1 + 12] = t[3];
        }
    }
};

template <bool INV> struct Dft<6, INV> {
    static CB_HD void run(cd *x) {
        // 6 = 2 x 3 (n = 2 n1 + n2, k = k1 + 3 k2): two radix-3 on even / odd samples, twiddles w6^k1, radix-2 across
        const double g = INV ? 1.0 : -1.0;
        cd e[3] = { x[0], x[2], x[4] }, o[3] = { x[1], x[3], x[5] };
        Dft<3, INV>::run(e); Dft<3, INV>::run(o);
        const cd w1 = make_double2(0.5, g * 0.86602540378443864676), w2 = make_double2(-0.5, g * 0.86602540378443864676);
        o[1] = cmul(o[1], w1); o[2] = cmul(o[2], w2);
        x[0] = cadd(e[0], o[0]); x[3] = csub(e[0], o[0]);
        x[1] = cadd(e[1], o[1]); x[4] = csub(e[1], o[1]);
        x[2] = cadd(e[2], o[2]); x[5] = csub(e[2], o[2]);
    }
};

template <bool INV> struct Dft<12, INV> {
    static CB_HD void run(cd *x) {
        // 12 = 3 x 4: n = 3 n1 + n2 (n1 < 4, n2 < 3), k = k1 + 4 k2 (k1 < 4, k2 < 3)
        const double g = INV ? 1.0 : -1.0;
        const double c30 = 0.86602540378443864676;
        cd y[3][4];
#pragma unroll
        for (int n2 = 0; n2 < 3; n2++) {
            cd t[4] = { x[n2], x[3 + n2], x[6 + n2], x[9 + n2] };
            Dft<4, INV>::run(t);
            y[n2][0] = t[0]; y[n2][1] = t[1]; y[n2][2] = t[2]; y[n2][3] = t[3];
        }
        // twiddles w12^(n2 k1)
        const cd w1 = make_double2(c30, g * 0.5), w2 = make_double2(0.5, g * c30), w4 = make_double2(-0.5, g * c30);
        y[1][1] = cmul(y[1][1], w1); y[1][2] = cmul(y[1][2], w2); y[1][3] = mul_mi<INV>(y[1][3]);          // w12^3 = -+i
        y[2][1] = cmul(y[2][1], w2); y[2][2] = cmul(y[2][2], w4); y[2][3] = make_double2(-y[2][3].x, -y[2][3].y);  // w12^6 = -1
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) {
            cd t[3] = { y[0][k1], y[1][k1], y[2][k1] };
            Dft<3, INV>::run(t);
            x[k1] = t[0]; x[k1 + 4] = t[1]; x[k1 + 8] = t[2];
        }
    }
};

// ---- input-pruned butterflies: the upper half of the inputs x[R/2 .. R-1] is known to be zero (first stage of a
// zero-padded transform: the tractions fill at most half of the padded array, m_aijpj.f90:932-939).  Same operations as
// Dft<R> on the non-zero terms, the additions of zeros left out => identical results (up to the sign of a zero).
template <bool INV> CB_HD void dft4_ab00(cd a, cd b, cd *t)
{
    const cd s3 = mul_mi<INV>(b);
    t[0] = cadd(a, b); t[2] = csub(a, b); t[1] = cadd(a, s3); t[3] = csub(a, s3);
}

template <int R, bool INV> struct DftHalfIn { static const bool ok = false; static CB_HD void run(cd *x) { Dft<R, INV>::run(x); } };

template <bool INV> struct DftHalfIn<4, INV> {
    static const bool ok = true;
    static CB_HD void run(cd *x) { cd t[4]; dft4_ab00<INV>(x[0], x[1], t); x[0] = t[0]; x[1] = t[1]; x[2] = t[2]; x[3] = t[3]; }
};

template <bool INV> struct DftHalfIn<8, INV> {
    static const bool ok = true;
    static CB_HD void run(cd *x) {
        const double h = 0.70710678118654752440;
        cd e[4], o[4];
        dft4_ab00<INV>(x[0], x[2], e); dft4_ab00<INV>(x[1], x[3], o);
        cd o1 = INV ? make_double2(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y))
                    : make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
        cd o2 = mul_mi<INV>(o[2]);
        cd o3 = INV ? make_double2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y))
                    : make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
        x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
        x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
        x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
        x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
    }
};

template <bool INV> struct DftHalfIn<12, INV> {
    static const bool ok = true;
    static CB_HD void run(cd *x) {
        const double g = INV ? 1.0 : -1.0;
        const double c30 = 0.86602540378443864676;
        cd y[3][4];
#pragma unroll
        for (int n2 = 0; n2 < 3; n2++) dft4_ab00<INV>(x[n2], x[3 + n2], y[n2]);
        const cd w1 = make_double2(c30, g * 0.5), w2 = make_double2(0.5, g * c30), w4 = make_double2(-0.5, g * c30);
        y[1][1] = cmul(y[1][1], w1); y[1][2] = cmul(y[1][2], w2); y[1][3] = mul_mi<INV>(y[1][3]);
        y[2][1] = cmul(y[2][1], w2); y[2][2] = cmul(y[2][2], w4); y[2][3] = make_double2(-y[2][3].x, -y[2][3].y);
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) {
            cd t[3] = { y[0][k1], y[1][k1], y[2][k1] };
            Dft<3, INV>::run(t);
            x[k1] = t[0]; x[k1 + 4] = t[1]; x[k1 + 8] = t[2];
        }
    }
};

template <bool INV> struct DftHalfIn<16, INV> {
    static const bool ok = true;
    static CB_HD void run(cd *x) {
        const double g = INV ? 1.0 : -1.0;
        const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
        cd y[4][4];
#pragma unroll
        for (int n2 = 0; n2 < 4; n2++) dft4_ab00<INV>(x[n2], x[4 + n2], y[n2]);
        const cd w1 = make_double2(c1, g * s1), w2 = make_double2(h, g * h), w3 = make_double2(s1, g * c1);
        const cd w6 = make_double2(-h, g * h), w9 = make_double2(-c1, -g * s1);
        y[1][1] = cmul(y[1][1], w1); y[1][2] = cmul(y[1][2], w2); y[1][3] = cmul(y[1][3], w3);
        y[2][1] = cmul(y[2][1], w2); y[2][2] = mul_mi<INV>(y[2][2]); y[2][3] = cmul(y[2][3], w6);
        y[3][1] = cmul(y[3][1], w3); y[3][2] = cmul(y[3][2], w6); y[3][3] = cmul(y[3][3], w9);
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) {
            cd t[4] = { y[0][k1], y[1][k1], y[2][k1], y[3][k1] };
            Dft<4, INV>::run(t);
            x[k1] = t[0]; x[k1 + 4] = t[1]; x[k1 + 8] = t[2]; x[k1 + 12] = t[3];
        }
    }
};

}  // namespace cb200
