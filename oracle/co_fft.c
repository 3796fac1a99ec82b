/*
 * co_fft.c -- ORACLE (test infrastructure, not product code).
 *
 * Double-precision FFTs for the CPU restatement of CONTACT's influence product.
 * The reference calls Intel MKL DFTI (un-vendored, version unpinned) at
 *   /root/reference/src/m_aijpj.f90:548-591, 633, 688 (fft_makePrec) and
 *   /root/reference/src/m_aijpj.f90:841-866, 907, 945, 970 (fft_VecAijPj):
 *   2-D real->complex forward (unscaled), complex->real backward scaled by 1/(n1*n2),
 *   CCE storage: half spectrum along the first (x, fastest) dimension, row stride n1/2+1.
 * MKL is absent, so this file restates the published DFT definition
 *   X[k] = sum_j x[j] exp(-2 pi i jk/n)  (forward),  conj kernel for backward
 * with a plain mixed-radix decimation-in-time algorithm (radix 4,2,3,5 + generic odd radix),
 * valid for ANY length (generic butterflies are O(p^2) per group) as fft_makePrec needs
 * un-optimised sizes such as 2*19, 2*91 = 2*7*13, 2*647.
 *
 * parity: cross-checked against numpy.fft (pocketfft) and torch.fft (oneMKL) in tests/test_oracle_fft.py.
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---------------------------------------------------------------- plan ---------------------------- */

static void factorize(int n, int *nfac, int *fac)
{
    int k = 0;
    while (n % 4 == 0) { fac[k++] = 4; n /= 4; }
    while (n % 2 == 0) { fac[k++] = 2; n /= 2; }
    for (int p = 3; p * p <= n; p += 2)
        while (n % p == 0) { fac[k++] = p; n /= p; }
    if (n > 1) fac[k++] = n;
    *nfac = k;
}

co_fftplan *co_fft_plan(int n)
{
    co_fftplan *pl = (co_fftplan *) calloc(1, sizeof(co_fftplan));
    pl->n = n;
    factorize(n, &pl->nfac, pl->fac);
    pl->tw = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) n);
    for (int k = 0; k < n; k++) {
        double ang = -2.0 * M_PI * (double) k / (double) n;
        pl->tw[k].re = cos(ang);
        pl->tw[k].im = sin(ang);
    }
    int pmax = 1;
    for (int i = 0; i < pl->nfac; i++) if (pl->fac[i] > pmax) pmax = pl->fac[i];
    pl->scratch = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) pmax);
    pl->work = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) n);
    return pl;
}

void co_fft_free(co_fftplan *pl)
{
    if (!pl) return;
    free(pl->tw); free(pl->scratch); free(pl->work); free(pl);
}

/* plan cache: plans are cheap but trig tables are not; keyed on n (single-threaded use per cache) */
co_fftplan *co_fft_cached(co_plancache *pc, int n)
{
    for (int i = 0; i < pc->nplan; i++)
        if (pc->plans[i]->n == n) return pc->plans[i];
    if (pc->nplan == CO_MAXPLAN) {           /* evict oldest */
        co_fft_free(pc->plans[0]);
        memmove(&pc->plans[0], &pc->plans[1], sizeof(pc->plans[0]) * (CO_MAXPLAN - 1));
        pc->nplan--;
    }
    pc->plans[pc->nplan] = co_fft_plan(n);
    return pc->plans[pc->nplan++];
}

void co_plancache_clear(co_plancache *pc)
{
    for (int i = 0; i < pc->nplan; i++) co_fft_free(pc->plans[i]);
    pc->nplan = 0;
}

/* ---------------------------------------------------------------- butterflies --------------------- */

#define CMUL(r, a, b) do { (r).re = (a).re*(b).re - (a).im*(b).im; (r).im = (a).re*(b).im + (a).im*(b).re; } while (0)

/* out[0..m*p) holds p interleaved sub-transforms of length m (out[q*m + k]); combine in place.
 * tw has stride fstride: w_n^(fstride*j). isign = +1 forward (tables hold exp(-i..)), -1 backward. */
static void bfly2(co_cplx *out, int fstride, const co_fftplan *pl, int m, int isign)
{
    const co_cplx *tw = pl->tw;
    for (int k = 0; k < m; k++) {
        co_cplx w = tw[k * fstride], t;
        if (isign < 0) w.im = -w.im;
        CMUL(t, out[m + k], w);
        out[m + k].re = out[k].re - t.re;  out[m + k].im = out[k].im - t.im;
        out[k].re += t.re;                 out[k].im += t.im;
    }
}

static void bfly4(co_cplx *out, int fstride, const co_fftplan *pl, int m, int isign)
{
    const co_cplx *tw = pl->tw;
    for (int k = 0; k < m; k++) {
        co_cplx w1 = tw[k * fstride], w2 = tw[2 * k * fstride], w3 = tw[3 * k * fstride];
        if (isign < 0) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
        co_cplx a = out[k], b, c, d;
        CMUL(b, out[m + k], w1);
        CMUL(c, out[2 * m + k], w2);
        CMUL(d, out[3 * m + k], w3);
        co_cplx s0 = { a.re + c.re, a.im + c.im }, s1 = { a.re - c.re, a.im - c.im };
        co_cplx s2 = { b.re + d.re, b.im + d.im }, s3 = { b.re - d.re, b.im - d.im };
        out[k].re         = s0.re + s2.re;  out[k].im         = s0.im + s2.im;
        out[2 * m + k].re = s0.re - s2.re;  out[2 * m + k].im = s0.im - s2.im;
        if (isign > 0) {        /* multiply s3 by -i */
            out[m + k].re     = s1.re + s3.im;  out[m + k].im     = s1.im - s3.re;
            out[3 * m + k].re = s1.re - s3.im;  out[3 * m + k].im = s1.im + s3.re;
        } else {                /* multiply s3 by +i */
            out[m + k].re     = s1.re - s3.im;  out[m + k].im     = s1.im + s3.re;
            out[3 * m + k].re = s1.re + s3.im;  out[3 * m + k].im = s1.im - s3.re;
        }
    }
}

static void bfly3(co_cplx *out, int fstride, const co_fftplan *pl, int m, int isign)
{
    const co_cplx *tw = pl->tw;
    const double s60 = isign * -0.86602540378443864676;    /* imag of exp(-+2 pi i/3) */
    for (int k = 0; k < m; k++) {
        co_cplx w1 = tw[k * fstride], w2 = tw[2 * k * fstride];
        if (isign < 0) { w1.im = -w1.im; w2.im = -w2.im; }
        co_cplx a = out[k], b, c;
        CMUL(b, out[m + k], w1);
        CMUL(c, out[2 * m + k], w2);
        co_cplx s = { b.re + c.re, b.im + c.im }, d = { b.re - c.re, b.im - c.im };
        co_cplx h = { a.re - 0.5 * s.re, a.im - 0.5 * s.im };
        out[k].re = a.re + s.re;  out[k].im = a.im + s.im;
        /* X1 = h + i*s60*d , X2 = h - i*s60*d */
        out[m + k].re     = h.re - s60 * d.im;  out[m + k].im     = h.im + s60 * d.re;
        out[2 * m + k].re = h.re + s60 * d.im;  out[2 * m + k].im = h.im - s60 * d.re;
    }
}

static void bfly5(co_cplx *out, int fstride, const co_fftplan *pl, int m, int isign)
{
    const co_cplx *tw = pl->tw;
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
    const double s1 = isign * -0.95105651629515357212, s2 = isign * -0.58778525229247312917;
    for (int k = 0; k < m; k++) {
        co_cplx w, x0 = out[k], x1, x2, x3, x4;
        w = tw[k * fstride];     if (isign < 0) w.im = -w.im;  CMUL(x1, out[m + k], w);
        w = tw[2 * k * fstride]; if (isign < 0) w.im = -w.im;  CMUL(x2, out[2 * m + k], w);
        w = tw[3 * k * fstride]; if (isign < 0) w.im = -w.im;  CMUL(x3, out[3 * m + k], w);
        w = tw[4 * k * fstride]; if (isign < 0) w.im = -w.im;  CMUL(x4, out[4 * m + k], w);
        co_cplx a1 = { x1.re + x4.re, x1.im + x4.im }, b1 = { x1.re - x4.re, x1.im - x4.im };
        co_cplx a2 = { x2.re + x3.re, x2.im + x3.im }, b2 = { x2.re - x3.re, x2.im - x3.im };
        out[k].re = x0.re + a1.re + a2.re;  out[k].im = x0.im + a1.im + a2.im;
        co_cplx p1 = { x0.re + c1 * a1.re + c2 * a2.re, x0.im + c1 * a1.im + c2 * a2.im };
        co_cplx p2 = { x0.re + c2 * a1.re + c1 * a2.re, x0.im + c2 * a1.im + c1 * a2.im };
        /* q1 = i*(s1*b1 + s2*b2), q2 = i*(s2*b1 - s1*b2) */
        co_cplx q1 = { -(s1 * b1.im + s2 * b2.im), s1 * b1.re + s2 * b2.re };
        co_cplx q2 = { -(s2 * b1.im - s1 * b2.im), s2 * b1.re - s1 * b2.re };
        out[m + k].re     = p1.re + q1.re;  out[m + k].im     = p1.im + q1.im;
        out[4 * m + k].re = p1.re - q1.re;  out[4 * m + k].im = p1.im - q1.im;
        out[2 * m + k].re = p2.re + q2.re;  out[2 * m + k].im = p2.im + q2.im;
        out[3 * m + k].re = p2.re - q2.re;  out[3 * m + k].im = p2.im - q2.im;
    }
}

static void bfly_generic(co_cplx *out, int fstride, const co_fftplan *pl, int m, int p, int isign)
{
    const co_cplx *tw = pl->tw;
    const int n = pl->n;
    co_cplx *scr = pl->scratch;
    for (int u = 0; u < m; u++) {
        for (int q = 0; q < p; q++) scr[q] = out[u + q * m];
        int k = u;
        for (int q1 = 0; q1 < p; q1++) {
            int twidx = 0;
            co_cplx acc = scr[0];
            for (int q = 1; q < p; q++) {
                twidx += fstride * k;
                if (twidx >= n) twidx %= n;
                co_cplx w = tw[twidx], t;
                if (isign < 0) w.im = -w.im;
                CMUL(t, scr[q], w);
                acc.re += t.re; acc.im += t.im;
            }
            out[k] = acc;
            k += m;
        }
    }
}

static void fft_work(co_cplx *out, const co_cplx *in, int fstride, int in_stride, const int *fac, int nleft,
                     const co_fftplan *pl, int isign)
{
    const int p = fac[0];
    const int m = nleft / p;
    if (m == 1) {
        for (int q = 0; q < p; q++) out[q] = in[(size_t) q * fstride * in_stride];
    } else {
        for (int q = 0; q < p; q++)
            fft_work(out + (size_t) q * m, in + (size_t) q * fstride * in_stride, fstride * p, in_stride,
                     fac + 1, m, pl, isign);
    }
    switch (p) {
    case 2:  bfly2(out, fstride, pl, m, isign); break;
    case 3:  bfly3(out, fstride, pl, m, isign); break;
    case 4:  bfly4(out, fstride, pl, m, isign); break;
    case 5:  bfly5(out, fstride, pl, m, isign); break;
    default: bfly_generic(out, fstride, pl, m, p, isign); break;
    }
}

/* complex 1-D transform, out-of-place, input with stride, output contiguous. isign +1: exp(-i), -1: exp(+i) */
void co_fft_c2c(const co_fftplan *pl, const co_cplx *in, int in_stride, co_cplx *out, int isign)
{
    if (pl->n == 1) { out[0] = in[0]; return; }
    fft_work(out, in, 1, in_stride, pl->fac, pl->n, pl, isign);
}

/* ---------------------------------------------------------------- 2-D real transforms -------------- */

/* forward: real a(n1,n2) (n1 fastest) -> complex A(n1/2+1, n2), unscaled. n1 must be even or 1.
 * Row transform via a length-n1/2 complex FFT of the packed even/odd samples + split step. */
void co_fft2_r2c(co_plancache *pc, int n1, int n2, const double *a, co_cplx *A)
{
    const int nh = n1 / 2, ldA = nh + 1;
    if (n1 % 2 != 0) {    /* odd first dimension: plain complex transform of each row */
        co_fftplan *p1 = co_fft_cached(pc, n1);
        co_cplx *row = (co_cplx *) malloc(sizeof(co_cplx) * 2 * (size_t) n1);
        for (int j = 0; j < n2; j++) {
            for (int i = 0; i < n1; i++) { row[i].re = a[(size_t) j * n1 + i]; row[i].im = 0.0; }
            co_fft_c2c(p1, row, 1, row + n1, +1);
            for (int k = 0; k < ldA; k++) A[(size_t) j * ldA + k] = row[n1 + k];
        }
        free(row);
    } else {
        co_fftplan *ph = co_fft_cached(pc, nh);
        co_fftplan *pf = co_fft_cached(pc, n1);     /* for the w_n1^k table */
        co_cplx *z = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) (nh + 1));
        for (int j = 0; j < n2; j++) {
            const co_cplx *packed = (const co_cplx *) (a + (size_t) j * n1);   /* (even, odd) pairs */
            co_fft_c2c(ph, packed, 1, z, +1);
            z[nh] = z[0];
            co_cplx *Aj = A + (size_t) j * ldA;
            for (int k = 0; k <= nh; k++) {
                co_cplx zk = z[k], zc = { z[nh - k].re, -z[nh - k].im };
                co_cplx e = { 0.5 * (zk.re + zc.re), 0.5 * (zk.im + zc.im) };       /* even part  */
                co_cplx o = { 0.5 * (zk.im - zc.im), -0.5 * (zk.re - zc.re) };      /* odd part = (zk-zc)/(2i) */
                co_cplx w = (k < n1) ? pf->tw[k] : pf->tw[0], t;
                CMUL(t, o, w);
                Aj[k].re = e.re + t.re;  Aj[k].im = e.im + t.im;
            }
        }
        free(z);
    }
    /* column transforms */
    if (n2 > 1) {
        co_fftplan *p2 = co_fft_cached(pc, n2);
        co_cplx *col = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) n2);
        for (int k = 0; k < ldA; k++) {
            co_fft_c2c(p2, A + k, ldA, col, +1);
            for (int j = 0; j < n2; j++) A[(size_t) j * ldA + k] = col[j];
        }
        free(col);
    }
}

/* backward: complex A(n1/2+1, n2) -> real a(n1,n2), multiplied by scale. A is overwritten. */
void co_fft2_c2r(co_plancache *pc, int n1, int n2, co_cplx *A, double *a, double scale)
{
    const int nh = n1 / 2, ldA = nh + 1;
    if (n2 > 1) {
        co_fftplan *p2 = co_fft_cached(pc, n2);
        co_cplx *col = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) n2);
        for (int k = 0; k < ldA; k++) {
            co_fft_c2c(p2, A + k, ldA, col, -1);
            for (int j = 0; j < n2; j++) A[(size_t) j * ldA + k] = col[j];
        }
        free(col);
    }
    if (n1 % 2 != 0) {
        co_fftplan *p1 = co_fft_cached(pc, n1);
        co_cplx *row = (co_cplx *) malloc(sizeof(co_cplx) * 2 * (size_t) n1);
        for (int j = 0; j < n2; j++) {
            const co_cplx *Aj = A + (size_t) j * ldA;
            for (int k = 0; k < ldA; k++) row[k] = Aj[k];
            for (int k = ldA; k < n1; k++) { row[k].re = Aj[n1 - k].re; row[k].im = -Aj[n1 - k].im; }
            co_fft_c2c(p1, row, 1, row + n1, -1);
            for (int i = 0; i < n1; i++) a[(size_t) j * n1 + i] = scale * row[n1 + i].re;
        }
        free(row);
    } else {
        co_fftplan *ph = co_fft_cached(pc, nh);
        co_fftplan *pf = co_fft_cached(pc, n1);
        co_cplx *z = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) (nh + 1));
        for (int j = 0; j < n2; j++) {
            const co_cplx *Aj = A + (size_t) j * ldA;
            for (int k = 0; k < nh; k++) {
                co_cplx xk = Aj[k], xc = { Aj[nh - k].re, -Aj[nh - k].im };
                co_cplx e = { xk.re + xc.re, xk.im + xc.im };
                co_cplx d = { xk.re - xc.re, xk.im - xc.im };
                co_cplx w = { pf->tw[k].re, -pf->tw[k].im }, o;      /* exp(+2 pi i k/n1) */
                CMUL(o, d, w);
                /* z[k] = e + i*o = 2 Z[k]; the unnormalised half-length inverse then yields n1 * x */
                z[k].re = e.re - o.im;  z[k].im = e.im + o.re;
            }
            co_cplx *outp = (co_cplx *) (a + (size_t) j * n1);
            co_fft_c2c(ph, z, 1, outp, -1);
            for (int i = 0; i < n1; i++) a[(size_t) j * n1 + i] *= scale;
        }
        free(z);
    }
}
