/*
 * co_aijpj.c -- ORACLE (test infrastructure, not product code).
 * The influence product u = A p: direct row sum (AijPj), FFT convolution (VecAijPj), FFT approximate
 * inverse (fft_makePrec), FFT size chooser.  Follows /root/reference/src/m_aijpj.f90 and
 * /root/reference/src/m_gridfunc.f90:582-637 (areas).
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <malloc.h>

co_ctx *co_ctx_new(void)
{
    /* keep the per-product work arrays (0.3 - 20 MB) on the heap: with glibc's default mmap threshold every
     * allocate/deallocate pair of fft_VecAijPj would be an mmap/munmap, which serialises threads on the mm lock */
    static int once = 0;
    if (!once) { mallopt(M_MMAP_THRESHOLD, 1 << 30); mallopt(M_TRIM_THRESHOLD, 1 << 30); once = 1; }
    co_ctx *cx = (co_ctx *) calloc(1, sizeof(co_ctx));
    return cx;
}

void co_ctx_free(co_ctx *cx)
{
    if (!cx) return;
    co_plancache_clear(&cx->pc);
    free(cx);
}

void co_eldiv_init(co_eldiv *e, int mx, int my)
{
    e->mx = mx; e->my = my;
    e->el = (int *) calloc((size_t) mx * my, sizeof(int));
    e->row1st = (int *) calloc((size_t) my, sizeof(int));
    e->rowlst = (int *) calloc((size_t) my, sizeof(int));
    for (int iy = 0; iy < my; iy++) { e->row1st[iy] = mx; e->rowlst[iy] = 0; }
    e->ixmin = mx; e->ixmax = 1; e->iymin = my; e->iymax = 1;
}

void co_eldiv_free(co_eldiv *e)
{
    free(e->el); free(e->row1st); free(e->rowlst);
    e->el = e->row1st = e->rowlst = NULL;
}

/* m_gridfunc.f90:582-637 */
void co_areas(co_eldiv *e)
{
    const int nx = e->mx, ny = e->my;
    e->ixmin = nx; e->ixmax = 1; e->iymin = ny; e->iymax = 1;
    for (int iy = 1; iy <= ny; iy++) {
        const int *row = e->el + (long) (iy - 1) * nx - 1;      /* 1-based ix */
        int first = 1;
        while (row[first] <= CO_EXTER && first != nx) first++;
        int last = 0;
        for (int ix = first; ix <= nx; ix++) if (row[ix] >= CO_ADHES) last = ix;
        e->row1st[iy - 1] = first;
        e->rowlst[iy - 1] = last;
        if (first <= last) {
            if (first < e->ixmin) e->ixmin = first;
            if (last > e->ixmax) e->ixmax = last;
            if (iy < e->iymin) e->iymin = iy;
            if (iy > e->iymax) e->iymax = iy;
        }
    }
}

/* m_aijpj.f90:67-95 */
static void igs_range(int iigs, int *igs0, int *igs1)
{
    if (iigs == CO_ALLELM)      { *igs0 = CO_EXTER; *igs1 = CO_PLAST; }
    else if (iigs == CO_ALLEXT) { *igs0 = CO_EXTER; *igs1 = CO_EXTER; }
    else if (iigs == CO_ALLINT) { *igs0 = CO_ADHES; *igs1 = CO_PLAST; }
    else if (iigs >= CO_EXTER && iigs <= CO_PLAST) { *igs0 = iigs; *igs1 = iigs; }
    else { *igs0 = CO_ADHES; *igs1 = CO_EXTER; }
}

/* m_aijpj.f90:1022-1119 */
int co_opt_fft_size(int fft_size)
{
    static const int kpattern[9][8] = {
        { -2,  1,  0,  0,  0,  4,  3,  3 },
        { -5,  3,  0,  0,  0, 32, 27,  3 },
        { -1, -1,  1,  0,  0,  6,  5,  5 },
        { -4,  1,  1,  0,  0, 16, 15,  5 },
        {  0, -3,  2,  0,  0, 27, 25,  5 },
        { -3,  0,  0,  1,  0,  8,  7,  7 },
        {  1, -1, -1,  1,  0, 15, 14,  7 },
        { -1,  2, -1,  0,  0, 10,  9,  5 },
        { -2, -1,  0,  0,  1, 12, 11, 11 } };
    const int use_fac = 7;
    int kfac[5] = { 0, 0, 0, 0, 0 };
    int k2 = (int) ceil(log(1.0 * fft_size) / log(2.0));
    long prod = 1L << k2;
    kfac[0] = k2;
    int changed = 1;
    while (changed) {
        changed = 0;
        for (int ip = 0; ip < 9; ip++) {
            const int *kp = kpattern[ip];
            if (kp[5] > 0 && kp[7] <= use_fac) {
                for (;;) {
                    int ok = 1;
                    for (int f = 0; f < 5; f++) if (kfac[f] + kp[f] < 0) ok = 0;
                    if (!(ok && (long) kp[6] * prod >= (long) kp[5] * fft_size)) break;
                    for (int f = 0; f < 5; f++) kfac[f] += kp[f];
                    prod = kp[6] * prod / kp[5];
                    changed = 1;
                }
            }
        }
    }
    return (int) prod;
}

/* m_aijpj.f90:99-254 (itypcf=0; no third-body layer / compressible sheet in the elastic half-space scope) */
double co_aijpj(int ii, int ik, const double *p, const co_eldiv *pel, int jkarg, const co_inflcf *c)
{
    int jk0, jk1;
    if (jkarg == -3) {
        jk0 = 1; jk1 = 3;
        if (!c->nt_cpl && ik <= 2) jk1 = 2;
        if (!c->nt_cpl && ik == 3) jk0 = 3;
    } else if (jkarg == -2) {
        jk0 = 1; jk1 = 2;
        if (!c->nt_cpl && ik == 3) jk1 = 0;
    } else if (jkarg >= 1 && jkarg <= 3) {
        jk0 = jkarg; jk1 = jkarg;
        if (!c->nt_cpl && ik <= 2) jk1 = (jk1 < 2 ? jk1 : 2);
        if (!c->nt_cpl && ik == 3) jk0 = 3;
    } else { jk0 = 1; jk1 = 0; }

    const int mx = pel->mx, my = pel->my, npot = mx * my;
    const int ix = (ii - 1) % mx + 1, iy = (ii - 1) / mx + 1;
    double partsum = 0.0;
    for (int jy = 1; jy <= my; jy++) {
        int j0 = pel->row1st[jy - 1] - 1; if (j0 < 1) j0 = 1;
        int j1 = pel->rowlst[jy - 1] + 1; if (j1 > mx) j1 = mx;
        for (int jk = jk0; jk <= jk1; jk++) {
            const double *blk = co_cf_ptr(c, ik, jk);
            const double *pj = p + (long) (jk - 1) * npot + (long) (jy - 1) * mx - 1;
            double rowsum = 0.0;
            for (int jx = j0; jx <= j1; jx++)
                rowsum = rowsum + CO_CF(c, blk, ix - jx, iy - jy) * pj[jx];
            partsum = partsum + rowsum;
        }
    }
    double res = partsum * c->ga_inv;
    if (c->use_flxz && ik == CO_Z && (jkarg == ik || jkarg == CO_ALL))
        res += c->flx_z * p[(long) (ik - 1) * npot + ii - 1];
    if (c->use_3bl && ik <= CO_Y && (jkarg == ik || jkarg <= CO_TANG))
        res += c->flx_3bl * p[(long) (ik - 1) * npot + ii - 1];
    return res;
}

static void ensure_fftcf(co_inflcf *c, int fft_mx, int fft_my)
{   /* m_aijpj.f90:873-884 */
    long len = (long) (fft_mx + 1) * 2 * fft_my;
    if (c->fft_mx != fft_mx || c->fft_my != fft_my || c->fft_len != len) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c->fft_ok[i][j] = 0;
        c->fft_mx = fft_mx; c->fft_my = fft_my; c->fft_len = len;
    }
}

static void transform_cf(co_ctx *cx, co_inflcf *c, int ik, int jk, int fft_mx, int fft_my, int lim_mx, int lim_my)
{   /* m_aijpj.f90:889-920 / 614-650: copy block to work array with cf(0,0) at 0-based (fft_mx, fft_my) */
    const long n1 = 2L * fft_mx, n2 = 2L * fft_my;
    double *csw = (double *) calloc((size_t) (n1 * n2), sizeof(double));
    const double *blk = co_cf_ptr(c, ik, jk);
    const int ly = fft_my < lim_my ? fft_my : lim_my, lx = fft_mx < lim_mx ? fft_mx : lim_mx;
    for (int iy = -ly; iy <= ly - 1; iy++)
        for (int ix = -lx; ix <= lx - 1; ix++)
            csw[(long) (iy + fft_my) * n1 + fft_mx + ix] = CO_CF(c, blk, ix, iy);
    free(c->fft_cf[ik - 1][jk - 1]);
    c->fft_cf[ik - 1][jk - 1] = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) c->fft_len);
    co_fft2_r2c(&cx->pc, (int) n1, (int) n2, csw, c->fft_cf[ik - 1][jk - 1]);
    free(csw);
    c->fft_ok[ik - 1][jk - 1] = 1;
    c->n_cfft++;
}

/* m_aijpj.f90:712-1015 */
static void fft_vecaijpj(co_ctx *cx, const co_eldiv *igs, int ladd, int iigs, double *u, int ik, const double *p,
                         const co_eldiv *pel, int jk, co_inflcf *c)
{
    const int mx = igs->mx, my = igs->my, npot = mx * my;
    int ix0, ix1, iy0, iy1, iarea = 0, igs0, igs1;
    double *uk = u + (long) (ik - 1) * npot;
    const double *pk = p + (long) (jk - 1) * npot;
    igs_range(iigs, &igs0, &igs1);

    ix0 = 1; ix1 = mx; iy0 = 1; iy1 = my;
    if (iigs == CO_ALLINT) {                                              /* :774-782 */
        int a = igs->ixmin - 1, b = pel->ixmin - 1;
        ix0 = a < b ? a : b; if (ix0 < 1) ix0 = 1;
        a = igs->ixmax + 1; b = pel->ixmax + 1;
        ix1 = a > b ? a : b; if (ix1 > mx) ix1 = mx;
        iy0 = igs->iymin < pel->iymin ? igs->iymin : pel->iymin;
        iy1 = igs->iymax > pel->iymax ? igs->iymax : pel->iymax;
        iarea = (ix1 - ix0 + 1) * (iy1 - iy0 + 1);
    }
    if (iigs != CO_ALLINT || 1.1 * iarea > (double) mx * my || cx->fullbox) {   /* :788-793 */
        ix0 = 1; ix1 = mx; iy0 = 1; iy1 = my;
    }
    if (ix1 < ix0 || iy1 < iy0) {                                         /* :795-801 */
        if (!ladd)
            for (int ii = 0; ii < npot; ii++)
                if (iigs == CO_ALLELM || igs->el[ii] >= CO_ADHES) uk[ii] = 0.0;
        return;
    }
    const int fft_mx = co_opt_fft_size(ix1 - ix0 + 1), fft_my = co_opt_fft_size(iy1 - iy0 + 1);
    const long n1 = 2L * fft_mx, n2 = 2L * fft_my, ld = fft_mx + 1;

    ensure_fftcf(c, fft_mx, fft_my);
    if (!c->fft_ok[ik - 1][jk - 1]) transform_cf(cx, c, ik, jk, fft_mx, fft_my, mx, my);

    double  *ps = (double *) calloc((size_t) (n1 * n2), sizeof(double));
    double  *us = (double *) malloc(sizeof(double) * (size_t) (n1 * n2));
    co_cplx *pf = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) c->fft_len);

    for (int iy = iy0; iy <= iy1; iy++)                                   /* :932-939 */
        for (int ix = ix0; ix <= ix1; ix++)
            ps[(long) (iy - iy0) * n1 + (ix - ix0)] = pk[(long) (iy - 1) * mx + ix - 1];

    co_fft2_r2c(&cx->pc, (int) n1, (int) n2, ps, pf);
    const co_cplx *cf = c->fft_cf[ik - 1][jk - 1];
    for (long i = 0; i < c->fft_len; i++) {                               /* :956-958 */
        co_cplx a = cf[i], b = pf[i];
        pf[i].re = a.re * b.re - a.im * b.im;
        pf[i].im = a.re * b.im + a.im * b.re;
    }
    co_fft2_c2r(&cx->pc, (int) n1, (int) n2, pf, us, 1.0 / (4.0 * fft_mx * fft_my));
    (void) ld;

    for (int iy = iy0; iy <= iy1; iy++)                                   /* :978-1005 */
        for (int ix = ix0; ix <= ix1; ix++) {
            const long ii = (long) (iy - 1) * mx + ix - 1;
            const double v = us[(long) (fft_my + iy - iy0) * n1 + fft_mx + (ix - ix0)];
            if (igs->el[ii] >= igs0 && igs->el[ii] <= igs1) {
                if (ladd) uk[ii] = uk[ii] + v; else uk[ii] = v;
            }
        }
    free(ps); free(us); free(pf);

    /* work accounting, SURVEY 8(d) */
    cx->st.n_prod++;
    {
        const double S = (double) (fft_mx + 1) * 2.0 * fft_my, N = 4.0 * fft_mx * fft_my;
        cx->st.alg_bytes += 8.0 * npot * 2.0 + npot;
        cx->st.alg_flops += 2.0 * 2.5 * N * log2(N) + 6.0 * S;
    }
}

/* m_aijpj.f90:258-451 with usefft = .true. (WITH_MKLFFT, itypcf=0) */
void co_vecaijpj(co_ctx *cx, const co_eldiv *igs, int iigs, double *u, int ikarg, const double *p,
                 const co_eldiv *pel, int jkarg, co_inflcf *c)
{
    const int npot = igs->mx * igs->my;
    int ik0, ik1, jk0, jk1, igs0, igs1;
    igs_range(iigs, &igs0, &igs1);
    if (ikarg == -3) { ik0 = 1; ik1 = 3; } else if (ikarg == -2) { ik0 = 1; ik1 = 2; }
    else if (ikarg >= 1 && ikarg <= 3) { ik0 = ik1 = ikarg; } else { ik0 = 1; ik1 = 0; }
    if (jkarg == -3) { jk0 = 1; jk1 = 3; } else if (jkarg == -2) { jk0 = 1; jk1 = 2; }
    else if (jkarg >= 1 && jkarg <= 3) { jk0 = jk1 = jkarg; } else { jk0 = 1; jk1 = 0; }

    for (int ik = ik0; ik <= ik1; ik++) {
        int ladd = 0;
        double *uk = u + (long) (ik - 1) * npot;
        for (int jk = jk0; jk <= jk1; jk++) {
            if (!c->nt_cpl && (ik * jk == 3 || ik * jk == 6)) {           /* :358-369 */
                if (!ladd)
                    for (int ii = 0; ii < npot; ii++)
                        if (iigs == CO_ALLELM || igs->el[ii] >= CO_ADHES) uk[ii] = 0.0;
            } else {
                fft_vecaijpj(cx, igs, ladd, iigs, u, ik, p, pel, jk, c);
                ladd = 1;
            }
        }
    }
    for (int ik = ik0; ik <= ik1; ik++) {                                 /* :395 gf3_scal(iigs, ga_inv) */
        double *uk = u + (long) (ik - 1) * npot;
        for (int ii = 0; ii < npot; ii++)
            if (iigs == CO_ALLELM || igs->el[ii] >= CO_ADHES) uk[ii] = c->ga_inv * uk[ii];
    }
    for (int ik = ik0; ik <= ik1; ik++) {                                 /* :400-422 */
        double *uk = u + (long) (ik - 1) * npot;
        const double *pk = p + (long) (ik - 1) * npot;
        if (c->use_flxz && ik == CO_Z && (jkarg == ik || jkarg == CO_ALL))
            for (int ii = 0; ii < npot; ii++)
                if (igs->el[ii] >= igs0 && igs->el[ii] <= igs1) uk[ii] += c->flx_z * pk[ii];
        if (c->use_3bl && ik <= CO_Y && (jkarg == ik || jkarg <= CO_TANG))
            for (int ii = 0; ii < npot; ii++)
                if (igs->el[ii] >= igs0 && igs->el[ii] <= igs1) uk[ii] += c->flx_3bl * pk[ii];
    }
}

/* m_aijpj.f90:428-446: the non-FFT implementation of VecAijPj -- the independent high-precision check */
void co_vecaijpj_direct(const co_eldiv *igs, int iigs, double *u, int ikarg, const double *p,
                        const co_eldiv *pel, int jkarg, const co_inflcf *c)
{
    const int npot = igs->mx * igs->my;
    int ik0, ik1, igs0, igs1;
    igs_range(iigs, &igs0, &igs1);
    if (ikarg == -3) { ik0 = 1; ik1 = 3; } else if (ikarg == -2) { ik0 = 1; ik1 = 2; }
    else if (ikarg >= 1 && ikarg <= 3) { ik0 = ik1 = ikarg; } else { ik0 = 1; ik1 = 0; }
    for (int ii = 1; ii <= npot; ii++)
        if (igs->el[ii - 1] >= igs0 && igs->el[ii - 1] <= igs1)
            for (int ik = ik0; ik <= ik1; ik++)
                u[(long) (ik - 1) * npot + ii - 1] = co_aijpj(ii, ik, p, pel, jkarg, c);
}

/* m_aijpj.f90:457-708 */
void co_fft_makeprec(co_ctx *cx, int ik, co_inflcf *c, int jk, co_inflcf *m)
{
    m->nt_cpl = c->nt_cpl;
    const int fft_mx = c->cf_mx, fft_my = c->cf_my;                        /* :504-518, un-optimised size */
    const long n1 = 2L * fft_mx, n2 = 2L * fft_my;

    ensure_fftcf(c, fft_mx, fft_my);
    if (!c->fft_ok[ik - 1][jk - 1]) transform_cf(cx, c, ik, jk, fft_mx, fft_my, c->cf_mx, c->cf_my);

    co_cplx *mf = (co_cplx *) malloc(sizeof(co_cplx) * (size_t) c->fft_len);
    double  *ms = (double *) malloc(sizeof(double) * (size_t) (n1 * n2));
    const co_cplx *cf = c->fft_cf[ik - 1][jk - 1];
    for (long i = 0; i < c->fft_len; i++) {                               /* :673-675  mf = 1/cf */
        const double d = cf[i].re * cf[i].re + cf[i].im * cf[i].im;
        mf[i].re = cf[i].re / d;
        mf[i].im = -cf[i].im / d;
    }
    co_fft2_c2r(&cx->pc, (int) n1, (int) n2, mf, ms, 1.0 / (4.0 * fft_mx * fft_my));
    double *blk = co_cf_ptr(m, ik, jk);
    const int ly = fft_my < c->cf_my ? fft_my : c->cf_my, lx = fft_mx < c->cf_mx ? fft_mx : c->cf_mx;
    for (int iy = -ly; iy <= ly - 1; iy++)                                /* :697-702 */
        for (int ix = -lx; ix <= lx - 1; ix++)
            CO_CF(m, blk, ix, iy) = c->ga * c->ga * ms[(long) (iy + fft_my) * n1 + fft_mx + ix];
    free(mf); free(ms);
}
