/*
 * co_norm.c -- ORACLE (test infrastructure, not product code).
 * Normal contact problem: NormCG (bound-constrained preconditioned CG) and the NORM active-set wrapper.
 * Follows /root/reference/src/m_solvpn.f90:24-527 (normcg, normcg_matvec; iplan /= 4 only) and
 * /root/reference/src/m_snorm.f90:31-378 (snorm; no sensitivities), with the masked BLAS-1 of
 * /root/reference/src/m_gridfunc.f90:936-1591 written out in the reference's sequential order.
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---- masked BLAS-1 on one column, gf3_* with AllInt = el >= Adhes (m_gridfunc.f90:1245-1591) ---- */
static double dot_int(const co_eldiv *g, int n, const double *a, const double *b)
{ double s = 0.0; for (int i = 0; i < n; i++) if (g->el[i] >= CO_ADHES) s = s + a[i] * b[i]; return s; }

static double sum_all(int n, const double *a)
{ double s = 0.0; for (int i = 0; i < n; i++) s = s + a[i]; return s; }

static double sum_int(const co_eldiv *g, int n, const double *a)
{ double s = 0.0; for (int i = 0; i < n; i++) if (g->el[i] >= CO_ADHES) s = s + a[i]; return s; }

static double rms_int(const co_eldiv *g, int n, const double *a)
{
    double s = 0.0; int cnt = 0;
    for (int i = 0; i < n; i++) if (g->el[i] >= CO_ADHES) { s = s + a[i] * a[i]; cnt++; }
    return sqrt(s / (cnt > 1 ? cnt : 1));
}

static void proj_avg_int(const co_eldiv *g, int n, double *a)
{
    double avg = 0.0; int nval = 0;
    for (int i = 0; i < n; i++) if (g->el[i] >= CO_ADHES) { avg = avg + a[i]; nval++; }
    avg = avg / (double) (float) (nval > 1 ? nval : 1);          /* real(max(1,nval)) is REAL(4) */
    for (int i = 0; i < n; i++) if (g->el[i] >= CO_ADHES) a[i] = a[i] - avg;
}

static double min_all(int n, const double *a)
{ double m = 1e20; for (int i = 0; i < n; i++) if (a[i] < m) m = a[i]; return m; }

static double max_all(int n, const double *a)
{ double m = -1e20; for (int i = 0; i < n; i++) if (a[i] > m) m = a[i]; return m; }

/* m_solvpn.f90:24-461. Grid functions here are single columns (the normal direction) of length npot.
 * Returns 0, or 1 when the reference would abort_run (MaxCG reached while diverging, :423-429). */
int co_normcg(co_ctx *cx, int ic_norm, int npot, double dxdy, int use_fftprec, int maxcg, double eps,
              co_inflcf *cs, co_inflcf *ms, const double *hstot, double *pen, double fntrue,
              co_eldiv *igs, double *ps, int *itcg_out, double *err)
{
    /* VecAijPj takes 3-column gf3's; wrap the normal column as column 3 of a virtual array */
#define COL3(a) ((a) - 2L * npot)
    double *rhs = (double *) calloc(npot, sizeof(double)), *res = (double *) calloc(npot, sizeof(double));
    double *r_prv = (double *) calloc(npot, sizeof(double)), *dd = (double *) calloc(npot, sizeof(double));
    double *z = (double *) calloc(npot, sizeof(double)), *v = (double *) calloc(npot, sizeof(double));
    double *q = (double *) calloc(npot, sizeof(double));
    int numinn, ncon, itcg = 0, itinn, itchg, lchanged, ret = 0;
    double alpha, beta, rms_upd, rms_upd1 = 0.0, rms_xk, conv, pn, rz1, rz2, rrprv, rv, vav, davg, fk,
           hsmin, hsmax, htrsh;

    if (npot <= 150) numinn = 3; else if (npot <= 400) numinn = 2; else numinn = 1;      /* :54-60 */

    if (ic_norm == 1) *pen = 0.0;                                                        /* :103-106 */
    for (int i = 0; i < npot; i++) rhs[i] = *pen;
    for (int i = 0; i < npot; i++) rhs[i] = rhs[i] + (-1.0) * hstot[i];
    davg = 0.0;

    ncon = 0;
    for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) ncon++;

    if (ic_norm == 0) {                                                                  /* :117-136 */
        hsmin = min_all(npot, hstot) - *pen;
        if (hsmin >= 0.0) {
            for (int i = 0; i < npot; i++) { ps[i] = 0.0; igs->el[i] = CO_EXTER; }
            itcg = 0; *err = 0.0;
            goto done;
        }
    } else {                                                                             /* :138-169 */
        if (ncon <= 0) {
            hsmin = min_all(npot, hstot);
            hsmax = max_all(npot, hstot);
            htrsh = hsmin + 0.1 * fmax(hsmax - hsmin, 1e-10);
            for (int i = 0; i < npot; i++) if (hstot[i] < htrsh) { igs->el[i] = CO_ADHES; ncon++; }
            co_areas(igs);
        }
        fk = dxdy * sum_all(npot, ps);
        if (fabs(fk) < (double) 1e-3f * fntrue) {                 /* 1e-3 is a REAL(4) literal :160 */
            pn = fntrue / (dxdy * (double) (float) ncon);
            for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) ps[i] = pn;
            fk = dxdy * sum_all(npot, ps);
        } else {
            const double f = fntrue / fk;
            for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) ps[i] = f * ps[i];
        }
    }

    /* res = rhs - A ps on interior :173-175 */
    co_vecaijpj(cx, igs, CO_ALLINT, COL3(res), CO_Z, COL3(ps), igs, CO_Z, cs);
    for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) res[i] = -1.0 * res[i];
    for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) res[i] = res[i] + 1.0 * rhs[i];
    if (ic_norm == 1) proj_avg_int(igs, npot, res);
    rz2 = 0.0;

    itcg = 0; itchg = 0; itinn = 0;
    rms_xk = 1.0;
    rms_upd = 2.0 * eps * rms_xk;
    lchanged = 0;

    while ((lchanged || rms_upd > eps * rms_xk) && itcg < maxcg) {                       /* :194 */
        itcg++; itinn++;

        if (use_fftprec)                                                                 /* :203-207 */
            co_vecaijpj(cx, igs, CO_ALLINT, COL3(z), CO_Z, COL3(res), igs, CO_Z, ms);
        else
            memcpy(z, res, sizeof(double) * npot);
        if (ic_norm == 1) proj_avg_int(igs, npot, z);

        rz1 = rz2;
        rz2 = dot_int(igs, npot, z, res);

        if (itcg <= 1 || rz1 < CO_TINY) {                                                /* :228-241 */
            for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) v[i] = z[i];
        } else {
            rrprv = dot_int(igs, npot, z, r_prv);
            beta = fmax(0.0, (rz2 - rrprv) / fmax(CO_TINY, rz1));
            for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) v[i] = beta * v[i];
            for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) v[i] = v[i] + 1.0 * z[i];
        }
        if (ic_norm == 1) proj_avg_int(igs, npot, v);

        co_vecaijpj(cx, igs, CO_ALLINT, COL3(q), CO_Z, COL3(v), igs, CO_Z, cs);          /* :249 */
        if (ic_norm == 1) proj_avg_int(igs, npot, q);

        rv  = dot_int(igs, npot, res, v);
        vav = dot_int(igs, npot, q, v);
        if (fabs(vav) > 1e-32 && ncon == 1) alpha = rv / vav;
        else alpha = rv / fmax(CO_TINY, vav);

        for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) ps[i] = ps[i] + alpha * v[i];

        rms_upd = fabs(alpha) * rms_int(igs, npot, v);
        if (itcg == 1) rms_upd1 = rms_upd;
        if (itcg <= 3 || itcg % 10 == 0) rms_xk = rms_int(igs, npot, ps);

        memcpy(r_prv, res, sizeof(double) * npot);                                       /* :294 */

        if (itinn < numinn && rms_upd >= eps * rms_xk) {                                 /* :298-303 */
            for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) res[i] = res[i] + (-alpha) * q[i];
        } else {
            int lchg_negpn = 0, lchg_intpen = 0;
            for (int i = 0; i < npot; i++)                                               /* :310-318 */
                if (igs->el[i] >= CO_ADHES && ps[i] < 0.0) {
                    igs->el[i] = CO_EXTER; ncon--; ps[i] = 0.0; lchg_negpn = 1;
                }
            if (ncon <= 0) {                                                             /* :323-333 */
                hsmin = min_all(npot, hstot);
                for (int i = 0; i < npot; i++)
                    if (hstot[i] <= hsmin + 1e-5) { igs->el[i] = CO_ADHES; ncon++; ps[i] = 0.0; lchg_negpn = 1; }
            }
            if (ic_norm == 1 && lchg_negpn) {                                            /* :337-344 */
                fk = dxdy * sum_all(npot, ps);
                if (fabs(fk) < (double) 1e-3f * fntrue) {
                    for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) ps[i] = 1.0;
                    fk = sum_all(npot, ps);
                }
                const double f = fntrue / fk;
                for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) ps[i] = f * ps[i];
            }
            if (lchg_negpn) co_areas(igs);

            co_vecaijpj(cx, igs, CO_ALLELM, COL3(dd), CO_Z, COL3(ps), igs, CO_Z, cs);    /* :351-352 */
            for (int i = 0; i < npot; i++) dd[i] = dd[i] + (-1.0) * rhs[i];

            if (ic_norm == 1) {                                                          /* :356-359 */
                davg = sum_int(igs, npot, dd) / (double) (float) ncon;
                proj_avg_int(igs, npot, dd);
            }
            for (int i = 0; i < npot; i++) res[i] = 0.0;
            for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) res[i] = res[i] + (-1.0) * dd[i];

            for (int i = 0; i < npot; i++)                                               /* :367-380 */
                if (igs->el[i] == CO_EXTER && dd[i] - davg < 0.0) {
                    igs->el[i] = CO_ADHES; ncon++;
                    res[i] = -(dd[i] - davg);
                    lchg_intpen = 1;
                }
            for (int i = 0; i < npot; i++) if (igs->el[i] == CO_EXTER) v[i] = 0.0;
            if (lchg_intpen) co_areas(igs);

            itinn = 0;
            lchanged = lchg_intpen || lchg_negpn;
            if (lchanged) itchg++;
        }
    }

    if (ic_norm == 1) *pen = davg;                                                       /* :414 */
    conv = 1.0;
    if (rms_upd * rms_upd1 > 0.0 && itcg > 1) conv = exp(log(rms_upd / rms_upd1) / (itcg - 1));
    if (rms_upd > rms_xk && conv > 1.0 && itcg >= maxcg) ret = 1;                        /* abort_run in ref. */
    *err = rms_upd;

done:
    *itcg_out = itcg;
    free(rhs); free(res); free(r_prv); free(dd); free(z); free(v); free(q);
    return ret;
#undef COL3
}

/* m_snorm.f90:31-378 (WITH_MKLFFT; elastic material so use_fftprec = true; iplan /= 4; no sensitivities) */
void co_snorm(co_ctx *cx, int ic_norm, int mx, int my, double dxdy, const co_solv *solv, const double *hs,
              co_inflcf *cs, co_inflcf *ms, double *pen, double *fntrue, co_eldiv *igs, double *ps,
              co_norm_info *info)
{
    const int npot = mx * my;
    double *psn = ps + 2L * npot;                    /* normal column of ps(npot,3) */
    const double *hsn = hs + 2L * npot;
    double *htang = (double *) calloc(3L * npot, sizeof(double));
    double *hstot = (double *) calloc(npot, sizeof(double));
    double *tmp = (double *) calloc(3L * npot, sizeof(double));
    double *unn = (double *) calloc(3L * npot, sizeof(double));
    int zready, ncon, newext, newcon, itcg = 0, it = 0, itnorm = 0;
    double errpn = 0.0, tol, dd;

    co_areas(igs);                                                                       /* :68 */
    const int use_vecfft = (npot >= 700);                                                /* :71 */
    const int use_fftprec = 1;
    info->diverged = 0;

    co_fft_makeprec(cx, 3, cs, 3, ms);                                                   /* :93-97 */

    memcpy(hstot, hsn, sizeof(double) * npot);                                           /* :112-119 */
    if (cs->nt_cpl) {
        co_vecaijpj(cx, igs, CO_ALLELM, htang, CO_Z, ps, igs, CO_TANG, cs);
        for (int i = 0; i < npot; i++) hstot[i] = hstot[i] + 1.0 * htang[2L * npot + i];
    }

    do {                                                                                 /* label 20 */
        itnorm++;
        zready = 1;
        ncon = 0;
        for (int i = 0; i < npot; i++) { if (igs->el[i] >= CO_ADHES) ncon++; else psn[i] = 0.0; }
        co_areas(igs);

        if (co_normcg(cx, ic_norm, npot, dxdy, use_fftprec, solv->maxgs, solv->eps, cs, ms, hstot, pen,
                      *fntrue, igs, psn, &it, &errpn)) info->diverged = 1;
        itcg += it;

        ncon = 0;
        for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) ncon++;

        newext = 0;                                                                      /* :197-219 */
        for (int i = 0; i < npot; i++)
            if (igs->el[i] >= CO_ADHES && psn[i] < -errpn) { igs->el[i] = CO_EXTER; newext++; psn[i] = 0.0; }
        ncon -= newext;
        if (newext > 0) zready = 0;

        if (zready) {                                                                    /* :227-292 */
            co_areas(igs);
            if (use_vecfft) co_vecaijpj(cx, igs, CO_ALLELM, unn, CO_Z, ps, igs, CO_Z, cs);

            int ix = mx / 2 > 1 ? mx / 2 : 1, iy = my / 2 > 1 ? my / 2 : 1;
            int ii = ix + (iy - 1) * mx;
            for (int i = 0; i < npot; i++) tmp[2L * npot + i] = 1.0;
            tol = fabs(errpn * co_aijpj(ii, CO_Z, tmp, igs, CO_Z, cs));
            cx->st.n_rowsum++;

            newcon = 0;
            for (int i = 0; i < npot; i++)
                if (igs->el[i] == CO_EXTER && hstot[i] - *pen < 0.0) {
                    if (use_vecfft) dd = hstot[i] - *pen + unn[2L * npot + i];
                    else { dd = hstot[i] - *pen + co_aijpj(i + 1, CO_Z, ps, igs, CO_Z, cs); cx->st.n_rowsum++; }
                    if (dd < -tol) { igs->el[i] = CO_ADHES; newcon++; }
                }
            ncon += newcon;
            if (newcon != 0) zready = 0;
        }
        if (it >= solv->maxgs) zready = 0;                                               /* :296 */
    } while (!zready && itnorm < solv->maxin);

    if (!zready) itnorm = -1;                                                            /* :305-311 */

    for (int i = 0; i < npot; i++)                                                       /* :315-325 */
        if (psn[i] < 0.0 && igs->el[i] >= CO_ADHES) { psn[i] = 0.0; igs->el[i] = CO_EXTER; ncon--; }

    co_areas(igs);
    if (ic_norm == 0) *fntrue = dxdy * sum_all(npot, psn);                               /* :352 */
    info->itcg = itcg;
    info->itnorm = itnorm;
    free(htang); free(hstot); free(tmp); free(unn);
}
