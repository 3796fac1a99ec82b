/*
 * co_gdsteady.c -- ORACLE (test infrastructure, not product code).
 * GDsteady: non-linear gradient-descent type solver for steady rolling (T=3, G=5) on the traction increments dp, with
 * the FFT product for A dv, a diagonal scaling and a Brent line search.  Follows /root/reference/src/gdsteady.f90:
 * gdsteady (:9-606), compute_dp (:610-635), project_searchdir (:639-803), apply_trcbnd (:807-847), solve_elmtrc
 * (:851-897), compute_diagscaling (:901-1042), perform_linesearch (:1046-1462); elastic material, chi = 0.
 *
 * Parity status: NO golden file of the reference exercises GDsteady on a module-3 grid (perfc_test/get_times.ref_out
 * predates the solver), so this restatement is pinned only by the property that it converges to the SteadyGS solution
 * of the same problem (tests/test_oracle_golden.py) -- "parity unpinned" for iteration counts.
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

static void compute_dp(int mx, int my, const double *ps, double *dp)
{
    const int n = mx * my;
    for (int iy = 0; iy < my; iy++) {
        int ii = mx - 1 + iy * mx;
        dp[ii] = ps[ii]; dp[n + ii] = ps[n + ii];
        for (int ix = mx - 2; ix >= 0; ix--) {
            ii = ix + iy * mx;
            dp[ii] = ps[ii] - ps[ii + 1]; dp[n + ii] = ps[n + ii] - ps[n + ii + 1];
        }
    }
}

/* :639-803.  dv in/out, v out; t = tangential direction per element; returns the iteration of the last fall-back */
static void project_searchdir(int mx, int my, const int *el, const double *g, int itgd, int *it_fb, double betaj, const double *t,
                              double *dv, double *v, const co_gdparams *sp, double fac_v)
{
    const int n = mx * my;
    int imeth, kdown;
    if (betaj > sp->betath || itgd - *it_fb <= 1) { imeth = sp->gd_meth; kdown = sp->kdown; }
    else { imeth = 2; kdown = sp->kdowfb; *it_fb = itgd; }
    double *dxin = (double *) malloc(sizeof(double) * mx), *dyin = (double *) malloc(sizeof(double) * mx);
    for (int iy = 0; iy < my; iy++) {
        const int i0 = iy * mx;
        if (imeth == 2) for (int ix = 0; ix < mx; ix++) { dxin[ix] = dv[i0 + ix]; dyin[ix] = dv[n + i0 + ix]; }
        int ii = i0 + mx - 1;
        v[ii] = 0.0; v[n + ii] = 0.0;
        for (int ix = mx - 2; ix >= 0; ix--) {
            ii = i0 + ix;
            if (el[ii] == CO_ADHES) {
                if (imeth == 1) { v[ii] = v[ii + 1] + dv[ii]; v[n + ii] = v[n + ii + 1] + dv[n + ii]; }
                else if (imeth == 2) {
                    if (ix + kdown <= mx - 1) { v[ii] = v[ii + 1] + dxin[ix] - dxin[ix + kdown]; v[n + ii] = v[n + ii + 1] + dyin[ix] - dyin[ix + kdown]; }
                    else { v[ii] = v[ii + 1] + dxin[ix]; v[n + ii] = v[n + ii + 1] + dyin[ix]; }
                    dv[ii] = v[ii] - v[ii + 1]; dv[n + ii] = v[n + ii] - v[n + ii + 1];
                } else {
                    v[ii] = sp->fdecay * v[ii + 1] + dv[ii]; v[n + ii] = sp->fdecay * v[n + ii + 1] + dv[n + ii];
                    dv[ii] = v[ii] - v[ii + 1]; dv[n + ii] = v[n + ii] - v[n + ii + 1];
                }
            } else if (el[ii] == CO_SLIP) {
                double vt = (imeth == 2) ? t[ii] * dxin[ix] + t[n + ii] * dyin[ix] : t[ii] * dv[ii] + t[n + ii] * dv[n + ii];
                vt = copysign(1.0, vt) * fmin(fabs(vt), fac_v * g[ii]);
                v[ii] = t[ii] * vt; v[n + ii] = t[n + ii] * vt;
                dv[ii] = v[ii] - v[ii + 1]; dv[n + ii] = v[n + ii] - v[n + ii + 1];
                if (imeth == 2) for (int k = 1; k <= kdown; k++) if (ix + k <= mx - 1) { dxin[ix + k] = 0.0; dyin[ix + k] = 0.0; }
            } else {
                v[ii] = 0.0; v[n + ii] = 0.0;
                dv[ii] = -v[ii + 1]; dv[n + ii] = -v[n + ii + 1];
                if (imeth == 2) for (int k = 1; k <= kdown; k++) if (ix + k <= mx - 1) { dxin[ix + k] = 0.0; dyin[ix + k] = 0.0; }
            }
        }
    }
    free(dxin); free(dyin);
}

/* :807-847: integrate the increments from the leading edge, clip at the traction bound */
static void apply_trcbnd(int mx, int my, const int *el, const double *g, double *dp, double *ps)
{
    const int n = mx * my;
    for (int iy = 0; iy < my; iy++) {
        int ii = mx - 1 + iy * mx;
        ps[ii] = 0.0; ps[n + ii] = 0.0;
        for (int ix = mx - 2; ix >= 0; ix--) {
            ii = ix + iy * mx;
            if (el[ii] <= CO_EXTER) { ps[ii] = 0.0; ps[n + ii] = 0.0; }
            else {
                ps[ii] = ps[ii + 1] + dp[ii]; ps[n + ii] = ps[n + ii + 1] + dp[n + ii];
                const double pa = sqrt(ps[ii] * ps[ii] + ps[n + ii] * ps[n + ii]);
                if (el[ii] == CO_SLIP || pa >= g[ii]) { ps[ii] = ps[ii] * g[ii] / pa; ps[n + ii] = ps[n + ii] * g[ii] / pa; }
            }
            dp[ii] = ps[ii] - ps[ii + 1]; dp[n + ii] = ps[n + ii] - ps[n + ii + 1];
        }
    }
}

static double wrap_pi(double e)
{
    if (fabs(e) >= CO_PI) e = e - nearbyint(e / (2.0 * CO_PI)) * 2.0 * CO_PI;
    return e;
}

/* :901-1042.  The reference's loop-carried `elnew` only matters for exterior elements (whose scaling is never used):
 * evaluated per element here. */
static void compute_diagscaling(int mx, int my, const int *el, const double *mus, const double coefs[2][2], const double *ps,
                                const double *ss, const co_gdparams *sp, double *dscl)
{
    const int n = mx * my;
    const double tiny_err = 1e-6, epselm = 1e-6;
    for (int ii = 0; ii < n; ii++) {
        int elnew = el[ii];
        double fac_s = 0.0;
        if (el[ii] == CO_SLIP) {
            double pr[3] = { ps[ii], ps[n + ii], ps[2L * n + ii] }, si[2] = { ss[ii], ss[n + ii] };
            double th_p0 = atan2(pr[1], pr[0]), th_s0 = atan2(-si[1], -si[0]);
            double err_0 = wrap_pi(th_s0 - th_p0);
            if (fabs(err_0) <= tiny_err) {
                pr[0] = mus[ii] * pr[2] * cos(th_p0 + 0.1);
                pr[1] = mus[ii] * pr[2] * sin(th_p0 + 0.1);
                si[0] = ss[ii] + coefs[0][0] * (pr[0] - ps[ii]) + coefs[0][1] * (pr[1] - ps[n + ii]);
                si[1] = ss[n + ii] + coefs[1][0] * (pr[0] - ps[ii]) + coefs[1][1] * (pr[1] - ps[n + ii]);
                th_p0 = atan2(pr[1], pr[0]); th_s0 = atan2(-si[1], -si[0]);
                err_0 = wrap_pi(th_s0 - th_p0);
            }
            if (fabs(err_0) > tiny_err) {
                co_plstrc(&elnew, coefs, epselm, 1.0, 1.0, pr, mus[ii], si);
                if (elnew == CO_SLIP) {
                    const double th_s1 = atan2(-si[1], -si[0]);
                    const double upd_s = wrap_pi(th_s1 - th_s0);
                    fac_s = fabs(upd_s) / fabs(err_0);
                }
            }
        }
        if (el[ii] == CO_SLIP && elnew == CO_SLIP) dscl[ii] = sp->d_slp * pow(fac_s, sp->pow_s);
        else if (el[ii] == CO_ADHES || (el[ii] == CO_SLIP && elnew == CO_ADHES)) {
            const int ix = ii % mx;
            int jx = ix;
            while (jx > 0 && el[ii - ix + jx] == CO_ADHES) jx--;
            double dnew;
            if (sp->d_cns < sp->d_ifc) dnew = fmax(sp->d_cns, sp->d_ifc + (ix - jx - 1) * fmin(0.0, sp->d_lin));
            else dnew = fmin(sp->d_cns, sp->d_ifc + (ix - jx - 1) * fmax(0.0, sp->d_lin));
            dscl[ii] = fmin(50.0 * dscl[ii], dnew);
        } else dscl[ii] = 0.0;
    }
}

static void set_nt(int n, const double *ps, double *nn, double *t)
{
    for (int i = 0; i < n; i++) {
        const double th = atan2(ps[n + i], ps[i]);
        nn[i] = cos(th); nn[n + i] = sin(th);
        t[i] = -nn[n + i]; t[n + i] = nn[i];
    }
}

static void set_residual(int n, const int *el, const double *t, const double *ss, const double *dscl, double *r)
{
    memset(r, 0, sizeof(double) * 2 * n);
    for (int i = 0; i < n; i++) {
        if (el[i] == CO_ADHES) { r[i] = -ss[i] * dscl[i]; r[n + i] = -ss[n + i] * dscl[i]; }
        else if (el[i] == CO_SLIP) {
            const double st = t[i] * ss[i] + t[n + i] * ss[n + i];
            r[i] = -st * t[i] * dscl[i]; r[n + i] = -st * t[n + i] * dscl[i];
        }
    }
}

/* :1046-1462: Brent line search on rho(alpha) = |D (s + alpha q)|^2 with the projected tractions */
static int perform_linesearch(int mx, int my, const int *el, const double *g, const double *dscl, const double *dp, const double *ss,
                              const double *dv, const double *v, const double *q, const double *told, double alpha0,
                              double *alpha_out, int *nobracket)
{
    enum { max_j = 99 };
    const int n = mx * my;
    const double tiny = 1e-12, eps_j = 0.001;
    double *dpnew = (double *) malloc(sizeof(double) * 2 * n), *psnew = (double *) calloc(2L * n, sizeof(double));
    double tbl_alpha[max_j + 2], tbl_rho[max_j + 2], tbl_drho[max_j + 2];
    int j = 0, stop_j = 0, has_bracket = 0, ita = 0, itb_ = 0, itx = 0, itw = 0, itv = 0, itu = 0, ibrack = 1;
    double alpha_j = 0.0, alpha_prv = -1.0, dst_cur = 0.0, dst_prv1 = 0.0, dst_prv2 = 0.0, ddrho = 10.0;
    (void) alpha_prv; (void) ddrho;
    while (!stop_j) {
        double rhonew = 0.0, c0adh = 0.0, c1adh = 0.0, c0slp = 0.0, c1slp = 0.0, c2slp = 0.0;
        memcpy(dpnew, dp, sizeof(double) * 2 * n);
        for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (el[i] >= CO_ADHES) dpnew[(long) k * n + i] += alpha_j * dv[(long) k * n + i];
        apply_trcbnd(mx, my, el, g, dpnew, psnew);
        for (int i = 0; i < n; i++) {
            const double vt = v[i] * told[i] + v[n + i] * told[n + i];
            const double dth_da = (g[i] * vt) / (fmax(tiny, g[i] * g[i]) + alpha_j * alpha_j * vt * vt);
            const double th = atan2(psnew[n + i], psnew[i]);
            const double nx = cos(th), ny = sin(th), tx = -ny, ty = nx;
            const double sn = ss[i] * nx + ss[n + i] * ny;
            int elnew = el[i];
            if (elnew == CO_SLIP && sn > 0.0) elnew = CO_ADHES;
            const double st = ss[i] * tx + ss[n + i] * ty, rt = -st;
            const double qn = q[i] * nx + q[n + i] * ny, qt = q[i] * tx + q[n + i] * ty;
            const double d2 = dscl[i] * dscl[i];
            if (elnew == CO_ADHES) {
                c0adh += 2.0 * d2 * (ss[i] * q[i] + ss[n + i] * q[n + i]);
                c1adh += 2.0 * d2 * (q[i] * q[i] + q[n + i] * q[n + i]);
                rhonew += d2 * ((ss[i] + alpha_j * q[i]) * (ss[i] + alpha_j * q[i]) + (ss[n + i] + alpha_j * q[n + i]) * (ss[n + i] + alpha_j * q[n + i]));
            } else if (elnew == CO_SLIP) {
                c0slp -= 2.0 * d2 * (rt * qt - rt * sn * dth_da);
                c1slp -= 2.0 * d2 * (-qt * qt + qt * sn * dth_da - rt * qn * dth_da);
                c2slp -= 2.0 * d2 * (qt * qn * dth_da);
                rhonew += d2 * (st + alpha_j * qt) * (st + alpha_j * qt);
            }
        }
        const double drho_da = c0adh + c0slp + alpha_j * (c1adh + c1slp) + alpha_j * alpha_j * c2slp;
        /* insert in the table (1-based, sorted on alpha) */
        int itb = 1;
        while (itb <= j && alpha_j > tbl_alpha[itb]) itb++;
        itu = itb;
        for (int k = j; k >= itb; k--) { tbl_alpha[k + 1] = tbl_alpha[k]; tbl_rho[k + 1] = tbl_rho[k]; tbl_drho[k + 1] = tbl_drho[k]; }
        if (has_bracket) {
            if (ita >= itu) ita++;
            if (itb_ >= itu) itb_++;
            if (itx >= itu) itx++;
            if (itw >= itu) itw++;
            if (itv >= itu) itv++;
        }
        j++;
        tbl_alpha[itb] = alpha_j; tbl_rho[itb] = rhonew; tbl_drho[itb] = drho_da;
        const int prev_bracket = has_bracket;
        has_bracket = 0;
        ibrack = 1;
        while (ibrack < j - 1 && !has_bracket) {
            has_bracket = (tbl_rho[ibrack + 1] <= fmin(tbl_rho[ibrack], tbl_rho[ibrack + 2]));
            if (!has_bracket) ibrack++;
        }
        if (has_bracket && !prev_bracket) {
            ita = ibrack; itx = ibrack + 1; itb_ = ibrack + 2;
            if (tbl_rho[ita] < tbl_rho[itb_]) { itw = ita; itv = itb_; } else { itw = itb_; itv = ita; }
            dst_prv1 = tbl_alpha[itx] - tbl_alpha[itw];
            dst_prv2 = tbl_alpha[itw] - tbl_alpha[itv];
        } else if (has_bracket) {
            if (tbl_rho[itu] <= tbl_rho[itx]) {
                if (tbl_alpha[itu] >= tbl_alpha[itx]) ita = itx; else itb_ = itx;
                itv = itw; itw = itx; itx = itu;
            } else {
                if (tbl_alpha[itu] < tbl_alpha[itx]) ita = itu; else itb_ = itu;
                if (tbl_rho[itu] <= tbl_rho[itw] || tbl_alpha[itw] == tbl_alpha[itx]) { itv = itw; itw = itu; }
                else if (tbl_rho[itu] <= tbl_rho[itv] || tbl_alpha[itv] == tbl_alpha[itx] || tbl_alpha[itv] == tbl_alpha[itw]) itv = itu;
            }
            dst_prv2 = dst_prv1;
            dst_prv1 = dst_cur;
        }
        alpha_prv = alpha_j;
        const double a_prev = alpha_j;
        if (has_bracket) {
            const double xmid = 0.5 * (tbl_alpha[ita] + tbl_alpha[itb_]);
            const double tolx = eps_j * fabs(tbl_alpha[itx]) + tiny;
            const int ldone = (fabs(tbl_alpha[itx] - xmid) <= 2.0 * tolx - 0.5 * (tbl_alpha[itb_] - tbl_alpha[ita]));
            if (!ldone) {
                int use_parab = 0;
                if (fabs(dst_prv2) > tolx) {
                    double br = (tbl_alpha[itx] - tbl_alpha[itw]) * (tbl_rho[itx] - tbl_rho[itv]);
                    double bq = (tbl_alpha[itx] - tbl_alpha[itv]) * (tbl_rho[itx] - tbl_rho[itw]);
                    double bp = (tbl_alpha[itx] - tbl_alpha[itv]) * bq - (tbl_alpha[itx] - tbl_alpha[itw]) * br;
                    bq = 2.0 * (bq - br);
                    if (bq > 0.0) bp = -bp;
                    bq = fabs(bq);
                    if (!(fabs(bp) >= fabs(0.5 * bq * dst_prv2) || bp <= bq * (tbl_alpha[ita] - tbl_alpha[itx]) ||
                          bp >= bq * (tbl_alpha[itb_] - tbl_alpha[itx]))) {
                        use_parab = 1;
                        dst_cur = bp / bq;
                        alpha_j = tbl_alpha[itx] + dst_cur;
                        if (alpha_j - tbl_alpha[ita] < 2.0 * tolx || tbl_alpha[itb_] - alpha_j < 2.0 * tolx)
                            dst_cur = tolx * copysign(1.0, xmid - tbl_alpha[itx]);
                    }
                }
                if (!use_parab) {
                    if (tbl_alpha[itx] >= xmid) dst_prv1 = tbl_alpha[ita] - tbl_alpha[itx];
                    else dst_prv1 = tbl_alpha[itb_] - tbl_alpha[itx];
                    dst_cur = 0.381966 * dst_prv1;
                }
                alpha_j = tbl_alpha[itx] + dst_cur;
            }
        } else if (j == 1) {
            alpha_j = alpha_j + 0.6 * alpha0;
        } else if (tbl_rho[j] < tbl_rho[1]) {
            const int it = j - 1;
            const double da = tbl_alpha[it + 1] - tbl_alpha[it];
            if (j < 3) { const double dd = (tbl_drho[it + 1] - tbl_drho[it]) / da; alpha_j = tbl_alpha[it + 1] + fmin(3.0 * da, -tbl_drho[it] / dd); }
            else alpha_j = tbl_alpha[it + 1] + 3.0 * da;
        } else {
            const int it = 1;
            const double da = tbl_alpha[it + 1] - tbl_alpha[it];
            if (j < 3) { const double dd = (tbl_drho[it + 1] - tbl_drho[it]) / da; alpha_j = tbl_alpha[it] + fmin(-3.0 * da, tbl_drho[it] / dd); }
            else alpha_j = tbl_alpha[it] - 3.0 * da;
        }
        stop_j = (j >= max_j || fabs(alpha_j - a_prev) < eps_j * fmax(fabs(alpha_j), fabs(a_prev)));
    }
    *nobracket = !has_bracket;
    if (j >= max_j) {
        int im = 1;
        for (int k = 2; k <= j; k++) if (tbl_rho[k] < tbl_rho[im]) im = k;
        alpha_j = tbl_alpha[im];
    }
    *alpha_out = alpha_j;
    free(dpnew); free(psnew);
    return j;
}

/* solv_input, m_sinput.f90:649-680: the .inp record (fdecay, betath, kdowfb, d_ifc, d_lin, d_cns, d_slp, pow_s) */
void co_gdparams_set(const double gd[8], co_gdparams *sp)
{
    sp->fdecay = gd[0]; sp->betath = gd[1]; sp->kdowfb = (int) gd[2];
    sp->d_ifc = fmax(0.01, gd[3]); sp->d_lin = gd[4]; sp->d_cns = fmax(0.01, gd[5]); sp->d_slp = fmax(0.01, gd[6]);
    sp->pow_s = fmax(0.01, fmin(10.0, gd[7]));
    if (sp->d_lin * (sp->d_cns - sp->d_ifc) < 0.0) { sp->d_lin = 0.0; sp->d_cns = sp->d_ifc; }
    sp->kdown = 1;
    if (sp->fdecay > 0.999) sp->gd_meth = 1;
    else if (sp->fdecay < 0.001) { sp->gd_meth = 2; sp->kdown = (int) fmax(1.0, nearbyint(-sp->fdecay)); }
    else sp->gd_meth = 3;
}

/* :9-606.  ps, ss: [3][npot]; ws: [2..3][npot]; returns itgd (negative: stagnation estimate), *lstagn */
int co_gdsteady(co_ctx *cx, int mx, int my, int maxgd, double eps, const double *ws, co_inflcf *cs, const double *mus,
                co_eldiv *igs, double *ps, double *ss, const co_gdparams *sp, double *err, int *lstagn)
{
    const int n = mx * my;
    const double tiny = 1e-12, fac_v = 1000.0;
    double coefs[2][2];
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) coefs[a][b] = cs->ga_inv * CO_CF(cs, co_cf_ptr(cs, a + 1, b + 1), 0, 0);
    double *g = (double *) calloc(n, sizeof(double)), *dp = (double *) calloc(3L * n, sizeof(double));
    double *dscl = (double *) malloc(sizeof(double) * n), *nn = (double *) calloc(2L * n, sizeof(double));
    double *t = (double *) calloc(2L * n, sizeof(double)), *r = (double *) calloc(3L * n, sizeof(double));
    double *dv = (double *) calloc(3L * n, sizeof(double)), *v = (double *) calloc(3L * n, sizeof(double));
    double *q = (double *) calloc(3L * n, sizeof(double)), *psopt = (double *) calloc(2L * n, sizeof(double));
    double *pold = (double *) calloc(2L * n, sizeof(double));
    int *elold = (int *) malloc(sizeof(int) * n), *elopt = (int *) malloc(sizeof(int) * n);
    int *el = igs->el;
    int nc = 0;
    for (int i = 0; i < n; i++) if (el[i] >= CO_ADHES) nc++;
    const double facnel = (double) sqrtf((float) n / (float) nc);
    for (int i = 0; i < n; i++) { g[i] = mus[i] * ps[2L * n + i]; dscl[i] = 1.0; }
    int itgd = 0, it_fb = -99, lchanged = 0;
    double dif = 2.0, difid = 1.0, dif1 = 0.0, beta = 1.0, alpha = 0.0, alpha0 = 0.0;
    *lstagn = 0;
    set_nt(n, ps, nn, t);
    compute_dp(mx, my, ps, dp);
    memset(ss, 0, sizeof(double) * 2 * n);
    co_vecaijpj(cx, igs, CO_ALLINT, ss, CO_TANG, dp, igs, CO_TANG, cs);
    for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (el[i] >= CO_ADHES) ss[(long) k * n + i] += ws[(long) k * n + i];
    compute_diagscaling(mx, my, el, mus, coefs, ps, ss, sp, dscl);
    set_residual(n, el, t, ss, dscl, r);
    {
        double m = 0.0;
        for (int i = 0; i < 2 * n; i++) m = fmax(m, fabs(r[i]));
        if (m < tiny) dif = 0.0;
    }
    while ((lchanged || dif > difid) && itgd < maxgd) {
        itgd++;
        memcpy(dv, r, sizeof(double) * 2 * n);
        project_searchdir(mx, my, el, g, itgd, &it_fb, beta, t, dv, v, sp, fac_v);
        co_vecaijpj(cx, igs, CO_ALLINT, q, CO_TANG, dv, igs, CO_TANG, cs);
        if (itgd <= 1) {
            double a0 = 0.0, a1 = 0.0;
            for (int i = 0; i < n; i++) { a0 += r[i] * r[i] + r[n + i] * r[n + i]; a1 += (r[i] * q[i] + r[n + i] * q[n + i]) * dscl[i]; }
            alpha0 = a0 / a1;
        } else alpha0 = 0.5 * (alpha0 + alpha);
        int nobr = 0;
        perform_linesearch(mx, my, el, g, dscl, dp, ss, dv, v, q, t, alpha0, &alpha, &nobr);
        {
            double sq = 0.0, sr = 0.0;
            for (int i = 0; i < n; i++) if (el[i] >= CO_ADHES) {
                sq += dscl[i] * dscl[i] * (q[i] * q[i] + q[n + i] * q[n + i]);
                sr += r[i] * r[i] + r[n + i] * r[n + i];
            }
            beta = alpha * sqrt(sq) / sqrt(sr);
        }
        lchanged = 0;
        memcpy(elold, el, sizeof(int) * n);
        memcpy(pold, ps, sizeof(double) * 2 * n);
        for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (el[i] >= CO_ADHES) dp[(long) k * n + i] += alpha * dv[(long) k * n + i];
        apply_trcbnd(mx, my, el, g, dp, ps);
        set_nt(n, ps, nn, t);
        memset(ss, 0, sizeof(double) * 2 * n);
        co_vecaijpj(cx, igs, CO_ALLINT, ss, CO_TANG, dp, igs, CO_TANG, cs);
        for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (el[i] >= CO_ADHES) ss[(long) k * n + i] += ws[(long) k * n + i];
        for (int i = 0; i < n; i++) {                                   /* solve_elmtrc :851-897 */
            double pr[3] = { ps[i], ps[n + i], ps[2L * n + i] }, si[2] = { ss[i], ss[n + i] };
            int e = el[i];
            co_plstrc(&e, coefs, 1e-6, 1.0, 1.0, pr, mus[i], si);
            elopt[i] = e; psopt[i] = pr[0]; psopt[n + i] = pr[1];
        }
        for (int i = 0; i < n; i++) if (el[i] == CO_ADHES) {
            const double pa = sqrt(ps[i] * ps[i] + ps[n + i] * ps[n + i]);
            if (pa >= g[i] - tiny) { el[i] = CO_SLIP; lchanged = 1; }
        }
        for (int i = 0; i < n; i++) if (elold[i] == CO_SLIP) {
            const double snrm = nn[i] * ss[i] + nn[n + i] * ss[n + i];
            const double cosdth = (nn[i] * psopt[i] + nn[n + i] * psopt[n + i]) / g[i];
            if ((snrm > 0.0 && (dscl[i] >= 0.001 || elopt[i] == CO_ADHES)) || cosdth <= -0.5) { el[i] = CO_ADHES; lchanged = 1; }
        }
        compute_diagscaling(mx, my, el, mus, coefs, ps, ss, sp, dscl);
        set_residual(n, el, t, ss, dscl, r);
        for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (el[i] >= CO_ADHES) pold[(long) k * n + i] -= ps[(long) k * n + i];
        const double facdif = beta >= 0.1 ? 1.0 : (beta >= 0.001 ? 0.1 / beta : 100.0);
        {
            double s1 = 0.0, s2 = 0.0;
            for (int i = 0; i < 2 * n; i++) { s1 += pold[i] * pold[i]; s2 += ps[i] * ps[i]; }
            dif = facnel * sqrt(s1 / (2.0 * n)) * facdif;
            difid = eps * fmax(1e-6, facnel * sqrt(s2 / (2.0 * n)));
        }
        if (itgd == 1) dif1 = dif;
    }
    double sr = 0.0; int cnt = 0;
    for (int i = 0; i < n; i++) if (el[i] >= CO_ADHES) { sr += r[i] * r[i] + r[n + i] * r[n + i]; cnt += 2; }
    const double resrms = sqrt(sr / (cnt > 1 ? cnt : 1));
    const double res_dp = resrms / fmax(coefs[0][0], coefs[1][1]);
    *err = dif;
    double conv = 1.0;
    if (dif * dif1 > 0.0 && itgd > 1) conv = exp(log(dif / dif1) / (itgd - 1));
    if (lchanged && itgd >= maxgd) *lstagn = 1;
    else if (dif > difid && conv > 1.0 && itgd >= maxgd) *lstagn = 1;
    else if (conv < 1.0 - tiny && res_dp > 5.0 * difid / (1.0 - conv)) { itgd = -itgd; *lstagn = 1; }
    free(g); free(dp); free(dscl); free(nn); free(t); free(r); free(dv); free(v); free(q); free(psopt); free(pold); free(elold); free(elopt);
    return itgd;
}
