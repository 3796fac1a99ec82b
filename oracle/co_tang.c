/*
 * co_tang.c -- ORACLE (test infrastructure, not product code).
 * Tangential contact problem for shifts (T=1) with the TangCG solver and the Newton-Raphson loop on the creepages,
 * plus the case driver contac/panprc.  Follows /root/reference/src/m_solvpt.f90:51-378 (solvpt), :1841-2442 (tangcg),
 * /root/reference/src/m_stang.f90:28-746 (stang, L=0 Coulomb friction), :749-951 (stang_rhs),
 * /root/reference/src/m_leadedge.f90:92-332 (sxbnd: for shifts facdt = 1, ii2j = 0),
 * /root/reference/src/m_sdis.f90:498-583 (set_tang_rhs), :587-768 (init_curr_data), m_scontc.f90:37-216, 356-553.
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

static void eldiv_count(const co_eldiv *g, int npot, int *nadh, int *nslip, int *nplast, int *nexter)
{
    *nadh = *nslip = *nplast = *nexter = 0;
    for (int i = 0; i < npot; i++) {
        if (g->el[i] == CO_ADHES) (*nadh)++; else if (g->el[i] == CO_SLIP) (*nslip)++;
        else if (g->el[i] == CO_PLAST) (*nplast)++; else (*nexter)++;
    }
}

static double dot_all2(int n, const double *a, const double *b)
{   /* gf3_dot(AllElm, a, b, ikTANG): ddot over x then y column */
    double s = 0.0, t;
    for (int k = 0; k < 2; k++) { t = 0.0; for (int i = 0; i < n; i++) t = t + a[(long) k * n + i] * b[(long) k * n + i]; s = s + t; }
    return s;
}

static double rms_all2(int n, const double *a)
{   /* gf3_rms(AllElm, a, ikTANG) = sqrt(sum / (2 n)) */
    double s = 0.0;
    for (int k = 0; k < 2; k++) { double t = 0.0; for (int i = 0; i < n; i++) t += a[(long) k * n + i] * a[(long) k * n + i]; s = s + t; }
    return sqrt(s / (2 * n > 1 ? 2 * n : 1));
}

/* m_solvpt.f90:1841-2442.  ps, ss, ws: [3][npot]; mu: [npot].  Elastic (no plasticity). */
void co_tangcg(co_ctx *cx, int npot, int maxcg, double eps, const double *ws, co_inflcf *cs, co_inflcf *ms,
               const double *mu, co_eldiv *igs, double *ps, double *ss, int *itcg_out, double *err)
{
    const int num_inn = 4, n = npot, my = igs->my;
    const double small = 1e-6, ga = cs->ga;
    double *g = (double *) calloc(n, sizeof(double)), *nn = (double *) calloc(3L * n, sizeof(double));
    double *t = (double *) calloc(3L * n, sizeof(double)), *r = (double *) calloc(3L * n, sizeof(double));
    double *z = (double *) calloc(3L * n, sizeof(double)), *v = (double *) calloc(3L * n, sizeof(double));
    double *q = (double *) calloc(3L * n, sizeof(double)), *pold = (double *) calloc(3L * n, sizeof(double));
    double *psx = ps, *psy = ps + n, *psn = ps + 2L * n, *ssx = ss, *ssy = ss + n;
    int use_fftprec = 1, lchanged = 0, nadh, nslip, nplast, nexter, itcg = 0, it_inn = 0;
    double alpha = 0.0, beta, rv, zq, vq, snrm, perp, ptabs, dif = 2.0, dif1 = 0.0, difid = 1.0, difinn = 0.0, ptang,
           trsinn = 0.0;
    const double c11 = CO_CF(cs, co_cf_ptr(cs, 1, 1), 0, 0), c22 = CO_CF(cs, co_cf_ptr(cs, 2, 2), 0, 0);

    co_fft_makeprec(cx, 1, cs, 1, ms);                                                   /* :1931-1934 */
    co_fft_makeprec(cx, 2, cs, 2, ms);
    eldiv_count(igs, n, &nadh, &nslip, &nplast, &nexter);
    const double facnel = (double) sqrtf((float) n / (float) (nadh + nslip + nplast));   /* REAL(4) arithmetic :1948 */
    for (int i = 0; i < n; i++) g[i] = mu[i] * psn[i];

#define SET_NT()                                                                          \
    for (int i = 0; i < n; i++) {                                                         \
        nn[2L * n + i] = atan2(psy[i], psx[i]);                                           \
        nn[i] = cos(nn[2L * n + i]); nn[n + i] = sin(nn[2L * n + i]);                     \
        t[i] = -nn[n + i]; t[n + i] = nn[i];                                              \
    }
#define PROJ_T(a)                                                                         \
    for (int i = 0; i < n; i++) if (igs->el[i] == CO_SLIP) {                              \
        perp = t[i] * (a)[i] + t[n + i] * (a)[n + i];                                     \
        (a)[i] = perp * t[i]; (a)[n + i] = perp * t[n + i];                               \
    }
    SET_NT();
    for (int i = 0; i < 2 * n; i++) ss[i] = 0.0;                                          /* :1980-1990 */
    co_vecaijpj(cx, igs, CO_ALLINT, ss, CO_TANG, ps, igs, CO_TANG, cs);
    for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (igs->el[i] >= CO_ADHES) ss[(long) k * n + i] += 1.0 * ws[(long) k * n + i];
    for (int i = 0; i < 2 * n; i++) r[i] = -1.0 * ss[i];
    PROJ_T(r);

    while ((lchanged || dif > difid) && itcg < maxcg) {                                   /* :2005 */
        itcg++; it_inn++;
        if (use_fftprec) {
            if (nslip + nplast > 3 * nadh && my > 1) use_fftprec = 0;
            if (itcg >= maxcg / 2) use_fftprec = 0;
            if (!use_fftprec) lchanged = 1;
        }
        if (use_fftprec) {
            co_vecaijpj(cx, igs, CO_ALLINT, z, CO_X, r, igs, CO_X, ms);
            co_vecaijpj(cx, igs, CO_ALLINT, z, CO_Y, r, igs, CO_Y, ms);
        } else memcpy(z, r, sizeof(double) * 2 * n);
        for (int i = 0; i < n; i++) {                                                     /* :2037-2046 */
            if (igs->el[i] == CO_SLIP) {
                snrm = -ssx[i] * nn[i] - ssy[i] * nn[n + i];
                z[i] = z[i] / (c11 + ga * snrm / g[i]);
                z[n + i] = z[n + i] / (c22 + ga * snrm / g[i]);
            } else { z[i] = z[i] / c11; z[n + i] = z[n + i] / c22; }
        }
        PROJ_T(z);
        if (itcg <= 1 || lchanged) memcpy(v, z, sizeof(double) * 2 * n);
        else {
            zq = dot_all2(n, z, q);
            vq = dot_all2(n, v, q);
            if (fabs(zq) < 1e-60 || fabs(vq) < small * fabs(zq)) beta = 0.0; else beta = -zq / vq;
            for (int i = 0; i < 2 * n; i++) v[i] = beta * v[i];
            for (int i = 0; i < 2 * n; i++) v[i] = v[i] + 1.0 * z[i];
        }
        co_vecaijpj(cx, igs, CO_ALLINT, q, CO_TANG, v, igs, CO_TANG, cs);                  /* :2094 */
        PROJ_T(q);
        for (int i = 0; i < n; i++) if (igs->el[i] == CO_SLIP) {
            snrm = (nn[i] * ssx[i] + nn[n + i] * ssy[i]) / g[i];
            q[i] = q[i] - snrm * v[i]; q[n + i] = q[n + i] - snrm * v[n + i];
        }
        rv = dot_all2(n, r, v);
        vq = dot_all2(n, v, q);
        if (fabs(rv) < 1e-60) alpha = 0.0; else if (fabs(vq) < small * fabs(rv)) alpha = 1.0; else alpha = rv / vq;

        memcpy(pold, ps, sizeof(double) * 2 * n);                                         /* :2170-2185 */
        for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (igs->el[i] >= CO_ADHES) ps[(long) k * n + i] += alpha * v[(long) k * n + i];
        for (int i = 0; i < n; i++) if (igs->el[i] == CO_SLIP) {
            ptabs = sqrt(psx[i] * psx[i] + psy[i] * psy[i]);
            psx[i] = psx[i] * g[i] / ptabs; psy[i] = psy[i] * g[i] / ptabs;
        }
        dif = alpha * facnel * rms_all2(n, v);
        difinn = difinn + dif;
        if (itcg <= 5 || itcg % 5 == 1) {
            ptang = facnel * rms_all2(n, ps);
            difid = eps * fmax(1e-6, ptang);
            trsinn = (double) 0.01f * ptang;
        }
        if (it_inn >= num_inn || dif <= difid || difinn > trsinn) {                        /* :2227 */
            lchanged = 0;
            for (int i = 0; i < n; i++) if (igs->el[i] == CO_ADHES) {
                ptabs = sqrt(psx[i] * psx[i] + psy[i] * psy[i]);
                if (ptabs > g[i]) {
                    igs->el[i] = CO_SLIP;
                    psx[i] = psx[i] * g[i] / ptabs; psy[i] = psy[i] * g[i] / ptabs;
                    lchanged = 1;
                }
            }
            co_vecaijpj(cx, igs, CO_ALLINT, ss, CO_TANG, ps, igs, CO_TANG, cs);
            for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (igs->el[i] >= CO_ADHES) ss[(long) k * n + i] += 1.0 * ws[(long) k * n + i];
            for (int i = 0; i < n; i++) if (igs->el[i] == CO_SLIP) {
                snrm = ssx[i] * psx[i] + ssy[i] * psy[i];
                if (snrm > 0.0) { igs->el[i] = CO_ADHES; lchanged = 1; }
            }
            for (int i = 0; i < 2 * n; i++) pold[i] = pold[i] + (-1.0) * ps[i];
            dif = facnel * rms_all2(n, pold);
            it_inn = 0;
            difinn = 0.0;
            if (lchanged) eldiv_count(igs, n, &nadh, &nslip, &nplast, &nexter);
        }
        if (itcg == 1) dif1 = dif;
        if ((lchanged || dif > difid) && itcg < maxcg) {                                   /* :2330-2380 */
            if (it_inn <= 0) {
                SET_NT();
                for (int i = 0; i < 2 * n; i++) r[i] = -1.0 * ss[i];
                PROJ_T(r);
            } else {
                for (int k = 0; k < 2; k++) for (int i = 0; i < n; i++) if (igs->el[i] >= CO_ADHES) r[(long) k * n + i] += (-alpha) * q[(long) k * n + i];
            }
        }
    }
    (void) dif1;
    *err = dif;
    *itcg_out = itcg;
    free(g); free(nn); free(t); free(r); free(z); free(v); free(q); free(pold);
#undef SET_NT
#undef PROJ_T
}

/* m_stang.f90:749-951: shifts (T=1: facdt = 1, previous tractions pv with cv = cs), transient rolling (T=2: near the
 * leading edge, ii2j > 0, equation (1b): u' replaced by ubnd = subnd(pv, cs), :888-925) and steady rolling (T=3: uvn from
 * the current pressures with the shifted coefficients cv, uvt and ubnd left to the solver) */
static void tang_rhs(co_ctx *cx, int npot, int is_ssrol, const double *facdt, const co_eldiv *igs, const double *hs,
                     const double *ps, const double *pv, co_inflcf *cs, co_inflcf *cv, co_leadedge *lg, double *wsfix)
{
    double *usn = (double *) calloc(3L * npot, sizeof(double)), *uvn = (double *) calloc(3L * npot, sizeof(double));
    double *uvt = (double *) calloc(3L * npot, sizeof(double));
    co_vecaijpj(cx, igs, CO_ALLINT, usn, CO_TANG, ps, igs, CO_Z, cs);
    if (!is_ssrol) {
        co_vecaijpj(cx, igs, CO_ALLINT, uvn, CO_TANG, pv, igs, CO_Z, cv);
        co_vecaijpj(cx, igs, CO_ALLINT, uvt, CO_TANG, pv, igs, CO_TANG, cv);
    } else
        co_vecaijpj(cx, igs, CO_ALLINT, uvn, CO_TANG, ps, igs, CO_Z, cv);
    if (!is_ssrol) {
        /* :892 subnd(cgrid, pv, cs, ledg): the row sums run over the element division of pv (previous time); pv is zero
         * outside it, so the whole grid gives the same sums */
        co_eldiv all;
        co_eldiv_init(&all, igs->mx, igs->my);
        for (int i = 0; i < npot; i++) all.el[i] = CO_ADHES;
        co_areas(&all);
        co_subnd(cx, igs->mx, igs->my, pv, &all, cs, lg);
        co_eldiv_free(&all);
    }
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < npot; i++)
            if (igs->el[i] >= CO_ADHES) {
                const long o = (long) k * npot + i;
                const double wsrig = -facdt[i] * hs[o];
                if (is_ssrol) wsfix[o] = wsrig + usn[o] - uvn[o];
                else if (lg->ii2j[i] <= 0) wsfix[o] = wsrig + usn[o] - uvn[o] - uvt[o];
                else wsfix[o] = wsrig + usn[o] - lg->ubnd[2 * lg->ii2j[i] + k];            /* :918-924 */
            }
    free(usn); free(uvn); free(uvt);
}

/* solver selection and relaxation parameters of stang (m_stang.f90:144-223) */
typedef struct { int solver; double omegah, omegas, dq; const double *facdt; int info;
                 co_inflcf *csv; co_leadedge *lg; const int *iel; int is_ssrol; } tang_opts;
enum { SOLV_TANGCG = 0, SOLV_STDYGS = 1, SOLV_CNVXGS = 2, SOLV_GDSTDY = 3 };

/* one call of the tangential solver + relative forces */
static void solve_once(co_ctx *cx, co_case *c, tang_opts *o, int npot, co_inflcf *cs, co_inflcf *ms, const double *wstot,
                       const double *mus, co_eldiv *igs, double *ps, double *ss, double dxdy, double muscal, double fntrue,
                       int *it, double *err, double *fx, double *fy)
{
    o->info = 0;
    if (o->solver == SOLV_TANGCG) co_tangcg(cx, npot, c->maxgs, c->eps, wstot, cs, ms, mus, igs, ps, ss, it, err);
    else {
        int k = 0;
        for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) k++;
        if (o->solver == SOLV_GDSTDY) {                    /* tang_solver, m_solvpt.f90:459-484 */
            co_gdparams sp;
            int lstagn = 0;
            co_gdparams_set(c->gd, &sp);
            *it = co_gdsteady(cx, igs->mx, igs->my, c->maxgs, c->eps, wstot, cs, mus, igs, ps, ss, &sp, err, &lstagn);
            if (lstagn) {
                c->gd_fallback++;
                co_stdygs(cx, igs->mx, igs->my, wstot, cs, mus, igs, ps, ss, k, c->eps, c->maxgs, o->omegah, o->omegas, &o->info, it, err);
            }
        } else if (o->solver == SOLV_STDYGS)
            co_stdygs(cx, igs->mx, igs->my, wstot, cs, mus, igs, ps, ss, k, c->eps, c->maxgs, o->omegah, o->omegas, &o->info, it, err);
        else
            co_cnvxgs(cx, igs->mx, igs->my, o->is_ssrol, wstot, cs, o->csv, o->lg, mus, igs, ps, ss, k, o->iel, c->eps, c->maxgs,
                      o->omegah, o->omegas, &o->info, it, err);
    }
    double sx = 0.0, sy = 0.0;
    for (int i = 0; i < npot; i++) sx = sx + ps[i];
    for (int i = 0; i < npot; i++) sy = sy + ps[npot + i];
    *fx = dxdy * sx / (muscal * fntrue);
    *fy = dxdy * sy / (muscal * fntrue);
    if (c->nr_n < CO_MAXNR) { c->nr_itcg[c->nr_n] = *it; c->nr_cksi[c->nr_n] = c->cksi; c->nr_ceta[c->nr_n] = c->ceta;
                              c->nr_fx[c->nr_n] = *fx; c->nr_fy[c->nr_n] = *fy; c->nr_n++; }
}

/* m_solvpt.f90:51-378 */
static void solvpt(co_ctx *cx, co_case *c, tang_opts *o, int npot, co_inflcf *cs, co_inflcf *ms, const double *wsfix, const double *mus,
                   co_eldiv *igs, double *ps, double *ss, double dxdy, double muscal, double fntrue, double sens[2][2],
                   int *itgs, double *err)
{
    double *wstot = (double *) calloc(3L * npot, sizeof(double));
    int it, nadh, nslip, nplast, nexter, itnr;
    double fxkp1, fykp1, fxk, fyk, df, dfx, dfy, dcksi = 0.0, dceta = 0.0, det, dfxk, dfyk, dpxavg, dpyavg;
    const double dq = o->dq, *facdt = o->facdt;
    *itgs = 0;
    co_areas(igs);
    memcpy(wstot, wsfix, sizeof(double) * 2 * npot);
    if (c->force3 >= 1) for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) wstot[i] = wstot[i] + facdt[i] * c->cksi * dq;
    if (c->force3 >= 2) for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) wstot[npot + i] = wstot[npot + i] + facdt[i] * c->ceta * dq;
    solve_once(cx, c, o, npot, cs, ms, wstot, mus, igs, ps, ss, dxdy, muscal, fntrue, &it, err, &fxkp1, &fykp1);
    *itgs += it;
    eldiv_count(igs, npot, &nadh, &nslip, &nplast, &nexter);
    if (c->force3 >= 1 && o->info <= 1) {
        itnr = 0;
        df = fabs(c->fxrel - fxkp1);
        if (c->force3 >= 2) df = df + fabs(c->fyrel - fykp1);
        while (df > c->eps && itnr < c->maxnr && o->info <= 1) {
            itnr++;
            dfx = c->fxrel - fxkp1;
            dfy = c->fyrel - fykp1;
            if (c->force3 == 1) {
                dceta = 0.0;
                if (fabs(sens[0][0]) > (double) 1e-6f) dcksi = dfx / sens[0][0]; else dcksi = (double) 0.00003f;
            } else {
                det = sens[0][0] * sens[1][1] - sens[1][0] * sens[0][1];          /* sens[out fx/fy][in ksi/eta] */
                if (det > c->eps) {
                    dcksi = (sens[1][1] * dfx - sens[1][0] * dfy) / det;
                    dceta = (-sens[0][1] * dfx + sens[0][0] * dfy) / det;
                } else { dcksi = 0.000003; dceta = 0.000003; }
            }
            for (int ifxy = 1; ifxy <= c->force3; ifxy++) {
                fxk = fxkp1; fyk = fykp1;
                if (ifxy == 1) {
                    c->cksi = c->cksi + dcksi;
                    for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) wstot[i] = wstot[i] + facdt[i] * dcksi * dq;
                } else {
                    c->ceta = c->ceta + dceta;
                    for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) wstot[npot + i] = wstot[npot + i] + facdt[i] * dceta * dq;
                }
                solve_once(cx, c, o, npot, cs, ms, wstot, mus, igs, ps, ss, dxdy, muscal, fntrue, &it, err, &fxkp1, &fykp1);
                *itgs += it;
                eldiv_count(igs, npot, &nadh, &nslip, &nplast, &nexter);
                dfxk = fxkp1 - fxk; dfyk = fykp1 - fyk;
                const double ncon = (double) (nadh + nslip + nplast);
                if (ifxy == 1) {
                    if (nadh > 0) {
                        dpxavg = dfxk * muscal * fntrue / (ncon * dxdy);
                        if (sens[0][0] == 0.0 || fabs(dpxavg) > 10.0 * *err) sens[0][0] = dfxk / dcksi;
                        dpyavg = dfyk * muscal * fntrue / (ncon * dxdy);
                        if (fabs(dpyavg) > 10.0 * *err) sens[1][0] = dfyk / dcksi;
                    } else if (fabs(sens[0][0]) < (double) 1e-6f) sens[0][0] = fxkp1 / c->cksi;
                } else {
                    if (nadh > 0) {
                        dpxavg = dfxk * muscal * fntrue / (ncon * dxdy);
                        if (fabs(dpxavg) > 10.0 * *err) sens[0][1] = dfxk / dceta;
                        dpyavg = dfyk * muscal * fntrue / (ncon * dxdy);
                        if (sens[1][1] == 0.0 || fabs(dpyavg) > 10.0 * *err) sens[1][1] = dfyk / dceta;
                    } else if (fabs(sens[1][1]) < (double) 1e-6f) sens[1][1] = fykp1 / c->ceta;
                }
                df = fabs(c->fxrel - fxkp1);
                if (c->force3 >= 2) df = df + fabs(c->fyrel - fykp1);
            }
        }
        *err = *err + 2.0 * df * muscal * fntrue / ((nadh + 2 + nplast) * dxdy);  /* reference uses the constant Slip=2, :361 */
    } else { sens[0][0] = sens[1][0] = sens[0][1] = sens[1][1] = 0.0; }
    free(wstot);
}

/* m_stang.f90:28-746 for L=0, elastic material: shifts with TangCG, steady rolling with SteadyGS */
static int stang(co_ctx *cx, co_case *c, int npot, co_inflcf *cs, co_inflcf *cv, co_inflcf *csv, co_inflcf *ms, const double *hs,
                 const double *pv, const double *x, co_eldiv *igs, double *ps, double *ss, double dxdy, double muscal,
                 double fntrue, double dq, double sens[2][2], int *itgs_out)
{
    const int mx = igs->mx, my = igs->my, is_ssrol = (c->tang == 3);
    double *wsfix = (double *) calloc(3L * npot, sizeof(double)), *mus = (double *) malloc(sizeof(double) * npot);
    double *tmp = (double *) calloc(3L * npot, sizeof(double)), *facdt = (double *) malloc(sizeof(double) * npot);
    int ittang = 0, itgs = 0, zready = 0, it, nadh, nslip, nplast, nexter;
    double errpt = 0.0, tol, tol1, tol2, pabs, ww;
    co_leadedge lg;
    memset(&lg, 0, sizeof(lg));
    int *iel = (int *) malloc(sizeof(int) * npot);
    tang_opts o = { SOLV_TANGCG, 1.0, 1.0, dq, facdt, 0, csv, &lg, iel, is_ssrol };
    int k = 0;
    for (int i = 0; i < npot; i++) if (igs->el[i] >= CO_ADHES) iel[k++] = i;
    {                                                                                      /* :129-223 */
        int icount = 0;
        for (int iy = 1; iy <= my; iy++) if (igs->el[1 + (iy - 1) * mx - 1] >= CO_ADHES) icount++;
        if (is_ssrol) o.solver = (c->gausei == 5) ? SOLV_GDSTDY : (c->gausei != 2 ? SOLV_STDYGS : SOLV_CNVXGS);
        else o.solver = (c->gausei != 2) ? SOLV_TANGCG : SOLV_CNVXGS;
        if (icount > 0 && (o.solver == SOLV_STDYGS || o.solver == SOLV_GDSTDY)) o.solver = SOLV_CNVXGS;
        if (c->gausei == 0 || c->gausei == 4 || c->gausei == 5) {
            if (k <= 25) { o.omegah = 1.0; o.omegas = 1.0; }
            else if (is_ssrol && o.solver == SOLV_CNVXGS) { o.omegah = 0.5; o.omegas = 0.5; }
            else if (is_ssrol) {
                if (c->dx / c->dy <= 5.0) { o.omegah = 0.9; o.omegas = 1.0; }
                else if (c->dx / c->dy <= 15.0) { o.omegah = 0.8; o.omegas = 0.8; }
                else { o.omegah = 0.8; o.omegas = 0.6; }
            } else { o.omegah = 0.5; o.omegas = 0.5; }
        } else { o.omegah = c->omegah; o.omegas = c->omegas; }
        co_sxbnd(mx, my, c->tang == 2 || c->tang == 3, o.solver != SOLV_STDYGS, igs, x, c->dx, dq, &lg);
        memcpy(facdt, lg.facdt, sizeof(double) * npot);
    }
    for (int i = 0; i < npot; i++) mus[i] = c->fstat;
    tang_rhs(cx, npot, is_ssrol, facdt, igs, hs, ps, pv, cs, cv, &lg, wsfix);
    while (!zready && ittang < c->maxin) {                                                /* :376 */
        ittang++;
        zready = 1;
        solvpt(cx, c, &o, npot, cs, ms, wsfix, mus, igs, ps, ss, dxdy, muscal, fntrue, sens, &it, &errpt);
        itgs += it;
        if (o.info >= 3) { zready = 1; ittang = -1; }                                     /* :420-427 */
        int newins = 0;
        tol = sqrt(2.0) * errpt;
        for (int i = 0; i < npot; i++) if (igs->el[i] == CO_ADHES) {                      /* :434-453 */
            pabs = sqrt(ps[i] * ps[i] + ps[npot + i] * ps[npot + i]);
            if (pabs >= mus[i] * ps[2L * npot + i] + tol) {
                newins++;
                igs->el[i] = CO_SLIP;
                ps[i] = ps[i] * mus[i] * ps[2L * npot + i] / pabs;
                ps[npot + i] = ps[npot + i] * mus[i] * ps[2L * npot + i] / pabs;
            }
        }
        if (newins != 0) zready = 0;
        if (zready) {                                                                     /* :463-510 */
            co_areas(igs);
            const int ii = (mx / 2 > 1 ? mx / 2 : 1) + ((my / 2 > 1 ? my / 2 : 1) - 1) * mx;
            for (int i = 0; i < 2 * npot; i++) tmp[i] = 1.0;
            tol1 = errpt * 2.0 * fabs(co_aijpj(ii, CO_X, tmp, igs, CO_X, cs));
            tol2 = errpt * 2.0 * fabs(co_aijpj(ii, CO_Y, tmp, igs, CO_Y, cs));
            int newadh = 0;
            for (int i = 0; i < npot; i++) if (igs->el[i] == CO_SLIP) {
                ww = ss[i] * ps[i] + ss[npot + i] * ps[npot + i];
                tol = tol1 * fabs(ps[i]) + tol2 * fabs(ps[npot + i]) + errpt * (fabs(ss[i]) + fabs(ss[npot + i]));
                if (ww > tol) { igs->el[i] = CO_ADHES; newadh++; }
            }
            if (newadh != 0) zready = 0;
            if (o.info > 0) zready = 1;                                                   /* :510 */
        }
    }
    eldiv_count(igs, npot, &nadh, &nslip, &nplast, &nexter);
    if (!zready) ittang = -1;
    *itgs_out = itgs;
    free(wsfix); free(mus); free(tmp); free(facdt); free(iel);
    co_leadedge_free(&lg);
    return ittang;
}

/* contac (m_scontc.f90:37-216) + panprc (:356-553) for module-3 cases: T = 0, 1 or 3 (SteadyGS), N = 0/1, F = 0/1/2, I = 0, P = 2 */
int co_contac(co_case *c)
{
    /* Hertzian input (hzsol + potcon_hertz, m_hertz.f90:29-91, m_hierarch_data.f90:1889-1959) */
    double hz_prm[6] = { 0, 0, 0, 0, 0, 0 };
    if (c->ipotcn == -1 || c->ipotcn == -3) {
        co_mater hm = { { c->gg[0], c->gg[1] }, { c->poiss[0], c->poiss[1] }, 0, 0, 0 };
        co_combin_mater(&hm);
        double cp, rho;
        co_hertz3d(hm.ga / (1.0 - hm.nu), c->ipotcn, &c->hz_a1, &c->hz_b1, &c->hz_aa, &c->hz_bb, c->norm, &c->pen, &c->fn, &cp, &rho);
        hz_prm[0] = c->hz_a1; hz_prm[2] = c->hz_b1;
        c->ibase = 1; c->prmudf = hz_prm;
        c->xl = -c->hz_scale * fmax(1e-9, c->hz_aa); c->yl = -c->hz_scale * fmax(1e-9, c->hz_bb);
        c->dx = (-c->xl - c->xl) / c->mx; c->dy = (-c->yl - c->yl) / c->my;
    }
    const int mx = c->mx, my = c->my, npot = mx * my;
    co_ctx *cx = co_ctx_new();
    cx->fullbox = c->fullbox;
    co_mater mat = { { c->gg[0], c->gg[1] }, { c->poiss[0], c->poiss[1] }, 0, 0, 0 };
    co_combin_mater(&mat);
    co_inflcf cs, cv, csv, ms;
    memset(&cs, 0, sizeof(cs)); memset(&cv, 0, sizeof(cv)); memset(&csv, 0, sizeof(csv)); memset(&ms, 0, sizeof(ms));
    /* check_roll_stepsize (m_sdis.f90:125-204): shifts chi = 0, dq = 1; SteadyGS forces chi = 0 (or pi), dq = dx */
    const int is_roll = (c->tang == 2 || c->tang == 3), is_ssrol = (c->tang == 3);
    double chi = 0.0, dq = 1.0;
    if (is_roll) {
        chi = c->chi; dq = c->dq;
        if (is_ssrol && c->gausei != 2) {
            if (fabs(chi) > 0.01 && fabs(chi - CO_PI) > 0.01) chi = 0.0;
            dq = c->dx;
        }
    }
    /* not restated: rolling directions other than +x */
    if (is_roll && fabs(chi) > 0.01) { co_ctx_free(cx); return -99; }
    co_sgencr(&mat, mx, my, c->dx, c->dy, is_roll, chi, dq, &cs, &cv, &csv, &ms);
    double *x = (double *) malloc(sizeof(double) * npot), *y = (double *) malloc(sizeof(double) * npot);
    double *hs = (double *) calloc(3L * npot, sizeof(double)), *ps = (double *) calloc(3L * npot, sizeof(double));
    double *pv = (double *) calloc(3L * npot, sizeof(double)), *ss = (double *) calloc(3L * npot, sizeof(double));
    double *po1 = (double *) calloc(3L * npot, sizeof(double));
    co_grid_coords(mx, my, c->xl, c->yl, c->dx, c->dy, x, y);
    co_set_norm_rhs(c->ibase, 1, npot, x, y, c->nn, c->prmudf, NULL, hs + 2L * npot);
    /* set_tang_rhs (m_sdis.f90:498-583): spin offset facphi*dq in rolling problems */
    const double facphi = (c->facphi > 0.0) ? c->facphi : 1.0 / 6.0;
    const double xofs_dq = is_roll ? cos(chi) * dq * facphi : 0.0, yofs_dq = is_roll ? sin(chi) * dq * facphi : 0.0;
    for (int i = 0; i < npot; i++) { hs[i] = 1.0 * -(y[i] + yofs_dq - 0.0) * c->cphi + 0.0; hs[npot + i] = 1.0 * (x[i] + xofs_dq - 0.0) * c->cphi + 0.0; }
    if (c->force3 == 0) for (int i = 0; i < npot; i++) hs[i] = hs[i] + c->cksi;
    if (c->force3 <= 1) for (int i = 0; i < npot; i++) hs[npot + i] = hs[npot + i] + c->ceta;
    for (int i = 0; i < 2 * npot; i++) hs[i] = -dq * hs[i];
    /* init_curr_data (I = 0) */
    co_eldiv igs;
    co_eldiv_init(&igs, mx, my);
    double pen = c->pen, fntrue = c->fn;
    if (c->pv_in) memcpy(pv, c->pv_in, sizeof(double) * 3 * npot);                            /* set_prev_data */
    const int iestim = (c->el_in && c->ps_in) ? c->iestim : 0;
    if (iestim == 0) {                                                                       /* init_curr_data, m_sdis.f90:587-768 */
        if (c->ipotcn == -1 || c->ipotcn == -3) {
            const double pnmax = 3.0 * fntrue / (2.0 * CO_PI * c->hz_aa * c->hz_bb);
            for (int i = 0; i < npot; i++) {
                const double f = 1.0 - (x[i] / c->hz_aa) * (x[i] / c->hz_aa) - (y[i] / c->hz_bb) * (y[i] / c->hz_bb);
                ps[2L * npot + i] = pnmax * sqrt(fmax(0.0, f));
                igs.el[i] = ps[2L * npot + i] > 1e-20 ? CO_ADHES : CO_EXTER;
            }
        } else
            co_eldiv0(c->norm, mx, my, c->dx, c->dy, c->ibase, c->prmudf, &mat, fntrue, &pen, hs + 2L * npot, &igs);
    } else {
        memcpy(igs.el, c->el_in, sizeof(int) * npot);
        memcpy(ps, c->ps_in, sizeof(double) * 3 * npot);
        if (iestim == 2 || c->tang == 0) memset(ps, 0, sizeof(double) * 2 * npot);
        if (iestim == 1)
            for (int i = 0; i < npot; i++) {
                if (igs.el[i] <= CO_EXTER) { ps[i] = 0.0; ps[npot + i] = 0.0; ps[2L * npot + i] = 0.0; }
                else if (igs.el[i] == CO_SLIP) {
                    const double pa = fmax(1e-10, sqrt(ps[i] * ps[i] + ps[npot + i] * ps[npot + i]));
                    ps[i] = c->fstat * ps[2L * npot + i] * ps[i] / pa; ps[npot + i] = c->fstat * ps[2L * npot + i] * ps[npot + i] / pa;
                }
            }
        if (iestim == 2) for (int i = 0; i < npot; i++) if (igs.el[i] >= CO_ADHES) igs.el[i] = CO_ADHES;
    }
    for (int i = 0; i < npot; i++) if (hs[2L * npot + i] > (double) 1e29f) igs.el[i] = CO_EXTER;
    co_areas(&igs);
    if (iestim == 0 || iestim == 2) {
        if (c->force3 >= 1) c->cksi = 1e-6;
        if (c->force3 == 2) c->ceta = 0.0;
    }
    double sens[2][2] = { { 0, 0 }, { 0, 0 } };

    /* panprc */
    co_solv solv = { c->maxgs, c->maxin, c->maxnr, c->maxout, c->eps };
    co_norm_info info;
    int itnorm = 0, ittang = 0, itout = 0, itgs = 0, itcg_norm = 0;
    double dif = 200.0, difid = 1.0;
    const double dxdy = c->dx * c->dy, muscal = c->fstat;
    c->nr_n = 0;
    int unsupported = 0;
    memcpy(po1, ps, sizeof(double) * 3 * npot);
    while (dif > difid && itout < c->maxout && itnorm >= 0 && ittang >= 0) {
        itout++;
        co_snorm(cx, c->norm, mx, my, dxdy, &solv, hs, &cs, &ms, &pen, &fntrue, &igs, ps, &info);
        itcg_norm += info.itcg;
        if (info.itnorm >= 0) itnorm += info.itnorm; else itnorm = -1;
        int ncon = 0;
        for (int i = 0; i < npot; i++) {
            if (igs.el[i] <= CO_EXTER) { ps[i] = 0.0; ps[npot + i] = 0.0; ps[2L * npot + i] = 0.0; } else ncon++;
        }
        if (itout <= 16) { c->pan_dif[itout - 1] = 0.0; c->pan_difid[itout - 1] = difid; }
        if (c->tang == 0 || ncon <= 0) dif = 0.0;
        else {
            int it = stang(cx, c, npot, &cs, &cv, &csv, &ms, hs, pv, x, &igs, ps, ss, dxdy, muscal, fntrue, dq, sens, &itgs);
            if (it == -99) { unsupported = 1; break; }
            if (it >= 0) ittang += it; else ittang = -1;
            for (int i = 0; i < 3 * npot; i++) po1[i] = po1[i] + (-1.0) * ps[i];
            double s1 = 0.0, s2 = 0.0; int cnt = 0;
            for (int k = 0; k < 3; k++) for (int i = 0; i < npot; i++) if (igs.el[i] >= CO_ADHES) { s1 += po1[(long) k * npot + i] * po1[(long) k * npot + i]; s2 += ps[(long) k * npot + i] * ps[(long) k * npot + i]; cnt++; }
            dif = sqrt(s1 / (cnt > 1 ? cnt : 1));
            difid = 5.0 * c->eps * sqrt(s2 / (cnt > 1 ? cnt : 1));
            memcpy(po1, ps, sizeof(double) * 3 * npot);
            if (itout <= 16) { c->pan_dif[itout - 1] = dif; c->pan_difid[itout - 1] = difid; }
        }
    }
    /* soutpt forces (m_soutpt.f90:424-436) */
    if (c->norm == 0) { double s = 0.0; for (int i = 0; i < npot; i++) s = s + ps[2L * npot + i]; fntrue = dxdy * s; }
    double sx = 0.0, sy = 0.0;
    for (int i = 0; i < npot; i++) sx = sx + ps[i];
    for (int i = 0; i < npot; i++) sy = sy + ps[npot + i];
    c->fx_out = (c->force3 == 0) ? dxdy * sx / (fntrue * muscal + CO_TINY) : c->fxrel;
    c->fy_out = (c->force3 <= 1) ? dxdy * sy / (fntrue * muscal + CO_TINY) : c->fyrel;
    c->itout = itout;
    c->pen_out = pen; c->fn_out = fntrue; c->itnorm = itnorm; c->ittang = ittang; c->itcg_norm = itcg_norm; c->itgs_tang = itgs;
    memcpy(c->el, igs.el, sizeof(int) * npot);
    memcpy(c->ps, ps, sizeof(double) * 3 * npot);
    if (c->ss) memcpy(c->ss, ss, sizeof(double) * 3 * npot);
    c->n_prod = cx->st.n_prod;
    co_eldiv_free(&igs); co_inflcf_free(&cs); co_inflcf_free(&cv); co_inflcf_free(&csv); co_inflcf_free(&ms);
    free(x); free(y); free(hs); free(ps); free(pv); free(ss); free(po1);
    co_ctx_free(cx);
    if (unsupported) return -99;
    return itnorm < 0 ? -27 : (ittang < 0 ? -28 : 0);
}
