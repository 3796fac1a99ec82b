/*
 * co_inflcf.c -- ORACLE (test infrastructure, not product code).
 * Influence coefficients of the elastic half-space for piecewise-constant tractions.
 * Follows /root/reference/src/m_visc.f90:127-376 (sgencr), :431-604 (elascf_pcwcns) and
 * /root/reference/src/m_hierarch_data.f90:1543-1570 (combin_mater, inflcf_mater).
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* m_hierarch_data.f90:1543-1553 */
void co_combin_mater(co_mater *m)
{
    m->ga = 2.0 / (1.0 / m->gg[0] + 1.0 / m->gg[1]);
    m->nu = m->ga * (m->poiss[0] / m->gg[0] + m->poiss[1] / m->gg[1]) / 2.0;
    m->ak = (m->ga / 4.0) * ((1.0 - 2.0 * m->poiss[0]) / m->gg[0] - (1.0 - 2.0 * m->poiss[1]) / m->gg[1]);
}

void co_inflcf_init(co_inflcf *c, int mx, int my, double dx, double dy)
{
    memset(c, 0, sizeof(*c));
    c->cf_mx = mx; c->cf_my = my; c->dx = dx; c->dy = dy;
    c->cf = (double *) calloc((size_t) 36 * mx * my, sizeof(double));
    c->nt_cpl = 1;                      /* inflcf_new default: coupling on until sgencr clears it */
    c->ga = 1.0; c->ga_inv = 1.0;
}

void co_inflcf_free(co_inflcf *c)
{
    free(c->cf); c->cf = NULL;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { free(c->fft_cf[i][j]); c->fft_cf[i][j] = NULL; }
}

/* m_hierarch_data.f90:1557-1570 (elastic half-space only: no third-body layer, no compressible sheet) */
void co_inflcf_mater(co_inflcf *c, const co_mater *m)
{
    c->ga = m->ga;
    c->ga_inv = 1.0 / m->ga;
    c->use_3bl = 0;  c->flx_3bl = 0.0;
    c->use_flxz = 0; c->flx_z = 0.0;
}

/* m_visc.f90:431-604 */
void co_elascf_pcwcns(double akv, double nuv, int mx, int my, double dx, double dy, double xshft, double yshft,
                      co_inflcf *cs)
{
    const int nxp = 2 * mx + 1, nyp = 2 * my + 1;
#define P(a, ix, iy) (a)[((long)((iy) + my)) * nxp + ((ix) + mx)]
    double *r = (double *) calloc((size_t) nxp * nyp, sizeof(double));
    double *x = (double *) calloc((size_t) nxp * nyp, sizeof(double));
    double *y = (double *) calloc((size_t) nxp * nyp, sizeof(double));
    double *xly2y1 = (double *) calloc((size_t) nxp * nyp, sizeof(double));
    double *ylx2x1 = (double *) calloc((size_t) nxp * nyp, sizeof(double));

    const double e1 = (1.0 - nuv) / CO_PI, e2 = 1.0 / CO_PI, e3 = nuv / CO_PI;
    const double tolx = dx * 1e-10, tolx2 = dx * 1e-20;

    memset(cs->cf, 0, sizeof(double) * (size_t) 36 * cs->cf_mx * cs->cf_my);
    double *c11 = co_cf_ptr(cs, 1, 1), *c12 = co_cf_ptr(cs, 1, 2), *c13 = co_cf_ptr(cs, 1, 3);
    double *c21 = co_cf_ptr(cs, 2, 1), *c22 = co_cf_ptr(cs, 2, 2), *c23 = co_cf_ptr(cs, 2, 3);
    double *c31 = co_cf_ptr(cs, 3, 1), *c32 = co_cf_ptr(cs, 3, 2), *c33 = co_cf_ptr(cs, 3, 3);

    int iy0 = -my;
    if (my == 1) iy0 = -my + 1;                                           /* :474-475 */

    for (int iy = iy0; iy <= my; iy++)                                    /* :483-489 */
        for (int ix = -mx; ix <= mx; ix++) {
            P(x, ix, iy) = (double) ix * dx + xshft - 0.5 * dx;
            P(y, ix, iy) = (double) iy * dy + yshft - 0.5 * dy;
            P(r, ix, iy) = sqrt(P(x, ix, iy) * P(x, ix, iy) + P(y, ix, iy) * P(y, ix, iy));
        }

    for (int iy = iy0; iy <= my - 1; iy++)                                /* :495-505 */
        for (int ix = -mx; ix <= mx - 1; ix++) {
            double v = CO_CF(cs, c21, ix, iy);
            v = v - e3 * P(r, ix, iy);
            v = v + e3 * P(r, ix, iy + 1);
            v = v + e3 * P(r, ix + 1, iy);
            v = v - e3 * P(r, ix + 1, iy + 1);
            CO_CF(cs, c21, ix, iy) = v;
        }

    if (fabs(akv) >= 1e-6) {                                              /* :507-546 */
        double *alr  = (double *) calloc((size_t) nxp * nyp, sizeof(double));
        double *atxy = (double *) calloc((size_t) nxp * nyp, sizeof(double));
        double *atyx = (double *) calloc((size_t) nxp * nyp, sizeof(double));
        for (int iy = iy0; iy <= my; iy++)
            for (int ix = -mx; ix <= mx; ix++) {
                P(alr, ix, iy)  = log(P(r, ix, iy));
                P(atyx, ix, iy) = atan(P(y, ix, iy) / (P(x, ix, iy) + tolx2));
                P(atxy, ix, iy) = atan(P(x, ix, iy) / (P(y, ix, iy) + tolx2));
            }
        for (int iy = iy0; iy <= my - 1; iy++)
            for (int ix = -mx; ix <= mx - 1; ix++) {
#define J5(i, j) (P(y, i, j) * P(alr, i, j) + P(x, i, j) * P(atyx, i, j))
#define J6(i, j) (P(x, i, j) * P(alr, i, j) + P(y, i, j) * P(atxy, i, j))
                double v = CO_CF(cs, c13, ix, iy);
                v = v - e2 * akv * J5(ix, iy);
                v = v + e2 * akv * J5(ix, iy + 1);
                v = v + e2 * akv * J5(ix + 1, iy);
                v = v - e2 * akv * J5(ix + 1, iy + 1);
                CO_CF(cs, c13, ix, iy) = v;
                v = CO_CF(cs, c23, ix, iy);
                v = v - e2 * akv * J6(ix, iy);
                v = v + e2 * akv * J6(ix, iy + 1);
                v = v + e2 * akv * J6(ix + 1, iy);
                v = v - e2 * akv * J6(ix + 1, iy + 1);
                CO_CF(cs, c23, ix, iy) = v;
#undef J5
#undef J6
            }
        free(alr); free(atxy); free(atyx);
    }

    for (int iy = iy0; iy <= my - 1; iy++)                                /* :550-558 */
        for (int ix = -mx; ix <= mx; ix++) {
            double a = fabs(P(y, ix, iy) + P(r, ix, iy)), b = fabs(P(y, ix, iy + 1) + P(r, ix, iy + 1));
            if ((a < b ? a : b) < tolx)
                P(xly2y1, ix, iy) = 0.0;
            else
                P(xly2y1, ix, iy) = P(x, ix, iy) * log((P(y, ix, iy + 1) + P(r, ix, iy + 1)) / (P(y, ix, iy) + P(r, ix, iy)));
        }

    for (int iy = iy0; iy <= my; iy++)                                    /* :560-568 */
        for (int ix = -mx; ix <= mx - 1; ix++) {
            double a = fabs(P(x, ix, iy) + P(r, ix, iy)), b = fabs(P(x, ix + 1, iy) + P(r, ix + 1, iy));
            if ((a < b ? a : b) < tolx)
                P(ylx2x1, ix, iy) = 0.0;
            else
                P(ylx2x1, ix, iy) = P(y, ix, iy) * log((P(x, ix + 1, iy) + P(r, ix + 1, iy)) / (P(x, ix, iy) + P(r, ix, iy)));
        }

    for (int iy = iy0; iy <= my - 1; iy++)                                /* :570-590 */
        for (int ix = -mx; ix <= mx - 1; ix++) {
            double v;
            v = CO_CF(cs, c33, ix, iy);
            v = v - e1 * (P(xly2y1, ix, iy) + P(ylx2x1, ix, iy));
            v = v + e1 * (P(xly2y1, ix + 1, iy) + P(ylx2x1, ix, iy + 1));
            CO_CF(cs, c33, ix, iy) = v;
            v = CO_CF(cs, c11, ix, iy);
            v = v - (e1 * P(xly2y1, ix, iy) + e2 * P(ylx2x1, ix, iy));
            v = v + (e1 * P(xly2y1, ix + 1, iy) + e2 * P(ylx2x1, ix, iy + 1));
            CO_CF(cs, c11, ix, iy) = v;
            v = CO_CF(cs, c22, ix, iy);
            v = v - (e1 * P(ylx2x1, ix, iy) + e2 * P(xly2y1, ix, iy));
            v = v + (e1 * P(ylx2x1, ix, iy + 1) + e2 * P(xly2y1, ix + 1, iy));
            CO_CF(cs, c22, ix, iy) = v;
        }

    for (int iy = iy0; iy <= my - 1; iy++)                                /* :594-600 */
        for (int ix = -mx; ix <= mx - 1; ix++) {
            CO_CF(cs, c31, ix, iy) = -CO_CF(cs, c13, ix, iy);
            CO_CF(cs, c32, ix, iy) = -CO_CF(cs, c23, ix, iy);
            CO_CF(cs, c12, ix, iy) =  CO_CF(cs, c21, ix, iy);
        }
#undef P
    free(r); free(x); free(y); free(xly2y1); free(ylx2x1);
}

/* m_visc.f90:127-376 for M=0 (elastic), C=2 (piecewise constant), no reuse logic (always recompute). */
void co_sgencr(const co_mater *m, int mx, int my, double dx, double dy, int is_roll, double chi, double dq,
               co_inflcf *cs, co_inflcf *cv, co_inflcf *csv, co_inflcf *ms)
{
    co_inflcf *all[4] = { cs, cv, csv, ms };
    for (int i = 0; i < 4; i++) {
        if (!all[i]) continue;
        co_inflcf_free(all[i]);
        co_inflcf_init(all[i], mx, my, dx, dy);
        co_inflcf_mater(all[i], m);
    }
    const double akv = m->ak, nuv = m->nu;           /* mater_set_visc for M=0: akv=ak, nuv=nu */
    double cc = 0.0, sc = 0.0;
    if (is_roll) { cc = cos(chi); sc = sin(chi); }
    if (fabs(akv) < 1e-6) {                           /* :239-243 */
        cs->nt_cpl = 0;
        if (cv) cv->nt_cpl = 0;
        if (csv) csv->nt_cpl = 0;
    }
    const double facdq = 1.0, facdqi = 1.0 / facdq;  /* use_dq_scaling = .false. :140,247-264 */
    if (cv) cv->dq = dq;
    if (csv) csv->dq = dq;

    co_elascf_pcwcns(akv, nuv, mx, my, dx, dy, 0.0, 0.0, cs);
    if (!cv || !csv) return;

    const size_t ntot = (size_t) 36 * mx * my, nblk = (size_t) 4 * mx * my;
    if (!is_roll) {                                   /* :294-308 */
        memcpy(cv->cf, cs->cf, sizeof(double) * ntot);
        memcpy(csv->cf, cs->cf, sizeof(double) * ntot);
    } else {                                          /* :310-358 */
        const double xshft = cc * facdq * dq, yshft = sc * facdq * dq;
        co_elascf_pcwcns(akv, nuv, mx, my, dx, dy, xshft, yshft, cv);
        for (int jk = 1; jk <= 3; jk++) {
            double *s1 = co_cf_ptr(cs, 1, jk), *s2 = co_cf_ptr(cs, 2, jk), *s3 = co_cf_ptr(cs, 3, jk);
            double *v1 = co_cf_ptr(cv, 1, jk), *v2 = co_cf_ptr(cv, 2, jk), *v3 = co_cf_ptr(cv, 3, jk);
            double *d1 = co_cf_ptr(csv, 1, jk), *d2 = co_cf_ptr(csv, 2, jk), *d3 = co_cf_ptr(csv, 3, jk);
            for (size_t i = 0; i < nblk; i++) {
                v3[i] = s3[i];
                d1[i] = facdqi * (s1[i] - v1[i]);
                d2[i] = facdqi * (s2[i] - v2[i]);
                d3[i] = s3[i];
                v1[i] = s1[i] - d1[i];
                v2[i] = s2[i] - d2[i];
            }
        }
    }
}
