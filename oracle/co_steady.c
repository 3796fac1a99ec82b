/*
 * co_steady.c -- ORACLE (test infrastructure, not product code).
 * Steady rolling (T=3) with the SteadyGS solver: per-element constrained 2x2 solve (plstrc, elastic branches),
 * Gauss-Seidel sweep on the traction differences dp with re-integration along the rows, leading-edge factors.
 * Follows /root/reference/src/m_solvpt.f90:2825-3254 (stdygs), :3278-3807 (plstrc; the plastic branch is reached with
 * taucs = 1e20 and only cycles the state, see SURVEY.md 7 "quirks"), /root/reference/src/m_leadedge.f90:92-332 (sxbnd),
 * :336-394 (subnd), and the ConvexGS solver /root/reference/src/m_solvpt.f90:2446-2821 (cnvxgs).
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* m_solvpt.f90:3278-3807 with use_plast = .false. (tau_c0 = taucv = 1e20, k_tau = 0).
 * el in/out, pr[3] in/out (pr[2] = pn), s[2] in/out (slip). */
void co_plstrc(int *el, const double coef[2][2], double eps, double omegah, double omegas, double pr[3], double mus,
               double s[2])
{
    const int maxnr = 10;
    const double prc = 0.0001, taucs = 1e20;
    const double pold1 = pr[0], pold2 = pr[1];
    double si0[2], si[2] = { 0.0, 0.0 }, dp[2], fk[2], gradf[2][2];
    si0[0] = s[0] - coef[0][0] * pr[0] - coef[0][1] * pr[1];
    si0[1] = s[1] - coef[1][0] * pr[0] - coef[1][1] * pr[1];
    int itry = 0, violat = 1;
    while (itry <= 3 && violat == 1) {
        itry++;
        const double trc_bound = mus * pr[2];
        if (*el == CO_ADHES) {
            const double detinv = 1.0 / (coef[0][0] * coef[1][1] - coef[0][1] * coef[1][0]);
            dp[0] = (-coef[1][1] * s[0] + coef[0][1] * s[1]) * detinv;
            dp[1] = (coef[1][0] * s[0] - coef[0][0] * s[1]) * detinv;
            pr[0] = pold1 + omegah * dp[0];
            pr[1] = pold2 + omegah * dp[1];
            const double pabs2 = sqrt(pr[0] * pr[0] + pr[1] * pr[1]);
            const double bnd = fmin(trc_bound, taucs);
            if (pabs2 <= bnd) violat = 0;
            else {
                const double pabs1 = sqrt((pold1 + dp[0]) * (pold1 + dp[0]) + (pold2 + dp[1]) * (pold2 + dp[1]));
                if (pabs1 <= bnd) { pr[0] = pr[0] * bnd / pabs2; pr[1] = pr[1] * bnd / pabs2; violat = 0; }
                else violat = 1;
            }
            si[0] = si0[0] + coef[0][0] * pr[0] + coef[0][1] * pr[1];
            si[1] = si0[1] + coef[1][0] * pr[0] + coef[1][1] * pr[1];
        } else if (*el == CO_SLIP) {
            violat = 0;
            const double sabs = sqrt(si0[0] * si0[0] + si0[1] * si0[1]);
            pr[0] = -trc_bound * si0[0] / sabs;
            pr[1] = -trc_bound * si0[1] / sabs;
            si[0] = si0[0] + coef[0][0] * pr[0] + coef[0][1] * pr[1];
            si[1] = si0[1] + coef[1][0] * pr[0] + coef[1][1] * pr[1];
            fk[0] = pr[1] * si[0] - pr[0] * si[1];
            fk[1] = pr[0] * pr[0] + pr[1] * pr[1] - trc_bound * trc_bound;
            int itnr = 0;
            while (itnr == 0 || (itnr < maxnr && fabs(fk[0]) + fabs(fk[1]) >= prc * eps * trc_bound)) {
                itnr++;
                gradf[0][0] = -2.0 * coef[1][0] * pr[0] - si0[1] + (coef[0][0] - coef[1][1]) * pr[1];
                gradf[0][1] = 2.0 * coef[0][1] * pr[1] + si0[0] + (coef[0][0] - coef[1][1]) * pr[0];
                gradf[1][0] = 2.0 * pr[0];
                gradf[1][1] = 2.0 * pr[1];
                const double det = gradf[0][0] * gradf[1][1] - gradf[0][1] * gradf[1][0];
                if (det == 0.0) itnr = maxnr;
                else {
                    dp[0] = -(gradf[1][1] * fk[0] - gradf[0][1] * fk[1]) / det;
                    dp[1] = -(-gradf[1][0] * fk[0] + gradf[0][0] * fk[1]) / det;
                    pr[0] += dp[0]; pr[1] += dp[1];
                }
                si[0] = si0[0] + coef[0][0] * pr[0] + coef[0][1] * pr[1];
                si[1] = si0[1] + coef[1][0] * pr[0] + coef[1][1] * pr[1];
                fk[0] = pr[1] * si[0] - pr[0] * si[1];
                fk[1] = pr[0] * pr[0] + pr[1] * pr[1] - trc_bound * trc_bound;
            }
            {   /* relaxation of the traction direction, :3612-3619 */
                const double alph0 = atan2(pold2, pold1);
                double alph1 = atan2(pr[1], pr[0]), dalph = alph1 - alph0;
                if (dalph < -CO_PI) dalph += 2.0 * CO_PI;
                if (dalph > CO_PI) dalph -= 2.0 * CO_PI;
                alph1 = alph0 + omegas * dalph;
                pr[0] = trc_bound * cos(alph1);
                pr[1] = trc_bound * sin(alph1);
            }
            if (fabs(pr[0]) > fabs(pr[1])) { if (pr[0] * si[0] > 0.0) violat = 1; }
            else { if (pr[1] * si[1] > 0.0) violat = 1; }
        } else {
            /* plastic branch with taucs = 1e20 >= trc_bound: always "violated"; its pr/si are rebuilt by the next try */
            si[0] = 0.0; si[1] = 0.0;
            violat = 1;
        }
        if (itry <= 3 && violat == 1) {
            if (*el == CO_ADHES) *el = CO_SLIP;            /* k_tau = 0: trc_bound <= taucs */
            else if (*el == CO_SLIP) *el = CO_PLAST;
            else *el = CO_ADHES;
        }
    }
    s[0] = si[0]; s[1] = si[1];
}

/* m_solvpt.f90:2825-3254, chi = 0 (rolling in +x: ixsta = 1 trailing edge, ixend = mx leading edge), elastic */
void co_stdygs(co_ctx *cx, int mx, int my, const double *ws, co_inflcf *cs, const double *mus, co_eldiv *igs, double *ps,
               double *ss, int k, double eps, int maxgs, double omegah, double omegas, int *info, int *itgs_out, double *err)
{
    const int npot = mx * my, ixsta = 1, ixinc = 1, ixend = mx;
    double *dp = (double *) calloc(3L * npot, sizeof(double));
    double *psx = ps, *psy = ps + npot, *psn = ps + 2L * npot;
    const double *c11 = co_cf_ptr(cs, 1, 1), *c12 = co_cf_ptr(cs, 1, 2), *c21 = co_cf_ptr(cs, 2, 1), *c22 = co_cf_ptr(cs, 2, 2);
    for (int iy = 1; iy <= my; iy++)
        for (int ix = 1; ix <= mx; ix++) {
            const int ii = ix + (iy - 1) * mx - 1;
            if (ix != ixend) { dp[ii] = psx[ii] - psx[ii + ixinc]; dp[npot + ii] = psy[ii] - psy[ii + ixinc]; }
            else { dp[ii] = psx[ii]; dp[npot + ii] = psy[ii]; }
        }
    int nadh = 0, nslip = 0, nplst = 0;
    for (int i = 0; i < npot; i++) { if (igs->el[i] == CO_ADHES) nadh++; else if (igs->el[i] == CO_SLIP) nslip++; else if (igs->el[i] == CO_PLAST) nplst++; }
    const double facnel = (double) sqrtf((float) npot / (float) (nadh + nslip + nplst));
    int itgs = 0;
    double dif = 2.0, difid = 1.0, dif1 = 0.0, conv;
    while (dif >= difid && itgs < maxgs) {
        itgs++;
        dif = 0.0;
        for (int iy = 1; iy <= my; iy++) {
            int ix = ixsta - ixinc;
            while (ix != ixend) {
                ix += ixinc;
                const int ii = ix + (iy - 1) * mx - 1;
                if (igs->el[ii] >= CO_ADHES) {
                    int jx = ix - ixinc, jj = jx + (iy - 1) * mx - 1;
                    while (igs->el[jj] == CO_ADHES && jx != ixsta) { jx -= ixinc; jj = jx + (iy - 1) * mx - 1; }
                    double coef[2][2], s[2], pr[3];
                    coef[0][0] = cs->ga_inv * (CO_CF(cs, c11, 0, 0) - CO_CF(cs, c11, jx - ix, 0));
                    coef[0][1] = cs->ga_inv * (CO_CF(cs, c12, 0, 0) - CO_CF(cs, c12, jx - ix, 0));
                    coef[1][0] = cs->ga_inv * (CO_CF(cs, c21, 0, 0) - CO_CF(cs, c21, jx - ix, 0));
                    coef[1][1] = cs->ga_inv * (CO_CF(cs, c22, 0, 0) - CO_CF(cs, c22, jx - ix, 0));
                    s[0] = ws[ii] + co_aijpj(ii + 1, CO_X, dp, igs, CO_TANG, cs);
                    s[1] = ws[npot + ii] + co_aijpj(ii + 1, CO_Y, dp, igs, CO_TANG, cs);
                    cx->st.n_rowsum += 2;
                    pr[0] = psx[ii]; pr[1] = psy[ii]; pr[2] = psn[ii];
                    co_plstrc(&igs->el[ii], coef, eps, omegah, omegas, pr, mus[ii], s);
                    dif = dif + (pr[0] - psx[ii]) * (pr[0] - psx[ii]) + (pr[1] - psy[ii]) * (pr[1] - psy[ii]);
                    dp[ii] = dp[ii] + pr[0] - psx[ii];
                    dp[npot + ii] = dp[npot + ii] + pr[1] - psy[ii];
                    psx[ii] = pr[0]; psy[ii] = pr[1];
                    ss[ii] = s[0]; ss[npot + ii] = s[1];
                    for (jx = ix - ixinc; jx >= ixsta; jx -= ixinc) {             /* re-integrate dp -> ps, :3089-3126 */
                        jj = jx + (iy - 1) * mx - 1;
                        if (igs->el[jj] == CO_ADHES) {
                            psx[jj] = psx[jj + ixinc] + dp[jj];
                            psy[jj] = psy[jj + ixinc] + dp[npot + jj];
                            const double ptabs = sqrt(psx[jj] * psx[jj] + psy[jj] * psy[jj]);
                            const double ptbnd = fmin(mus[jj] * psn[jj], 1e20);
                            if (ptabs > ptbnd) {
                                psx[jj] = psx[jj] * ptbnd / ptabs; psy[jj] = psy[jj] * ptbnd / ptabs;
                                dp[jj] = psx[jj] - psx[jj + ixinc]; dp[npot + jj] = psy[jj] - psy[jj + ixinc];
                            }
                        } else if (igs->el[jj] == CO_SLIP || igs->el[jj] == CO_PLAST) {
                            dp[jj] = psx[jj] - psx[jj + ixinc]; dp[npot + jj] = psy[jj] - psy[jj + ixinc];
                        } else if (igs->el[jj] <= CO_EXTER && igs->el[jj + ixinc] >= CO_ADHES) {
                            dp[jj] = -psx[jj + ixinc]; dp[npot + jj] = -psy[jj + ixinc];
                        }
                    }
                }
            }
        }
        dif = sqrt(dif / (2.0 * k));
        {
            double sq = 0.0;
            for (int i = 0; i < 2 * npot; i++) sq += ps[i] * ps[i];
            difid = eps * fmax(1e-6, facnel * sqrt(sq / (2.0 * npot)));
        }
        if (itgs == 1) dif1 = dif;
    }
    conv = 1.0;
    if (dif * dif1 != 0.0 && itgs > 1) conv = exp(log(dif / dif1) / (itgs - 1));
    *err = dif;
    *info = 0;
    if (itgs >= maxgs) *info = 1;
    if (itgs >= maxgs && conv > 0.997) *info = 2;
    if (itgs >= maxgs && conv > 1.0) *info = 3;
    *itgs_out = itgs;
    free(dp);
}

/* m_leadedge.f90:92-332 for chi = 0 and solvers without the leading-edge correction (SteadyGS: fxdfac = 2):
 * facdt(ii) = 0 in the exterior, 1 near the end of the grid, min(1, (xbnd - x)/dq) otherwise */
void co_sxbnd_facdt(int mx, int my, const co_eldiv *igs, const double *x, double dx, double dq, double *facdt)
{
    for (int iy = 1; iy <= my; iy++) {
        /* positions ixbnd of the transitions C -> E in this row, ascending */
        int nb = 0, *ixb = (int *) malloc(sizeof(int) * (mx + 1));
        for (int ix = 1; ix <= mx - 1; ix++) {
            const int ii = ix + (iy - 1) * mx - 1;
            if (igs->el[ii] >= CO_ADHES && igs->el[ii + 1] <= CO_EXTER) ixb[nb++] = ix;
        }
        if (igs->el[mx + (iy - 1) * mx - 1] >= CO_ADHES) ixb[nb++] = mx;
        int j = 0;
        for (int ix = 1; ix <= mx; ix++) {
            while (j < nb && ix > ixb[j]) j++;
            const int ii = ix + (iy - 1) * mx - 1;
            if (igs->el[ii] <= CO_EXTER) facdt[ii] = 0.0;
            else if (ix + 2 > mx) facdt[ii] = 1.0;
            else {
                const double xbnd = x[ixb[j] + (iy - 1) * mx - 1] + 2.0 * dx;
                facdt[ii] = fmin(1.0, (xbnd - x[ii]) / dq);
            }
        }
        free(ixb);
    }
}

/* m_leadedge.f90:92-332 for chi = 0: leading-edge positions per row (jbnd, ixbnd), facdx = fxdfac (2 without, 1 with the
 * leading-edge correction), xbnd, facdt and ii2j (0: interior equation, j > 0: leading-edge equation at position j) */
void co_sxbnd(int mx, int my, int is_roll, int use_ledg, const co_eldiv *igs, const double *x, double dx, double dq,
              co_leadedge *lg)
{
    const int npot = mx * my;
    lg->jbnd = (int *) realloc(lg->jbnd, sizeof(int) * (my + 2));
    lg->jbnd[1] = 1;
    for (int iy = 1; iy <= my; iy++) {
        int np = 0;
        for (int ix = 1; ix <= mx - 1; ix++) {
            const int ii = ix + (iy - 1) * mx - 1;
            if (igs->el[ii] >= CO_ADHES && igs->el[ii + 1] <= CO_EXTER) np++;
        }
        if (igs->el[mx + (iy - 1) * mx - 1] >= CO_ADHES) np++;
        lg->jbnd[iy + 1] = lg->jbnd[iy] + np;
    }
    lg->npos = lg->jbnd[my + 1];
    lg->ixbnd = (int *) realloc(lg->ixbnd, sizeof(int) * (lg->npos + 1));
    lg->xbnd = (double *) realloc(lg->xbnd, sizeof(double) * (lg->npos + 1));
    lg->facdx = (double *) realloc(lg->facdx, sizeof(double) * (lg->npos + 1));
    lg->ubnd = (double *) realloc(lg->ubnd, sizeof(double) * 2 * (lg->npos + 1));
    lg->ii2j = (int *) realloc(lg->ii2j, sizeof(int) * npot);
    lg->facdt = (double *) realloc(lg->facdt, sizeof(double) * npot);
    for (int iy = 1; iy <= my; iy++) {
        int j = lg->jbnd[iy];
        for (int ix = 1; ix <= mx - 1; ix++) {
            const int ii = ix + (iy - 1) * mx - 1;
            if (igs->el[ii] >= CO_ADHES && igs->el[ii + 1] <= CO_EXTER) lg->ixbnd[j++] = ix;
        }
        if (igs->el[mx + (iy - 1) * mx - 1] >= CO_ADHES) lg->ixbnd[j++] = mx;
    }
    const double fxdfac = use_ledg ? 1.0 : 2.0;
    for (int iy = 1; iy <= my; iy++)
        for (int j = lg->jbnd[iy]; j < lg->jbnd[iy + 1]; j++) {
            lg->facdx[j] = fxdfac;
            lg->xbnd[j] = x[lg->ixbnd[j] + (iy - 1) * mx - 1] + 1 * lg->facdx[j] * dx;
        }
    for (int iy = 1; iy <= my; iy++) {
        int j = lg->jbnd[iy];
        for (int ix = 1; ix <= mx; ix++) {
            while (j < lg->jbnd[iy + 1] && ix > lg->ixbnd[j]) j++;
            const int ii = ix + (iy - 1) * mx - 1;
            if (!is_roll) { lg->ii2j[ii] = 0; lg->facdt[ii] = 1.0; }
            else if (igs->el[ii] <= CO_EXTER) { lg->ii2j[ii] = 0; lg->facdt[ii] = 0.0; }
            else if (ix + 2 > mx) { lg->ii2j[ii] = 0; lg->facdt[ii] = 1.0; }
            else {
                lg->facdt[ii] = fmin(1.0, 1 * (lg->xbnd[j] - x[ii]) / dq);
                lg->ii2j[ii] = (lg->facdt[ii] < 0.9999) ? j : 0;
            }
        }
    }
}

void co_leadedge_free(co_leadedge *lg)
{
    free(lg->jbnd); free(lg->ixbnd); free(lg->xbnd); free(lg->facdx); free(lg->ubnd); free(lg->ii2j); free(lg->facdt);
    memset(lg, 0, sizeof(*lg));
}

/* m_leadedge.f90:336-394: displacement difference at the leading-edge positions, interpolated between the last interior
 * element and its exterior neighbour (all three traction directions) */
void co_subnd(co_ctx *cx, int mx, int my, const double *p, const co_eldiv *pel, const co_inflcf *c, co_leadedge *lg)
{
    for (int iy = 1; iy <= my; iy++)
        for (int j = lg->jbnd[iy]; j < lg->jbnd[iy + 1]; j++) {
            const int ix0 = lg->ixbnd[j];
            if (ix0 + 2 > mx) { lg->ubnd[2 * j] = 0.0; lg->ubnd[2 * j + 1] = 0.0; continue; }
            const int ii = ix0 + (iy - 1) * mx;
            for (int ik = 1; ik <= 2; ik++) {
                const double u0 = co_aijpj(ii, ik, p, pel, CO_ALL, c), u1 = co_aijpj(ii + 1, ik, p, pel, CO_ALL, c);
                cx->st.n_rowsum += 2;
                lg->ubnd[2 * j + ik - 1] = (1.0 - lg->facdx[j]) * u0 + lg->facdx[j] * u1;
            }
        }
}

/* m_solvpt.f90:2446-2821, chi = 0, elastic: block Gauss-Seidel with the per-element constrained solve; shifts use cs,
 * steady rolling csv (= cs - cv) in the interior and cs - ubnd at the leading-edge elements (ii2j > 0) */
void co_cnvxgs(co_ctx *cx, int mx, int my, int is_ssrol, const double *ws, co_inflcf *cs, co_inflcf *csv, co_leadedge *lg,
               const double *mus, co_eldiv *igs, double *ps, double *ss, int k, const int *iel, double eps, int maxgs,
               double omegah, double omegas, int *info, int *itgs_out, double *err)
{
    const int npot = mx * my;
    double coefs[2][2], coefsv[2][2];
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
            coefs[a][b] = cs->ga_inv * CO_CF(cs, co_cf_ptr(cs, a + 1, b + 1), 0, 0);
            coefsv[a][b] = csv->ga_inv * CO_CF(csv, co_cf_ptr(csv, a + 1, b + 1), 0, 0);
        }
    int nadh = 0, nslip = 0, nplst = 0;
    for (int i = 0; i < npot; i++) { if (igs->el[i] == CO_ADHES) nadh++; else if (igs->el[i] == CO_SLIP) nslip++; else if (igs->el[i] == CO_PLAST) nplst++; }
    const double facnel = (double) sqrtf((float) npot / (float) (nadh + nslip + nplst));
    int itgs = 0;
    double dif = 2.0, difid = 1.0, dif1 = 0.0;
    while (dif >= difid && itgs < maxgs) {
        itgs++;
        dif = 0.0;
        if (is_ssrol) co_subnd(cx, mx, my, ps, igs, cs, lg);
        for (int i = 0; i < k; i++) {
            const int ii = iel[i];                         /* 0-based element number */
            if (itgs <= 1000 || itgs % 2 == 0 || igs->el[ii] == CO_ADHES) {
                double pr[3] = { ps[ii], ps[npot + ii], ps[2L * npot + ii] }, s[2];
                const int zledge = is_ssrol && lg->ii2j[ii] > 0;
                if (!is_ssrol) {
                    s[0] = ws[ii] + co_aijpj(ii + 1, CO_X, ps, igs, CO_TANG, cs);
                    s[1] = ws[npot + ii] + co_aijpj(ii + 1, CO_Y, ps, igs, CO_TANG, cs);
                } else if (!zledge) {
                    s[0] = ws[ii] + co_aijpj(ii + 1, CO_X, ps, igs, CO_TANG, csv);
                    s[1] = ws[npot + ii] + co_aijpj(ii + 1, CO_Y, ps, igs, CO_TANG, csv);
                } else {
                    s[0] = ws[ii] + co_aijpj(ii + 1, CO_X, ps, igs, CO_TANG, cs) - lg->ubnd[2 * lg->ii2j[ii]];
                    s[1] = ws[npot + ii] + co_aijpj(ii + 1, CO_Y, ps, igs, CO_TANG, cs) - lg->ubnd[2 * lg->ii2j[ii] + 1];
                }
                cx->st.n_rowsum += 2;
                co_plstrc(&igs->el[ii], (!is_ssrol || zledge) ? coefs : coefsv, eps, omegah, omegas, pr, mus[ii], s);
                dif = dif + (ps[ii] - pr[0]) * (ps[ii] - pr[0]) + (ps[npot + ii] - pr[1]) * (ps[npot + ii] - pr[1]);
                ps[ii] = pr[0]; ps[npot + ii] = pr[1];
                ss[ii] = s[0]; ss[npot + ii] = s[1];
            }
        }
        dif = sqrt(dif / (2.0 * k));
        {
            double sq = 0.0;
            for (int i = 0; i < 2 * npot; i++) sq += ps[i] * ps[i];
            difid = eps * fmax(1e-6, facnel * sqrt(sq / (2.0 * npot)));
        }
        if (itgs == 1) dif1 = dif;
    }
    double conv = 1.0;
    if (dif * dif1 != 0.0 && itgs > 1) conv = exp(log(dif / dif1) / (itgs - 1));
    *err = dif;
    *info = 0;
    if (itgs >= maxgs) *info = 1;
    if (itgs >= maxgs && conv > 0.997) *info = 2;
    if (itgs >= maxgs && conv > 1.0) *info = 3;
    *itgs_out = itgs;
}
