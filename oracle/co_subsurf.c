/*
 * co_subsurf.c -- ORACLE (test infrastructure, not product code).
 * Subsurface displacements and stresses in the elastic half-space.  Follows /root/reference/src/m_subsurf.f90:
 *   stres1_pcwcns :1633-1852 (Kalker 1986, "Numerical calculation of the elastic field in a half-space"),
 *   sstres_inflcf :1261-1408, sstres_fft :1097-1257, sstres :1412-1515, sstres_derived :1519-1629.
 * The principal stresses use the trigonometric solution of the characteristic cubic of a symmetric 3x3 matrix,
 * derived here from scratch (the reference's dsyevc3 :2439-2498 is LGPL-derived and is not restated).
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* v[i][j][k]: k = 0 displacement u_j, k = 1..3 gradient u_{j,k}, due to unit load in direction i (0-based i, j) */
void co_stres1_pcwcns(double dx, double dy, double gg, double v[3][3][4], double vnu[3][3][4], const double xw[3],
                      const double xp[2])
{
    const double epsrel = 5e-7, pi = 4.0 * atan(1.0);
    static const int sgn[3] = { -1, -1, 1 }, ip[3] = { 1, 2, 0 };
    double y[3], yeps[3], w, weps, epsy, al[3], at[3], wm[4], a[3][3][4], t[3][3][4], q;
    memset(v, 0, sizeof(double) * 36);
    memset(vnu, 0, sizeof(double) * 36);
    for (int jx = -1; jx <= 1; jx += 2)
        for (int jy = -1; jy <= 1; jy += 2) {
            y[0] = xp[0] + jx * dx / 2.0 - xw[0];
            y[1] = xp[1] + jy * dy / 2.0 - xw[1];
            y[2] = xw[2];
            w = fmax(1e-12, sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]));
            epsy = epsrel * w;
            for (int i = 0; i < 3; i++) yeps[i] = (y[i] >= 0.0) ? fmax(epsy, y[i]) : fmin(-epsy, y[i]);
            weps = fmax(1e-12, sqrt(yeps[0] * yeps[0] + yeps[1] * yeps[1] + yeps[2] * yeps[2]));
            wm[0] = weps;
            for (int k = 0; k < 3; k++) {
                al[k] = log(y[k] + weps);
                at[k] = atan((y[ip[k]] + y[ip[ip[k]]] + w) / yeps[k]);
                wm[k + 1] = sgn[k] * y[k] / weps;
            }
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    a[i][j][0] = y[i] * al[j];
                    t[i][j][0] = y[i] * at[j];
                    for (int k = 0; k < 3; k++) {
                        const double kik = (i == k), kjk = (j == k);
                        a[i][j][k + 1] = sgn[k] * (kik * al[j] + y[i] * (kjk * w + y[k]) / (weps * (y[j] + weps)));
                        t[i][j][k + 1] = sgn[k] * (kik * at[j] + y[i] * (y[j] * (y[k] + w) - kjk * w * (y[0] + y[1] + y[2] + w))
                                                                  / (2 * weps * (y[ip[j]] + weps) * (y[ip[ip[j]]] + weps)));
                    }
                }
            const double s = (double) (jx * jy);
            for (int k = 0; k < 4; k++) {
                for (int i = 0; i < 2; i++) {
                    const int l = 1 - i;
                    q = a[i][l][k] + 2 * a[l][i][k] + 4 * t[2][2][k] + a[i][l][k] - 2 * t[2][i][k];
                    v[i][i][k] += s * q;
                    q = -2 * (a[i][l][k] - 2 * t[2][i][k]);
                    vnu[i][i][k] += q * s;
                    q = -a[2][2][k];
                    v[i][l][k] += s * q;
                    q = 2 * (a[2][2][k] - wm[k]);
                    vnu[i][l][k] += q * s;
                    q = 2 * (a[l][2][k] + a[2][l][k] + 2 * t[i][i][k]);
                    vnu[i][2][k] += q * s;
                    vnu[2][i][k] = -vnu[i][2][k];
                    q = -a[l][2][k] - 2 * t[i][i][k];
                    v[i][2][k] += q * s;
                    v[2][i][k] += (2 * a[2][l][k] - q) * s;
                }
                q = -2 * (a[0][1][k] + a[1][0][k] + 2 * t[2][2][k]);
                vnu[2][2][k] += q * s;
                q = -2 * t[2][2][k] + 2 * (a[0][1][k] + a[1][0][k] + 2 * t[2][2][k]);
                v[2][2][k] += q * s;
            }
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 4; k++) { v[i][j][k] /= (4 * pi * gg); vnu[i][j][k] /= (4 * pi * gg); }
}

/* eigenvalues of a symmetric 3x3 matrix, sorted descending: trigonometric root formula of the depressed cubic.
 * With q = tr(A)/3, p = sqrt(tr((A-qI)^2)/6), B = (A-qI)/p: det(B)/2 = cos(3 phi), eigenvalues q + 2 p cos(phi + 2 pi k/3). */
void co_sym3_eigenvalues(const double s[3][3], double ev[3])
{
    const double q = (s[0][0] + s[1][1] + s[2][2]) / 3.0;
    const double p1 = s[0][1] * s[0][1] + s[0][2] * s[0][2] + s[1][2] * s[1][2];
    const double d0 = s[0][0] - q, d1 = s[1][1] - q, d2 = s[2][2] - q;
    const double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
    if (p2 <= 0.0) { ev[0] = ev[1] = ev[2] = q; return; }
    const double p = sqrt(p2 / 6.0);
    const double b00 = d0 / p, b11 = d1 / p, b22 = d2 / p, b01 = s[0][1] / p, b02 = s[0][2] / p, b12 = s[1][2] / p;
    double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
    if (r < -1.0) r = -1.0;
    if (r > 1.0) r = 1.0;
    const double phi = acos(r) / 3.0;
    ev[0] = q + 2.0 * p * cos(phi);
    ev[2] = q + 2.0 * p * cos(phi + 2.0 * CO_PI / 3.0);
    ev[1] = 3.0 * q - ev[0] - ev[2];
}

/* m_subsurf.f90:1519-1629. vr[j][k] (0-based j, k = 0..3) is modified (sign flips for the lower body).
 * out[18] = uw(3), sighyd, sigvm, sigtr, sigmaj(3), sigma(3,3) column-major  == table columns 4..21 */
void co_sstres_derived(double gg, double poiss, int neg, double vr[3][4], double out[18])
{
    const double tolsml = 1e-15;
    double er[3][3], sigma[3][3], sigmaj[3], uw[3];
    for (int j = 0; j < 3; j++) {
        const int ifacj = (j >= 1) ? neg : 1;
        for (int k = 0; k < 4; k++) {
            const int ifack = (k >= 2) ? neg : 1;
            vr[j][k] = ifacj * ifack * vr[j][k];
        }
    }
    for (int i = 0; i < 3; i++) uw[i] = vr[i][0];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) er[i][j] = (vr[i][j + 1] + vr[j][i + 1]) / 2.0;
    const double dil = er[0][0] + er[1][1] + er[2][2];
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) sigma[i][j] = 2.0 * gg * er[i][j];
    for (int i = 0; i < 3; i++) sigma[i][i] = sigma[i][i] + 2.0 * gg * dil * poiss / fmax(1e-6, 1.0 - 2.0 * poiss);
    const double sigii = sigma[0][0] + sigma[1][1] + sigma[2][2];
    const double sighyd = sigii / 3.0;
    double sijsij = -(1.0 / 3.0) * sigii * sigii;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) sijsij = sijsij + sigma[i][j] * sigma[i][j];
    sijsij = 0.5 * sijsij;
    const double sigvm = sqrt(3.0 * sijsij);
    co_sym3_eigenvalues(sigma, sigmaj);
    const double sigtr = sigmaj[0] - sigmaj[2];
    for (int i = 0; i < 3; i++) if (fabs(uw[i]) < gg * tolsml) uw[i] = 0.0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) if (fabs(sigma[i][j]) < gg * tolsml) sigma[i][j] = 0.0;
    out[0] = uw[0]; out[1] = uw[1]; out[2] = uw[2];
    out[3] = sighyd; out[4] = sigvm; out[5] = sigtr;
    out[6] = sigmaj[0]; out[7] = sigmaj[1]; out[8] = sigmaj[2];
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) out[9 + j * 3 + i] = sigma[i][j];    /* reshape(sigma,(/9/)) */
}

/* m_subsurf.f90:1412-1515: direct evaluation in one point xw (ISUBS = 9). ps: [3][npot]. out[18] as above. */
void co_sstres_point(int mx, int my, double dx, double dy, const double *x, const double *y, const double gg[2],
                     const double poiss[2], const double *ps, const double xw_in[3], double out[18])
{
    const int npot = mx * my;
    double vr[3][4], v[3][3][4], vnu[3][3][4], xw[3] = { xw_in[0], xw_in[1], xw_in[2] }, xp[2];
    memset(vr, 0, sizeof(vr));
    int ia, neg;
    if (xw[2] >= 0) { ia = 0; neg = 1; } else { ia = 1; neg = -1; }
    xw[2] = neg * xw[2];
    xw[1] = neg * xw[1];
    for (int ii = 0; ii < npot; ii++) {
        if (ps[2L * npot + ii] > 0.0) {
            xp[0] = x[ii];
            xp[1] = neg * y[ii];
            co_stres1_pcwcns(dx, dy, gg[ia], v, vnu, xw, xp);
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 4; k++)
                    for (int i = 0; i < 3; i++) {
                        const int ifaci = (i == 0) ? neg : 1;
                        vr[j][k] = vr[j][k] + ifaci * ps[(long) i * npot + ii] * (v[i][j][k] + poiss[ia] * vnu[i][j][k]);
                    }
        }
    }
    co_sstres_derived(gg[ia], poiss[ia], neg, vr, out);
}

/* m_subsurf.f90:1261-1408 for ISUBS 1/5 blocks covering the whole grid (nx = mx, ny = my): ck[0..3] */
void co_sstres_inflcf(int mx, int my, double dx, double dy, const double gg[2], const double poiss[2], double zw,
                      co_inflcf ck[4])
{
    int ia, neg;
    if (zw >= 0.0) { ia = 0; neg = 1; } else { ia = 1; neg = -1; }
    const int nx = mx, ny = my;
    double v[3][3][4], vnu[3][3][4], xw[3], xp[2] = { 0.0, 0.0 };
    for (int k = 0; k < 4; k++) {
        co_inflcf_init(&ck[k], mx, my, dx, dy);
        ck[k].nt_cpl = 1; ck[k].ga = 1.0; ck[k].ga_inv = 1.0;
    }
    for (int iy = -ny; iy <= 0; iy++)
        for (int ix = -nx; ix <= 0; ix++) {
            xw[0] = (double) ix * dx;
            xw[1] = neg * (double) iy * dy;
            xw[2] = neg * zw;
            co_stres1_pcwcns(dx, dy, gg[ia], v, vnu, xw, xp);
            for (int k = 0; k < 4; k++)
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++)                       /* cf(ix,iy, j, i) = v(i,j,k) + nu vnu(i,j,k) */
                        CO_CF(&ck[k], co_cf_ptr(&ck[k], j + 1, i + 1), ix, iy) = v[i][j][k] + poiss[ia] * vnu[i][j][k];
        }
    for (int k = 0; k < 4; k++) {                                     /* mirror in x */
        const int ifack = (k == 1) ? -1 : 1;
        for (int ik = 1; ik <= 3; ik++) {
            const int ifaci = (ik == 1) ? -1 : 1;
            for (int jk = 1; jk <= 3; jk++) {
                const int ifacj = (jk == 1) ? -1 : 1;
                double *blk = co_cf_ptr(&ck[k], jk, ik);
                for (int iy = -ny; iy <= 0; iy++)
                    for (int ix = 1; ix <= nx - 1; ix++)
                        CO_CF(&ck[k], blk, ix, iy) = ifaci * ifacj * ifack * CO_CF(&ck[k], blk, -ix, iy);
            }
        }
    }
    for (int k = 0; k < 4; k++) {                                     /* mirror in y */
        const int ifack = (k == 2) ? -1 : 1;
        for (int ik = 1; ik <= 3; ik++) {
            const int ifaci = (ik == 2) ? -1 : 1;
            for (int jk = 1; jk <= 3; jk++) {
                const int ifacj = (jk == 2) ? -1 : 1;
                double *blk = co_cf_ptr(&ck[k], jk, ik);
                for (int iy = 1; iy <= ny - 1; iy++)
                    for (int ix = -nx; ix <= nx - 1; ix++)
                        CO_CF(&ck[k], blk, ix, iy) = ifaci * ifacj * ifack * CO_CF(&ck[k], blk, ix, -iy);
            }
        }
    }
}

/* m_subsurf.f90:1097-1257 for ISUBS = 1/5 (all elements of the potential contact, nz depths).
 * ps: [3][npot] (NOT modified: the reference flips px in place for z < 0 and never restores it, :1170; the oracle
 * works on a copy).  table: [nz*npot][18] (columns 4..21 of the reference's table). use_fft: 1 = VecAijPj, 0 = AijPj sums */
void co_subsurf_block_fft(co_ctx *cx, int mx, int my, double dx, double dy, const double gg[2], const double poiss[2],
                          const int *el, const double *ps_in, int nz, const double *z, int use_fft, double *table)
{
    const int npot = mx * my;
    double *ps = (double *) malloc(sizeof(double) * 3 * npot);
    double *tmp = (double *) calloc(3L * npot, sizeof(double));
    double *vr = (double *) calloc(12L * npot, sizeof(double));      /* vr[k][ik][ii] */
    co_eldiv igs;
    co_eldiv_init(&igs, mx, my);
    memcpy(igs.el, el, sizeof(int) * npot);
    co_areas(&igs);
    for (int iz = 0; iz < nz; iz++) {
        const double zw = z[iz];
        const int ia = (zw >= 0) ? 0 : 1, neg = (zw >= 0) ? 1 : -1;
        co_inflcf ck[4];
        co_sstres_inflcf(mx, my, dx, dy, gg, poiss, zw, ck);
        memcpy(ps, ps_in, sizeof(double) * 3 * npot);
        if (neg < 0) for (int i = 0; i < npot; i++) ps[i] = -ps[i];
        for (int k = 0; k < 4; k++) {
            if (use_fft) co_vecaijpj(cx, &igs, CO_ALLELM, tmp, CO_ALL, ps, &igs, CO_ALL, &ck[k]);
            else {
                co_eldiv full; co_eldiv_init(&full, mx, my);
                for (int i = 0; i < npot; i++) full.el[i] = 1;
                co_areas(&full);
                co_vecaijpj_direct(&full, CO_ALLELM, tmp, CO_ALL, ps, &full, CO_ALL, &ck[k]);
                co_eldiv_free(&full);
            }
            memcpy(vr + (long) k * 3 * npot, tmp, sizeof(double) * 3 * npot);
        }
        for (int ii = 0; ii < npot; ii++) {
            double vr_ii[3][4];
            for (int j = 0; j < 3; j++) for (int k = 0; k < 4; k++) vr_ii[j][k] = vr[(long) k * 3 * npot + (long) j * npot + ii];
            co_sstres_derived(gg[ia], poiss[ia], neg, vr_ii, table + ((long) iz * npot + ii) * 18);
        }
        for (int k = 0; k < 4; k++) co_inflcf_free(&ck[k]);
    }
    co_eldiv_free(&igs);
    free(ps); free(tmp); free(vr);
}
