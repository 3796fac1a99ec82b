/*
 * co_sdis.c -- ORACLE (test infrastructure, not product code).
 * Solver inputs: grid coordinates, undeformed distance, initial element division.
 * Follows /root/reference/src/m_sdis.f90:321-494 (set_norm_rhs), :818-1007 (eldiv0) and
 * /root/reference/src/m_hierarch_data.f90:1801-1885 (element centres).
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <math.h>

/* element centres x(ix) = xl + dx/2 + (ix-1) dx: m_hierarch_data.f90:1834-1835,1875-1885, m_grids.f90:686-693 */
void co_grid_coords(int mx, int my, double xl, double yl, double dx, double dy, double *x, double *y)
{
    const double xc1 = xl + 0.5 * dx, yc1 = yl + 0.5 * dy;
    for (int iy = 1; iy <= my; iy++)
        for (int ix = 1; ix <= mx; ix++) {
            const long ii = (long) (iy - 1) * mx + ix - 1;
            x[ii] = xc1 + (ix - 1) * dx;
            y[ii] = yc1 + (iy - 1) * dy;
        }
}

/* m_sdis.f90:321-494; iplan 1 (unrestricted), 2 (quadratic), 3 (two rectangles) */
void co_set_norm_rhs(int ibase, int iplan, int npot, const double *x, const double *y, int nn,
                     const double *prmudf, const double *prmpln, double *hs)
{
    if (ibase == 1) {
        for (int i = 0; i < npot; i++)
            hs[i] = prmudf[0] * x[i] * x[i] + prmudf[1] * x[i] * y[i] + prmudf[2] * y[i] * y[i]
                  + prmudf[3] * x[i] + prmudf[4] * y[i] + prmudf[5];
    } else if (ibase == 2) {
        const double xm = prmudf[1], rm = prmudf[2], y1 = prmudf[3], dy1 = prmudf[4];
        const double yn = y1 + (nn - 1) * dy1;
        for (int i = 0; i < npot; i++) {
            int mleft; double yleft;
            if (y[i] < y1) { yleft = y1; mleft = 1; }
            else if (y[i] >= yn) { yleft = yn - dy1; mleft = nn - 1; }
            else { mleft = (int) ((y[i] - y1) / dy1) + 1; yleft = y1 + (mleft - 1) * dy1; }
            mleft += 5;                                        /* 1-based index into prmudf */
            const double rc = (prmudf[mleft] - prmudf[mleft - 1]) / dy1;
            hs[i] = prmudf[mleft - 1] + rc * (y[i] - yleft);
            hs[i] = hs[i] + (x[i] - xm) * (x[i] - xm) / (2.0 * rm);
        }
    } else if (ibase == 3) {
        for (int i = 0; i < npot; i++)
            hs[i] = prmudf[0] * sin(prmudf[1] * (x[i] - prmudf[2])) - prmudf[3] * sin(prmudf[4] * (x[i] - prmudf[5]))
                  + x[i] * x[i] / prmudf[6] + y[i] * y[i] / prmudf[7];
    } else if (ibase == 9) {
        for (int i = 0; i < npot; i++) hs[i] = prmudf[i];
    }
    if (iplan == 2) {
        for (int i = 0; i < npot; i++) {
            const double a = prmpln[0] * x[i] * x[i] + prmpln[1] * x[i] * y[i] + prmpln[2] * y[i] * y[i]
                           + prmpln[3] * x[i] + prmpln[4] * y[i] + prmpln[5];
            if (a >= 0) hs[i] = 1e30;
        }
    } else if (iplan == 3) {
        for (int i = 0; i < npot; i++) {
            const int z1 = (prmpln[0] <= x[i] && x[i] <= prmpln[1]) && (prmpln[2] <= y[i] && y[i] <= prmpln[3]);
            const int z2 = (prmpln[4] <= x[i] && x[i] <= prmpln[5]) && (prmpln[6] <= y[i] && y[i] <= prmpln[7]);
            if (!(z1 || z2)) hs[i] = 1e30;
        }
    }
}

/* m_sdis.f90:818-1007.  facpen, reltol are REAL(4) literals stored in doubles (:832). */
void co_eldiv0(int ic_norm, int mx, int my, double dx, double dy, int ibase, const double *prmudf,
               const co_mater *m, double fntrue, double *pen, const double *hs, co_eldiv *igs)
{
    const int npot = mx * my, maxit = 50;
    const double facpen = (double) 0.60f, reltol = (double) 0.01f;
    const double dxdy = dx * dy, fnscal = fntrue / m->ga;
    int imin = 0;
    for (int i = 1; i < npot; i++) if (hs[i] < hs[imin]) imin = i;          /* idmin: first minimum */
    const double hsmin = hs[imin];
    double pentru = *pen - hsmin;

    if (ic_norm == 1 && my == 1) {                                          /* case A: 2-D, :858-891 */
        double rm;
        if (ibase == 1) rm = 0.5 / fmax(1e-6, prmudf[0]);
        else if (ibase == 2) rm = prmudf[2];
        else if (ibase == 3) rm = 0.5 * prmudf[6];
        else rm = 1.0;
        const double cdy = 0.35 - 0.05 * (log(dy) - log(200.0));
        pentru = pow(fnscal / dy * (1.0 - m->nu) / 0.6 / cdy / pow(rm, (double) 0.1f), (double) 0.926f);
        *pen = pentru + hsmin;
    } else if (ic_norm == 1) {                                              /* case B: 3-D, :893-991 */
        const double rmn = 1.0;
        const double fac = rmn * (double) 0.75f * sqrt(CO_PI) * (1.0 - m->nu);
        double penmin = 0.0, fnmin = 0.0, penmax, fnmax, penmid, fnmid;
        int ncnmax = 0, ncnmid, iter = 0;
        penmax = fac * fnscal / sqrt(dxdy);
        for (int i = 0; i < npot; i++) if (hs[i] - hsmin < facpen * penmax) ncnmax++;
        fnmax = penmax * sqrt(dxdy * ncnmax) / fac;
        while (iter < maxit && fabs(fnmax - fnmin) > reltol * fnscal) {
            iter++;
            penmid = 0.5 * (penmax + penmin);
            ncnmid = 0;
            for (int i = 0; i < npot; i++) if (hs[i] - hsmin < facpen * penmid) ncnmid++;
            fnmid = penmid * sqrt(dxdy * ncnmid) / fac;
            if (fnmid < fnscal) { penmin = penmid; fnmin = fnmid; }
            else { penmax = penmid; ncnmax = ncnmid; fnmax = fnmid; }
        }
        pentru = penmax;
        *pen = pentru + hsmin;
    }
    for (int i = 0; i < npot; i++)                                          /* :999-1005 */
        igs->el[i] = (hs[i] - hsmin < facpen * pentru) ? CO_ADHES : CO_EXTER;
}

/* complete elliptic integrals by the AGM: K, E, B = (E - mc K)/m, D = (K - E)/m (finite for m -> 0).  The reference
 * evaluates the same functions with Fukushima's series (m_hertz.f90:446-520, ellip_bd). */
void co_ellip_kebd(double mc, double *K, double *E, double *B, double *D)
{
    const double m = 1.0 - mc;
    double a = 1.0, b = sqrt(mc);
    double t = (1.0 - b) / (4.0 * (1.0 + b));
    double sum = 1.0 + 2.0 * t, pw = 2.0;
    double an = 0.5 * (a + b), bn = sqrt(a * b);
    a = an; b = bn;
    for (int it = 0; it < 40 && fabs(a - b) > 1e-17 * a; it++) {
        an = 0.5 * (a + b); bn = sqrt(a * b);
        const double cn = 0.5 * (a - b);
        pw *= 2.0;
        t = (m > 1e-300) ? cn * cn / m : 0.0;
        sum += pw * t;
        a = an; b = bn;
    }
    *K = CO_PI / (2.0 * a);
    *D = 0.5 * *K * sum;
    *B = *K - *D;
    *E = *B + mc * *D;
}

static double eli_(double k, double *K, double *E)
{
    double B, D;
    co_ellip_kebd(1.0 - k * k, K, E, &B, &D);
    return (*E - B) / B;
}

/* m_hertz.f90:267-385 (hzcalc3d) with bisec (:386-440) */
void co_hertz3d(double e_star, int ipotcn, double *a1, double *b1, double *aa, double *bb, int ic_norm, double *pen,
                double *fn, double *cp, double *rho)
{
    double K, E;
    if (ipotcn == -1) {
        const int zbla = *b1 <= *a1;
        const double y = zbla ? *b1 / *a1 : *a1 / *b1;
        double xl = 0.0, xr = 1.0, elr = eli_(xr, &K, &E), x = 0.5;
        while (fabs(xr - xl) > 1e-9) {
            x = 0.5 * (xl + xr);
            const double elx = eli_(x, &K, &E);
            if ((y - elr) * (y - elx) <= 0.0) xl = x; else { xr = x; elr = elx; }
        }
        eli_(x, &K, &E);
        const double g = sqrt(fmax(1e-40, 1.0 - x * x)), sg = sqrt(g);
        *rho = 2.0 / (*a1 + *b1);
        if (ic_norm == 1) { *cp = pow(3.0 * fmax(0.0, *fn) * *rho * E / (4.0 * CO_PI * e_star * sg), 1.0 / 3.0); *pen = 2.0 * (*cp * sg) * (*cp * sg) * K / (*rho * E); }
        else { *cp = sqrt(fmax(0.0, *pen) * *rho * E / (2.0 * K * sg * sg)); *fn = 4.0 * CO_PI * *cp * *cp * *cp * e_star * sg / (3.0 * *rho * E); }
        if (zbla) { *aa = *cp * sg; *bb = *cp / sg; } else { *aa = *cp / sg; *bb = *cp * sg; }
    } else {
        const int zbla = *bb <= *aa;
        const double g = zbla ? *bb / *aa : *aa / *bb, sg = sqrt(g), k = sqrt(1.0 - g * g);
        const double y = eli_(k, &K, &E);
        *cp = sqrt(*aa * *bb);
        if (ic_norm == 1) { *rho = 4.0 * CO_PI * *cp * *cp * *cp * e_star * sg / (3.0 * *fn * E); *pen = 2.0 * (*cp * sg) * (*cp * sg) * K / (*rho * E); }
        else {
            *pen = fmax(*pen, 1e-9);
            *rho = 2.0 * (*cp * sg) * (*cp * sg) * K / (*pen * E);
            *fn = 4.0 * CO_PI * *cp * *cp * *cp * e_star * sg / (3.0 * *rho * E);
        }
        const double apb = 2.0 / *rho, ama = apb / (y + 1.0), ami = y * ama;
        if (zbla) { *a1 = ami; *b1 = ama; } else { *a1 = ama; *b1 = ami; }
    }
}
