/*
 * co_case.c -- ORACLE (test infrastructure, not product code).
 * One contact case, restating the driver contac() of /root/reference/src/m_scontc.f90:37-216 for module-3
 * problems in the hot-path scope: combin_mater -> sgencr -> set_norm_rhs -> eldiv0 (I=0) -> panprc(snorm).
 */
#include "contact_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* Normal problem only (T=0), non-Hertzian geometry (IPOTCN=1), I=0, P=2.
 * stats_out[0..3] = n_prod, n_rowsum, alg_bytes, alg_flops. Returns 0 / -27 (NORM not converged). */
int co_norm_case(int mx, int my, double xl, double yl, double dx, double dy,
                 double gg1, double gg2, double poiss1, double poiss2,
                 int ibase, int nn, const double *prmudf, int ic_norm, double pen_in, double fn_in,
                 int maxgs, int maxin, double eps, int fullbox,
                 int *el_out, double *pn_out, double *pen_out, double *fn_out, int *itcg_out, int *itnorm_out,
                 double *stats_out)
{
    const int npot = mx * my;
    co_ctx *cx = co_ctx_new();
    cx->fullbox = fullbox;
    co_mater mat = { { gg1, gg2 }, { poiss1, poiss2 }, 0, 0, 0 };
    co_combin_mater(&mat);
    co_inflcf cs, ms;
    memset(&cs, 0, sizeof(cs)); memset(&ms, 0, sizeof(ms));
    co_sgencr(&mat, mx, my, dx, dy, 0, 0.0, 1.0, &cs, NULL, NULL, &ms);

    double *x = (double *) malloc(sizeof(double) * npot), *y = (double *) malloc(sizeof(double) * npot);
    double *hs = (double *) calloc(3L * npot, sizeof(double)), *ps = (double *) calloc(3L * npot, sizeof(double));
    co_grid_coords(mx, my, xl, yl, dx, dy, x, y);
    co_set_norm_rhs(ibase, 1, npot, x, y, nn, prmudf, NULL, hs + 2L * npot);

    co_eldiv igs;
    co_eldiv_init(&igs, mx, my);
    double pen = pen_in, fntrue = fn_in;
    co_eldiv0(ic_norm, mx, my, dx, dy, ibase, prmudf, &mat, fntrue, &pen, hs + 2L * npot, &igs);
    for (int i = 0; i < npot; i++) if (hs[2L * npot + i] > (double) 1e29f) igs.el[i] = CO_EXTER;   /* m_sdis.f90:748-750 */
    co_areas(&igs);

    co_solv solv = { maxgs, maxin, 30, 1, eps };
    co_norm_info info;
    co_snorm(cx, ic_norm, mx, my, dx * dy, &solv, hs, &cs, &ms, &pen, &fntrue, &igs, ps, &info);
    for (int i = 0; i < npot; i++) if (igs.el[i] <= CO_EXTER) ps[2L * npot + i] = 0.0;               /* m_scontc.f90:456-466 */

    memcpy(el_out, igs.el, sizeof(int) * npot);
    memcpy(pn_out, ps + 2L * npot, sizeof(double) * npot);
    *pen_out = pen; *fn_out = fntrue; *itcg_out = info.itcg; *itnorm_out = info.itnorm;
    if (stats_out) {
        stats_out[0] = (double) cx->st.n_prod; stats_out[1] = (double) cx->st.n_rowsum;
        stats_out[2] = cx->st.alg_bytes; stats_out[3] = cx->st.alg_flops;
    }
    co_eldiv_free(&igs); co_inflcf_free(&cs); co_inflcf_free(&ms);
    free(x); free(y); free(hs); free(ps);
    co_ctx_free(cx);
    return info.itnorm < 0 ? -27 : 0;
}

/* Batch of independent normal-contact cases sharing grid and material, one case per OpenMP thread
 * (the pattern of /root/reference/src/test_table.f90:196-292: one result element per thread).  Each thread computes
 * the influence coefficients once and re-uses them for its cases (sgencr's reuse rule, m_visc.f90:153-161), while
 * the preconditioner and all coefficient transforms are redone per case as in the reference (m_snorm.f90:93-97).
 * fn_or_pen[ncase]; outputs el[ncase][npot], pn[ncase][npot], scal[ncase][4] = pen, fn, itcg, n_prod. */
int co_norm_batch(int ncase, int nthreads, int mx, int my, double xl, double yl, double dx, double dy,
                  double gg1, double gg2, double poiss1, double poiss2, int ibase, int nn, const double *prmudf,
                  int ic_norm, const double *fn_or_pen, int maxgs, int maxin, double eps, int fullbox,
                  int *el_out, double *pn_out, double *scal_out)
{
    const long npot = (long) mx * my;
    int nfail = 0;
#pragma omp parallel num_threads(nthreads) reduction(+ : nfail)
    {
        co_ctx *cx = co_ctx_new();
        cx->fullbox = fullbox;
        co_mater mat = { { gg1, gg2 }, { poiss1, poiss2 }, 0, 0, 0 };
        co_combin_mater(&mat);
        co_inflcf cs, ms;
        memset(&cs, 0, sizeof(cs)); memset(&ms, 0, sizeof(ms));
        co_sgencr(&mat, mx, my, dx, dy, 0, 0.0, 1.0, &cs, NULL, NULL, &ms);
        double *x = (double *) malloc(sizeof(double) * npot), *y = (double *) malloc(sizeof(double) * npot);
        double *hs = (double *) calloc(3L * npot, sizeof(double)), *ps = (double *) calloc(3L * npot, sizeof(double));
        co_grid_coords(mx, my, xl, yl, dx, dy, x, y);
        co_set_norm_rhs(ibase, 1, (int) npot, x, y, nn, prmudf, NULL, hs + 2L * npot);
        co_eldiv igs;
        co_eldiv_init(&igs, mx, my);
#pragma omp for schedule(dynamic, 1)
        for (int ic = 0; ic < ncase; ic++) {
            double pen = ic_norm ? 0.0 : fn_or_pen[ic], fntrue = ic_norm ? fn_or_pen[ic] : 0.0;
            const long np0 = cx->st.n_prod;
            memset(ps, 0, sizeof(double) * 3 * npot);
            co_eldiv0(ic_norm, mx, my, dx, dy, ibase, prmudf, &mat, fntrue, &pen, hs + 2L * npot, &igs);
            co_areas(&igs);
            co_solv solv = { maxgs, maxin, 30, 1, eps };
            co_norm_info info;
            co_snorm(cx, ic_norm, mx, my, dx * dy, &solv, hs, &cs, &ms, &pen, &fntrue, &igs, ps, &info);
            for (long i = 0; i < npot; i++) if (igs.el[i] <= CO_EXTER) ps[2L * npot + i] = 0.0;
            memcpy(el_out + ic * npot, igs.el, sizeof(int) * npot);
            memcpy(pn_out + ic * npot, ps + 2L * npot, sizeof(double) * npot);
            scal_out[ic * 4 + 0] = pen; scal_out[ic * 4 + 1] = fntrue; scal_out[ic * 4 + 2] = info.itcg;
            scal_out[ic * 4 + 3] = (double) (cx->st.n_prod - np0);
            if (info.itnorm < 0) nfail++;
        }
        co_eldiv_free(&igs); co_inflcf_free(&cs); co_inflcf_free(&ms);
        free(x); free(y); free(hs); free(ps);
        co_ctx_free(cx);
    }
    return nfail;
}
