// gdsteady_solver.cuh -- device-resident GDsteady (T=3, G=5): the non-linear gradient-descent type solver for steady
// rolling on the traction increments dp, with FFT products for A dv, a diagonal scaling and a Brent line search.
//
// Mirrors /root/reference/src/gdsteady.f90: gdsteady (:9-606), compute_dp (:610-635), project_searchdir (:639-803),
// apply_trcbnd (:807-847), solve_elmtrc (:851-897), compute_diagscaling (:901-1042), perform_linesearch (:1046-1462);
// elastic material, chi = 0.  Written once against the execution context X (BlockCtx: one CTA per case; GridCtx: the
// whole GPU per case), like the other tangential solvers.
//
// Parallel structure: everything per element is an element loop (first()/stride()); the three recurrences along the
// rolling direction (integration dp -> ps from the leading edge with clipping at the traction bound, the projected
// search direction, the leading-edge factor) run one grid row per lane-group: rows are independent, so a row is owned
// by one warp; the warp loads 32 elements of the row coalesced and walks them from the leading edge, every lane
// evaluating the same scalar recurrence on values broadcast by shuffles and keeping the result of its own element.
// The scalars of the line search (tables of alpha, rho, drho; bracket bookkeeping) are computed redundantly and
// identically by all threads from fixed-order reductions.
#pragma once

namespace cb200 {

// cycle counters of GDsteady, added by the leader thread of every solver call (development aid, cb200_gd_prof):
// [0] iterations, [1] cycles in the products A dv / A dp, [2] search direction, [3] line search: integration along the
// rows, [4] line search: element pass + reduction + bracketing, [5] step, active set, diagonal scaling, residual,
// [6] line-search trials, [7] total cycles of the solver calls
__device__ unsigned long long g_gd_prof[12];
#define GD_TICK(slot) do { if (x.leader()) { const long long t_ = clock64(); prof[slot] += (unsigned long long) (t_ - tprev); tprev = t_; } } while (0)

__device__ __forceinline__ double gd_wrap_pi(double e)
{
    const double pi = 3.14159265358979323846;
    if (fabs(e) >= pi) e = e - nearbyint(e / (2.0 * pi)) * 2.0 * pi;
    return e;
}

// leading-edge factor facdt (m_leadedge.f90:92-332, chi = 0) for any execution context; see sxbnd_facdt_dev
template <class X>
__device__ void sxbnd_facdt_x(const X &x, int mx, int my, const int *el, double dx, double dq, double fxdfac, double *facdt)
{
    for (int iy = (int) x.row_first(); iy < my; iy += (int) x.row_stride()) {
        const int *e = el + (size_t) iy * mx;
        double *f = facdt + (size_t) iy * mx;
        int ixb = -1;
        for (int ix = 0; ix < mx; ix++) {
            if (e[ix] < 1) { f[ix] = 0.0; continue; }
            if (ixb < ix) { ixb = ix; while (ixb < mx - 1 && e[ixb + 1] >= 1) ixb++; }
            if (ix + 3 > mx) f[ix] = 1.0;
            else f[ix] = fmin(1.0, (((double) (ixb - ix) + fxdfac) * dx) / dq);
        }
    }
    x.sync();
}

// broadcast within a group of `gw` lanes (gw = 32: the whole warp)
__device__ __forceinline__ double shfl_d(double v, int src, int gw) { return __shfl_sync(0xffffffffu, v, src, gw); }

// The walks below give one grid row to a group of gw lanes (gw = 32, 16, 8 or 4: the largest width with which all rows are in
// flight at once, X::row_group_width).  A group loads gw consecutive elements of its row, then its lanes walk them from the
// leading edge; the 32/gw groups of a warp execute the same instruction stream on different rows, so on the one-CTA path
// (12 warps, 81 rows of tang_problm_1c) the rows take 1 round of walks instead of 7.

// apply_trcbnd (:807-847) on dpnew = dp + alpha dv (on C): integrate the increments from the leading edge (high x) and
// clip at the traction bound g.  ps_out receives the tractions, dp_out (may be null: trial step of the line search) the
// consistent increments, pold (may be null) the previous content of ps_out.
template <class X>
__device__ void gd_apply_trcbnd(const X &x, int mx, int my, const int *el, const double *g, const double *dp, double alpha,
                                const double *dv, double *dp_out, double *ps_out, double *pold)
{
    const int n = mx * my, lane = threadIdx.x & 31;
    const int gw = x.row_group_width(my), lig = lane & (gw - 1), grp = lane / gw, ngrp = 32 / gw;
    for (int r0 = (int) x.warp_first() * ngrp; r0 < my; r0 += (int) x.warp_stride() * ngrp) {
        const int iy = r0 + grp;
        const bool row_ok = iy < my;
        const int i0 = iy * mx;
        double prx = 0.0, pry = 0.0;                          // ps of the element to the right (uniform over the group)
        for (int base = ((mx - 1) / gw) * gw; base >= 0; base -= gw) {
            __syncwarp();
            const int ix = base + lig, ii = i0 + ix;
            const bool have = row_ok && ix < mx;
            int e = 0; double gg = 0.0, dx_ = 0.0, dy_ = 0.0;
            if (have) {
                e = el[ii]; gg = g[ii]; dx_ = dp[ii]; dy_ = dp[n + ii];
                if (e >= EL_ADHES) { dx_ += alpha * dv[ii]; dy_ += alpha * dv[n + ii]; }
                if (pold) { pold[ii] = ps_out[ii]; pold[n + ii] = ps_out[n + ii]; }
            }
            double mypx = 0.0, mypy = 0.0, mydx = dx_, mydy = dy_;
            const int kmax = min(gw - 1, mx - 1 - base);
            if (__all_sync(0xffffffffu, e <= EL_EXTER)) {
                // exterior elements only: no walk; zero tractions, and only the element next to the chunk on the right
                // sees a non-zero neighbour (about half of a potential contact area is exterior)
                if (ix != mx - 1) { mydx = (lig == kmax) ? 0.0 - prx : 0.0; mydy = (lig == kmax) ? 0.0 - pry : 0.0; }
                prx = 0.0; pry = 0.0;
            } else
            for (int k = kmax; k >= 0; k--) {
                const int ek = __shfl_sync(0xffffffffu, e, k, gw);
                const double gk = shfl_d(gg, k, gw), dkx = shfl_d(dx_, k, gw), dky = shfl_d(dy_, k, gw);
                double px = 0.0, py = 0.0;
                const bool last = (base + k == mx - 1);
                if (!last && ek > EL_EXTER) {
                    px = prx + dkx; py = pry + dky;
                    const double p2 = px * px + py * py;
                    // (square root and divisions only where the bound can be active: the filter is conservative
                    //  by more than the rounding of the square root, the decision itself is the reference's)
                    if (ek == EL_SLIP || p2 >= gk * gk * (1.0 - 1e-15)) {
                        const double pa = sqrt(p2);
                        if (ek == EL_SLIP || pa >= gk) { px = px * gk / pa; py = py * gk / pa; }
                    }
                }
                if (lig == k) { mypx = px; mypy = py; if (!last) { mydx = px - prx; mydy = py - pry; } }
                prx = px; pry = py;
                __syncwarp();
            }
            if (have) {
                ps_out[ii] = mypx; ps_out[n + ii] = mypy;
                if (dp_out) { dp_out[ii] = mydx; dp_out[n + ii] = mydy; }
            }
        }
    }
    x.sync();
}

// project_searchdir (:639-803): dv in/out, v out.  imeth 1: E_trl, 2: E_down(kdown), 3: E_keep(fdecay).  One group of lanes
// per row, like gd_apply_trcbnd.  E_down looks kdown elements ahead in a copy of dv in which every non-adhesion element
// zeroes the kdown entries to its right as the row is walked (:700-790); seen from the adhesion element at ix the entry
// ix + kdown is still intact exactly when the elements ix+1 .. ix+kdown-1 are all in adhesion, so the walk only carries
// the length of the adhesion run to the right (`clean`) and reads the look-ahead from the unmodified copy scr.
template <class X>
__device__ void gd_project_searchdir(const X &x, int mx, int my, const int *el, const double *g, int imeth, int kdown, double fdecay,
                                     const double *nn, double *dv, double *v, double *scr, double fac_v,
                                     unsigned long long *prof = nullptr)
{
    const int n = mx * my;
    long long tq0 = clock64();
    if (imeth == 2) {
        for (size_t i = x.first(); i < (size_t) (2 * n); i += x.stride()) scr[i] = dv[i];
        x.sync();
        if (prof) { prof[11] += 1; }
    }
    long long tq1 = clock64();
    const int lane = threadIdx.x & 31;
    const int gw = x.row_group_width(my), lig = lane & (gw - 1), grp = lane / gw, ngrp = 32 / gw;
    for (int r0 = (int) x.warp_first() * ngrp; r0 < my; r0 += (int) x.warp_stride() * ngrp) {
        const int iy = r0 + grp;
        const bool row_ok = iy < my;
        const int i0 = iy * mx;
        double vrx = 0.0, vry = 0.0;                          // v of the element to the right (uniform over the group)
        int clean = 0;                                        // adhesion run to the right of the current element
        for (int base = ((mx - 1) / gw) * gw; base >= 0; base -= gw) {
            __syncwarp();
            const int ix = base + lig, ii = i0 + ix;
            const bool have = row_ok && ix < mx;
            int e = 0; double dx_ = 0.0, dy_ = 0.0, tx = 0.0, ty = 0.0, lim = 0.0, lx_ = 0.0, ly_ = 0.0;
            if (have) {
                e = el[ii]; dx_ = dv[ii]; dy_ = dv[n + ii];
                if (e == EL_SLIP) { tx = -nn[n + ii]; ty = nn[ii]; lim = fac_v * g[ii]; }
                if (imeth == 2 && e == EL_ADHES && ix + kdown <= mx - 1) { lx_ = scr[ii + kdown]; ly_ = scr[n + ii + kdown]; }
            }
            double myvx = 0.0, myvy = 0.0, mydx = dx_, mydy = dy_;
            const int kmax = min(gw - 1, mx - 1 - base);
            if (__all_sync(0xffffffffu, e <= EL_EXTER)) {                 // exterior elements only: no walk (see gd_apply_trcbnd)
                if (ix != mx - 1) { mydx = (lig == kmax) ? 0.0 - vrx : 0.0; mydy = (lig == kmax) ? 0.0 - vry : 0.0; }
                vrx = 0.0; vry = 0.0;
                clean = (kmax == 0 && base == mx - 1) ? 1 : 0;
            } else
            // the walk: all broadcasts of a step are issued unconditionally at its top and the three element kinds are
            // evaluated by selection, so that the loop body is straight-line code for every group of the warp
            for (int k = kmax; k >= 0; k--) {
                const int ek = __shfl_sync(0xffffffffu, e, k, gw);
                const double dkx = shfl_d(dx_, k, gw), dky = shfl_d(dy_, k, gw);
                const double tkx = shfl_d(tx, k, gw), tky = shfl_d(ty, k, gw), lk = shfl_d(lim, k, gw);
                double lkx = 0.0, lky = 0.0;
                if (imeth == 2) { lkx = shfl_d(lx_, k, gw); lky = shfl_d(ly_, k, gw); if (clean < kdown - 1) { lkx = 0.0; lky = 0.0; } }
                const bool last = (base + k == mx - 1), adh = (ek == EL_ADHES) && !last, slp = (ek == EL_SLIP) && !last;
                // adhesion: accumulate (E_keep damps the running sum, E_down subtracts the look-ahead entry)
                const double fr = (imeth == 3) ? fdecay : 1.0;
                const double ax = (fr * vrx + dkx) - lkx, ay = (fr * vry + dky) - lky;
                // slip: component along the tangent, limited
                double vt = tkx * dkx + tky * dky;
                vt = copysign(1.0, vt) * fmin(fabs(vt), lk);
                const double sx = tkx * vt, sy = tky * vt;
                const double vx = adh ? ax : (slp ? sx : 0.0), vy = adh ? ay : (slp ? sy : 0.0);
                if (lig == k) {
                    myvx = vx; myvy = vy;
                    if (!last && !(adh && imeth == 1)) { mydx = vx - vrx; mydy = vy - vry; }
                }
                clean = last ? 1 : (adh ? clean + 1 : 0);
                vrx = vx; vry = vy;
            }
            if (have) { v[ii] = myvx; v[n + ii] = myvy; dv[ii] = mydx; dv[n + ii] = mydy; }
        }
    }
    long long tq2 = clock64();
    x.sync();
    if (prof) { prof[8] += (unsigned long long) (tq1 - tq0); prof[9] += (unsigned long long) (tq2 - tq1); prof[10] += (unsigned long long) (clock64() - tq2); }
}

// compute_diagscaling (:901-1042) + residual r = -D s (adhesion) / -D (s.t) t (slip), per element
template <class X>
__device__ void gd_diagscaling_residual(const X &x, int mx, int my, const int *el, const double *g, double c00, double c01, double c10,
                                        double c11, const double *ps, const double *ss, const double *nn, const GdParams &sp,
                                        double *dscl, double *r)
{
    const int n = mx * my;
    const double tiny_err = 1e-6, epselm = 1e-6;
    for (size_t i = x.first(); i < (size_t) n; i += x.stride()) {
        const int e0 = el[i];
        int elnew = e0;
        double fac_s = 0.0, d = 0.0;
        const double sx0 = ss[i], sy0 = ss[n + i];
        if (e0 == EL_SLIP) {
            double px = ps[i], py = ps[n + i], sx = sx0, sy = sy0;
            const double bound = g[i];
            double th_p0 = atan2(py, px), th_s0 = atan2(-sy, -sx);
            double err_0 = gd_wrap_pi(th_s0 - th_p0);
            if (fabs(err_0) <= tiny_err) {
                px = bound * cos(th_p0 + 0.1); py = bound * sin(th_p0 + 0.1);
                sx = sx0 + c00 * (px - ps[i]) + c01 * (py - ps[n + i]);
                sy = sy0 + c10 * (px - ps[i]) + c11 * (py - ps[n + i]);
                th_p0 = atan2(py, px); th_s0 = atan2(-sy, -sx);
                err_0 = gd_wrap_pi(th_s0 - th_p0);
            }
            if (fabs(err_0) > tiny_err) {
                plstrc_dev(elnew, c00, c01, c10, c11, epselm, 1.0, 1.0, px, py, bound, sx, sy);
                if (elnew == EL_SLIP) {
                    const double th_s1 = atan2(-sy, -sx);
                    fac_s = fabs(gd_wrap_pi(th_s1 - th_s0)) / fabs(err_0);
                }
            }
        }
        if (e0 == EL_SLIP && elnew == EL_SLIP) d = sp.d_slp * pow(fac_s, sp.pow_s);
        else if (e0 == EL_ADHES || (e0 == EL_SLIP && elnew == EL_ADHES)) {
            const int iy = (int) (i / mx), ix = (int) i - iy * mx;
            int jx = ix;
            while (jx > 0 && el[i - ix + jx] == EL_ADHES) jx--;
            double dnew;
            if (sp.d_cns < sp.d_ifc) dnew = fmax(sp.d_cns, sp.d_ifc + (ix - jx - 1) * fmin(0.0, sp.d_lin));
            else dnew = fmin(sp.d_cns, sp.d_ifc + (ix - jx - 1) * fmax(0.0, sp.d_lin));
            d = fmin(50.0 * dscl[i], dnew);
        }
        dscl[i] = d;
        double rx = 0.0, ry = 0.0;
        if (e0 == EL_ADHES) { rx = -sx0 * d; ry = -sy0 * d; }
        else if (e0 == EL_SLIP) {
            const double tx = -nn[n + i], ty = nn[i], st = tx * sx0 + ty * sy0;
            rx = -st * tx * d; ry = -st * ty * d;
        }
        r[i] = rx; r[n + i] = ry;
    }
    x.sync();
}

// ss = A_tt dp + ws on C, zero elsewhere
template <class X>
__device__ void gd_slip(const X &x, ContactCase &c, const double *dp, const double *ws, double *ss, int &nprod)
{
    const int n = x.n();
    const int *el = c.nrm.el;
    nprod += conv_multi(x, c.chatA, dp, 0, 1, ss, 0, 1, el, 1, 0);
    for (size_t i = x.first(); i < (size_t) n; i += x.stride()) {
        if (el[i] >= EL_ADHES) { ss[i] += ws[i]; ss[n + i] += ws[n + i]; } else { ss[i] = 0.0; ss[n + i] = 0.0; }
    }
    x.sync();
}

// gdsteady (:9-606).  ws: [2][n] right-hand side; returns itgd (negative: stagnation estimated from the convergence rate)
template <class X>
__device__ __noinline__ int gdsteady_dev(const X &x, ContactCase &c, const double *ws, int maxgd, double eps, double &err_out, int &lstagn, int &nprod)
{
    const int n = x.n(), mx = x.plan().mx, my = x.plan().my;
    int *el = c.nrm.el;
    double *ps = c.ps, *ss = c.ss;
    const double *psn = c.ps + 2 * (size_t) n;
    double *g = c.gwork, *dp = g + n, *dscl = dp + 2 * (size_t) n, *nn = dscl + n, *r = nn + 2 * (size_t) n, *dv = r + 2 * (size_t) n,
           *v = dv + 2 * (size_t) n, *q = v + 2 * (size_t) n, *scr = q + 2 * (size_t) n;      // scr: pold / psnew / dxin (16 n in total)
    const GdParams sp = c.gd;
    const double tiny = 1e-12, fac_v = 1000.0, mu = c.fstat;
    const size_t ctr = (size_t) c.nrm.cmy * 2 * c.nrm.cmx + c.nrm.cmx;
    const double c00 = c.c11 * c.nrm.ga_inv, c11 = c.c22 * c.nrm.ga_inv, c01 = c.cf12[ctr] * c.nrm.ga_inv, c10 = c01;
    int nadh, nslip;
    count_el(x, el, n, nadh, nslip);
    const double facnel = (double) sqrtf(__fdiv_rn((float) n, (float) (nadh + nslip)));
    int itgd = 0, it_fb = -99;
    bool lchanged = false;
    unsigned long long prof[12] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    long long tprev = clock64();
    const long long tstart = tprev;
    double dif = 2.0, difid = 1.0, dif1 = 0.0, beta = 1.0, alpha = 0.0, alpha0 = 0.0;
    lstagn = 0;

    for (size_t i = x.first(); i < (size_t) n; i += x.stride()) {
        g[i] = mu * psn[i]; dscl[i] = 1.0;
        const double th = atan2(ps[n + i], ps[i]);
        nn[i] = cos(th); nn[n + i] = sin(th);
        const int ix = (int) (i % mx);
        dp[i] = (ix == mx - 1) ? ps[i] : ps[i] - ps[i + 1];                      // compute_dp (:610-635)
        dp[n + i] = (ix == mx - 1) ? ps[n + i] : ps[n + i] - ps[n + i + 1];
        ss[i] = 0.0; ss[n + i] = 0.0; q[i] = 0.0; q[n + i] = 0.0; v[i] = 0.0; v[n + i] = 0.0;
    }
    x.sync();
    gd_slip(x, c, dp, ws, ss, nprod);
    gd_diagscaling_residual(x, mx, my, el, g, c00, c01, c10, c11, ps, ss, nn, sp, dscl, r);
    {
        double m[1] = { 0.0 };
        for (size_t i = x.first(); i < (size_t) (2 * n); i += x.stride()) if (fabs(r[i]) >= tiny) m[0] += 1.0;
        x.template sum<1>(m);
        if (m[0] == 0.0) dif = 0.0;
    }

    while ((lchanged || dif > difid) && itgd < maxgd) {
        itgd++;
        for (size_t i = x.first(); i < (size_t) (2 * n); i += x.stride()) dv[i] = r[i];
        x.sync();
        int imeth, kdown;
        if (beta > sp.betath || itgd - it_fb <= 1) { imeth = sp.gd_meth; kdown = sp.kdown; }
        else { imeth = 2; kdown = sp.kdowfb; it_fb = itgd; }
        GD_TICK(5);
        gd_project_searchdir(x, mx, my, el, g, imeth, kdown, sp.fdecay, nn, dv, v, scr, fac_v, x.leader() ? prof : nullptr);
        GD_TICK(2);
        nprod += conv_multi(x, c.chatA, dv, 0, 1, q, 0, 1, el, 1, 0);
        GD_TICK(1);
        double sm4[4] = { 0.0, 0.0, 0.0, 0.0 };                                  // r.r, (r.q) d over all; d^2 q.q, r.r over C
        for (size_t i = x.first(); i < (size_t) n; i += x.stride()) {
            const bool in = el[i] >= EL_ADHES;
            const double qx = in ? q[i] : 0.0, qy = in ? q[n + i] : 0.0;        // q = A dv is defined on C only
            if (!in) { q[i] = 0.0; q[n + i] = 0.0; }
            const double rr = r[i] * r[i] + r[n + i] * r[n + i];
            sm4[0] += rr; sm4[1] += (r[i] * qx + r[n + i] * qy) * dscl[i];
            if (in) { sm4[2] += dscl[i] * dscl[i] * (qx * qx + qy * qy); sm4[3] += rr; }
        }
        x.template sum<4>(sm4);
        if (itgd <= 1) alpha0 = sm4[0] / sm4[1]; else alpha0 = 0.5 * (alpha0 + alpha);

        // ---- perform_linesearch (:1046-1462): Brent's method on rho(alpha) = |D (s + alpha q)|^2 with projected tractions
        {
            const int max_j = 99;
            const double eps_j = 0.001;
            double tbl_alpha[max_j + 2], tbl_rho[max_j + 2], tbl_drho[max_j + 2];
            int j = 0, ita = 0, itb_ = 0, itx = 0, itw = 0, itv = 0, itu = 0, ibrack = 1;
            bool stop_j = false, has_bracket = false;
            double alpha_j = 0.0, dst_cur = 0.0, dst_prv1 = 0.0, dst_prv2 = 0.0;
            while (!stop_j) {
                GD_TICK(4);
                gd_apply_trcbnd(x, mx, my, el, g, dp, alpha_j, dv, (double *) nullptr, scr, (double *) nullptr);
                GD_TICK(3);
                double s6[6] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };               // rho, c0adh, c1adh, c0slp, c1slp, c2slp
                for (size_t i = x.first(); i < (size_t) n; i += x.stride()) {
                    const int e0 = el[i];
                    if (e0 < EL_ADHES) continue;
                    const double gi = g[i], sx = ss[i], sy = ss[n + i], qx = q[i], qy = q[n + i];
                    const double vt = v[i] * (-nn[n + i]) + v[n + i] * nn[i];
                    const double dth_da = (gi * vt) / (fmax(tiny, gi * gi) + alpha_j * alpha_j * vt * vt);
                    // direction of the trial traction: (cos, sin)(atan2(py, px)) of the reference is the normalised
                    // vector ((1, 0) for the zero vector), evaluated here without the three transcendentals -- this pass
                    // runs once per line-search trial
                    const double qx_ = scr[i], qy_ = scr[n + i], q2_ = qx_ * qx_ + qy_ * qy_;
                    double nx = 1.0, ny = 0.0;
                    if (q2_ > 0.0) { const double qi = 1.0 / sqrt(q2_); nx = qx_ * qi; ny = qy_ * qi; }
                    const double tx = -ny, ty = nx;
                    const double sn = sx * nx + sy * ny;
                    int elnew = e0;
                    if (elnew == EL_SLIP && sn > 0.0) elnew = EL_ADHES;
                    const double st = sx * tx + sy * ty, rt = -st;
                    const double qn = qx * nx + qy * ny, qt = qx * tx + qy * ty;
                    const double d2 = dscl[i] * dscl[i];
                    if (elnew == EL_ADHES) {
                        s6[1] += 2.0 * d2 * (sx * qx + sy * qy);
                        s6[2] += 2.0 * d2 * (qx * qx + qy * qy);
                        s6[0] += d2 * ((sx + alpha_j * qx) * (sx + alpha_j * qx) + (sy + alpha_j * qy) * (sy + alpha_j * qy));
                    } else if (elnew == EL_SLIP) {
                        s6[3] -= 2.0 * d2 * (rt * qt - rt * sn * dth_da);
                        s6[4] -= 2.0 * d2 * (-qt * qt + qt * sn * dth_da - rt * qn * dth_da);
                        s6[5] -= 2.0 * d2 * (qt * qn * dth_da);
                        s6[0] += d2 * (st + alpha_j * qt) * (st + alpha_j * qt);
                    }
                }
                x.template sum<6>(s6);
                const double rhonew = s6[0];
                const double drho_da = s6[1] + s6[3] + alpha_j * (s6[2] + s6[4]) + alpha_j * alpha_j * s6[5];
                int itb = 1;                                                    // insert in the table (1-based, sorted on alpha)
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
                const bool prev_bracket = has_bracket;
                has_bracket = false;
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
                const double a_prev = alpha_j;
                if (has_bracket) {
                    const double xmid = 0.5 * (tbl_alpha[ita] + tbl_alpha[itb_]);
                    const double tolx = eps_j * fabs(tbl_alpha[itx]) + tiny;
                    const bool ldone = (fabs(tbl_alpha[itx] - xmid) <= 2.0 * tolx - 0.5 * (tbl_alpha[itb_] - tbl_alpha[ita]));
                    if (!ldone) {
                        bool use_parab = false;
                        if (fabs(dst_prv2) > tolx) {
                            double br = (tbl_alpha[itx] - tbl_alpha[itw]) * (tbl_rho[itx] - tbl_rho[itv]);
                            double bq = (tbl_alpha[itx] - tbl_alpha[itv]) * (tbl_rho[itx] - tbl_rho[itw]);
                            double bp = (tbl_alpha[itx] - tbl_alpha[itv]) * bq - (tbl_alpha[itx] - tbl_alpha[itw]) * br;
                            bq = 2.0 * (bq - br);
                            if (bq > 0.0) bp = -bp;
                            bq = fabs(bq);
                            if (!(fabs(bp) >= fabs(0.5 * bq * dst_prv2) || bp <= bq * (tbl_alpha[ita] - tbl_alpha[itx]) ||
                                  bp >= bq * (tbl_alpha[itb_] - tbl_alpha[itx]))) {
                                use_parab = true;
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
            if (j >= max_j) {
                int im = 1;
                for (int k = 2; k <= j; k++) if (tbl_rho[k] < tbl_rho[im]) im = k;
                alpha_j = tbl_alpha[im];
            }
            alpha = alpha_j;
            if (x.leader()) { c.gd_ntrial += j; prof[6] += j; }
            GD_TICK(4);
        }

        beta = alpha * sqrt(sm4[2]) / sqrt(sm4[3]);
        // the step: dp += alpha dv on C, integrate and clip (scr keeps the previous tractions)
        gd_apply_trcbnd(x, mx, my, el, g, dp, alpha, dv, dp, ps, scr);
        for (size_t i = x.first(); i < (size_t) n; i += x.stride()) {
            const double th = atan2(ps[n + i], ps[i]);
            nn[i] = cos(th); nn[n + i] = sin(th);
        }
        x.sync();
        GD_TICK(5);
        gd_slip(x, c, dp, ws, ss, nprod);
        GD_TICK(1);
        double ch[1] = { 0.0 };
        for (size_t i = x.first(); i < (size_t) n; i += x.stride()) {
            const int e0 = el[i];
            if (e0 == EL_ADHES) {
                const double pa = sqrt(ps[i] * ps[i] + ps[n + i] * ps[n + i]);
                if (pa >= g[i] - tiny) { el[i] = EL_SLIP; ch[0] += 1.0; }
            } else if (e0 == EL_SLIP) {                                          // solve_elmtrc (:851-897) for this element
                double px = ps[i], py = ps[n + i], sx = ss[i], sy = ss[n + i];
                int eo = e0;
                plstrc_dev(eo, c00, c01, c10, c11, 1e-6, 1.0, 1.0, px, py, g[i], sx, sy);
                const double snrm = nn[i] * ss[i] + nn[n + i] * ss[n + i];
                const double cosdth = (nn[i] * px + nn[n + i] * py) / g[i];
                if ((snrm > 0.0 && (dscl[i] >= 0.001 || eo == EL_ADHES)) || cosdth <= -0.5) { el[i] = EL_ADHES; ch[0] += 1.0; }
            }
        }
        x.template sum<1>(ch);                                                   // (also orders the el updates before their use)
        lchanged = ch[0] > 0.0;
        gd_diagscaling_residual(x, mx, my, el, g, c00, c01, c10, c11, ps, ss, nn, sp, dscl, r);
        double s2[2] = { 0.0, 0.0 };
        for (size_t i = x.first(); i < (size_t) n; i += x.stride()) {
            double dx_ = scr[i], dy_ = scr[n + i];
            if (el[i] >= EL_ADHES) { dx_ -= ps[i]; dy_ -= ps[n + i]; }
            s2[0] += dx_ * dx_ + dy_ * dy_;
            s2[1] += ps[i] * ps[i] + ps[n + i] * ps[n + i];
        }
        x.template sum<2>(s2);
        const double facdif = beta >= 0.1 ? 1.0 : (beta >= 0.001 ? 0.1 / beta : 100.0);
        dif = facnel * sqrt(s2[0] / (2.0 * n)) * facdif;
        difid = eps * fmax(1e-6, facnel * sqrt(s2[1] / (2.0 * n)));
        if (itgd == 1) dif1 = dif;
    }

    double s2[2] = { 0.0, 0.0 };
    for (size_t i = x.first(); i < (size_t) n; i += x.stride()) if (el[i] >= EL_ADHES) { s2[0] += r[i] * r[i] + r[n + i] * r[n + i]; s2[1] += 2.0; }
    x.template sum<2>(s2);
    const double resrms = sqrt(s2[0] / fmax(1.0, s2[1]));
    const double res_dp = resrms / fmax(c00, c11);
    err_out = dif;
    double conv = 1.0;
    if (dif * dif1 > 0.0 && itgd > 1) conv = exp(log(dif / dif1) / (itgd - 1));
    if (lchanged && itgd >= maxgd) lstagn = 1;
    else if (dif > difid && conv > 1.0 && itgd >= maxgd) lstagn = 1;
    else if (conv < 1.0 - tiny && res_dp > 5.0 * difid / (1.0 - conv)) { itgd = -itgd; lstagn = 1; }
    GD_TICK(5);
    if (x.leader()) {
        prof[0] = (unsigned long long) abs(itgd); prof[7] = (unsigned long long) (clock64() - tstart);
        for (int k = 0; k < 12; k++) atomicAdd(&g_gd_prof[k], prof[k]);
    }
    return itgd;
}

}  // namespace cb200
