// norm_solver.cuh -- device-resident NORM / NormCG: one CTA owns one contact problem for its whole solve.
//
// Mirrors the algorithm of the reference's normcg (/root/reference/src/m_solvpn.f90:24-461) and snorm
// (/root/reference/src/m_snorm.f90:31-378): bound-constrained preconditioned CG (Polak-Ribiere) on the normal
// pressures with contact/exterior active set, prescribed approach (N=0) or prescribed force (N=1, mean deflation).
// B200-first: every influence product is the shared-memory FFT convolution (conv_dev), the masked BLAS-1 steps of
// m_gridfunc.f90:936-1591 are fused into a handful of block-wide passes with fixed-tree reductions, and the
// iteration control (active-set flips, convergence test) never leaves the SM.  Like the reference
// (m_aijpj.f90:774-793), products on the contact area use the bounding box of the contact area with a smaller
// transform when that pays (a ladder of sizes per coefficient set); products on all elements use the full grid.
#pragma once
#include "device_core.cuh"

namespace cb200 {

#define CB_TINY 1e-20

// Products on the contact area (AllInt) are restricted to the bounding box of the contact area, as in the reference
// (m_aijpj.f90:774-793): a ladder of smaller transform sizes with their own transformed coefficients is prepared per
// coefficient set; lev[ly * nlx + lx], sizes descending, lev[0] = the full grid.
struct ConvLevel {
    ConvPlan P;
    const cd *chat[2][3][3];   // [0: cs, 1: ms (preconditioner)][ik][jk]; null = not prepared (the full grid is used then)
};

struct NormCase {
    // inputs
    const double *hs;        // undeformed distance, normal direction (npot)
    const double *ptx, *pty; // tangential tractions for the n-t coupling term (may be null)
    const cd *chatA;         // transformed cs(3,3) * ga_inv / (4 Fx Fy)
    const cd *chatM;         // transformed ms(3,3) * ga_inv / (4 Fx Fy)
    const cd *chatA31, *chatA32;   // transformed cs(3,1), cs(3,2) (null when nt_cpl is false)
    const double *cf33;      // spatial block cs(3,3), cf(-cmx:cmx-1, -cmy:cmy-1)
    int cmx, cmy;
    double ga_inv;
    int ic_norm, maxgs, maxin;
    double eps, dxdy;
    // in/out
    double pen, fntrue;
    int *el;                 // element division (npot)
    double *pn;              // normal pressure (npot)
    double *work;            // 9 * npot doubles
    // outputs
    int itcg, itnorm, ncon, status;     // status bit 0: NormCG diverged at MaxCG (reference: abort_run)
    double err;
    int nprod;               // number of single-block influence products performed (work accounting)
    const ConvLevel *lev;    // ladder of transform sizes (null: always the full grid)
    int nlx, nly;
};

struct ContactBox { int x0, y0, bw, bh; const ConvLevel *lv; };

// bounding box of the contact area (+1 column on either side, m_aijpj.f90:774-777) and the smallest level that holds it
__device__ ContactBox contact_box_dev(const ConvPlan &P, const NormCase &c, const int *el, double *red)
{
    const int mx = P.mx, n = P.npot;
    int x0 = mx, x1 = -1, y0 = P.my, y1 = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (el[i] >= 1) { const int iy = i / mx, ix = i - iy * mx; x0 = min(x0, ix); x1 = max(x1, ix); y0 = min(y0, iy); y1 = max(y1, iy); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
        y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    int *ired = reinterpret_cast<int *>(red);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) { ired[4 * wid] = x0; ired[4 * wid + 1] = x1; ired[4 * wid + 2] = y0; ired[4 * wid + 3] = y1; }
    __syncthreads();
    for (int w = 0; w < nw; w++) { x0 = min(x0, ired[4 * w]); x1 = max(x1, ired[4 * w + 1]); y0 = min(y0, ired[4 * w + 2]); y1 = max(y1, ired[4 * w + 3]); }
    __syncthreads();
    ContactBox b = { 0, 0, mx, P.my, nullptr };
    if (c.lev == nullptr || x1 < x0) return b;
    x0 = max(0, x0 - 1); x1 = min(mx - 1, x1 + 1);
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    if (1.1 * (double) bw * bh > (double) mx * P.my) return b;              // not much smaller: full grid (:783-789)
    int lx = 0, ly = 0;
    while (lx + 1 < c.nlx && c.lev[lx + 1].P.mx >= bw) lx++;
    while (ly + 1 < c.nly && c.lev[(ly + 1) * c.nlx].P.my >= bh) ly++;
    if (lx == 0 && ly == 0) return b;
    b.x0 = x0; b.y0 = y0; b.bw = bw; b.bh = bh; b.lv = &c.lev[ly * c.nlx + lx];
    return b;
}

// AllInt product with A_zz (which = 0) or the preconditioner M_zz (which = 1), on the contact box when a smaller level fits
__device__ __forceinline__ void conv_int_dev(const ConvPlan &P, const Smem &sm, const NormCase &c, const ContactBox &b, int which,
                                             const double *p, double *u, const int *el)
{
    if (b.lv && b.lv->chat[which][2][2]) conv_box_dev(b.lv->P, sm, p, b.lv->chat[which][2][2], u, el, 1, 0, b.x0, b.y0, b.bw, b.bh, P.mx);
    else conv_dev(P, sm, p, which ? c.chatM : c.chatA, u, el, 1, 0);
}

// Masked vector pass with all loads of a batch in flight before any use: the work vectors live in global memory
// (L2-resident), so a pass is latency bound unless its loads are issued back to back.  Each thread handles elements
// tid + k*nt; per batch of CB_VB elements it first loads el[] and the NA input arrays, then calls f(i, el, a[]).
#define CB_VB 8
template <int NA, class F>
__device__ __forceinline__ void vec_pass(int n, const int *el, const double *a0, const double *a1, const double *a2,
                                         const double *a3, F f)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int base = tid; base < n; base += CB_VB * nt) {
        int e[CB_VB];
        double a[CB_VB][NA > 0 ? NA : 1];
#pragma unroll
        for (int k = 0; k < CB_VB; k++) {
            const int i = base + k * nt;
            const bool ok = i < n;
            e[k] = ok ? el[i] : 0;
            if (NA > 0) a[k][0] = ok ? a0[i] : 0.0;
            if (NA > 1) a[k][1] = ok ? a1[i] : 0.0;
            if (NA > 2) a[k][2] = ok ? a2[i] : 0.0;
            if (NA > 3) a[k][3] = ok ? a3[i] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < CB_VB; k++) {
            const int i = base + k * nt;
            if (i < n) f(i, e[k], a[k]);
        }
    }
}

__device__ __forceinline__ void proj_avg_dev(const int *el, double *a, int n, double *red)
{   // gf3_proj_avg(AllInt): m_gridfunc.f90:1325-1370
    double s[2] = { 0.0, 0.0 };
    vec_pass<1>(n, el, a, nullptr, nullptr, nullptr, [&](int, int e, const double *x) { if (e >= 1) { s[0] += x[0]; s[1] += 1.0; } });
    block_sum<2>(s, red);
    const double avg = s[0] / fmax(1.0, s[1]);
    vec_pass<1>(n, el, a, nullptr, nullptr, nullptr, [&](int i, int e, const double *x) { if (e >= 1) a[i] = x[0] - avg; });
    __syncthreads();
}

// returns 1 when the reference would abort (MaxCG reached while diverging)
__device__ int normcg_dev(const ConvPlan &P, const Smem &sm, const NormCase &c, const double *hstot, double &pen,
                          int *el, double *ps, double *wk, int &itcg_out, double &err_out, int &nprod)
{
    const int n = P.npot, tid = threadIdx.x, nt = blockDim.x;
    double *__restrict__ rhs = wk, *__restrict__ res = wk + n, *__restrict__ r_prv = wk + 2 * n,
           *__restrict__ dd = wk + 3 * n, *__restrict__ z = wk + 4 * n, *__restrict__ v = wk + 5 * n,
           *__restrict__ q = wk + 6 * n;
    double *red = sm.red;
    const int ic_norm = c.ic_norm, maxcg = c.maxgs;
    const double eps = c.eps, dxdy = c.dxdy, fntrue = c.fntrue;
    const int numinn = n <= 150 ? 3 : (n <= 400 ? 2 : 1);

    if (ic_norm == 1) pen = 0.0;
    double hmin = 1e20, hmaxn = 1e20;
    double cnt[2] = { 0.0, 0.0 };
    for (int i = tid; i < n; i += nt) {
        const double h = hstot[i];
        rhs[i] = pen - h;
        res[i] = 0.0; r_prv[i] = 0.0; dd[i] = 0.0; z[i] = 0.0; v[i] = 0.0; q[i] = 0.0;
        hmin = fmin(hmin, h); hmaxn = fmin(hmaxn, -h);
        if (el[i] >= 1) cnt[0] += 1.0;
        cnt[1] += ps[i];
    }
    block_sum<2>(cnt, red);
    int ncon = (int) cnt[0];
    const double hsmin0 = block_min(hmin, red);
    double davg = 0.0;

    if (ic_norm == 0) {
        if (hsmin0 - pen >= 0.0) {                                   // m_solvpn.f90:121-136: no contact at all
            for (int i = tid; i < n; i += nt) { ps[i] = 0.0; el[i] = 0; }
            __syncthreads();
            itcg_out = 0; err_out = 0.0;
            return 0;
        }
    } else {
        if (ncon <= 0) {                                             // :144-155
            const double hsmax = -block_min(hmaxn, red);
            const double htrsh = hsmin0 + 0.1 * fmax(hsmax - hsmin0, 1e-10);
            double k[1] = { 0.0 };
            for (int i = tid; i < n; i += nt) if (hstot[i] < htrsh) { el[i] = 1; k[0] += 1.0; }
            block_sum<1>(k, red);
            ncon += (int) k[0];
        }
        double fk = dxdy * cnt[1];                                   // :159-168
        if (fabs(fk) < (double) 1e-3f * fntrue) {
            const double pn = fntrue / (dxdy * (double) ncon);
            for (int i = tid; i < n; i += nt) if (el[i] >= 1) ps[i] = pn;
        } else {
            const double f = fntrue / fk;
            for (int i = tid; i < n; i += nt) if (el[i] >= 1) ps[i] = f * ps[i];
        }
        __syncthreads();
    }

    ContactBox box = contact_box_dev(P, c, el, red);
    conv_int_dev(P, sm, c, box, 0, ps, res, el); nprod++;             // :173-175 res = rhs - A ps on C
    vec_pass<2>(n, el, rhs, res, nullptr, nullptr, [&](int i, int e, const double *a) { if (e >= 1) res[i] = a[0] - a[1]; });
    __syncthreads();
    if (ic_norm == 1) proj_avg_dev(el, res, n, red);

    double rz1, rz2 = 0.0, rms_xk = 1.0, rms_upd = 2.0 * eps * rms_xk, rms_upd1 = 0.0;
    int itcg = 0, itinn = 0;
    bool lchanged = false;

    while ((lchanged || rms_upd > eps * rms_xk) && itcg < maxcg) {   // :194
        itcg++; itinn++;
        conv_int_dev(P, sm, c, box, 1, res, z, el); nprod++;          // z = M res on C
        if (ic_norm == 1) proj_avg_dev(el, z, n, red);

        double d2[2] = { 0.0, 0.0 };
        vec_pass<3>(n, el, z, res, r_prv, nullptr, [&](int, int e, const double *a) {
            if (e >= 1) { d2[0] += a[0] * a[1]; d2[1] += a[0] * a[2]; }
        });
        block_sum<2>(d2, red);
        rz1 = rz2; rz2 = d2[0];

        if (itcg <= 1 || rz1 < CB_TINY) {                            // :228-241
            vec_pass<1>(n, el, z, nullptr, nullptr, nullptr, [&](int i, int e, const double *a) { if (e >= 1) v[i] = a[0]; });
        } else {
            const double beta = fmax(0.0, (rz2 - d2[1]) / fmax(CB_TINY, rz1));
            vec_pass<2>(n, el, z, v, nullptr, nullptr, [&](int i, int e, const double *a) { if (e >= 1) v[i] = beta * a[1] + a[0]; });
        }
        __syncthreads();
        if (ic_norm == 1) proj_avg_dev(el, v, n, red);

        conv_int_dev(P, sm, c, box, 0, v, q, el); nprod++;            // q = A v on C
        if (ic_norm == 1) proj_avg_dev(el, q, n, red);

        double d4[4] = { 0.0, 0.0, 0.0, 0.0 };
        vec_pass<3>(n, el, v, res, q, nullptr, [&](int, int e, const double *a) {
            if (e >= 1) { d4[0] += a[1] * a[0]; d4[1] += a[2] * a[0]; d4[2] += a[0] * a[0]; d4[3] += 1.0; }
        });
        block_sum<4>(d4, red);
        const double rv = d4[0], vav = d4[1];
        double alpha;
        if (fabs(vav) > 1e-32 && ncon == 1) alpha = rv / vav;
        else alpha = rv / fmax(CB_TINY, vav);
        rms_upd = fabs(alpha) * sqrt(d4[2] / fmax(1.0, d4[3]));
        if (itcg == 1) rms_upd1 = rms_upd;
        const bool need_xk = (itcg <= 3 || itcg % 10 == 0);

        double p2[1] = { 0.0 };
        vec_pass<3>(n, el, res, ps, v, nullptr, [&](int i, int e, const double *a) {
            r_prv[i] = a[0];                                          // :294 (AllElm copy)
            if (e >= 1) { const double pi = a[1] + alpha * a[2]; ps[i] = pi; p2[0] += pi * pi; }
        });
        if (need_xk) { block_sum<1>(p2, red); rms_xk = sqrt(p2[0] / fmax(1.0, d4[3])); }
        else __syncthreads();

        if (itinn < numinn && rms_upd >= eps * rms_xk) {             // :298-303
            vec_pass<2>(n, el, res, q, nullptr, nullptr, [&](int i, int e, const double *a) { if (e >= 1) res[i] = a[0] - alpha * a[1]; });
            __syncthreads();
        } else {
            double k2[1] = { 0.0 };
            vec_pass<1>(n, el, ps, nullptr, nullptr, nullptr, [&](int i, int e, const double *a) {   // :310-318
                if (e >= 1 && a[0] < 0.0) { el[i] = 0; ps[i] = 0.0; k2[0] += 1.0; }
            });
            block_sum<1>(k2, red);
            bool lchg_negpn = k2[0] > 0.0;
            ncon -= (int) k2[0];
            if (ncon <= 0) {                                          // :323-333
                double k[1] = { 0.0 };
                for (int i = tid; i < n; i += nt)
                    if (hstot[i] <= hsmin0 + 1e-5) { el[i] = 1; ps[i] = 0.0; k[0] += 1.0; }
                block_sum<1>(k, red);
                ncon += (int) k[0];
                lchg_negpn = true;
            }
            if (ic_norm == 1 && lchg_negpn) {                         // :337-344
                double s[1] = { 0.0 };
                for (int i = tid; i < n; i += nt) s[0] += ps[i];
                block_sum<1>(s, red);
                double fk = dxdy * s[0];
                if (fabs(fk) < (double) 1e-3f * fntrue) {
                    for (int i = tid; i < n; i += nt) if (el[i] >= 1) ps[i] = 1.0;
                    fk = (double) ncon;
                }
                const double f = fntrue / fk;
                for (int i = tid; i < n; i += nt) if (el[i] >= 1) ps[i] = f * ps[i];
                __syncthreads();
            }

            conv_dev(P, sm, ps, c.chatA, dd, el, 0, 0); nprod++;     // :351-352 dd = A ps - rhs, whole grid
            double sd[1] = { 0.0 };
            vec_pass<2>(n, el, dd, rhs, nullptr, nullptr, [&](int i, int e, const double *a) {
                const double d = a[0] - a[1];
                dd[i] = d;
                if (e >= 1) sd[0] += d;
            });
            if (ic_norm == 1) {                                       // :356-359
                block_sum<1>(sd, red);
                davg = sd[0] / (double) ncon;
                for (int i = tid; i < n; i += nt) if (el[i] >= 1) dd[i] -= davg;
            }
            __syncthreads();

            double ke[1] = { 0.0 };
            vec_pass<1>(n, el, dd, nullptr, nullptr, nullptr, [&](int i, int e, const double *a) {   // :361-384
                const double di = a[0];
                double r = 0.0;
                if (e >= 1) r = -di;
                else if (di - davg < 0.0) { el[i] = 1; r = -(di - davg); ke[0] += 1.0; }
                else v[i] = 0.0;
                res[i] = r;
            });
            block_sum<1>(ke, red);
            const bool lchg_intpen = ke[0] > 0.0;
            ncon += (int) ke[0];
            itinn = 0;
            lchanged = lchg_intpen || lchg_negpn;
            if (lchanged) box = contact_box_dev(P, c, el, red);
        }
    }

    if (ic_norm == 1) pen = davg;                                     // :414
    double conv = 1.0;
    if (rms_upd * rms_upd1 > 0.0 && itcg > 1) conv = exp(log(rms_upd / rms_upd1) / (itcg - 1));
    itcg_out = itcg; err_out = rms_upd;
    return (rms_upd > rms_xk && conv > 1.0 && itcg >= maxcg) ? 1 : 0;
}

// |row sum of A_zz| at the central element over the columns AijPj would visit (m_snorm.f90:244-248 with
// m_aijpj.f90:196-211: jx in [row1st(jy)-1, rowlst(jy)+1], empty row: first = mx, last = 0)
__device__ double centre_rowsum_dev(const ConvPlan &P, const NormCase &c, const int *el, double *red)
{
    const int mx = P.mx, my = P.my;
    const int ixm = max(1, mx / 2), iym = max(1, my / 2);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double s[1] = { 0.0 };
    for (int jy = 1 + wid; jy <= my; jy += nw) {
        int first = mx + 1, last = 0;
        for (int jx = 1 + lane; jx <= mx; jx += 32)
            if (el[(jy - 1) * mx + jx - 1] >= 1) { first = min(first, jx); last = max(last, jx); }
        for (int o = 16; o > 0; o >>= 1) {
            first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        }
        if (last == 0) first = mx;
        const int j0 = max(1, first - 1), j1 = min(mx, last + 1);
        const double *row = c.cf33 + (size_t) (iym - jy + c.cmy) * (2 * c.cmx) + c.cmx;
        for (int jx = j0 + lane; jx <= j1; jx += 32) s[0] += row[ixm - jx];
    }
    block_sum<1>(s, red);
    return s[0] * c.ga_inv;
}

__device__ void snorm_dev(const ConvPlan &P, const Smem &sm, NormCase &c)
{
    const int n = P.npot, tid = threadIdx.x, nt = blockDim.x;
    double *wk = c.work, *hstot = wk + 7 * n, *unn = wk + 3 * n /* = dd */, *tmp = wk + 8 * n;
    double *red = sm.red;
    int *el = c.el;
    double *ps = c.pn;
    double pen = c.pen;
    int nprod = 0;

    if (c.chatA31 != nullptr && c.ptx != nullptr) {                   // m_snorm.f90:112-119 hstot = hs + A_zt p_t
        conv_dev(P, sm, c.ptx, c.chatA31, tmp, el, 0, 0);
        conv_dev(P, sm, c.pty, c.chatA32, tmp, el, 0, 1);
        nprod += 2;
        for (int i = tid; i < n; i += nt) hstot[i] = c.hs[i] + tmp[i];
    } else {
        for (int i = tid; i < n; i += nt) hstot[i] = c.hs[i];
    }
    __syncthreads();

    int itnorm = 0, itcg = 0, it = 0, status = 0;
    bool zready;
    double errpn = 0.0;
    do {                                                              // :142-300
        itnorm++;
        zready = true;
        for (int i = tid; i < n; i += nt) if (el[i] < 1) ps[i] = 0.0;
        __syncthreads();

        if (normcg_dev(P, sm, c, hstot, pen, el, ps, wk, it, errpn, nprod)) status |= 1;
        itcg += it;

        double k[1] = { 0.0 };
        for (int i = tid; i < n; i += nt)                             // :197-219 contract
            if (el[i] >= 1 && ps[i] < -errpn) { el[i] = 0; ps[i] = 0.0; k[0] += 1.0; }
        block_sum<1>(k, red);
        if (k[0] > 0.0) zready = false;

        if (zready) {                                                 // :227-292 expand
            conv_dev(P, sm, ps, c.chatA, unn, el, 0, 0); nprod++;
            const double tol = fabs(errpn * centre_rowsum_dev(P, c, el, red));
            double kc[1] = { 0.0 };
            for (int i = tid; i < n; i += nt)
                if (el[i] == 0 && hstot[i] - pen < 0.0) {
                    const double d = hstot[i] - pen + unn[i];
                    if (d < -tol) { el[i] = 1; kc[0] += 1.0; }
                }
            block_sum<1>(kc, red);
            if (kc[0] > 0.0) zready = false;
        }
        if (it >= c.maxgs) zready = false;
    } while (!zready && itnorm < c.maxin);
    if (!zready) itnorm = -1;

    double s[2] = { 0.0, 0.0 };
    for (int i = tid; i < n; i += nt) {                               // :315-325
        if (ps[i] < 0.0 && el[i] >= 1) { ps[i] = 0.0; el[i] = 0; }
        if (el[i] < 1) ps[i] = 0.0;                                   // m_scontc.f90:456-466
        s[0] += ps[i];
        if (el[i] >= 1) s[1] += 1.0;
    }
    block_sum<2>(s, red);
    if (tid == 0) {
        c.pen = pen;
        if (c.ic_norm == 0) c.fntrue = c.dxdy * s[0];                 // :352
        c.itcg = itcg; c.itnorm = itnorm; c.ncon = (int) s[1]; c.status = status; c.err = errpn; c.nprod = nprod;
    }
}

}  // namespace cb200
