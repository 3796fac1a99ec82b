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
    int stage_bytes;         // bytes at the bottom of the shared-memory window that no plan of the ladder keeps tables in
                             // (room for the staged vector passes, 0: none)
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

// Product on the box (x0, y0, bw x bh) with element-wise work fused into its first and last stage (ConvFuse) when the
// plan has the warp-resident path; the same work as explicit passes around the block-wide product otherwise.  Returns the
// fused masked sum.  p is written when fuse.in_mode != 0.
__device__ __noinline__ double conv_fx_dev(const ConvPlan &P, const Smem &sm, double *p, const cd *chat, double *u, const int *el,
                              int mask_mode, int x0, int y0, int bw, int bh, int stride, const ConvFuse f)
{
    if (P.c2.ok) return conv2_box_dev(P, sm, p, chat, u, el, mask_mode, 0, x0, y0, bw, bh, stride, f);
    const int tid = threadIdx.x, nt = blockDim.x, nbox = bw * bh;
    if (f.in_mode != 0) {
        for (int w = tid; w < nbox; w += nt) {
            const int iy = w / bw, ix = w - iy * bw;
            const size_t i = (size_t) (y0 + iy) * stride + x0 + ix;
            if (el[i] >= 1) p[i] = f.in_mode == 2 ? f.in_shift * p[i] : p[i] - f.in_shift;
        }
        __syncthreads();
    }
    conv_box_dev(P, sm, p, chat, u, el, mask_mode, 0, x0, y0, bw, bh, stride);
    double s[1] = { 0.0 };
    if (f.out_sub != nullptr || f.out_sum) {
        for (int w = tid; w < nbox; w += nt) {
            const int iy = w / bw, ix = w - iy * bw;
            const size_t i = (size_t) (y0 + iy) * stride + x0 + ix;
            const int e = el[i];
            if (mask_mode == 1 && e < 1) continue;
            double v = u[i];
            if (f.out_sub != nullptr) { v -= f.out_sub[i]; u[i] = v; }
            if (e >= 1) s[0] += v;
        }
        block_sum<1>(s, sm.red);
    }
    return s[0];
}

// the same on the contact box with A_zz (which = 0) or M_zz (which = 1) when a smaller level fits, else on the full grid
__device__ __forceinline__ double conv_int_fx_dev(const ConvPlan &P, const Smem &sm, const NormCase &c, const ContactBox &b, int which,
                                                  double *p, double *u, const int *el, const ConvFuse f)
{
    if (b.lv && b.lv->chat[which][2][2])
        return conv_fx_dev(b.lv->P, sm, p, b.lv->chat[which][2][2], u, el, 1, b.x0, b.y0, b.bw, b.bh, P.mx, f);
    return conv_fx_dev(P, sm, p, which ? c.chatM : c.chatA, u, el, 1, 0, 0, P.mx, P.my, P.mx, f);
}

// Masked vector pass with all loads of a batch in flight before any use: the work vectors live in global memory
// (L2-resident), so a pass is latency bound unless its loads are issued back to back.  Each thread handles elements
// tid + k*nt; per batch of CB_VB elements it first loads el[] and the NA input arrays, then calls f(i, el, a[]).
#define CB_VB 11
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

// ---- vector passes on shared-memory copies fetched by bulk-async copies (TMA engine) ----
// Between two products the S / W window of the plan is free.  A pass over up to three work vectors asks the copy engine
// for the whole vectors (one cp.async.bulk each, completion on an mbarrier), prefetches its slice of the element division
// into registers meanwhile, and then computes from shared memory: one L2 round trip per pass instead of one per batch of
// per-thread loads (a thread can keep ~8 loads of 8 bytes in flight: 384 threads x 64 B / ~1000 cycles = 25 B per cycle,
// which made every pass cost 12 - 35 k cycles).  Results go straight to global memory (stores do not wait).
struct VecStage {
    uint32_t a0;             // shared-window address of the window base
    uint32_t bar;            // mbarrier of the copies (shared address)
    uint32_t par;            // parity of its next phase (uniform over the CTA)
    uint32_t bstride;        // bytes per buffer (multiple of 16); 0: staging not available (window too small)
};
#define CB_VS_MAXK 24        // elements per thread held in registers (element division): n <= CB_VS_MAXK * blockDim.x

__device__ __forceinline__ VecStage vec_stage_init(const Smem &sm, int n, int window_bytes)
{
    VecStage vs;
    vs.a0 = sm.a0; vs.bar = conv_hdr_bar(sm) + 16u; vs.par = 0u;
    // the barrier lives for the whole kernel: the parity of its next phase is kept beside it (red[115]), written by thread 0
    // after every copy it waited for
    __syncthreads();
    vs.par = (uint32_t) *reinterpret_cast<volatile int *>(sm.red + CB_HDR_SLOT + 3);
    const uint32_t b = ((uint32_t) n * 8u + 16u + 15u) & ~15u;      // the copy starts at the 16-byte boundary below the vector
    vs.bstride = (3u * b <= (uint32_t) window_bytes && n <= CB_VS_MAXK * (int) blockDim.x) ? b : 0u;
    return vs;
}

template <int NA, class F>
__device__ __forceinline__ void staged_pass(VecStage &vs, int n, const int *el, const double *a0, const double *a1,
                                            const double *a2, F f)
{
    static_assert(NA >= 1 && NA <= 3, "up to three staged vectors");
    if (vs.bstride == 0u) { vec_pass<NA>(n, el, a0, a1, a2, nullptr, f); return; }
    const int tid = threadIdx.x, nt = blockDim.x;
    // everything written so far (this window by the product, the vectors by earlier passes) must be visible to the copy engine
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    const double *src[3] = { a0, a1, a2 };
    if (tid == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int k = 0; k < NA; k++) tot += ((uint32_t) (((uintptr_t) src[k]) & 15u) + (uint32_t) n * 8u + 15u) & ~15u;
        mbar_expect_tx(vs.bar, tot);
#pragma unroll
        for (int k = 0; k < NA; k++) {
            const uint32_t sh = (uint32_t) (((uintptr_t) src[k]) & 15u);
            bulk_g2s(vs.a0 + (uint32_t) k * vs.bstride, reinterpret_cast<const char *>(src[k]) - sh, (sh + (uint32_t) n * 8u + 15u) & ~15u, vs.bar);
        }
    }
    // the thread's slice of the element division (values 0..3), all loads in flight at once, packed two bits each
    unsigned long long em = 0ull;
#pragma unroll
    for (int k = 0; k < CB_VS_MAXK; k++) { const int i = tid + k * nt; em |= (unsigned long long) ((i < n ? el[i] : 0) & 3) << (2 * k); }
    mbar_wait(vs.bar, vs.par);
    vs.par ^= 1u;
    if (tid == 0) *reinterpret_cast<volatile int *>(cb_smem_window + (vs.bar - vs.a0) + 8u) = (int) vs.par;     // red[115]
    const double *s0 = reinterpret_cast<const double *>(cb_smem_window) + ((((uintptr_t) a0) & 15u) >> 3);
    const double *s1 = reinterpret_cast<const double *>(cb_smem_window + vs.bstride) + ((((uintptr_t) a1) & 15u) >> 3);
    const double *s2 = reinterpret_cast<const double *>(cb_smem_window + 2u * vs.bstride) + ((((uintptr_t) a2) & 15u) >> 3);
#pragma unroll 2
    for (int k = 0; k < CB_VS_MAXK; k++) {
        const int i = tid + k * nt;
        if (i < n) {
            double a[3];
            a[0] = s0[i];
            if (NA > 1) a[1] = s1[i];
            if (NA > 2) a[2] = s2[i];
            f(i, (int) ((em >> (2 * k)) & 3ull), a);
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

// contact box from the bounding box (ix in [x0, x1], iy in [y0, y1]) of the contact area held by every thread's partial
// extremes: block reduction, then the rules of contact_box_dev
__device__ ContactBox contact_box_reduce(const ConvPlan &P, const NormCase &c, int x0, int x1, int y0, int y1, double *red)
{
    const int mx = P.mx;
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

// returns 1 when the reference would abort (MaxCG reached while diverging)
//
// Statement order of normcg (m_solvpn.f90:24-461).  The masked BLAS-1 steps between the three products of an iteration
// are fused: the means of the projections (N=1) come out of the products' last stage, the shift of the search direction
// and the rescaling of the pressures go into their first stage (ConvFuse), and the remaining element-wise work is five
// passes per iteration, each with its reductions.
__device__ __noinline__ int normcg_dev(const ConvPlan &P, const Smem &sm, const NormCase &c, const double *hstot, double &pen,
                          int *el, double *ps, double *wk, int &itcg_out, double &err_out, int &nprod)
{
    const int n = P.npot, tid = threadIdx.x, nt = blockDim.x, mx = P.mx;
    double *__restrict__ rhs = wk, *__restrict__ res = wk + n, *__restrict__ r_prv = wk + 2 * n,
           *__restrict__ dd = wk + 3 * n, *__restrict__ z = wk + 4 * n, *__restrict__ v = wk + 5 * n,
           *__restrict__ q = wk + 6 * n;
    double *red = sm.red;
    const int ic_norm = c.ic_norm, maxcg = c.maxgs;
    const bool n1 = ic_norm == 1;
    const double eps = c.eps, dxdy = c.dxdy, fntrue = c.fntrue;
    const int numinn = n <= 150 ? 3 : (n <= 400 ? 2 : 1);
    const uint32_t mg_mx = div_magic((uint32_t) mx);
    VecStage vs = vec_stage_init(sm, n, c.stage_bytes);
    CB_T_INIT();

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

    CB_T(4);
    ContactBox box = contact_box_dev(P, c, el, red);
    conv_int_dev(P, sm, c, box, 0, ps, res, el); nprod++;             // :173-175 res = rhs - A ps on C
    vec_pass<2>(n, el, rhs, res, nullptr, nullptr, [&](int i, int e, const double *a) { if (e >= 1) res[i] = a[0] - a[1]; });
    __syncthreads();
    if (ic_norm == 1) proj_avg_dev(el, res, n, red);
    CB_T(5);

    double rz1, rz2 = 0.0, rms_xk = 1.0, rms_upd = 2.0 * eps * rms_xk, rms_upd1 = 0.0;
    int itcg = 0, itinn = 0;
    bool lchanged = false;

    while ((lchanged || rms_upd > eps * rms_xk) && itcg < maxcg) {   // :194
        itcg++; itinn++;
        const double rncon = fmax(1.0, (double) ncon);                // number of elements with el >= 1 (gf3 counts them)
        ConvFuse fz = conv_no_fuse(); fz.out_sum = n1 ? 1 : 0;
        const double sz = conv_int_fx_dev(P, sm, c, box, 1, res, z, el, fz); nprod++;      // z = M res on C (+ its sum)
        const double zavg = n1 ? sz / rncon : 0.0;                    // gf3_proj_avg z, :207-209
        CB_T(6);

        // pass A: z -= mean ; (z, res), (z, r_prv)
        double d2[2] = { 0.0, 0.0 };
        staged_pass<3>(vs, n, el, z, res, r_prv, [&](int i, int e, const double *a) {
            if (e >= 1) { const double zz = n1 ? a[0] - zavg : a[0]; if (n1) z[i] = zz; d2[0] += zz * a[1]; d2[1] += zz * a[2]; }
        });
        block_sum<2>(d2, red);
        rz1 = rz2; rz2 = d2[0];
        CB_T(7);

        // pass B: v = z (+ beta v) ; sum of v for its projection
        double sv[1] = { 0.0 };
        if (itcg <= 1 || rz1 < CB_TINY) {                            // :228-241
            staged_pass<1>(vs, n, el, z, nullptr, nullptr, [&](int i, int e, const double *a) { if (e >= 1) { v[i] = a[0]; sv[0] += a[0]; } });
        } else {
            const double beta = fmax(0.0, (rz2 - d2[1]) / fmax(CB_TINY, rz1));
            staged_pass<2>(vs, n, el, z, v, nullptr, [&](int i, int e, const double *a) {
                if (e >= 1) { const double vn = beta * a[1] + a[0]; v[i] = vn; sv[0] += vn; }
            });
        }
        double vavg = 0.0;
        if (n1) { block_sum<1>(sv, red); vavg = sv[0] / rncon; }
        else __syncthreads();
        CB_T(8);

        // q = A v on C: the projection of v goes into the product's first stage, the sum of q comes out of its last
        ConvFuse fq = conv_no_fuse();
        if (n1) { fq.in_mode = 1; fq.in_shift = vavg; fq.out_sum = 1; }
        const double sq = conv_int_fx_dev(P, sm, c, box, 0, v, q, el, fq); nprod++;
        const double qavg = n1 ? sq / rncon : 0.0;
        CB_T(9);

        // pass C: q -= mean ; (res, v), (q, v), (v, v)
        double d4[3] = { 0.0, 0.0, 0.0 };
        staged_pass<3>(vs, n, el, v, res, q, [&](int i, int e, const double *a) {
            if (e >= 1) { const double qq = n1 ? a[2] - qavg : a[2]; if (n1) q[i] = qq; d4[0] += a[1] * a[0]; d4[1] += qq * a[0]; d4[2] += a[0] * a[0]; }
        });
        block_sum<3>(d4, red);
        const double rv = d4[0], vav = d4[1];
        CB_T(10);
        double alpha;
        if (fabs(vav) > 1e-32 && ncon == 1) alpha = rv / vav;
        else alpha = rv / fmax(CB_TINY, vav);
        rms_upd = fabs(alpha) * sqrt(d4[2] / rncon);
        if (itcg == 1) rms_upd1 = rms_upd;
        const bool need_xk = (itcg <= 3 || itcg % 10 == 0);
        const bool outer_known = numinn == 1;                         // itinn < numinn never holds: every iteration is an outer one

        // pass D: r_prv = res ; ps += alpha v ; |ps|^2 ; (outer iteration) negative pressures leave the contact area, :310-318
        double p2[3] = { 0.0, 0.0, 0.0 };                             // |ps|^2, elements released, sum of the kept pressures
        staged_pass<3>(vs, n, el, res, ps, v, [&](int i, int e, const double *a) {
            r_prv[i] = a[0];                                          // :294 (AllElm copy)
            if (e >= 1) {
                const double pi = a[1] + alpha * a[2];
                p2[0] += pi * pi;
                if (outer_known && pi < 0.0) { el[i] = 0; ps[i] = 0.0; p2[1] += 1.0; }
                else { ps[i] = pi; p2[2] += pi; }
            }
        });
        block_sum<3>(p2, red);
        if (need_xk) rms_xk = sqrt(p2[0] / rncon);
        CB_T(11);

        if (!outer_known && itinn < numinn && rms_upd >= eps * rms_xk) {             // :298-303
            vec_pass<2>(n, el, res, q, nullptr, nullptr, [&](int i, int e, const double *a) { if (e >= 1) res[i] = a[0] - alpha * a[1]; });
            __syncthreads();
        } else {
            double nneg = p2[1], spos = p2[2];
            if (!outer_known) {                                       // small grids: the release pass on its own, :310-318
                double k2[2] = { 0.0, 0.0 };
                vec_pass<1>(n, el, ps, nullptr, nullptr, nullptr, [&](int i, int e, const double *a) {
                    if (e >= 1) { if (a[0] < 0.0) { el[i] = 0; ps[i] = 0.0; k2[0] += 1.0; } else k2[1] += a[0]; }
                });
                block_sum<2>(k2, red);
                nneg = k2[0]; spos = k2[1];
            }
            bool lchg_negpn = nneg > 0.0;
            ncon -= (int) nneg;
            bool recount = false;
            if (ncon <= 0) {                                          // :323-333
                double k[1] = { 0.0 };
                for (int i = tid; i < n; i += nt)
                    if (hstot[i] <= hsmin0 + 1e-5) { el[i] = 1; ps[i] = 0.0; k[0] += 1.0; }
                block_sum<1>(k, red);
                ncon += (int) k[0];
                lchg_negpn = true; recount = true;
            }
            ConvFuse fd = conv_no_fuse();
            fd.out_sub = rhs; fd.out_sum = 1;
            if (ic_norm == 1 && lchg_negpn) {                         // :337-344 rescale to the prescribed force
                double s0 = spos;
                if (recount) {
                    double s[1] = { 0.0 };
                    for (int i = tid; i < n; i += nt) s[0] += ps[i];
                    block_sum<1>(s, red);
                    s0 = s[0];
                }
                double fk = dxdy * s0;
                if (fabs(fk) < (double) 1e-3f * fntrue) {
                    for (int i = tid; i < n; i += nt) if (el[i] >= 1) ps[i] = 1.0;
                    __syncthreads();
                    fk = (double) ncon;
                }
                fd.in_mode = 2; fd.in_shift = fntrue / fk;            // ps = f ps on C: in the product's first stage
            }

            CB_T(12);
            // dd = A ps - rhs on all elements (+ its sum over C), :351-359
            const double sd = conv_fx_dev(P, sm, ps, c.chatA, dd, el, 0, 0, 0, mx, P.my, mx, fd); nprod++;
            if (ic_norm == 1) davg = sd / (double) ncon;
            CB_T(13);

            // pass E: residual, elements entering the contact area, bounding box of the new contact area, :361-384
            double ke[1] = { 0.0 };
            int bx0 = mx, bx1 = -1, by0 = P.my, by1 = -1;
            staged_pass<1>(vs, n, el, dd, nullptr, nullptr, [&](int i, int e, const double *a) {
                double di = a[0], r = 0.0;
                bool in = e >= 1;
                if (in) { if (n1) { di -= davg; dd[i] = di; } r = -di; }
                else if (di - davg < 0.0) { el[i] = 1; r = -(di - davg); ke[0] += 1.0; in = true; }
                else v[i] = 0.0;
                res[i] = r;
                if (in) {
                    const int iy = (int) fdiv((uint32_t) i, mg_mx), ix = i - iy * mx;
                    bx0 = min(bx0, ix); bx1 = max(bx1, ix); by0 = min(by0, iy); by1 = max(by1, iy);
                }
            });
            block_sum<1>(ke, red);
            const bool lchg_intpen = ke[0] > 0.0;
            ncon += (int) ke[0];
            itinn = 0;
            lchanged = lchg_intpen || lchg_negpn;
            box = contact_box_reduce(P, c, bx0, bx1, by0, by1, red);
            CB_T(14);
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
    CB_T_INIT();

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

        CB_T(15);
        if (normcg_dev(P, sm, c, hstot, pen, el, ps, wk, it, errpn, nprod)) status |= 1;
        if (threadIdx.x == 0 && blockIdx.x == 0) cb_tl_ = clock64();
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
    CB_T(16);
    if (tid == 0) {
        c.pen = pen;
        if (c.ic_norm == 0) c.fntrue = c.dxdy * s[0];                 // :352
        c.itcg = itcg; c.itnorm = itnorm; c.ncon = (int) s[1]; c.status = status; c.err = errpn; c.nprod = nprod;
    }
}

}  // namespace cb200
