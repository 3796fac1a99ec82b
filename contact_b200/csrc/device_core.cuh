// device_core.cuh -- shared-memory view, fused convolution as a device function, deterministic block reductions.
#pragma once
#include "fftconv_warp.cuh"
#include "fftconv2.cuh"

#ifndef CB_THREADS
#define CB_THREADS 384
#endif

namespace cb200 {

struct Smem {
    cd *S, *W, *twx, *twy;
    unsigned short *posx;
    double *red;       // 128 doubles of reduction scratch
    uint32_t a0;                  // shared-window address of the dynamic shared memory base
};

__device__ __forceinline__ Smem smem_view(const ConvPlan &P, unsigned char *base)
{
    Smem s;
    s.S = reinterpret_cast<cd *>(base + P.off_S);
    s.W = reinterpret_cast<cd *>(base + P.off_W);
    s.twx = reinterpret_cast<cd *>(base + P.off_twx);
    s.twy = reinterpret_cast<cd *>(base + P.off_twy);
    s.posx = reinterpret_cast<unsigned short *>(base + P.off_posx);
    s.red = reinterpret_cast<double *>(base + P.off_red);
    s.a0 = (uint32_t) __cvta_generic_to_shared(base);
    return s;
}

// ---- typed shared-memory accessor of the warp-resident product: element i = 16-byte unit i of the dynamic shared
//      memory window (ld.shared.v2.f64 / st.shared.v2.f64, never a generic load) ----
extern __shared__ __align__(16) unsigned char cb_smem_window[];
struct ShBuf {
    __device__ __forceinline__ cd ld(uint32_t i) const { return reinterpret_cast<const cd *>(cb_smem_window)[i]; }
    __device__ __forceinline__ void st(uint32_t i, cd v) const { reinterpret_cast<cd *>(cb_smem_window)[i] = v; }
    __device__ __forceinline__ cd ldz(uint32_t i, bool ok) const { return ok ? ld(i) : make_double2(0.0, 0.0); }
    __device__ __forceinline__ void stp(uint32_t i, cd v, bool ok) const { if (ok) st(i, v); }
    __device__ __forceinline__ int ldi(uint32_t w) const { return reinterpret_cast<const int *>(cb_smem_window)[w]; }
};

// ---- bulk-async copy (TMA engine, cp.async.bulk) + mbarrier: used to fetch a plan's twiddle tables ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// try_wait suspends the thread for a bounded time per attempt.  A copy that never completes (a byte-count or parity
// bug) traps after ~2 s of waiting instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t tries = 0; !done; tries++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && (tries & 1023u) == 1023u) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ll) __trap();
        }
    }
}

// Control words of the table cache, kept in the reduction scratch (red[112..113] of the kernel's own plan, common to
// the products of all level plans of a launch): one mbarrier, the id of the plan whose tables are resident, the parity
// of the barrier's next phase.
#define CB_HDR_SLOT 112
__device__ __forceinline__ volatile int *conv_hdr(const Smem &s) { return reinterpret_cast<volatile int *>(s.red + CB_HDR_SLOT + 1); }
__device__ __forceinline__ uint32_t conv_hdr_bar(const Smem &s) { return (uint32_t) __cvta_generic_to_shared(s.red + CB_HDR_SLOT); }

// anything else that writes into the S / W window of the plan (the Gauss-Seidel sweep arrays) must forget the tables
__device__ __forceinline__ void conv_tables_invalidate(const Smem &s) { if (threadIdx.x == 0) conv_hdr(s)[0] = -1; }

__device__ __forceinline__ void smem_load_tables(const ConvPlan &P, const Smem &s)
{
    if (threadIdx.x == 0) {
        mbar_init(conv_hdr_bar(s), 1);
        mbar_init(conv_hdr_bar(s) + 16u, 1);                        // staged vector passes (VecStage)
        conv_hdr(s)[0] = -1; conv_hdr(s)[1] = 0;
        *reinterpret_cast<volatile int *>(s.red + CB_HDR_SLOT + 3) = 0;       // parity of the staging barrier's next phase
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// cycle counters of CTA 0 (development aid, read with cb200_conv_prof): [0] products, [1] cycles inside conv_dev,
// [2] cycles of the whole solver kernels
__device__ unsigned long long g_conv_prof[32];
// section timer of CTA 0, thread 0: CB_T(k) adds the cycles since the previous mark to g_conv_prof[k] (slots 4..31: solver sections)
#define CB_T_INIT() long long cb_tl_ = clock64()
#define CB_T(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t__ = clock64(); g_conv_prof[k] += (unsigned long long) (t__ - cb_tl_); cb_tl_ = t__; } } while (0)

// work accounting of ALL CTAs (read with cb200_work_counters): one atomic per product by thread 0 --
// [0] products, [1] nominal flops at the transform size actually used (the contact-box level, not the full grid),
// [2] algorithmic bytes 17 per element of the box (8 in, 8 out, 1 mask: SURVEY 8(d) B_batched at the box size)
__device__ unsigned long long g_work[4];
__device__ __forceinline__ void work_count(const ConvPlan &P, int bw, int bh)
{
    if (threadIdx.x == 0) {
        atomicAdd(&g_work[0], 1ull); atomicAdd(&g_work[1], (unsigned long long) P.nom_flops);
        atomicAdd(&g_work[2], 17ull * (unsigned long long) (bw * bh));
    }
}

#define CB_PHASE(call) do { call; __syncthreads(); } while (0)
#include "conv_sequence.inc"

// The same product restricted to the box (x0, y0, bw x bh) of a grid with row stride `stride`, with the plan P of a
// (smaller) transform size Fx >= bw, Fy >= bh and ITS transformed coefficients: what the reference does for AllInt
// products (bounding box of the contact area, m_aijpj.f90:774-793).  p must vanish outside the box; u is written inside
// the box only.  The plan's tables are (re)loaded into shared memory first -- plans of different sizes share the buffer.
// Warp-resident product (fftconv2.cuh) on the box (x0, y0, bw x bh): rows | columns | rows with one block barrier after
// each; the plan's twiddle tables are fetched by a bulk-async copy when another plan's are resident.
// Optional fused element-wise work (ConvFuse, by value: registers across the call); returns the fused masked sum (or 0).
__device__ __noinline__ double conv2_box_dev(const ConvPlan &P, const Smem &sm, const double *p, const cd *chat, double *u,
                                             const int *el, int mask_mode, int add, int x0, int y0, int bw, int bh, int stride,
                                             const ConvFuse fuse)
{
    const int tid = threadIdx.x, warp = tid >> 5;
    const long long t_in = (tid == 0 && blockIdx.x == 0) ? clock64() : 0;
    const Conv2Plan &c = P.c2;
    volatile int *hdr = conv_hdr(sm);
    if (hdr[0] != c.id) {                                       // uniform over the CTA
        const int par = hdr[1];
        const uint32_t bar = conv_hdr_bar(sm);
        __syncthreads();                                        // all have read the control words
        if (tid == 0) {
            hdr[0] = c.id; hdr[1] = par ^ 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (uint32_t) c.tab_len * 16u);
            bulk_g2s(sm.a0 + (uint32_t) c.off_tab, c.tab, (uint32_t) c.tab_len * 16u, bar);
        }
        mbar_wait(bar, (uint32_t) par);
        if (tid == 0 && blockIdx.x == 0) { g_conv_prof[18] += (unsigned long long) (clock64() - t_in); g_conv_prof[19] += 1; }
    }
    const ShBuf buf;
    const bool f_in = fuse.in_mode != 0, f_sum = fuse.out_sum != 0;
    c2_rows_fwd(P, buf, p + (size_t) y0 * stride + x0, bw, bh, stride, warp,
                f_in ? el + (size_t) y0 * stride + x0 : nullptr, fuse.in_shift, fuse.in_mode);
    __syncthreads();
    c2_cols(P, buf, chat, bh, bh, warp);
    __syncthreads();
    c2_rows_inv(P, buf, u, el, mask_mode, add, x0, y0, bw, bh, stride, warp, fuse.out_sub, f_sum ? 1 : 0);
    __syncthreads();
    double s_ = 0.0;
    if (f_sum) {                                                // per-warp partials, summed in warp order by every thread
        const uint32_t osum = (uint32_t) (c.off_tab / 16 + c.tab_len);
        const int nw = c.nslot;
        for (int w = 0; w < nw; w++) s_ += buf.ld(osum + (uint32_t) w).x;
    }
    work_count(P, bw, bh);                                      // at the end: no register pressure on the passes above
    if (tid == 0 && blockIdx.x == 0) { g_conv_prof[0] += 1; g_conv_prof[1] += (unsigned long long) (clock64() - t_in); }
    return s_;
}

__device__ __forceinline__ ConvFuse conv_no_fuse() { ConvFuse f; f.in_mode = 0; f.in_shift = 0.0; f.out_sub = nullptr; f.out_sum = 0; f.sum = 0.0; return f; }

__device__ __noinline__ void conv_box_dev(const ConvPlan &P, const Smem &sm, const double *p, const cd *chat, double *u,
                                          const int *el, int mask_mode, int add, int x0, int y0, int bw, int bh, int stride)
{
#ifndef CB_NO_CONV2
    if (P.c2.ok) { conv2_box_dev(P, sm, p, chat, u, el, mask_mode, add, x0, y0, bw, bh, stride, conv_no_fuse()); return; }
#endif
    conv_tables_invalidate(sm);                                 // this path overwrites the window of the other one
    work_count(P, bw, bh);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const long long t_in = (tid == 0 && blockIdx.x == 0) ? clock64() : 0;
    typedef MemBuf<cd> CB_BUF;
    const CB_BUF BUF = { reinterpret_cast<cd *>(__cvta_shared_to_generic(sm.a0)) };
    const uint32_t oS = P.off_S / 16, oW = P.off_W / 16;
    {
        cd *tx = reinterpret_cast<cd *>(__cvta_shared_to_generic(sm.a0 + P.off_twx));
        cd *ty = reinterpret_cast<cd *>(__cvta_shared_to_generic(sm.a0 + P.off_twy));
        unsigned short *px = reinterpret_cast<unsigned short *>(__cvta_shared_to_generic(sm.a0 + P.off_posx));
        for (int k = tid; k < 2 * P.Fx; k += nthr) tx[k] = P.twx[k];
        for (int k = tid; k < 2 * P.Fy; k += nthr) ty[k] = P.twy[k];
        for (int k = tid; k < P.Lx; k += nthr) px[k] = P.posx[k];
        __syncthreads();
    }
    const MemBuf<const cd> twx = { reinterpret_cast<const cd *>(__cvta_shared_to_generic(sm.a0 + P.off_twx)) };
    const MemBuf<const cd> twy = { reinterpret_cast<const cd *>(__cvta_shared_to_generic(sm.a0 + P.off_twy)) };
    const MemBuf<const unsigned short> posx = { reinterpret_cast<const unsigned short *>(__cvta_shared_to_generic(sm.a0 + P.off_posx)) };
    const int SY = P.SY;
#ifndef CB_WARP_CONV
    // block-wide phase sequence (default: measured faster inside the solvers)
    RowSrc src;
    src.base = p + (size_t) y0 * stride + x0; src.kind = 0; src.mx = bw; src.my = bh; src.cmx = 0; src.cmy = 0; src.Fx = P.Fx; src.Fy = P.Fy;
    src.row0 = 0; src.stride = stride;
    CB_CONV_FORWARD_ROWS(bh, src);
    CB_CONV_COLUMNS_PRODUCT2(bh, bh, chat);
    CB_CONV_INVERSE_ROWS_STORE(bh, u, el, mask_mode, add, x0, y0, bw, stride);
#else
    // warp-scheduled variant (fftconv_warp.cuh, -DCB_WARP_CONV): a transform never leaves its warp, three block barriers
    // per product; bit-identical results
    const int warp = tid >> 5, nwarps = nthr >> 5;
    warp_rows_fwd(P, BUF, oS, SY, p + (size_t) y0 * stride + x0, bw, bh, stride, twx, posx, warp, nwarps);
    __syncthreads();
    warp_cols(P, BUF, oS, oW, SY, bh, bh, chat, twy, warp, nwarps);
    __syncthreads();
    warp_rows_inv(P, BUF, oS, SY, u, el, mask_mode, add, x0, y0, bw, bh, stride, twx, posx, warp, nwarps);
    __syncthreads();
#endif
    if (tid == 0 && blockIdx.x == 0) { g_conv_prof[0] += 1; g_conv_prof[1] += (unsigned long long) (clock64() - t_in); }
}

// u (masked) = conv(p) with transformed coefficients chat on the full grid of plan P.  All threads of the CTA must call.
__device__ __forceinline__ void conv_dev(const ConvPlan &P, const Smem &sm, const double *p, const cd *chat,
                                         double *u, const int *el, int mask_mode, int add)
{
    conv_box_dev(P, sm, p, chat, u, el, mask_mode, add, 0, 0, P.mx, P.my, P.mx);
}

// ---- deterministic block reductions (fixed shuffle tree; result broadcast to all threads) ----
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double *red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncwarp();                          // shuffles of a warp that is still split after a ragged loop take a slow path
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    __syncthreads();                       // protect red[] from a previous use
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < N; i++) red[wid * N + i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        for (int w = 0; w < nw; w++) s += red[w * N + i];
        v[i] = s;
    }
}

__device__ __forceinline__ double block_min(double v, double *red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double s = red[0];
    for (int w = 1; w < nw; w++) s = fmin(s, red[w]);
    return s;
}

}  // namespace cb200
