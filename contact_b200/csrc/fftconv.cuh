// fftconv.cuh -- the influence-coefficient product u = A p as a pruned, zero-padded 2-D FP64 FFT convolution that
// lives entirely in one CTA's shared memory (sm_100a: up to 227 KB per CTA).
//
// Replaces fft_VecAijPj of the reference (/root/reference/src/m_aijpj.f90:712-1015): pad p(mx,my) into a
// 2Fx x 2Fy array (Fx = opt_fft_size(mx)), real 2-D FFT, pointwise multiply with the transformed coefficients,
// inverse FFT, read the result at offset (Fx,Fy), store where the element-division mask selects.
//
// B200-first formulation (not MKL's):
//   * row pass    : the my non-zero rows only; each real row of length 2Fx is transformed as a packed complex FFT of
//                   length Fx (+ split step) -> S[kx][iy], kx = 0..Fx, batch index iy fastest (bank-conflict free).
//   * column pass : chunks of C columns kx; zero-padded complex FFT of length 2Fy in W[j][c] (c fastest), multiply by
//                   C^ in the last stage's registers, inverse FFT, keep rows Fy..Fy+my-1 only.
//   * inverse rows: Hermitian merge, packed inverse FFT of length Fx, keep columns Fx..Fx+mx-1 only.
//   All FFT stages are in place: decimation in frequency forward, decimation in time backward, so the spectrum
//   stays in digit-reversed order and C^ is simply stored in that same order (built by the same code) -- no
//   reordering pass, no second buffer.  1/(4 Fx Fy) and 1/G are folded into C^.
//
// Every phase is a plain function of (tid, nthr) templated on a buffer accessor, so that the identical code runs
// (a) on shared memory through explicit ld.shared/st.shared (ShBuf), (b) on global scratch for the one-off
// coefficient transforms (MemBuf) and (c) on the host (tests/host_emul) for validation without a GPU.
// Index arithmetic uses multiply-high "magic" division (no integer divide in the hot loops).
#pragma once
#include <stdint.h>
#include "fft_radix.cuh"

namespace cb200 {

#define CB_MAXSTAGE 8

// per-stage constants of a transform of length L at sub-length ns with radix R (host-computed: no integer divisions
// in the device loops)
struct StageK {
    int m;                      // ns / R
    int cnt;                    // L / R   (butterflies per transform)
    int tstep;                  // (L / ns) * twmul   (twiddle index step)
    uint32_t mg_m;              // div_magic(m)
};

// Plan of the warp-resident product (fftconv2.cuh): row transforms Lx = Ax * Bx, column transforms Ly = Ay * By, every
// transform stays inside one warp (private W slot, __syncwarp only); filled by make_plan, ok = 0: not available for
// this size (the block-wide phase sequence of conv_sequence.inc serves it then).
struct Conv2Plan {
    int ok, id;                 // id: unique per plan (tables cached in shared memory between products of the same plan)
    int Ax, Bx, blkx, nux;      // rows: radices, slot block stride (Bx | 1), units per row in the split stage
    int RG, rowlen;             // rows per warp group, slot elements per row (Ax * blkx + padding)
    int Ay, By, blky;           // columns: radices, slot block stride (By | 1)
    int G, ngrp, collen;        // columns per warp group, groups, slot elements per column (Ay * blky + padding)
    int SY;                     // row stride of S[kx][iy] (>= my; padded so that column and row stages hit distinct banks)
    uint32_t mg_G, mg_RG;
    int slot_len, nslot;        // elements per warp slot, warps that own one
    int chat_len;               // ngrp * G * Ly  (cd elements per coefficient block, layout [grp][k2][k1][c])
    int tab_len;                // cd elements of the twiddle tables
    int o_t1x, o_tsx, o_tmx, o_tay;   // offsets (elements) of the tables inside tab
    int o_kr, o_kc;             // offsets (elements) of the stage constants of the row / column stages inside tab
    int off_S, off_W, off_tab;  // byte offsets into dynamic shared memory
    uint32_t mg_Bx, mg_nux, mg_By, mg_Ay;
    const cd *tab;              // device copy of the tables
};

struct ConvPlan {
    int mx, my, npot;
    int Fx, Fy;                 // half sizes of the padded array (opt_fft_size)
    int Lx, Ly;                 // transform lengths: Lx = Fx (packed real rows), Ly = 2 Fy
    int nsx, nsy;               // number of stages
    int rx[CB_MAXSTAGE], ry[CB_MAXSTAGE];   // radices, DIF order
    int SY;                     // batch stride of S (>= number of rows held, odd)
    int C;                      // columns per chunk
    int nchunk;                 // ceil((Fx+1)/C)
    int chat_len;               // nchunk * Ly * C  (cd elements per coefficient block)
    // byte offsets into dynamic shared memory
    int off_S, off_W, off_twx, off_twy, off_posx, off_red, smem_bytes;
    // device tables
    const cd *twx;              // [2Fx] exp(-2 pi i k / 2Fx)
    const cd *twy;              // [2Fy] exp(-2 pi i k / 2Fy)
    const unsigned short *posx; // [Lx] position of frequency k after the DIF stages
    // warp-scheduled product (fftconv_warp.cuh): per-stage constants and the column grouping, all from make_plan
    StageK kx[CB_MAXSTAGE], ky[CB_MAXSTAGE];
    int G;                      // columns per warp group (<= C)
    int gpc;                    // groups per chunk = ceil(C / G)
    int wslots;                 // warps that own a W slot
    int wslot_len;              // cd elements per W slot (padded layout, ViewSk)
    uint32_t mg_G, mg_gpc;      // division magics
    Conv2Plan c2;               // warp-resident product (fftconv2.cuh)
    unsigned int nom_flops;     // nominal flops of ONE product with this plan: 2 * 2.5 N log2 N + 6 S, N = 4 Fx Fy, S = (Fx+1) 2Fy (work accounting)
};

// ---- exact division of small non-negative integers by a run-time constant: q = (n * magic) >> 32 ----
// magic = floor(2^32/d) + 1 (32-bit divide only: 0xFFFFFFFF/d equals floor(2^32/d) unless d is a power of two, where the
// +1 lands exactly on 2^32/d); exact for n*d < 2^32; 0 encodes divide-by-1
CB_HD uint32_t div_magic(uint32_t d) { return d <= 1u ? 0u : 0xFFFFFFFFu / d + 1u; }
CB_HD uint32_t fdiv(uint32_t n, uint32_t magic)
{
#ifdef __CUDA_ARCH__
    return magic ? __umulhi(n, magic) : n;
#else
    return magic ? (uint32_t) (((uint64_t) n * magic) >> 32) : n;
#endif
}

// ---- buffer accessors ----
template <class T> struct MemBuf {          // any memory through a (restrict-free) pointer: global scratch, host, or
    T *p;                                   // shared memory when the pointer was derived with __cvta_shared_to_generic
    CB_HD T ld(uint32_t i) const { return p[i]; }
    CB_HD void st(uint32_t i, T v) const { p[i] = v; }
    CB_HD T ldz(uint32_t i, bool ok) const { return ok ? p[i] : T(); }      // predicated load, zero when off
    CB_HD void stp(uint32_t i, T v, bool ok) const { if (ok) p[i] = v; }    // predicated store
    CB_HD int ldi(uint32_t w) const { return reinterpret_cast<const int *>(p)[w]; }   // 32-bit word w of the window
};

// ---- 2-D views used by the FFT stages: element e of transform c ----
// (templated, inlined views and a compile-time forward/inverse switch measured faster on B200 than one shared
//  run-time body per radix: 83.6k vs 100.5k cycles per 91x91 product, tools/phase_timer.cu, profiles/)
template <class B> struct ViewLin {          // buf[off + e*estride + c]
    B buf; uint32_t off, estride;
    CB_HD cd ld(uint32_t e, uint32_t c) const { return buf.ld(off + e * estride + c); }
    CB_HD void st(uint32_t e, uint32_t c, cd v) const { buf.st(off + e * estride + c, v); }
};
template <class B> struct ViewPad {          // column chunk read straight from S, zero beyond n_in rows (pruned input)
    B buf; uint32_t off, SY, n_in;           // off = oS + row0*SY
    CB_HD cd ld(uint32_t e, uint32_t c) const { return e < n_in ? buf.ld(off + c * SY + e) : make_double2(0.0, 0.0); }
};
template <class B> struct ViewCrop {         // column chunk written straight back to S, rows e0..e0+ne-1 only (pruned output)
    B buf; uint32_t off, SY, e0, ne;
    CB_HD void st(uint32_t e, uint32_t c, cd v) const { const uint32_t r = e - e0; if (r < ne) buf.st(off + c * SY + r, v); }
};

// ------------------------------------------------------------------------------------------------------------
// one FFT stage over a batch of transforms (in place when in and out view the same buffer).
// INV = false: decimation in frequency (butterfly, then twiddle); INV = true: its inverse, decimation in time.
// ------------------------------------------------------------------------------------------------------------
// core loop with all constants supplied: items = cnt * nbatch butterflies, m = ns / R, twiddle step tstep.
// PR = 1: first DIF stage of a transform whose upper half of the input is zero (zero-padded tractions: n_in <= L/2), the
//         inputs x[R/2..R-1] are not loaded and the butterfly leaves their additions out;
// PR = 2: last DIT stage of a transform of which only the upper half of the output is kept (result rows Fy..Fy+my-1): the
//         outputs x[0..R/2-1] are neither completed nor stored.  Both give the results of PR = 0.
template <int R, bool INV, int PR, class VI, class VO, class TW>
CB_HD void fft_stage_core_p(VI in, VO out, int nbatch, int items, int m, int tstep, uint32_t mg_b, uint32_t mg_m, TW tw,
                            int tid, int nthr)
{
    const int ns = m * R;
    constexpr bool half_in = (PR == 1) && DftHalfIn<R, INV>::ok;
    constexpr int NL = half_in ? R / 2 : R;
    constexpr int S0 = (PR == 2 && R % 2 == 0) ? R / 2 : 0;
    for (int w = tid; w < items; w += nthr) {
        const uint32_t g = fdiv(w, mg_b), c = w - g * nbatch;
        const uint32_t blk = fdiv(g, mg_m), j = g - blk * m;
        const uint32_t e0 = blk * ns + j;
        cd x[R];
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] = in.ld(e0 + q * m, c);
        if (INV && m > 1) {
            const int t = tstep * j;
#pragma unroll
            for (int q = 1; q < NL; q++) x[q] = cmulc(x[q], tw.ld(t * q));
        }
        if (half_in) DftHalfIn<R, INV>::run(x); else Dft<R, INV>::run(x);
        if (!INV && m > 1) {
            const int t = tstep * j;
#pragma unroll
            for (int q = (S0 > 1 ? S0 : 1); q < R; q++) x[q] = cmul(x[q], tw.ld(t * q));
        }
#pragma unroll
        for (int q = S0; q < R; q++) out.st(e0 + q * m, c, x[q]);
    }
}

template <int R, bool INV, class VI, class VO, class TW>
CB_HD void fft_stage_core(VI in, VO out, int nbatch, int items, int m, int tstep, uint32_t mg_b, uint32_t mg_m, TW tw,
                          int tid, int nthr)
{
    fft_stage_core_p<R, INV, 0>(in, out, nbatch, items, m, tstep, mg_b, mg_m, tw, tid, nthr);
}

template <int R, bool INV, class VI, class VO, class TW>
CB_HD void fft_stage(VI in, VO out, int nbatch, int L, int ns, TW tw, int twmul, int tid, int nthr)
{
    const int m = ns / R;
    fft_stage_core<R, INV>(in, out, nbatch, (L / R) * nbatch, m, (L / ns) * twmul, div_magic(nbatch), div_magic(m), tw,
                           tid, nthr);
}

#ifdef CB_MAXRADIX8
#define CB_RADIX_BIG(CALL)
#else
#define CB_RADIX_BIG(CALL) case 12: { CALL(12); } break; case 16: { CALL(16); } break;
#endif
#define CB_RADIX_SWITCH(r, CALL)          \
    switch (r) {                           \
    case 2:  { CALL(2); } break;           \
    case 3:  { CALL(3); } break;           \
    case 4:  { CALL(4); } break;           \
    case 5:  { CALL(5); } break;           \
    case 6:  { CALL(6); } break;           \
    case 7:  { CALL(7); } break;           \
    case 8:  { CALL(8); } break;           \
    case 9:  { CALL(9); } break;           \
    CB_RADIX_BIG(CALL)                     \
    default: break;                        \
    }

template <bool INV, class VI, class VO, class TW>
CB_HD void fft_stage_r(int r, VI in, VO out, int nbatch, int L, int ns, TW tw, int twmul, int tid, int nthr)
{
#define CB_CALL_(RR) fft_stage<RR, INV>(in, out, nbatch, L, ns, tw, twmul, tid, nthr)
    CB_RADIX_SWITCH(r, CB_CALL_)
#undef CB_CALL_
}

// last forward stage (m = 1, no twiddles) + pointwise multiply with C^ + first inverse stage, in registers.
// chat is laid out like W: [e][c] with row stride cstride.
template <int R, class VI, class VO>
CB_HD void fft_stage_mid_core(VI in, VO out, int nbatch, int items, uint32_t mg_b, const cd *chat, uint32_t cstride,
                              int tid, int nthr)
{
    for (int w = tid; w < items; w += nthr) {
        const uint32_t g = fdiv(w, mg_b), c = w - g * nbatch;
        const uint32_t e0 = g * R;
        cd x[R], h[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
#ifdef __CUDA_ARCH__
            h[q] = __ldg(reinterpret_cast<const double2 *>(chat + (e0 + q) * cstride + c));
#else
            h[q] = chat[(e0 + q) * cstride + c];
#endif
        }
#pragma unroll
        for (int q = 0; q < R; q++) x[q] = in.ld(e0 + q, c);
        Dft<R, false>::run(x);
#pragma unroll
        for (int q = 0; q < R; q++) x[q] = cmul(x[q], h[q]);
        Dft<R, true>::run(x);
#pragma unroll
        for (int q = 0; q < R; q++) out.st(e0 + q, c, x[q]);
    }
}

template <int R, class VI, class VO>
CB_HD void fft_stage_mid(VI in, VO out, int nbatch, int L, const cd *chat, uint32_t cstride, int tid, int nthr)
{
    fft_stage_mid_core<R>(in, out, nbatch, (L / R) * nbatch, div_magic(nbatch), chat, cstride, tid, nthr);
}

template <class VI, class VO>
CB_HD void fft_stage_mid_r(int r, VI in, VO out, int nbatch, int L, const cd *chat, uint32_t cstride, int tid, int nthr)
{
#define CB_CALL_(RR) fft_stage_mid<RR>(in, out, nbatch, L, chat, cstride, tid, nthr)
    CB_RADIX_SWITCH(r, CB_CALL_)
#undef CB_CALL_
}

// the same stages with the per-stage constants of the plan (StageK) and a ready-made magic for nbatch
template <bool INV, class VI, class VO, class TW>
CB_HD void fft_stage_k(int r, VI in, VO out, int nbatch, uint32_t mg_b, const StageK &k, TW tw, int tid, int nthr)
{
#define CB_CALL_(RR) fft_stage_core<RR, INV>(in, out, nbatch, k.cnt * nbatch, k.m, k.tstep, mg_b, k.mg_m, tw, tid, nthr)
    CB_RADIX_SWITCH(r, CB_CALL_)
#undef CB_CALL_
}

// pruned first (PR = 1, forward) / last (PR = 2, inverse) stage of the column transforms of a product
template <bool INV, int PR, class VI, class VO, class TW>
CB_HD void fft_stage_kp(int r, VI in, VO out, int nbatch, uint32_t mg_b, const StageK &k, TW tw, int tid, int nthr)
{
#define CB_CALL_(RR) fft_stage_core_p<RR, INV, PR>(in, out, nbatch, k.cnt * nbatch, k.m, k.tstep, mg_b, k.mg_m, tw, tid, nthr)
    CB_RADIX_SWITCH(r, CB_CALL_)
#undef CB_CALL_
}

template <class VI, class VO>
CB_HD void fft_stage_mid_k(int r, VI in, VO out, int nbatch, uint32_t mg_b, const StageK &k, const cd *chat,
                           uint32_t cstride, int tid, int nthr)
{
#define CB_CALL_(RR) fft_stage_mid_core<RR>(in, out, nbatch, k.cnt * nbatch, mg_b, chat, cstride, tid, nthr)
    CB_RADIX_SWITCH(r, CB_CALL_)
#undef CB_CALL_
}

// ------------------------------------------------------------------------------------------------------------
// row pass pieces
// ------------------------------------------------------------------------------------------------------------

// Source of real rows: either a traction column p(mx,my) placed at the low corner (m_aijpj.f90:932-939), or an
// influence-coefficient block cf(-cmx:cmx-1,-cmy:cmy-1) placed with cf(0,0) at (Fx,Fy) (m_aijpj.f90:894-901).
struct RowSrc {
    const double *base;
    int kind;            // 0: tractions, 1: coefficients
    int mx, my;          // tractions: grid size ; coefficients: limits min(F, m) used by the reference copy loop
    int cmx, cmy;        // coefficients: allocated half sizes of cf
    int Fx, Fy;
    int row0;            // first padded row handled in this batch
    int stride;          // tractions: row stride of the source array (0: = mx); lets a sub-box of a larger grid be the input
};

CB_HD double rowsrc_get(const RowSrc &s, int row, int col)
{
    if (s.kind == 0) {
        return (row < s.my && col < s.mx) ? s.base[(size_t) row * (s.stride ? s.stride : s.mx) + col] : 0.0;
    } else {
        const int iy = row + s.row0 - s.Fy, ix = col - s.Fx;
        if (iy < -s.my || iy >= s.my || ix < -s.mx || ix >= s.mx) return 0.0;
        return s.base[(size_t) (iy + s.cmy) * (2 * s.cmx) + (ix + s.cmx)];
    }
}

// S[j][b] = x[b][2j] + i x[b][2j+1]  for j < Lx, b < nbatch      (S = buf + oS)
// The source lives in global memory: the loads of CB_RLB items are issued back to back before the first store, so a
// thread has CB_RLB loads in flight instead of one (the loop was bound by the L2/HBM latency: 9.6k of the 88k cycles of
// a 91x91 product, tools/phase_timer.cu).
#ifndef CB_RLB
#define CB_RLB 8
#endif
template <class B>
CB_HD void row_load(const ConvPlan &P, B buf, uint32_t oS, int SY, int nbatch, const RowSrc &src, int tid, int nthr)
{
    const int items = P.Lx * nbatch;
    const uint32_t mg = div_magic(P.Lx);
    for (int w0 = tid; w0 < items; w0 += CB_RLB * nthr) {
        double re[CB_RLB], im[CB_RLB];
#pragma unroll
        for (int k = 0; k < CB_RLB; k++) {
            const int w = w0 + k * nthr;
            re[k] = 0.0; im[k] = 0.0;
            if (w < items) {
                const uint32_t b = fdiv(w, mg), j = w - b * P.Lx;
                re[k] = rowsrc_get(src, b, 2 * j); im[k] = rowsrc_get(src, b, 2 * j + 1);
            }
        }
#pragma unroll
        for (int k = 0; k < CB_RLB; k++) {
            const int w = w0 + k * nthr;
            if (w < items) {
                const uint32_t b = fdiv(w, mg), j = w - b * P.Lx;
                buf.st(oS + j * SY + b, make_double2(re[k], im[k]));
            }
        }
    }
}

// ---- fused row passes ----
// The tractions fill at most the lower half of a padded row (mx <= Fx), so in the first DIF stage of the packed row
// transform the inputs x[R/2..R-1] of every butterfly are zero, and of the last DIT stage of the inverse only the outputs
// x[R/2..R-1] (columns Fx..2Fx-1) are read.  For an even first radix the staging pass through S (row_load) and the separate
// store pass (row_store_box) therefore merge with those stages: the butterfly reads its R/2 non-zero inputs straight from
// the traction row and writes its R/2 wanted outputs straight to u.  Items run with the column index j fastest, so global
// accesses are contiguous per row.  Same operations on the non-zero terms => same results as the unfused passes.
CB_HD bool rows_fusable(const ConvPlan &P)
{
    const int r = P.nsx > 0 ? P.rx[0] : 0;
    return r == 4 || r == 8 || r == 12 || r == 16;
}

template <int R, class B, class TW>
CB_HD void row_first_stage(const ConvPlan &P, B buf, uint32_t oS, int SY, int nbatch, const RowSrc &src, TW tw,
                           int tid, int nthr)
{
    const StageK &k = P.kx[0];
    const int m = k.m, items = m * nbatch;
    const int rs = src.stride ? src.stride : src.mx;
    for (int w = tid; w < items; w += nthr) {
        const uint32_t b = fdiv(w, k.mg_m), j = w - b * m;
        const double *row = src.base + (size_t) b * rs;
        cd x[R];
#pragma unroll
        for (int q = 0; q < R / 2; q++) {
            const int col = 2 * (int) (j + q * m);
            x[q] = make_double2(col < src.mx ? row[col] : 0.0, col + 1 < src.mx ? row[col + 1] : 0.0);
        }
        DftHalfIn<R, false>::run(x);
        if (m > 1) {
            const int t = k.tstep * j;
#pragma unroll
            for (int q = 1; q < R; q++) x[q] = cmul(x[q], tw.ld(t * q));
        }
#pragma unroll
        for (int q = 0; q < R; q++) buf.st(oS + (j + q * m) * SY + b, x[q]);
    }
}

template <class B, class TW>
CB_HD void row_first_stage_r(const ConvPlan &P, B buf, uint32_t oS, int SY, int nbatch, const RowSrc &src, TW tw,
                             int tid, int nthr)
{
    switch (P.rx[0]) {
    case 4:  row_first_stage<4>(P, buf, oS, SY, nbatch, src, tw, tid, nthr); break;
    case 8:  row_first_stage<8>(P, buf, oS, SY, nbatch, src, tw, tid, nthr); break;
#ifndef CB_MAXRADIX8
    case 12: row_first_stage<12>(P, buf, oS, SY, nbatch, src, tw, tid, nthr); break;
    case 16: row_first_stage<16>(P, buf, oS, SY, nbatch, src, tw, tid, nthr); break;
#endif
    default: break;
    }
}

// last DIT stage of the inverse row transforms + masked store of the box (x0, y0, bw x nbatch) of u (row stride `stride`);
// semantics of row_store_box
template <int R, class B, class TW>
CB_HD void row_last_stage_store(const ConvPlan &P, B buf, uint32_t oS, int SY, int nbatch, TW tw, double *u, const int *el,
                                int mask_mode, int add, int x0, int y0, int bw, int stride, int tid, int nthr)
{
    const StageK &k = P.kx[0];
    const int m = k.m, items = m * nbatch;
    for (int w = tid; w < items; w += nthr) {
        const uint32_t b = fdiv(w, k.mg_m), j = w - b * m;
        cd x[R];
#pragma unroll
        for (int q = 0; q < R; q++) x[q] = buf.ld(oS + (j + q * m) * SY + b);
        if (m > 1) {
            const int t = k.tstep * j;
#pragma unroll
            for (int q = 1; q < R; q++) x[q] = cmulc(x[q], tw.ld(t * q));
        }
        Dft<R, true>::run(x);
        const size_t r0 = (size_t) (y0 + b) * stride + x0;
#pragma unroll
        for (int q = R / 2; q < R; q++) {
            const int ix = 2 * (int) (j + q * m) - P.Fx;           // real column 2e -> grid column ix (Fx is even here)
            if (ix < bw && !(mask_mode == 1 && el[r0 + ix] < 1)) u[r0 + ix] = add ? u[r0 + ix] + x[q].x : x[q].x;
            if (ix + 1 < bw && !(mask_mode == 1 && el[r0 + ix + 1] < 1)) u[r0 + ix + 1] = add ? u[r0 + ix + 1] + x[q].y : x[q].y;
        }
    }
}

template <class B, class TW>
CB_HD void row_last_stage_store_r(const ConvPlan &P, B buf, uint32_t oS, int SY, int nbatch, TW tw, double *u, const int *el,
                                  int mask_mode, int add, int x0, int y0, int bw, int stride, int tid, int nthr)
{
    switch (P.rx[0]) {
    case 4:  row_last_stage_store<4>(P, buf, oS, SY, nbatch, tw, u, el, mask_mode, add, x0, y0, bw, stride, tid, nthr); break;
    case 8:  row_last_stage_store<8>(P, buf, oS, SY, nbatch, tw, u, el, mask_mode, add, x0, y0, bw, stride, tid, nthr); break;
#ifndef CB_MAXRADIX8
    case 12: row_last_stage_store<12>(P, buf, oS, SY, nbatch, tw, u, el, mask_mode, add, x0, y0, bw, stride, tid, nthr); break;
    case 16: row_last_stage_store<16>(P, buf, oS, SY, nbatch, tw, u, el, mask_mode, add, x0, y0, bw, stride, tid, nthr); break;
#endif
    default: break;
    }
}

// split step of the packed real transform: Z (scrambled) -> X[k], k = 0..Lx, X[k] at row posx[k] (k<Lx), X[Lx] at row Lx
// items per thread whose loads are issued together in the split / merge steps (measured on B200: 1 is best -- with 4 or
// 8 the index arrays go to local memory and the step gets 4x slower, gpurun_out/pt_w5: 3.5k -> 14.6k cycles)
#ifndef CB_SMB
#define CB_SMB 1
#endif
template <class B, class TW, class PX>
CB_HD void row_split(const ConvPlan &P, B buf, uint32_t oS, int SY, int nbatch, TW twx, PX posx, int tid, int nthr)
{
    const int L = P.Lx;
    // k = 0 (-> X[0], X[L]) and the self-paired k = L/2
    for (int b = tid; b < nbatch; b += nthr) {
        cd a = buf.ld(oS + b);
        buf.st(oS + b, make_double2(a.x + a.y, 0.0));
        buf.st(oS + (uint32_t) L * SY + b, make_double2(a.x - a.y, 0.0));
        if (L > 1 && (L & 1) == 0) {
            const uint32_t ia = oS + (uint32_t) posx.ld(L / 2) * SY + b;
            buf.st(ia, cconj(buf.ld(ia)));
        }
    }
    // pairs (k, L-k), k = 1 .. (L-1)/2: the loads of CB_SMB items are issued before the first store
    const int npair = (L - 1) / 2;
    const int items = npair * nbatch;
    const uint32_t mg = div_magic(nbatch);
    for (int w0 = tid; w0 < items; w0 += CB_SMB * nthr) {
        cd a[CB_SMB], bb[CB_SMB], tw[CB_SMB];
        uint32_t ia[CB_SMB], ib[CB_SMB];
#pragma unroll
        for (int i = 0; i < CB_SMB; i++) {
            const int w = w0 + i * nthr;
            if (w < items) {
                const uint32_t k = fdiv(w, mg) + 1, b = w - (k - 1) * nbatch;
                ia[i] = oS + (uint32_t) posx.ld(k) * SY + b; ib[i] = oS + (uint32_t) posx.ld(L - k) * SY + b;
                tw[i] = twx.ld(k);
            }
        }
#pragma unroll
        for (int i = 0; i < CB_SMB; i++)
            if (w0 + i * nthr < items) { a[i] = buf.ld(ia[i]); bb[i] = buf.ld(ib[i]); }
#pragma unroll
        for (int i = 0; i < CB_SMB; i++) {
            if (w0 + i * nthr < items) {
                cd e = make_double2(0.5 * (a[i].x + bb[i].x), 0.5 * (a[i].y - bb[i].y));          // (a + conj b)/2
                cd d = make_double2(0.5 * (a[i].x - bb[i].x), 0.5 * (a[i].y + bb[i].y));          // (a - conj b)/2
                cd t = cmul(tw[i], make_double2(d.y, -d.x));                                 // w^k * (-i) d
                buf.st(ia[i], cadd(e, t));
                buf.st(ib[i], cconj(csub(e, t)));
            }
        }
    }
}

// merge step of the inverse packed real transform: X -> Z' = 2 Z (scrambled positions)
template <class B, class TW, class PX>
CB_HD void row_merge(const ConvPlan &P, B buf, uint32_t oS, int SY, int nbatch, TW twx, PX posx, int tid, int nthr)
{
    const int L = P.Lx;
    for (int b = tid; b < nbatch; b += nthr) {
        cd p = buf.ld(oS + b), q = buf.ld(oS + (uint32_t) L * SY + b);
        cd u = make_double2(p.x + q.x, p.y - q.y), d = make_double2(p.x - q.x, p.y + q.y);
        buf.st(oS + b, make_double2(u.x - d.y, u.y + d.x));                      // u + i d
        if (L > 1 && (L & 1) == 0) {
            const uint32_t ia = oS + (uint32_t) posx.ld(L / 2) * SY + b;
            cd a = buf.ld(ia);
            buf.st(ia, make_double2(2.0 * a.x, -2.0 * a.y));
        }
    }
    const int npair = (L - 1) / 2;
    const int items = npair * nbatch;
    const uint32_t mg = div_magic(nbatch);
    for (int w0 = tid; w0 < items; w0 += CB_SMB * nthr) {
        cd p[CB_SMB], q[CB_SMB], tw[CB_SMB];
        uint32_t ia[CB_SMB], ib[CB_SMB];
#pragma unroll
        for (int i = 0; i < CB_SMB; i++) {
            const int w = w0 + i * nthr;
            if (w < items) {
                const uint32_t k = fdiv(w, mg) + 1, b = w - (k - 1) * nbatch;
                ia[i] = oS + (uint32_t) posx.ld(k) * SY + b; ib[i] = oS + (uint32_t) posx.ld(L - k) * SY + b;
                tw[i] = twx.ld(k);
            }
        }
#pragma unroll
        for (int i = 0; i < CB_SMB; i++)
            if (w0 + i * nthr < items) { p[i] = buf.ld(ia[i]); q[i] = buf.ld(ib[i]); }
#pragma unroll
        for (int i = 0; i < CB_SMB; i++) {
            if (w0 + i * nthr < items) {
                cd u = make_double2(p[i].x + q[i].x, p[i].y - q[i].y);                       // p + conj q
                cd d = make_double2(p[i].x - q[i].x, p[i].y + q[i].y);                       // p - conj q
                cd v = cmulc(make_double2(-d.y, d.x), tw[i]);                            // i d conj(w^k)
                buf.st(ia[i], cadd(u, v));
                buf.st(ib[i], cconj(csub(u, v)));
            }
        }
    }
}

// store the wanted part of the result: u(ix,iy) = x[iy][Fx+ix], ix < mx  (m_aijpj.f90:978-1005)
// mask_mode 0: all elements (AllElm), 1: el >= Adhes (AllInt).  add: u += result.
template <class B>
CB_HD void row_store(const ConvPlan &P, B buf, uint32_t oS, int SY, double *u, const int *el, int mask_mode, int add,
                     int tid, int nthr)
{
    const uint32_t mg = div_magic(P.mx);
    for (int ii = tid; ii < P.npot; ii += nthr) {
        const uint32_t iy = fdiv(ii, mg), ix = ii - iy * P.mx;
        if (mask_mode == 1 && el[ii] < 1) continue;
        const uint32_t xi = P.Fx + ix;
        const cd z = buf.ld(oS + (xi >> 1) * SY + iy);
        const double v = (xi & 1) ? z.y : z.x;
        u[ii] = add ? u[ii] + v : v;
    }
}

// the same for a sub-box (x0, y0, bw x bh) of a grid with row stride `stride`: element (ix, iy) of the box is element
// (x0+ix, y0+iy) of u / el.  Used when the product is restricted to the bounding box of the contact area
// (m_aijpj.f90:774-793).
template <class B>
CB_HD void row_store_box(const ConvPlan &P, B buf, uint32_t oS, int SY, double *u, const int *el, int mask_mode, int add,
                         int x0, int y0, int bw, int bh, int stride, int tid, int nthr)
{
    const uint32_t mg = div_magic(bw);
    const int nbox = bw * bh;
    for (int w = tid; w < nbox; w += nthr) {
        const uint32_t iy = fdiv(w, mg), ix = w - iy * bw;
        const size_t ii = (size_t) (y0 + iy) * stride + x0 + ix;
        if (mask_mode == 1 && el[ii] < 1) continue;
        const uint32_t xi = P.Fx + ix;
        const cd z = buf.ld(oS + (xi >> 1) * SY + iy);
        const double v = (xi & 1) ? z.y : z.x;
        u[ii] = add ? u[ii] + v : v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// coefficient builder pieces.  Chunk ch holds S rows r = ch*C + c, c < C (rows beyond Fx are treated as zero).
// ------------------------------------------------------------------------------------------------------------
template <class B>
CB_HD void col_load(const ConvPlan &P, B buf, uint32_t oS, int SY, int n_in, uint32_t oW, int ch, int tid, int nthr)
{
    const int items = P.Ly * P.C;
    const uint32_t mg = div_magic(P.C);
    for (int w = tid; w < items; w += nthr) {
        const uint32_t j = fdiv(w, mg), c = w - j * P.C;
        const uint32_t r = ch * P.C + c;
        buf.st(oW + w, ((int) j < n_in && (int) r <= P.Fx) ? buf.ld(oS + r * SY + j) : make_double2(0.0, 0.0));
    }
}

// dump the fully forward-transformed chunk, scaled, as C^
template <class B>
CB_HD void col_dump(const ConvPlan &P, B buf, uint32_t oW, cd *chat, int ch, double scale, int tid, int nthr)
{
    const int items = P.Ly * P.C;
    for (int w = tid; w < items; w += nthr)
        chat[(size_t) ch * items + w] = cscale(buf.ld(oW + w), scale);
}

}  // namespace cb200
