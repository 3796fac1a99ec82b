// fftconv2.cuh -- warp-resident form of the fused influence product u = A p (FP64, one CTA, sm_100a).
//
// Same mathematics as fftconv.cuh (pruned, zero-padded 2-D real FFT convolution replacing fft_VecAijPj of the reference,
// /root/reference/src/m_aijpj.f90:712-1015), re-laid for the SM:
//   * a transform never leaves its warp.  Row transforms Lx = Ax * Bx and column transforms Ly = Ay * By are two
//     register butterflies with ONE exchange through a warp-private slot of shared memory; only __syncwarp() separates
//     the stages, so the 12 warps of a CTA drift apart and the loads/stores of one overlap the FP64 chains of the
//     others.  Three block barriers per product remain (rows | columns | rows).
//   * rows forward : stage 1 = input-pruned radix-Ax butterflies read straight from the traction rows (the upper half
//     of a padded row is zero), twiddle, slot; stage 2 = two radix-Bx butterflies per "unit" (blocks k1 and Ax - k1) whose
//     outputs are exactly the pairs (k, Lx - k) of the packed-real split step, which is therefore done in registers;
//     the half spectrum lands in S[kx][iy] in natural order (no digit reversal, no position table).
//   * columns      : a warp owns G spectrum columns: stage A = input-pruned radix-Ay (rows beyond the grid are zero),
//     stage M = radix-By forward, multiply by C^ (read coalesced as [group][k2][column][k1]), radix-By inverse,
//     stage C = output-pruned radix-Ay inverse (only rows Fy..Fy+my-1 are kept) back into S.
//   * rows inverse : mirror image: merge step + radix-Bx inverse per unit, then radix-Ax inverse whose outputs are the
//     wanted columns Fx..Fx+mx-1, written masked to u.
//   * twiddles come from per-stage tables laid out [q][j] (lanes read consecutive entries: no bank conflicts), fetched
//     per plan by one bulk-async copy (cp.async.bulk + mbarrier) and kept while the plan does not change.
//   * all shared-memory traffic is typed (ld.shared / st.shared, 16 B per lane); slot blocks are padded to an odd stride.
// Every stage is a plain function of the lane id templated on a buffer accessor, so the identical code is stepped through
// on the host by tests/host_emul (lanes looped) and checked against the oracle's direct sum without a GPU.
#pragma once
#include "fftconv.cuh"
#include "fft_radix2.cuh"

namespace cb200 {

#ifdef __CUDA_ARCH__
#define CB2_LANES(call) do { { const int lane = (int) (threadIdx.x & 31u); call; } __syncwarp(); } while (0)
#else
// host emulation: the lanes of a warp are stepped one after the other; tests/host_emul hooks a shared-memory
// wavefront model in through CB2_LANE_BEGIN / CB2_LANES_END
#ifndef CB2_LANE_BEGIN
#define CB2_LANE_BEGIN(lane)
#define CB2_LANES_END()
#endif
#define CB2_LANES(call) do { for (int lane = 0; lane < 32; lane++) { CB2_LANE_BEGIN(lane); call; } CB2_LANES_END(); } while (0)
#endif

// Start-up stagger of the passes: the warps leave a block barrier together and would run load -> FP64 -> store in lock
// step (the LSU and the FP64 pipe busy in turns); warp w of SM sub-partition w % 4 starts (w / 4) * CB2_STAGGER cycles
// late, so that on every sub-partition one warp computes while the others move data.
#ifndef CB2_STAGGER
#define CB2_STAGGER 0
#endif
#if defined(__CUDA_ARCH__) && CB2_STAGGER > 0
#define CB2_STAGGER_WAIT(warp) do { const long long t0_ = clock64(), d_ = (long long) ((warp) >> 2) * CB2_STAGGER; while (clock64() - t0_ < d_) { } } while (0)
#else
#define CB2_STAGGER_WAIT(warp)
#endif

// development hook (tools/conv2_bench.cu): per-stage cycle counters of one warp; empty in the product
#ifndef CB2_TICK
#define CB2_TICK(i)
#endif
#ifndef CB2_TICK_INIT
#define CB2_TICK_INIT()
#endif

CB_HD bool c2_radix_a(int r) { return (r >= 2 && r <= 10) || r == 12 || r == 16 || r == 18; }
CB_HD bool c2_radix_b(int r) { return (r >= 1 && r <= 10) || r == 12 || r == 16; }

// The constants of the stage functions.  They are plan-static, so the host stores them behind the twiddle tables (rows
// block, columns block) and they arrive in shared memory with the same bulk copy; every stage reads the few it needs at
// its top (LDS, ~30 cycles) into registers.  Nothing of this is live across stages, and nothing is re-read from the plan
// (a generic reference: every shared-memory store would force a reload) or from a stack copy (local memory = L2 trips).
struct C2K {
    int A, B, blk, nu, len;       // radices of the two stages, slot block stride, units per row, slot elements per transform
    uint32_t mg_B, mg_nu, mg_A, mg_G, mg_RG;
    int RG;
    uint32_t o_t1, o_ts, o_tm;    // table offsets (elements of the shared window)
    int L, SY, Fx, G;
    int pad_[2];                  // 20 words = 5 x 16 bytes
};
#define CB2_KWORDS 20
CB_HD C2K c2k_rows(const ConvPlan &P)
{
    const Conv2Plan &c = P.c2;
    const uint32_t otab = (uint32_t) c.off_tab / 16;
    C2K k;
    k.A = c.Ax; k.B = c.Bx; k.blk = c.blkx; k.nu = c.nux; k.len = c.rowlen;
    k.mg_B = c.mg_Bx; k.mg_nu = c.mg_nux; k.mg_A = 0;
    k.o_t1 = otab + c.o_t1x; k.o_ts = otab + c.o_tsx; k.o_tm = otab + c.o_tmx;
    k.L = P.Lx; k.SY = c.SY; k.Fx = P.Fx; k.G = 0; k.mg_G = 0; k.mg_RG = c.mg_RG; k.RG = c.RG;
    k.pad_[0] = k.pad_[1] = 0;
    return k;
}
CB_HD C2K c2k_cols(const ConvPlan &P)
{
    const Conv2Plan &c = P.c2;
    const uint32_t otab = (uint32_t) c.off_tab / 16;
    C2K k;
    k.A = c.Ay; k.B = c.By; k.blk = c.blky; k.nu = 0; k.len = c.collen;
    k.mg_B = c.mg_By; k.mg_nu = 0; k.mg_A = c.mg_Ay;
    k.o_t1 = otab + c.o_tay; k.o_ts = 0; k.o_tm = 0;
    k.L = P.Ly; k.SY = c.SY; k.Fx = P.Fx; k.G = c.G; k.mg_G = c.mg_G; k.mg_RG = 0; k.RG = 0;
    k.pad_[0] = k.pad_[1] = 0;
    return k;
}
// fetch the block at element offset ok of the shared window (the compiler drops the loads of unused fields)
template <class B> CB_HD C2K c2k_load(B buf, uint32_t ok)
{
    C2K k;
    const uint32_t w = ok * 4;
    k.A = buf.ldi(w + 0); k.B = buf.ldi(w + 1); k.blk = buf.ldi(w + 2); k.nu = buf.ldi(w + 3); k.len = buf.ldi(w + 4);
    k.mg_B = (uint32_t) buf.ldi(w + 5); k.mg_nu = (uint32_t) buf.ldi(w + 6); k.mg_A = (uint32_t) buf.ldi(w + 7);
    k.mg_G = (uint32_t) buf.ldi(w + 8); k.mg_RG = (uint32_t) buf.ldi(w + 9); k.RG = buf.ldi(w + 10);
    k.o_t1 = (uint32_t) buf.ldi(w + 11); k.o_ts = (uint32_t) buf.ldi(w + 12); k.o_tm = (uint32_t) buf.ldi(w + 13);
    k.L = buf.ldi(w + 14); k.SY = buf.ldi(w + 15); k.Fx = buf.ldi(w + 16); k.G = buf.ldi(w + 17);
    k.pad_[0] = k.pad_[1] = 0;
    return k;
}

// Read-only global load of one complex coefficient that stays where it is written: the compiler otherwise sinks these
// loads down to their first use (after the forward butterfly), one L2 round trip per coefficient instead of one per stage.
CB_HD cd c2_ldg_early(const cd *p)
{
#ifdef __CUDA_ARCH__
    cd v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}

// split of the packed real transform for the pair (k, L-k): a = Z[k], b = Z[L-k], w = exp(-i pi k / L)
//   X[k] = (a + conj b)/2 + w (-i)(a - conj b)/2 ,   X[L-k] = conj((a + conj b)/2 - w (-i)(a - conj b)/2)
CB_HD void c2_split_pair(cd a, cd b, cd w, cd &xk, cd &xlk)
{
    const cd e = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
    const cd d = make_double2(0.5 * (a.x - b.x), 0.5 * (a.y + b.y));
    const cd t = cmul(w, make_double2(d.y, -d.x));
    xk = cadd(e, t);
    xlk = cconj(csub(e, t));
}
// merge (inverse of the split up to the factor 2): p = X[k], q = X[L-k]  ->  Z'[k] = 2 Z[k], Z'[L-k] = 2 Z[L-k]
CB_HD void c2_merge_pair(cd p, cd q, cd w, cd &zk, cd &zlk)
{
    const cd u = make_double2(p.x + q.x, p.y - q.y);
    const cd d = make_double2(p.x - q.x, p.y + q.y);
    const cd v = cmulc(make_double2(-d.y, d.x), w);
    zk = cadd(u, v);
    zlk = cconj(csub(u, v));
}

// ------------------------------------------------------------------------------------------------------------
// Element-wise work fused into the first and the last stage of a product (the solvers' masked BLAS-1 passes that touch
// exactly what the product reads or writes, m_gridfunc.f90:936-1591): every input element is loaded by exactly one lane
// of the first row stage and every output element is produced by exactly one lane of the last one.
//   input : on the elements with el_i >= 1, written back: p_i := p_i - in_shift (in_mode 1, gf3_proj_avg of the search
//           direction) or p_i := in_shift * p_i (in_mode 2, rescaling the pressures to the prescribed force);
//   output: u_i := u_i - out_sub[i] (a right-hand side), and the sum of the new u_i over the elements with el_i >= 1
//           (the mean that the next projection needs) -- reduced per warp by a fixed shuffle tree, per CTA in warp order.
struct ConvFuse {
    int in_mode;             // 0: none, 1: shift, 2: scale the masked input and store it back
    double in_shift;
    const double *out_sub;   // null: nothing subtracted
    int out_sum;             // 1: return the masked sum of the output in `sum`
    double sum;
};

#ifdef __CUDA_ARCH__
#define CB2_NL 1
#define CB2_LI(lane) 0
#else
#define CB2_NL 32
#define CB2_LI(lane) (lane)
#endif

// Local memory is poison here: with 227 KB of the SM given to shared memory, L1 keeps ~24 KB, a local word of 384
// threads is 12 cache lines, so every spill or stack access is an L2 round trip (measured: plan fields read from a stack
// copy cost 15 %, 2 KB of spills 17 %).  Hence: plan constants by value in registers (C2K), stage functions inlined into
// ONE function whose spill count is checked at build time (ptxas -v: conv2_box_dev must report 0 bytes), loads issued in
// batches that fit the 168 registers of a 384-thread CTA.
// ------------------------------------------------------------------------------------------------------------
// rows, forward
// ------------------------------------------------------------------------------------------------------------
// Source of the real rows: tractions (box of a grid, row stride `stride`, zero beyond bw columns) -- the coefficient
// transforms are built by dense DFTs (k_chat2_*), not through this path.
//
// stage 1: item (r, j), j < Bx: x[q] = z[j + Bx q] = (row[2n], row[2n+1]); radix-Ax butterfly; y[k1] *= w_Lx^(j k1);
//          slot[r][k1 * blkx + j]
template <int A, class B>
CB_HD void c2_rowf1(uint32_t ok, B buf, uint32_t oslot, const double *src, int nrows, int bw, int stride, int lane,
                    const int *el_src = nullptr, double shift = 0.0, int in_mode = 1)
{
    const C2K k = c2k_load(buf, ok);
    const int Bx = k.B, items = nrows * Bx;
    constexpr bool half = DftHalfIn<A, false>::ok;
    constexpr int NL = half ? A / 2 : A;
    for (int i = lane; i < items; i += 32) {
        const uint32_t r = fdiv((uint32_t) i, k.mg_B), j = (uint32_t) i - r * Bx;
        const double *row = src + (size_t) r * stride;
        cd x[A], tw[A > 1 ? A - 1 : 1];
#pragma unroll
        for (int q = 0; q < NL; q++) {
            const int col = 2 * (int) (j + q * Bx);
            x[q] = make_double2(col < bw ? row[col] : 0.0, col + 1 < bw ? row[col + 1] : 0.0);
        }
        if (el_src != nullptr) {                       // fused input pass: masked shift, written back (ConvFuse::in_mode)
            const int *erow = el_src + (size_t) r * stride;
            double *wrow = const_cast<double *>(row);
#pragma unroll
            for (int q = 0; q < NL; q++) {
                const int col = 2 * (int) (j + q * Bx);
                if (col < bw && erow[col] >= 1) { x[q].x = in_mode == 2 ? shift * x[q].x : x[q].x - shift; wrow[col] = x[q].x; }
                if (col + 1 < bw && erow[col + 1] >= 1) { x[q].y = in_mode == 2 ? shift * x[q].y : x[q].y - shift; wrow[col + 1] = x[q].y; }
            }
        }
        const uint32_t o = oslot + r * k.len + j, t = k.o_t1 + j;
#pragma unroll
        for (int q = 1; q < A; q++) tw[q - 1] = buf.ld(t + (q - 1) * Bx);        // all loads in flight before the butterfly
        if (half) DftHalfIn<A, false>::run(x); else Dft<A, false>::run(x);
#pragma unroll
        for (int q = 1; q < A; q++) x[q] = cmul(x[q], tw[q - 1]);
#pragma unroll
        for (int q = 0; q < A; q++) buf.st(o + q * k.blk, x[q]);
    }
}

// stage 2: item (r, u): unit u >= 1 holds blocks (u, Ax - u); unit 0 holds block 0 and (Ax even) block Ax/2.
//          xa[k2] = Z[ka + Ax k2], xb[k2] = Z[kb + Ax k2]; split in registers; X[k] -> S[k * SY + row]
template <int Bq, class B>
CB_HD void c2_rowf2(uint32_t ok, B buf, uint32_t oslot, uint32_t oS, int row0, int nrows, int lane)
{
    const C2K k = c2k_load(buf, ok);
    const int A = k.A, nu = k.nu, items = nrows * nu, SY = k.SY;
    const bool aeven = (A & 1) == 0;
    const uint32_t mg_nr = nrows == k.RG ? k.mg_RG : div_magic((uint32_t) nrows);
    for (int i = lane; i < items; i += 32) {
        const uint32_t u = fdiv((uint32_t) i, mg_nr), r = (uint32_t) i - u * nrows;       // rows fastest over the lanes
        const int ka = (int) u, kb = u == 0 ? (aeven ? A / 2 : 0) : A - (int) u;
        const uint32_t o = oslot + r * k.len;
        cd xa[Bq], xb[Bq];
#pragma unroll
        for (int q = 0; q < Bq; q++) { xa[q] = buf.ld(o + ka * k.blk + q); xb[q] = buf.ld(o + kb * k.blk + q); }
        // twiddles of the split step: loaded before the butterflies run (unit 0 reads both of its small tables)
        cd ts[Bq], tm[Bq / 2 > 0 ? Bq / 2 : 1];
        {
            const uint32_t t = k.o_ts + u;
#pragma unroll
            for (int k2 = 0; k2 < Bq; k2++) ts[k2] = buf.ld(t + k2 * nu);
        }
        if (u == 0 && aeven) {
#pragma unroll
            for (int k2 = 0; k2 < Bq / 2; k2++) tm[k2] = buf.ld(k.o_tm + k2);
        }
        Dft<Bq, false>::run(xa); Dft<Bq, false>::run(xb);
        const uint32_t so = oS + (uint32_t) (row0 + (int) r);
        if (u != 0) {
            // pairs (ka + A k2, L - that = kb + A (Bq-1-k2)): results back into xa / xb, then one run of stores
#pragma unroll
            for (int k2 = 0; k2 < Bq; k2++) c2_split_pair(xa[k2], xb[Bq - 1 - k2], ts[k2], xa[k2], xb[Bq - 1 - k2]);
#pragma unroll
            for (int k2 = 0; k2 < Bq; k2++) {
                buf.st(so + (uint32_t) ((ka + A * k2) * SY), xa[k2]);
                buf.st(so + (uint32_t) ((kb + A * k2) * SY), xb[k2]);
            }
        } else {
            // block 0: k = A k2 pairs with A (Bq - k2); k = 0 gives the real X[0] and X[L], stored as ONE complex column
            // (row 0 of S): the column pass transforms both at once, see the packed stage M
            xa[0] = make_double2(xa[0].x + xa[0].y, xa[0].x - xa[0].y);
#pragma unroll
            for (int k2 = 1; k2 <= (Bq - 1) / 2; k2++) c2_split_pair(xa[k2], xa[Bq - k2], ts[k2], xa[k2], xa[Bq - k2]);
            if (Bq % 2 == 0 && Bq > 1) xa[Bq / 2] = cconj(xa[Bq / 2]);
#pragma unroll
            for (int k2 = 0; k2 < Bq; k2++) buf.st(so + (uint32_t) ((A * k2) * SY), xa[k2]);
            if (aeven) {
                // block A/2: k = A/2 + A k2 pairs with A/2 + A (Bq - 1 - k2)
#pragma unroll
                for (int k2 = 0; k2 < Bq / 2; k2++) c2_split_pair(xb[k2], xb[Bq - 1 - k2], tm[k2], xb[k2], xb[Bq - 1 - k2]);
                if (Bq % 2 == 1) xb[(Bq - 1) / 2] = cconj(xb[(Bq - 1) / 2]);
#pragma unroll
                for (int k2 = 0; k2 < Bq; k2++) buf.st(so + (uint32_t) ((A / 2 + A * k2) * SY), xb[k2]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// rows, inverse
// ------------------------------------------------------------------------------------------------------------
// stage 1': item (r, u): gather X of the unit's blocks from S, merge, inverse radix-Bx; slot[r][k1 * blkx + j]
template <int Bq, class B>
CB_HD void c2_rowi1(uint32_t ok, B buf, uint32_t oslot, uint32_t oS, int row0, int nrows, int lane)
{
    const C2K k = c2k_load(buf, ok);
    const int A = k.A, nu = k.nu, items = nrows * nu, SY = k.SY;
    const bool aeven = (A & 1) == 0;
    const uint32_t mg_nr = nrows == k.RG ? k.mg_RG : div_magic((uint32_t) nrows);
    for (int i = lane; i < items; i += 32) {
        const uint32_t u = fdiv((uint32_t) i, mg_nr), r = (uint32_t) i - u * nrows;
        const int ka = (int) u, kb = u == 0 ? (aeven ? A / 2 : 0) : A - (int) u;
        const uint32_t so = oS + (uint32_t) (row0 + (int) r);
        cd xa[Bq], xb[Bq], ts[Bq], tm[Bq / 2 > 0 ? Bq / 2 : 1];
        // all loads first: the unit's two blocks of the spectrum (block kb of unit 0 with odd Ax is block 0 again) and the
        // twiddles of the merge step
#pragma unroll
        for (int k2 = 0; k2 < Bq; k2++) {
            xa[k2] = buf.ld(so + (uint32_t) ((ka + A * k2) * SY));
            xb[k2] = buf.ld(so + (uint32_t) ((kb + A * k2) * SY));
            ts[k2] = buf.ld(k.o_ts + u + k2 * nu);
        }
        if (u == 0 && aeven) {
#pragma unroll
            for (int k2 = 0; k2 < Bq / 2; k2++) tm[k2] = buf.ld(k.o_tm + k2);
        }
        if (u != 0) {
#pragma unroll
            for (int k2 = 0; k2 < Bq; k2++) c2_merge_pair(xa[k2], xb[Bq - 1 - k2], ts[k2], xa[k2], xb[Bq - 1 - k2]);
        } else {
            xa[0] = make_double2(xa[0].x + xa[0].y, xa[0].x - xa[0].y);       // packed u(kx=0) + i u(kx=L), both real
#pragma unroll
            for (int k2 = 1; k2 <= (Bq - 1) / 2; k2++) c2_merge_pair(xa[k2], xa[Bq - k2], ts[k2], xa[k2], xa[Bq - k2]);
            if (Bq % 2 == 0 && Bq > 1) xa[Bq / 2] = make_double2(2.0 * xa[Bq / 2].x, -2.0 * xa[Bq / 2].y);
            if (aeven) {
#pragma unroll
                for (int k2 = 0; k2 < Bq / 2; k2++) c2_merge_pair(xb[k2], xb[Bq - 1 - k2], tm[k2], xb[k2], xb[Bq - 1 - k2]);
                if (Bq % 2 == 1) xb[(Bq - 1) / 2] = make_double2(2.0 * xb[(Bq - 1) / 2].x, -2.0 * xb[(Bq - 1) / 2].y);
            }
        }
        Dft<Bq, true>::run(xa); Dft<Bq, true>::run(xb);
        const uint32_t o = oslot + r * k.len;
#pragma unroll
        for (int q = 0; q < Bq; q++) buf.st(o + ka * k.blk + q, xa[q]);
        if (kb != ka) {
#pragma unroll
            for (int q = 0; q < Bq; q++) buf.st(o + kb * k.blk + q, xb[q]);
        }
    }
}

// stage 2': item (r, j): conj twiddle, inverse radix-Ax, masked store of the wanted columns of the box (x0, y0, bw x .)
// of u / el (row stride `stride`); mask_mode 1: only elements with el >= 1 (AllInt), add: u += result
template <int A, class B>
CB_HD void c2_rowi2(uint32_t ok, B buf, uint32_t oslot, double *u, const int *el, int mask_mode,
                    int add, int x0, int y0, int bw, int stride, int row0, int nrows, int lane,
                    const double *sub = nullptr, double *acc = nullptr)
{
    const C2K k = c2k_load(buf, ok);
    const int Bx = k.B, items = nrows * Bx, Fx = k.Fx;
    constexpr int Q0 = (A % 2 == 0) ? A / 2 : 0;        // even Ax: outputs q < Ax/2 lie left of column Fx
    for (int i = lane; i < items; i += 32) {
        const uint32_t r = fdiv((uint32_t) i, k.mg_B), j = (uint32_t) i - r * Bx;
        const uint32_t o = oslot + r * k.len + j, t = k.o_t1 + j;
        // what the masked store needs from global memory (element division, right-hand side) is requested first: the L2
        // round trip runs while the butterfly does
        const size_t r0 = (size_t) (y0 + row0 + (int) r) * stride + x0;
        constexpr int NZ = 2 * (A - Q0);
        int ev[NZ]; double sv[NZ]; uint32_t okv = 0u;
#pragma unroll
        for (int q = Q0; q < A; q++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int z_ = 2 * (q - Q0) + h, ic = 2 * (int) (j + q * Bx) - Fx + h;
                const bool inb = ic >= 0 && ic < bw;
                if (inb) okv |= 1u << z_;
                ev[z_] = (inb && el != nullptr) ? el[r0 + ic] : 1;
                sv[z_] = (inb && sub != nullptr) ? sub[r0 + ic] : 0.0;
            }
        }
        cd x[A], tw[A > 1 ? A - 1 : 1];
#pragma unroll
        for (int q = 0; q < A; q++) x[q] = buf.ld(o + q * k.blk);
#pragma unroll
        for (int q = 1; q < A; q++) tw[q - 1] = buf.ld(t + (q - 1) * Bx);
#pragma unroll
        for (int q = 1; q < A; q++) x[q] = cmulc(x[q], tw[q - 1]);
        Dft<A, true>::run(x);
        // masked store (+ fused output pass, ConvFuse: right-hand side subtracted, sum over the contact area)
        double a_ = 0.0;
#pragma unroll
        for (int q = Q0; q < A; q++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int z_ = 2 * (q - Q0) + h;
                if (okv & (1u << z_)) {
                    const int ic = 2 * (int) (j + q * Bx) - Fx + h, e = ev[z_];
                    if (!(mask_mode == 1 && e < 1)) {
                        double v = h ? x[q].y : x[q].x;
                        if (add) v += u[r0 + ic];
                        if (sub != nullptr) v -= sv[z_];
                        u[r0 + ic] = v;
                        if (e >= 1) a_ += v;
                    }
                }
            }
        }
        if (acc != nullptr) *acc += a_;
    }
}

// ------------------------------------------------------------------------------------------------------------
// columns
// ------------------------------------------------------------------------------------------------------------
// stage A: item (cc, j), j < By: x[q] = S[col cc][j + By q] (zero beyond n_in rows), radix-Ay, y[k1] *= w_Ly^(j k1),
//          slot[cc][k1 * blky + j]
template <int A, class B>
CB_HD void c2_colA(uint32_t ok, B buf, uint32_t oslot, uint32_t oScol, int ncols, int n_in, int lane)
{
    const C2K k = c2k_load(buf, ok);
    const int By = k.B, items = ncols * By, SY = k.SY;
    const uint32_t mg_nc = ncols == k.G ? k.mg_G : div_magic((uint32_t) ncols);
    constexpr bool half = DftHalfIn<A, false>::ok;
    constexpr int NL = half ? A / 2 : A;
    for (int i = lane; i < items; i += 32) {
        const uint32_t j = fdiv((uint32_t) i, mg_nc), cc = (uint32_t) i - j * ncols;      // columns fastest over the lanes
        const uint32_t s = oScol + cc * SY + j;
        cd x[A];
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] = buf.ldz(s + q * By, (int) (j + q * By) < n_in);
        const uint32_t o = oslot + cc * k.len + j, t = k.o_t1 + j;
        // twiddles in two batches: the first is in flight while the butterfly runs, the second while the first is used
        constexpr int H = A / 2;
        cd tw0[H], tw1[A - 1 - H > 0 ? A - 1 - H : 1];
#pragma unroll
        for (int q = 1; q <= H; q++) tw0[q - 1] = buf.ld(t + (q - 1) * By);
        if (half) DftHalfIn<A, false>::run(x); else Dft<A, false>::run(x);
#pragma unroll
        for (int q = H + 1; q < A; q++) tw1[q - H - 1] = buf.ld(t + (q - 1) * By);
        buf.st(o, x[0]);
#pragma unroll
        for (int q = 1; q <= H; q++) buf.st(o + q * k.blk, cmul(x[q], tw0[q - 1]));
#pragma unroll
        for (int q = H + 1; q < A; q++) buf.st(o + q * k.blk, cmul(x[q], tw1[q - H - 1]));
    }
}

// stage M: item (cc, k1), k1 < Ay: radix-By forward over the block, multiply by C^[k2][cc][k1] (frequency k1 + Ay k2),
//          radix-By inverse, back in place.  chat_g points at the group's [By][Ay][G] coefficients.
template <int Bq, class B>
CB_HD void c2_colM(uint32_t ok, B buf, uint32_t oslot, const cd *chat_g, int ncols, int lane)
{
    const C2K k = c2k_load(buf, ok);
    const int A = k.A, items = ncols * A, GA = k.G * A;
    const uint32_t mg_nc = ncols == k.G ? k.mg_G : div_magic((uint32_t) ncols);
    for (int i = lane; i < items; i += 32) {
        const uint32_t k1 = fdiv((uint32_t) i, mg_nc), cc = (uint32_t) i - k1 * ncols;
        const cd *hp = chat_g + k1 * k.G + cc;        // (k2 * A + k1) * G + cc
        cd h[Bq], x[Bq];
#pragma unroll
        for (int q = 0; q < Bq; q++) h[q] = c2_ldg_early(hp + q * GA);
        const uint32_t o = oslot + cc * k.len + k1 * k.blk;
#pragma unroll
        for (int q = 0; q < Bq; q++) x[q] = buf.ld(o + q);
        Dft<Bq, false>::run(x);
#pragma unroll
        for (int q = 0; q < Bq; q++) x[q] = cmul(x[q], h[q]);
        Dft<Bq, true>::run(x);
#pragma unroll
        for (int q = 0; q < Bq; q++) buf.st(o + q, x[q]);
    }
}

// stage M of the packed column (S row 0 = X[0] + i X[L], two real sequences in y): its transform is T = A + i B with A, B
// Hermitian, and the product wanted is U[ky] = C^_0 A + i C^_L B = P[ky] T[ky] + Q[ky] conj(T[N - ky]) with
// P = (C^_0 + C^_L)/2, Q = (C^_0 - C^_L)/2 (stored by the builder in the places of columns 0 and L).  The partner N - ky of
// ky = k1 + Ay q sits in block (Ay - k1) % Ay, i.e. in another lane: the forward butterflies therefore write T to a scratch
// copy (the slot of a neighbour column that has already left the slot), and a second step reads own and partner values
// from there.  Two short steps for one column of the whole product.
template <int Bq, class B>
CB_HD void c2_colM0a(uint32_t ok, B buf, uint32_t oslot, uint32_t oscr, int lane)
{
    const C2K k = c2k_load(buf, ok);
    const int A = k.A;
    for (int k1 = lane; k1 < A; k1 += 32) {
        cd x[Bq];
#pragma unroll
        for (int q = 0; q < Bq; q++) x[q] = buf.ld(oslot + k1 * k.blk + q);
        Dft<Bq, false>::run(x);
#pragma unroll
        for (int q = 0; q < Bq; q++) buf.st(oscr + k1 * k.blk + q, x[q]);
    }
}
template <int Bq, class B>
CB_HD void c2_colM0b(uint32_t ok, B buf, uint32_t oslot, uint32_t oscr, const cd *chat_p, const cd *chat_q, int lane)
{
    const C2K k = c2k_load(buf, ok);
    const int A = k.A, GA = k.G * A;
    for (int k1 = lane; k1 < A; k1 += 32) {
        const int kb = (A - k1) % A;
        cd pc[Bq], qc[Bq], x[Bq];
#pragma unroll
        for (int q = 0; q < Bq; q++) {
            pc[q] = c2_ldg_early(chat_p + k1 * k.G + q * GA);
            qc[q] = c2_ldg_early(chat_q + k1 * k.G + q * GA);
        }
#pragma unroll
        for (int q = 0; q < Bq; q++) {
            const int qp = k1 == 0 ? (Bq - q) % Bq : Bq - 1 - q;               // N - (k1 + A q) = kb + A qp
            const cd t = buf.ld(oscr + k1 * k.blk + q), tp = buf.ld(oscr + kb * k.blk + qp);
            x[q] = cadd(cmul(pc[q], t), cmulc(qc[q], tp));
        }
        Dft<Bq, true>::run(x);
#pragma unroll
        for (int q = 0; q < Bq; q++) buf.st(oslot + k1 * k.blk + q, x[q]);
    }
}

// stage C: item (cc, j): conj twiddle, inverse radix-Ay, rows Fy .. Fy + n_out - 1 (outputs q >= Ay/2) back to S
template <int A, class B>
CB_HD void c2_colC(uint32_t ok, B buf, uint32_t oslot, uint32_t oScol, int ncols, int n_out, int lane)
{
    const C2K k = c2k_load(buf, ok);
    const int By = k.B, items = ncols * By, SY = k.SY;
    const uint32_t mg_nc = ncols == k.G ? k.mg_G : div_magic((uint32_t) ncols);
    for (int i = lane; i < items; i += 32) {
        const uint32_t j = fdiv((uint32_t) i, mg_nc), cc = (uint32_t) i - j * ncols;
        const uint32_t o = oslot + cc * k.len + j, t = k.o_t1 + j;
        cd x[A];
#pragma unroll
        for (int q = 0; q < A; q++) x[q] = buf.ld(o + q * k.blk);
        {   // twiddles in two batches: all loads of a batch are in flight before its first use
            constexpr int H = (A - 1 + 1) / 2;
            cd tw[H];
#pragma unroll
            for (int q = 1; q <= H; q++) tw[q - 1] = buf.ld(t + (q - 1) * By);
#pragma unroll
            for (int q = 1; q <= H; q++) x[q] = cmulc(x[q], tw[q - 1]);
#pragma unroll
            for (int q = H + 1; q < A; q++) tw[q - H - 1] = buf.ld(t + (q - 1) * By);
#pragma unroll
            for (int q = H + 1; q < A; q++) x[q] = cmulc(x[q], tw[q - H - 1]);
        }
        Dft<A, true>::run(x);
        const uint32_t s = oScol + cc * SY + j;
#pragma unroll
        for (int q = A / 2; q < A; q++) buf.stp(s + (q - A / 2) * By, x[q], (int) (j + (q - A / 2) * By) < n_out);
    }
}

// ------------------------------------------------------------------------------------------------------------
// the three passes of one product as seen by one warp (device: all warps call; host emulation: one call per warp)
// ------------------------------------------------------------------------------------------------------------
// Each pass is a function of its own per RADIX PAIR (template), never inlined on the device: ptxas optimises a function
// that holds two or three butterfly bodies far better than one function that holds the bodies of every radix (measured
// on B200, same source, 91x91: 53 k cycles per product with only the radices of that plan instantiated, 82 k with all of
// them inlined behind switches -- and the old block-wide path showed the same dependence).  The plans are therefore
// restricted to the radix pairs listed in CB2_ROW_CLASSES / CB2_COL_CLASSES; other sizes use the block-wide path.
struct C2Pass {                   // what a pass needs from the plan, by value (registers across the call)
    uint32_t oS, oW, ok;          // element offsets of S, of the slots, of the pass's stage constants
    int nslot, slot_len;
    int RG;                       // rows
    int G, ngrp, ncol, Ly, SY, collen;   // columns
};
CB_HD C2Pass c2_pass_rows(const ConvPlan &P)
{
    const Conv2Plan &c = P.c2;
    C2Pass a;
    a.oS = (uint32_t) c.off_S / 16; a.oW = (uint32_t) c.off_W / 16; a.ok = (uint32_t) (c.off_tab / 16 + c.o_kr);
    a.nslot = c.nslot; a.slot_len = c.slot_len; a.RG = c.RG;
    a.G = 0; a.ngrp = 0; a.ncol = 0; a.Ly = 0; a.SY = c.SY; a.collen = 0;
    return a;
}
CB_HD C2Pass c2_pass_cols(const ConvPlan &P)
{
    const Conv2Plan &c = P.c2;
    C2Pass a;
    a.oS = (uint32_t) c.off_S / 16; a.oW = (uint32_t) c.off_W / 16; a.ok = (uint32_t) (c.off_tab / 16 + c.o_kc);
    a.nslot = c.nslot; a.slot_len = c.slot_len; a.RG = 0;
    a.G = c.G; a.ngrp = c.ngrp; a.ncol = P.Fx; a.Ly = P.Ly; a.SY = c.SY; a.collen = c.collen;
    return a;
}

template <int AX, int BX, class B>
CB_HNI void c2_rows_fwd_t(const C2Pass a, B buf, const double *base, int bw, int bh, int stride, int warp,
                          const int *el_in, double shift, int in_mode)
{
    if (warp >= a.nslot) return;
    const uint32_t ok = a.ok, oS = a.oS, oslot = a.oW + (uint32_t) (warp * a.slot_len);
    const int RG = a.RG, nslot = a.nslot;
    CB2_STAGGER_WAIT(warp);
    CB2_TICK_INIT();
    for (int r0 = warp * RG; r0 < bh; r0 += nslot * RG) {
        const int nr = bh - r0 < RG ? bh - r0 : RG;
        const double *src = base + (size_t) r0 * stride;
        CB2_TICK(0);
        CB2_LANES((c2_rowf1<AX>(ok, buf, oslot, src, nr, bw, stride, lane, el_in ? el_in + (size_t) r0 * stride : nullptr, shift, in_mode)));
        CB2_TICK(1);
        CB2_LANES((c2_rowf2<BX>(ok, buf, oslot, oS, r0, nr, lane)));
        CB2_TICK(2);
    }
}

template <int AY, int BY, class B>
CB_HNI void c2_cols_t(const C2Pass a, B buf, const cd *chat, int n_in, int n_out, int warp)
{
    if (warp >= a.nslot) return;
    const uint32_t ok = a.ok, oS = a.oS, oslot = a.oW + (uint32_t) (warp * a.slot_len);
    const int nslot = a.nslot, G = a.G, ngrp = a.ngrp, ncol = a.ncol, Ly = a.Ly, SYc = a.SY, collen = a.collen;
    CB2_STAGGER_WAIT(warp);
    CB2_TICK_INIT();
    for (int g = warp; g < ngrp; g += nslot) {
        const int left = ncol - g * G, nc = left < G ? left : G;
        const uint32_t oScol = oS + (uint32_t) (g * G * SYc);
        const cd *chat_g = chat + (size_t) g * G * Ly;
        CB2_TICK(3);
        CB2_LANES((c2_colA<AY>(ok, buf, oslot, oScol, nc, n_in, lane)));
        CB2_TICK(4);
        if (g != 0) {
            // (requesting the coefficients before stage A was tried: the 48 registers they hold through stage A spill to
            //  local memory, columns 41 k -> 47 k cycles per 91x91 product)
            CB2_LANES((c2_colM<BY>(ok, buf, oslot, chat_g, nc, lane)));
            CB2_TICK(5);
            CB2_LANES((c2_colC<AY>(ok, buf, oslot, oScol, nc, n_out, lane)));
        } else {
            // group 0: its first column is the packed pair (kx = 0, kx = Fx).  The other columns of the group go first;
            // the slot of column 1 is free then and serves as the scratch copy of the packed column's transform
            const uint32_t oscr = oslot + (uint32_t) collen;
            if (nc > 1) {
                CB2_LANES((c2_colM<BY>(ok, buf, oscr, chat_g + 1, nc - 1, lane)));
                CB2_LANES((c2_colC<AY>(ok, buf, oscr, oScol + (uint32_t) SYc, nc - 1, n_out, lane)));
            }
            const cd *chat_q = chat + (size_t) (ncol / G) * G * Ly + (ncol % G);      // coefficients of "column Fx": Q
            CB2_LANES((c2_colM0a<BY>(ok, buf, oslot, oscr, lane)));
            CB2_LANES((c2_colM0b<BY>(ok, buf, oslot, oscr, chat_g, chat_q, lane)));
            CB2_TICK(5);
            CB2_LANES((c2_colC<AY>(ok, buf, oslot, oScol, 1, n_out, lane)));
        }
        CB2_TICK(6);
    }
}

template <int AX, int BX, class B>
CB_HNI void c2_rows_inv_t(const C2Pass a, B buf, double *u, const int *el, int mask_mode, int add, int x0, int y0, int bw,
                          int bh, int stride, int warp, const double *sub, int want_sum, uint32_t osum)
{
    double acc[CB2_NL];
    for (int l = 0; l < CB2_NL; l++) acc[l] = 0.0;
    if (warp >= a.nslot) return;
    const uint32_t ok = a.ok, oS = a.oS, oslot = a.oW + (uint32_t) (warp * a.slot_len);
    const int RG = a.RG, nslot = a.nslot;
    CB2_STAGGER_WAIT(warp);
    CB2_TICK_INIT();
    for (int r0 = warp * RG; r0 < bh; r0 += nslot * RG) {
        const int nr = bh - r0 < RG ? bh - r0 : RG;
        CB2_TICK(7);
        CB2_LANES((c2_rowi1<BX>(ok, buf, oslot, oS, r0, nr, lane)));
        CB2_TICK(8);
        CB2_LANES((c2_rowi2<AX>(ok, buf, oslot, u, el, mask_mode, add, x0, y0, bw, stride, r0, nr, lane, sub,
                                want_sum ? &acc[CB2_LI(lane)] : nullptr)));
        CB2_TICK(9);
    }
    if (want_sum) {
        // per-warp partial of the fused masked sum: fixed shuffle tree, lane 0 stores it behind the tables (element osum + warp)
#ifdef __CUDA_ARCH__
        double s_ = acc[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s_ += __shfl_xor_sync(0xffffffffu, s_, o);
        if ((threadIdx.x & 31u) == 0) buf.st(osum + (uint32_t) warp, make_double2(s_, 0.0));
#else
        double s_ = 0.0;
        for (int l = 0; l < 32; l++) s_ += acc[l];
        buf.st(osum + (uint32_t) warp, make_double2(s_, 0.0));
#endif
    }
}

// The radix pairs served (rows: Lx = AX * BX, columns: 2 Fy = AY * BY): the transform sizes of the reference's test and
// benchmark grids (opt_fft_size of 11, 19, 43/45, 71, 81, 91/93) and of the ladder of contact-box sizes.
#define CB2_ROW_CLASSES(X) X(12, 8) X(9, 9) X(10, 8) X(12, 6) X(8, 8) X(6, 8) X(9, 5) X(4, 8) X(4, 6) X(4, 5) X(4, 4) X(4, 3)
#define CB2_COL_CLASSES(X) X(16, 12) X(18, 9) X(16, 10) X(12, 12) X(16, 8) X(12, 8) X(10, 9) X(8, 8) X(8, 6) X(8, 5) X(8, 4) X(6, 4)

CB_HD bool c2_row_class(int A, int Bq)
{
#define CB2_X_(a, b) if (A == a && Bq == b) return true;
    CB2_ROW_CLASSES(CB2_X_)
#undef CB2_X_
    return false;
}
CB_HD bool c2_col_class(int A, int Bq)
{
#define CB2_X_(a, b) if (A == a && Bq == b) return true;
    CB2_COL_CLASSES(CB2_X_)
#undef CB2_X_
    return false;
}

template <class B>
CB_HD void c2_rows_fwd(const ConvPlan &P, B buf, const double *base, int bw, int bh, int stride, int warp,
                       const int *el_in = nullptr, double shift = 0.0, int in_mode = 1)
{
    const C2Pass a = c2_pass_rows(P);
    switch (P.c2.Ax * 32 + P.c2.Bx) {
#define CB2_X_(ax, bx) case ax * 32 + bx: c2_rows_fwd_t<ax, bx>(a, buf, base, bw, bh, stride, warp, el_in, shift, in_mode); break;
    CB2_ROW_CLASSES(CB2_X_)
#undef CB2_X_
    default: break;
    }
}

template <class B>
CB_HD void c2_cols(const ConvPlan &P, B buf, const cd *chat, int n_in, int n_out, int warp)
{
    const C2Pass a = c2_pass_cols(P);
    switch (P.c2.Ay * 32 + P.c2.By) {
#define CB2_X_(ay, by) case ay * 32 + by: c2_cols_t<ay, by>(a, buf, chat, n_in, n_out, warp); break;
    CB2_COL_CLASSES(CB2_X_)
#undef CB2_X_
    default: break;
    }
}

template <class B>
CB_HD void c2_rows_inv(const ConvPlan &P, B buf, double *u, const int *el, int mask_mode, int add, int x0, int y0, int bw,
                       int bh, int stride, int warp, const double *sub = nullptr, int want_sum = 0)
{
    const C2Pass a = c2_pass_rows(P);
    const uint32_t osum = (uint32_t) (P.c2.off_tab / 16 + P.c2.tab_len);        // per-warp partial sums behind the tables
    switch (P.c2.Ax * 32 + P.c2.Bx) {
#define CB2_X_(ax, bx) case ax * 32 + bx: c2_rows_inv_t<ax, bx>(a, buf, u, el, mask_mode, add, x0, y0, bw, bh, stride, warp, sub, want_sum, osum); break;
    CB2_ROW_CLASSES(CB2_X_)
#undef CB2_X_
    default: break;
    }
}

// position of frequency (kx, ky) inside a coefficient block: [group][k2][k1][column in group], ky = k1 + Ay k2
CB_HD size_t c2_chat_index(const ConvPlan &P, int kx, int ky)
{
    const Conv2Plan &c = P.c2;
    const int g = kx / c.G, cc = kx - g * c.G, k2 = ky / c.Ay, k1 = ky - k2 * c.Ay;
    return ((size_t) (g * c.By + k2) * c.Ay + k1) * c.G + cc;
}

// ---- coefficient transform for the warp-resident layout, as two dense DFT passes (one-off per grid/material/block) ----
// T[iy][kx] = sum_x pad(iy, x) exp(-2 pi i x kx / 2Fx),  iy < 2Fy, kx <= Fx   (pad: cf(0,0) at (Fx,Fy), m_aijpj.f90:894-901)
CB_HD cd c2_chat_row_entry(const ConvPlan &P, const RowSrc &src, int iy, int kx, const cd *twx)
{
    const int N = 2 * P.Fx;
    double re = 0.0, im = 0.0;
    int idx = 0;
    for (int x = 0; x < N; x++) {
        const double v = rowsrc_get(src, iy, x);
        re += v * twx[idx].x; im += v * twx[idx].y;
        idx += kx; if (idx >= N) idx -= N;
    }
    return make_double2(re, im);
}
// C^(kx, ky) = scale * sum_iy T[iy][kx] exp(-2 pi i iy ky / 2Fy)
CB_HD cd c2_chat_col_entry(const ConvPlan &P, const cd *T, int kx, int ky, const cd *twy, double scale)
{
    const int N = 2 * P.Fy, ld = P.Fx + 1;
    double re = 0.0, im = 0.0;
    int idx = 0;
    for (int iy = 0; iy < N; iy++) {
        const cd t = T[(size_t) iy * ld + kx], w = twy[idx];
        re += t.x * w.x - t.y * w.y; im += t.x * w.y + t.y * w.x;
        idx += ky; if (idx >= N) idx -= N;
    }
    return make_double2(scale * re, scale * im);
}

// value stored at the layout position of (kx, ky): the plain spectrum for 0 < kx < Fx; P and Q of the packed column pair
// (see c2_colM_packed) in the places of kx = 0 and kx = Fx
CB_HD cd c2_chat_value(const ConvPlan &P, const cd *T, int kx, int ky, const cd *twy, double scale)
{
    if (kx != 0 && kx != P.Fx) return c2_chat_col_entry(P, T, kx, ky, twy, scale);
    const cd a = c2_chat_col_entry(P, T, 0, ky, twy, scale), b = c2_chat_col_entry(P, T, P.Fx, ky, twy, scale);
    return kx == 0 ? make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y)) : make_double2(0.5 * (a.x - b.x), 0.5 * (a.y - b.y));
}

}  // namespace cb200
