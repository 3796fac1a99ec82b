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
#define CB2_LANES(call) do { for (int lane = 0; lane < 32; lane++) { call; } } while (0)
#endif

// radices served: rows stage 1 / columns stage A, C (even ones input/output pruned); rows stage 2 / columns stage M
#define CB2_SWITCH_A(r, CALL)                                                                                   \
    switch (r) {                                                                                                \
    case 2: { CALL(2); } break;   case 3: { CALL(3); } break;   case 4: { CALL(4); } break;                     \
    case 5: { CALL(5); } break;   case 6: { CALL(6); } break;   case 7: { CALL(7); } break;                     \
    case 8: { CALL(8); } break;   case 9: { CALL(9); } break;   case 10: { CALL(10); } break;                   \
    case 12: { CALL(12); } break; case 16: { CALL(16); } break; case 18: { CALL(18); } break;                   \
    default: break;                                                                                             \
    }
#define CB2_SWITCH_B(r, CALL)                                                                                   \
    switch (r) {                                                                                                \
    case 1: { CALL(1); } break;   case 2: { CALL(2); } break;   case 3: { CALL(3); } break;                     \
    case 4: { CALL(4); } break;   case 5: { CALL(5); } break;   case 6: { CALL(6); } break;                     \
    case 7: { CALL(7); } break;   case 8: { CALL(8); } break;   case 9: { CALL(9); } break;                     \
    case 10: { CALL(10); } break; case 12: { CALL(12); } break; case 16: { CALL(16); } break;                   \
    default: break;                                                                                             \
    }

CB_HD bool c2_radix_a(int r) { return (r >= 2 && r <= 10) || r == 12 || r == 16 || r == 18; }
CB_HD bool c2_radix_b(int r) { return (r >= 1 && r <= 10) || r == 12 || r == 16; }

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
// rows, forward
// ------------------------------------------------------------------------------------------------------------
// Source of the real rows: tractions (box of a grid, row stride `stride`, zero beyond bw columns) -- the coefficient
// transforms are built by dense DFTs (k_chat2_*), not through this path.
//
// stage 1: item (r, j), j < Bx: x[q] = z[j + Bx q] = (row[2n], row[2n+1]); radix-Ax butterfly; y[k1] *= w_Lx^(j k1);
//          slot[r][k1 * blkx + j]
template <int A, class B>
CB_HD void c2_rowf1(const ConvPlan &P, B buf, uint32_t oslot, uint32_t otab, const double *src, int nrows, int bw,
                    int stride, int lane)
{
    const Conv2Plan &c = P.c2;
    const int Bx = c.Bx, items = nrows * Bx;
    constexpr bool half = DftHalfIn<A, false>::ok;
    constexpr int NL = half ? A / 2 : A;
    for (int i = lane; i < items; i += 32) {
        const uint32_t r = fdiv((uint32_t) i, c.mg_Bx), j = (uint32_t) i - r * Bx;
        const double *row = src + (size_t) r * stride;
        cd x[A];
#pragma unroll
        for (int q = 0; q < NL; q++) {
            const int col = 2 * (int) (j + q * Bx);
            x[q] = make_double2(col < bw ? row[col] : 0.0, col + 1 < bw ? row[col + 1] : 0.0);
        }
        if (half) DftHalfIn<A, false>::run(x); else Dft<A, false>::run(x);
        const uint32_t o = oslot + r * c.rowlen + j, t = otab + c.o_t1x + j;
        buf.st(o, x[0]);
#pragma unroll
        for (int q = 1; q < A; q++) buf.st(o + q * c.blkx, cmul(x[q], buf.ld(t + (q - 1) * Bx)));
    }
}

// stage 2: item (r, u): unit u >= 1 holds blocks (u, Ax - u); unit 0 holds block 0 and (Ax even) block Ax/2.
//          xa[k2] = Z[ka + Ax k2], xb[k2] = Z[kb + Ax k2]; split in registers; X[k] -> S[k * SY + row]
template <int Bq, class B>
CB_HD void c2_rowf2(const ConvPlan &P, B buf, uint32_t oslot, uint32_t otab, uint32_t oS, int SY, int row0, int nrows,
                    int lane)
{
    const Conv2Plan &c = P.c2;
    const int A = c.Ax, nu = c.nux, items = nrows * nu, L = P.Lx;
    const bool aeven = (A & 1) == 0;
    for (int i = lane; i < items; i += 32) {
        const uint32_t r = fdiv((uint32_t) i, c.mg_nux), u = (uint32_t) i - r * nu;
        const int ka = (int) u, kb = u == 0 ? (aeven ? A / 2 : 0) : A - (int) u;
        const uint32_t o = oslot + r * c.rowlen;
        cd xa[Bq], xb[Bq];
#pragma unroll
        for (int q = 0; q < Bq; q++) { xa[q] = buf.ld(o + ka * c.blkx + q); xb[q] = buf.ld(o + kb * c.blkx + q); }
        Dft<Bq, false>::run(xa); Dft<Bq, false>::run(xb);
        const uint32_t so = oS + (uint32_t) (row0 + (int) r);
        if (u != 0) {
            const uint32_t t = otab + c.o_tsx + u;
#pragma unroll
            for (int k2 = 0; k2 < Bq; k2++) {
                cd xk, xlk;
                c2_split_pair(xa[k2], xb[Bq - 1 - k2], buf.ld(t + k2 * nu), xk, xlk);
                buf.st(so + (uint32_t) ((ka + A * k2) * SY), xk);
                buf.st(so + (uint32_t) ((kb + A * (Bq - 1 - k2)) * SY), xlk);
            }
        } else {
            // block 0: k = A k2 pairs with A (Bq - k2); k = 0 gives X[0] and X[L]
            buf.st(so, make_double2(xa[0].x + xa[0].y, 0.0));
            buf.st(so + (uint32_t) (L * SY), make_double2(xa[0].x - xa[0].y, 0.0));
            const uint32_t t = otab + c.o_tsx;
#pragma unroll
            for (int k2 = 1; k2 <= (Bq - 1) / 2; k2++) {
                cd xk, xlk;
                c2_split_pair(xa[k2], xa[Bq - k2], buf.ld(t + k2 * nu), xk, xlk);
                buf.st(so + (uint32_t) ((A * k2) * SY), xk);
                buf.st(so + (uint32_t) ((A * (Bq - k2)) * SY), xlk);
            }
            if (Bq % 2 == 0 && Bq > 1) buf.st(so + (uint32_t) ((A * (Bq / 2)) * SY), cconj(xa[Bq / 2]));
            if (aeven) {
                // block A/2: k = A/2 + A k2 pairs with A/2 + A (Bq - 1 - k2)
                const uint32_t tm = otab + c.o_tmx;
#pragma unroll
                for (int k2 = 0; k2 < Bq / 2; k2++) {
                    cd xk, xlk;
                    c2_split_pair(xb[k2], xb[Bq - 1 - k2], buf.ld(tm + k2), xk, xlk);
                    buf.st(so + (uint32_t) ((A / 2 + A * k2) * SY), xk);
                    buf.st(so + (uint32_t) ((A / 2 + A * (Bq - 1 - k2)) * SY), xlk);
                }
                if (Bq % 2 == 1) buf.st(so + (uint32_t) ((A / 2 + A * ((Bq - 1) / 2)) * SY), cconj(xb[(Bq - 1) / 2]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// rows, inverse
// ------------------------------------------------------------------------------------------------------------
// stage 1': item (r, u): gather X of the unit's blocks from S, merge, inverse radix-Bx; slot[r][k1 * blkx + j]
template <int Bq, class B>
CB_HD void c2_rowi1(const ConvPlan &P, B buf, uint32_t oslot, uint32_t otab, uint32_t oS, int SY, int row0, int nrows,
                    int lane)
{
    const Conv2Plan &c = P.c2;
    const int A = c.Ax, nu = c.nux, items = nrows * nu, L = P.Lx;
    const bool aeven = (A & 1) == 0;
    for (int i = lane; i < items; i += 32) {
        const uint32_t r = fdiv((uint32_t) i, c.mg_nux), u = (uint32_t) i - r * nu;
        const int ka = (int) u, kb = u == 0 ? (aeven ? A / 2 : 0) : A - (int) u;
        const uint32_t so = oS + (uint32_t) (row0 + (int) r);
        cd xa[Bq], xb[Bq];
        if (u != 0) {
            const uint32_t t = otab + c.o_tsx + u;
#pragma unroll
            for (int k2 = 0; k2 < Bq; k2++)
                c2_merge_pair(buf.ld(so + (uint32_t) ((ka + A * k2) * SY)), buf.ld(so + (uint32_t) ((kb + A * (Bq - 1 - k2)) * SY)),
                              buf.ld(t + k2 * nu), xa[k2], xb[Bq - 1 - k2]);
        } else {
            const cd p = buf.ld(so), q = buf.ld(so + (uint32_t) (L * SY));
            const cd uu = make_double2(p.x + q.x, p.y - q.y), d = make_double2(p.x - q.x, p.y + q.y);
            xa[0] = make_double2(uu.x - d.y, uu.y + d.x);
            const uint32_t t = otab + c.o_tsx;
#pragma unroll
            for (int k2 = 1; k2 <= (Bq - 1) / 2; k2++)
                c2_merge_pair(buf.ld(so + (uint32_t) ((A * k2) * SY)), buf.ld(so + (uint32_t) ((A * (Bq - k2)) * SY)),
                              buf.ld(t + k2 * nu), xa[k2], xa[Bq - k2]);
            if (Bq % 2 == 0 && Bq > 1) {
                const cd a = buf.ld(so + (uint32_t) ((A * (Bq / 2)) * SY));
                xa[Bq / 2] = make_double2(2.0 * a.x, -2.0 * a.y);
            }
            if (aeven) {
                const uint32_t tm = otab + c.o_tmx;
#pragma unroll
                for (int k2 = 0; k2 < Bq / 2; k2++)
                    c2_merge_pair(buf.ld(so + (uint32_t) ((A / 2 + A * k2) * SY)),
                                  buf.ld(so + (uint32_t) ((A / 2 + A * (Bq - 1 - k2)) * SY)), buf.ld(tm + k2), xb[k2], xb[Bq - 1 - k2]);
                if (Bq % 2 == 1) {
                    const cd a = buf.ld(so + (uint32_t) ((A / 2 + A * ((Bq - 1) / 2)) * SY));
                    xb[(Bq - 1) / 2] = make_double2(2.0 * a.x, -2.0 * a.y);
                }
            } else {
#pragma unroll
                for (int k2 = 0; k2 < Bq; k2++) xb[k2] = xa[k2];
            }
        }
        Dft<Bq, true>::run(xa); Dft<Bq, true>::run(xb);
        const uint32_t o = oslot + r * c.rowlen;
#pragma unroll
        for (int q = 0; q < Bq; q++) buf.st(o + ka * c.blkx + q, xa[q]);
        if (kb != ka) {
#pragma unroll
            for (int q = 0; q < Bq; q++) buf.st(o + kb * c.blkx + q, xb[q]);
        }
    }
}

// stage 2': item (r, j): conj twiddle, inverse radix-Ax, masked store of the wanted columns of the box (x0, y0, bw x .)
// of u / el (row stride `stride`); mask_mode 1: only elements with el >= 1 (AllInt), add: u += result
template <int A, class B>
CB_HD void c2_rowi2(const ConvPlan &P, B buf, uint32_t oslot, uint32_t otab, double *u, const int *el, int mask_mode,
                    int add, int x0, int y0, int bw, int stride, int row0, int nrows, int lane)
{
    const Conv2Plan &c = P.c2;
    const int Bx = c.Bx, items = nrows * Bx, Fx = P.Fx;
    constexpr int Q0 = (A % 2 == 0) ? A / 2 : 0;        // even Ax: outputs q < Ax/2 lie left of column Fx
    for (int i = lane; i < items; i += 32) {
        const uint32_t r = fdiv((uint32_t) i, c.mg_Bx), j = (uint32_t) i - r * Bx;
        const uint32_t o = oslot + r * c.rowlen + j, t = otab + c.o_t1x + j;
        cd x[A];
        x[0] = buf.ld(o);
#pragma unroll
        for (int q = 1; q < A; q++) x[q] = cmulc(buf.ld(o + q * c.blkx), buf.ld(t + (q - 1) * Bx));
        Dft<A, true>::run(x);
        const size_t r0 = (size_t) (y0 + row0 + (int) r) * stride + x0;
#pragma unroll
        for (int q = Q0; q < A; q++) {
            const int ix = 2 * (int) (j + q * Bx) - Fx;
            if (ix >= 0 && ix < bw && !(mask_mode == 1 && el[r0 + ix] < 1)) u[r0 + ix] = add ? u[r0 + ix] + x[q].x : x[q].x;
            if (ix + 1 >= 0 && ix + 1 < bw && !(mask_mode == 1 && el[r0 + ix + 1] < 1))
                u[r0 + ix + 1] = add ? u[r0 + ix + 1] + x[q].y : x[q].y;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// columns
// ------------------------------------------------------------------------------------------------------------
// stage A: item (cc, j), j < By: x[q] = S[col cc][j + By q] (zero beyond n_in rows), radix-Ay, y[k1] *= w_Ly^(j k1),
//          slot[cc][k1 * blky + j]
template <int A, class B>
CB_HD void c2_colA(const ConvPlan &P, B buf, uint32_t oslot, uint32_t otab, uint32_t oScol, int SY, int ncols, int n_in,
                   int lane)
{
    const Conv2Plan &c = P.c2;
    const int By = c.By, items = ncols * By;
    constexpr bool half = DftHalfIn<A, false>::ok;
    constexpr int NL = half ? A / 2 : A;
    for (int i = lane; i < items; i += 32) {
        const uint32_t cc = fdiv((uint32_t) i, c.mg_By), j = (uint32_t) i - cc * By;
        const uint32_t s = oScol + cc * SY + j;
        cd x[A];
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] = (int) (j + q * By) < n_in ? buf.ld(s + q * By) : make_double2(0.0, 0.0);
        if (half) DftHalfIn<A, false>::run(x); else Dft<A, false>::run(x);
        const uint32_t o = oslot + cc * c.collen + j, t = otab + c.o_tay + j;
        buf.st(o, x[0]);
#pragma unroll
        for (int q = 1; q < A; q++) buf.st(o + q * c.blky, cmul(x[q], buf.ld(t + (q - 1) * By)));
    }
}

// stage M: item (cc, k1), k1 < Ay: radix-By forward over the block, multiply by C^[k2][cc][k1] (frequency k1 + Ay k2),
//          radix-By inverse, back in place.  chat_g points at the group's [By][G][Ay] coefficients.
template <int Bq, class B>
CB_HD void c2_colM(const ConvPlan &P, B buf, uint32_t oslot, const cd *chat_g, int ncols, int lane)
{
    const Conv2Plan &c = P.c2;
    const int A = c.Ay, items = ncols * A, GA = c.G * A;
    for (int i = lane; i < items; i += 32) {
        const uint32_t cc = fdiv((uint32_t) i, c.mg_Ay), k1 = (uint32_t) i - cc * A;
        const cd *hp = chat_g + i;                    // (k2 * G + cc) * A + k1 = k2 * G A + i
        cd h[Bq], x[Bq];
#pragma unroll
        for (int q = 0; q < Bq; q++) {
#ifdef __CUDA_ARCH__
            h[q] = __ldg(reinterpret_cast<const double2 *>(hp + q * GA));
#else
            h[q] = hp[q * GA];
#endif
        }
        const uint32_t o = oslot + cc * c.collen + k1 * c.blky;
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

// stage C: item (cc, j): conj twiddle, inverse radix-Ay, rows Fy .. Fy + n_out - 1 (outputs q >= Ay/2) back to S
template <int A, class B>
CB_HD void c2_colC(const ConvPlan &P, B buf, uint32_t oslot, uint32_t otab, uint32_t oScol, int SY, int ncols, int n_out,
                   int lane)
{
    const Conv2Plan &c = P.c2;
    const int By = c.By, items = ncols * By;
    for (int i = lane; i < items; i += 32) {
        const uint32_t cc = fdiv((uint32_t) i, c.mg_By), j = (uint32_t) i - cc * By;
        const uint32_t o = oslot + cc * c.collen + j, t = otab + c.o_tay + j;
        cd x[A];
        x[0] = buf.ld(o);
#pragma unroll
        for (int q = 1; q < A; q++) x[q] = cmulc(buf.ld(o + q * c.blky), buf.ld(t + (q - 1) * By));
        Dft<A, true>::run(x);
        const uint32_t s = oScol + cc * SY + j;
#pragma unroll
        for (int q = A / 2; q < A; q++)
            if ((int) (j + (q - A / 2) * By) < n_out) buf.st(s + (q - A / 2) * By, x[q]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// the three passes of one product as seen by one warp (device: all warps call; host emulation: one call per warp)
// ------------------------------------------------------------------------------------------------------------
template <class B>
CB_HD void c2_rows_fwd(const ConvPlan &P, B buf, const double *base, int bw, int bh, int stride, int warp)
{
    const Conv2Plan &c = P.c2;
    if (warp >= c.nslot) return;
    const uint32_t oS = c.off_S / 16, otab = c.off_tab / 16, oslot = c.off_W / 16 + (uint32_t) (warp * c.slot_len);
    for (int r0 = warp * c.RG; r0 < bh; r0 += c.nslot * c.RG) {
        const int nr = bh - r0 < c.RG ? bh - r0 : c.RG;
        const double *src = base + (size_t) r0 * stride;
#define CB2_CALL_(RR) CB2_LANES((c2_rowf1<RR>(P, buf, oslot, otab, src, nr, bw, stride, lane)))
        CB2_SWITCH_A(c.Ax, CB2_CALL_)
#undef CB2_CALL_
#define CB2_CALL_(RR) CB2_LANES((c2_rowf2<RR>(P, buf, oslot, otab, oS, P.SY, r0, nr, lane)))
        CB2_SWITCH_B(c.Bx, CB2_CALL_)
#undef CB2_CALL_
    }
}

template <class B>
CB_HD void c2_cols(const ConvPlan &P, B buf, const cd *chat, int n_in, int n_out, int warp)
{
    const Conv2Plan &c = P.c2;
    if (warp >= c.nslot) return;
    const uint32_t oS = c.off_S / 16, otab = c.off_tab / 16, oslot = c.off_W / 16 + (uint32_t) (warp * c.slot_len);
    for (int g = warp; g < c.ngrp; g += c.nslot) {
        const int left = P.Fx + 1 - g * c.G, nc = left < c.G ? left : c.G;
        const uint32_t oScol = oS + (uint32_t) (g * c.G * P.SY);
        const cd *chat_g = chat + (size_t) g * c.G * P.Ly;
#define CB2_CALL_(RR) CB2_LANES((c2_colA<RR>(P, buf, oslot, otab, oScol, P.SY, nc, n_in, lane)))
        CB2_SWITCH_A(c.Ay, CB2_CALL_)
#undef CB2_CALL_
#define CB2_CALL_(RR) CB2_LANES((c2_colM<RR>(P, buf, oslot, chat_g, nc, lane)))
        CB2_SWITCH_B(c.By, CB2_CALL_)
#undef CB2_CALL_
#define CB2_CALL_(RR) CB2_LANES((c2_colC<RR>(P, buf, oslot, otab, oScol, P.SY, nc, n_out, lane)))
        CB2_SWITCH_A(c.Ay, CB2_CALL_)
#undef CB2_CALL_
    }
}

template <class B>
CB_HD void c2_rows_inv(const ConvPlan &P, B buf, double *u, const int *el, int mask_mode, int add, int x0, int y0, int bw,
                       int bh, int stride, int warp)
{
    const Conv2Plan &c = P.c2;
    if (warp >= c.nslot) return;
    const uint32_t oS = c.off_S / 16, otab = c.off_tab / 16, oslot = c.off_W / 16 + (uint32_t) (warp * c.slot_len);
    for (int r0 = warp * c.RG; r0 < bh; r0 += c.nslot * c.RG) {
        const int nr = bh - r0 < c.RG ? bh - r0 : c.RG;
#define CB2_CALL_(RR) CB2_LANES((c2_rowi1<RR>(P, buf, oslot, otab, oS, P.SY, r0, nr, lane)))
        CB2_SWITCH_B(c.Bx, CB2_CALL_)
#undef CB2_CALL_
#define CB2_CALL_(RR) CB2_LANES((c2_rowi2<RR>(P, buf, oslot, otab, u, el, mask_mode, add, x0, y0, bw, stride, r0, nr, lane)))
        CB2_SWITCH_A(c.Ax, CB2_CALL_)
#undef CB2_CALL_
    }
}

// position of frequency (kx, ky) inside a coefficient block: [group][k2][column in group][k1], ky = k1 + Ay k2
CB_HD size_t c2_chat_index(const ConvPlan &P, int kx, int ky)
{
    const Conv2Plan &c = P.c2;
    const int g = kx / c.G, cc = kx - g * c.G, k2 = ky / c.Ay, k1 = ky - k2 * c.Ay;
    return ((size_t) (g * c.By + k2) * c.G + cc) * c.Ay + k1;
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

}  // namespace cb200
