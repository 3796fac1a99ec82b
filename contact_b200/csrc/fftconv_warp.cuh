// fftconv_warp.cuh -- warp-scheduled form of the fused influence product (same butterflies, same arithmetic order and
// therefore bit-identical results as the block-wide phase sequence of conv_sequence.inc; only WHO does WHAT WHEN differs).
//
// Why: in the block-wide sequence every FFT stage is one phase of the whole CTA followed by __syncthreads().  A 91x91
// product has 20 such phases with 300 - 1100 butterflies each: at most one radix-16 butterfly per thread, so all 12
// warps load, then all compute, then all store, in lock step -- the shared-memory pipe and the FP64 pipe are never busy
// at the same time and every phase pays a barrier (tools/phase_timer.cu: 3.4 - 6.2 k cycles per phase, 84 k per product
// against an FP64 issue floor of ~25 k).  Here a transform never leaves its warp:
//   * forward rows : a warp owns GR grid rows: load, all DIF stages, split step -- only __syncwarp() in between;
//   * columns      : a warp owns G spectrum columns at a time in a private W slot: first stage (zero-padded read of S),
//                    middle stages, multiply by C^ in registers, inverse stages, pruned write back to S;
//   * inverse rows : merge step, DIT stages, masked store of the result rows -- again per warp.
// Three block barriers per product remain (rows -> columns -> rows -> caller); between them the warps drift apart, so
// the loads of one warp overlap the FP64 chain of another.
// Reference semantics unchanged: fft_VecAijPj, /root/reference/src/m_aijpj.f90:712-1015.
//
// Host emulation (tests/host_emul): the same functions compile for the host, where a "warp" is a loop over 32 lanes per
// stage (CB_LANES) -- the test drives one call per warp.
#pragma once
#include "fftconv.cuh"

namespace cb200 {

#ifdef __CUDA_ARCH__
#define CB_LANES(call) do { { const int lane = (int) (threadIdx.x & 31u); call; } __syncwarp(); } while (0)
#else
#define CB_LANES(call) do { for (int lane = 0; lane < 32; lane++) { call; } } while (0)
#endif

// W slot of one warp: element e of column c at a = e*nb + c, one spare element after every 8 so that the stride between
// the butterflies of the last stages (R*nb elements, a multiple of 8 for R = 8, 12, 16 and even nb) does not put all
// lanes on the same banks
template <class B> struct ViewSk {
    B buf; uint32_t off, nb;
    CB_HD uint32_t at(uint32_t e, uint32_t c) const { const uint32_t a = e * nb + c; return off + a + (a >> 3); }
    CB_HD cd ld(uint32_t e, uint32_t c) const { return buf.ld(at(e, c)); }
    CB_HD void st(uint32_t e, uint32_t c, cd v) const { buf.st(at(e, c), v); }
};

// rows per warp group: enough groups for all warps, a power of two <= 8 (full lanes in the row stages)
CB_HD int warp_row_group(int nrows, int nwarps)
{
    int g = 1;
    while (g < 8 && g * nwarps < nrows) g *= 2;
    return g;
}

// ---- forward rows of the tractions p (box bw x bh at `base`, row stride `stride`) -> S[kx][iy] ----
template <class B, class TW, class PX>
CB_HD void warp_rows_fwd(const ConvPlan &P, B buf, uint32_t oS, int SY, const double *base, int bw, int bh, int stride,
                         TW twx, PX posx, int warp, int nwarps)
{
    const int GR = warp_row_group(bh, nwarps);
    const uint32_t mg_gr = div_magic((uint32_t) GR);
    for (int b0 = warp * GR; b0 < bh; b0 += nwarps * GR) {
        const int nb = bh - b0 < GR ? bh - b0 : GR;
        const uint32_t mg_b = nb == GR ? mg_gr : div_magic((uint32_t) nb);
        RowSrc src;
        src.base = base + (size_t) b0 * stride; src.kind = 0; src.mx = bw; src.my = nb; src.cmx = 0; src.cmy = 0;
        src.Fx = P.Fx; src.Fy = P.Fy; src.row0 = 0; src.stride = stride;
        const uint32_t o = oS + (uint32_t) b0;
        const ViewLin<B> vS = { buf, o, (uint32_t) SY };
        CB_LANES(row_load(P, buf, o, SY, nb, src, lane, 32));
        for (int s = 0; s < P.nsx; s++)
            CB_LANES(fft_stage_k<false>(P.rx[s], vS, vS, nb, mg_b, P.kx[s], twx, lane, 32));
        CB_LANES(row_split(P, buf, o, SY, nb, twx, posx, lane, 32));
    }
}

// ---- columns: forward stages, multiply by C^ (layout [chunk][Ly][C]), inverse stages; rows Fy..Fy+n_out-1 back to S ----
template <class B, class TW>
CB_HD void warp_cols(const ConvPlan &P, B buf, uint32_t oS, uint32_t oW, int SY, int n_in, int n_out, const cd *chat,
                     TW twy, int warp, int nwarps)
{
    const int wslots = P.wslots < nwarps ? P.wslots : nwarps;
    if (warp >= wslots) return;
    const int G = P.G, ngid = P.nchunk * P.gpc, nsy = P.nsy;
    const uint32_t oWw = oW + (uint32_t) (warp * P.wslot_len);
    for (int gid = warp; gid < ngid; gid += wslots) {
        const int ch = (int) fdiv((uint32_t) gid, P.mg_gpc), c0 = (gid - ch * P.gpc) * G;
        const int left = P.Fx + 1 - ch * P.C;
        const int ncch = left < P.C ? left : P.C;
        if (c0 >= ncch) continue;
        const int nb = ncch - c0 < G ? ncch - c0 : G;
        const uint32_t mg_b = nb == G ? P.mg_G : div_magic((uint32_t) nb);
        const uint32_t ocol = oS + (uint32_t) ((ch * P.C + c0) * SY);
        const cd *chat_ = chat + (size_t) ch * P.Ly * P.C + c0;
        const ViewPad<B> vin = { buf, ocol, (uint32_t) SY, (uint32_t) n_in };
        const ViewCrop<B> vout = { buf, ocol, (uint32_t) SY, (uint32_t) P.Fy, (uint32_t) n_out };
        const ViewSk<B> vW = { buf, oWw, (uint32_t) nb };
        if (nsy == 1) {
            CB_LANES(fft_stage_mid_k(P.ry[0], vin, vout, nb, mg_b, P.ky[0], chat_, (uint32_t) P.C, lane, 32));
        } else {
            CB_LANES(fft_stage_k<false>(P.ry[0], vin, vW, nb, mg_b, P.ky[0], twy, lane, 32));
            for (int s = 1; s < nsy - 1; s++)
                CB_LANES(fft_stage_k<false>(P.ry[s], vW, vW, nb, mg_b, P.ky[s], twy, lane, 32));
            CB_LANES(fft_stage_mid_k(P.ry[nsy - 1], vW, vW, nb, mg_b, P.ky[nsy - 1], chat_, (uint32_t) P.C, lane, 32));
            for (int s = nsy - 2; s >= 1; s--)
                CB_LANES(fft_stage_k<true>(P.ry[s], vW, vW, nb, mg_b, P.ky[s], twy, lane, 32));
            CB_LANES(fft_stage_k<true>(P.ry[0], vW, vout, nb, mg_b, P.ky[0], twy, lane, 32));
        }
    }
}

// ---- inverse rows and masked store of the box (x0, y0, bw x bh) of u / el (row stride `stride`) ----
template <class B, class TW, class PX>
CB_HD void warp_rows_inv(const ConvPlan &P, B buf, uint32_t oS, int SY, double *u, const int *el, int mask_mode, int add,
                         int x0, int y0, int bw, int bh, int stride, TW twx, PX posx, int warp, int nwarps)
{
    const int GR = warp_row_group(bh, nwarps);
    const uint32_t mg_gr = div_magic((uint32_t) GR);
    for (int b0 = warp * GR; b0 < bh; b0 += nwarps * GR) {
        const int nb = bh - b0 < GR ? bh - b0 : GR;
        const uint32_t mg_b = nb == GR ? mg_gr : div_magic((uint32_t) nb);
        const uint32_t o = oS + (uint32_t) b0;
        const ViewLin<B> vS = { buf, o, (uint32_t) SY };
        CB_LANES(row_merge(P, buf, o, SY, nb, twx, posx, lane, 32));
        for (int s = P.nsx - 1; s >= 0; s--)
            CB_LANES(fft_stage_k<true>(P.rx[s], vS, vS, nb, mg_b, P.kx[s], twx, lane, 32));
        CB_LANES(row_store_box(P, buf, o, SY, u, el, mask_mode, add, x0, y0 + b0, bw, nb, stride, lane, 32));
    }
}

}  // namespace cb200
