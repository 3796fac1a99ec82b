// chat_kernels.cuh -- __global__ entry points that transform influence-coefficient blocks into C^ (one-off per
// grid / material / block; reference: lazy coefficient transforms of fft_VecAijPj, /root/reference/src/m_aijpj.f90:873-920).
#pragma once
#include "device_core.cuh"

namespace cb200 {

// ---- coefficient transform C^ (one CTA, scratch in global memory; runs once per grid/material/block) ----
__global__ void __launch_bounds__(CB_THREADS, 1)
k_build_chat(ConvPlan P, const double *cfblk0, int cmx, int cmy, double scale, cd *SWg0, cd *chat0)
{
    // one CTA per coefficient block: blockIdx.x selects the block, its scratch and its output
    const size_t nscr = (size_t) (P.Lx + 1) * 2 * P.Fy + (size_t) P.Ly * P.C;
    const double *cfblk = cfblk0 + (size_t) blockIdx.x * 4 * cmx * cmy;
    cd *SWg = SWg0 + (size_t) blockIdx.x * nscr;
    cd *chat = chat0 + (size_t) blockIdx.x * P.chat_len;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // only the tables live in shared memory here; reuse the plan's offsets relative to off_twx
    const MemBuf<const cd> twx = { P.twx }, twy = { P.twy };        // tables straight from global (one-off kernel)
    const MemBuf<const unsigned short> posx = { P.posx };
    const int tid = threadIdx.x, nthr = blockDim.x;
    typedef MemBuf<cd> CB_BUF;
    const CB_BUF BUF = { SWg };                                   // scratch: S region, then W region
    const int SY = 2 * P.Fy;
    const uint32_t oS = 0u, oW = (uint32_t) (P.Lx + 1) * SY;
    RowSrc src;
    src.base = cfblk; src.kind = 1;
    src.mx = min(P.Fx, P.mx); src.my = min(P.Fy, P.my);          // m_aijpj.f90:896-898
    src.cmx = cmx; src.cmy = cmy; src.Fx = P.Fx; src.Fy = P.Fy; src.row0 = 0; src.stride = 0;
    CB_CONV_FORWARD_ROWS(2 * P.Fy, src);
    CB_CONV_COLUMNS_DUMP(2 * P.Fy, chat, scale);
}

// ---- coefficient transform in the layout of the warp-resident product: two dense DFT passes over global scratch.
//      blockIdx.y selects the coefficient block (its spatial block, scratch and output follow at fixed strides) ----
__global__ void k_chat2_rows(ConvPlan P, const double *cfblk0, int cmx, int cmy, cd *T0)
{
    const int ld = P.Fx + 1, n = 2 * P.Fy * ld;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int iy = t / ld, kx = t - iy * ld;
    RowSrc src;
    src.base = cfblk0 + (size_t) blockIdx.y * 4 * cmx * cmy; src.kind = 1;
    src.mx = min(P.Fx, cmx); src.my = min(P.Fy, cmy);            // m_aijpj.f90:896-898
    src.cmx = cmx; src.cmy = cmy; src.Fx = P.Fx; src.Fy = P.Fy; src.row0 = 0; src.stride = 0;
    T0[(size_t) blockIdx.y * n + t] = c2_chat_row_entry(P, src, iy, kx, P.twx);
}

__global__ void k_chat2_cols(ConvPlan P, const cd *T0, double scale, cd *chat0)
{
    const int ld = P.Fx + 1, n = 2 * P.Fy * ld;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int ky = t / ld, kx = t - ky * ld;
    chat0[(size_t) blockIdx.y * P.c2.chat_len + c2_chat_index(P, kx, ky)] =
        c2_chat_value(P, T0 + (size_t) blockIdx.y * n, kx, ky, P.twy, scale);
}

}  // namespace cb200
