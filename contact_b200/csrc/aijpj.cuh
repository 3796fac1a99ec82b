// aijpj.cuh -- direct row sum of the influence product: u_i = (1/G) sum_jk sum_j cf(i - j, ik, jk) p_j(jk), the reference's
// gf3_AijPj (/root/reference/src/m_aijpj.f90:99-254).  It is the O(ncon) twin of the FFT product for single elements (the
// reference uses it for point selections and inside the Gauss-Seidel solvers) and the independent check of the FFT path.
//
// One CTA per requested element.  Warp w takes the grid rows jy = w, w + nw, ...; its lanes cover the row's column range
// [row1st(jy) - 1, rowlst(jy) + 1] -- the elements in contact plus one neighbour either side, as the reference (:170-180; the
// range comes from the element division of p) -- coalesced along x, one fixed-tree shuffle sum per (row, direction).  The
// per-row sums go to shared memory and are added by one thread in the reference's order (jy outer, jk inner), so the result
// does not depend on the number of warps.
#pragma once
#include "device_core.cuh"

namespace cb200 {

// cf: the 9 spatial blocks [jk][ik][2 my][2 mx] of one coefficient set (x G, as the reference stores them); p: [3][npot];
// el: [npot]; ii: 0-based element indices; out[k] = row sum of element ii[k]
__global__ void k_aijpj(int mx, int my, const double *__restrict__ cf, double ga_inv, int ik, int jk0, int jk1,
                        const double *__restrict__ p, const int *__restrict__ el, const int *__restrict__ ii, int npts,
                        double *__restrict__ out)
{
    extern __shared__ double rows[];                                 // [my][3] per-row sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t npot = (size_t) mx * my, nblk = 4 * npot;
    for (int k = blockIdx.x; k < npts; k += gridDim.x) {
        const int ix = ii[k] % mx, iy = ii[k] / mx;
        for (int jy = warp; jy < my; jy += nw) {
            // row1st / rowlst of the element division (m_gridfunc.f90 areas: first / last element with el >= Adhes)
            int first = mx, last = -1;
            for (int jx = lane; jx < mx; jx += 32)
                if (el[(size_t) jy * mx + jx] >= 1) { first = min(first, jx); last = max(last, jx); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
                last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
            }
            const int j0 = max(0, first - 1), j1 = last < 0 ? -1 : min(mx - 1, last + 1);
            for (int jk = jk0; jk <= jk1; jk++) {
                const double *blk = cf + (size_t) ((ik - 1) + 3 * (jk - 1)) * nblk + (size_t) (iy - jy + my) * (2 * mx) + mx + ix;
                const double *pj = p + (size_t) (jk - 1) * npot + (size_t) jy * mx;
                double s = 0.0;
                for (int jx = j0 + lane; jx <= j1; jx += 32) s += blk[-jx] * pj[jx];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) rows[jy * 3 + (jk - 1)] = s;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double part = 0.0;
            for (int jy = 0; jy < my; jy++) for (int jk = jk0; jk <= jk1; jk++) part += rows[jy * 3 + (jk - 1)];
            out[k] = part * ga_inv;
        }
        __syncthreads();
    }
}

}  // namespace cb200
