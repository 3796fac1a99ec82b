// phase_timer.cu -- development tool (not product, not test): runs the fused convolution with a cycle counter after
// every phase barrier and prints the per-phase cycle counts of CTA 0.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/phase_timer tools/phase_timer.cu && /tmp/phase_timer 91 91 148
#include <cstdio>
#include <vector>
#include "../contact_b200/csrc/plan.h"
#include "../contact_b200/csrc/fftconv_warp.cuh"
using namespace cb200;

#ifndef CB_THREADS
#define CB_THREADS 384
#endif
__device__ long long g_t[512];
__device__ int g_n;
#define CB_PHASE(call) do { call; __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0 && np_ < 512) { tt_[np_++] = clock64(); } } while (0)
#include "../contact_b200/csrc/conv_sequence.inc"

__global__ void __launch_bounds__(CB_THREADS, 1)
k_prof(ConvPlan P, const double *p, const cd *chat, double *u, int ncase)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t a0 = (uint32_t) __cvta_generic_to_shared(smem_raw);
    typedef MemBuf<cd> CB_BUF;
    const CB_BUF BUF = { reinterpret_cast<cd *>(__cvta_shared_to_generic(a0)) };
    const uint32_t oS = P.off_S / 16, oW = P.off_W / 16;
    const MemBuf<const cd> twx = { reinterpret_cast<const cd *>(__cvta_shared_to_generic(a0 + P.off_twx)) };
    const MemBuf<const cd> twy = { reinterpret_cast<const cd *>(__cvta_shared_to_generic(a0 + P.off_twy)) };
    const MemBuf<const unsigned short> posx = { reinterpret_cast<const unsigned short *>(__cvta_shared_to_generic(a0 + P.off_posx)) };
    cd *ptx = reinterpret_cast<cd *>(smem_raw + P.off_twx), *pty = reinterpret_cast<cd *>(smem_raw + P.off_twy);
    unsigned short *pp = reinterpret_cast<unsigned short *>(smem_raw + P.off_posx);
    for (int k = threadIdx.x; k < 2 * P.Fx; k += blockDim.x) ptx[k] = P.twx[k];
    for (int k = threadIdx.x; k < 2 * P.Fy; k += blockDim.x) pty[k] = P.twy[k];
    for (int k = threadIdx.x; k < P.Lx; k += blockDim.x) pp[k] = P.posx[k];
    __syncthreads();
    const int tid = threadIdx.x, nthr = blockDim.x, SY = P.SY;
    long long tt_[128]; int np_ = 0;
    for (int ic = blockIdx.x; ic < ncase; ic += gridDim.x) {
        np_ = 0;
        if (tid == 0 && blockIdx.x == 0) tt_[np_++] = clock64();
        RowSrc src; src.base = p + (size_t) ic * P.npot; src.kind = 0; src.mx = P.mx; src.my = P.my; src.cmx = 0; src.cmy = 0; src.Fx = P.Fx; src.Fy = P.Fy; src.row0 = 0; src.stride = 0;
#ifdef PT_WARP
        // warp-scheduled product (fftconv_warp.cuh): three passes, one barrier after each
        const int warp = tid >> 5, nwarps = nthr >> 5;
        CB_PHASE(warp_rows_fwd(P, BUF, oS, SY, src.base, P.mx, P.my, P.mx, twx, posx, warp, nwarps));
        CB_PHASE(warp_cols(P, BUF, oS, oW, SY, P.my, P.my, chat, twy, warp, nwarps));
        CB_PHASE(warp_rows_inv(P, BUF, oS, SY, u + (size_t) ic * P.npot, (const int *) nullptr, 0, 0, 0, 0, P.mx, P.my, P.mx, twx, posx, warp, nwarps));
#else
        CB_CONV_FORWARD_ROWS(P.my, src);
        CB_CONV_COLUMNS_PRODUCT(P.my, chat);
        CB_CONV_INVERSE_ROWS_STORE(P.my, u + (size_t) ic * P.npot, (const int *) nullptr, 0, 0, 0, 0, P.mx, P.mx);
#endif
    }
    if (tid == 0 && blockIdx.x == 0) { for (int i = 0; i < np_; i++) g_t[i] = tt_[i]; g_n = np_; }
}

int main(int argc, char **argv)
{
    int mx = argc > 1 ? atoi(argv[1]) : 91, my = argc > 2 ? atoi(argv[2]) : 91, ncase = argc > 3 ? atoi(argv[3]) : 148;
    HostPlan hp; make_plan(mx, my, hp);
    ConvPlan &P = hp.p;
    cd *twx, *twy; unsigned short *posx; double *p, *u; cd *chat;
    cudaMalloc(&twx, 16 * hp.twx.size()); cudaMalloc(&twy, 16 * hp.twy.size()); cudaMalloc(&posx, 2 * hp.posx.size());
    cudaMemcpy(twx, hp.twx.data(), 16 * hp.twx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(twy, hp.twy.data(), 16 * hp.twy.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(posx, hp.posx.data(), 2 * hp.posx.size(), cudaMemcpyHostToDevice);
    P.twx = twx; P.twy = twy; P.posx = posx;
    cudaMalloc(&p, 8L * ncase * P.npot); cudaMalloc(&u, 8L * ncase * P.npot); cudaMalloc(&chat, 16L * P.chat_len);
    cudaMemset(p, 0, 8L * ncase * P.npot); cudaMemset(chat, 0, 16L * P.chat_len);
    cudaFuncSetAttribute(k_prof, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    for (int rep = 0; rep < 2; rep++) k_prof<<<ncase < 148 ? ncase : 148, CB_THREADS, P.smem_bytes>>>(P, p, chat, u, ncase);
    cudaError_t e = cudaDeviceSynchronize();
    {   // whole-kernel time per product (all CTAs busy), CUDA events
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int big = 148 * 8;
        double *pb, *ub; cudaMalloc(&pb, 8L * big * P.npot); cudaMalloc(&ub, 8L * big * P.npot); cudaMemset(pb, 0, 8L * big * P.npot);
        k_prof<<<148, CB_THREADS, P.smem_bytes>>>(P, pb, chat, ub, big);
        cudaEventRecord(e0);
        k_prof<<<148, CB_THREADS, P.smem_bytes>>>(P, pb, chat, ub, big);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("kernel: %d products in %.3f ms = %.2f us per product per SM\n", big, ms, 1e3 * ms / 8);
    }
    printf("status %s; plan Fx %d Fy %d C %d nchunk %d rx", cudaGetErrorString(e), P.Fx, P.Fy, P.C, P.nchunk);
    for (int i = 0; i < P.nsx; i++) printf(" %d", P.rx[i]);
    printf(" ry");
    for (int i = 0; i < P.nsy; i++) printf(" %d", P.ry[i]);
    printf("\n");
    long long t[512]; int n;
    cudaMemcpyFromSymbol(t, g_t, sizeof(t)); cudaMemcpyFromSymbol(&n, g_n, sizeof(int));
    printf("phases %d total cycles %lld\n", n - 1, t[n - 1] - t[0]);
    for (int i = 1; i < n; i++) printf("%lld ", t[i] - t[i - 1]);
    printf("\n");
    return 0;
}
