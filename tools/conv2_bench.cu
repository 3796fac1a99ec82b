// conv2_bench.cu -- development tool (not product, not test): the fused influence product through the block-wide phase
// sequence (conv_sequence.inc) and through the warp-resident passes (fftconv2.cuh) on the same inputs.
// Prints the largest difference between the two results, the device time per product (CUDA events, all SMs busy) and
// the cycles of the three passes of CTA 0.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -lineinfo -o /tmp/conv2_bench tools/conv2_bench.cu
//   /tmp/conv2_bench 91 91 [cases per SM] [box w] [box h]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
// per-stage cycle counters of warp CB2_TW of CTA 0: g_tk[i] accumulates the cycles since the previous tick
__device__ unsigned long long g_tk[16];
#ifndef CB2_TW
#define CB2_TW 0
#endif
#if defined(__CUDA_ARCH__) && !defined(CB2_NOTICK)
// fire-and-forget reductions (RED): the instrumented warp must not wait for global memory, it would become the slowest
#define CB2_TICK_INIT() long long c2_tl_ = clock64()
#define CB2_TICK(i) do { if (blockIdx.x == 0 && threadIdx.x == 32 * CB2_TW) { const long long t_ = clock64(); atomicAdd(&g_tk[i], (unsigned long long) (t_ - c2_tl_)); c2_tl_ = t_; } } while (0)
#else
#define CB2_TICK_INIT()
#define CB2_TICK(i)
#endif
#include "../contact_b200/csrc/plan.h"
#include "../contact_b200/csrc/chat_kernels.cuh"
using namespace cb200;

__device__ unsigned long long g_ph[8];

__global__ void __launch_bounds__(CB_THREADS, 1)
k_prod(const __grid_constant__ ConvPlan P, const double *p, const cd *chat, double *u, const int *el, int mask_mode, int ncase, int bw, int bh)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = smem_view(P, smem_raw);
    smem_load_tables(P, sm);
    for (int ic = blockIdx.x; ic < ncase; ic += gridDim.x)
        conv_box_dev(P, sm, p + (size_t) ic * P.npot, chat, u + (size_t) ic * P.npot, el, mask_mode, 0, 0, 0, bw, bh, P.mx);
}

// the warp-resident passes with a clock after each block barrier
__global__ void __launch_bounds__(CB_THREADS, 1)
k_prod_phases(const __grid_constant__ ConvPlan P, const double *p, const cd *chat, double *u, int ncase)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = smem_view(P, smem_raw);
    smem_load_tables(P, sm);
    const int tid = threadIdx.x, warp = tid >> 5;
    const Conv2Plan &c = P.c2;
    const ShBuf buf;
    {   // tables
        const uint32_t bar = conv_hdr_bar(sm);
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (uint32_t) c.tab_len * 16u);
            bulk_g2s(sm.a0 + (uint32_t) c.off_tab, c.tab, (uint32_t) c.tab_len * 16u, bar);
        }
        mbar_wait(bar, 0u);
    }
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    unsigned long long a0 = 0, a1 = 0, a2 = 0, n = 0;
    for (int ic = blockIdx.x; ic < ncase; ic += gridDim.x) {
        __syncthreads();
        t0 = clock64();
        c2_rows_fwd(P, buf, p + (size_t) ic * P.npot, P.mx, P.my, P.mx, warp);
        __syncthreads();
        t1 = clock64();
        c2_cols(P, buf, chat, P.my, P.my, warp);
        __syncthreads();
        t2 = clock64();
        c2_rows_inv(P, buf, u + (size_t) ic * P.npot, nullptr, 0, 0, 0, 0, P.mx, P.my, P.mx, warp);
        __syncthreads();
        t3 = clock64();
        a0 += t1 - t0; a1 += t2 - t1; a2 += t3 - t2; n++;
    }
    if (tid == 0 && blockIdx.x == 0) { g_ph[0] = a0; g_ph[1] = a1; g_ph[2] = a2; g_ph[3] = n; }
}

__global__ void k_spin(long long cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) { } }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

static int upload_plan(HostPlan &hp)
{
    cd *twx, *twy; unsigned short *posx;
    CK(cudaMalloc(&twx, sizeof(cd) * hp.twx.size())); CK(cudaMalloc(&twy, sizeof(cd) * hp.twy.size()));
    CK(cudaMalloc(&posx, 2 * hp.posx.size()));
    CK(cudaMemcpy(twx, hp.twx.data(), sizeof(cd) * hp.twx.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(twy, hp.twy.data(), sizeof(cd) * hp.twy.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(posx, hp.posx.data(), 2 * hp.posx.size(), cudaMemcpyHostToDevice));
    hp.p.twx = twx; hp.p.twy = twy; hp.p.posx = posx;
    if (hp.p.c2.ok) {
        cd *tab;
        CK(cudaMalloc(&tab, sizeof(cd) * hp.tab2.size()));
        CK(cudaMemcpy(tab, hp.tab2.data(), sizeof(cd) * hp.tab2.size(), cudaMemcpyHostToDevice));
        hp.p.c2.tab = tab;
    }
    return 0;
}

int main(int argc, char **argv)
{
    const int mx = argc > 1 ? atoi(argv[1]) : 91, my = argc > 2 ? atoi(argv[2]) : 91, per_sm = argc > 3 ? atoi(argv[3]) : 8;
    const int bw = argc > 4 ? atoi(argv[4]) : mx, bh = argc > 5 ? atoi(argv[5]) : my;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount, ncase = nsm * per_sm, npot = mx * my;
    HostPlan h2; if (!make_plan(mx, my, h2)) { printf("no plan\n"); return 1; }
    if (argc > 6 && atoi(argv[6]) > 0 && atoi(argv[6]) < h2.p.c2.nslot) h2.p.c2.nslot = atoi(argv[6]);   // experiment: fewer warps at work
    HostPlan h1 = h2; h1.p.c2.ok = 0;                                // the block-wide path with the same layout offsets
    if (upload_plan(h2)) return 1;
    h1.p.twx = h2.p.twx; h1.p.twy = h2.p.twy; h1.p.posx = h2.p.posx;
    const ConvPlan &P2 = h2.p, &P1 = h1.p;
    printf("grid %dx%d F=(%d,%d) c2.ok=%d rows %dx%d cols %dx%d G=%d RG=%d nslot=%d slot=%d smem=%d\n", mx, my, P2.Fx, P2.Fy,
           P2.c2.ok, P2.c2.Ax, P2.c2.Bx, P2.c2.Ay, P2.c2.By, P2.c2.G, P2.c2.RG, P2.c2.nslot, P2.c2.slot_len, P2.smem_bytes);
    // a smooth decaying kernel as coefficient block (2mx x 2my), tractions on a disc
    std::vector<double> cf((size_t) 4 * npot), p((size_t) ncase * npot, 0.0);
    for (int iy = -my; iy < my; iy++) for (int ix = -mx; ix < mx; ix++)
        cf[(size_t) (iy + my) * 2 * mx + ix + mx] = 1.0 / sqrt(0.3 + 0.01 * ix * ix + 0.013 * iy * iy + 0.002 * ix);
    std::vector<int> el(npot, 0);
    srand(7);
    for (int i = 0; i < npot; i++) {
        const int iy = i / mx, ix = i % mx;
        const double rx = (ix - 0.5 * bw) / (0.45 * bw), ry = (iy - 0.5 * bh) / (0.45 * bh);
        el[i] = (ix < bw && iy < bh && rx * rx + ry * ry < 1.0) ? 1 : 0;
    }
    for (size_t i = 0; i < p.size(); i++) p[i] = el[i % npot] ? (rand() / (double) RAND_MAX - 0.3) : 0.0;
    double *d_cf, *d_p, *d_u1, *d_u2; int *d_el; cd *d_c1, *d_c2, *d_scr;
    CK(cudaMalloc(&d_cf, 8 * cf.size())); CK(cudaMalloc(&d_p, 8 * p.size())); CK(cudaMalloc(&d_u1, 8 * p.size()));
    CK(cudaMalloc(&d_u2, 8 * p.size())); CK(cudaMalloc(&d_el, 4 * npot));
    CK(cudaMemcpy(d_cf, cf.data(), 8 * cf.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_p, p.data(), 8 * p.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_el, el.data(), 4 * npot, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_u1, 0, 8 * p.size())); CK(cudaMemset(d_u2, 0, 8 * p.size()));
    const double scale = 1.0 / (4.0 * P2.Fx * P2.Fy);
    // C^ for both layouts
    CK(cudaMalloc(&d_c1, sizeof(cd) * P1.chat_len));
    CK(cudaMalloc(&d_scr, sizeof(cd) * ((size_t) (P1.Lx + 1) * 2 * P1.Fy + (size_t) P1.Ly * P1.C + (size_t) 2 * P1.Fy * (P1.Fx + 1))));
    k_build_chat<<<1, CB_THREADS, 64>>>(P1, d_cf, mx, my, scale, d_scr, d_c1);
    CK(cudaDeviceSynchronize());
    if (P2.c2.ok) {
        const int n = 2 * P2.Fy * (P2.Fx + 1);
        CK(cudaMalloc(&d_c2, sizeof(cd) * P2.c2.chat_len));
        CK(cudaMemset(d_c2, 0, sizeof(cd) * P2.c2.chat_len));
        k_chat2_rows<<<dim3((n + 127) / 128, 1), 128>>>(P2, d_cf, mx, my, d_scr);
        k_chat2_cols<<<dim3((n + 127) / 128, 1), 128>>>(P2, d_scr, scale, d_c2);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaFuncSetAttribute(k_prod, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CK(cudaFuncSetAttribute(k_prod_phases, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms1 = 0, ms2 = 0;
    {   // SM clock actually delivered: 20 M cycles of spinning on every SM against the event clock
        k_spin<<<nsm, 32>>>(2000000); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k_spin<<<nsm, 32>>>(20000000); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms1, e0, e1));
        printf("SM clock under a light load: %.0f MHz (nominal %.0f)\n", 20000000.0 / ms1 * 1e-3, prop.clockRate * 1e-3);
    }
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            k_prod<<<nsm, CB_THREADS, P1.smem_bytes>>>(P1, d_p, d_c1, d_u1, d_el, mode, ncase, bw, bh);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms1, e0, e1));
            if (P2.c2.ok) {
                CK(cudaEventRecord(e0));
                k_prod<<<nsm, CB_THREADS, P2.smem_bytes>>>(P2, d_p, d_c2, d_u2, d_el, mode, ncase, bw, bh);
                CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms2, e0, e1));
            }
        }
        std::vector<double> u1(p.size()), u2(p.size());
        CK(cudaMemcpy(u1.data(), d_u1, 8 * p.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(u2.data(), d_u2, 8 * p.size(), cudaMemcpyDeviceToHost));
        double dmax = 0, umax = 0;
        for (size_t i = 0; i < p.size(); i++) { dmax = fmax(dmax, fabs(u1[i] - u2[i])); umax = fmax(umax, fabs(u1[i])); }
        const double clk = prop.clockRate * 1e3;
        printf("mask_mode %d: block-wide %.3f ms = %.1f us/product/SM (%.0f cycles at %.0f MHz) | warp-resident %.3f ms = %.1f us (%.0f cycles)"
               " | max |du| %.3e of %.3e\n", mode, ms1, 1e3 * ms1 / per_sm, 1e-3 * ms1 / per_sm * clk, clk * 1e-6, ms2,
               1e3 * ms2 / per_sm, 1e-3 * ms2 / per_sm * clk, dmax, umax);
    }
    if (P2.c2.ok && bw == mx && bh == my) {
        k_prod_phases<<<nsm, CB_THREADS, P2.smem_bytes>>>(P2, d_p, d_c2, d_u2, ncase);
        CK(cudaDeviceSynchronize());
        unsigned long long ph[8];
        CK(cudaMemcpyFromSymbol(ph, g_ph, sizeof(ph)));
        printf("warp-resident passes of CTA 0, cycles per product: rows fwd %.0f | columns %.0f | rows inv %.0f\n",
               (double) ph[0] / ph[3], (double) ph[1] / ph[3], (double) ph[2] / ph[3]);
        unsigned long long tk[16];
        CK(cudaMemcpyFromSymbol(tk, g_tk, sizeof(tk)));
        const double np_ = (double) ph[3] + 2.0 * 3 * ((ncase + nsm - 1) / nsm);   // k_prod launches of CTA 0 tick too
        printf("warp %d of CTA 0, cycles per product and stage (all launches): rowf1 %.0f rowf2 %.0f | colA %.0f colM %.0f colC %.0f | rowi1 %.0f rowi2 %.0f | waiting at barriers etc. %.0f\n",
               CB2_TW, tk[1] / np_, tk[2] / np_, tk[4] / np_, tk[5] / np_, tk[6] / np_, tk[8] / np_, tk[9] / np_,
               (tk[0] + tk[3] + tk[7]) / np_);
    }
    return 0;
}
