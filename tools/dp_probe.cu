// dp_probe.cu -- development tool: FP64 latency/throughput on one SM at low occupancy (clock64 around FMA chains).
#include <cstdio>
template <int ILP>
__global__ void k(double *out, long long *cyc, int iters)
{
    double a[ILP];
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int threads, double *out, long long *cyc)
{
    const int iters = 4096;
    k<ILP><<<1, threads>>>(out, cyc, iters);
    k<ILP><<<1, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %4d ILP %2d: %.2f cycles per FMA-instr per warp (warp0), SM rate %.1f DFMA lanes/clk\n", threads, ILP,
           (double) c / (iters * ILP), (double) threads * iters * ILP / c);
}
int main()
{
    double *out; long long *cyc; cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 8);
    for (int th : {32, 128, 384, 512, 1024}) { run<1>(th, out, cyc); run<2>(th, out, cyc); run<4>(th, out, cyc); run<8>(th, out, cyc); run<16>(th, out, cyc); }
    return 0;
}
