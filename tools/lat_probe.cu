// lat_probe.cu -- development tool: dependent-chain latencies of FP64 ops, shared-memory loads and shuffles on one warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lat_probe tools/lat_probe.cu && /tmp/lat_probe
#include <cstdio>
__global__ void k(double *out, long long *t, double x0, int n)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double) ((i * 7 + 1) & 1023);
    __syncthreads();
    double a = x0, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) a = fma(a, b, c);
    long long t1 = clock64();
    for (int i = 0; i < n; i++) a = a + b;
    long long t2 = clock64();
    for (int i = 0; i < n; i++) a = 1.0 / (a + 1.5);
    long long t3 = clock64();
    for (int i = 0; i < n; i++) a = sqrt(a + 2.0);
    long long t4 = clock64();
    int idx = (int) x0 & 1023;
    for (int i = 0; i < n; i++) idx = (int) sm[idx] & 1023;
    long long t5 = clock64();
    for (int i = 0; i < n; i++) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0;
    long long t6 = clock64();
    float f = (float) x0;
    for (int i = 0; i < n; i++) f = fmaf(f, 1.0000001f, 1e-9f);
    long long t7 = clock64();
    if (threadIdx.x == 0) { out[0] = a + idx + f; t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; t[6] = t7 - t6; }
}
int main()
{
    double *o; long long *t, h[7];
    cudaMalloc(&o, 8); cudaMalloc(&t, 56);
    const int n = 4096;
    for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(o, t, 1.25, n);
    cudaMemcpy(h, t, 56, cudaMemcpyDeviceToHost);
    const char *nm[7] = { "DFMA", "DADD", "1/x (+DADD)", "sqrt (+DADD)", "LDS->cvt->and", "SHFL(double)+DADD", "FFMA" };
    for (int i = 0; i < 7; i++) printf("%-20s %.1f cycles per dependent op\n", nm[i], (double) h[i] / n);
    return 0;
}
