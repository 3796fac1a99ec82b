// kernels.cuh -- __global__ entry points of the B200 hot path.
#pragma once
#include "chat_kernels.cuh"
#include "norm_solver.cuh"
#include "subsurf.cuh"
#include "tang_solver.cuh"
#include "large_solver.cuh"

namespace cb200 {

// ---- Boussinesq-Cerruti influence coefficients, piecewise-constant elements ----
// Device restatement of the closed forms of elascf_pcwcns (/root/reference/src/m_visc.f90:431-604); one thread per
// offset (ix,iy) in [-mx,mx-1] x [-my,my-1], all nine blocks.  cf layout: [jk][ik][iy+my][ix+mx].
struct ElascfArgs { double akv, nuv, dx, dy, xshft, yshft; int mx, my; };

__device__ __forceinline__ void corner(const ElascfArgs &a, int ix, int iy, double &x, double &y, double &r)
{
    x = (double) ix * a.dx + a.xshft - 0.5 * a.dx;
    y = (double) iy * a.dy + a.yshft - 0.5 * a.dy;
    r = sqrt(x * x + y * y);
}

__global__ void k_elascf_pcwcns(ElascfArgs a, double *cf)
{
    const int mx = a.mx, my = a.my;
    const long nblk = 4L * mx * my;
    const long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
    if (t >= nblk) return;
    const int ix = (int) (t % (2 * mx)) - mx, iy = (int) (t / (2 * mx)) - my;
    const double pi = 3.14159265358979323846;
    const double e1 = (1.0 - a.nuv) / pi, e2 = 1.0 / pi, e3 = a.nuv / pi;
    const double tolx = a.dx * 1e-10, tolx2 = a.dx * 1e-20;
    double c11 = 0, c22 = 0, c33 = 0, c21 = 0, c13 = 0, c23 = 0;
    const int iy0 = (my == 1) ? 0 : -my;
    if (iy >= iy0) {
        double x0, y0, r00, x1, y1, r01, r10, r11, xd, yd;
        corner(a, ix, iy, x0, y0, r00);
        corner(a, ix, iy + 1, xd, y1, r01);
        corner(a, ix + 1, iy, x1, yd, r10);
        corner(a, ix + 1, iy + 1, xd, yd, r11);
        c21 = ((0.0 - e3 * r00) + e3 * r01 + e3 * r10) - e3 * r11;
        if (fabs(a.akv) >= 1e-6) {
            const double l00 = log(r00), l01 = log(r01), l10 = log(r10), l11 = log(r11);
            const double j5_00 = y0 * l00 + x0 * atan(y0 / (x0 + tolx2)), j5_01 = y1 * l01 + x0 * atan(y1 / (x0 + tolx2));
            const double j5_10 = y0 * l10 + x1 * atan(y0 / (x1 + tolx2)), j5_11 = y1 * l11 + x1 * atan(y1 / (x1 + tolx2));
            const double j6_00 = x0 * l00 + y0 * atan(x0 / (y0 + tolx2)), j6_01 = x0 * l01 + y1 * atan(x0 / (y1 + tolx2));
            const double j6_10 = x1 * l10 + y0 * atan(x1 / (y0 + tolx2)), j6_11 = x1 * l11 + y1 * atan(x1 / (y1 + tolx2));
            const double f = e2 * a.akv;
            c13 = ((0.0 - f * j5_00) + f * j5_01 + f * j5_10) - f * j5_11;
            c23 = ((0.0 - f * j6_00) + f * j6_01 + f * j6_10) - f * j6_11;
        }
        // xly2y1(ix,iy), xly2y1(ix+1,iy), ylx2x1(ix,iy), ylx2x1(ix,iy+1)
        const double xl0 = (fmin(fabs(y0 + r00), fabs(y1 + r01)) < tolx) ? 0.0 : x0 * log((y1 + r01) / (y0 + r00));
        const double xl1 = (fmin(fabs(y0 + r10), fabs(y1 + r11)) < tolx) ? 0.0 : x1 * log((y1 + r11) / (y0 + r10));
        const double yl0 = (fmin(fabs(x0 + r00), fabs(x1 + r10)) < tolx) ? 0.0 : y0 * log((x1 + r10) / (x0 + r00));
        const double yl1 = (fmin(fabs(x0 + r01), fabs(x1 + r11)) < tolx) ? 0.0 : y1 * log((x1 + r11) / (x0 + r01));
        c33 = (0.0 - e1 * (xl0 + yl0)) + e1 * (xl1 + yl1);
        c11 = (0.0 - (e1 * xl0 + e2 * yl0)) + (e1 * xl1 + e2 * yl1);
        c22 = (0.0 - (e1 * yl0 + e2 * xl0)) + (e1 * yl1 + e2 * xl1);
    }
    // block (ik,jk) at offset ((jk-1)*3 + (ik-1)) * nblk
    cf[0 * nblk + t] = c11;      // (1,1)
    cf[1 * nblk + t] = c21;      // (2,1)
    cf[2 * nblk + t] = -c13;     // (3,1)
    cf[3 * nblk + t] = c21;      // (1,2)
    cf[4 * nblk + t] = c22;      // (2,2)
    cf[5 * nblk + t] = -c23;     // (3,2)
    cf[6 * nblk + t] = c13;      // (1,3)
    cf[7 * nblk + t] = c23;      // (2,3)
    cf[8 * nblk + t] = c33;      // (3,3)
}

// csv = cs - cv for the tangential displacement rows, csv(3,:) = cs(3,:) (sgencr, m_visc.f90:310-359); n = 4 mx my
__global__ void k_csv_blocks(const double *cs, const double *cv, double *csv, long n)
{
    const long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
    if (t >= 9 * n) return;
    const int blk = (int) (t / n), ik = blk % 3;                 // block index = (jk-1)*3 + (ik-1)
    csv[t] = (ik == 2) ? cs[t] : cs[t] - cv[t];
}

// ---- dense DFT along one axis of a complex (n2 x n1) array (used by the preconditioner builder only) ----
// out[k] = sum_n in[n] tw[(n k) mod N] (sign +1: conj table).  axis 0: along x (fastest), axis 1: along y.
__global__ void k_dft_axis(const cd *in, cd *out, int n1, int n2, int axis, int inverse, const cd *tw)
{
    const long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
    if (t >= (long) n1 * n2) return;
    const int kx = (int) (t % n1), ky = (int) (t / n1);
    const int N = axis == 0 ? n1 : n2, k = axis == 0 ? kx : ky;
    const cd *src = axis == 0 ? in + (size_t) ky * n1 : in + kx;
    const size_t stride = axis == 0 ? 1 : n1;
    double re = 0.0, im = 0.0;
    int idx = 0;
    for (int n = 0; n < N; n++) {
        const cd x = src[n * stride];
        cd w = tw[idx];
        if (inverse) w.y = -w.y;
        re += x.x * w.x - x.y * w.y;
        im += x.x * w.y + x.y * w.x;
        idx += k; if (idx >= N) idx -= N;
    }
    out[t] = make_double2(re, im);
}

__global__ void k_real_to_cplx(const double *in, cd *out, long n)
{
    const long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
    if (t < n) out[t] = make_double2(in[t], 0.0);
}

__global__ void k_cplx_recip(cd *a, long n)
{
    const long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
    if (t < n) { const cd z = a[t]; const double d = z.x * z.x + z.y * z.y; a[t] = make_double2(z.x / d, -z.y / d); }
}

__global__ void k_cplx_real_scaled(const cd *in, double *out, long n, double scale)
{
    const long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
    if (t < n) out[t] = scale * in[t].x;
}

// ---- batched NORM solve: one CTA per contact problem ----
__global__ void __launch_bounds__(CB_THREADS, 1)
k_snorm_batch(const __grid_constant__ ConvPlan P, NormCase *cases, int ncase, int *next_case)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = smem_view(P, smem_raw);
    smem_load_tables(P, sm);
    volatile int *s_case_p = reinterpret_cast<volatile int *>(sm.red + 127);   // last slot of the reduction scratch
    const long long t_in = clock64();
    // dynamic case queue: iteration counts differ between cases, so CTAs pull the next case when they finish one
    for (;;) {
        if (threadIdx.x == 0) *s_case_p = atomicAdd(next_case, 1);
        __syncthreads();
        const int ic = *s_case_p;
        __syncthreads();
        if (ic >= ncase) break;
        snorm_dev(P, sm, cases[ic]);
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) g_conv_prof[2] += (unsigned long long) (clock64() - t_in);
}

// ---- batched contact cases (NORM + TANG alternation), one CTA per case, dynamic queue ----
__global__ void __launch_bounds__(CB_THREADS, 1)
k_contac_batch(const __grid_constant__ ConvPlan P, ContactCase *cases, int ncase, int *next_case)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = smem_view(P, smem_raw);
    smem_load_tables(P, sm);
    volatile int *s_case_p = reinterpret_cast<volatile int *>(sm.red + 127);
    for (;;) {
        if (threadIdx.x == 0) *s_case_p = atomicAdd(next_case, 1);
        __syncthreads();
        const int ic = *s_case_p;
        __syncthreads();
        if (ic >= ncase) break;
        panprc_dev(BlockCtx{ P, sm }, cases[ic]);
        __syncthreads();
    }
}

// ---- FP64 FMA throughput probe (roofline denominator for the FP64-bound batched small-grid products) ----
__global__ void k_fp64_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace cb200
