// subsurf.cuh -- subsurface displacements / stresses on the device.
//
// Mirrors the reference's subsurface evaluator (/root/reference/src/m_subsurf.f90): per depth z four 3x3 sets of
// influence coefficients (displacement and its x-, y-, z-derivatives) from Kalker's closed forms (stres1_pcwcns
// :1633-1852, sstres_inflcf :1261-1408), 36 influence products with the surface tractions (sstres_fft :1097-1257),
// then strains -> Hooke stresses -> invariants and principal stresses per point (sstres_derived :1519-1629).
// B200-first: coefficients and their transforms are built on the device once per (grid, material, z) and shared by
// the whole batch of cases; the 36 products per (case, depth) run back to back inside one persistent CTA through the
// same shared-memory FFT convolution as the contact solvers, followed by the fused per-point stress evaluation.
// ISUBS = 9 (arbitrary points) uses the direct sum (sstres :1412-1515), one thread per point.
#pragma once
#include "device_core.cuh"

namespace cb200 {

// v[i][j][k]: k = 0 displacement u_j, k = 1..3 gradient u_{j,k} due to a unit load in direction i on the rectangle
// dx x dy centred at xp; vnu: the part multiplied by Poisson's ratio.  Kalker (1986), Comm.Appl.Num.Meth. 2, 401-410.
__device__ void stres1_dev(double dx, double dy, double gg, double (&v)[3][3][4], double (&vnu)[3][3][4],
                           const double (&xw)[3], const double (&xp)[2])
{
    const double epsrel = 5e-7, pi = 3.14159265358979323846;
    const int sgn[3] = { -1, -1, 1 }, ip[3] = { 1, 2, 0 };
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 4; k++) { v[i][j][k] = 0.0; vnu[i][j][k] = 0.0; }
    for (int corner = 0; corner < 4; corner++) {
        const int jx = (corner & 1) ? 1 : -1, jy = (corner & 2) ? 1 : -1;
        double y[3], yeps[3], al[3], at[3], wm[4], a[3][3][4], t[3][3][4];
        y[0] = xp[0] + jx * dx / 2.0 - xw[0];
        y[1] = xp[1] + jy * dy / 2.0 - xw[1];
        y[2] = xw[2];
        const double w = fmax(1e-12, sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]));
        const double epsy = epsrel * w;
#pragma unroll
        for (int i = 0; i < 3; i++) yeps[i] = (y[i] >= 0.0) ? fmax(epsy, y[i]) : fmin(-epsy, y[i]);
        const double weps = fmax(1e-12, sqrt(yeps[0] * yeps[0] + yeps[1] * yeps[1] + yeps[2] * yeps[2]));
        wm[0] = weps;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            al[k] = log(y[k] + weps);
            at[k] = atan((y[ip[k]] + y[ip[ip[k]]] + w) / yeps[k]);
            wm[k + 1] = sgn[k] * y[k] / weps;
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                a[i][j][0] = y[i] * al[j];
                t[i][j][0] = y[i] * at[j];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double kik = (i == k) ? 1.0 : 0.0, kjk = (j == k) ? 1.0 : 0.0;
                    a[i][j][k + 1] = sgn[k] * (kik * al[j] + y[i] * (kjk * w + y[k]) / (weps * (y[j] + weps)));
                    t[i][j][k + 1] = sgn[k] * (kik * at[j] + y[i] * (y[j] * (y[k] + w) - kjk * w * (y[0] + y[1] + y[2] + w))
                                                              / (2 * weps * (y[ip[j]] + weps) * (y[ip[ip[j]]] + weps)));
                }
            }
        const double s = (double) (jx * jy);
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int l = 1 - i;
                double q = a[i][l][k] + 2 * a[l][i][k] + 4 * t[2][2][k] + a[i][l][k] - 2 * t[2][i][k];
                v[i][i][k] += s * q;
                q = -2 * (a[i][l][k] - 2 * t[2][i][k]);
                vnu[i][i][k] += q * s;
                q = -a[2][2][k];
                v[i][l][k] += s * q;
                q = 2 * (a[2][2][k] - wm[k]);
                vnu[i][l][k] += q * s;
                q = 2 * (a[l][2][k] + a[2][l][k] + 2 * t[i][i][k]);
                vnu[i][2][k] += q * s;
                vnu[2][i][k] = -vnu[i][2][k];
                q = -a[l][2][k] - 2 * t[i][i][k];
                v[i][2][k] += q * s;
                v[2][i][k] += (2 * a[2][l][k] - q) * s;
            }
            double q = -2 * (a[0][1][k] + a[1][0][k] + 2 * t[2][2][k]);
            vnu[2][2][k] += q * s;
            q = -2 * t[2][2][k] + 2 * (a[0][1][k] + a[1][0][k] + 2 * t[2][2][k]);
            v[2][2][k] += q * s;
        }
    }
    const double f = 1.0 / (4 * pi * gg);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 4; k++) { v[i][j][k] *= f; vnu[i][j][k] *= f; }
}

// Coefficient arrays ck(k)%cf(ix, iy, ik, jk), k = 0..3, for depth zw: cf layout [k][jk][ik][iy+my][ix+mx].
// The reference evaluates one quadrant and mirrors with signs (m_subsurf.f90:1314-1389); here every offset is
// evaluated at its quadrant image (-|ix|, -|iy|) and multiplied by the same sign products.
__global__ void k_subsurf_coef(int mx, int my, double dx, double dy, double gg, double poiss, double zw, int neg, double *cf)
{
    const long nblk = 4L * mx * my;
    const long tq = blockIdx.x * (long) blockDim.x + threadIdx.x;
    if (tq >= nblk) return;
    const int ix = (int) (tq % (2 * mx)) - mx, iy = (int) (tq / (2 * mx)) - my;
    const int sx = ix > 0, sy = iy > 0;
    double v[3][3][4], vnu[3][3][4];
    const double xw[3] = { (double) (-abs(ix)) * dx, neg * (double) (-abs(iy)) * dy, neg * zw }, xp[2] = { 0.0, 0.0 };
    stres1_dev(dx, dy, gg, v, vnu, xw, xp);
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < 3; i++)          // i: load direction = jk-1
            for (int j = 0; j < 3; j++) {    // j: displacement direction = ik-1
                int sign = 1;
                if (sx) sign *= ((j == 0) ? -1 : 1) * ((i == 0) ? -1 : 1) * ((k == 1) ? -1 : 1);
                if (sy) sign *= ((j == 1) ? -1 : 1) * ((i == 1) ? -1 : 1) * ((k == 2) ? -1 : 1);
                cf[((size_t) k * 9 + (size_t) i * 3 + j) * nblk + tq] = sign * (v[i][j][k] + poiss * vnu[i][j][k]);
            }
}

// principal stresses: trigonometric roots of the characteristic cubic of a symmetric 3x3 matrix, descending
__device__ __forceinline__ void sym3_eigenvalues(const double (&s)[3][3], double (&ev)[3])
{
    const double q = (s[0][0] + s[1][1] + s[2][2]) / 3.0;
    const double p1 = s[0][1] * s[0][1] + s[0][2] * s[0][2] + s[1][2] * s[1][2];
    const double d0 = s[0][0] - q, d1 = s[1][1] - q, d2 = s[2][2] - q;
    const double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
    if (p2 <= 0.0) { ev[0] = ev[1] = ev[2] = q; return; }
    const double p = sqrt(p2 / 6.0);
    const double b00 = d0 / p, b11 = d1 / p, b22 = d2 / p, b01 = s[0][1] / p, b02 = s[0][2] / p, b12 = s[1][2] / p;
    double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) / 3.0;
    ev[0] = q + 2.0 * p * cos(phi);
    ev[2] = q + 2.0 * p * cos(phi + 2.0 * 3.14159265358979323846 / 3.0);
    ev[1] = 3.0 * q - ev[0] - ev[2];
}

// sstres_derived (m_subsurf.f90:1519-1629): vr[j][k] -> out[18] = uw(3), sighyd, sigvm, sigtr, sigmaj(3), sigma(3,3)
__device__ __forceinline__ void sstres_derived_dev(double gg, double poiss, int neg, double (&vr)[3][4], double *out)
{
    const double tolsml = 1e-15;
    double er[3][3], sigma[3][3], sigmaj[3], uw[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) vr[j][k] = ((j >= 1) ? neg : 1) * ((k >= 2) ? neg : 1) * vr[j][k];
#pragma unroll
    for (int i = 0; i < 3; i++) uw[i] = vr[i][0];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) er[i][j] = (vr[i][j + 1] + vr[j][i + 1]) / 2.0;
    const double dil = er[0][0] + er[1][1] + er[2][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) sigma[i][j] = 2.0 * gg * er[i][j];
#pragma unroll
    for (int i = 0; i < 3; i++) sigma[i][i] = sigma[i][i] + 2.0 * gg * dil * poiss / fmax(1e-6, 1.0 - 2.0 * poiss);
    const double sigii = sigma[0][0] + sigma[1][1] + sigma[2][2];
    double sijsij = -(1.0 / 3.0) * sigii * sigii;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) sijsij += sigma[i][j] * sigma[i][j];
    sijsij *= 0.5;
    sym3_eigenvalues(sigma, sigmaj);
#pragma unroll
    for (int i = 0; i < 3; i++) if (fabs(uw[i]) < gg * tolsml) uw[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) if (fabs(sigma[i][j]) < gg * tolsml) sigma[i][j] = 0.0;
    out[0] = uw[0]; out[1] = uw[1]; out[2] = uw[2];
    out[3] = sigii / 3.0; out[4] = sqrt(3.0 * sijsij); out[5] = sigmaj[0] - sigmaj[2];
    out[6] = sigmaj[0]; out[7] = sigmaj[1]; out[8] = sigmaj[2];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) out[9 + j * 3 + i] = sigma[i][j];
}

struct SubsArgs {
    int ncase, nz, neg_mask;        // neg_mask bit iz: depth iz lies in the lower body (z < 0)
    const double *ps;               // [ncase][3][npot] surface tractions
    const cd *chat;                 // [nz][4][9][chat_len] transformed coefficients, block index (jk-1)*3 + (ik-1)
    double *vr;                     // scratch [gridDim][13][npot]: 12 products + sign-flipped px
    double *table;                  // [ncase][nz][npot][18]
    double gg[2], poiss[2];
    int *next;
    long chat_len;
};

// one CTA per (case, depth) work item, dynamic queue
__global__ void __launch_bounds__(CB_THREADS, 1)
k_subsurf_batch(const __grid_constant__ ConvPlan P, SubsArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = smem_view(P, smem_raw);
    smem_load_tables(P, sm);
    volatile int *s_item = reinterpret_cast<volatile int *>(sm.red + 127);
    const int n = P.npot;
    double *vr = A.vr + (size_t) blockIdx.x * 13 * n;
    double *pxneg = vr + (size_t) 12 * n;
    for (;;) {
        if (threadIdx.x == 0) *s_item = atomicAdd(A.next, 1);
        __syncthreads();
        const int item = *s_item;
        __syncthreads();
        if (item >= A.ncase * A.nz) break;
        const int ic = item / A.nz, iz = item % A.nz;
        const int neg = ((A.neg_mask >> iz) & 1) ? -1 : 1, ia = neg < 0 ? 1 : 0;
        const double *ps = A.ps + (size_t) ic * 3 * n;
        if (neg < 0) {                                              // m_subsurf.f90:1170 (on a copy: the caller's ps is kept)
            for (int i = threadIdx.x; i < n; i += blockDim.x) pxneg[i] = -ps[i];
            __syncthreads();
        }
        for (int k = 0; k < 4; k++)
            for (int ik = 0; ik < 3; ik++)
                for (int jk = 0; jk < 3; jk++) {
                    const double *pj = (jk == 0 && neg < 0) ? pxneg : ps + (size_t) jk * n;
                    const cd *ch = A.chat + (((size_t) iz * 4 + k) * 9 + (size_t) jk * 3 + ik) * A.chat_len;
                    conv_dev(P, sm, pj, ch, vr + (size_t) (k * 3 + ik) * n, nullptr, 0, jk > 0 ? 1 : 0);
                }
        double *tbl = A.table + ((size_t) ic * A.nz + iz) * n * 18;
        for (int ii = threadIdx.x; ii < n; ii += blockDim.x) {
            double v[3][4];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int k = 0; k < 4; k++) v[j][k] = vr[(size_t) (k * 3 + j) * n + ii];
            sstres_derived_dev(A.gg[ia], A.poiss[ia], neg, v, tbl + (size_t) ii * 18);
        }
        __syncthreads();
    }
}

// ISUBS = 9: direct sum in arbitrary points (sstres, m_subsurf.f90:1412-1515); one thread per point
__global__ void k_subsurf_points(int mx, int my, double xc1, double yc1, double dx, double dy, double gg0, double gg1,
                                 double poiss0, double poiss1, const double *ps, const double *xyz, int npoint, double *table)
{
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npoint) return;
    const int npot = mx * my;
    double xw[3] = { xyz[3 * ip], xyz[3 * ip + 1], xyz[3 * ip + 2] };
    const int neg = xw[2] >= 0 ? 1 : -1;
    const double gg = neg > 0 ? gg0 : gg1, poiss = neg > 0 ? poiss0 : poiss1;
    xw[2] = neg * xw[2]; xw[1] = neg * xw[1];
    double vr[3][4] = {};
    for (int ii = 0; ii < npot; ii++) {
        if (ps[2 * (size_t) npot + ii] > 0.0) {
            const double xp[2] = { xc1 + (ii % mx) * dx, neg * (yc1 + (ii / mx) * dy) };
            double v[3][3][4], vnu[3][3][4];
            stres1_dev(dx, dy, gg, v, vnu, xw, xp);
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 4; k++)
                    for (int i = 0; i < 3; i++)
                        vr[j][k] += ((i == 0) ? neg : 1) * ps[(size_t) i * npot + ii] * (v[i][j][k] + poiss * vnu[i][j][k]);
        }
    }
    sstres_derived_dev(gg, poiss, neg, vr, table + (size_t) ip * 18);
}

}  // namespace cb200
