// large_solver.cuh -- the influence product and the NORM / NormCG solver for grids that do not fit one CTA's shared
// memory (143x163 ... 575x647 and beyond): ONE contact problem uses the WHOLE GPU.
//
// Reference: fft_VecAijPj (/root/reference/src/m_aijpj.f90:712-1015), normcg (/root/reference/src/m_solvpn.f90:24-461),
// snorm (/root/reference/src/m_snorm.f90:31-378) -- same algorithm and padding (opt_fft_size) as the single-CTA path.
//
// B200-first formulation: a product is three grid-wide phases over a spectrum workspace T[kx][iy] that never leaves
// the 126 MB L2 (575x647: T = 6 MB, C^ = 12 MB per block, the 9 NormCG vectors 27 MB):
//   rows   : each CTA takes RB grid rows: packed real FFT of length Fx in shared memory -> T[kx][iy0..iy0+RB)
//   columns: each CTA takes CB columns kx: zero-padded FFT of length 2Fy in shared memory, multiply by C^ in the
//            registers of the last forward stage, inverse FFT, rows Fy..Fy+my-1 written back to T in place
//   rows^-1: Hermitian merge + packed inverse FFT, masked store of u
// HBM sees p, u and (first touch) C^ only.  The whole NORM loop -- NormCG iterations, active-set changes, convergence
// test -- runs in ONE persistent cooperative kernel (one CTA per SM, grid.sync() between phases, fixed-order grid
// reductions), so there is no host round trip per iteration.
#pragma once
#include <cooperative_groups.h>
#include "tang_solver.cuh"

namespace cb200 {
namespace cg = cooperative_groups;

struct LargePlan {
    ConvPlan P;            // sizes, radices, global tables (twx, twy, posx); P.SY / P.C / chunk fields are unused here
    int RB, CB;            // rows per row task, columns per column task
    int ntr, ntc;          // row tasks for the my grid rows, column tasks for the Fx+1 columns
    int ldT;               // leading dimension of T[kx][iy]  (>= my)
    int smem_bytes;
};

template <class B> struct ViewPadLin {       // W[e][c] with implicit zeros beyond n_in rows
    B buf; uint32_t off, estride, n_in;
    CB_HD cd ld(uint32_t e, uint32_t c) const { return e < n_in ? buf.ld(off + e * estride + c) : make_double2(0.0, 0.0); }
};

// ---- row task, forward: rows iy0..iy0+nb-1 of the source -> T[k][iy0+b], k = 0..Lx ----
__device__ void lg_rows_fwd_task(const ConvPlan &P, uint32_t a0, int iy0, int nb, RowSrc src, cd *T, int ldT)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    typedef MemBuf<cd> CB_BUF;
    const CB_BUF BUF = { reinterpret_cast<cd *>(__cvta_shared_to_generic(a0)) };
    const uint32_t oS = 0;
    const MemBuf<const cd> twx = { P.twx };
    const MemBuf<const unsigned short> posx = { P.posx };
    const int SY = nb | 1;
    if (src.kind == 0) { src.base += (size_t) iy0 * src.mx; src.my = nb; } else src.row0 = iy0;
    CB_CONV_FORWARD_ROWS(nb, src);
    const int items = (P.Lx + 1) * nb;
    const uint32_t mg = div_magic(nb);
    for (int w = tid; w < items; w += nthr) {
        const uint32_t k = fdiv(w, mg), b = w - k * nb;
        const uint32_t pos = (int) k < P.Lx ? posx.ld(k) : (uint32_t) P.Lx;
        T[(size_t) k * ldT + iy0 + b] = BUF.ld(pos * SY + b);
    }
    __syncthreads();
}

// ---- row task, inverse: T[k][iy0+b] -> u(ix, iy0+b) on the selected elements ----
__device__ void lg_rows_inv_task(const ConvPlan &P, uint32_t a0, int iy0, int nb, const cd *T, int ldT, double *u,
                                 const int *el, int mask_mode, int add)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    typedef MemBuf<cd> CB_BUF;
    const CB_BUF BUF = { reinterpret_cast<cd *>(__cvta_shared_to_generic(a0)) };
    const uint32_t oS = 0;
    const MemBuf<const cd> twx = { P.twx };
    const MemBuf<const unsigned short> posx = { P.posx };
    const int SY = nb | 1;
    const int items = (P.Lx + 1) * nb;
    const uint32_t mg = div_magic(nb);
    for (int w = tid; w < items; w += nthr) {
        const uint32_t k = fdiv(w, mg), b = w - k * nb;
        const uint32_t pos = (int) k < P.Lx ? posx.ld(k) : (uint32_t) P.Lx;
        BUF.st(pos * SY + b, T[(size_t) k * ldT + iy0 + b]);
    }
    __syncthreads();
    CB_CONV_INVERSE_ROWS(nb);
    const int nst = nb * P.mx;
    const uint32_t mgx = div_magic(P.mx);
    for (int w = tid; w < nst; w += nthr) {
        const uint32_t b = fdiv(w, mgx), ix = w - b * P.mx;
        const size_t ii = (size_t) (iy0 + b) * P.mx + ix;
        if (mask_mode == 1 && el[ii] < 1) continue;
        const uint32_t xi = P.Fx + ix;
        const cd z = BUF.ld((xi >> 1) * SY + b);
        const double v = (xi & 1) ? z.y : z.x;
        u[ii] = add ? u[ii] + v : v;
    }
    __syncthreads();
}

// ---- column task: columns kx0..kx0+nc-1 of T.  chat != null: product (in place on T, rows Fy..Fy+n_out-1 kept);
//      chat_out != null: coefficient builder (forward only, scaled spectrum dumped as [Ly][C]) ----
__device__ void lg_cols_task(const ConvPlan &P, uint32_t a0, int C, int kx0, int nc, int n_in, int n_out, cd *T, int ldT,
                             const cd *chat, cd *chat_out, double scale)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    typedef MemBuf<cd> CB_BUF;
    const CB_BUF BUF = { reinterpret_cast<cd *>(__cvta_shared_to_generic(a0)) };
    const MemBuf<const cd> twy = { P.twy };
    const int Ly = P.Ly;
    {   // coalesced load along iy, transposed into W[e][c]
        const int items = nc * n_in;
        const uint32_t mg = div_magic(n_in);
        for (int w = tid; w < items; w += nthr) {
            const uint32_t c = fdiv(w, mg), e = w - c * n_in;
            BUF.st(e * C + c, T[(size_t) (kx0 + c) * ldT + e]);
        }
    }
    __syncthreads();
    int nsv[CB_MAXSTAGE];
    for (int s = 0, ns = Ly; s < P.nsy; ns /= P.ry[s], s++) nsv[s] = ns;
    const ViewLin<CB_BUF> vW = { BUF, 0u, (uint32_t) C };
    const ViewPadLin<CB_BUF> vin = { BUF, 0u, (uint32_t) C, (uint32_t) n_in };
    if (chat_out) {
        CB_PHASE(fft_stage_r<false>(P.ry[0], vin, vW, nc, Ly, nsv[0], twy, 1, tid, nthr));
        for (int s = 1; s < P.nsy; s++) CB_PHASE(fft_stage_r<false>(P.ry[s], vW, vW, nc, Ly, nsv[s], twy, 1, tid, nthr));
        const int items = Ly * C;
        const uint32_t mg = div_magic(C);
        for (int w = tid; w < items; w += nthr) {
            const uint32_t e = fdiv(w, mg), c = w - e * C;
            chat_out[w] = (int) c < nc ? cscale(BUF.ld(w), scale) : make_double2(0.0, 0.0);
        }
        __syncthreads();
        return;
    }
    if (P.nsy == 1) {
        CB_PHASE(fft_stage_mid_r(P.ry[0], vin, vW, nc, Ly, chat, C, tid, nthr));
    } else {
        CB_PHASE(fft_stage_r<false>(P.ry[0], vin, vW, nc, Ly, nsv[0], twy, 1, tid, nthr));
        for (int s = 1; s < P.nsy - 1; s++) CB_PHASE(fft_stage_r<false>(P.ry[s], vW, vW, nc, Ly, nsv[s], twy, 1, tid, nthr));
        CB_PHASE(fft_stage_mid_r(P.ry[P.nsy - 1], vW, vW, nc, Ly, chat, C, tid, nthr));
        for (int s = P.nsy - 2; s >= 0; s--) CB_PHASE(fft_stage_r<true>(P.ry[s], vW, vW, nc, Ly, nsv[s], twy, 1, tid, nthr));
    }
    {
        const int items = nc * n_out;
        const uint32_t mg = div_magic(n_out);
        for (int w = tid; w < items; w += nthr) {
            const uint32_t c = fdiv(w, mg), r = w - c * n_out;
            T[(size_t) (kx0 + c) * ldT + r] = BUF.ld((P.Fy + r) * C + c);
        }
    }
    __syncthreads();
}

// ---- plain kernels (stand-alone products, coefficient transforms): one launch per phase ----
__global__ void __launch_bounds__(CB_THREADS, 1)
k_lg_rows_fwd(LargePlan L, RowSrc src, int nrows, int RB, cd *T, int ldT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t a0 = (uint32_t) __cvta_generic_to_shared(smem_raw);
    const int ntask = (nrows + RB - 1) / RB;
    for (int t = blockIdx.x; t < ntask; t += gridDim.x)
        lg_rows_fwd_task(L.P, a0, t * RB, min(RB, nrows - t * RB), src, T, ldT);
}

__global__ void __launch_bounds__(CB_THREADS, 1)
k_lg_cols(LargePlan L, int n_in, int n_out, cd *T, int ldT, const cd *chat, cd *chat_out, double scale)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t a0 = (uint32_t) __cvta_generic_to_shared(smem_raw);
    const int ncol = L.P.Fx + 1;
    for (int t = blockIdx.x; t < L.ntc; t += gridDim.x)
        lg_cols_task(L.P, a0, L.CB, t * L.CB, min(L.CB, ncol - t * L.CB), n_in, n_out, T, ldT,
                     chat ? chat + (size_t) t * L.P.Ly * L.CB : nullptr,
                     chat_out ? chat_out + (size_t) t * L.P.Ly * L.CB : nullptr, scale);
}

__global__ void __launch_bounds__(CB_THREADS, 1)
k_lg_rows_inv(LargePlan L, const cd *T, double *u, const int *el, int mask_mode, int add)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t a0 = (uint32_t) __cvta_generic_to_shared(smem_raw);
    for (int t = blockIdx.x; t < L.ntr; t += gridDim.x)
        lg_rows_inv_task(L.P, a0, t * L.RB, min(L.RB, L.P.my - t * L.RB), T, L.ldT, u, el, mask_mode, add);
}

// ---- grid-wide product inside a cooperative kernel; ends with a grid barrier (u visible to every CTA) ----
struct LargeCtx {
    LargePlan L;
    cd *T;
    double *gpart;         // [2][gridDim][8] partial sums of the grid reductions
    unsigned long long *prof;   // optional cycle counters
    int gs_off;                 // byte offset of the Gauss-Seidel row arrays in the dynamic shared memory (stdygs_dev<1, true>)
};

__device__ void lg_conv(const LargeCtx &X, uint32_t a0, const double *p, const cd *chat, double *u, const int *el,
                        int mask_mode, int add, cg::grid_group &grid)
{
    const LargePlan &L = X.L;
    const ConvPlan &P = L.P;
    RowSrc src;
    src.base = p; src.kind = 0; src.mx = P.mx; src.my = P.my; src.cmx = 0; src.cmy = 0; src.Fx = P.Fx; src.Fy = P.Fy; src.row0 = 0; src.stride = 0;
    for (int t = blockIdx.x; t < L.ntr; t += gridDim.x)
        lg_rows_fwd_task(P, a0, t * L.RB, min(L.RB, P.my - t * L.RB), src, X.T, L.ldT);
    grid.sync();
    const int ncol = P.Fx + 1;
    for (int t = blockIdx.x; t < L.ntc; t += gridDim.x)
        lg_cols_task(P, a0, L.CB, t * L.CB, min(L.CB, ncol - t * L.CB), P.my, P.my, X.T, L.ldT,
                     chat + (size_t) t * P.Ly * L.CB, nullptr, 1.0);
    grid.sync();
    for (int t = blockIdx.x; t < L.ntr; t += gridDim.x)
        lg_rows_inv_task(P, a0, t * L.RB, min(L.RB, P.my - t * L.RB), X.T, L.ldT, u, el, mask_mode, add);
    grid.sync();
}

// ---- fixed-order grid reductions (bit-reproducible for a given grid size); every thread gets the result ----
template <int N>
__device__ void grid_sum(double (&v)[N], double *red, double *gpart, int &phase, cg::grid_group &grid)
{
    block_sum<N>(v, red);
    double *slot = gpart + (size_t) (phase & 1) * gridDim.x * 8;
    phase++;
    if (threadIdx.x == 0)
#pragma unroll
        for (int i = 0; i < N; i++) slot[blockIdx.x * 8 + i] = v[i];
    grid.sync();
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        for (int b = lane; b < (int) gridDim.x; b += 32) s += __ldcg(slot + b * 8 + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        v[i] = s;
    }
}

__device__ double grid_min(double v, double *red, double *gpart, int &phase, cg::grid_group &grid)
{
    v = block_min(v, red);
    double *slot = gpart + (size_t) (phase & 1) * gridDim.x * 8;
    phase++;
    if (threadIdx.x == 0) slot[blockIdx.x * 8] = v;
    grid.sync();
    const int lane = threadIdx.x & 31;
    double s = 1e300;
    for (int b = lane; b < (int) gridDim.x; b += 32) s = fmin(s, __ldcg(slot + b * 8));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmin(s, __shfl_xor_sync(0xffffffffu, s, o));
    return s;
}

#define CB_GLOOP(i, n) for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < (size_t) (n); i += (size_t) gridDim.x * blockDim.x)

__device__ void lg_proj_avg(const int *el, double *a, int n, double *red, double *gpart, int &phase, cg::grid_group &grid)
{   // gf3_proj_avg(AllInt): m_gridfunc.f90:1325-1370
    double s[2] = { 0.0, 0.0 };
    CB_GLOOP(i, n) if (el[i] >= 1) { s[0] += a[i]; s[1] += 1.0; }
    grid_sum<2>(s, red, gpart, phase, grid);
    const double avg = s[0] / fmax(1.0, s[1]);
    CB_GLOOP(i, n) if (el[i] >= 1) a[i] -= avg;
}

// normcg (m_solvpn.f90:24-461) on the whole GPU.  Same statement order as normcg_dev; element-wise passes keep the
// same thread -> element mapping, so a grid barrier is needed only in front of products (and inside reductions).
__device__ int lg_normcg(const LargeCtx &X, uint32_t a0, double *red, const NormCase &c, const double *hstot, double &pen,
                         int *el, double *ps, double *wk, int &itcg_out, double &err_out, int &nprod, int &phase,
                         cg::grid_group &grid)
{
    const int n = X.L.P.npot;
    double *rhs = wk, *res = wk + n, *r_prv = wk + 2 * (size_t) n, *dd = wk + 3 * (size_t) n, *z = wk + 4 * (size_t) n,
           *v = wk + 5 * (size_t) n, *q = wk + 6 * (size_t) n;
    double *gp = X.gpart;
    const int ic_norm = c.ic_norm, maxcg = c.maxgs;
    const double eps = c.eps, dxdy = c.dxdy, fntrue = c.fntrue;
    const int numinn = n <= 150 ? 3 : (n <= 400 ? 2 : 1);

    if (ic_norm == 1) pen = 0.0;
    double hmin = 1e20, hmaxn = 1e20;
    double cnt[2] = { 0.0, 0.0 };
    CB_GLOOP(i, n) {
        const double h = hstot[i];
        rhs[i] = pen - h;
        res[i] = 0.0; r_prv[i] = 0.0; dd[i] = 0.0; z[i] = 0.0; v[i] = 0.0; q[i] = 0.0;
        hmin = fmin(hmin, h); hmaxn = fmin(hmaxn, -h);
        if (el[i] >= 1) cnt[0] += 1.0;
        cnt[1] += ps[i];
    }
    grid_sum<2>(cnt, red, gp, phase, grid);
    int ncon = (int) cnt[0];
    const double hsmin0 = grid_min(hmin, red, gp, phase, grid);
    double davg = 0.0;

    if (ic_norm == 0) {
        if (hsmin0 - pen >= 0.0) {                                   // :121-136 no contact at all
            CB_GLOOP(i, n) { ps[i] = 0.0; el[i] = 0; }
            grid.sync();
            itcg_out = 0; err_out = 0.0;
            return 0;
        }
    } else {
        if (ncon <= 0) {                                             // :144-155
            const double hsmax = -grid_min(hmaxn, red, gp, phase, grid);
            const double htrsh = hsmin0 + 0.1 * fmax(hsmax - hsmin0, 1e-10);
            double k[1] = { 0.0 };
            CB_GLOOP(i, n) if (hstot[i] < htrsh) { el[i] = 1; k[0] += 1.0; }
            grid_sum<1>(k, red, gp, phase, grid);
            ncon += (int) k[0];
        }
        const double fk = dxdy * cnt[1];                             // :159-168
        if (fabs(fk) < (double) 1e-3f * fntrue) {
            const double pn = fntrue / (dxdy * (double) ncon);
            CB_GLOOP(i, n) if (el[i] >= 1) ps[i] = pn;
        } else {
            const double f = fntrue / fk;
            CB_GLOOP(i, n) if (el[i] >= 1) ps[i] = f * ps[i];
        }
    }
    grid.sync();
    lg_conv(X, a0, ps, c.chatA, res, el, 1, 0, grid); nprod++;       // :173-175 res = rhs - A ps on C
    CB_GLOOP(i, n) if (el[i] >= 1) res[i] = rhs[i] - res[i];
    if (ic_norm == 1) lg_proj_avg(el, res, n, red, gp, phase, grid);

    double rz1, rz2 = 0.0, rms_xk = 1.0, rms_upd = 2.0 * eps * rms_xk, rms_upd1 = 0.0;
    int itcg = 0, itinn = 0;
    bool lchanged = false;

    while ((lchanged || rms_upd > eps * rms_xk) && itcg < maxcg) {   // :194
        itcg++; itinn++;
        grid.sync();
        lg_conv(X, a0, res, c.chatM, z, el, 1, 0, grid); nprod++;    // z = M res on C
        if (ic_norm == 1) lg_proj_avg(el, z, n, red, gp, phase, grid);

        double d2[2] = { 0.0, 0.0 };
        CB_GLOOP(i, n) if (el[i] >= 1) { d2[0] += z[i] * res[i]; d2[1] += z[i] * r_prv[i]; }
        grid_sum<2>(d2, red, gp, phase, grid);
        rz1 = rz2; rz2 = d2[0];
        if (itcg <= 1 || rz1 < CB_TINY) {                            // :228-241
            CB_GLOOP(i, n) if (el[i] >= 1) v[i] = z[i];
        } else {
            const double beta = fmax(0.0, (rz2 - d2[1]) / fmax(CB_TINY, rz1));
            CB_GLOOP(i, n) if (el[i] >= 1) v[i] = beta * v[i] + z[i];
        }
        if (ic_norm == 1) lg_proj_avg(el, v, n, red, gp, phase, grid);

        grid.sync();
        lg_conv(X, a0, v, c.chatA, q, el, 1, 0, grid); nprod++;      // q = A v on C
        if (ic_norm == 1) lg_proj_avg(el, q, n, red, gp, phase, grid);

        double d4[4] = { 0.0, 0.0, 0.0, 0.0 };
        CB_GLOOP(i, n) if (el[i] >= 1) { const double vi = v[i]; d4[0] += res[i] * vi; d4[1] += q[i] * vi; d4[2] += vi * vi; d4[3] += 1.0; }
        grid_sum<4>(d4, red, gp, phase, grid);
        const double rv = d4[0], vav = d4[1];
        double alpha;
        if (fabs(vav) > 1e-32 && ncon == 1) alpha = rv / vav;
        else alpha = rv / fmax(CB_TINY, vav);
        rms_upd = fabs(alpha) * sqrt(d4[2] / fmax(1.0, d4[3]));
        if (itcg == 1) rms_upd1 = rms_upd;
        const bool need_xk = (itcg <= 3 || itcg % 10 == 0);

        double p2[1] = { 0.0 };
        CB_GLOOP(i, n) {
            r_prv[i] = res[i];                                        // :294 (AllElm copy)
            if (el[i] >= 1) { const double pi = ps[i] + alpha * v[i]; ps[i] = pi; p2[0] += pi * pi; }
        }
        if (need_xk) { grid_sum<1>(p2, red, gp, phase, grid); rms_xk = sqrt(p2[0] / fmax(1.0, d4[3])); }

        if (itinn < numinn && rms_upd >= eps * rms_xk) {             // :298-303
            CB_GLOOP(i, n) if (el[i] >= 1) res[i] -= alpha * q[i];
        } else {
            double k2[1] = { 0.0 };
            CB_GLOOP(i, n) if (el[i] >= 1 && ps[i] < 0.0) { el[i] = 0; ps[i] = 0.0; k2[0] += 1.0; }   // :310-318
            grid_sum<1>(k2, red, gp, phase, grid);
            bool lchg_negpn = k2[0] > 0.0;
            ncon -= (int) k2[0];
            if (ncon <= 0) {                                          // :323-333
                double k[1] = { 0.0 };
                CB_GLOOP(i, n) if (hstot[i] <= hsmin0 + 1e-5) { el[i] = 1; ps[i] = 0.0; k[0] += 1.0; }
                grid_sum<1>(k, red, gp, phase, grid);
                ncon += (int) k[0];
                lchg_negpn = true;
            }
            if (ic_norm == 1 && lchg_negpn) {                         // :337-344
                double s[1] = { 0.0 };
                CB_GLOOP(i, n) s[0] += ps[i];
                grid_sum<1>(s, red, gp, phase, grid);
                double fk = dxdy * s[0];
                if (fabs(fk) < (double) 1e-3f * fntrue) {
                    CB_GLOOP(i, n) if (el[i] >= 1) ps[i] = 1.0;
                    fk = (double) ncon;
                }
                const double f = fntrue / fk;
                CB_GLOOP(i, n) if (el[i] >= 1) ps[i] = f * ps[i];
            }
            grid.sync();
            lg_conv(X, a0, ps, c.chatA, dd, el, 0, 0, grid); nprod++; // :351-352 dd = A ps - rhs, whole grid
            double sd[1] = { 0.0 };
            CB_GLOOP(i, n) { const double d = dd[i] - rhs[i]; dd[i] = d; if (el[i] >= 1) sd[0] += d; }
            if (ic_norm == 1) {                                       // :356-359
                grid_sum<1>(sd, red, gp, phase, grid);
                davg = sd[0] / (double) ncon;
                CB_GLOOP(i, n) if (el[i] >= 1) dd[i] -= davg;
            }
            double ke[1] = { 0.0 };
            CB_GLOOP(i, n) {                                          // :361-384
                const double di = dd[i];
                double r = 0.0;
                if (el[i] >= 1) r = -di;
                else if (di - davg < 0.0) { el[i] = 1; r = -(di - davg); ke[0] += 1.0; }
                else v[i] = 0.0;
                res[i] = r;
            }
            grid_sum<1>(ke, red, gp, phase, grid);
            const bool lchg_intpen = ke[0] > 0.0;
            ncon += (int) ke[0];
            itinn = 0;
            lchanged = lchg_intpen || lchg_negpn;
        }
    }
    if (ic_norm == 1) pen = davg;                                     // :414
    double conv = 1.0;
    if (rms_upd * rms_upd1 > 0.0 && itcg > 1) conv = exp(log(rms_upd / rms_upd1) / (itcg - 1));
    itcg_out = itcg; err_out = rms_upd;
    return (rms_upd > rms_xk && conv > 1.0 && itcg >= maxcg) ? 1 : 0;
}

// snorm (m_snorm.f90:31-378) on the whole GPU
__device__ void lg_snorm(const LargeCtx &X, uint32_t a0, double *red, NormCase &c, int &phase, cg::grid_group &grid)
{
    const ConvPlan &P = X.L.P;
    const int n = P.npot;
    double *wk = c.work, *hstot = wk + 7 * (size_t) n, *unn = wk + 3 * (size_t) n, *tmp = wk + 8 * (size_t) n;
    double *gp = X.gpart;
    int *el = c.el;
    double *ps = c.pn;
    double pen = c.pen;
    int nprod = 0;

    if (c.chatA31 != nullptr && c.ptx != nullptr) {                   // :112-119 hstot = hs + A_zt p_t
        lg_conv(X, a0, c.ptx, c.chatA31, tmp, el, 0, 0, grid);
        lg_conv(X, a0, c.pty, c.chatA32, tmp, el, 0, 1, grid);
        nprod += 2;
        CB_GLOOP(i, n) hstot[i] = c.hs[i] + tmp[i];
    } else {
        CB_GLOOP(i, n) hstot[i] = c.hs[i];
    }

    int itnorm = 0, itcg = 0, it = 0, status = 0;
    bool zready;
    double errpn = 0.0;
    do {                                                              // :142-300
        itnorm++;
        zready = true;
        CB_GLOOP(i, n) if (el[i] < 1) ps[i] = 0.0;
        if (lg_normcg(X, a0, red, c, hstot, pen, el, ps, wk, it, errpn, nprod, phase, grid)) status |= 1;
        itcg += it;

        double k[1] = { 0.0 };
        CB_GLOOP(i, n) if (el[i] >= 1 && ps[i] < -errpn) { el[i] = 0; ps[i] = 0.0; k[0] += 1.0; }    // :197-219 contract
        grid_sum<1>(k, red, gp, phase, grid);
        if (k[0] > 0.0) zready = false;

        if (zready) {                                                 // :227-292 expand
            lg_conv(X, a0, ps, c.chatA, unn, el, 0, 0, grid); nprod++;
            const double tol = fabs(errpn * centre_rowsum_dev(P, c, el, red));
            double kc[1] = { 0.0 };
            CB_GLOOP(i, n) if (el[i] == 0 && hstot[i] - pen < 0.0) {
                const double d = hstot[i] - pen + unn[i];
                if (d < -tol) { el[i] = 1; kc[0] += 1.0; }
            }
            grid_sum<1>(kc, red, gp, phase, grid);
            if (kc[0] > 0.0) zready = false;
        }
        if (it >= c.maxgs) zready = false;
    } while (!zready && itnorm < c.maxin);
    if (!zready) itnorm = -1;

    double s[2] = { 0.0, 0.0 };
    CB_GLOOP(i, n) {                                                  // :315-325
        if (ps[i] < 0.0 && el[i] >= 1) { ps[i] = 0.0; el[i] = 0; }
        if (el[i] < 1) ps[i] = 0.0;
        s[0] += ps[i];
        if (el[i] >= 1) s[1] += 1.0;
    }
    grid_sum<2>(s, red, gp, phase, grid);
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c.pen = pen;
        if (c.ic_norm == 0) c.fntrue = c.dxdy * s[0];                 // :352
        c.itcg = itcg; c.itnorm = itnorm; c.ncon = (int) s[1]; c.status = status; c.err = errpn; c.nprod = nprod;
    }
}

// one contact problem, NORM + u_n, on the whole GPU (cooperative launch, one CTA per SM)
__global__ void __launch_bounds__(CB_THREADS, 1)
k_lg_snorm(LargeCtx X, NormCase *cp, double *un)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    const uint32_t a0 = (uint32_t) __cvta_generic_to_shared(smem_raw);
    double *red = reinterpret_cast<double *>(smem_raw + X.L.smem_bytes - 1024);
    NormCase c = *cp;
    int phase = 0;
    grid.sync();                                                      // everybody has read the case before it is updated
    lg_snorm(X, a0, red, c, phase, grid);
    if (blockIdx.x == 0 && threadIdx.x == 0) *cp = c;
    if (un) lg_conv(X, a0, c.pn, c.chatA, un, c.el, 1, 0, grid);      // soutpt: u_n = A_zz p_n on the contact area
}

// execution context "whole GPU per problem" for the solver text of tang_solver.cuh (see BlockCtx there)
struct GridCtx {
    const LargeCtx &X;
    uint32_t a0;
    double *redp;
    cg::grid_group &grid;
    int &phase;
    unsigned char *sraw;        // dynamic shared memory of this CTA
    static constexpr bool kBlock = false;
    __device__ __forceinline__ int n() const { return X.L.P.npot; }
    __device__ __forceinline__ const ConvPlan &plan() const { return X.L.P; }
    __device__ __forceinline__ double *red() const { return redp; }
    __device__ __forceinline__ size_t first() const { return (size_t) blockIdx.x * blockDim.x + threadIdx.x; }
    __device__ __forceinline__ size_t stride() const { return (size_t) gridDim.x * blockDim.x; }
    __device__ __forceinline__ bool leader() const { return blockIdx.x == 0 && threadIdx.x == 0; }
    // rows interleaved over the CTAs so that a few hundred rows spread over all SMs
    __device__ __forceinline__ size_t row_first() const { return (size_t) threadIdx.x * gridDim.x + blockIdx.x; }
    __device__ __forceinline__ size_t row_stride() const { return (size_t) gridDim.x * blockDim.x; }
    __device__ __forceinline__ size_t warp_first() const { return (size_t) (threadIdx.x >> 5) * gridDim.x + blockIdx.x; }
    __device__ __forceinline__ size_t warp_stride() const { return (size_t) gridDim.x * (blockDim.x >> 5); }
    __device__ __forceinline__ int row_group_width(int nrows) const
    { const int nw = (int) (gridDim.x * (blockDim.x >> 5)); return nrows <= nw ? 32 : (nrows <= 2 * nw ? 16 : (nrows <= 4 * nw ? 8 : 4)); }
    __device__ __forceinline__ void sync() const { grid.sync(); }
    template <int N> __device__ __forceinline__ void sum(double (&v)[N]) const { grid_sum<N>(v, redp, X.gpart, phase, grid); }
    __device__ __forceinline__ void conv(const double *p, const cd *chat, double *u, const int *el, int mask_mode, int add) const
    { lg_conv(X, a0, p, chat, u, el, mask_mode, add, grid); }
    __device__ __forceinline__ void snorm(NormCase &c) const { lg_snorm(X, a0, redp, c, phase, grid); grid.sync(); }
    __device__ __forceinline__ void update_box(const NormCase &, const int *) const {}
    __device__ __forceinline__ void conv_int(int, int, int, const double *p, const cd *chat_full, double *u, const int *el, int add) const
    { lg_conv(X, a0, p, chat_full, u, el, 1, add, grid); }
};

// one contact case (NORM / TANG alternation: T = 0 or shifts T = 1 with TangCG) on the whole GPU
__global__ void __launch_bounds__(CB_THREADS, 1)
k_lg_contac(LargeCtx X, ContactCase *cp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    const uint32_t a0 = (uint32_t) __cvta_generic_to_shared(smem_raw);
    double *red = reinterpret_cast<double *>(smem_raw + X.L.smem_bytes - 1024);
    int phase = 0;
    const GridCtx x = { X, a0, red, grid, phase, smem_raw };
    panprc_dev(x, *cp);
}

}  // namespace cb200
