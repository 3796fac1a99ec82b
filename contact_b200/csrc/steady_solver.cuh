// steady_solver.cuh -- device-resident SteadyGS for steady rolling (T=3): one CTA owns one contact problem.
//
// Mirrors the reference's stdygs (/root/reference/src/m_solvpt.f90:2825-3254), plstrc (:3278-3807, elastic branches)
// and the leading-edge factor of sxbnd (/root/reference/src/m_leadedge.f90:92-332): a Gauss-Seidel sweep over the
// contact elements in rolling order; per element the 2x2 constrained solve, then re-integration of the traction
// differences dp along the row.
//
// B200-first formulation.  The reference evaluates two O(ncon) row sums AijPj per element per sweep.  Here the
// displacement differences U = A_tt dp of all contact elements live in REGISTERS (each thread owns <= KMAX contact
// elements), the tangential coefficient blocks live in shared memory as one quadrant table (c11, c22 even, c12 odd
// in x and y: 3 npot doubles instead of 16 npot), and every change of dp is applied to all elements as a rank-1
// update (same flops as the row sum, no reduction, no global-memory traffic in the sweep).  The sequential part of
// the Gauss-Seidel step (plstrc + re-integration of the current row) runs on one thread out of shared memory.
// U is (re)computed from scratch with four FFT products at the start of every solver call.
#pragma once
#include "norm_solver.cuh"

namespace cb200 {

enum { EL_EXTER = 0, EL_ADHES = 1, EL_SLIP = 2, EL_PLAST = 3 };

// cycle counters of the SteadyGS step (thread 0 of every CTA adds its totals at the end of a solver call):
// [0] element steps, [1] plstrc, [2] re-integration, [3] waiting + rank-1 updates + hand-over, [4] solver calls
__device__ unsigned long long g_steady_prof[8];

// plstrc (m_solvpt.f90:3278-3807) without plasticity (tau_c = 1e20, k_tau = 0): el, (px,py), (sx,sy) in/out.
// c00..c11: 2x2 influence matrix of the element, bound = mus * pn.
__device__ void plstrc_dev(int &el, double c00, double c01, double c10, double c11, double eps, double omegah,
                           double omegas, double &px, double &py, double bound, double &sx, double &sy)
{
    const double pi = 3.14159265358979323846, prc = 0.0001;
    const double pox = px, poy = py;
    const double s0x = sx - c00 * pox - c01 * poy, s0y = sy - c10 * pox - c11 * poy;   // shift without this element
    double six = 0.0, siy = 0.0;
    bool violated = true;
    for (int itry = 1; itry <= 4 && violated; itry++) {
        if (el == EL_ADHES) {                                                  // linear equations S = 0
            const double dinv = 1.0 / (c00 * c11 - c01 * c10);
            const double dx = (-c11 * sx + c01 * sy) * dinv, dy = (c10 * sx - c00 * sy) * dinv;
            px = pox + omegah * dx; py = poy + omegah * dy;
            const double q2 = px * px + py * py, b2 = bound * bound;           // compare squares: no square root on the
            if (q2 <= b2) violated = false;                                    // common path
            else {
                const double fx = pox + dx, fy = poy + dy;
                if (fx * fx + fy * fy <= b2) { const double t = bound / sqrt(q2); px *= t; py *= t; violated = false; }
                else violated = true;
            }
            six = s0x + c00 * px + c01 * py; siy = s0y + c10 * px + c11 * py;
        } else if (el == EL_SLIP) {                                            // |P| = bound, S anti-parallel to P
            violated = false;
            // (one reciprocal instead of two divisions here and in the Newton step: a device division costs ~20
            //  dependent instructions and this code runs on a single thread)
            const double t0 = -bound / sqrt(s0x * s0x + s0y * s0y);
            px = t0 * s0x; py = t0 * s0y;
            six = s0x + c00 * px + c01 * py; siy = s0y + c10 * px + c11 * py;
            double f0 = py * six - px * siy, f1 = px * px + py * py - bound * bound;
            int itnr = 0;
            while (itnr == 0 || (itnr < 10 && fabs(f0) + fabs(f1) >= prc * eps * bound)) {
                itnr++;
                const double g00 = -2.0 * c10 * px - s0y + (c00 - c11) * py, g01 = 2.0 * c01 * py + s0x + (c00 - c11) * px;
                const double g10 = 2.0 * px, g11 = 2.0 * py;
                const double det = g00 * g11 - g01 * g10;
                if (det == 0.0) itnr = 10;
                else {
                    const double rdet = 1.0 / det;
                    px -= (g11 * f0 - g01 * f1) * rdet;
                    py -= (-g10 * f0 + g00 * f1) * rdet;
                }
                six = s0x + c00 * px + c01 * py; siy = s0y + c10 * px + c11 * py;
                f0 = py * six - px * siy; f1 = px * px + py * py - bound * bound;
            }
            if (omegas == 1.0) {
                // relaxation of the direction with omega = 1 (:3612-3619): bound * (cos, sin)(atan2(py, px)) is the
                // traction scaled back onto the bound -- no trigonometry needed
                const double t1 = bound / sqrt(px * px + py * py);
                px *= t1; py *= t1;
            } else {
                const double a0 = atan2(poy, pox);
                double da = atan2(py, px) - a0;
                if (da < -pi) da += 2.0 * pi;
                if (da > pi) da -= 2.0 * pi;
                const double a1 = a0 + omegas * da;
                px = bound * cos(a1); py = bound * sin(a1);
            }
            if (fabs(px) > fabs(py)) { if (px * six > 0.0) violated = true; }
            else { if (py * siy > 0.0) violated = true; }
        } else {                                                               // plasticity with an infinite yield limit
            six = 0.0; siy = 0.0;
            violated = true;
        }
        if (itry <= 3 && violated) el = (el == EL_ADHES) ? EL_SLIP : (el == EL_SLIP ? EL_PLAST : EL_ADHES);
    }
    sx = six; sy = siy;
}

// leading-edge factor facdt (m_leadedge.f90:92-332, chi = 0, no leading-edge correction): 0 in the exterior, 1 near
// the end of the grid, min(1, (xbnd - x)/dq) otherwise with xbnd two elements beyond the next C->E transition
__device__ void sxbnd_facdt_dev(int mx, int my, const int *el, double dx, double dq, double *facdt)
{
    for (int iy = threadIdx.x; iy < my; iy += blockDim.x) {
        const int *e = el + (size_t) iy * mx;
        double *f = facdt + (size_t) iy * mx;
        int ixb = -1;                                   // next transition at or to the right of ix (0-based), found lazily
        for (int ix = 0; ix < mx; ix++) {
            if (e[ix] < 1) { f[ix] = 0.0; continue; }
            if (ixb < ix) { ixb = ix; while (ixb < mx - 1 && e[ixb + 1] >= 1) ixb++; }
            if (ix + 3 > mx) f[ix] = 1.0;
            else f[ix] = fmin(1.0, ((double) (ixb - ix + 2) * dx) / dq);
        }
    }
    __syncthreads();
}

struct SteadyArgs {
    const double *ws;       // [2][n] right-hand side
    double *dp;             // [2][n] traction differences (global scratch)
    double *ug;             // [2][n] scratch for the FFT evaluation of U
    int *iel;               // [n] compact list of contact elements
    const cd *(*chatA)[3];  // transformed cs blocks
    const double *cf11, *cf12, *cf22;
    int cmx, cmy;
    double ga_inv, mu, eps, omegah, omegas;
    int maxgs;
};

// shared-memory carve-up for the sweep (bytes); tab = 0: coefficient table does not fit, read cf from global memory
struct SteadySmem { double *q, *psx, *psy, *dpx, *dpy, *bnd, *wsx, *wsy, *ssx, *ssy, *chx, *chy, *scal; int *el, *chj, *rowk, *ictl; };

__device__ __forceinline__ bool steady_carve(const ConvPlan &P, unsigned char *base, SteadySmem &s)
{
    const size_t n = P.npot, mx = P.mx, my = P.my;
    const size_t fixed = (11 * mx + 8) * 8 + (2 * mx + my + 8) * 4 + 64;
    const bool tab = 3 * n * 8 + fixed <= (size_t) P.off_twx;
    double *d = reinterpret_cast<double *>(base);
    s.q = tab ? d : nullptr;
    if (tab) d += 3 * n;
    s.psx = d; s.psy = d + mx; s.dpx = d + 2 * mx; s.dpy = d + 3 * mx; s.bnd = d + 4 * mx; s.wsx = d + 5 * mx; s.wsy = d + 6 * mx;
    s.ssx = d + 7 * mx; s.ssy = d + 8 * mx; s.chx = d + 9 * mx; s.chy = d + 10 * mx; s.scal = d + 11 * mx;
    int *i = reinterpret_cast<int *>(d + 11 * mx + 8);
    s.el = i; s.chj = i + mx; s.rowk = i + 2 * mx; s.ictl = i + 2 * mx + my + 2;
    return fixed <= (size_t) P.off_twx;
}

// stdygs: returns info (0 ok, 1 maxgs reached, 2 stagnation, 3 divergence).  All threads of the CTA must call.
template <int KMAX>
__device__ __noinline__ int stdygs_dev(const ConvPlan &P, const Smem &sm, const SteadyArgs &a, int *el, double *ps, double *ss,
                          int ncon, int &itgs_out, double &err_out, int &nprod)
{
    const int n = P.npot, mx = P.mx, my = P.my, tid = threadIdx.x, nt = blockDim.x;
    double *psx = ps, *psy = ps + n, *psn = ps + 2 * (size_t) n;
    double *red = sm.red;

    // traction differences along the rolling direction (x ascending = towards the leading edge), :2900-2915
    for (int i = tid; i < n; i += nt) {
        const int ix = i % mx;
        a.dp[i] = (ix != mx - 1) ? psx[i] - psx[i + 1] : psx[i];
        a.dp[n + i] = (ix != mx - 1) ? psy[i] - psy[i + 1] : psy[i];
    }
    __syncthreads();
    const double facnel = (double) sqrtf(__fdiv_rn((float) n, (float) ncon));

    // U = A_tt dp on the contact area by four FFT products (fresh at every solver call)
    for (int ik = 0; ik < 2; ik++) {
        bool ladd = false;
        for (int jk = 0; jk < 2; jk++) {
            if (a.chatA[ik][jk] == nullptr) continue;
            conv_dev(P, sm, a.dp + (size_t) jk * n, a.chatA[ik][jk], a.ug + (size_t) ik * n, el, 1, ladd ? 1 : 0);
            ladd = true; nprod++;
        }
    }

    // shared memory is ours now (S and W regions of the FFT layout)
    SteadySmem s;
    steady_carve(P, reinterpret_cast<unsigned char *>(sm.S), s);
    if (s.q) {
        for (int i = tid; i < n; i += nt) {
            const int ay = i / mx, ax = i - ay * mx;
            const size_t o = (size_t) (ay + a.cmy) * (2 * a.cmx) + ax + a.cmx;
            s.q[i] = a.cf11[o] * a.ga_inv; s.q[n + i] = a.cf12[o] * a.ga_inv; s.q[2 * n + i] = a.cf22[o] * a.ga_inv;
        }
    }
    // compact list of contact elements in sweep order + row offsets
    for (int iy = tid; iy < my; iy += nt) {
        int cnt = 0;
        for (int ix = 0; ix < mx; ix++) cnt += (el[iy * mx + ix] >= 1);
        s.rowk[iy + 1] = cnt;
    }
    __syncthreads();
    if (tid == 0) { s.rowk[0] = 0; for (int iy = 0; iy < my; iy++) s.rowk[iy + 1] += s.rowk[iy]; }
    __syncthreads();
    for (int iy = tid; iy < my; iy += nt) {
        int k = s.rowk[iy];
        for (int ix = 0; ix < mx; ix++) if (el[iy * mx + ix] >= 1) a.iel[k++] = iy * mx + ix;
    }
    __syncthreads();

    // registers: my contact elements k = tid + m nt
    double Ux[KMAX], Uy[KMAX];
    int ixy[KMAX];
#pragma unroll
    for (int m = 0; m < KMAX; m++) {
        const int k = tid + m * nt;
        Ux[m] = 0.0; Uy[m] = 0.0; ixy[m] = 0;
        if (k < ncon) {
            const int ii = a.iel[k], iy = ii / mx;
            ixy[m] = (ii - iy * mx) | (iy << 16);
            Ux[m] = a.ug[ii]; Uy[m] = a.ug[n + ii];
        }
    }
    if (tid == 0) { s.scal[0] = Ux[0]; s.scal[1] = Uy[0]; }
    __syncthreads();

    const double q00 = s.q ? s.q[0] : a.cf11[(size_t) a.cmy * 2 * a.cmx + a.cmx] * a.ga_inv;
    const double q11 = s.q ? s.q[2 * n] : a.cf22[(size_t) a.cmy * 2 * a.cmx + a.cmx] * a.ga_inv;
    const double q01 = s.q ? s.q[n] : a.cf12[(size_t) a.cmy * 2 * a.cmx + a.cmx] * a.ga_inv;

    int itgs = 0;
    double dif = 2.0, difid = 1.0, dif1 = 0.0;
    unsigned long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0, tlast = clock64();
    while (dif >= difid && itgs < a.maxgs) {
        itgs++;
        double dsum = 0.0;                                     // thread 0 only
        int own_t = 0, own_m = 0;                              // owner (thread, slot) of the NEXT element k+1
        for (int iy = 0; iy < my; iy++) {
            const int k0 = s.rowk[iy], k1 = s.rowk[iy + 1];
            if (k1 == k0) continue;
            for (int jx = tid; jx < mx; jx += nt) {            // stage the row in shared memory
                const int ii = iy * mx + jx;
                s.psx[jx] = psx[ii]; s.psy[jx] = psy[ii]; s.dpx[jx] = a.dp[ii]; s.dpy[jx] = a.dp[n + ii];
                s.bnd[jx] = a.mu * psn[ii]; s.wsx[jx] = a.ws[ii]; s.wsy[jx] = a.ws[n + ii];
                s.ssx[jx] = ss[ii]; s.ssy[jx] = ss[n + ii]; s.el[jx] = el[ii];
            }
            __syncthreads();
            int ixc = -1;                                      // thread 0: position of the current element
            for (int k = k0; k < k1; k++) {
                if (tid == 0) {
                    const unsigned long long ta = clock64();
                    tp3 += ta - tlast; tp0++;
                    do ixc++; while (s.el[ixc] < 1);
                    const int ix = ixc;
                    int jx = ix - 1;
                    while (jx > 0 && s.el[jx] == EL_ADHES) jx--;
                    const int off = ix - jx;
                    double t00, t01, t11;
                    if (s.q) { t00 = s.q[off]; t01 = -s.q[n + off]; t11 = s.q[2 * n + off]; }
                    else {
                        const size_t o = (size_t) a.cmy * 2 * a.cmx + a.cmx - off;
                        t00 = a.cf11[o] * a.ga_inv; t01 = a.cf12[o] * a.ga_inv; t11 = a.cf22[o] * a.ga_inv;
                    }
                    const double c00 = q00 - t00, c01 = q01 - t01, c11 = q11 - t11;
                    double sx = s.wsx[ix] + s.scal[0], sy = s.wsy[ix] + s.scal[1];
                    const double pox = s.psx[ix], poy = s.psy[ix];
                    double px = pox, py = poy;
                    int e = s.el[ix];
                    plstrc_dev(e, c00, c01, c01, c11, a.eps, a.omegah, a.omegas, px, py, s.bnd[ix], sx, sy);
                    const unsigned long long tb = clock64();
                    tp1 += tb - ta;
                    const double ex = px - pox, ey = py - poy;
                    dsum += ex * ex + ey * ey;
                    int nch = 0;
                    if (ex != 0.0 || ey != 0.0) { s.chj[nch] = ix; s.chx[nch] = ex; s.chy[nch] = ey; nch++; }
                    s.dpx[ix] += ex; s.dpy[ix] += ey;
                    s.psx[ix] = px; s.psy[ix] = py; s.ssx[ix] = sx; s.ssy[ix] = sy; s.el[ix] = e;
                    // re-integrate dp -> ps to the left (:3089-3126); stops where nothing can change any more
                    double rx = px, ry = py;                   // tractions of element jj+1
                    for (int jj = ix - 1; jj >= 0; jj--) {
                        const int ej = s.el[jj];
                        if (ej == EL_ADHES) {
                            double nx = rx + s.dpx[jj], ny = ry + s.dpy[jj];
                            const double pa2 = nx * nx + ny * ny, pb = fmin(s.bnd[jj], 1e20);
                            if (pa2 > pb * pb) {
                                const double t = pb / sqrt(pa2);
                                nx *= t; ny *= t;
                                const double ndx = nx - rx, ndy = ny - ry;
                                s.chj[nch] = jj; s.chx[nch] = ndx - s.dpx[jj]; s.chy[nch] = ndy - s.dpy[jj]; nch++;
                                s.dpx[jj] = ndx; s.dpy[jj] = ndy;
                            }
                            const bool same = (nx == s.psx[jj] && ny == s.psy[jj]);
                            s.psx[jj] = nx; s.psy[jj] = ny;
                            rx = nx; ry = ny;
                            if (same) break;
                        } else {
                            double ndx, ndy;
                            if (ej >= EL_SLIP) { ndx = s.psx[jj] - rx; ndy = s.psy[jj] - ry; }
                            else if (s.el[jj + 1] >= EL_ADHES) { ndx = -rx; ndy = -ry; }
                            else break;
                            const double cx = ndx - s.dpx[jj], cy = ndy - s.dpy[jj];
                            if (cx != 0.0 || cy != 0.0) { s.chj[nch] = jj; s.chx[nch] = cx; s.chy[nch] = cy; nch++; }
                            s.dpx[jj] = ndx; s.dpy[jj] = ndy;
                            break;
                        }
                    }
                    s.ictl[0] = nch;
                    tlast = clock64();
                    tp2 += tlast - tb;
                }
                __syncthreads();
                // rank-1 updates of U for every changed dp, all elements
                const int nch = s.ictl[0];
                for (int c = 0; c < nch; c++) {
                    const int jx = s.chj[c];
                    const double ex = s.chx[c], ey = s.chy[c];
#pragma unroll
                    for (int m = 0; m < KMAX; m++) {
                        if (tid + m * nt < ncon) {
                            const int dx = (ixy[m] & 0xffff) - jx, dy = (ixy[m] >> 16) - iy;
                            double c11, c12, c22;
                            if (s.q) {
                                const int o = abs(dy) * mx + abs(dx);
                                c11 = s.q[o]; c12 = s.q[n + o]; c22 = s.q[2 * n + o];
                                if ((dx < 0) != (dy < 0)) c12 = -c12;
                            } else {
                                const size_t o = (size_t) (dy + a.cmy) * (2 * a.cmx) + dx + a.cmx;
                                c11 = a.cf11[o] * a.ga_inv; c12 = a.cf12[o] * a.ga_inv; c22 = a.cf22[o] * a.ga_inv;
                            }
                            Ux[m] += c11 * ex + c12 * ey;
                            Uy[m] += c12 * ex + c22 * ey;
                        }
                    }
                }
                // hand U of the next element to thread 0
                own_t++; if (own_t == nt) { own_t = 0; own_m++; }
                if (k + 1 == ncon) { own_t = 0; own_m = 0; }
                if (tid == own_t) {
#pragma unroll
                    for (int m = 0; m < KMAX; m++) if (m == own_m) { s.scal[0] = Ux[m]; s.scal[1] = Uy[m]; }
                }
                __syncthreads();
            }
            for (int jx = tid; jx < mx; jx += nt) {            // write the row back
                const int ii = iy * mx + jx;
                psx[ii] = s.psx[jx]; psy[ii] = s.psy[jx]; a.dp[ii] = s.dpx[jx]; a.dp[n + ii] = s.dpy[jx];
                ss[ii] = s.ssx[jx]; ss[n + ii] = s.ssy[jx]; el[ii] = s.el[jx];
            }
            __syncthreads();
        }
        if (tid == 0) s.scal[2] = dsum;
        double p2[1] = { 0.0 };
        for (int i = tid; i < n; i += nt) p2[0] += psx[i] * psx[i] + psy[i] * psy[i];
        block_sum<1>(p2, red);
        dif = sqrt(s.scal[2] / (2.0 * ncon));
        difid = a.eps * fmax(1e-6, facnel * sqrt(p2[0] / (2.0 * n)));
        if (itgs == 1) dif1 = dif;
        __syncthreads();
    }
    double conv = 1.0;
    if (dif * dif1 != 0.0 && itgs > 1) conv = exp(log(dif / dif1) / (itgs - 1));
    int info = 0;
    if (itgs >= a.maxgs) info = 1;
    if (itgs >= a.maxgs && conv > 0.997) info = 2;
    if (itgs >= a.maxgs && conv > 1.0) info = 3;
    itgs_out = itgs; err_out = dif;
    if (tid == 0) {
        atomicAdd(&g_steady_prof[0], tp0); atomicAdd(&g_steady_prof[1], tp1); atomicAdd(&g_steady_prof[2], tp2);
        atomicAdd(&g_steady_prof[3], tp3); atomicAdd(&g_steady_prof[4], 1ull);
    }
    __syncthreads();
    return info;
}

}  // namespace cb200
