// tang_solver.cuh -- device-resident tangential problem for shifts (T=1): TangCG, the Newton-Raphson loop on the
// creepages for prescribed tangential forces, the TANG adhesion/slip active-set loop and the Panagiotopoulos
// NORM/TANG alternation -- one CTA owns one contact problem for the whole case.
//
// Mirrors the reference's tangcg (/root/reference/src/m_solvpt.f90:1841-2442), solvpt (:51-378), stang / stang_rhs
// (/root/reference/src/m_stang.f90:28-746, 749-951; Coulomb friction L=0, elastic material) and panprc
// (/root/reference/src/m_scontc.f90:356-553).  Every influence product is the shared-memory FFT convolution; a 2x2
// tangential product is four single-block products accumulated in place.
#pragma once
#include "norm_solver.cuh"
#include "steady_solver.cuh"

namespace cb200 {

#define CB_MAXNR_LOG 32

// parameters of GDsteady (solv_input, m_sinput.f90:649-680; cntc_setsolverflags G=5, contact_addon.f90:1220-1250)
struct GdParams { int gd_meth, kdown, kdowfb; double fdecay, betath, d_ifc, d_lin, d_cns, d_slp, pow_s; };

struct ContactCase {
    NormCase nrm;                   // normal problem (hs normal, el, pn = ps + 2n, work, ...)
    // tangential inputs
    int tang, force3, maxnr, maxout;
    double cksi, ceta, fxrel, fyrel, fstat;
    const double *hst;              // [2][n] tangential right-hand side -dq*(rigid slip), without unknown creepages
    const double *pv;               // [3][n] tractions of the previous time instance (null = zero)
    double *ps;                     // [3][n] tractions (in/out), nrm.pn == ps + 2n
    double *ss;                     // [2][n] shift / slip distance (out)
    double *twork;                  // 24 n doubles
    const cd *chatA[3][3];          // transformed cs blocks (null where not needed / no n-t coupling)
    const cd *chatV[3][3];          // transformed cv blocks (shifts: same as cs)
    const cd *chatM11, *chatM22;    // transformed preconditioner blocks
    const double *cf11, *cf22;      // spatial blocks cs(1,1), cs(2,2)
    double c11, c22, ga;
    const double *cf12;             // spatial block cs(1,2)
    double dq, dx;                  // rolling step (shifts: 1), grid step in x
    int gausei;                     // G-digit
    double omegah, omegas;          // relaxation factors of the Gauss-Seidel solvers (set by stang)
    const cd *chatSV[2][2];         // transformed csv = cs - cv blocks (steady rolling with ConvexGS)
    const double *cfv11, *cfv12, *cfv22;   // spatial blocks of csv
    const double *cf13, *cf23;      // spatial blocks cs(1,3), cs(2,3) (null without n-t coupling): ubnd of the leading edge
    int solver_eff;                 // 0 TangCG, 1 SteadyGS, 2 ConvexGS, 3 GDsteady (set by stang, m_stang.f90:144-191)
    GdParams gd;                    // G = 5
    double *gwork;                  // 16 n doubles of GDsteady work space (null unless G = 5)
    int gd_fallback, gd_ntrial;     // out: GDsteady stagnated -> SteadyGS used (m_solvpt.f90:474-484); line-search trials
    // outputs
    int ittang, itgs, itout, nr_n;
    int nr_itcg[CB_MAXNR_LOG];
    double nr_cksi[CB_MAXNR_LOG], nr_ceta[CB_MAXNR_LOG], nr_fx[CB_MAXNR_LOG], nr_fy[CB_MAXNR_LOG];
    double fx, fy, sens[2][2];
    int nadh, nslip;
    double pan_dif[16], pan_difid[16];  // dif / difid of panprc's convergence test per outer iteration (m_scontc.f90:510-513)
    int gs_info, gs_it; double gs_err;  // result of a Gauss-Seidel solve by CTA 0 on the whole-GPU path, for the other CTAs
    int tstatus;                    // bit 0: the case needs a solver outside this path (ConvexGS / GDsteady)
};

// Execution context of the solver code below: one CTA per problem (BlockCtx, the batched path) or the whole GPU per
// problem (GridCtx in large_solver.cuh).  The solver text is written once against this interface: element loops run
// from first() with stride(), sum<N>() is a fixed-order reduction whose result every thread receives, sync() orders
// element passes against products, conv() is the influence product.
struct BlockCtx {
    const ConvPlan &P;
    const Smem &sm;
    mutable ContactBox box = { 0, 0, 0, 0, nullptr };   // bounding box of the contact area + its transform level (null: full grid)
    static constexpr bool kBlock = true;
    __device__ __forceinline__ int n() const { return P.npot; }
    __device__ __forceinline__ void update_box(const NormCase &c, const int *el) const { box = contact_box_dev(P, c, el, sm.red); }
    // AllInt product with block (ik, jk) of cs (lset 0) or of the preconditioner (lset 1): on the contact box when a smaller
    // level with that block is prepared; lset < 0: inputs that may be non-zero outside the contact area -> full grid
    __device__ __forceinline__ void conv_int(int lset, int ik, int jk, const double *p, const cd *chat_full, double *u, const int *el, int add) const
    {
        const cd *ch = (lset >= 0 && box.lv) ? box.lv->chat[lset][ik][jk] : nullptr;
        if (ch) conv_box_dev(box.lv->P, sm, p, ch, u, el, 1, add, box.x0, box.y0, box.bw, box.bh, P.mx);
        else conv_dev(P, sm, p, chat_full, u, el, 1, add);
    }
    __device__ __forceinline__ const ConvPlan &plan() const { return P; }
    __device__ __forceinline__ const Smem &smem() const { return sm; }
    __device__ __forceinline__ double *red() const { return sm.red; }
    __device__ __forceinline__ size_t first() const { return threadIdx.x; }
    __device__ __forceinline__ size_t stride() const { return blockDim.x; }
    __device__ __forceinline__ bool leader() const { return threadIdx.x == 0; }
    // loops over grid rows: one thread per row / one warp per row
    __device__ __forceinline__ size_t row_first() const { return threadIdx.x; }
    __device__ __forceinline__ size_t row_stride() const { return blockDim.x; }
    __device__ __forceinline__ size_t warp_first() const { return threadIdx.x >> 5; }
    __device__ __forceinline__ size_t warp_stride() const { return blockDim.x >> 5; }
    // lanes per grid row in the row walks of GDsteady: the largest of 32, 16, 8, 4 with which all rows are in flight at once
    __device__ __forceinline__ int row_group_width(int nrows) const
    { const int nw = blockDim.x >> 5; return nrows <= nw ? 32 : (nrows <= 2 * nw ? 16 : (nrows <= 4 * nw ? 8 : 4)); }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    template <int N> __device__ __forceinline__ void sum(double (&v)[N]) const { block_sum<N>(v, sm.red); }
    __device__ __forceinline__ void conv(const double *p, const cd *chat, double *u, const int *el, int mask_mode, int add) const
    { conv_dev(P, sm, p, chat, u, el, mask_mode, add); }
    __device__ __forceinline__ void snorm(NormCase &c) const { snorm_dev(P, sm, c); __syncthreads(); }
};

// u(ik) = sum_jk A(ik,jk) p(jk) over the given direction ranges, masked (AllInt when el given), blocks with a null
// transform are skipped (no normal-tangential coupling: m_aijpj.f90:358-369); returns the number of products done
template <class X>
__device__ int conv_multi(const X &x, const cd *(&chat)[3][3], const double *p, int jk0, int jk1,
                          double *u, int ik0, int ik1, const int *el, int mask_mode, int lset = -1)
{
    const int n = x.n();
    int np = 0;
    for (int ik = ik0; ik <= ik1; ik++) {
        bool ladd = false;
        for (int jk = jk0; jk <= jk1; jk++) {
            if (chat[ik][jk] == nullptr) continue;
            if (mask_mode == 1) x.conv_int(lset, ik, jk, p + (size_t) jk * n, chat[ik][jk], u + (size_t) ik * n, el, ladd ? 1 : 0);
            else x.conv(p + (size_t) jk * n, chat[ik][jk], u + (size_t) ik * n, el, mask_mode, ladd ? 1 : 0);
            ladd = true; np++;
        }
        if (!ladd) {
            for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (mask_mode == 0 || el[i] >= 1) u[(size_t) ik * n + i] = 0.0;
            x.sync();
        }
    }
    return np;
}

template <class X>
__device__ __forceinline__ void count_el(const X &x, const int *el, int n, int &nadh, int &nslip)
{
    double c[2] = { 0.0, 0.0 };
    for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) { const int e = el[i]; if (e == 1) c[0] += 1.0; else if (e >= 2) c[1] += 1.0; }
    x.template sum<2>(c);
    nadh = (int) c[0]; nslip = (int) c[1];
}

}  // namespace cb200
#include "gdsteady_solver.cuh"
namespace cb200 {

// m_solvpt.f90:1841-2442 (elastic material).  ws: [2][n] right-hand side; mu = fstat (uniform).
template <class X>
__device__ void tangcg_dev(const X &x, ContactCase &c, const double *ws, int maxcg, double eps,
                           int &itcg_out, double &err_out, int &nprod)
{
    const int n = x.n();
    int *el = c.nrm.el;
    double *ps = c.ps, *psn = c.ps + 2 * (size_t) n, *ss = c.ss;
    double *g = c.twork, *nx = g + n, *ny = nx + n, *r = ny + n, *z = r + 2 * n, *v = z + 2 * n, *q = v + 2 * n,
           *pold = q + 2 * n;
    const double small = 1e-6, ga = c.ga, c11 = c.c11, c22 = c.c22, mu = c.fstat;
    const int num_inn = 4;
    int nadh, nslip;
    count_el(x, el, n, nadh, nslip);
    const double facnel = (double) sqrtf(__fdiv_rn((float) n, (float) (nadh + nslip)));      // REAL(4) arithmetic, :1948
    bool use_fftprec = true, lchanged = false;
    int itcg = 0, it_inn = 0;
    double alpha = 0.0, dif = 2.0, difid = 1.0, difinn = 0.0, trsinn = 0.0;

    // g, n, t; ss = A ps + ws on C; r = -ss projected on the tangent in S
    for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
        g[i] = mu * psn[i];
        const double a = atan2(ps[n + i], ps[i]);
        nx[i] = cos(a); ny[i] = sin(a);
        ss[i] = 0.0; ss[n + i] = 0.0;
        v[i] = 0.0; v[n + i] = 0.0; q[i] = 0.0; q[n + i] = 0.0; z[i] = 0.0; z[n + i] = 0.0; pold[i] = 0.0; pold[n + i] = 0.0;
    }
    x.sync();
    nprod += conv_multi(x, c.chatA, ps, 0, 1, ss, 0, 1, el, 1, 0);
    for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
        const int e = el[i];
        double sx = ss[i], sy = ss[n + i];
        if (e >= 1) { sx += ws[i]; sy += ws[n + i]; ss[i] = sx; ss[n + i] = sy; }
        double rx = -sx, ry = -sy;
        if (e == 2) { const double tx = -ny[i], ty = nx[i], perp = tx * rx + ty * ry; rx = perp * tx; ry = perp * ty; }
        r[i] = rx; r[n + i] = ry;
    }
    x.sync();

    while ((lchanged || dif > difid) && itcg < maxcg) {
        itcg++; it_inn++;
        if (use_fftprec) {
            if (nslip > 3 * nadh && x.plan().my > 1) use_fftprec = false;
            if (itcg >= maxcg / 2) use_fftprec = false;
            if (!use_fftprec) lchanged = true;
        }
        if (use_fftprec) {
            x.conv_int(1, 0, 0, r, c.chatM11, z, el, 0);
            x.conv_int(1, 1, 1, r + n, c.chatM22, z + n, el, 0);
            nprod += 2;
        }
        double d2[2] = { 0.0, 0.0 };
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {                                   // diagonal scaling + tangent projection
            const int e = el[i];
            double zx = use_fftprec ? z[i] : r[i], zy = use_fftprec ? z[n + i] : r[n + i];
            if (e == 2) {
                const double snrm = -ss[i] * nx[i] - ss[n + i] * ny[i];
                zx = zx / (c11 + ga * snrm / g[i]); zy = zy / (c22 + ga * snrm / g[i]);
                const double tx = -ny[i], ty = nx[i], perp = tx * zx + ty * zy;
                zx = perp * tx; zy = perp * ty;
            } else { zx = zx / c11; zy = zy / c22; }
            z[i] = zx; z[n + i] = zy;
            d2[0] += zx * q[i] + zy * q[n + i];
            d2[1] += v[i] * q[i] + v[n + i] * q[n + i];
        }
        x.template sum<2>(d2);
        double beta = 0.0;
        const bool restart = (itcg <= 1 || lchanged);
        if (!restart) {
            const double zq = d2[0], vq = d2[1];
            if (fabs(zq) < 1e-60 || fabs(vq) < small * fabs(zq)) beta = 0.0; else beta = -zq / vq;
        }
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
            if (restart) { v[i] = z[i]; v[n + i] = z[n + i]; }
            else { v[i] = beta * v[i] + z[i]; v[n + i] = beta * v[n + i] + z[n + i]; }
        }
        x.sync();

        nprod += conv_multi(x, c.chatA, v, 0, 1, q, 0, 1, el, 1, 0);       // q = A_tt v on C
        double d3[3] = { 0.0, 0.0, 0.0 };
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
            double qx = q[i], qy = q[n + i];
            const double vx = v[i], vy = v[n + i];
            if (el[i] == 2) {
                const double tx = -ny[i], ty = nx[i], perp = tx * qx + ty * qy;
                qx = perp * tx; qy = perp * ty;
                const double snrm = (nx[i] * ss[i] + ny[i] * ss[n + i]) / g[i];
                qx = qx - snrm * vx; qy = qy - snrm * vy;
                q[i] = qx; q[n + i] = qy;
            }
            d3[0] += r[i] * vx + r[n + i] * vy;
            d3[1] += vx * qx + vy * qy;
            d3[2] += vx * vx + vy * vy;
        }
        x.template sum<3>(d3);
        const double rv = d3[0], vq = d3[1];
        if (fabs(rv) < 1e-60) alpha = 0.0; else if (fabs(vq) < small * fabs(rv)) alpha = 1.0; else alpha = rv / vq;

        double p2[1] = { 0.0 };
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {                                   // :2170-2185
            const int e = el[i];
            double px = ps[i], py = ps[n + i];
            pold[i] = px; pold[n + i] = py;
            if (e >= 1) { px += alpha * v[i]; py += alpha * v[n + i]; }
            if (e == 2) { const double pa = sqrt(px * px + py * py); px = px * g[i] / pa; py = py * g[i] / pa; }
            ps[i] = px; ps[n + i] = py;
            p2[0] += px * px + py * py;
        }
        dif = alpha * facnel * sqrt(d3[2] / (2.0 * n));
        difinn += dif;
        if (itcg <= 5 || itcg % 5 == 1) {
            x.template sum<1>(p2);
            const double ptang = facnel * sqrt(p2[0] / (2.0 * n));
            difid = eps * fmax(1e-6, ptang);
            trsinn = (double) 0.01f * ptang;
        } else x.sync();

        if (it_inn >= num_inn || dif <= difid || difinn > trsinn) {            // :2227 check constraints
            double ch[1] = { 0.0 };
            for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] == 1) {
                const double px = ps[i], py = ps[n + i], pa = sqrt(px * px + py * py);
                if (pa > g[i]) { el[i] = 2; ps[i] = px * g[i] / pa; ps[n + i] = py * g[i] / pa; ch[0] += 1.0; }
            }
            x.sync();
            nprod += conv_multi(x, c.chatA, ps, 0, 1, ss, 0, 1, el, 1, 0);
            double dp[1] = { 0.0 };
            for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
                const int e = el[i];
                if (e >= 1) { ss[i] += ws[i]; ss[n + i] += ws[n + i]; }
                if (e == 2 && ss[i] * ps[i] + ss[n + i] * ps[n + i] > 0.0) { el[i] = 1; ch[0] += 1.0; }
                const double dx = pold[i] - ps[i], dy = pold[n + i] - ps[n + i];
                pold[i] = dx; pold[n + i] = dy;
                dp[0] += dx * dx + dy * dy;
            }
            double both[2] = { ch[0], dp[0] };
            x.template sum<2>(both);
            lchanged = both[0] > 0.0;
            dif = facnel * sqrt(both[1] / (2.0 * n));
            it_inn = 0; difinn = 0.0;
            if (lchanged) count_el(x, el, n, nadh, nslip);
        }
        if ((lchanged || dif > difid) && itcg < maxcg) {                       // :2330-2380
            if (it_inn <= 0) {
                for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
                    const double a = atan2(ps[n + i], ps[i]);
                    const double cx = cos(a), sy = sin(a);
                    nx[i] = cx; ny[i] = sy;
                    double rx = -ss[i], ry = -ss[n + i];
                    if (el[i] == 2) { const double tx = -sy, ty = cx, perp = tx * rx + ty * ry; rx = perp * tx; ry = perp * ty; }
                    r[i] = rx; r[n + i] = ry;
                }
            } else {
                for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] >= 1) { r[i] -= alpha * q[i]; r[n + i] -= alpha * q[n + i]; }
            }
            x.sync();
        }
    }
    itcg_out = itcg; err_out = dif;
}

// |row sum| of a spatial coefficient block at the central element over the columns AijPj visits (see centre_rowsum_dev)
__device__ double centre_rowsum_blk(const ConvPlan &P, const double *blk, int cmx, int cmy, const int *el, double *red)
{
    const int mx = P.mx, my = P.my;
    const int ixm = max(1, mx / 2), iym = max(1, my / 2);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double s[1] = { 0.0 };
    for (int jy = 1 + wid; jy <= my; jy += nw) {
        int first = mx + 1, last = 0;
        for (int jx = 1 + lane; jx <= mx; jx += 32)
            if (el[(jy - 1) * mx + jx - 1] >= 1) { first = min(first, jx); last = max(last, jx); }
        for (int o = 16; o > 0; o >>= 1) {
            first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        }
        if (last == 0) first = mx;
        const int j0 = max(1, first - 1), j1 = min(mx, last + 1);
        const double *row = blk + (size_t) (iym - jy + cmy) * (2 * cmx) + cmx;
        for (int jx = j0 + lane; jx <= j1; jx += 32) s[0] += row[ixm - jx];
    }
    block_sum<1>(s, red);
    return s[0];
}

// one tangential solve + relative forces (+ log of the Newton-Raphson process)
template <class X>
__device__ int solve_once_dev(const X &x, ContactCase &c, const double *wstot, double fntrue,
                              int &it, double &err, double &fx, double &fy, int &nprod)
{
    const int n = x.n();
    int info = 0;
    bool use_gs = (c.solver_eff == 1 || c.solver_eff == 2);
    if (c.solver_eff == 3) {                                                   // GDsteady, tang_solver m_solvpt.f90:459-484
        int lstagn = 0;
        it = gdsteady_dev(x, c, wstot, c.nrm.maxgs, c.nrm.eps, err, lstagn, nprod);
        if (lstagn) {                                                          // stagnation: SteadyGS takes over
            if (x.leader()) c.gd_fallback++;
            x.sync();
            use_gs = true;
        }
    }
    if (use_gs) {                                                              // SteadyGS / ConvexGS
        {
        int nadh, nslip;
        count_el(x, c.nrm.el, n, nadh, nslip);
        const int ncon = nadh + nslip;
        const bool convex = c.solver_eff == 2, sv = convex && c.tang == 3;     // csv in steady rolling, cs otherwise
        const cd *chat_sv[3][3] = { { c.chatSV[0][0], c.chatSV[0][1], nullptr }, { c.chatSV[1][0], c.chatSV[1][1], nullptr },
                                    { nullptr, nullptr, nullptr } };
        SteadyArgs a;
        a.ws = wstot; a.dp = c.twork + 3 * (size_t) n; a.ug = c.twork + 5 * (size_t) n;
        a.iel = reinterpret_cast<int *>(c.twork + 7 * (size_t) n);
        a.isp = reinterpret_cast<int *>(c.twork + 8 * (size_t) n);          // (TangCG's vectors live here; it does not run beside a sweep)
        a.chatA = sv ? chat_sv : c.chatA;
        a.cf11 = sv ? c.cfv11 : c.cf11; a.cf12 = sv ? c.cfv12 : c.cf12; a.cf22 = sv ? c.cfv22 : c.cf22;
        a.cmx = c.nrm.cmx; a.cmy = c.nrm.cmy;
        a.ga_inv = c.nrm.ga_inv; a.mu = c.fstat; a.eps = c.nrm.eps; a.omegah = c.omegah; a.omegas = c.omegas; a.maxgs = c.nrm.maxgs;
        a.convex = convex ? 1 : 0; a.sym = sv ? 0 : 1;
        a.ledge = (sv && c.dq > c.dx) ? 1 : 0; a.facdt = c.twork;             // facdt of stang_dev
        a.cs11 = c.cf11; a.cs12 = c.cf12; a.cs22 = c.cf22; a.cs13 = c.cf13; a.cs23 = c.cf23; a.ub = a.ug;
        if constexpr (X::kBlock) {
            const int nown = CB_THREADS - 32;                    // warp 0 owns no elements (it walks the rows)
            if (ncon <= 6 * nown) info = stdygs_dev<6>(x.plan(), x.smem(), a, c.nrm.el, c.ps, c.ss, ncon, it, err, nprod);
            else if (ncon <= 12 * nown) info = stdygs_dev<12>(x.plan(), x.smem(), a, c.nrm.el, c.ps, c.ss, ncon, it, err, nprod);
            else if (ncon <= 22 * nown) info = stdygs_dev<22>(x.plan(), x.smem(), a, c.nrm.el, c.ps, c.ss, ncon, it, err, nprod);
            else info = stdygs_dev<1, true>(x.plan(), x.smem(), a, c.nrm.el, c.ps, c.ss, ncon, it, err, nprod);   // direct row sums
        } else {
            // whole-GPU path: the sweep is one sequential chain of element steps, each with an O(ncon) row sum -- CTA 0 runs it
            // (direct form), the other CTAs wait at the grid barrier; the result reaches them through the case record
            if (blockIdx.x == 0) {
                Smem sm0;
                sm0.S = nullptr; sm0.W = nullptr; sm0.twx = nullptr; sm0.twy = nullptr; sm0.posx = nullptr; sm0.red = x.red(); sm0.a0 = 0;
                int it0 = 0; double err0 = 0.0;
                const int inf0 = stdygs_dev<1, true>(x.plan(), sm0, a, c.nrm.el, c.ps, c.ss, ncon, it0, err0, nprod, x.sraw + x.X.gs_off);
                if (threadIdx.x == 0) { c.gs_info = inf0; c.gs_it = it0; c.gs_err = err0; }
            }
            x.sync();
            info = *reinterpret_cast<volatile int *>(&c.gs_info); it = *reinterpret_cast<volatile int *>(&c.gs_it);
            err = *reinterpret_cast<volatile double *>(&c.gs_err);
            x.sync();
        }
        }
    } else if (c.solver_eff == 0)
        tangcg_dev(x, c, wstot, c.nrm.maxgs, c.nrm.eps, it, err, nprod);
    double s[2] = { 0.0, 0.0 };
    for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) { s[0] += c.ps[i]; s[1] += c.ps[n + i]; }
    x.template sum<2>(s);
    fx = c.nrm.dxdy * s[0] / (c.fstat * fntrue);
    fy = c.nrm.dxdy * s[1] / (c.fstat * fntrue);
    if (x.leader() && c.nr_n < CB_MAXNR_LOG) {
        const int k = c.nr_n;
        c.nr_itcg[k] = it; c.nr_cksi[k] = c.cksi; c.nr_ceta[k] = c.ceta; c.nr_fx[k] = fx; c.nr_fy[k] = fy;
    }
    x.sync();
    if (x.leader()) c.nr_n++;
    x.sync();
    return info;
}

// solvpt (m_solvpt.f90:51-378): Newton-Raphson on (cksi[, ceta]) for prescribed tangential forces; shifts: dq = 1
template <class X>
__device__ int solvpt_dev(const X &x, ContactCase &c, const double *wsfix, double *wstot,
                          const double *facdt, double fntrue, int &itgs, double &err, int &nprod)
{
    const int n = x.n();
    const int *el = c.nrm.el;
    const double dq = c.dq, dxdy = c.nrm.dxdy, muscal = c.fstat, eps = c.nrm.eps;
    double cksi = c.cksi, ceta = c.ceta;                     // block-uniform copies; c.cksi/c.ceta updated by thread 0
    int it, nadh, nslip;
    double fxkp1, fykp1;
    itgs = 0;
    for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
        double wx = wsfix[i], wy = wsfix[n + i];
        if (el[i] >= 1) {
            const double f = facdt ? facdt[i] : 1.0;
            if (c.force3 >= 1) wx += f * cksi * dq;
            if (c.force3 >= 2) wy += f * ceta * dq;
        }
        wstot[i] = wx; wstot[n + i] = wy;
    }
    x.sync();
    int info = solve_once_dev(x, c, wstot, fntrue, it, err, fxkp1, fykp1, nprod);
    itgs += it;
    count_el(x, el, n, nadh, nslip);
    if (c.force3 >= 1 && info <= 1) {
        int itnr = 0;
        double df = fabs(c.fxrel - fxkp1);
        if (c.force3 >= 2) df += fabs(c.fyrel - fykp1);
        double s00 = c.sens[0][0], s01 = c.sens[0][1], s10 = c.sens[1][0], s11 = c.sens[1][1];
        while (df > eps && itnr < c.maxnr && info <= 1) {
            itnr++;
            const double dfx = c.fxrel - fxkp1, dfy = c.fyrel - fykp1;
            double dcksi, dceta;
            if (c.force3 == 1) {
                dceta = 0.0;
                dcksi = (fabs(s00) > (double) 1e-6f) ? dfx / s00 : (double) 0.00003f;
            } else {
                const double det = s00 * s11 - s10 * s01;
                if (det > eps) { dcksi = (s11 * dfx - s10 * dfy) / det; dceta = (-s01 * dfx + s00 * dfy) / det; }
                else { dcksi = 0.000003; dceta = 0.000003; }
            }
            for (int ifxy = 1; ifxy <= c.force3; ifxy++) {
                const double fxk = fxkp1, fyk = fykp1;
                if (ifxy == 1) { cksi += dcksi; for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] >= 1) wstot[i] += (facdt ? facdt[i] : 1.0) * dcksi * dq; }
                else { ceta += dceta; for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] >= 1) wstot[n + i] += (facdt ? facdt[i] : 1.0) * dceta * dq; }
                x.sync();
                if (x.leader()) { c.cksi = cksi; c.ceta = ceta; }
                x.sync();
                info = solve_once_dev(x, c, wstot, fntrue, it, err, fxkp1, fykp1, nprod);
                itgs += it;
                count_el(x, el, n, nadh, nslip);
                const double dfxk = fxkp1 - fxk, dfyk = fykp1 - fyk, ncon = (double) (nadh + nslip);
                if (ifxy == 1) {
                    if (nadh > 0) {
                        const double dpx = dfxk * muscal * fntrue / (ncon * dxdy), dpy = dfyk * muscal * fntrue / (ncon * dxdy);
                        if (s00 == 0.0 || fabs(dpx) > 10.0 * err) s00 = dfxk / dcksi;
                        if (fabs(dpy) > 10.0 * err) s10 = dfyk / dcksi;
                    } else if (fabs(s00) < (double) 1e-6f) s00 = fxkp1 / cksi;
                } else {
                    if (nadh > 0) {
                        const double dpx = dfxk * muscal * fntrue / (ncon * dxdy), dpy = dfyk * muscal * fntrue / (ncon * dxdy);
                        if (fabs(dpx) > 10.0 * err) s01 = dfxk / dceta;
                        if (s11 == 0.0 || fabs(dpy) > 10.0 * err) s11 = dfyk / dceta;
                    } else if (fabs(s11) < (double) 1e-6f) s11 = fykp1 / ceta;
                }
                df = fabs(c.fxrel - fxkp1);
                if (c.force3 >= 2) df += fabs(c.fyrel - fykp1);
            }
        }
        err = err + 2.0 * df * muscal * fntrue / ((nadh + 2) * dxdy);          // the reference's constant Slip=2, :361
        if (x.leader()) { c.sens[0][0] = s00; c.sens[0][1] = s01; c.sens[1][0] = s10; c.sens[1][1] = s11; }
    } else if (x.leader()) { c.sens[0][0] = 0.0; c.sens[0][1] = 0.0; c.sens[1][0] = 0.0; c.sens[1][1] = 0.0; }
    if (x.leader()) { c.fx = fxkp1; c.fy = fykp1; }
    x.sync();
    return info;
}

// stang (m_stang.f90:28-746) for shifts with uniform Coulomb friction; returns ittang (-1: MaxIn reached)
template <class X>
__device__ int stang_dev(const X &x, ContactCase &c, double fntrue, int &itgs_tot, int &nprod)
{
    const int n = x.n();
    int *el = c.nrm.el;
    double *ps = c.ps, *ss = c.ss;
    double *wsfix = c.twork + 13 * (size_t) n, *wstot = wsfix + 2 * n, *u1 = wstot + 2 * n, *u2 = u1 + 2 * n;
    const double mu = c.fstat;
    const bool ssrol = (c.tang == 3);
    double *facdt = nullptr;
    {                                                                          // m_stang.f90:129-223
        double cnt[2] = { 0.0, 0.0 };
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] >= 1) { cnt[0] += 1.0; if (i % x.plan().mx == 0) cnt[1] += 1.0; }
        x.template sum<2>(cnt);
        const int k = (int) cnt[0];
        int solver = ssrol ? (c.gausei == 5 ? 3 : (c.gausei != 2 ? 1 : 2)) : (c.gausei != 2 ? 0 : 2);
        if (cnt[1] > 0.0 && (solver == 1 || solver == 3)) solver = 2;          // no exterior elements at the trailing edge
        // the Gauss-Seidel solvers exist on the one-CTA-per-case path only (ConvexGS with dq > dx: leading-edge equations
        // inside stdygs_dev)
        bool refuse = (solver == 3 && c.gwork == nullptr);
        if (refuse) { if (x.leader()) c.tstatus |= 1; x.sync(); itgs_tot = 0; return -1; }
        double oh = c.omegah, os = c.omegas;
        if (c.gausei == 0 || c.gausei == 4 || c.gausei == 5) {
            const double r = c.dx / (c.nrm.dxdy / c.dx);                       // dx / dy
            if (k <= 25) { oh = 1.0; os = 1.0; }
            else if (ssrol && solver == 2) { oh = 0.5; os = 0.5; }
            else if (ssrol) {
                if (r <= 5.0) { oh = 0.9; os = 1.0; }
                else if (r <= 15.0) { oh = 0.8; os = 0.8; }
                else { oh = 0.8; os = 0.6; }
            } else { oh = 0.5; os = 0.5; }
        }
        x.sync();
        if (x.leader()) { c.omegah = oh; c.omegas = os; c.solver_eff = solver; }
        x.sync();
        if (ssrol) {
            facdt = c.twork;
            // sxbnd (m_leadedge.f90:92-332): the leading edge sits 2 dx (SteadyGS) or 1 dx (ConvexGS, GDsteady) beyond the
            // last interior element
            sxbnd_facdt_x(x, x.plan().mx, x.plan().my, el, c.dx, c.dq, solver == 1 ? 2.0 : 1.0, facdt);
        }
    }
    // stang_rhs (:749-951): wsfix = -facdt hs_t + A_tn pn - A'_tn p'n - A'_tt p'_t on C; shifts: facdt = 1, previous
    // tractions p'; steady rolling: p' = p with the shifted coefficients cv, A'_tt p'_t left to the solver
    nprod += conv_multi(x, c.chatA, ps, 2, 2, u1, 0, 1, el, 1, 0);
    // transient rolling with dq > dx: near the leading edge (ii2j > 0, i.e. facdt < 0.9999) equation (1b) replaces the
    // displacements of the previous time u' = A' p' by ubnd = subnd(p', cs) (m_stang.f90:888-925, m_leadedge.f90:336-394):
    // the displacement of p' (all three directions, coefficients cs) at the first exterior element behind the C -> E
    // transition of the row -- taken from ONE unmasked product A_cs p' instead of a row sum per transition
    const bool tledge = (c.tang == 2 && c.dq > c.dx);
    if (ssrol) {
        nprod += conv_multi(x, c.chatV, ps, 2, 2, u2, 0, 1, el, 1);
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] >= 1) { u1[i] -= u2[i]; u1[n + i] -= u2[n + i]; }
        x.sync();
    } else if (c.pv) {
        if (tledge) {
            const int mx = x.plan().mx;
            nprod += conv_multi(x, c.chatA, c.pv, 0, 2, u2, 0, 1, el, 0);
            for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] >= 1) {      // wstot: usn - ubnd, used below
                const int ix = (int) (i % mx);
                int ixb = ix;
                while (ixb < mx - 1 && el[i - ix + ixb + 1] >= 1) ixb++;
                const bool ok = ixb + 3 <= mx;                                               // else ubnd = 0 (:364-366)
                wstot[i] = u1[i] - (ok ? u2[i - ix + ixb + 1] : 0.0);
                wstot[n + i] = u1[n + i] - (ok ? u2[n + i - ix + ixb + 1] : 0.0);
            }
            x.sync();
        }
        nprod += conv_multi(x, c.chatV, c.pv, 2, 2, u2, 0, 1, el, 1);
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] >= 1) { u1[i] -= u2[i]; u1[n + i] -= u2[n + i]; }
        x.sync();
        nprod += conv_multi(x, c.chatV, c.pv, 0, 1, u2, 0, 1, el, 1);
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] >= 1) { u1[i] -= u2[i]; u1[n + i] -= u2[n + i]; }
        x.sync();
        if (tledge) {                                       // u2 is free now: facdt lives there during the TANG loop
            facdt = u2;
            sxbnd_facdt_x(x, x.plan().mx, x.plan().my, el, c.dx, c.dq, 1.0, facdt);
            for (size_t i = x.first(); i < (size_t) (n); i += x.stride())
                if (el[i] >= 1 && facdt[i] < 0.9999) { u1[i] = wstot[i]; u1[n + i] = wstot[n + i]; }
            x.sync();
        }
    } else if (tledge) {                                    // from rest (p' = 0): u' = ubnd = 0, only facdt differs from 1
        facdt = u2;
        sxbnd_facdt_x(x, x.plan().mx, x.plan().my, el, c.dx, c.dq, 1.0, facdt);
    }
    for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
        const bool in = el[i] >= 1;
        const double f = facdt ? facdt[i] : 1.0;
        wsfix[i] = in ? -f * c.hst[i] + u1[i] : 0.0;
        wsfix[n + i] = in ? -f * c.hst[n + i] + u1[n + i] : 0.0;
    }
    x.sync();

    int ittang = 0, it;
    bool zready = false;
    double errpt = 0.0;
    itgs_tot = 0;
    while (!zready && ittang < c.nrm.maxin) {
        ittang++;
        zready = true;
        const int info = solvpt_dev(x, c, wsfix, wstot, facdt, fntrue, it, errpt, nprod);
        itgs_tot += it;
        if (info >= 3) { ittang = -1; break; }                                 // :420-427 divergence
        double k[1] = { 0.0 };
        const double tol = sqrt(2.0) * errpt;
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] == 1) {                    // :434-453 adhesion -> slip
            const double px = ps[i], py = ps[n + i], pn = ps[2 * (size_t) n + i], pabs = sqrt(px * px + py * py);
            if (pabs >= mu * pn + tol) { el[i] = 2; ps[i] = px * mu * pn / pabs; ps[n + i] = py * mu * pn / pabs; k[0] += 1.0; }
        }
        x.template sum<1>(k);
        if (k[0] > 0.0) zready = false;
        if (zready) {                                                          // :463-510 slip -> adhesion
            const double tol1 = errpt * 2.0 * fabs(centre_rowsum_blk(x.plan(), c.cf11, c.nrm.cmx, c.nrm.cmy, el, x.red()) * c.nrm.ga_inv);
            const double tol2 = errpt * 2.0 * fabs(centre_rowsum_blk(x.plan(), c.cf22, c.nrm.cmx, c.nrm.cmy, el, x.red()) * c.nrm.ga_inv);
            double ka[1] = { 0.0 };
            for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] == 2) {
                const double ww = ss[i] * ps[i] + ss[n + i] * ps[n + i];
                const double tl = tol1 * fabs(ps[i]) + tol2 * fabs(ps[n + i]) + errpt * (fabs(ss[i]) + fabs(ss[n + i]));
                if (ww > tl) { el[i] = 1; ka[0] += 1.0; }
            }
            x.template sum<1>(ka);
            if (ka[0] > 0.0) zready = false;
            if (info > 0) zready = true;                                       // :510
        }
    }
    if (!zready) ittang = -1;
    return ittang;
}

// contac's computing part: panprc (m_scontc.f90:356-553) = NORM / TANG alternation until the tractions settle
template <class X>
__device__ void panprc_dev(const X &x, ContactCase &c)
{
    const int n = x.n();
    double *ps = c.ps, *po1 = c.twork + 21 * (size_t) n;
    int *el = c.nrm.el;
    int itnorm = 0, ittang = 0, itout = 0, itcg = 0, itgs = 0, nprod = 0;
    double dif = 200.0, difid = 1.0;
    if (x.leader()) c.nr_n = 0;
    for (size_t i = x.first(); i < (size_t) (3 * n); i += x.stride()) po1[i] = ps[i];
    x.sync();
    while (dif > difid && itout < c.maxout && itnorm >= 0 && ittang >= 0) {
        itout++;
        x.snorm(c.nrm);
        x.update_box(c.nrm, el);
        itcg += c.nrm.itcg; nprod += c.nrm.nprod;
        if (c.nrm.itnorm >= 0) itnorm += c.nrm.itnorm; else itnorm = -1;
        const int ncon = c.nrm.ncon;
        for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) if (el[i] < 1) { ps[i] = 0.0; ps[n + i] = 0.0; }
        x.sync();
        if (c.tang == 0 || ncon <= 0) dif = 0.0;
        else {
            int it_gs;
            const int it = stang_dev(x, c, c.nrm.fntrue, it_gs, nprod);
            itgs = it_gs;                                                      // solv%itgs is that of the last TANG call (m_stang.f90:274,709)
            if (it >= 0) ittang += it; else ittang = -1;
            double s[3] = { 0.0, 0.0, 0.0 };
            for (size_t i = x.first(); i < (size_t) (n); i += x.stride()) {
                const bool in = el[i] >= 1;
                for (int k = 0; k < 3; k++) {
                    const double pk = ps[(size_t) k * n + i], d = po1[(size_t) k * n + i] - pk;
                    if (in) { s[0] += d * d; s[1] += pk * pk; s[2] += 1.0; }
                    po1[(size_t) k * n + i] = pk;
                }
            }
            x.template sum<3>(s);
            dif = sqrt(s[0] / fmax(1.0, s[2]));
            difid = 5.0 * c.nrm.eps * sqrt(s[1] / fmax(1.0, s[2]));
        }
        if (x.leader() && itout <= 16) { c.pan_dif[itout - 1] = dif; c.pan_difid[itout - 1] = difid; }
    }
    int nadh, nslip;
    count_el(x, el, n, nadh, nslip);
    if (x.leader()) {
        c.nrm.itnorm = itnorm; c.nrm.itcg = itcg; c.nrm.nprod = nprod;
        c.ittang = ittang; c.itgs = itgs; c.itout = itout; c.nadh = nadh; c.nslip = nslip;
    }
    x.sync();
}

}  // namespace cb200
