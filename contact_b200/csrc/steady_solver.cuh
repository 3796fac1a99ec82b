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
// the Gauss-Seidel step (plstrc + re-integration of the current row) is walked by warp 0 alone, in an instantiation of
// its own (gs_sweeps<1, false, true>: no U registers, no local memory on the chain), while warps 1.. own the contact
// elements and apply the finished rows' net changes (gs_sweeps<KMAX, false, false>); same barrier sequence in both.
// U is (re)computed from scratch with four FFT products at the start of every solver call.
#pragma once
#include "norm_solver.cuh"

namespace cb200 {

enum { EL_EXTER = 0, EL_ADHES = 1, EL_SLIP = 2, EL_PLAST = 3 };

// cycle counters of the SteadyGS step (thread 0 of every CTA adds its totals at the end of a solver call):
// [0] element steps, [1] per-element solve (plstrc), [2] re-integration of the row, [3] in-row rank-1 updates,
// [4] solver calls, [5] per-row updates of all other rows (whole CTA)
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

// leading-edge factor facdt (m_leadedge.f90:92-332, chi = 0): 0 in the exterior, 1 near the end of the grid,
// min(1, (xbnd - x)/dq) otherwise with xbnd fxdfac elements beyond the next C->E transition (2 without the leading-edge
// correction -- SteadyGS --, 1 with it -- ConvexGS)
__device__ void sxbnd_facdt_dev(int mx, int my, const int *el, double dx, double dq, double fxdfac, double *facdt)
{
    for (int iy = threadIdx.x; iy < my; iy += blockDim.x) {
        const int *e = el + (size_t) iy * mx;
        double *f = facdt + (size_t) iy * mx;
        int ixb = -1;                                   // next transition at or to the right of ix (0-based), found lazily
        for (int ix = 0; ix < mx; ix++) {
            if (e[ix] < 1) { f[ix] = 0.0; continue; }
            if (ixb < ix) { ixb = ix; while (ixb < mx - 1 && e[ixb + 1] >= 1) ixb++; }
            if (ix + 3 > mx) f[ix] = 1.0;
            else f[ix] = fmin(1.0, (((double) (ixb - ix) + fxdfac) * dx) / dq);
        }
    }
    __syncthreads();
}

struct SteadyArgs {
    const double *ws;       // [2][n] right-hand side
    double *dp;             // [2][n] traction differences (global scratch)
    double *ug;             // [2][n] scratch for the FFT evaluation of U
    int *iel;               // [n] compact list of contact elements
    int *isp;               // direct form: [my + 2] row offsets, then the list of all elements of the column ranges
                            // [row1st - 1, rowlst + 1] (the reference's range of gf3_AijPj: the traction differences of SteadyGS do
                            // not vanish at the exterior element in front of a contact run); room for my + 2 + n ints
    const cd *(*chatA)[3];  // transformed cs blocks
    const double *cf11, *cf12, *cf22;
    int cmx, cmy;
    double ga_inv, mu, eps, omegah, omegas;
    int maxgs;
    int convex;             // 0: SteadyGS on the traction differences; 1: ConvexGS on the tractions (cnvxgs)
    int sym;                // coefficient blocks have the quadrant symmetry of cs (false for csv = cs - cv)
    // leading-edge equations of cnvxgs in steady rolling with dq > dx (m_solvpt.f90:2583, 2632-2668; m_leadedge.f90:336-394)
    int ledge;              // 1: elements with facdt < 0.9999 use s = ws + A_cs p_t - ubnd and the 2x2 matrix of cs
    const double *facdt;    // [n] fraction of the step spent in contact (sxbnd), defines ii2j > 0
    const double *cs11, *cs12, *cs22, *cs13, *cs23;   // spatial blocks of cs (13, 23: null without n-t coupling)
    double *ub;             // [2][n] ubnd per leading-edge position, stored at the last interior element of the run
};

// shared-memory carve-up for the sweep.  q: coefficient table (quadrant [3][my][mx] for the symmetric cs, half plane
// [3][my][2mx] for csv, null when it does not fit: blocks are then read from global memory); r0: the dy = 0 rows of the
// three blocks for dx in [-mx, mx) (in-row updates and the element's own 2x2 matrix)
struct SteadySmem {
    // All arrays are addressed as 32-bit element offsets into the CTA's shared-memory window (typed ld.shared / st.shared: the
    // sweep is one warp's dependent instruction chain, and generic 64-bit addressing cost a third of its instructions).
    uint32_t oq, or0, orow;    // (doubles) table (hasq), dy = 0 rows [3][2mx], per-row arrays [15][mx] + 8 scalars
    uint32_t oirow;            // (ints) int arrays [3][mx], then rowk[my+2], ictl[8]
    uint32_t olst;             // (doubles) two net-change lists of finished rows [2][2][mx] (x, y): pipelined update of the other rows
    uint32_t oilst;            // (ints) their element positions [2][mx] and lengths [2]
    int hasq;
    __device__ __forceinline__ double *wd() const { return reinterpret_cast<double *>(cb_smem_window); }
    __device__ __forceinline__ int *wi() const { return reinterpret_cast<int *>(cb_smem_window); }
    __device__ __forceinline__ double *q() const { return wd() + oq; }
    __device__ __forceinline__ double *r0() const { return wd() + or0; }
    __device__ __forceinline__ double *row() const { return wd() + orow; }
    __device__ __forceinline__ int *irow() const { return wi() + oirow; }
    __device__ __forceinline__ double &lx(int b, int j) const { return wd()[olst + (2 * b) * mx + j]; }
    __device__ __forceinline__ double &ly(int b, int j) const { return wd()[olst + (2 * b + 1) * mx + j]; }
    __device__ __forceinline__ int &lj(int b, int j) const { return wi()[oilst + b * mx + j]; }
    __device__ __forceinline__ int &lcnt(int b) const { return wi()[oilst + 2 * mx + b]; }
    int mx;
    // accessors: one base pointer + multiples of mx (keeps the register footprint small)
    __device__ __forceinline__ double &psx(int j) const { return row()[j]; }
    __device__ __forceinline__ double &psy(int j) const { return row()[mx + j]; }
    __device__ __forceinline__ double &dpx(int j) const { return row()[2 * mx + j]; }
    __device__ __forceinline__ double &dpy(int j) const { return row()[3 * mx + j]; }
    __device__ __forceinline__ double &bnd(int j) const { return row()[4 * mx + j]; }
    __device__ __forceinline__ double &wsx(int j) const { return row()[5 * mx + j]; }
    __device__ __forceinline__ double &wsy(int j) const { return row()[6 * mx + j]; }
    __device__ __forceinline__ double &ssx(int j) const { return row()[7 * mx + j]; }
    __device__ __forceinline__ double &ssy(int j) const { return row()[8 * mx + j]; }
    __device__ __forceinline__ double &urx(int j) const { return row()[9 * mx + j]; }
    __device__ __forceinline__ double &ury(int j) const { return row()[10 * mx + j]; }
    __device__ __forceinline__ double &ddx(int j) const { return row()[11 * mx + j]; }
    __device__ __forceinline__ double &ddy(int j) const { return row()[12 * mx + j]; }
    __device__ __forceinline__ double &chx(int j) const { return row()[13 * mx + j]; }
    __device__ __forceinline__ double &chy(int j) const { return row()[14 * mx + j]; }
    __device__ __forceinline__ double &scal(int j) const { return row()[15 * mx + j]; }
    __device__ __forceinline__ int &el(int j) const { return irow()[j]; }
    __device__ __forceinline__ int &chj(int j) const { return irow()[mx + j]; }
    __device__ __forceinline__ int &cix(int j) const { return irow()[2 * mx + j]; }
    __device__ __forceinline__ int &rowk(int j) const { return irow()[3 * mx + j]; }
    __device__ __forceinline__ int &ictl(int j, int my) const { return irow()[3 * mx + my + 2 + j]; }
};

// bytes of the row arrays etc. (without the coefficient table)
__host__ __device__ __forceinline__ size_t steady_fixed_bytes(int mx, int my)
{
    return ((size_t) (6 * mx + 15 * mx + 8)) * 8 + ((((size_t) (3 * mx + my + 10)) * 4 + 7) & ~(size_t) 7) + (size_t) (4 * mx) * 8 +
           (size_t) (2 * mx + 4) * 4 + 64;
}

// extra dynamic shared memory that a launch must provide beyond the FFT plan's: for small grids the S/W regions of the
// plan are too small for the sweep arrays, which then live behind the plan's own layout
__host__ __device__ __forceinline__ size_t steady_extra_smem(const ConvPlan &P)
{
    const size_t fixed = steady_fixed_bytes(P.mx, P.my);
    return fixed <= (size_t) P.off_twx ? 0 : fixed;
}

// offsets of the sweep arrays for a region that starts at `base` (a pointer into this CTA's dynamic shared memory)
__device__ __forceinline__ void steady_offsets(unsigned char *base, size_t tab_doubles, int mx, int my, SteadySmem &s)
{
    const uint32_t b = (uint32_t) __cvta_generic_to_shared(base) - (uint32_t) __cvta_generic_to_shared(cb_smem_window);
    uint32_t d = b >> 3;                                            // base is 16-byte aligned
    s.hasq = tab_doubles > 0 ? 1 : 0;
    s.oq = d; d += (uint32_t) tab_doubles;
    s.or0 = d; d += 6u * mx;
    s.orow = d;
    s.oirow = (d + 15u * mx + 8u) * 2u;
    s.olst = (((s.oirow + 3u * mx + my + 10u) * 4u + 7u) & ~7u) >> 3;
    s.oilst = (s.olst + 4u * mx) * 2u;
    s.mx = mx;
}

__device__ __forceinline__ void steady_carve(const ConvPlan &P, unsigned char *base, int sym, SteadySmem &s)
{
    const size_t n = P.npot, mx = P.mx, my = P.my;
    const size_t fixed = steady_fixed_bytes(P.mx, P.my);
    const size_t tabsz = (sym ? 3 : 6) * n * 8;
    const bool tab = tabsz + fixed <= (size_t) P.off_twx;
    if (fixed > (size_t) P.off_twx) base += P.smem_bytes;      // behind the plan's layout (see steady_extra_smem)
    steady_offsets(base, tab ? (sym ? 3 : 6) * n : 0, (int) mx, (int) my, s);
}

// load from a table in shared memory that is constant during the sweeps, by its 32-bit shared-space byte address (no address
// derivation per access, free to be scheduled across the sweep's stores)
__device__ __forceinline__ double lds_const(uint32_t addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// coefficient lookup A_tt(dx, dy): c11, c12 (= c21), c22, already times 1/G
struct SteadyTab {
    uint32_t oq, or0;                       // (doubles) offsets of the table (0xffffffff: none) and of the dy = 0 rows in the shared window
    const double *cf11, *cf12, *cf22;
    int n, mx, cmx, cmy, sym;
    double ga_inv;
    __device__ __forceinline__ void get(int dx, int dy, double &c11, double &c12, double &c22) const
    {
        const double *q = reinterpret_cast<const double *>(cb_smem_window) + oq;
        if (oq != 0xffffffffu) {
            if (sym) {                                      // even (c11, c22) / odd (c12) in x and in y
                const int o = abs(dy) * mx + abs(dx);
                c11 = q[o]; c12 = q[n + o]; c22 = q[2 * n + o];
                if ((dx < 0) != (dy < 0)) c12 = -c12;
            } else {                                        // csv: symmetric in y only
                const int o = abs(dy) * 2 * mx + dx + mx;
                c11 = q[o]; c12 = q[2 * n + o]; c22 = q[4 * n + o];
                if (dy < 0) c12 = -c12;
            }
        } else {
            const size_t o = (size_t) (dy + cmy) * (2 * cmx) + dx + cmx;
            c11 = cf11[o] * ga_inv; c12 = cf12[o] * ga_inv; c22 = cf22[o] * ga_inv;
        }
    }
    __device__ __forceinline__ void row0(int dx, double &c11, double &c12, double &c22) const
    { const double *r0 = reinterpret_cast<const double *>(cb_smem_window) + or0; const int o = dx + mx; c11 = r0[o]; c12 = r0[2 * mx + o]; c22 = r0[4 * mx + o]; }
};

// profiling counters of the element step (tools/steady_timing*.py): compiled in with -DCB_STEADY_PROF only -- six 64-bit counters and
// four clock reads per element step cost the walking warp ~25 local-memory accesses per step (they do not fit its registers)
#ifdef CB_STEADY_PROF
#define CB_GS_PROF(...) __VA_ARGS__
#else
#define CB_GS_PROF(...)
#endif

// row-invariant state of a sweep for the warp that walks the rows
struct GsWalk {
    SteadySmem s;
    SteadyTab T;
    const SteadyArgs *a;
    const double *xp, *psx, *psy;      // global memory: traction differences / tractions of the other rows (direct form, leading edge)
    double *red;
    double q00, q01, q11, l00, l01, l11;
    int n, mx, my, ncon, nsp;
    uint32_t mg_mx;
    int convex, ledge;
};

// The Gauss-Seidel steps of grid row iy (contact elements k0..k1-1 of the compact list) and the compaction of the row's net changes
// into list buffer `par`: called by warp 0 alone (register form) or by the whole CTA (DIRECT: block-wide row sum before every step).
// Inlined into the WALKER / DIRECT instantiations of gs_sweeps (see there).  Returns the updated sum of squared changes.
template <bool DIRECT>
__device__ __forceinline__ double gs_walk_row(const GsWalk &w, const int iy, const int k0, const int k1, const int itgs, const int par,
                                           double dsum, unsigned long long *tp)
{
    const SteadySmem s = w.s;
    const SteadyTab T = w.T;
    const SteadyArgs &a = *w.a;
    const int n = w.n, mx = w.mx, my = w.my, ncon = w.ncon, nsp = w.nsp, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const uint32_t mg_mx = w.mg_mx;
    const unsigned full = 0xffffffffu;
    const bool convex = w.convex != 0, ledge = w.ledge != 0;
    const double *xp = w.xp, *psx = w.psx, *psy = w.psy;
    double *red = w.red;
    const double q00 = w.q00, q01 = w.q01, q11 = w.q11, l00 = w.l00, l01 = w.l01, l11 = w.l11;
    (void) tp; (void) ncon; (void) nsp; (void) nt; (void) mg_mx; (void) xp; (void) red;
    // rows of up to 96 elements: the row's own U and its accumulated net changes live in the registers of warp 0
    // (lane l holds elements l, l + 32, l + 64) for the whole walk -- no shared-memory round trip per change
    const bool rr3 = !DIRECT && mx <= 96;
    // byte addresses of the dy = 0 rows (constant during the sweeps): entry dx of block b is at r0a + 8 dx + r0s b
    const uint32_t r0a = (uint32_t) __cvta_generic_to_shared(cb_smem_window) + 8u * (T.or0 + (uint32_t) mx), r0s = 16u * (uint32_t) mx;
    uint32_t r0l[3];                                // ... of this lane's three columns of the row
#pragma unroll
    for (int r = 0; r < 3; r++) r0l[r] = r0a + 8u * (uint32_t) min(lane + 32 * r, mx - 1);
    double urx_[3] = { 0.0, 0.0, 0.0 }, ury_[3] = { 0.0, 0.0, 0.0 }, ddx_[3] = { 0.0, 0.0, 0.0 }, ddy_[3] = { 0.0, 0.0, 0.0 };
    if (rr3) {
#pragma unroll
        for (int r = 0; r < 3; r++) { const int ixp = min(lane + 32 * r, mx - 1); urx_[r] = s.urx(ixp); ury_[r] = s.ury(ixp); }
    }
    for (int k = k0; k < k1; k++) {
        if constexpr (DIRECT) {                    // whole CTA: row sum of element k over the contact list
            const int ixd = s.cix(k - k0);
            double us[2] = { 0.0, 0.0 };
            const int *spl = a.isp + my + 2;
            // batches of 6 pairs per thread with all loads of a batch in flight at once: the list, the tractions of the
            // other rows and the coefficients come from L2 (two dependent round trips per pair otherwise)
            for (int kk = tid; kk < nsp; kk += 6 * nt) {
                int jj[6];
#pragma unroll
                for (int u = 0; u < 6; u++) jj[u] = (kk + u * nt < nsp) ? spl[kk + u * nt] : -1;
                double qx[6], qy[6], c11[6], c12[6], c22[6];
#pragma unroll
                for (int u = 0; u < 6; u++) {
                    const int j = jj[u] < 0 ? 0 : jj[u];
                    const int jy = (int) fdiv((uint32_t) j, mg_mx), jx = j - jy * mx;
                    const size_t o = (size_t) (iy - jy + a.cmy) * (2 * a.cmx) + (ixd - jx) + a.cmx;
                    if (jy == iy) { qx[u] = convex ? s.psx(jx) : s.dpx(jx); qy[u] = convex ? s.psy(jx) : s.dpy(jx); }
                    else { qx[u] = xp[j]; qy[u] = xp[n + j]; }
                    c11[u] = a.cf11[o]; c12[u] = a.cf12[o]; c22[u] = a.cf22[o];
                }
#pragma unroll
                for (int u = 0; u < 6; u++)
                    if (jj[u] >= 0) { us[0] += c11[u] * qx[u] + c12[u] * qy[u]; us[1] += c12[u] * qx[u] + c22[u] * qy[u]; }
            }
            block_sum<2>(us, red);
            if (tid == 0) { s.urx(ixd) = us[0] * a.ga_inv; s.ury(ixd) = us[1] * a.ga_inv; }
            __syncthreads();
        }
        if (!DIRECT || tid < 32) {
        CB_GS_PROF(const unsigned long long ta = clock64();)
        const int ix = s.cix(k - k0);
        int e = s.el(ix);
        // cnvxgs :2626: after 1000 iterations the slip elements are skipped every other iteration
        const bool active = (!convex || itgs <= 1000 || (itgs & 1) == 0 || e == EL_ADHES);
        // run of adhesion elements directly to the left of ix (lanes look at ix-1, ix-2, ...): gives the
        // element jx of the 2x2 matrix (stdygs :3010-3040) and the first window of the re-integration
        int ej0 = -1, L0 = 0, jxs = ix - 1;
        if (!convex) {
            const int jj = ix - 1 - lane;
            ej0 = (jj >= 0) ? s.el(jj) : -1;
            const unsigned na = __ballot_sync(full, ej0 != EL_ADHES);
            L0 = na ? __ffs(na) - 1 : 32;
            jxs = ix - 1 - L0;
            if (L0 == 32) { while (jxs > 0 && s.el(jxs) == EL_ADHES) jxs--; }
            if (jxs < 0) jxs = 0;
        }
        // leading-edge element (ii2j > 0): the shift is A_cs p_t - ubnd instead of A_csv p_t (:2654-2658); the
        // row sum over the contact list is taken by the warp -- this row from shared memory, the others
        // from the tractions in global memory, which are current after every row
        const bool zl = ledge && active && a.facdt[iy * mx + ix] < 0.9999;
        double lsx = 0.0, lsy = 0.0;
        if (zl) {
            for (int kk = lane; kk < ncon; kk += 32) {
                const int jj = a.iel[kk], jy = jj / mx, jx = jj - jy * mx;
                const size_t o = (size_t) (iy - jy + a.cmy) * (2 * a.cmx) + (ix - jx) + a.cmx;
                const double qx = (jy == iy) ? s.psx(jx) : psx[jj], qy = (jy == iy) ? s.psy(jx) : psy[jj];
                const double c12 = a.cs12[o];
                lsx += a.cs11[o] * qx + c12 * qy; lsy += c12 * qx + a.cs22[o] * qy;
            }
            for (int o = 16; o > 0; o >>= 1) { lsx += __shfl_xor_sync(full, lsx, o); lsy += __shfl_xor_sync(full, lsy, o); }
            lsx *= a.ga_inv; lsy *= a.ga_inv;
        }
        double px = 0.0, py = 0.0;
        double ucx, ucy;                                // U of this element
        if (rr3) {
            const int rq = ix >> 5;
            const double vx = rq == 0 ? urx_[0] : (rq == 1 ? urx_[1] : urx_[2]), vy = rq == 0 ? ury_[0] : (rq == 1 ? ury_[1] : ury_[2]);
            ucx = __shfl_sync(full, vx, ix & 31); ucy = __shfl_sync(full, vy, ix & 31);
        } else { ucx = s.urx(ix); ucy = s.ury(ix); }
        const double pox = s.psx(ix), poy = s.psy(ix);  // all lanes: the element's own change is applied from registers (rr3)
        __syncwarp();                                   // every lane has read el, ps of this element before lane 0 rewrites them
        if (lane == 0) {
            s.ictl(0, my) = 0;
            if (active) {
                double c00, c01, c11;
                if (zl) { c00 = l00; c01 = l01; c11 = l11; }
                else if (convex) { c00 = q00; c01 = q01; c11 = q11; }   // cnvxgs: coefs / coefsv, :2519-2541
                else {                                              // stdygs: c(0) - c(jx - ix), :3010-3040
                    const uint32_t ad = r0a + 8u * (uint32_t) (jxs - ix);
                    const double t00 = lds_const(ad), t01 = lds_const(ad + r0s), t11 = lds_const(ad + 2u * r0s);
                    c00 = q00 - t00; c01 = q01 - t01; c11 = q11 - t11;
                }
                double sx = s.wsx(ix) + ucx, sy = s.wsy(ix) + ucy;
                if (zl) {
                    int ixb = ix;
                    while (ixb < mx - 1 && s.el(ixb + 1) >= 1) ixb++;
                    sx = s.wsx(ix) + lsx - a.ub[iy * mx + ixb]; sy = s.wsy(ix) + lsy - a.ub[n + iy * mx + ixb];
                }
                px = pox; py = poy;
                plstrc_dev(e, c00, c01, c01, c11, a.eps, a.omegah, a.omegas, px, py, s.bnd(ix), sx, sy);
                const double ex = px - pox, ey = py - poy;
                dsum += ex * ex + ey * ey;
                if (!rr3 && (ex != 0.0 || ey != 0.0)) { s.chj(0) = ix; s.chx(0) = ex; s.chy(0) = ey; s.ictl(0, my) = 1; }
                s.psx(ix) = px; s.psy(ix) = py; s.ssx(ix) = sx; s.ssy(ix) = sy; s.el(ix) = e;
                if (!convex) { s.dpx(ix) += ex; s.dpy(ix) += ey; }
            }
        }
        CB_GS_PROF(const unsigned long long tb = clock64();)
        __syncwarp();
        if (rr3 && active) {
            // the element's own change, first in the order of the change list: every lane forms it from the broadcast new traction
            // and updates its three columns of the row at once -- no trip through the list in shared memory, and the update's
            // loads and FMAs are independent of the re-integration scan that follows
            const double ex = __shfl_sync(full, px, 0) - pox, ey = __shfl_sync(full, py, 0) - poy;
            if (ex != 0.0 || ey != 0.0) {
                double c11[3], c12[3], c22[3];
                const uint32_t jo = 8u * (uint32_t) ix;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const uint32_t ad = r0l[r] - jo;
                    c11[r] = lds_const(ad); c12[r] = lds_const(ad + r0s); c22[r] = lds_const(ad + 2u * r0s);
                }
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    urx_[r] = urx_[r] + (c11[r] * ex + c12[r] * ey); ury_[r] = ury_[r] + (c12[r] * ex + c22[r] * ey);
                    if (lane + 32 * r == ix) { ddx_[r] += ex; ddy_[r] += ey; }
                }
            }
        }
        if (!convex && active) {
            // re-integrate dp -> ps to the left (:3089-3126) with a warp scan: lanes 0.. take the elements
            // ix-1, ix-2, ...; the chain of adhesion elements follows the new traction, it ends at the
            // first element that does not change, at a clamped element or at the first non-adhesion element
            double rx = __shfl_sync(full, px, 0), ry = __shfl_sync(full, py, 0);
            int jj0 = ix - 1;
            bool done = false, first = true;
            while (!done && jj0 >= 0) {
                const int jj = jj0 - lane;
                int ej, L;
                if (first) { ej = ej0; L = L0; first = false; }
                else {
                    ej = (jj >= 0) ? s.el(jj) : -1;
                    const unsigned notadh = __ballot_sync(full, ej != EL_ADHES);
                    L = notadh ? __ffs(notadh) - 1 : 32;
                }
                double ax = (lane < L) ? s.dpx(jj) : 0.0, ay = (lane < L) ? s.dpy(jj) : 0.0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double tx = __shfl_up_sync(full, ax, o), ty = __shfl_up_sync(full, ay, o);
                    if (lane >= o) { ax += tx; ay += ty; }
                }
                double nx = rx + ax, ny = ry + ay;
                const double pb = (lane < L) ? fmin(s.bnd(jj), 1e20) : 0.0;
                const bool clamp = lane < L && (nx * nx + ny * ny > pb * pb);
                const unsigned cm = __ballot_sync(full, clamp);
                const int Lc = cm ? __ffs(cm) - 1 : L;              // lanes < Lc keep the scanned values
                const bool same = lane < Lc && nx == s.psx(jj) && ny == s.psy(jj);
                const unsigned smk = __ballot_sync(full, same);
                const int Ls = smk ? __ffs(smk) - 1 : 32;
                if (lane < Lc && lane <= Ls) { s.psx(jj) = nx; s.psy(jj) = ny; }
                if (Ls < Lc) { done = true; break; }
                // traction of the element to the right of lane Lc
                const int src = Lc > 0 ? Lc - 1 : 0;
                const double qx = __shfl_sync(full, nx, src), qy = __shfl_sync(full, ny, src);
                const double prx = Lc > 0 ? qx : rx, pry = Lc > 0 ? qy : ry;
                if (Lc < L) {                                       // lane Lc: adhesion element beyond its bound
                    double cxn = 0.0, cyn = 0.0;
                    int stop = 0;
                    if (lane == Lc) {
                        double mx_ = prx + s.dpx(jj), my_ = pry + s.dpy(jj);
                        const double t = pb / sqrt(mx_ * mx_ + my_ * my_);
                        mx_ *= t; my_ *= t;
                        const double ndx = mx_ - prx, ndy = my_ - pry;
                        const int c = s.ictl(0, my);
                        s.chj(c) = jj; s.chx(c) = ndx - s.dpx(jj); s.chy(c) = ndy - s.dpy(jj); s.ictl(0, my) = c + 1;
                        s.dpx(jj) = ndx; s.dpy(jj) = ndy;
                        stop = (mx_ == s.psx(jj) && my_ == s.psy(jj));
                        s.psx(jj) = mx_; s.psy(jj) = my_;
                        cxn = mx_; cyn = my_;
                    }
                    rx = __shfl_sync(full, cxn, Lc); ry = __shfl_sync(full, cyn, Lc);
                    done = __shfl_sync(full, stop, Lc) != 0;
                    jj0 -= Lc + 1;
                    __syncwarp();
                } else if (L < 32) {                                // lane L: the chain's terminator
                    if (lane == L && jj >= 0) {
                        double ndx, ndy;
                        bool upd = true;
                        if (ej >= EL_SLIP) { ndx = s.psx(jj) - prx; ndy = s.psy(jj) - pry; }
                        else if (s.el(jj + 1) >= EL_ADHES) { ndx = -prx; ndy = -pry; }
                        else { upd = false; ndx = 0.0; ndy = 0.0; }
                        if (upd) {
                            const double cx = ndx - s.dpx(jj), cy = ndy - s.dpy(jj);
                            if (cx != 0.0 || cy != 0.0) { const int c = s.ictl(0, my); s.chj(c) = jj; s.chx(c) = cx; s.chy(c) = cy; s.ictl(0, my) = c + 1; }
                            s.dpx(jj) = ndx; s.dpy(jj) = ndy;
                        }
                    }
                    done = true;
                } else {                                            // 32 adhesion elements done, next window
                    rx = __shfl_sync(full, nx, 31); ry = __shfl_sync(full, ny, 31);
                    jj0 -= 32;
                    __syncwarp();
                }
            }
            __syncwarp();
        }
        CB_GS_PROF(const unsigned long long tc = clock64();)
        // in-row rank-1 updates: keep the displacement differences of this row current, accumulate the net
        // change of the row for the other rows
        const int nch = DIRECT ? 0 : s.ictl(0, my);
        for (int c = 0; c < nch; c++) {
            const int jx = s.chj(c);
            const double ex = s.chx(c), ey = s.chy(c);
            if (rr3) {
                double c11[3], c12[3], c22[3];
                const uint32_t jo = 8u * (uint32_t) jx;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const uint32_t ad = r0l[r] - jo;
                    c11[r] = lds_const(ad); c12[r] = lds_const(ad + r0s); c22[r] = lds_const(ad + 2u * r0s);
                }
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    urx_[r] = urx_[r] + (c11[r] * ex + c12[r] * ey); ury_[r] = ury_[r] + (c12[r] * ex + c22[r] * ey);
                    if (lane + 32 * r == jx) { ddx_[r] += ex; ddy_[r] += ey; }
                }
                continue;
            }
            if (lane == 0) { s.ddx(jx) += ex; s.ddy(jx) += ey; }
            for (int base = 0; base < mx; base += 96) {             // 3 elements per lane, loads first; the
                double c11[3], c12[3], c22[3], u0[3], u1[3];        // exterior elements are updated too (unused)
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const int ixp = min(base + lane + 32 * r, mx - 1);
                    T.row0(ixp - jx, c11[r], c12[r], c22[r]); u0[r] = s.urx(ixp); u1[r] = s.ury(ixp);
                }
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const int ixp = base + lane + 32 * r;
                    if (ixp < mx) { s.urx(ixp) = u0[r] + (c11[r] * ex + c12[r] * ey); s.ury(ixp) = u1[r] + (c12[r] * ex + c22[r] * ey); }
                }
            }
        }
        __syncwarp();
        CB_GS_PROF(if (lane == 0) { const unsigned long long td = clock64(); tp[1] += tb - ta; tp[2] += tc - tb; tp[3] += td - tc; tp[5] += nch; })
        }
        if constexpr (DIRECT) __syncthreads();     // the row arrays changed by warp 0 feed the next row sum
    }
    if (rr3) {                                      // the row's U back to shared memory for its owners
#pragma unroll
        for (int r = 0; r < 3; r++) { const int ixp = lane + 32 * r; if (ixp < mx) { s.urx(ixp) = urx_[r]; s.ury(ixp) = ury_[r]; } }
    }
    // compact the net changes of this row for the update of the other rows
    int cnt = 0;
    if (!DIRECT)
    for (int base = 0; base < mx; base += 32) {
        const int jx = base + lane;
        const int rb = base >> 5;
        const double dxr = rb == 0 ? ddx_[0] : (rb == 1 ? ddx_[1] : ddx_[2]), dyr = rb == 0 ? ddy_[0] : (rb == 1 ? ddy_[1] : ddy_[2]);
        const double ex = jx < mx ? (rr3 ? dxr : s.ddx(jx)) : 0.0, ey = jx < mx ? (rr3 ? dyr : s.ddy(jx)) : 0.0;
        const bool nz = (ex != 0.0 || ey != 0.0);
        const unsigned mk = __ballot_sync(full, nz);
        if (nz) { const int pos = cnt + __popc(mk & ((1u << lane) - 1u)); s.lj(par, pos) = jx; s.lx(par, pos) = ex; s.ly(par, pos) = ey; }
        cnt += __popc(mk);
    }
    if (!DIRECT && lane == 0) { s.lcnt(par) = cnt; CB_GS_PROF(tp[6] += cnt;) }
    return dsum;
}

// what the sweeps need from the set-up in stdygs_dev, and their results (identical in all threads)
struct GsSweep {
    const ConvPlan *P;
    const SteadyArgs *a;
    SteadySmem s;
    SteadyTab T;
    int *el;
    double *ps, *ss, *red;
    const double *xp;
    int ncon, nsp;
    double facnel;
    int itgs;
    double dif, dif1;
};

// The sweeps themselves, in three roles with the same barrier sequence:
//   WALKER (warp 0 of the register form): walks the rows (gs_walk_row), owns no contact elements -- an instantiation of its own so
//          that the dependent instruction chain of the walk has the whole register file (inside one function with the owners the
//          chain made ~110 local-memory accesses per element step: the row's U and the counters did not fit beside 5 KMAX registers
//          of register-resident U that warp 0 never uses);
//   owner  (warps 1.., register form): KMAX contact elements per thread with their U in registers, rank-1 updates by apply_list;
//   DIRECT (whole CTA): no U, block-wide row sum before every element step.
template <int KMAX, bool DIRECT, bool WALKER>
__device__ __noinline__ void gs_sweeps(GsSweep &g)
{
    constexpr bool OWN = !DIRECT && !WALKER;
    const ConvPlan &P = *g.P;
    const SteadyArgs &a = *g.a;
    const SteadySmem s = g.s;
    const SteadyTab T = g.T;
    int *el = g.el;
    double *ps = g.ps, *ss = g.ss, *red = g.red;
    const double *xp = g.xp;
    const int ncon = g.ncon, nsp = g.nsp;
    const double facnel = g.facnel;
    const int n = P.npot, mx = P.mx, my = P.my, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const unsigned full = 0xffffffffu;
    double *psx = ps, *psy = ps + n, *psn = ps + 2 * (size_t) n;
    const bool convex = a.convex != 0;
    (void) lane; (void) full; (void) nsp;
    // registers: my contact elements k = (tid - 32) + m (nt - 32).  Warp 0 owns none: it walks the rows while the other warps
    // apply the net change of the previous row to their elements (pipelined update below)
    double Ux[KMAX], Uy[KMAX];
    int ixy[KMAX];
#pragma unroll
    for (int m = 0; m < KMAX; m++) {
        const int k = (tid - 32) + m * (nt - 32);
        Ux[m] = 0.0; Uy[m] = 0.0; ixy[m] = -1;
        if (OWN && tid >= 32 && k < ncon) {
            const int ii = a.iel[k], iy = ii / mx;
            ixy[m] = (ii - iy * mx) | (iy << 16);
            Ux[m] = a.ug[ii]; Uy[m] = a.ug[n + ii];
        }
    }
    double q00, q01, q11;
    T.row0(0, q00, q01, q11);
    const uint32_t mg_mx = div_magic((uint32_t) mx);

    int itgs = 0;
    double dif = 2.0, difid = 1.0, dif1 = 0.0;
#ifdef CB_STEADY_PROF
    unsigned long long tpw_[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, *tpw = tpw_;
#else
    unsigned long long *tpw = nullptr;
#endif
    unsigned long long nstep = 0;                              // element steps of this call (thread 0)
    // leading-edge elements: 2x2 matrix of cs (coefs instead of coefsv, :2665-2669)
    const bool ledge = convex && a.ledge != 0;
    const size_t oc = (size_t) a.cmy * (2 * a.cmx) + a.cmx;
    const double l00 = ledge ? a.cs11[oc] * a.ga_inv : 0.0, l01 = ledge ? a.cs12[oc] * a.ga_inv : 0.0,
                 l11 = ledge ? a.cs22[oc] * a.ga_inv : 0.0;
    // pipelined update of the other rows (register form): the compacted net-change list of a finished row lives in one of two
    // shared-memory buffers; apply_list adds its effect to this thread's elements in the rows selected by `sel`
    int par = 0, pend_par = 0, pend_row = -1;
    auto apply_list = [&](int b, int src, auto sel) {
        const int ncl = s.lcnt(b);
        if (ncl == 0) return;
        if (s.hasq && a.sym) {                                 // quadrant table in shared memory: the fast path
#pragma unroll
            for (int m = 0; m < KMAX; m++) {
                const int iym = ixy[m] >> 16, ixm = ixy[m] & 0xffff;
                if (ixy[m] >= 0 && sel(iym)) {
                    const int dy = iym - src;
                    const double *r11 = s.q() + abs(dy) * mx, *r12 = r11 + n, *r22 = r12 + n;
                    const bool ny = dy < 0;
                    double ux0 = 0.0, uy0 = 0.0, ux1 = 0.0, uy1 = 0.0;
                    int c = 0;
                    for (; c + 1 < ncl; c += 2) {
                        const int d0 = ixm - s.lj(b, c), d1 = ixm - s.lj(b, c + 1);
                        const int a0 = abs(d0), a1 = abs(d1);
                        const double e0x = s.lx(b, c), e0y = s.ly(b, c), e1x = s.lx(b, c + 1), e1y = s.ly(b, c + 1);
                        const double g0 = r11[a0], h0 = ((d0 < 0) != ny) ? -r12[a0] : r12[a0], k0_ = r22[a0];
                        const double g1 = r11[a1], h1 = ((d1 < 0) != ny) ? -r12[a1] : r12[a1], k1_ = r22[a1];
                        ux0 += g0 * e0x + h0 * e0y; uy0 += h0 * e0x + k0_ * e0y;
                        ux1 += g1 * e1x + h1 * e1y; uy1 += h1 * e1x + k1_ * e1y;
                    }
                    if (c < ncl) {
                        const int d0 = ixm - s.lj(b, c), a0 = abs(d0);
                        const double e0x = s.lx(b, c), e0y = s.ly(b, c);
                        const double g0 = r11[a0], h0 = ((d0 < 0) != ny) ? -r12[a0] : r12[a0], k0_ = r22[a0];
                        ux0 += g0 * e0x + h0 * e0y; uy0 += h0 * e0x + k0_ * e0y;
                    }
                    Ux[m] += ux0 + ux1; Uy[m] += uy0 + uy1;
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < KMAX; m++) {
                const int iym = ixy[m] >> 16, ixm = ixy[m] & 0xffff;
                if (ixy[m] >= 0 && sel(iym)) {
                    double ux = 0.0, uy = 0.0;
                    for (int c = 0; c < ncl; c++) {
                        double c11, c12, c22;
                        T.get(ixm - s.lj(b, c), iym - src, c11, c12, c22);
                        const double ex = s.lx(b, c), ey = s.ly(b, c);
                        ux += c11 * ex + c12 * ey;
                        uy += c12 * ex + c22 * ey;
                    }
                    Ux[m] += ux; Uy[m] += uy;
                }
            }
        }
    };
    GsWalk w;                                                  // what the walking warp needs (gs_walk_row)
    w.s = s; w.T = T; w.a = &a; w.xp = xp; w.psx = psx; w.psy = psy; w.red = red;
    w.q00 = q00; w.q01 = q01; w.q11 = q11; w.l00 = l00; w.l01 = l01; w.l11 = l11;
    w.n = n; w.mx = mx; w.my = my; w.ncon = ncon; w.nsp = nsp; w.mg_mx = mg_mx; w.convex = convex ? 1 : 0; w.ledge = ledge ? 1 : 0;
    while (dif >= difid && itgs < a.maxgs) {
        itgs++;
        nstep += ncon;
        // work accounting: a sweep is 2 ncon row sums over (ncon + 2 my) 2 columns in the reference
        if (threadIdx.x == 0) atomicAdd(&g_work[3], (unsigned long long) ncon * (unsigned long long) (ncon + 2 * P.my));
        double dsum = 0.0;                                     // lane 0 of warp 0 only
        if (ledge) {
            // subnd (m_leadedge.f90:336-394) at the start of every sweep (:2583): displacement difference of the current
            // tractions (all three directions, cs) at the first exterior element behind each C -> E transition
            // (facdx = 1); one warp per transition, row sum over the compact contact list with a fixed shuffle tree
            const int wid = tid >> 5, nw = nt >> 5;
            for (int k = wid; k < ncon; k += nw) {
                const int ii = a.iel[k], iy = ii / mx, ix = ii - iy * mx;
                if (ix != mx - 1 && el[ii + 1] >= 1) continue;
                double ux = 0.0, uy = 0.0;
                if (ix + 3 <= mx) {
                    for (int kk = lane; kk < ncon; kk += 32) {
                        const int jj = a.iel[kk], jy = jj / mx, jx = jj - jy * mx;
                        const size_t o = (size_t) (iy - jy + a.cmy) * (2 * a.cmx) + (ix + 1 - jx) + a.cmx;
                        const double px = psx[jj], py = psy[jj], c12 = a.cs12[o];
                        ux += a.cs11[o] * px + c12 * py; uy += c12 * px + a.cs22[o] * py;
                        if (a.cs13) { const double pn = psn[jj]; ux += a.cs13[o] * pn; uy += a.cs23[o] * pn; }
                    }
                    for (int o = 16; o > 0; o >>= 1) { ux += __shfl_xor_sync(full, ux, o); uy += __shfl_xor_sync(full, uy, o); }
                    ux *= a.ga_inv; uy *= a.ga_inv;
                }
                if (lane == 0) { a.ub[ii] = ux; a.ub[n + ii] = uy; }
            }
            __syncthreads();
        }
        for (int iy = 0; iy < my; iy++) {
            const int k0 = s.rowk(iy), k1 = s.rowk(iy + 1);
            if (k1 == k0) continue;
            for (int jx = tid; jx < mx; jx += nt) {            // stage the row in shared memory
                const int ii = iy * mx + jx;
                s.psx(jx) = psx[ii]; s.psy(jx) = psy[ii];
                if (!convex) { s.dpx(jx) = a.dp[ii]; s.dpy(jx) = a.dp[n + ii]; }
                s.bnd(jx) = a.mu * psn[ii]; s.wsx(jx) = a.ws[ii]; s.wsy(jx) = a.ws[n + ii];
                s.ssx(jx) = ss[ii]; s.ssy(jx) = ss[n + ii]; s.el(jx) = el[ii];
                s.ddx(jx) = 0.0; s.ddy(jx) = 0.0;
            }
            for (int k = k0 + tid; k < k1; k += nt) s.cix(k - k0) = a.iel[k] - iy * mx;
            if constexpr (OWN) {
#pragma unroll
                for (int m = 0; m < KMAX; m++)
                    if (ixy[m] >= 0 && (ixy[m] >> 16) == iy) { s.urx(ixy[m] & 0xffff) = Ux[m]; s.ury(ixy[m] & 0xffff) = Uy[m]; }
            }
            __syncthreads();

            if constexpr (!OWN) {                              // ---- warp 0 (whole CTA: direct form): the Gauss-Seidel steps of this row ----
                dsum = gs_walk_row<DIRECT>(w, iy, k0, k1, itgs, par, dsum, tpw);
            } else if (pend_row >= 0) {
                // ---- warps 1.. meanwhile: the net change of the PREVIOUS row goes to the elements of all other rows.  This pass is
                // shared-memory-bandwidth bound (three table entries per pair), the chain of warp 0 is latency bound: they overlap.
                // (This row received it before it was staged; the previous row keeps its own U in shared memory.)
                apply_list(pend_par, pend_row, [&](int r) { return r != pend_row && r != iy; });
            }
            __syncthreads();

            // ---- whole CTA: the net change of row iy goes to the NEXT row with contact only (it is staged next); all other rows
            // receive it while warp 0 walks that next row ----
            CB_GS_PROF(const unsigned long long te = clock64();)
            if constexpr (!DIRECT) {
                int nx = iy;
                do { nx = (nx + 1 == my) ? 0 : nx + 1; } while (nx != iy && s.rowk(nx + 1) == s.rowk(nx));
                if constexpr (OWN) { if (nx != iy) apply_list(par, iy, [&](int r) { return r == nx; }); }
                pend_row = iy; pend_par = par; par ^= 1;
            }
            if constexpr (OWN) {
#pragma unroll
                for (int m = 0; m < KMAX; m++)
                    if (ixy[m] >= 0 && (ixy[m] >> 16) == iy) { Ux[m] = s.urx(ixy[m] & 0xffff); Uy[m] = s.ury(ixy[m] & 0xffff); }
            }
            for (int jx = tid; jx < mx; jx += nt) {            // write the row back
                const int ii = iy * mx + jx;
                psx[ii] = s.psx(jx); psy[ii] = s.psy(jx);
                if (!convex) { a.dp[ii] = s.dpx(jx); a.dp[n + ii] = s.dpy(jx); }
                ss[ii] = s.ssx(jx); ss[n + ii] = s.ssy(jx); el[ii] = s.el(jx);
            }
            CB_GS_PROF(if (tid == 0) tpw[4] += clock64() - te;)
            __syncthreads();
        }
        if (tid == 0) s.scal(2) = dsum;
        double p2[1] = { 0.0 };
        for (int i = tid; i < n; i += nt) p2[0] += psx[i] * psx[i] + psy[i] * psy[i];
        block_sum<1>(p2, red);
        dif = sqrt(s.scal(2) / (2.0 * ncon));
        difid = a.eps * fmax(1e-6, facnel * sqrt(p2[0] / (2.0 * n)));
        if (itgs == 1) dif1 = dif;
        __syncthreads();
    }
    g.itgs = itgs; g.dif = dif; g.dif1 = dif1;
    if (tid == 0) {
        atomicAdd(&g_steady_prof[0], nstep); atomicAdd(&g_steady_prof[4], 1ull);
        CB_GS_PROF(atomicAdd(&g_steady_prof[1], tpw[1]); atomicAdd(&g_steady_prof[2], tpw[2]); atomicAdd(&g_steady_prof[3], tpw[3]);
                   atomicAdd(&g_steady_prof[5], tpw[4]); atomicAdd(&g_steady_prof[6], tpw[5]); atomicAdd(&g_steady_prof[7], tpw[6]);)
    }
}

// Gauss-Seidel sweeps of stdygs (convex = 0) or cnvxgs (convex = 1): returns info (0 ok, 1 maxgs reached,
// 2 stagnation, 3 divergence).  All threads of the CTA must call.
//
// Row-blocked organisation: the elements of one grid row are processed by warp 0 alone -- lane 0 runs the scalar
// per-element solve, all 32 lanes keep the row's own displacement differences up to date and re-integrate the row
// with a warp scan -- while the net change of the row is applied to the register-resident U of all other rows once
// per row by the whole CTA.
//
// DIRECT = true: the form for contact areas that do not fit the register-resident U (more than 22 x 352 elements, or a grid on
// the whole-GPU path).  No U is kept: before every element step the whole CTA evaluates the reference's row sum
// U_i = (1/G) sum_j A(i - j) xp_j over the compact contact list (gf3_AijPj, m_aijpj.f90:99-254; the current row from shared
// memory, the other rows from global memory / L2, coefficients from the spatial blocks in L2), then warp 0 performs the
// element step exactly as in the register form.  No rank-1 updates, no FFT products, no coefficient table in shared
// memory: O(ncon) work per element like the reference, any grid size.  `sbase`: shared memory for the row arrays
// (steady_fixed_bytes), used instead of the plan's layout.
template <int KMAX, bool DIRECT = false>
__device__ __noinline__ int stdygs_dev(const ConvPlan &P, const Smem &sm, const SteadyArgs &a, int *el, double *ps, double *ss,
                          int ncon, int &itgs_out, double &err_out, int &nprod, unsigned char *sbase = nullptr)
{
    const int n = P.npot, mx = P.mx, my = P.my, tid = threadIdx.x, nt = blockDim.x;
    double *psx = ps, *psy = ps + n;
    double *red = sm.red;

    // SteadyGS: traction differences along the rolling direction (x ascending = towards the leading edge), :2900-2915;
    // ConvexGS works on the tractions themselves
    const bool convex = a.convex != 0;
    const double *xp = convex ? ps : a.dp;
    if (!convex) {
        for (int i = tid; i < n; i += nt) {
            const int ix = i % mx;
            a.dp[i] = (ix != mx - 1) ? psx[i] - psx[i + 1] : psx[i];
            a.dp[n + i] = (ix != mx - 1) ? psy[i] - psy[i + 1] : psy[i];
        }
    }
    __syncthreads();
    const double facnel = (double) sqrtf(__fdiv_rn((float) n, (float) ncon));

    // U = A_tt xp on the contact area by four FFT products (fresh at every solver call)
    if constexpr (!DIRECT)
    for (int ik = 0; ik < 2; ik++) {
        bool ladd = false;
        for (int jk = 0; jk < 2; jk++) {
            if (a.chatA[ik][jk] == nullptr) continue;
            conv_dev(P, sm, xp + (size_t) jk * n, a.chatA[ik][jk], a.ug + (size_t) ik * n, el, 1, ladd ? 1 : 0);
            ladd = true; nprod++;
        }
    }

    // shared memory is ours now (S and W regions of the FFT layout)
    SteadySmem s;
    if constexpr (DIRECT) {
        if (sbase == nullptr) {                                     // one-CTA path: the place steady_carve would choose
            sbase = reinterpret_cast<unsigned char *>(sm.S);
            if (steady_fixed_bytes(P.mx, P.my) > (size_t) P.off_twx) sbase += P.smem_bytes;
            conv_tables_invalidate(sm);
        }
        steady_offsets(sbase, 0, mx, my, s);
    } else {
        steady_carve(P, reinterpret_cast<unsigned char *>(sm.S), a.sym, s);
        conv_tables_invalidate(sm);                                 // the sweep arrays overwrite the product's window
    }
    if (s.hasq) {
        if (a.sym) {
            for (int i = tid; i < n; i += nt) {
                const int ay = i / mx, ax = i - ay * mx;
                const size_t o = (size_t) (ay + a.cmy) * (2 * a.cmx) + ax + a.cmx;
                s.q()[i] = a.cf11[o] * a.ga_inv; s.q()[n + i] = a.cf12[o] * a.ga_inv; s.q()[2 * n + i] = a.cf22[o] * a.ga_inv;
            }
        } else {
            for (int i = tid; i < 2 * n; i += nt) {
                const int ay = i / (2 * mx), dx = i - ay * 2 * mx - mx;
                const size_t o = (size_t) (ay + a.cmy) * (2 * a.cmx) + dx + a.cmx;
                s.q()[i] = a.cf11[o] * a.ga_inv; s.q()[2 * n + i] = a.cf12[o] * a.ga_inv; s.q()[4 * n + i] = a.cf22[o] * a.ga_inv;
            }
        }
    }
    for (int i = tid; i < 2 * mx; i += nt) {
        const size_t o = (size_t) a.cmy * (2 * a.cmx) + (i - mx) + a.cmx;
        s.r0()[i] = a.cf11[o] * a.ga_inv; s.r0()[2 * mx + i] = a.cf12[o] * a.ga_inv; s.r0()[4 * mx + i] = a.cf22[o] * a.ga_inv;
    }
    SteadyTab T;
    T.oq = s.hasq ? s.oq : 0xffffffffu; T.or0 = s.or0; T.cf11 = a.cf11; T.cf12 = a.cf12; T.cf22 = a.cf22; T.n = n; T.mx = mx; T.cmx = a.cmx; T.cmy = a.cmy;
    T.sym = a.sym; T.ga_inv = a.ga_inv;
    // compact list of contact elements in sweep order + row offsets
    for (int iy = tid; iy < my; iy += nt) {
        int cnt = 0;
        for (int ix = 0; ix < mx; ix++) cnt += (el[iy * mx + ix] >= 1);
        s.rowk(iy + 1) = cnt;
    }
    __syncthreads();
    if (tid == 0) { s.rowk(0) = 0; for (int iy = 0; iy < my; iy++) s.rowk(iy + 1) += s.rowk(iy); }
    __syncthreads();
    for (int iy = tid; iy < my; iy += nt) {
        int k = s.rowk(iy);
        for (int ix = 0; ix < mx; ix++) if (el[iy * mx + ix] >= 1) a.iel[k++] = iy * mx + ix;
    }
    __syncthreads();

    int nsp = 0;
    if constexpr (DIRECT) {                                       // column ranges of the row sums (contact area is fixed in TANG)
        int *spk = a.isp, *spl = a.isp + my + 2;
        for (int iy = tid; iy < my; iy += nt) {
            int first = mx, last = -1;
            for (int ix = 0; ix < mx; ix++) if (el[iy * mx + ix] >= 1) { if (first == mx) first = ix; last = ix; }
            spk[iy + 1] = last < 0 ? 0 : min(mx - 1, last + 1) - max(0, first - 1) + 1;
        }
        __syncthreads();
        if (tid == 0) { spk[0] = 0; for (int iy = 0; iy < my; iy++) spk[iy + 1] += spk[iy]; }
        __syncthreads();
        for (int iy = tid; iy < my; iy += nt) {
            const int cnt = spk[iy + 1] - spk[iy];
            if (cnt > 0) {
                int first = 0;
                while (el[iy * mx + first] < 1) first++;
                const int j0 = max(0, first - 1);
                for (int q = 0; q < cnt; q++) spl[spk[iy] + q] = iy * mx + j0 + q;
            }
        }
        __syncthreads();
        nsp = spk[my];
    }
    GsSweep g;
    g.P = &P; g.a = &a; g.s = s; g.T = T; g.el = el; g.ps = ps; g.ss = ss; g.red = red; g.xp = xp; g.ncon = ncon; g.nsp = nsp;
    g.facnel = facnel; g.itgs = 0; g.dif = 2.0; g.dif1 = 0.0;
    if constexpr (DIRECT) gs_sweeps<1, true, false>(g);
    else if (tid < 32) gs_sweeps<1, false, true>(g);           // warp 0 walks the rows,
    else gs_sweeps<KMAX, false, false>(g);                     // warps 1.. own the contact elements
    const int itgs = g.itgs;
    const double dif = g.dif, dif1 = g.dif1;
    double conv = 1.0;
    if (dif * dif1 != 0.0 && itgs > 1) conv = exp(log(dif / dif1) / (itgs - 1));
    int info = 0;
    if (itgs >= a.maxgs) info = 1;
    if (itgs >= a.maxgs && conv > 0.997) info = 2;
    if (itgs >= a.maxgs && conv > 1.0) info = 3;
    itgs_out = itgs; err_out = dif;
    __syncthreads();
    return info;
}

}  // namespace cb200
