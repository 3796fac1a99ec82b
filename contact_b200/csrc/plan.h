// plan.h -- host-side planning for the single-CTA FFT convolution (sizes, radices, tables, shared-memory layout).
// opt_fft_size restates the reference's size chooser (/root/reference/src/m_aijpj.f90:1022-1119) because the padded
// size is part of the reference's semantics (it fixes which transform lengths occur); everything else is new.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "fftconv2.cuh"

namespace cb200 {

static const int kSmemMax = 232448;      // 227 KB opt-in maximum per CTA on sm_100
static const int kPlanWarps = 12;        // warps per CTA the plans are laid out for (CB_THREADS / 32)

inline int opt_fft_size(int n)
{
    // factor-trading patterns (delta exponents of 2,3,5,7 | numerator, denominator | largest prime used)
    static const int pat[8][7] = {
        { -2,  1,  0,  0,   4,  3, 3 }, { -5,  3,  0,  0,  32, 27, 3 }, { -1, -1,  1,  0,   6,  5, 5 },
        { -4,  1,  1,  0,  16, 15, 5 }, {  0, -3,  2,  0,  27, 25, 5 }, { -3,  0,  0,  1,   8,  7, 7 },
        {  1, -1, -1,  1,  15, 14, 7 }, { -1,  2, -1,  0,  10,  9, 5 } };
    int e[4] = { (int) std::ceil(std::log(1.0 * n) / std::log(2.0)), 0, 0, 0 };
    long prod = 1L << e[0];
    bool changed = true;
    while (changed) {
        changed = false;
        for (int ip = 0; ip < 8; ip++) {
            const int *k = pat[ip];
            for (;;) {
                bool ok = true;
                for (int f = 0; f < 4; f++) ok = ok && (e[f] + k[f] >= 0);
                if (!ok || (long) k[5] * prod < (long) k[4] * n) break;
                for (int f = 0; f < 4; f++) e[f] += k[f];
                prod = k[5] * prod / k[4];
                changed = true;
            }
        }
    }
    return (int) prod;
}

// split L = 2^a 3^b 5^c 7^d into register-sized radices; returns false if another prime occurs
inline bool choose_radices(int L, int *nst, int *rad)
{
    int a = 0, b = 0, c = 0, d = 0, n = L;
    while (n % 2 == 0) { a++; n /= 2; }
    while (n % 3 == 0) { b++; n /= 3; }
    while (n % 5 == 0) { c++; n /= 5; }
    while (n % 7 == 0) { d++; n /= 7; }
    if (n != 1) return false;
    // register-sized radices {16, 12, 9, 8, 7, 6, 5, 4, 3, 2}; fewest stages first (every stage is one full pass
    // over shared memory), largest radix first (the last forward stage is fused with the multiply and the
    // first inverse stage, so it should be a mid-size radix)
    int k = 0;
#ifdef CB_MAXRADIX8
    // experiment: radices <= 9 only (smaller register footprint per butterfly => more warps per SM)
    while (a >= 3) { rad[k++] = 8; a -= 3; }
#endif
    while (a >= 2 && b >= 1 && (a + b > 3 || c + d > 0 || a == 2)) { rad[k++] = 12; a -= 2; b -= 1; if (a < 2 || b < 1) break; }
    while (a >= 4) { rad[k++] = 16; a -= 4; }
    if (a == 3) { rad[k++] = 8; a = 0; }
    while (b >= 2) { rad[k++] = 9; b -= 2; }
    if (a >= 1 && b >= 1) { rad[k++] = 6; a -= 1; b -= 1; }
    if (a == 2) { rad[k++] = 4; a = 0; }
    if (a == 1) { rad[k++] = 2; a = 0; }
    if (b == 1) { rad[k++] = 3; b = 0; }
    while (d-- > 0) rad[k++] = 7;
    while (c-- > 0) rad[k++] = 5;
    // sort descending
    for (int i = 0; i < k; i++) for (int j = i + 1; j < k; j++) if (rad[j] > rad[i]) { int t = rad[i]; rad[i] = rad[j]; rad[j] = t; }
    if (k > CB_MAXSTAGE) return false;
    *nst = k;
    return true;
}

struct HostPlan {
    ConvPlan p;
    std::vector<cd> tab2;          // twiddle tables of the warp-resident product (Conv2Plan::tab)
    std::vector<cd> twx, twy;
    std::vector<unsigned short> posx;
    bool fits;           // whole product fits one CTA's shared memory
};

// position of frequency k after in-place DIF with radices r[0..ns-1]
inline void dif_positions(int L, int ns, const int *r, std::vector<unsigned short> &pos)
{
    pos.resize(L > 0 ? L : 1);
    for (int k = 0; k < L; k++) {
        int rem = k, span = L, p = 0;
        for (int s = 0; s < ns; s++) {
            const int q = rem % r[s];
            rem /= r[s];
            span /= r[s];
            p += q * span;
        }
        pos[k] = (unsigned short) p;
    }
}

// ---- plan of the warp-resident product (fftconv2.cuh) ----
inline int c2_tab_end(const ConvPlan &P) { return P.c2.ok ? P.c2.off_tab + (P.c2.tab_len + kPlanWarps) * 16 : 0; }

// cost model used to pick radices and group sizes: FP64 instructions of a radix-R butterfly incl. its twiddles
inline double c2_bfly_cost(int R) { return R <= 1 ? 0.0 : 3.3 * R * std::log2((double) R) + 4.0 * R; }

inline int c2_next_id() { static int id = 0; return ++id; }

// fills hp.p.c2 and hp.tab2; returns false (c2.ok = 0) when the size is not served (the block-wide path runs then).
// `end` returns the first byte behind the layout.
inline bool make_plan2(HostPlan &hp, int smem_limit, long *end, int tab_end = 0)
{
    ConvPlan &P = hp.p;
    Conv2Plan &c = P.c2;
    std::memset(&c, 0, sizeof(c));
    *end = 0;
    if (!hp.fits || P.Lx < 4 || P.Ly < 8) return false;
    // radices: rows Lx = Ax * Bx, columns Ly = Ay * By (Ay even: the pruned column stages need it)
    int bestx = 1 << 30, besty = 1 << 30;
    for (int A = 2; A <= 18; A++) {
        if (P.Lx % A == 0 && c2_row_class(A, P.Lx / A)) {
            const int B = P.Lx / A;
            const int sc = ((A & 1) ? 100 : 0) + std::abs(A - B) + (B > 8 ? 10 : 0) + (B == 1 ? 50 : 0);
            if (sc < bestx) { bestx = sc; c.Ax = A; c.Bx = B; }
        }
        if (P.Ly % A == 0 && (A & 1) == 0 && c2_col_class(A, P.Ly / A)) {
            const int B = P.Ly / A;
            const int sc = std::abs(A - B) + (B > 12 ? 10 : 0) + (B > A ? 5 : 0) + (B == 1 ? 50 : 0);
            if (sc < besty) { besty = sc; c.Ay = A; c.By = B; }
        }
    }
    if (c.Ax == 0 || c.Ay == 0) return false;
    c.blkx = c.Bx | 1; c.blky = c.By | 1;
    c.nux = (c.Ax & 1) ? (c.Ax + 1) / 2 : c.Ax / 2;
    // tables
    c.o_t1x = 0;
    c.o_tsx = c.o_t1x + (c.Ax - 1) * c.Bx;
    c.o_tmx = c.o_tsx + c.nux * c.Bx;
    c.o_tay = c.o_tmx + c.Bx;
    c.o_kr = c.o_tay + (c.Ay - 1) * c.By;                       // stage constants behind the twiddles (20 words = 5 elements each)
    c.o_kc = c.o_kr + CB2_KWORDS / 4;
    c.tab_len = c.o_kc + CB2_KWORDS / 4;
    const long bytesT = (long) (c.tab_len + kPlanWarps) * 16;       // + one partial sum per warp (fused output reductions)
    // group sizes: minimise the estimated FP64 issue time of one product over the warps that get a slot (the paddings
    // below may still take a slot away on the largest grids; they are chosen afterwards)
    const double cA = c2_bfly_cost(c.Ay), cM = 2.0 * c2_bfly_cost(c.By) + 4.0 * c.By;
    const double cR1 = c2_bfly_cost(c.Ax), cR2 = 2.0 * c2_bfly_cost(c.Bx) + 14.0 * c.Bx;
    {
        const long avail = (long) smem_limit - 1024 - (long) P.Fx * P.my * 16 - bytesT;
        double best = 1e300;
        for (int G = 1; G <= 8; G++)
            for (int RG = 1; RG <= 8 && G != 6; RG *= 2) {       // G = 6: no stride keeps 8 lanes on 8 bank groups                  // rows per group: a power of two (lane order of the split stage)
                const long slot = std::max((long) RG * c.Ax * c.blkx, (long) std::max(G, 2) * c.Ay * c.blky);
                long ns = avail / (slot * 16);
                if (ns > kPlanWarps) ns = kPlanWarps;
                if (ns < 1) continue;
                const int ngrp = (P.Fx + G - 1) / G, nrg = (P.my + RG - 1) / RG;
                auto rounds = [](int items) { return (double) ((items + 31) / 32); };
                const double tc = std::ceil((double) ngrp / ns) * (2.0 * rounds(G * c.By) * cA + rounds(G * c.Ay) * cM);
                const double tr = std::ceil((double) nrg / ns) * 2.0 * (rounds(RG * c.Bx) * cR1 + rounds(RG * c.nux) * cR2);
                if (tc + tr < best * (1.0 - 1e-9)) { best = tc + tr; c.G = G; c.RG = RG; }
            }
        if (c.G == 0) { std::memset(&c, 0, sizeof(c)); return false; }
    }
    // paddings of the S row stride and of the slot strides: the 16-byte accesses of a quarter warp should fall into 8
    // distinct bank groups.  Every candidate is scored with the wavefronts of the access patterns of all stages (lane ->
    // element address, exactly as the stage functions of fftconv2.cuh compute them), under the shared-memory budget.
    {
        auto wf = [](const std::vector<long> &addr) {            // wavefronts of one warp-wide 16-byte access
            long tot = 0;
            for (size_t q0 = 0; q0 < addr.size(); q0 += 8) {
                int cnt[8] = { 0 }; long seen[8][8];
                for (size_t l = q0; l < q0 + 8 && l < addr.size(); l++) {
                    const int b = (int) (((addr[l] % 8) + 8) % 8); bool dup = false;
                    for (int z = 0; z < cnt[b]; z++) dup = dup || seen[b][z] == addr[l];
                    if (!dup) seen[b][cnt[b]++] = addr[l];
                }
                int m = 0; for (int b = 0; b < 8; b++) m = std::max(m, cnt[b]);
                tot += m;
            }
            return tot;
        };
        const int nlc = std::min(32, c.G * c.By), nlm = std::min(32, c.G * c.Ay);
        const int nl1 = std::min(32, c.RG * c.Bx), nl2 = std::min(32, c.RG * c.nux);
        auto ka = [&](int u) { return u; };
        auto kb = [&](int u) { return u == 0 ? ((c.Ax & 1) ? 0 : c.Ax / 2) : c.Ax - u; };
        double best = 1e300;
        int bSY = 0, brow = 0, bcol = 0, bns = 0;
        for (int ps = 0; ps < 8; ps++)
            for (int pr = 0; pr < 8; pr++)
                for (int pc = 0; pc < 8; pc++) {
                    const long SY = P.my + ps, rowlen = (long) c.Ax * c.blkx + pr, collen = (long) c.Ay * c.blky + pc;
                    const long slot = std::max((long) c.RG * rowlen, (long) std::max(c.G, 2) * collen);
                    const long avail = (long) smem_limit - 1024 - (long) P.Fx * SY * 16 - bytesT;
                    long ns = avail / (slot * 16);
                    if (ns > kPlanWarps) ns = kPlanWarps;
                    if (ns < 1) continue;
                    std::vector<long> a;
                    double cost = 0.0;
                    // columns, stages A / C: lane -> (cc = L % G, j = L / G)
                    a.assign(nlc, 0); for (int L = 0; L < nlc; L++) a[L] = (L % c.G) * SY + L / c.G;
                    cost += (double) c.Ay * wf(a);                                   // S: Ay/2 loads + Ay/2 stores
                    a.assign(nlc, 0); for (int L = 0; L < nlc; L++) a[L] = (L % c.G) * collen + L / c.G;
                    cost += 2.0 * c.Ay * wf(a);                                      // slot: Ay stores + Ay loads
                    // columns, stage M: lane -> (cc = L % G, k1 = L / G)
                    a.assign(nlm, 0); for (int L = 0; L < nlm; L++) a[L] = (L % c.G) * collen + (long) (L / c.G) * c.blky;
                    cost += 2.0 * c.By * wf(a) * (double) nlm / std::max(1, nlc);    // per column the M items outnumber the A items
                    cost *= (double) P.Fx / c.G;
                    double rc = 0.0;
                    // rows, first / last stage: lane -> (j = L % Bx, r = L / Bx)
                    a.assign(nl1, 0); for (int L = 0; L < nl1; L++) a[L] = (L % c.Bx) + (L / c.Bx) * rowlen;
                    rc += 2.0 * c.Ax * wf(a);
                    // rows, split / merge stage: lane -> (r = L % RG, u = L / RG), blocks ka(u) and kb(u)
                    double w2 = 0.0;
                    a.assign(nl2, 0); for (int L = 0; L < nl2; L++) a[L] = (L % c.RG) * rowlen + (long) ka(L / c.RG) * c.blkx;
                    w2 += wf(a);
                    a.assign(nl2, 0); for (int L = 0; L < nl2; L++) a[L] = (L % c.RG) * rowlen + (long) kb(L / c.RG) * c.blkx;
                    w2 += wf(a);
                    rc += 2.0 * c.Bx * w2;                                           // slot: Bx loads per block (fwd) + Bx stores (inv)
                    double w3 = 0.0;
                    a.assign(nl2, 0); for (int L = 0; L < nl2; L++) a[L] = (long) ka(L / c.RG) * SY + L % c.RG;
                    w3 += wf(a);
                    a.assign(nl2, 0); for (int L = 0; L < nl2; L++) a[L] = (long) kb(L / c.RG) * SY + L % c.RG;
                    w3 += wf(a);
                    rc += 2.0 * c.Bx * w3;                                           // S: Bx stores per block (fwd) + Bx loads (inv)
                    rc *= (double) P.my / c.RG;
                    cost = (cost + rc) * (double) kPlanWarps / (double) ns;          // fewer slots: fewer warps at work
                    cost *= 1.0 + 1e-4 * (ps + pr + pc);                             // ties: least padding
                    if (cost < best) { best = cost; bSY = (int) SY; brow = (int) rowlen; bcol = (int) collen; bns = (int) ns; }
                }
        if (bns < 4) { std::memset(&c, 0, sizeof(c)); return false; }
        c.SY = bSY; c.rowlen = brow; c.collen = bcol; c.nslot = bns;
        c.slot_len = std::max(c.RG * c.rowlen, std::max(c.G, 2) * c.collen);      // two columns at least: scratch of the packed column
    }
    // S holds the columns kx = 0 .. Fx-1: column 0 is the packed pair (kx = 0, kx = Fx), both real after the row split
    const long bytesS = (long) P.Fx * c.SY * 16;
    c.ngrp = (P.Fx + c.G - 1) / c.G;
    c.chat_len = ((P.Fx + 1 + c.G - 1) / c.G) * c.G * P.Ly;       // coefficient columns 0 .. Fx (column Fx holds Q)
    c.off_S = 0;
    c.off_W = (int) bytesS;
    c.off_tab = c.off_W + c.nslot * c.slot_len * 16;
    if (tab_end > 0) {
        // plan of a ladder level: its tables end where those of the full-grid plan end, so that the window below the tables
        // of ALL plans of a launch is free between products (staged vector passes, norm_solver.cuh)
        if (tab_end - (int) bytesT < c.off_tab) { std::memset(&c, 0, sizeof(c)); return false; }
        c.off_tab = tab_end - (int) bytesT;
    }
    *end = (long) c.off_tab + bytesT;
    c.mg_Bx = div_magic((uint32_t) c.Bx); c.mg_nux = div_magic((uint32_t) c.nux);
    c.mg_By = div_magic((uint32_t) c.By); c.mg_Ay = div_magic((uint32_t) c.Ay);
    c.mg_G = div_magic((uint32_t) c.G); c.mg_RG = div_magic((uint32_t) c.RG);
    // twiddle tables, laid out [q][j] so that the lanes of a warp read consecutive entries
    const double pi = 3.14159265358979323846;
    hp.tab2.assign((size_t) c.tab_len, make_double2(1.0, 0.0));
    auto w = [&](double num, double den) { return make_double2(std::cos(-2.0 * pi * num / den), std::sin(-2.0 * pi * num / den)); };
    for (int q = 1; q < c.Ax; q++) for (int j = 0; j < c.Bx; j++) hp.tab2[c.o_t1x + (q - 1) * c.Bx + j] = w((double) j * q, P.Lx);
    for (int k2 = 0; k2 < c.Bx; k2++) for (int u = 0; u < c.nux; u++) hp.tab2[c.o_tsx + k2 * c.nux + u] = w(u + c.Ax * k2, 2.0 * P.Lx);
    for (int k2 = 0; k2 < c.Bx; k2++) hp.tab2[c.o_tmx + k2] = w(0.5 * c.Ax + c.Ax * k2, 2.0 * P.Lx);
    for (int q = 1; q < c.Ay; q++) for (int j = 0; j < c.By; j++) hp.tab2[c.o_tay + (q - 1) * c.By + j] = w((double) j * q, P.Ly);
    {   // the constants of the stage functions, as the device reads them (C2K, CB2_KWORDS 32-bit words per block)
        static_assert(sizeof(C2K) == CB2_KWORDS * 4, "C2K layout");
        const C2K kr = c2k_rows(P), kc = c2k_cols(P);
        std::memcpy(&hp.tab2[c.o_kr], &kr, sizeof(C2K));
        std::memcpy(&hp.tab2[c.o_kc], &kc, sizeof(C2K));
    }
    c.id = c2_next_id();
    c.ok = 1;
    return true;
}

inline bool make_plan(int mx, int my, HostPlan &hp, int smem_limit = kSmemMax, int tab_end = 0)
{
    ConvPlan &P = hp.p;
    std::memset(&P, 0, sizeof(P));
    P.mx = mx; P.my = my; P.npot = mx * my;
    P.Fx = opt_fft_size(mx); P.Fy = opt_fft_size(my);
    P.Lx = P.Fx; P.Ly = 2 * P.Fy;
    if (P.Lx == 1) P.nsx = 0;
    else if (!choose_radices(P.Lx, &P.nsx, P.rx)) return false;
    if (!choose_radices(P.Ly, &P.nsy, P.ry)) return false;
    if (P.Lx >= 65535 || P.Ly >= 65535) return false;
    P.SY = my | 1;
    hp.twx.resize(2 * P.Fx); hp.twy.resize(2 * P.Fy);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < 2 * P.Fx; k++) hp.twx[k] = make_double2(std::cos(-2.0 * pi * k / (2.0 * P.Fx)), std::sin(-2.0 * pi * k / (2.0 * P.Fx)));
    for (int k = 0; k < 2 * P.Fy; k++) hp.twy[k] = make_double2(std::cos(-2.0 * pi * k / (2.0 * P.Fy)), std::sin(-2.0 * pi * k / (2.0 * P.Fy)));
    dif_positions(P.Lx, P.nsx, P.rx, hp.posx);

    // shared-memory layout: S | W | twx | twy | posx | red
    const long bytesS = (long) (P.Lx + 1) * P.SY * 16;
    const long bytesT = (long) (2 * P.Fx + 2 * P.Fy) * 16 + ((P.Lx * 2 + 15) / 16) * 16 + 1024;
    const long avail = (long) smem_limit - bytesS - bytesT;
    const int ncol = P.Fx + 1;
    int C = (int) (avail / ((long) P.Ly * 16));
    hp.fits = C >= 1;
    if (C < 1) C = 1;
    if (C > ncol) C = ncol;
    P.nchunk = (ncol + C - 1) / C;
    P.C = (ncol + P.nchunk - 1) / P.nchunk;
    P.chat_len = P.nchunk * P.Ly * P.C;
    // per-stage constants (twiddle tables: twx has 2Fx entries for transforms of length Fx => index step x2)
    for (int s = 0, ns = P.Lx; s < P.nsx; ns /= P.rx[s], s++) {
        StageK &k = P.kx[s];
        k.m = ns / P.rx[s]; k.cnt = P.Lx / P.rx[s]; k.tstep = (P.Lx / ns) * 2; k.mg_m = div_magic((uint32_t) k.m);
    }
    for (int s = 0, ns = P.Ly; s < P.nsy; ns /= P.ry[s], s++) {
        StageK &k = P.ky[s];
        k.m = ns / P.ry[s]; k.cnt = P.Ly / P.ry[s]; k.tstep = P.Ly / ns; k.mg_m = div_magic((uint32_t) k.m);
    }
    // warp-scheduled column pass: every warp owns a W slot for G columns (padded layout: one spare element per 8);
    // the W region is the larger of the block-wide chunk and the warp slots
    const int nwarp = kPlanWarps;
    int G = P.C / nwarp; if (G > 4) G = 4; if (G < 1) G = 1;
    int slot = P.Ly * G + (P.Ly * G) / 8 + 1;
    int wslots = (int) ((avail / 16) / slot); if (wslots > nwarp) wslots = nwarp; if (wslots > (P.C + G - 1) / G * P.nchunk) wslots = (P.C + G - 1) / G * P.nchunk;
    if (wslots < 1) { wslots = 1; if (hp.fits && (long) slot * 16 > avail) hp.fits = false; }
    P.G = G; P.gpc = (P.C + G - 1) / G; P.wslots = wslots; P.wslot_len = slot;
    P.mg_G = div_magic((uint32_t) G); P.mg_gpc = div_magic((uint32_t) P.gpc);
    long bytesW = (long) P.Ly * P.C * 16;
    if ((long) wslots * slot * 16 > bytesW) bytesW = (long) wslots * slot * 16;
    P.off_S = 0;
    P.off_W = (int) bytesS;
    P.off_twx = P.off_W + (int) bytesW;
    P.off_twy = P.off_twx + 2 * P.Fx * 16;
    P.off_posx = P.off_twy + 2 * P.Fy * 16;
    P.off_red = P.off_posx + ((P.Lx * 2 + 15) / 16) * 16;
    P.smem_bytes = P.off_red + 1024;
    { const double N = 4.0 * P.Fx * P.Fy, S = (P.Fx + 1) * 2.0 * P.Fy; P.nom_flops = (unsigned int) (5.0 * N * log2(N) + 6.0 * S); }
    // warp-resident product: same shared-memory window, the reduction scratch sits behind both layouts
    long end2 = 0;
    if (make_plan2(hp, smem_limit, &end2, tab_end)) {
        const int e2 = (int) ((end2 + 15) / 16) * 16;
        if (e2 > P.off_red) { P.off_red = e2; P.smem_bytes = P.off_red + 1024; }
    }
    return true;
}

}  // namespace cb200
