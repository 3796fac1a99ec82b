// Host emulation of the fused convolution: steps the *product's* phase functions (fftconv.cuh) through all thread
// ids sequentially on the CPU.  Test infrastructure only -- lets the CPU-only test suite validate the exact device
// algorithm (indexing, digit-reversed C^ layout, split/merge steps) against the oracle without a GPU.
#include <vector>
#include <cstdio>
#include <cstdint>
#include <algorithm>
// ---- shared-memory wavefront model of the warp-resident product ----
// TraceBuf records, per lane, the sequence of 16-byte shared-memory accesses of one stage (the lanes are stepped one after
// the other, every lane issues the same instruction sequence).  A 16-byte access is served per quarter-warp (8 lanes); a
// quarter needs as many wavefronts as the largest number of DISTINCT addresses that fall into one of the 8 bank groups.
struct TraceLog { std::vector<std::vector<uint32_t>> lane; long wf = 0, ideal = 0, instr = 0; long swf[10] = { 0 }, sideal[10] = { 0 }; int stage = 0; };
static TraceLog *g_trace = nullptr;
static int g_trace_lane = 0;
static void trace_flush(TraceLog &t)
{
    size_t n = 0;
    for (auto &l : t.lane) n = std::max(n, l.size());
    for (size_t k = 0; k < n; k++) {
        int active = 0;
        for (int q = 0; q < 4; q++) {
            int cnt[8] = { 0 }; uint32_t seen[8][8];
            for (int l = 8 * q; l < 8 * q + 8; l++) {
                if (t.lane[l].size() <= k) continue;
                const uint32_t a = t.lane[l][k]; if (a == 0xFFFFFFFFu) continue;
                const int b = a & 7; bool dup = false;
                for (int z = 0; z < cnt[b]; z++) dup = dup || seen[b][z] == a;
                if (!dup) seen[b][cnt[b]++] = a;
                active++;
            }
            int m = 0; for (int b = 0; b < 8; b++) m = std::max(m, cnt[b]);
            t.wf += m; t.swf[t.stage] += m;
        }
        t.ideal += (active + 7) / 8; t.sideal[t.stage] += (active + 7) / 8; t.instr++;
    }
    for (auto &l : t.lane) l.clear();
}
#define CB2_TICK(i) do { if (g_trace) g_trace->stage = (i); } while (0)
#define CB2_LANE_BEGIN(lane) (g_trace_lane = (lane))
#define CB2_LANES_END() do { if (g_trace) trace_flush(*g_trace); } while (0)
#include "../../contact_b200/csrc/plan.h"
#include "../../contact_b200/csrc/fftconv_warp.cuh"

using namespace cb200;

struct TraceBuf {
    cd *p;
    cd ld(uint32_t i) const { if (g_trace) g_trace->lane[g_trace_lane].push_back(i); return p[i]; }
    void st(uint32_t i, cd v) const { if (g_trace) g_trace->lane[g_trace_lane].push_back(i); p[i] = v; }
    int ldi(uint32_t w) const { return reinterpret_cast<const int *>(p)[w]; }
    cd ldz(uint32_t i, bool ok) const { if (g_trace) g_trace->lane[g_trace_lane].push_back(ok ? i : 0xFFFFFFFFu); return ok ? p[i] : make_double2(0.0, 0.0); }
    void stp(uint32_t i, cd v, bool ok) const { if (g_trace) g_trace->lane[g_trace_lane].push_back(ok ? i : 0xFFFFFFFFu); if (ok) p[i] = v; }
};

#define CB_PHASE(call) do { for (int tid = 0; tid < nthr; tid++) { call; } } while (0)
#include "../../contact_b200/csrc/conv_sequence.inc"

extern "C" int emul_plan_info(int mx, int my, int *out)
{
    HostPlan hp;
    if (!make_plan(mx, my, hp)) return -1;
    out[0] = hp.p.Fx; out[1] = hp.p.Fy; out[2] = hp.p.C; out[3] = hp.p.nchunk; out[4] = hp.p.smem_bytes;
    out[5] = hp.fits; out[6] = hp.p.nsx; out[7] = hp.p.nsy;
    for (int i = 0; i < CB_MAXSTAGE; i++) { out[8 + i] = hp.p.rx[i]; out[16 + i] = hp.p.ry[i]; }
    return 0;
}

// u = mask (.) conv(cf block, p) * scale ; cf block has half sizes (cmx, cmy)
static int emul_conv_impl(int mx, int my, const double *p, const double *cfblk, int cmx, int cmy, double scale,
                          const int *el, int mask_mode, int add, double *u, int nthr, int warp_sched)
{
    HostPlan hp;
    if (!make_plan(mx, my, hp)) return -1;
    const ConvPlan &P = hp.p;
    const MemBuf<const cd> twx = { hp.twx.data() }, twy = { hp.twy.data() };
    const MemBuf<const unsigned short> posx = { hp.posx.data() };

    // --- build C^ with the coefficient sequence (global-scratch style: S holds all 2Fy rows)
    std::vector<cd> chat(P.chat_len);
    {
        const int SY = 2 * P.Fy;
        std::vector<cd> SW((size_t) (P.Lx + 1) * SY + (size_t) P.Ly * P.C);
        typedef MemBuf<cd> CB_BUF;
        const CB_BUF BUF = { SW.data() };
        const uint32_t oS = 0u, oW = (uint32_t) (P.Lx + 1) * SY;
        RowSrc src = { cfblk, 1, P.Fx < mx ? P.Fx : mx, P.Fy < my ? P.Fy : my, cmx, cmy, P.Fx, P.Fy, 0 };
        CB_CONV_FORWARD_ROWS(2 * P.Fy, src);
        CB_CONV_COLUMNS_DUMP(2 * P.Fy, chat.data(), scale / (4.0 * P.Fx * P.Fy));
    }
    // --- the product
    {
        const int SY = P.SY;
        std::vector<cd> SW((size_t) (P.off_twx / 16) + 16);        // S | W exactly as laid out in shared memory
        typedef MemBuf<cd> CB_BUF;
        const CB_BUF BUF = { SW.data() };
        const uint32_t oS = 0u, oW = (uint32_t) (P.off_W / 16);
        if (warp_sched) {
            // the warp-scheduled product (fftconv_warp.cuh): one call per warp and pass, lanes looped inside
            const int nwarps = nthr / 32 > 0 ? nthr / 32 : 1;
            for (int w = 0; w < nwarps; w++) warp_rows_fwd(P, BUF, oS, SY, p, mx, my, mx, twx, posx, w, nwarps);
            for (int w = 0; w < nwarps; w++) warp_cols(P, BUF, oS, oW, SY, my, my, chat.data(), twy, w, nwarps);
            for (int w = 0; w < nwarps; w++) warp_rows_inv(P, BUF, oS, SY, u, el, mask_mode, add, 0, 0, mx, my, mx, twx, posx, w, nwarps);
            return 0;
        }
        RowSrc src = { p, 0, mx, my, 0, 0, P.Fx, P.Fy, 0 };
        CB_CONV_FORWARD_ROWS(P.my, src);
        CB_CONV_COLUMNS_PRODUCT(P.my, chat.data());
        CB_CONV_INVERSE_ROWS_STORE(P.my, u, el, mask_mode, add, 0, 0, mx, mx);
    }
    return 0;
}

extern "C" int emul_conv(int mx, int my, const double *p, const double *cfblk, int cmx, int cmy, double scale,
                         const int *el, int mask_mode, int add, double *u, int nthr)
{ return emul_conv_impl(mx, my, p, cfblk, cmx, cmy, scale, el, mask_mode, add, u, nthr, 0); }

// the same product through the warp-scheduled passes with nthr/32 warps
extern "C" int emul_conv_warp(int mx, int my, const double *p, const double *cfblk, int cmx, int cmy, double scale,
                              const int *el, int mask_mode, int add, double *u, int nthr)
{ return emul_conv_impl(mx, my, p, cfblk, cmx, cmy, scale, el, mask_mode, add, u, nthr, 1); }

// radix butterflies against a naive DFT
extern "C" double emul_radix_error(int R, int inv)
{
    cd x[18], y[18];
    double err = 0.0;
    for (int q = 0; q < R; q++) { x[q] = make_double2(0.3 + 0.7 * q - 0.05 * q * q, -0.2 + 0.11 * q * q * q / 7.0); y[q] = x[q]; }
#define RUN(RR) case RR: if (inv) Dft<RR, true>::run(y); else Dft<RR, false>::run(y); break;
    switch (R) { RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(12) RUN(16) RUN(10) RUN(18) default: return -1; }
    const double pi = 3.14159265358979323846, sg = inv ? 1.0 : -1.0;
    for (int k = 0; k < R; k++) {
        double re = 0, im = 0;
        for (int q = 0; q < R; q++) {
            const double a = sg * 2.0 * pi * q * k / R;
            re += x[q].x * cos(a) - x[q].y * sin(a);
            im += x[q].x * sin(a) + x[q].y * cos(a);
        }
        err = fmax(err, fmax(fabs(re - y[k].x), fabs(im - y[k].y)));
    }
    return err;
}

// ---- warp-resident product (fftconv2.cuh): the three passes stepped warp by warp, lanes looped; coefficient transform by
//      the dense DFT entries the device builder kernels use.  Box (x0, y0, bw x bh) of the mx x my grid; plan of (pmx, pmy).
extern "C" int emul_plan2_info(int mx, int my, int *out)
{
    HostPlan hp;
    if (!make_plan(mx, my, hp)) return -1;
    const Conv2Plan &c = hp.p.c2;
    const int v[16] = { c.ok, c.Ax, c.Bx, c.Ay, c.By, c.G, c.RG, c.nslot, c.slot_len, c.ngrp, c.chat_len, c.tab_len,
                        hp.p.smem_bytes, hp.p.off_red, c.off_tab, c.nux };
    for (int i = 0; i < 16; i++) out[i] = v[i];
    return 0;
}

// fuse: in_mode (0 none, 1 shift, 2 scale) with in_val on the elements el >= 1 of the box, written back to p; out_sub
// (null: none) subtracted from the result; *out_sum receives the sum of the stored result over el >= 1 (null: not wanted)
extern "C" int emul_conv2_fused(int pmx, int pmy, int mx, double *p, const double *cfblk, int cmx, int cmy, double scale,
                                const int *el, int mask_mode, int add, double *u, int x0, int y0, int bw, int bh, int nwarps,
                                int in_mode, double in_val, const double *out_sub, double *out_sum);

extern "C" int emul_conv2(int pmx, int pmy, int mx, const double *p, const double *cfblk, int cmx, int cmy, double scale,
                          const int *el, int mask_mode, int add, double *u, int x0, int y0, int bw, int bh, int nwarps)
{
    return emul_conv2_fused(pmx, pmy, mx, const_cast<double *>(p), cfblk, cmx, cmy, scale, el, mask_mode, add, u, x0, y0, bw, bh,
                            nwarps, 0, 0.0, nullptr, nullptr);
}

extern "C" int emul_conv2_fused(int pmx, int pmy, int mx, double *p, const double *cfblk, int cmx, int cmy, double scale,
                                const int *el, int mask_mode, int add, double *u, int x0, int y0, int bw, int bh, int nwarps,
                                int in_mode, double in_val, const double *out_sub, double *out_sum)
{
    HostPlan hp;
    if (!make_plan(pmx, pmy, hp)) return -1;
    const ConvPlan &P = hp.p;
    const Conv2Plan &c = P.c2;
    if (!c.ok) return -2;
    // C^ in the layout [group][k2][column][k1]
    std::vector<cd> chat((size_t) c.chat_len, make_double2(0.0, 0.0));
    {
        RowSrc src = { cfblk, 1, P.Fx < cmx ? P.Fx : cmx, P.Fy < cmy ? P.Fy : cmy, cmx, cmy, P.Fx, P.Fy, 0, 0 };
        std::vector<cd> T((size_t) 2 * P.Fy * (P.Fx + 1));
        for (int iy = 0; iy < 2 * P.Fy; iy++)
            for (int kx = 0; kx <= P.Fx; kx++) T[(size_t) iy * (P.Fx + 1) + kx] = c2_chat_row_entry(P, src, iy, kx, hp.twx.data());
        for (int kx = 0; kx <= P.Fx; kx++)
            for (int ky = 0; ky < 2 * P.Fy; ky++)
                chat[c2_chat_index(P, kx, ky)] = c2_chat_value(P, T.data(), kx, ky, hp.twy.data(), scale / (4.0 * P.Fx * P.Fy));
    }
    std::vector<cd> sm((size_t) (P.off_red / 16) + 64, make_double2(1e300, -1e300));   // poisoned: stale reads show up
    for (int i = 0; i < c.tab_len; i++) sm[c.off_tab / 16 + i] = hp.tab2[i];
    const MemBuf<cd> buf = { sm.data() };
    for (int w = 0; w < nwarps; w++)
        c2_rows_fwd(P, buf, p + (size_t) y0 * mx + x0, bw, bh, mx, w, in_mode ? el + (size_t) y0 * mx + x0 : nullptr, in_val, in_mode);
    for (int w = 0; w < nwarps; w++) c2_cols(P, buf, chat.data(), bh, bh, w);
    for (int w = 0; w < nwarps; w++) c2_rows_inv(P, buf, u, el, mask_mode, add, x0, y0, bw, bh, mx, w, out_sub, out_sum ? 1 : 0);
    if (out_sum) {                                   // per-warp partials behind the tables, summed in warp order (conv2_box_dev)
        double s_ = 0.0;
        for (int w = 0; w < c.nslot && w < nwarps; w++) s_ += sm[c.off_tab / 16 + c.tab_len + w].x;
        *out_sum = s_;
    }
    return 0;
}

// input-pruned butterflies against the full ones on half-zero input
extern "C" double emul_halfin_error(int R, int inv)
{
    cd x[18], y[18];
    for (int q = 0; q < R; q++) { x[q] = q < R / 2 ? make_double2(0.3 + 0.7 * q, -0.2 + 0.11 * q * q) : make_double2(0.0, 0.0); y[q] = x[q]; }
#define RUNH(RR) case RR: if (inv) { DftHalfIn<RR, true>::run(y); Dft<RR, true>::run(x); } else { DftHalfIn<RR, false>::run(y); Dft<RR, false>::run(x); } break;
    switch (R) { RUNH(4) RUNH(6) RUNH(8) RUNH(10) RUNH(12) RUNH(16) RUNH(18) default: return -1; }
    double err = 0.0;
    for (int k = 0; k < R; k++) err = fmax(err, fmax(fabs(x[k].x - y[k].x), fabs(x[k].y - y[k].y)));
    return err;
}

// out[0..5]: modelled / ideal wavefronts of rows fwd, columns, rows inv of one full-grid product; out[6..25]: per stage
extern "C" int emul_conv2_wavefronts(int mx, int my, long *out)
{
    HostPlan hp;
    if (!make_plan(mx, my, hp) || !hp.p.c2.ok) return -1;
    const ConvPlan &P = hp.p;
    const Conv2Plan &c = P.c2;
    std::vector<cd> chat((size_t) c.chat_len, make_double2(1.0, 0.0));
    std::vector<cd> sm((size_t) (P.off_red / 16) + 64, make_double2(0.0, 0.0));
    for (int i = 0; i < c.tab_len; i++) sm[c.off_tab / 16 + i] = hp.tab2[i];
    std::vector<double> p((size_t) mx * my, 1.0), u((size_t) mx * my, 0.0);
    const TraceBuf buf = { sm.data() };
    TraceLog t; t.lane.resize(32);
    g_trace = &t;
    for (int w = 0; w < 12; w++) c2_rows_fwd(P, buf, p.data(), mx, my, mx, w);
    out[0] = t.wf; out[1] = t.ideal; t.wf = t.ideal = 0;
    for (int w = 0; w < 12; w++) c2_cols(P, buf, chat.data(), my, my, w);
    out[2] = t.wf; out[3] = t.ideal; t.wf = t.ideal = 0;
    for (int w = 0; w < 12; w++) c2_rows_inv(P, buf, u.data(), nullptr, 0, 0, 0, 0, mx, my, mx, w);
    out[4] = t.wf; out[5] = t.ideal;
    for (int i = 0; i < 10; i++) { out[6 + 2 * i] = t.swf[i]; out[7 + 2 * i] = t.sideal[i]; }
    g_trace = nullptr;
    return 0;
}
