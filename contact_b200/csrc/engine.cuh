// engine.cuh -- host-side engine: coefficient sets cached per (grid, material, rolling step), their transforms,
// the FFT preconditioner, launch wrappers.  Everything numeric runs on the device; the host only plans and launches.
//
// "coefficients cached per grid and material": the reference recomputes or reuses cs/cv/csv in sgencr
// (/root/reference/src/m_visc.f90:127-376) and re-transforms them whenever the FFT size changes
// (m_aijpj.f90:873-920); here one CoefSet owns the spatial blocks and the lazily built C^ per (set, ik, jk).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "kernels.cuh"
#include "plan.h"

namespace cb200 {



inline std::string &last_error() { static thread_local std::string e; return e; }

#define CB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof(b_), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            last_error() = b_;                                                                     \
            return -99;                                                                            \
        }                                                                                          \
    } while (0)

enum { SET_CS = 0, SET_CV = 1, SET_CSV = 2, SET_MS = 3 };

struct Material { double gg[2], poiss[2], ga, nu, ak; };

inline void combine_material(Material &m)
{   // m_hierarch_data.f90:1543-1553
    m.ga = 2.0 / (1.0 / m.gg[0] + 1.0 / m.gg[1]);
    m.nu = m.ga * (m.poiss[0] / m.gg[0] + m.poiss[1] / m.gg[1]) / 2.0;
    m.ak = (m.ga / 4.0) * ((1.0 - 2.0 * m.poiss[0]) / m.gg[0] - (1.0 - 2.0 * m.poiss[1]) / m.gg[1]);
}

struct CoefKey {
    int mx, my; double dx, dy, ga, nu, ak; int is_roll; double chi, dq;
    int whole_gpu;     // 1: plan and transforms for the whole-GPU path even though the grid fits one CTA (latency mode)
    bool operator<(const CoefKey &o) const {
        const double a[] = { (double) mx, (double) my, dx, dy, ga, nu, ak, (double) is_roll, chi, dq, (double) whole_gpu };
        const double b[] = { (double) o.mx, (double) o.my, o.dx, o.dy, o.ga, o.nu, o.ak, (double) o.is_roll, o.chi, o.dq, (double) o.whole_gpu };
        for (int i = 0; i < 11; i++) { if (a[i] < b[i]) return true; if (a[i] > b[i]) return false; }
        return false;
    }
};

struct CoefSet {
    CoefKey key;
    int mx, my;
    bool nt_cpl;
    double ga, ga_inv;
    HostPlan hp;                  // plan for products on the full mx x my grid
    cd *d_twx = nullptr, *d_twy = nullptr;
    unsigned short *d_posx = nullptr;
    double *d_cf[4] = { nullptr, nullptr, nullptr, nullptr };     // spatial coefficients, 9 blocks each
    cd *d_chat[4][3][3] = {};                                      // transformed blocks (lazy)
    bool prec_ready[3] = { false, false, false };
    long n_chat_built = 0;
    // subsurface: transformed coefficients per depth list / material, [nz][4][9][chat_len]
    std::map<std::vector<double>, cd *> subs_chat;
    // grids beyond one CTA's shared memory (hp.fits == false): whole-GPU plan, spectrum workspace, reduction scratch;
    // d_chat then holds the transforms in the column-task layout [ntc][Ly][CB]
    LargePlan lp;
    cd *d_T = nullptr;
    double *d_gpart = nullptr;
    cudaEvent_t ev_l0 = nullptr, ev_l1 = nullptr;   // bracket the three phase kernels of the last stand-alone product
    // ladder of smaller transform sizes for products on the bounding box of the contact area (single-CTA path)
    std::vector<HostPlan> lev_hp;
    ConvLevel *d_lev = nullptr;
    int nlx = 0, nly = 0;
    bool lev_tang = false;
    // bytes at the bottom of the shared-memory window below the tables of every plan of this set (staged vector passes)
    int stage_bytes() const {
        int b = hp.p.c2.ok ? hp.p.c2.off_tab : hp.p.off_red;
        for (const HostPlan &l : lev_hp) if (l.p.c2.ok && l.p.c2.off_tab < b) b = l.p.c2.off_tab;
        return b;
    }
};

struct Engine {
    std::mutex mu;
    std::map<CoefKey, CoefSet *> sets;
    std::vector<CoefSet *> by_handle;
    int device = -1;
    int num_sms = 0;
    bool attr_set = false;
    long launches = 0;            // kernels launched by this library (for bench "gpu_launches")
};

// One instance per CUDA device: the calling thread's current device selects it.  cntc_calculate_batch can fan a batch out over
// several devices (one host thread per device, cb200_set_devices): each device has its own coefficient cache, work-space pool,
// events and kernel attributes, so the per-device threads share nothing but the problem registry.
#define CB_MAX_DEVICES 16
inline int current_device() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) d = 0; return (d >= 0 && d < CB_MAX_DEVICES) ? d : 0; }
inline Engine &engine() { static Engine e[CB_MAX_DEVICES]; return e[current_device()]; }

inline int engine_init()
{
    Engine &E = engine();
    if (E.device >= 0) return 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        last_error() = "contact_addon_b200: no CUDA device available (this library has no CPU fallback)";
        return -99;
    }
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CB_CUDA(cudaGetDeviceProperties(&prop, dev));
    E.device = dev;
    E.num_sms = prop.multiProcessorCount;
    CB_CUDA(cudaFuncSetAttribute(k_snorm_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CB_CUDA(cudaFuncSetAttribute(k_build_chat, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CB_CUDA(cudaFuncSetAttribute(k_subsurf_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CB_CUDA(cudaFuncSetAttribute(k_contac_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CB_CUDA(cudaFuncSetAttribute(k_lg_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CB_CUDA(cudaFuncSetAttribute(k_lg_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CB_CUDA(cudaFuncSetAttribute(k_lg_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CB_CUDA(cudaFuncSetAttribute(k_lg_snorm, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CB_CUDA(cudaFuncSetAttribute(k_lg_contac, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    return 0;
}

inline int grid1d(long n, int b) { return (int) ((n + b - 1) / b); }

// ---- whole-GPU plan for grids that do not fit one CTA ----
inline bool make_large_plan(const ConvPlan &P, int nsm, LargePlan &L)
{
    L.P = P;
    const int ncol = P.Fx + 1;
    const long budget = kSmemMax - 1024;
    L.RB = (P.my + nsm - 1) / nsm;
    while (L.RB > 1 && (long) (P.Lx + 1) * (L.RB | 1) * 16 > budget) L.RB--;
    L.CB = (ncol + nsm - 1) / nsm;
    while (L.CB > 1 && (long) P.Ly * L.CB * 16 > budget) L.CB--;
    if ((long) (P.Lx + 1) * (L.RB | 1) * 16 > budget || (long) P.Ly * L.CB * 16 > budget) return false;
    L.ntr = (P.my + L.RB - 1) / L.RB;
    L.ntc = (ncol + L.CB - 1) / L.CB;
    L.ldT = P.my;
    // the coefficient transform runs the row pass over 2Fy rows with the same RB
    const long sr = (long) (P.Lx + 1) * (L.RB | 1) * 16, sc = (long) P.Ly * L.CB * 16;
    L.smem_bytes = (int) (sr > sc ? sr : sc) + 1024;
    L.P.chat_len = L.ntc * P.Ly * L.CB;
    return true;
}

inline int build_chat_large(CoefSet &cs, int set, int ik, int jk, cudaStream_t st)
{
    Engine &E = engine();
    const LargePlan &L = cs.lp;
    const ConvPlan &P = L.P;
    cd *chat = nullptr, *T2 = nullptr;
    CB_CUDA(cudaMalloc(&chat, sizeof(cd) * (size_t) P.chat_len));
    CB_CUDA(cudaMalloc(&T2, sizeof(cd) * (size_t) (P.Lx + 1) * 2 * P.Fy));
    RowSrc src;
    src.base = cs.d_cf[set] + (size_t) ((jk - 1) * 3 + (ik - 1)) * 4 * cs.mx * cs.my;
    src.kind = 1; src.mx = std::min(P.Fx, P.mx); src.my = std::min(P.Fy, P.my); src.cmx = cs.mx; src.cmy = cs.my;
    src.Fx = P.Fx; src.Fy = P.Fy; src.row0 = 0; src.stride = 0;
    const double scale = cs.ga_inv / (4.0 * P.Fx * P.Fy);
    const int nrows = 2 * P.Fy, ntask = (nrows + L.RB - 1) / L.RB;
    k_lg_rows_fwd<<<std::min(ntask, 4 * E.num_sms), CB_THREADS, L.smem_bytes, st>>>(L, src, nrows, L.RB, T2, nrows);
    k_lg_cols<<<L.ntc, CB_THREADS, L.smem_bytes, st>>>(L, nrows, 0, T2, nrows, nullptr, chat, scale);
    E.launches += 2;
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaStreamSynchronize(st));
    CB_CUDA(cudaFree(T2));
    cs.d_chat[set][ik - 1][jk - 1] = chat;
    cs.n_chat_built++;
    return 0;
}

// cd elements of one transformed block in the layout the plan's product reads
inline size_t plan_chat_len(const ConvPlan &P) { return P.c2.ok ? (size_t) P.c2.chat_len : (size_t) P.chat_len; }

// transform nblk consecutive spatial blocks (4 cmx cmy doubles each) into chat (plan_chat_len each); synchronises st
inline int launch_build_chat(const ConvPlan &P, const double *cf, int nblk, int cmx, int cmy, double scale, cd *chat,
                             cudaStream_t st)
{
    Engine &E = engine();
    cd *scr = nullptr;
    if (P.c2.ok) {
        const int n = 2 * P.Fy * (P.Fx + 1);
        CB_CUDA(cudaMalloc(&scr, sizeof(cd) * (size_t) n * nblk));
        CB_CUDA(cudaMemsetAsync(chat, 0, sizeof(cd) * plan_chat_len(P) * nblk, st));      // padding columns of the last group
        k_chat2_rows<<<dim3(grid1d(n, 128), nblk), 128, 0, st>>>(P, cf, cmx, cmy, scr);
        k_chat2_cols<<<dim3(grid1d(n, 128), nblk), 128, 0, st>>>(P, scr, scale, chat);
        E.launches += 2;
    } else {
        const size_t nscr = (size_t) (P.Lx + 1) * 2 * P.Fy + (size_t) P.Ly * P.C;
        CB_CUDA(cudaMalloc(&scr, sizeof(cd) * nscr * nblk));
        k_build_chat<<<nblk, CB_THREADS, 64, st>>>(P, cf, cmx, cmy, scale, scr, chat);
        E.launches++;
    }
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaStreamSynchronize(st));
    CB_CUDA(cudaFree(scr));
    return 0;
}

// ---- coefficient transform of one block ----
inline int build_chat(CoefSet &cs, int set, int ik, int jk, cudaStream_t st)
{
    Engine &E = engine();
    if (cs.d_chat[set][ik - 1][jk - 1]) return 0;
    if (!cs.hp.fits) return build_chat_large(cs, set, ik, jk, st);
    const ConvPlan &P = cs.hp.p;
    cd *chat = nullptr;
    CB_CUDA(cudaMalloc(&chat, sizeof(cd) * plan_chat_len(P)));
    const double *blk = cs.d_cf[set] + (size_t) ((jk - 1) * 3 + (ik - 1)) * 4 * cs.mx * cs.my;
    const double scale = cs.ga_inv / (4.0 * P.Fx * P.Fy);
    int rc = launch_build_chat(P, blk, 1, cs.mx, cs.my, scale, chat, st);
    if (rc) return rc;
    cs.d_chat[set][ik - 1][jk - 1] = chat;
    cs.n_chat_built++;
    return 0;
}

// ---- ladder of transform sizes for the contact-box products (m_aijpj.f90:774-793): sizes ~ 1, .8, .64, .5, .4, .3 of
//      the grid in each direction, every combination with its own plan, tables and transformed cs(3,3), ms(3,3) ----
inline int build_levels(CoefSet &cs, cudaStream_t st, bool tang = false)
{
    Engine &E = engine();
    if (!cs.hp.fits) return 0;
    if (cs.d_lev && (!tang || cs.lev_tang)) return 0;
    const size_t nblk = (size_t) 4 * cs.mx * cs.my;
    if (!cs.d_lev) {
        // sizes with two large radix stages per direction (48, 64, 72, 96 ...) are preferred over the tightest fit
        auto ladder = [](int n) {
            static const int nice[] = { 192, 144, 128, 96, 80, 72, 64, 48, 32, 24, 16, 12, 8 };
            std::vector<int> v(1, n);
            for (int c : nice) if (c < n && (int) v.size() < 8) v.push_back(c);
            return v;
        };
        const std::vector<int> fx = ladder(cs.mx), fy = ladder(cs.my);
        cs.nlx = (int) fx.size(); cs.nly = (int) fy.size();
        if (cs.nlx * cs.nly <= 1) return 0;
        cs.lev_hp.resize((size_t) cs.nlx * cs.nly);
        std::vector<ConvLevel> h((size_t) cs.nlx * cs.nly);
        for (int ly = 0; ly < cs.nly; ly++)
            for (int lx = 0; lx < cs.nlx; lx++) {
                HostPlan &hp = cs.lev_hp[(size_t) ly * cs.nlx + lx];
                ConvLevel &L = h[(size_t) ly * cs.nlx + lx];
                memset(&L, 0, sizeof(L));
                if (lx == 0 && ly == 0) { L.P = cs.hp.p; continue; }
                // the level plans live in the dynamic shared memory that the kernels are launched with for the full-size plan
                if (!make_plan(fx[lx], fy[ly], hp, cs.hp.p.smem_bytes, c2_tab_end(cs.hp.p)) || !hp.fits) { last_error() = "internal: level plan does not fit"; return -99; }
                ConvPlan &P = hp.p;
                cd *twx, *twy; unsigned short *posx;
                CB_CUDA(cudaMalloc(&twx, sizeof(cd) * hp.twx.size()));
                CB_CUDA(cudaMalloc(&twy, sizeof(cd) * hp.twy.size()));
                CB_CUDA(cudaMalloc(&posx, sizeof(unsigned short) * hp.posx.size()));
                CB_CUDA(cudaMemcpyAsync(twx, hp.twx.data(), sizeof(cd) * hp.twx.size(), cudaMemcpyHostToDevice, st));
                CB_CUDA(cudaMemcpyAsync(twy, hp.twy.data(), sizeof(cd) * hp.twy.size(), cudaMemcpyHostToDevice, st));
                CB_CUDA(cudaMemcpyAsync(posx, hp.posx.data(), sizeof(unsigned short) * hp.posx.size(), cudaMemcpyHostToDevice, st));
                P.twx = twx; P.twy = twy; P.posx = posx;
                if (P.c2.ok) {
                    cd *tab2;
                    CB_CUDA(cudaMalloc(&tab2, sizeof(cd) * hp.tab2.size()));
                    CB_CUDA(cudaMemcpyAsync(tab2, hp.tab2.data(), sizeof(cd) * hp.tab2.size(), cudaMemcpyHostToDevice, st));
                    P.c2.tab = tab2;
                }
                L.P = P;
            }
        CB_CUDA(cudaMalloc(&cs.d_lev, sizeof(ConvLevel) * h.size()));
        CB_CUDA(cudaMemcpy(cs.d_lev, h.data(), sizeof(ConvLevel) * h.size(), cudaMemcpyHostToDevice));
    }
    // transformed blocks per level: all nine blocks of cs (and the diagonal of ms) in one launch per (level, set); the
    // normal problem needs cs(3,3), ms(3,3), the tangential problem cs(1:2,1:3), ms(1,1), ms(2,2)
    std::vector<ConvLevel> h((size_t) cs.nlx * cs.nly);
    CB_CUDA(cudaMemcpy(h.data(), cs.d_lev, sizeof(ConvLevel) * h.size(), cudaMemcpyDeviceToHost));
    for (size_t l = 1; l < h.size(); l++) {
        const ConvPlan &P = h[l].P;
        for (int set = 0; set < 2; set++) {
            const bool have33 = h[l].chat[set][2][2] != nullptr, have11 = h[l].chat[set][0][0] != nullptr;
            if (have33 && (!tang || have11)) continue;
            cd *chat = nullptr;
            CB_CUDA(cudaMalloc(&chat, sizeof(cd) * 9 * plan_chat_len(P)));
            int rcb = launch_build_chat(P, cs.d_cf[set ? SET_MS : SET_CS], 9, cs.mx, cs.my, cs.ga_inv / (4.0 * P.Fx * P.Fy), chat, st);
            if (rcb) return rcb;
            for (int jk = 0; jk < 3; jk++) for (int ik = 0; ik < 3; ik++) {
                const bool diag = ik == jk, nt = (ik == 2) != (jk == 2);
                if (set == 1 && !diag) continue;                       // the preconditioner has diagonal blocks only
                if (nt && !cs.nt_cpl) continue;
                if (set == 1 && !cs.prec_ready[ik]) continue;          // ms(ik,ik) not built: leave null -> full grid
                h[l].chat[set][ik][jk] = chat + (size_t) (jk * 3 + ik) * plan_chat_len(P);
            }
        }
    }
    CB_CUDA(cudaMemcpy(cs.d_lev, h.data(), sizeof(ConvLevel) * h.size(), cudaMemcpyHostToDevice));
    if (tang) cs.lev_tang = true;
    (void) nblk;
    return 0;
}

// ---- FFT preconditioner: ms(3,3) = G^2 IFFT2(1 / FFT2(cs(3,3))) on the un-optimised (2mx x 2my) array ----
// Follows fft_makePrec (/root/reference/src/m_aijpj.f90:457-708).  The size is not FFT friendly (e.g. 2*7*13), so
// this one-off transform is done as dense separable DFTs.
inline int build_prec(CoefSet &cs, cudaStream_t st, int ik = 3)
{
    Engine &E = engine();
    if (cs.prec_ready[ik - 1]) return 0;
    const int n1 = 2 * cs.mx, n2 = 2 * cs.my;
    const long n = (long) n1 * n2;
    std::vector<cd> t1(n1), t2(n2);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < n1; k++) t1[k] = make_double2(cos(-2.0 * pi * k / n1), sin(-2.0 * pi * k / n1));
    for (int k = 0; k < n2; k++) t2[k] = make_double2(cos(-2.0 * pi * k / n2), sin(-2.0 * pi * k / n2));
    cd *d_t1, *d_t2, *a, *b;
    CB_CUDA(cudaMalloc(&d_t1, sizeof(cd) * n1));
    CB_CUDA(cudaMalloc(&d_t2, sizeof(cd) * n2));
    CB_CUDA(cudaMalloc(&a, sizeof(cd) * n));
    CB_CUDA(cudaMalloc(&b, sizeof(cd) * n));
    CB_CUDA(cudaMemcpyAsync(d_t1, t1.data(), sizeof(cd) * n1, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(d_t2, t2.data(), sizeof(cd) * n2, cudaMemcpyHostToDevice, st));
    const size_t blk_ = (size_t) ((ik - 1) * 3 + (ik - 1));
    const double *c33 = cs.d_cf[SET_CS] + blk_ * n;
    if (!cs.d_cf[SET_MS]) {
        CB_CUDA(cudaMalloc(&cs.d_cf[SET_MS], sizeof(double) * 9 * n));
        CB_CUDA(cudaMemsetAsync(cs.d_cf[SET_MS], 0, sizeof(double) * 9 * n, st));
    }
    const int B = 256, G = grid1d(n, B);
    k_real_to_cplx<<<G, B, 0, st>>>(c33, a, n);
    k_dft_axis<<<G, B, 0, st>>>(a, b, n1, n2, 0, 0, d_t1);
    k_dft_axis<<<G, B, 0, st>>>(b, a, n1, n2, 1, 0, d_t2);
    k_cplx_recip<<<G, B, 0, st>>>(a, n);
    k_dft_axis<<<G, B, 0, st>>>(a, b, n1, n2, 1, 1, d_t2);
    k_dft_axis<<<G, B, 0, st>>>(b, a, n1, n2, 0, 1, d_t1);
    k_cplx_real_scaled<<<G, B, 0, st>>>(a, cs.d_cf[SET_MS] + blk_ * n, n, cs.ga * cs.ga / (double) n);
    E.launches += 7;
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_t1); cudaFree(d_t2); cudaFree(a); cudaFree(b);
    cs.prec_ready[ik - 1] = true;
    return 0;
}

// ---- create / look up a coefficient set ----
inline int get_coefset(int mx, int my, double dx, double dy, Material mat, int is_roll, double chi, double dq,
                       cudaStream_t st, CoefSet **out, int whole_gpu = 0)
{
    int rc = engine_init();
    if (rc) return rc;
    Engine &E = engine();
    combine_material(mat);
    CoefKey key = { mx, my, dx, dy, mat.ga, mat.nu, mat.ak, is_roll, is_roll ? chi : 0.0, is_roll ? dq : 0.0, whole_gpu };
    std::lock_guard<std::mutex> lk(E.mu);
    auto it = E.sets.find(key);
    if (it != E.sets.end()) { *out = it->second; return 0; }

    CoefSet *cs = new CoefSet();
    cs->key = key; cs->mx = mx; cs->my = my;
    cs->ga = mat.ga; cs->ga_inv = 1.0 / mat.ga;
    cs->nt_cpl = !(fabs(mat.ak) < 1e-6);                             // m_visc.f90:239-243
    if (!make_plan(mx, my, cs->hp)) { last_error() = "unsupported grid size for the FFT product"; delete cs; return -34; }
    if (whole_gpu) { cs->hp.fits = false; cs->hp.p.c2.ok = 0; }
    ConvPlan &P = cs->hp.p;
    CB_CUDA(cudaMalloc(&cs->d_twx, sizeof(cd) * cs->hp.twx.size()));
    CB_CUDA(cudaMalloc(&cs->d_twy, sizeof(cd) * cs->hp.twy.size()));
    CB_CUDA(cudaMalloc(&cs->d_posx, sizeof(unsigned short) * cs->hp.posx.size()));
    CB_CUDA(cudaMemcpyAsync(cs->d_twx, cs->hp.twx.data(), sizeof(cd) * cs->hp.twx.size(), cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(cs->d_twy, cs->hp.twy.data(), sizeof(cd) * cs->hp.twy.size(), cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(cs->d_posx, cs->hp.posx.data(), sizeof(unsigned short) * cs->hp.posx.size(), cudaMemcpyHostToDevice, st));
    P.twx = cs->d_twx; P.twy = cs->d_twy; P.posx = cs->d_posx;
    if (P.c2.ok) {
        cd *tab2;
        CB_CUDA(cudaMalloc(&tab2, sizeof(cd) * cs->hp.tab2.size()));
        CB_CUDA(cudaMemcpyAsync(tab2, cs->hp.tab2.data(), sizeof(cd) * cs->hp.tab2.size(), cudaMemcpyHostToDevice, st));
        P.c2.tab = tab2;
    }
    if (!cs->hp.fits) {
        if (!make_large_plan(P, E.num_sms, cs->lp)) { last_error() = "grid too large for the whole-GPU FFT product"; delete cs; return -34; }
        CB_CUDA(cudaMalloc(&cs->d_T, sizeof(cd) * (size_t) (P.Lx + 1) * cs->lp.ldT));
        CB_CUDA(cudaMalloc(&cs->d_gpart, sizeof(double) * 2 * 8 * (size_t) E.num_sms));
    }

    const long nblk = 4L * mx * my;
    CB_CUDA(cudaMalloc(&cs->d_cf[SET_CS], sizeof(double) * 9 * nblk));
    ElascfArgs a = { mat.ak, mat.nu, dx, dy, 0.0, 0.0, mx, my };
    k_elascf_pcwcns<<<grid1d(nblk, 128), 128, 0, st>>>(a, cs->d_cf[SET_CS]);
    E.launches++;
    if (is_roll) {
        // sgencr for rolling (m_visc.f90:310-359): cv = coefficients at the offset dq along the rolling direction for
        // the tangential displacements, the normal rows are those of cs
        CB_CUDA(cudaMalloc(&cs->d_cf[SET_CV], sizeof(double) * 9 * nblk));
        ElascfArgs av = { mat.ak, mat.nu, dx, dy, cos(chi) * dq, sin(chi) * dq, mx, my };
        k_elascf_pcwcns<<<grid1d(nblk, 128), 128, 0, st>>>(av, cs->d_cf[SET_CV]);
        E.launches++;
        for (int jk = 0; jk < 3; jk++)
            CB_CUDA(cudaMemcpyAsync(cs->d_cf[SET_CV] + (size_t) (jk * 3 + 2) * nblk, cs->d_cf[SET_CS] + (size_t) (jk * 3 + 2) * nblk,
                                    sizeof(double) * nblk, cudaMemcpyDeviceToDevice, st));
        CB_CUDA(cudaMalloc(&cs->d_cf[SET_CSV], sizeof(double) * 9 * nblk));
        k_csv_blocks<<<grid1d(9 * nblk, 256), 256, 0, st>>>(cs->d_cf[SET_CS], cs->d_cf[SET_CV], cs->d_cf[SET_CSV], nblk);
        E.launches++;
    }
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaStreamSynchronize(st));
    E.sets[key] = cs;
    E.by_handle.push_back(cs);
    *out = cs;
    return 0;
}

// ---- subsurface coefficients for a list of depths: built once per (grid, material, z-list), shared by all cases ----
inline int build_subsurf_chat(CoefSet &cs, const std::vector<double> &z, const double gg[2], const double poiss[2],
                              cudaStream_t st, const cd **out)
{
    Engine &E = engine();
    std::vector<double> key(z);
    key.push_back(gg[0]); key.push_back(gg[1]); key.push_back(poiss[0]); key.push_back(poiss[1]);
    auto it = cs.subs_chat.find(key);
    if (it != cs.subs_chat.end()) { *out = it->second; return 0; }
    const ConvPlan &P = cs.hp.p;
    const int nz = (int) z.size();
    const long nblk = 4L * cs.mx * cs.my;
    cd *chat = nullptr; double *cf = nullptr;
    const size_t clen = plan_chat_len(P);
    CB_CUDA(cudaMalloc(&chat, sizeof(cd) * (size_t) nz * 36 * clen));
    CB_CUDA(cudaMalloc(&cf, sizeof(double) * 36 * nblk));
    for (int iz = 0; iz < nz; iz++) {
        const int neg = z[iz] >= 0.0 ? 1 : -1, ia = neg > 0 ? 0 : 1;
        k_subsurf_coef<<<grid1d(nblk, 64), 64, 0, st>>>(cs.mx, cs.my, cs.key.dx, cs.key.dy, gg[ia], poiss[ia], z[iz], neg, cf);
        E.launches++;
        int rcb = launch_build_chat(P, cf, 36, cs.mx, cs.my, 1.0 / (4.0 * P.Fx * P.Fy), chat + (size_t) iz * 36 * clen, st);
        if (rcb) return rcb;
    }
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaStreamSynchronize(st));
    cudaFree(cf);
    cs.subs_chat[key] = chat;
    *out = chat;
    return 0;
}

inline int launch_blocks(int ncase) { Engine &E = engine(); return ncase < E.num_sms ? ncase : E.num_sms; }

}  // namespace cb200
