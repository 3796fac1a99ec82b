// api_lowlevel.cuh -- cb200_* entry points: kernel-level C-ABI (plain pointers and sizes).
#pragma once
#include "engine.cuh"
#include "aijpj.cuh"

namespace cb200 {

inline CoefSet *set_from_handle(int h)
{
    Engine &E = engine();
    std::lock_guard<std::mutex> lk(E.mu);
    if (h < 0 || h >= (int) E.by_handle.size()) { last_error() = "invalid coefficient-set handle"; return nullptr; }
    return E.by_handle[h];
}

inline int handle_of(CoefSet *cs)
{
    Engine &E = engine();
    std::lock_guard<std::mutex> lk(E.mu);
    for (size_t i = 0; i < E.by_handle.size(); i++) if (E.by_handle[i] == cs) return (int) i;
    return -1;
}

inline void dir_range(int arg, int &a0, int &a1)
{   // m_aijpj.f90:292-336
    if (arg == -3) { a0 = 1; a1 = 3; } else if (arg == -2) { a0 = 1; a1 = 2; }
    else if (arg >= 1 && arg <= 3) { a0 = a1 = arg; } else { a0 = 1; a1 = 0; }
}

__global__ void k_fill_masked(double *u, const int *el, int mask_mode, double val, long n)
{
    const long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
    if (t < n && (mask_mode == 0 || el[t] >= 1)) u[t] = val;
}

// Strided batched product: case ic reads p + ic*pstride, writes u + ic*ustride, mask el + ic*npot
__global__ void __launch_bounds__(CB_THREADS, 1)
k_conv_batch_strided(const __grid_constant__ ConvPlan P, const double *p, long pstride, const cd *chat, double *u, long ustride,
                     const int *el, int mask_mode, int add, int ncase)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = smem_view(P, smem_raw);
    smem_load_tables(P, sm);
    for (int ic = blockIdx.x; ic < ncase; ic += gridDim.x)
        conv_dev(P, sm, p + ic * pstride, chat, u + ic * ustride, el ? el + (size_t) ic * P.npot : nullptr,
                 el ? mask_mode : 0, add);
}

// one single-block product on a grid beyond one CTA: three grid-wide phases (rows, columns x C^, inverse rows)
inline int conv_large_launch(CoefSet &cs, const double *d_p, const cd *chat, double *d_u, const int *d_el, int mask_mode,
                             int add, cudaStream_t st)
{
    Engine &E = engine();
    const LargePlan &L = cs.lp;
    RowSrc src;
    src.base = d_p; src.kind = 0; src.mx = L.P.mx; src.my = L.P.my; src.cmx = 0; src.cmy = 0; src.Fx = L.P.Fx; src.Fy = L.P.Fy; src.row0 = 0; src.stride = 0;
    if (!cs.ev_l0) { CB_CUDA(cudaEventCreate(&cs.ev_l0)); CB_CUDA(cudaEventCreate(&cs.ev_l1)); }
    CB_CUDA(cudaEventRecord(cs.ev_l0, st));
    k_lg_rows_fwd<<<L.ntr, CB_THREADS, L.smem_bytes, st>>>(L, src, L.P.my, L.RB, cs.d_T, L.ldT);
    k_lg_cols<<<L.ntc, CB_THREADS, L.smem_bytes, st>>>(L, L.P.my, L.P.my, cs.d_T, L.ldT, chat, nullptr, 1.0);
    k_lg_rows_inv<<<L.ntr, CB_THREADS, L.smem_bytes, st>>>(L, cs.d_T, d_u, d_el, mask_mode, add);
    CB_CUDA(cudaEventRecord(cs.ev_l1, st));
    E.launches += 3;
    return 0;
}

// gf3_VecAijPj semantics (m_aijpj.f90:346-395) on device buffers laid out [ncase][3][npot]
inline int vecaijpj_dev(CoefSet &cs, int set, int ncase, int iigs, int ikarg, int jkarg, const double *d_p,
                        const int *d_el, double *d_u, cudaStream_t st)
{
    Engine &E = engine();
    const ConvPlan &P = cs.hp.p;
    if (iigs != -9 && iigs != -8) { last_error() = "FFT product allowed only for AllElm or AllInt"; return -99; }
    if (!cs.d_cf[set]) { last_error() = "coefficient set not available"; return -99; }
    const int mask_mode = (iigs == -8) ? 1 : 0;
    if (mask_mode == 1 && !d_el) { last_error() = "AllInt product needs an element division"; return -99; }
    static bool attr[CB_MAX_DEVICES] = { false };               // function attributes are per device
    if (!attr[current_device()]) { CB_CUDA(cudaFuncSetAttribute(k_conv_batch_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)); attr[current_device()] = true; }
    int ik0, ik1, jk0, jk1;
    dir_range(ikarg, ik0, ik1); dir_range(jkarg, jk0, jk1);
    const long cstride = 3L * P.npot;
    const int nblk = launch_blocks(ncase);
    for (int ik = ik0; ik <= ik1; ik++) {
        bool ladd = false;
        for (int jk = jk0; jk <= jk1; jk++) {
            if (!cs.nt_cpl && (ik * jk == 3 || ik * jk == 6)) continue;     // :358-369
            int rc = build_chat(cs, set, ik, jk, st);
            if (rc) return rc;
            if (!cs.hp.fits) {
                for (int ic = 0; ic < ncase && !rc; ic++)
                    rc = conv_large_launch(cs, d_p + ic * cstride + (size_t) (jk - 1) * P.npot, cs.d_chat[set][ik - 1][jk - 1],
                                           d_u + ic * cstride + (size_t) (ik - 1) * P.npot,
                                           d_el ? d_el + (size_t) ic * P.npot : nullptr, mask_mode, ladd ? 1 : 0, st);
                if (rc) return rc;
                ladd = true;
                continue;
            }
            k_conv_batch_strided<<<nblk, CB_THREADS, P.smem_bytes, st>>>(
                P, d_p + (size_t) (jk - 1) * P.npot, cstride, cs.d_chat[set][ik - 1][jk - 1],
                d_u + (size_t) (ik - 1) * P.npot, cstride, d_el, mask_mode, ladd ? 1 : 0, ncase);
            E.launches++;
            ladd = true;
        }
        if (!ladd) {                                                        // all blocks skipped: u = 0 on the selection
            for (int ic = 0; ic < ncase; ic++) {
                k_fill_masked<<<grid1d(P.npot, 256), 256, 0, st>>>(d_u + ic * cstride + (size_t) (ik - 1) * P.npot,
                                                                  d_el ? d_el + (size_t) ic * P.npot : nullptr,
                                                                  mask_mode, 0.0, P.npot);
                E.launches++;
            }
        }
    }
    CB_CUDA(cudaGetLastError());
    return 0;
}

struct NormBatch {            // persistent device workspace of the batched NORM solve
    int cap = 0;
    NormCase *d_cases = nullptr;
    double *d_work = nullptr;
    int *d_next = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;     // bracket the solver kernel alone (roofline timing)
    // staging for the host-buffer entry point
    long stage_cap = 0;
    int scal_cap = 0;
    double *s_hs = nullptr, *s_pn = nullptr, *s_un = nullptr, *s_scal = nullptr;
    int *s_el = nullptr;
    cudaStream_t pipe[3] = { nullptr, nullptr, nullptr };   // host-buffer entry point: chunks of the batch alternate over these
};
#define CB_NEXT_SLOTS 64

__global__ void k_norm_pack(NormCase *cases, int ncase, int npot, const double *hs, int *el, double *pn, double *un,
                            const double *scal, double *work, NormCase proto)
{
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= ncase) return;
    NormCase c = proto;
    c.hs = hs + (size_t) ic * npot;
    c.el = el + (size_t) ic * npot;
    c.pn = pn + (size_t) ic * npot;
    c.work = work + (size_t) ic * 9 * npot;
    c.pen = scal[ic * 8 + 0];
    c.fntrue = scal[ic * 8 + 1];
    c.ptx = nullptr; c.pty = nullptr;
    (void) un;
    cases[ic] = c;
}

__global__ void k_norm_unpack(const NormCase *cases, int ncase, double *scal)
{
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= ncase) return;
    const NormCase &c = cases[ic];
    double *s = scal + ic * 8;
    s[0] = c.pen; s[1] = c.fntrue; s[2] = c.itcg; s[3] = c.itnorm; s[4] = c.ncon; s[5] = c.status; s[6] = c.err; s[7] = c.nprod;
}

// u_n = A_zz p_n on the contact area after the solve (soutpt, m_soutpt.f90:378-385), batched
inline NormBatch &norm_batch() { static NormBatch b[CB_MAX_DEVICES]; return b[current_device()]; }

// grids beyond one CTA: one cooperative whole-GPU launch per case (cases run one after the other)
inline int snorm_large_dev(CoefSet &cs, int ncase, NormCase proto, const double *d_hs, int *d_el, double *d_pn,
                           double *d_un, double *d_scal, cudaStream_t st)
{
    Engine &E = engine();
    NormBatch &B = norm_batch();
    const LargePlan &L = cs.lp;
    const int npot = L.P.npot;
    LargeCtx X;
    X.L = L; X.T = cs.d_T; X.gpart = cs.d_gpart; X.prof = nullptr;
    k_norm_pack<<<grid1d(ncase, 128), 128, 0, st>>>(B.d_cases, ncase, npot, d_hs, d_el, d_pn, d_un, d_scal, B.d_work, proto);
    E.launches++;
    CB_CUDA(cudaEventRecord(B.ev0, st));
    for (int ic = 0; ic < ncase; ic++) {
        NormCase *cp = B.d_cases + ic;
        double *un = d_un ? d_un + (size_t) ic * npot : nullptr;
        if (un) CB_CUDA(cudaMemsetAsync(un, 0, sizeof(double) * npot, st));
        void *args[] = { (void *) &X, (void *) &cp, (void *) &un };
        CB_CUDA(cudaLaunchCooperativeKernel((void *) k_lg_snorm, dim3(E.num_sms), dim3(CB_THREADS), args, (size_t) L.smem_bytes, st));
        E.launches++;
    }
    CB_CUDA(cudaEventRecord(B.ev1, st));
    k_norm_unpack<<<grid1d(ncase, 128), 128, 0, st>>>(B.d_cases, ncase, d_scal);
    E.launches++;
    CB_CUDA(cudaGetLastError());
    return 0;
}

// c0 / total / slot: this call handles cases c0 .. c0+ncase-1 of a batch of `total` cases whose chunks are in flight on
// different streams (host-buffer entry point); the workspace is sized for the whole batch, `slot` selects the work queue
inline int snorm_batch_dev(CoefSet &cs, int ncase, int ic_norm, int maxgs, int maxin, double eps, const double *d_hs,
                           int *d_el, double *d_pn, double *d_un, double *d_scal, cudaStream_t st, int c0 = 0, int total = 0,
                           int slot = 0)
{
    Engine &E = engine();
    const ConvPlan &P = cs.hp.p;
    int rc;
    if ((rc = build_prec(cs, st))) return rc;
    if ((rc = build_chat(cs, SET_CS, 3, 3, st))) return rc;
    if ((rc = build_chat(cs, SET_MS, 3, 3, st))) return rc;
    if ((rc = build_levels(cs, st))) return rc;
    NormBatch &B = norm_batch();
    if (!B.d_next) CB_CUDA(cudaMalloc(&B.d_next, sizeof(int) * CB_NEXT_SLOTS));
    if (!B.ev0) { CB_CUDA(cudaEventCreate(&B.ev0)); CB_CUDA(cudaEventCreate(&B.ev1)); }
    static long cap_bytes_dev[CB_MAX_DEVICES] = { 0 };
    long &cap_bytes = cap_bytes_dev[current_device()];
    const int ntot = total > ncase ? total : ncase;
    const long need = (long) ntot * 9 * P.npot * sizeof(double);
    if (ntot > B.cap || need > cap_bytes) {
        if (c0 > 0) { last_error() = "snorm_batch_dev: workspace must be reserved by the first chunk"; return -99; }
        CB_CUDA(cudaDeviceSynchronize());
        if (B.d_cases) cudaFree(B.d_cases);
        if (B.d_work) cudaFree(B.d_work);
        CB_CUDA(cudaMalloc(&B.d_cases, sizeof(NormCase) * ntot));
        CB_CUDA(cudaMalloc(&B.d_work, need));
        B.cap = ntot; cap_bytes = need;
    }
    NormCase *d_cases = B.d_cases + c0;
    double *d_work = B.d_work + (size_t) c0 * 9 * P.npot;
    int *d_next = B.d_next + (slot % CB_NEXT_SLOTS);
    NormCase proto;
    memset(&proto, 0, sizeof(proto));
    proto.chatA = cs.d_chat[SET_CS][2][2];
    proto.chatM = cs.d_chat[SET_MS][2][2];
    proto.chatA31 = nullptr; proto.chatA32 = nullptr;
    proto.cf33 = cs.d_cf[SET_CS] + (size_t) 8 * 4 * cs.mx * cs.my;
    proto.cmx = cs.mx; proto.cmy = cs.my;
    proto.ga_inv = cs.ga_inv;
    proto.ic_norm = ic_norm; proto.maxgs = maxgs; proto.maxin = maxin; proto.eps = eps;
    proto.dxdy = cs.key.dx * cs.key.dy;
    proto.lev = cs.d_lev; proto.nlx = cs.nlx; proto.nly = cs.nly; proto.stage_bytes = cs.stage_bytes();
    if (!cs.hp.fits) return snorm_large_dev(cs, ncase, proto, d_hs, d_el, d_pn, d_un, d_scal, st);
    k_norm_pack<<<grid1d(ncase, 128), 128, 0, st>>>(d_cases, ncase, P.npot, d_hs, d_el, d_pn, d_un, d_scal, d_work, proto);
    CB_CUDA(cudaMemsetAsync(d_next, 0, sizeof(int), st));
    if (c0 == 0) CB_CUDA(cudaEventRecord(B.ev0, st));
    k_snorm_batch<<<launch_blocks(ncase), CB_THREADS, P.smem_bytes, st>>>(P, d_cases, ncase, d_next);
    if (total <= ncase || c0 + ncase >= total) CB_CUDA(cudaEventRecord(B.ev1, st));
    k_norm_unpack<<<grid1d(ncase, 128), 128, 0, st>>>(d_cases, ncase, d_scal);
    E.launches += 3;
    if (d_un) {
        static bool attr[CB_MAX_DEVICES] = { false };
        if (!attr[current_device()]) { CB_CUDA(cudaFuncSetAttribute(k_conv_batch_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)); attr[current_device()] = true; }
        CB_CUDA(cudaMemsetAsync(d_un, 0, sizeof(double) * (size_t) ncase * P.npot, st));
        k_conv_batch_strided<<<launch_blocks(ncase), CB_THREADS, P.smem_bytes, st>>>(
            P, d_pn, P.npot, proto.chatA, d_un, P.npot, d_el, 1, 0, ncase);
        E.launches++;
    }
    CB_CUDA(cudaGetLastError());
    return 0;
}

// ---- batched subsurface evaluation on device buffers ----
struct SubsBatch { double *d_vr = nullptr; long cap = 0; int *d_next = nullptr; double *s_ps = nullptr, *s_t = nullptr; long cap_ps = 0, cap_t = 0; };
inline SubsBatch &subs_batch() { static SubsBatch b[CB_MAX_DEVICES]; return b[current_device()]; }

inline int subsurf_batch_dev(CoefSet &cs, int ncase, int nz, const double *z, const double gg[2], const double poiss[2],
                             const double *d_ps, double *d_table, cudaStream_t st)
{
    Engine &E = engine();
    const ConvPlan &P = cs.hp.p;
    if (!cs.hp.fits) { last_error() = "grid too large for the single-CTA product"; return -34; }
    if (nz > 31) { last_error() = "at most 31 depths per subsurface block"; return -99; }
    std::vector<double> zz(z, z + nz);
    const cd *chat = nullptr;
    int rc = build_subsurf_chat(cs, zz, gg, poiss, st, &chat);
    if (rc) return rc;
    SubsBatch &B = subs_batch();
    const int nblk = launch_blocks(ncase * nz);
    const long need = (long) nblk * 13 * P.npot;
    if (!B.d_next) CB_CUDA(cudaMalloc(&B.d_next, sizeof(int)));
    if (need > B.cap) { if (B.d_vr) cudaFree(B.d_vr); CB_CUDA(cudaMalloc(&B.d_vr, sizeof(double) * need)); B.cap = need; }
    SubsArgs A;
    A.ncase = ncase; A.nz = nz; A.neg_mask = 0;
    for (int i = 0; i < nz; i++) if (z[i] < 0.0) A.neg_mask |= (1 << i);
    A.ps = d_ps; A.chat = chat; A.vr = B.d_vr; A.table = d_table;
    A.gg[0] = gg[0]; A.gg[1] = gg[1]; A.poiss[0] = poiss[0]; A.poiss[1] = poiss[1];
    A.next = B.d_next; A.chat_len = (long) plan_chat_len(P);
    CB_CUDA(cudaMemsetAsync(B.d_next, 0, sizeof(int), st));
    k_subsurf_batch<<<nblk, CB_THREADS, P.smem_bytes, st>>>(P, A);
    E.launches++;
    CB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

const char *cb200_last_error(void) { return last_error().c_str(); }
long cb200_num_launches(void) { return engine().launches; }
int cb200_conv_prof(unsigned long long *out, int reset)
{   // cycle counters of CTA 0: out[0] products, [1] cycles inside the fused product, [2] cycles of k_snorm_batch
    int rc = engine_init();
    if (rc) return rc;
    CB_CUDA(cudaMemcpyFromSymbol(out, g_conv_prof, sizeof(unsigned long long) * 4));
    if (reset) { unsigned long long z[4] = { 0 }; CB_CUDA(cudaMemcpyToSymbol(g_conv_prof, z, sizeof(z))); }
    return 0;
}

// work counters of all CTAs since the last reset: out[0] products, [1] nominal flops of the products at the transform sizes used,
// [2] algorithmic bytes of the products (17 per box element), [3] Gauss-Seidel row-sum units ncon (ncon + 2 my) summed over sweeps
int cb200_work_counters(unsigned long long *out, int reset)
{
    int rc = engine_init();
    if (rc) return rc;
    CB_CUDA(cudaMemcpyFromSymbol(out, g_work, sizeof(unsigned long long) * 4));
    if (reset) { unsigned long long z[4] = { 0 }; CB_CUDA(cudaMemcpyToSymbol(g_work, z, sizeof(z))); }
    return 0;
}

// all 32 counters (slots 4..31: sections of the solver kernels of CTA 0, see CB_T)
int cb200_solver_prof(unsigned long long *out, int reset)
{
    int rc = engine_init();
    if (rc) return rc;
    CB_CUDA(cudaMemcpyFromSymbol(out, g_conv_prof, sizeof(unsigned long long) * 32));
    if (reset) { unsigned long long z[32] = { 0 }; CB_CUDA(cudaMemcpyToSymbol(g_conv_prof, z, sizeof(z))); }
    return 0;
}
int cb200_steady_prof(unsigned long long *out, int reset)
{   // cycle counters of the SteadyGS step, summed over all CTAs since the last reset (see steady_solver.cuh)
    int rc = engine_init();
    if (rc) return rc;
    CB_CUDA(cudaMemcpyFromSymbol(out, g_steady_prof, sizeof(unsigned long long) * 8));
    if (reset) { unsigned long long z[8] = { 0 }; CB_CUDA(cudaMemcpyToSymbol(g_steady_prof, z, sizeof(z))); }
    return 0;
}
int cb200_gd_prof(unsigned long long *out, int reset)
{   // cycle counters of GDsteady (leader thread of every solver call) since the last reset (see gdsteady_solver.cuh)
    int rc = engine_init();
    if (rc) return rc;
    CB_CUDA(cudaMemcpyFromSymbol(out, g_gd_prof, sizeof(unsigned long long) * 12));
    if (reset) { unsigned long long z[12] = { 0 }; CB_CUDA(cudaMemcpyToSymbol(g_gd_prof, z, sizeof(z))); }
    return 0;
}
int cb200_num_sms(void) { int rc = engine_init(); return rc ? rc : engine().num_sms; }
int cb200_opt_fft_size(int n) { return opt_fft_size(n); }

int cb200_coefset_create(int mx, int my, double dx, double dy, double gg1, double gg2, double poiss1, double poiss2,
                         int is_roll, double chi, double dq)
{
    if (mx < 1 || my < 1 || dx <= 0 || dy <= 0) { last_error() = "invalid grid"; return -34; }
    Material m = { { gg1, gg2 }, { poiss1, poiss2 }, 0, 0, 0 };
    CoefSet *cs = nullptr;
    int rc = get_coefset(mx, my, dx, dy, m, is_roll, chi, dq, 0, &cs);
    if (rc) return rc;
    return handle_of(cs);
}

int cb200_coefset_plan(int handle, int *out)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    const ConvPlan &P = cs->hp.p;
    out[0] = P.Fx; out[1] = P.Fy; out[2] = P.C; out[3] = P.nchunk; out[4] = P.smem_bytes; out[5] = cs->hp.fits;
    out[6] = P.nsx; out[7] = P.nsy;
    return 0;
}

int cb200_coefset_get_block(int handle, int set, int ik, int jk, double *out)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    if (set < 0 || set > 3 || ik < 1 || ik > 3 || jk < 1 || jk > 3) { last_error() = "invalid block"; return -99; }
    if (set == SET_MS) { int rc = build_prec(*cs, 0, ik); if (rc) return rc; }
    if (!cs->d_cf[set]) { last_error() = "coefficient set not available"; return -99; }
    const size_t nblk = (size_t) 4 * cs->mx * cs->my;
    CB_CUDA(cudaMemcpy(out, cs->d_cf[set] + ((jk - 1) * 3 + (ik - 1)) * nblk, sizeof(double) * nblk, cudaMemcpyDeviceToHost));
    return 0;
}

int cb200_vecaijpj_dev(int handle, int set, int ncase, int iigs, int ikarg, int jkarg, const double *d_p,
                       const int *d_el, double *d_u, void *stream)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    if (set == SET_MS) { int rc = build_prec(*cs, (cudaStream_t) stream); if (rc) return rc; }
    return vecaijpj_dev(*cs, set, ncase, iigs, ikarg, jkarg, d_p, d_el, d_u, (cudaStream_t) stream);
}

int cb200_vecaijpj(int handle, int set, int ncase, int iigs, int ikarg, int jkarg, const double *p, const int *el,
                   double *u)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    const size_t n3 = (size_t) ncase * 3 * cs->hp.p.npot, n1 = (size_t) ncase * cs->hp.p.npot;
    double *d_p = nullptr, *d_u = nullptr; int *d_el = nullptr;
    CB_CUDA(cudaMalloc(&d_p, sizeof(double) * n3));
    CB_CUDA(cudaMalloc(&d_u, sizeof(double) * n3));
    CB_CUDA(cudaMemcpy(d_p, p, sizeof(double) * n3, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_u, u, sizeof(double) * n3, cudaMemcpyHostToDevice));   // unselected elements keep their value
    if (el) { CB_CUDA(cudaMalloc(&d_el, sizeof(int) * n1)); CB_CUDA(cudaMemcpy(d_el, el, sizeof(int) * n1, cudaMemcpyHostToDevice)); }
    int rc = cb200_vecaijpj_dev(handle, set, ncase, iigs, ikarg, jkarg, d_p, d_el, d_u, nullptr);
    if (!rc) { cudaError_t e = cudaMemcpy(u, d_u, sizeof(double) * n3, cudaMemcpyDeviceToHost); if (e != cudaSuccess) { last_error() = cudaGetErrorString(e); rc = -99; } }
    cudaFree(d_p); cudaFree(d_u); if (d_el) cudaFree(d_el);
    return rc;
}

// gf3_AijPj (m_aijpj.f90:99-254) for npts elements of one case: out[k] = displacement in direction ik of element ii[k] (0-based)
// due to the tractions p [3][npot] in the directions jkarg (1..3, -2 tangential, -3 all), with the coefficient set `set`;
// the n-t blocks are skipped when the materials are similar (nt_cpl false), as in the reference (:121-141).  Host arrays.
int cb200_aijpj(int handle, int set, int ik, int jkarg, int npts, const int *ii, const double *p, const int *el, double *out)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    if (set < 0 || set > 3 || !cs->d_cf[set]) { last_error() = "coefficient set not available"; return -99; }
    if (ik < 1 || ik > 3) { last_error() = "invalid direction ik"; return -39; }
    if (npts < 1) return 0;
    int jk0, jk1;
    dir_range(jkarg, jk0, jk1);
    if (!cs->nt_cpl) {
        if (jkarg == -3) { if (ik <= 2) jk1 = 2; else jk0 = 3; }
        else if (jkarg == -2) { if (ik == 3) jk1 = 0; }
        else if (jk0 <= jk1) { if (ik <= 2) jk1 = std::min(2, jk1); else jk0 = 3; }
    }
    const int mx = cs->mx, my = cs->my, npot = mx * my;
    for (int k = 0; k < npts; k++) if (ii[k] < 0 || ii[k] >= npot) { last_error() = "cb200_aijpj: element index out of range"; return -39; }
    double *d_p = nullptr, *d_out = nullptr; int *d_el = nullptr, *d_ii = nullptr;
    CB_CUDA(cudaMalloc(&d_p, sizeof(double) * 3 * npot)); CB_CUDA(cudaMalloc(&d_out, sizeof(double) * npts));
    CB_CUDA(cudaMalloc(&d_el, sizeof(int) * npot)); CB_CUDA(cudaMalloc(&d_ii, sizeof(int) * npts));
    cudaMemcpy(d_p, p, sizeof(double) * 3 * npot, cudaMemcpyHostToDevice);
    cudaMemcpy(d_el, el, sizeof(int) * npot, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ii, ii, sizeof(int) * npts, cudaMemcpyHostToDevice);
    const int blocks = std::min(npts, 8 * std::max(1, engine().num_sms));
    k_aijpj<<<blocks, 256, sizeof(double) * 3 * my>>>(mx, my, cs->d_cf[set], cs->ga_inv, ik, jk0, jk1, d_p, d_el, d_ii, npts, d_out);
    engine().launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(double) * npts, cudaMemcpyDeviceToHost);
    cudaFree(d_p); cudaFree(d_out); cudaFree(d_el); cudaFree(d_ii);
    if (e != cudaSuccess) { last_error() = std::string("k_aijpj: ") + cudaGetErrorString(e); return -99; }
    return 0;
}

int cb200_snorm_batch_dev(int handle, int ncase, int ic_norm, int maxgs, int maxin, double eps, const double *d_hs,
                          int *d_el, double *d_pn, double *d_un, double *d_scal, void *stream)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    if (ncase < 1) return 0;
    return snorm_batch_dev(*cs, ncase, ic_norm, maxgs, maxin, eps, d_hs, d_el, d_pn, d_un, d_scal, (cudaStream_t) stream);
}

// host-buffer variant: copies in, solves, copies out (the e2e path of bench.py and of cntc_calculate_batch)
int cb200_snorm_batch(int handle, int ncase, int ic_norm, int maxgs, int maxin, double eps, const double *hs,
                      int *el, double *pn, double *un, double *scal)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    if (ncase < 1) return 0;
    NormBatch &B = norm_batch();
    const long n = (long) ncase * cs->hp.p.npot;
    if (n > B.stage_cap) {
        cudaFree(B.s_hs); cudaFree(B.s_pn); cudaFree(B.s_un); cudaFree(B.s_el);
        B.stage_cap = 0;
        CB_CUDA(cudaMalloc(&B.s_hs, sizeof(double) * n));
        CB_CUDA(cudaMalloc(&B.s_pn, sizeof(double) * n));
        CB_CUDA(cudaMalloc(&B.s_un, sizeof(double) * n));
        CB_CUDA(cudaMalloc(&B.s_el, sizeof(int) * n));
        B.stage_cap = n;
    }
    if (ncase > B.scal_cap) {                       // per-case scalars: their own capacity (many cases on a small grid)
        cudaFree(B.s_scal);
        B.scal_cap = 0;
        CB_CUDA(cudaMalloc(&B.s_scal, sizeof(double) * 8 * (size_t) ncase));
        B.scal_cap = ncase;
    }
    // The batch is cut into chunks (multiples of the SM count: one CTA per case) that alternate over three streams, so
    // that the host->device copy of chunk k+1 and the device->host copy of chunk k-1 run on the copy engines while the
    // solver kernel of chunk k occupies the SMs.  Pageable host buffers still work (the copies then serialise).
    Engine &E = engine();
    const int npot = cs->hp.p.npot, nsm = E.num_sms > 0 ? E.num_sms : 1;
    int csz = ncase;
    if (cs->hp.fits && ncase >= 4 * nsm) csz = ((ncase / 4 + nsm - 1) / nsm) * nsm;        // about four chunks
    const int nchunk = (ncase + csz - 1) / csz;
    if (nchunk > 1) {
        for (int k = 0; k < 3; k++) if (!B.pipe[k]) CB_CUDA(cudaStreamCreateWithFlags(&B.pipe[k], cudaStreamNonBlocking));
        // coefficient transforms (cached per grid and material) are built once, before the chunks fan out over the streams
        int rc0 = build_prec(*cs, 0);
        if (!rc0) rc0 = build_chat(*cs, SET_CS, 3, 3, 0);
        if (!rc0) rc0 = build_chat(*cs, SET_MS, 3, 3, 0);
        if (!rc0) rc0 = build_levels(*cs, 0);
        if (rc0) return rc0;
        CB_CUDA(cudaStreamSynchronize(0));
    }
    for (int ch = 0; ch < nchunk; ch++) {
        const int c0 = ch * csz, nc = (ncase - c0 < csz) ? ncase - c0 : csz;
        const size_t o = (size_t) c0 * npot, m = (size_t) nc * npot;
        cudaStream_t st = nchunk > 1 ? B.pipe[ch % 3] : (cudaStream_t) 0;
        CB_CUDA(cudaMemcpyAsync(B.s_hs + o, hs + o, sizeof(double) * m, cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(B.s_pn + o, pn + o, sizeof(double) * m, cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(B.s_el + o, el + o, sizeof(int) * m, cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(B.s_scal + (size_t) c0 * 8, scal + (size_t) c0 * 8, sizeof(double) * 8 * nc, cudaMemcpyHostToDevice, st));
        int rc = snorm_batch_dev(*cs, nc, ic_norm, maxgs, maxin, eps, B.s_hs + o, B.s_el + o, B.s_pn + o, un ? B.s_un + o : nullptr,
                                 B.s_scal + (size_t) c0 * 8, st, c0, ncase, ch);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        CB_CUDA(cudaMemcpyAsync(pn + o, B.s_pn + o, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        CB_CUDA(cudaMemcpyAsync(el + o, B.s_el + o, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
        if (un) CB_CUDA(cudaMemcpyAsync(un + o, B.s_un + o, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        CB_CUDA(cudaMemcpyAsync(scal + (size_t) c0 * 8, B.s_scal + (size_t) c0 * 8, sizeof(double) * 8 * nc, cudaMemcpyDeviceToHost, st));
    }
    if (nchunk > 1) { for (int k = 0; k < 3; k++) CB_CUDA(cudaStreamSynchronize(B.pipe[k])); }
    else CB_CUDA(cudaStreamSynchronize(0));
    return 0;
}

// device time of the most recent solver kernel(s) alone (k_snorm_batch, k_contac_batch, k_lg_*), ms; synchronises on
// the end event
double cb200_snorm_kernel_ms(void)
{
    NormBatch &B = norm_batch();
    if (!B.ev1) return -1.0;
    if (cudaEventSynchronize(B.ev1) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, B.ev0, B.ev1) != cudaSuccess) return -1.0;
    return (double) ms;
}

// measured FP64 FMA throughput of the device in TFLOP/s (2 flops per FMA), best of `reps`
double cb200_fp64_peak_tflops(int reps)
{
    if (engine_init()) return -1.0;
    Engine &E = engine();
    const int blocks = E.num_sms * 8, threads = 256, iters = 1 << 15;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(e0, 0);
        k_fp64_peak<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        E.launches++;
        const double tf = 2.0 * 8.0 * (double) iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return best;
}

// Subsurface block of type ISUBS 1/5: all elements x nz depths, for ncase traction fields (device buffers).
int cb200_subsurf_batch_dev(int handle, int ncase, int nz, const double *z, double gg1, double gg2, double poiss1,
                            double poiss2, const double *d_ps, double *d_table, void *stream)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    if (ncase < 1 || nz < 1) return 0;
    const double gg[2] = { gg1, gg2 }, poiss[2] = { poiss1, poiss2 };
    return subsurf_batch_dev(*cs, ncase, nz, z, gg, poiss, d_ps, d_table, (cudaStream_t) stream);
}

// same with HOST buffers: ps [ncase][3][npot] -> table [ncase][nz][npot][18]
int cb200_subsurf_batch(int handle, int ncase, int nz, const double *z, double gg1, double gg2, double poiss1,
                        double poiss2, const double *ps, double *table)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    if (ncase < 1 || nz < 1) return 0;
    const size_t n3 = (size_t) ncase * 3 * cs->hp.p.npot, nt = (size_t) ncase * nz * cs->hp.p.npot * 18;
    // staging buffers kept between calls (a sequence of cases asks for the same block after every case: cudaMalloc / cudaFree per
    // call cost far more than the kernel once the process holds gigabytes of work space)
    SubsBatch &B = subs_batch();
    if ((long) n3 > B.cap_ps) { if (B.s_ps) cudaFree(B.s_ps); B.s_ps = nullptr; B.cap_ps = 0; CB_CUDA(cudaMalloc(&B.s_ps, sizeof(double) * n3)); B.cap_ps = (long) n3; }
    if ((long) nt > B.cap_t) { if (B.s_t) cudaFree(B.s_t); B.s_t = nullptr; B.cap_t = 0; CB_CUDA(cudaMalloc(&B.s_t, sizeof(double) * nt)); B.cap_t = (long) nt; }
    double *d_ps = B.s_ps, *d_t = B.s_t;
    CB_CUDA(cudaMemcpy(d_ps, ps, sizeof(double) * n3, cudaMemcpyHostToDevice));
    int rc = cb200_subsurf_batch_dev(handle, ncase, nz, z, gg1, gg2, poiss1, poiss2, d_ps, d_t, nullptr);
    if (!rc) { cudaError_t e = cudaMemcpy(table, d_t, sizeof(double) * nt, cudaMemcpyDeviceToHost); if (e != cudaSuccess) { last_error() = cudaGetErrorString(e); rc = -99; } }
    return rc;
}

// ISUBS 9: direct evaluation in npoint arbitrary points xyz[npoint][3] for one traction field (HOST buffers);
// grid given by first element centre (xc1, yc1) and spacing; table [npoint][18]
int cb200_subsurf_points(int mx, int my, double xc1, double yc1, double dx, double dy, double gg1, double gg2,
                         double poiss1, double poiss2, const double *ps, int npoint, const double *xyz, double *table)
{
    int rc = engine_init();
    if (rc) return rc;
    const size_t n3 = (size_t) 3 * mx * my;
    double *d_ps = nullptr, *d_x = nullptr, *d_t = nullptr;
    CB_CUDA(cudaMalloc(&d_ps, sizeof(double) * n3));
    CB_CUDA(cudaMalloc(&d_x, sizeof(double) * 3 * npoint));
    CB_CUDA(cudaMalloc(&d_t, sizeof(double) * 18 * npoint));
    CB_CUDA(cudaMemcpy(d_ps, ps, sizeof(double) * n3, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_x, xyz, sizeof(double) * 3 * npoint, cudaMemcpyHostToDevice));
    k_subsurf_points<<<grid1d(npoint, 64), 64>>>(mx, my, xc1, yc1, dx, dy, gg1, gg2, poiss1, poiss2, d_ps, d_x, npoint, d_t);
    engine().launches++;
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaMemcpy(table, d_t, sizeof(double) * 18 * npoint, cudaMemcpyDeviceToHost));
    cudaFree(d_ps); cudaFree(d_x); cudaFree(d_t);
    return 0;
}

long cb200_snorm_workspace_bytes(int handle, int ncase)
{
    CoefSet *cs = set_from_handle(handle);
    if (!cs) return -99;
    return (long) ncase * (9L * cs->hp.p.npot * sizeof(double) + sizeof(NormCase));
}

}  // extern "C"
