/*
 * contact_addon_b200.h -- C-ABI of the B200-native CONTACT hot path (libcontact_addon_b200.so).
 *
 * Part 1 re-declares, with identical names, argument lists and calling convention (every scalar by pointer, arrays
 * with explicit lengths, all void), the entry points of the reference library that lie on the hot path and that a
 * caller needs to drive it.  Canonical prototypes: /root/reference/matlab_intfc/contact_addon.h:7-148; Fortran
 * implementations: /root/reference/src/contact_addon.f90 (line numbers cited per function).
 *
 * Part 2 (cntc_calculate_batch, cb200_*) is new: the batched case scheduler behind cntc_calculate and the
 * kernel-level entry points (influence product, NORM solve, coefficient access) used by the parity tests and the
 * benchmark.  They take plain pointers and sizes only.
 *
 * There is no CPU fallback: every entry point that computes needs a CUDA device and reports ierror = -99 otherwise.
 */
#ifndef CONTACT_ADDON_B200_H
#define CONTACT_ADDON_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------------
 * Part 1: the reference's cntc_* / subs_* interface (module 3, contact problems on a given grid)
 * ---------------------------------------------------------------------------------------------------------- */

/* contact_addon.f90:235-442 */
void cntc_initializefirst(int *ifcver, int *ierror, int *ioutput, const char *c_wrkdir, const char *c_outdir,
                          const char *c_expnam, int *len_wrkdir, int *len_outdir, int *len_expnam);
/* contact_addon.f90:446-472 */
void cntc_initializefirst_new(int *ifcver, int *ierror, int *ioutput, const char *c_wrkdir, const char *c_outdir,
                          const char *c_expnam, int *len_wrkdir, int *len_outdir, int *len_expnam);
/* contact_addon.f90:476-576 */
void cntc_initialize(int *ire, int *imodul, int *ifcver, int *ierror, const char *c_outdir, int *len_outdir);
/* contact_addon.f90:580-725 */
void cntc_setglobalflags(int *lenflg, int *params, int *values);
/* contact_addon.f90:780-973 */
void cntc_setflags(int *ire, int *icp, int *lenflg, int *params, int *values);
/* contact_addon.f90:1055-1124 */
void cntc_setmetadata(int *ire, int *icp, int *lenmta, int *params, double *values);
/* contact_addon.f90:1128-1283 */
void cntc_setsolverflags(int *ire, int *icp, int *gdigit, int *nints, int *iparam, int *nreals, double *rparam);
/* contact_addon.f90:1287-1473 (M = 0: [nu1, nu2, g1, g2]) */
void cntc_setmaterialparameters(int *ire, int *icp, int *mdigit, int *nparam, double *rparam);
/* contact_addon.f90:1554-1583 */
void cntc_settimestep(int *ire, int *icp, double *dt);
/* contact_addon.f90:1587-1615 */
void cntc_setreferencevelocity(int *ire, int *icp, double *veloc);
/* contact_addon.f90:1619-1667 */
void cntc_setrollingstepsize(int *ire, int *icp, double *chi, double *dq);
/* contact_addon.f90:1671-1968 (L = 0: [fstat, fkin]) */
void cntc_setfrictionmethod(int *ire, int *icp, int *imeth, int *nparam, double *params);
/* contact_addon.f90:1972-2093 */
void cntc_sethertzcontact(int *ire, int *icp, int *ipotcn, int *nparam, double *rparam);
/* contact_addon.f90:2097-2257 */
void cntc_setpotcontact(int *ire, int *icp, int *ipotcn, int *nparam, double *rparam);
/* contact_addon.f90:2294-2323 */
void cntc_setpenetration(int *ire, int *icp, double *pen);
/* contact_addon.f90:2327-2357 */
void cntc_setnormalforce(int *ire, int *icp, double *fn);
/* contact_addon.f90:2361-2487 */
void cntc_setundeformeddistc(int *ire, int *icp, int *ibase, int *nparam, double *prmudf);
/* contact_addon.f90:2491-2530 */
void cntc_setcreepages(int *ire, int *icp, double *vx, double *vy, double *phi);
/* contact_addon.f90:2597-2650 */
void cntc_settangentialforces(int *ire, int *icp, double *fx, double *fy);
/* contact_addon.f90:3330-3515 */
void subs_addblock(int *ire, int *icp, int *iblk, int *isubs, int *nx, int *ny, int *nz, double *xparam,
                   double *yparam, double *zparam);
/* contact_addon.f90:3519-3552 -> cntc_calculate3 :3754-3900 -> contac (m_scontc.f90:37-216) */
void cntc_calculate(int *ire, int *icp, int *ierror);
/* contact_addon.f90:3904-3997 */
void subs_calculate(int *ire, int *icp, int *ierror);
/* contact_addon.f90:4001-4160 */
void cntc_getflags(int *ire, int *icp, int *nparam, int *iparam, int *values);
/* contact_addon.f90:5070-5099 */
void cntc_getnumelements(int *ire, int *icp, int *mx, int *my);
/* contact_addon.f90:5103-5140 */
void cntc_getgriddiscretization(int *ire, int *icp, double *dx, double *dy);
/* contact_addon.f90:5144-5187 */
void cntc_getpotcontact(int *ire, int *icp, int *lenarr, double *values);
/* contact_addon.f90:5191-5220 */
void cntc_getpenetration(int *ire, int *icp, double *pen);
/* contact_addon.f90:5224-5271 */
void cntc_getcreepages(int *ire, int *icp, double *vx, double *vy, double *phi);
/* contact_addon.f90:5275-5318 */
void cntc_getcontactforces(int *ire, int *icp, double *fn, double *tx, double *ty, double *mz);
/* contact_addon.f90:5460-5497 */
void cntc_getcontactpatchareas(int *ire, int *icp, double *carea, double *harea, double *sarea);
/* contact_addon.f90:5501-5552 */
void cntc_getelementdivision(int *ire, int *icp, int *lenarr, int *eldiv);
/* contact_addon.f90:5556-5587 */
void cntc_getmaximumpressure(int *ire, int *icp, double *pnmax);
/* contact_addon.f90:5591-5626 */
void cntc_getmaximumtraction(int *ire, int *icp, double *ptmax);
/* contact_addon.f90:5672-5823 */
void cntc_getfielddata(int *ire, int *icp, int *ifld, int *lenarr, double *fld);
/* contact_addon.f90:5827-5855 */
void cntc_gettractions(int *ire, int *icp, int *lenarr, double *pn, double *px, double *py);
/* contact_addon.f90:5859-5885 */
void cntc_getmicroslip(int *ire, int *icp, int *lenarr, double *sx, double *sy);
/* contact_addon.f90:5889-5915 */
void cntc_getdisplacements(int *ire, int *icp, int *lenarr, double *un, double *ux, double *uy);
/* contact_addon.f90:6011-6067 */
void cntc_getcalculationtime(int *ire, int *icp, double *tcpu, double *twall);
/* contact_addon.f90:6071-6109 */
void subs_getblocksize(int *ire, int *icp, int *iblk, int *nx, int *ny, int *nz);
/* contact_addon.f90:6113-6178 */
void subs_getresults(int *ire, int *icp, int *iblk, int *lenarr, int *ncol, int *icol, double *values);
/* ---- the remaining prototypes of matlab_intfc/contact_addon.h:7-148, so that a caller built against the reference's
 *      header links unchanged.  Module-3 getters are served; the wheel/rail (module 1) entry points are outside the
 *      hot-path scope: they record an error (cb200_last_error) and return. ---- */
/* contact_addon.f90:4164-4251: itask 1 kinematic constants, 2 material, 3 friction */
void cntc_getparameters(int *ire, int *icp, int *itask, int *lenarr, double *values);
/* contact_addon.h:100 */
void cntc_getreferencevelocity(int *ire, int *icp, double *veloc);
/* contact_addon.f90:5021-5066: a1, b1, aa, bb, rho, cp, scale, bneg, bpos, aob */
void cntc_gethertzcontact(int *ire, int *icp, int *lenarr, double *values);
/* contact_addon.h:126 (H = 0 on this path: both zero) */
void cntc_getmaximumtemperature(int *ire, int *icp, double *t1max, double *t2max);
/* contact_addon.f90:5919-6007: sens(lenout, lenin); filled: d(fx, fy)/d(cksi, ceta) of the Newton-Raphson process */
void cntc_getsensitivities(int *ire, int *icp, int *lenout, int *lenin, double *sens);
/* contact_addon.h:140 */
void cntc_resetcalculationtime(int *ire, int *icp);
/* contact_addon.h:55 (E = 9), :31 (H-digit): stored, the digits themselves are refused at cntc_calculate */
void cntc_setextrarigidslip(int *ire, int *icp, int *lenarr, double *wx, double *wy);
void cntc_settemperaturedata(int *ire, int *icp, int *imeth, int *nparam, double *rparam);
/* contact_addon.h:20: the .inp reader of this library is contact_b200/inp.py; returns CNTC_err_other */
void cntc_readinpfile(int *ire, int *inp_type, const char *c_fname, int *len_fname, int *ierror);
/* contact_addon.h:45-73, 86-98, 116: module 1 only */
void cntc_setverticalforce(int *ire, double *fz);
void cntc_setprofileinputfname(int *ire, const char *c_fname, int *len_fname, int *nints, int *iparam, int *nreals,
                               double *rparam);
void cntc_setprofileinputvalues(int *ire, int *npoint, double *values, int *nints, int *iparam, int *nreals,
                                double *rparam);
void cntc_settrackdimensions(int *ire, int *ztrack, int *nparam, double *rparam);
void cntc_setwheelsetdimensions(int *ire, int *ewheel, int *nparam, double *rparam);
void cntc_setwheelsetposition(int *ire, int *ewheel, int *nparam, double *rparam);
void cntc_setwheelsetvelocity(int *ire, int *ewheel, int *nparam, double *rparam);
void cntc_setwheelsetflexibility(int *ire, int *ewheel, int *nparam, double *rparam);
void cntc_getprofilevalues(int *ire, int *itask, int *nints, int *iparam, int *nreals, double *rparam, int *lenarr,
                           double *values);
void cntc_getprofilevalues_new(int *ire, int *itask, int *nints, int *iparam, int *nreals, double *rparam, int *lenarr,
                               double *values);
void cntc_getwheelsetposition(int *ire, int *lenarr, double *values);
void cntc_getwheelsetvelocity(int *ire, int *lenarr, double *values);
void cntc_getnumcontactpatches(int *ire, int *npatch);
void cntc_getcontactlocation(int *ire, int *icp, int *lenarr, double *values);
void cntc_getglobalforces(int *ire, int *icp, int *lenarr, double *values);
/* contact_addon.f90:6243 */
void cntc_finalize(int *ire);
void cntc_finalizelast(void);

/* ------------------------------------------------------------------------------------------------------------
 * Part 2: B200 extensions
 * ---------------------------------------------------------------------------------------------------------- */

/* Batched case scheduler behind cntc_calculate: solve nre result elements (same icp) in ONE launch per grid class.
 * ierror[k] receives what cntc_calculate(ire[k], icp) would have returned. */
void cntc_calculate_batch(int *nre, int *ire, int *icp, int *ierror);

/* iteration counters of the last case of (ire, icp): out[0..7] = itnorm, ItCG of NormCG, ittang, iterations of the
 * tangential solver, ncon, number of tangential solver calls, outer (Panagiotopoulos) iterations, GDsteady: number of
 * line-search trials (negated - 1 when a stagnating GDsteady fell back to SteadyGS, m_solvpt.f90:474-484); nr_itcg =
 * iterations per solver call (Newton-Raphson log) */
int cb200_get_iterations(int ire, int icp, int *out, int lenarr, int *nr_itcg);
/* soutpt results that the reference prints to the .out file only (m_soutpt.f90:424-500): out[0..8] = Fn, Fx, Fy, Mx, My, Mz,
 * elastic energy, frictional power, maximum pressure (CONTACT units) */
int cb200_get_soutpt(int ire, int icp, int lenarr, double *out);
/* deformed distance hs - pen + un of all elements (m_soutpt.f90:459-468) */
int cb200_get_deformed_distance(int ire, int icp, int lenarr, double *out);
/* Devices that cntc_calculate_batch spreads a batch over (the batched case scheduler across the GPUs of one box: contiguous shards
 * of the case list, one host thread per device, no inter-GPU traffic).  n <= 1: the calling thread's current device only (default);
 * n > 1: devices 0..n-1, or devs[0..n-1] when devs is not NULL.  The environment variable CONTACT_B200_DEVICES = all | n | "0,2,3"
 * does the same for callers that cannot be changed (Fortran, python_intfc).  Returns the number of devices in use or a negative code. */
int cb200_set_devices(int n, const int *devs);
/* test hook: replace the stored solution (element division, tractions [3][npot] x, y, n) that the next case of a sequence
 * starts from (I and P digits) */
int cb200_set_state(int ire, int icp, int npot, const int *el, const double *ps);
/* dif, difid of the convergence test of the outer (Panagiotopoulos) loop, m_scontc.f90:510-513, one pair per outer iteration
 * of the last cntc_calculate (at most 16); returns their number */
int cb200_get_outer_history(int ire, int icp, int lenarr, double *dif, double *difid);

/* wall-clock split (s) of the last cntc_calculate / cntc_calculate_batch call: out[0] host set-up of the cases, [1] coefficient
 * transforms (cached per grid and material), [2] device allocation + uploads, [3] solver kernel(s), [4] output products +
 * downloads, [5] total */
int cb200_batch_timing(double *out);
int cb200_batch_timing_output(double *out);   /* split of the output phase: us products, downloads, host post-processing, device frees (s) */

/* last error message of the calling thread (NUL-terminated, owned by the library) */
const char *cb200_last_error(void);
/* number of kernels launched by the library so far (bench.py: "gpu_launches") */
long cb200_num_launches(void);
/* cycle counters of CTA 0 since the last reset: out[0] fused products, [1] cycles inside them, [2] cycles of
 * k_snorm_batch (development aid for the roofline analysis) */
int cb200_conv_prof(unsigned long long *out, int reset);
/* work accounting over all CTAs since the last reset: out[0] products, out[1] nominal flops of those products at the transform
 * size each one used (contact-box level), out[2] algorithmic bytes (17 per box element), out[3] Gauss-Seidel row-sum units */
int cb200_work_counters(unsigned long long *out, int reset);
/* all 32 cycle counters of CTA 0 (slots 4..31: sections of the solver kernels); development aid */
int cb200_solver_prof(unsigned long long *out, int reset);
/* cycle counters of the SteadyGS element step summed over all CTAs since the last reset: out[0] element steps,
 * [1] cycles in the per-element solve (plstrc), [2] re-integration, [3] rank-1 updates + barriers, [4] solver calls */
int cb200_steady_prof(unsigned long long *out, int reset);
/* cycle counters of GDsteady (leader thread of every solver call) since the last reset: out[0] iterations, [1] cycles in
 * the FFT products, [2] search direction, [3] line search: integration along the rows, [4] line search: element pass,
 * reduction, bracketing, [5] step + active set + diagonal scaling + residual, [6] line-search trials, [7] total cycles,
 * [8..10] search direction: copy for E_down, the leader's own rows, wait for the other rows, [11] E_down fall-backs (out: 12 values) */
int cb200_gd_prof(unsigned long long *out, int reset);
/* number of SMs of the device in use, or -99 */
int cb200_num_sms(void);

/* opt_fft_size (m_aijpj.f90:1022-1119) */
int cb200_opt_fft_size(int n);

/* Coefficient set for (grid, material, rolling step): returns handle >= 0 or a negative error.
 * gg, poiss: per body; is_roll/chi/dq as in sgencr (m_visc.f90:127-376). */
int cb200_coefset_create(int mx, int my, double dx, double dy, double gg1, double gg2, double poiss1, double poiss2,
                         int is_roll, double chi, double dq);
/* copy the spatial block cf(-mx:mx-1,-my:my-1,ik,jk) of set (0 cs, 1 cv, 2 csv, 3 ms) to the host buffer (4*mx*my) */
int cb200_coefset_get_block(int handle, int set, int ik, int jk, double *out);
/* plan information: out[0..7] = Fx, Fy, C, nchunk, smem_bytes, fits, nsx, nsy */
int cb200_coefset_plan(int handle, int *out);

/* VecAijPj (m_aijpj.f90:258-451) for ncase independent right-hand sides sharing one coefficient set.
 * p, u: [ncase][3][npot] (direction-major per case, x fastest); el: [ncase][npot] or NULL;
 * iigs -9 (AllElm) / -8 (AllInt); ikarg, jkarg: 1..3, -2 (tangential), -3 (all).  HOST buffers. */
/* gf3_AijPj (m_aijpj.f90:99-254), the direct row sum: out[k] = displacement in direction ik (1..3) of element ii[k] (0-based) due to
 * the tractions p [3][npot] in the directions jkarg (1..3, -2 = tangential, -3 = all) over the column range of the element division
 * el [npot] (contact elements plus one neighbour per row), times 1/G.  Host arrays; npts elements of one case per call. */
int cb200_aijpj(int handle, int set, int ik, int jkarg, int npts, const int *ii, const double *p, const int *el, double *out);
int cb200_vecaijpj(int handle, int set, int ncase, int iigs, int ikarg, int jkarg, const double *p, const int *el,
                   double *u);
/* same with DEVICE buffers, asynchronous on `stream` (a cudaStream_t passed as void*) */
int cb200_vecaijpj_dev(int handle, int set, int ncase, int iigs, int ikarg, int jkarg, const double *d_p,
                       const int *d_el, double *d_u, void *stream);

/* Batched NORM solve (snorm + NormCG, m_snorm.f90:31-378, m_solvpn.f90:24-461) on DEVICE buffers.
 *   d_hs [ncase][npot] undeformed distance; d_el [ncase][npot] in/out element division;
 *   d_pn [ncase][npot] in/out pressures; d_un [ncase][npot] out (may be NULL): u_n = A_zz p_n on the contact area;
 *   d_scal [ncase][8] in/out doubles: pen, fn, (out) itcg, itnorm, ncon, status, err, number of products.
 * ic_norm: 0 approach prescribed (pen in), 1 force prescribed (fn in). */
int cb200_snorm_batch_dev(int handle, int ncase, int ic_norm, int maxgs, int maxin, double eps, const double *d_hs,
                          int *d_el, double *d_pn, double *d_un, double *d_scal, void *stream);
/* initial element division + approach estimate (eldiv0, m_sdis.f90:818-1007) for a gap h[npot]; host buffers */
int cb200_eldiv0(int mx, int my, double dx, double dy, double gg1, double gg2, double poiss1, double poiss2,
                 int ibase, const double *prmudf, int ic_norm, double fn, double pen_in, const double *h, int *el,
                 double *pen_out);
/* same with HOST buffers (copies in, solves, copies out, synchronous) */
int cb200_snorm_batch(int handle, int ncase, int ic_norm, int maxgs, int maxin, double eps, const double *hs,
                      int *el, double *pn, double *un, double *scal);
/* device time [ms] of the most recent solver kernel alone (CUDA events on its stream) */
double cb200_snorm_kernel_ms(void);
/* measured FP64 FMA throughput [TFLOP/s] (roofline denominator for FP64-bound kernels) */
double cb200_fp64_peak_tflops(int reps);
/* Subsurface stresses (sstres_fft, m_subsurf.f90:1097-1257), block type ISUBS 1/5: all npot elements x nz depths z[],
 * for ncase traction fields ps [ncase][3][npot]; table [ncase][nz][npot][18] = columns 4..21 of the reference's
 * table (ux,uy,uz, sighyd,sigvm,sigtr, sigma1..3, sigma(3,3) column-major).  DEVICE / HOST buffer variants. */
int cb200_subsurf_batch_dev(int handle, int ncase, int nz, const double *z, double gg1, double gg2, double poiss1,
                            double poiss2, const double *d_ps, double *d_table, void *stream);
int cb200_subsurf_batch(int handle, int ncase, int nz, const double *z, double gg1, double gg2, double poiss1,
                        double poiss2, const double *ps, double *table);
/* Subsurface stresses in arbitrary points (ISUBS 9, direct sum sstres, m_subsurf.f90:1412-1515); HOST buffers */
int cb200_subsurf_points(int mx, int my, double xc1, double yc1, double dx, double dy, double gg1, double gg2,
                         double poiss1, double poiss2, const double *ps, int npoint, const double *xyz, double *table);
/* workspace the batched solve keeps per case, bytes (for memory planning) */
long cb200_snorm_workspace_bytes(int handle, int ncase);

#ifdef __cplusplus
}
#endif
#endif
