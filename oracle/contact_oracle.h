/*
 * contact_oracle.h -- ORACLE (test infrastructure, not product code).
 *
 * CPU restatement (plain C, double precision) of the hot path of eve70a/CONTACT: influence-coefficient
 * product (AijPj / VecAijPj), NormCG / snorm, tangential solvers and subsurface stresses.  Every function
 * cites the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use anything in oracle/.
 *
 * Parity status: the reference (Intel Fortran + static MKL) cannot be built here (no Fortran compiler), so the
 * oracle is pinned against the reference's golden files (examples .ref_out, testbank .ref_fx, ...) to their
 * printed precision (3-6 digits, exact element pictures and iteration counts) -- see tests/test_oracle_*.py.
 * At 1e-9 the FFT product is pinned only by the restatement's own direct-sum twin (AijPj): "parity unpinned"
 * by the reference's tests at that precision.
 */
#ifndef CONTACT_ORACLE_H
#define CONTACT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* element states and selectors: /root/reference/src/m_gridfunc.f90:30-37 */
enum { CO_EXTER = 0, CO_ADHES = 1, CO_SLIP = 2, CO_PLAST = 3 };
enum { CO_ALLELM = -9, CO_ALLINT = -8, CO_ALLEXT = -7 };
/* coordinate directions: ikXDIR=1, ikYDIR=2, ikZDIR=3; jkALL=-3, jkTANG=-2 */
enum { CO_X = 1, CO_Y = 2, CO_Z = 3, CO_TANG = -2, CO_ALL = -3 };

#define CO_PI   3.14159265358979323846   /* m_globals.f90: pi = 4*atan(1) */
#define CO_TINY 1e-20                    /* m_globals.f90: tiny */

typedef struct { double re, im; } co_cplx;

/* ---- FFT (co_fft.c) ---- */
#define CO_MAXPLAN 24
typedef struct co_fftplan {
    int n, nfac, fac[40];
    co_cplx *tw, *scratch, *work;
} co_fftplan;
typedef struct { int nplan; co_fftplan *plans[CO_MAXPLAN]; } co_plancache;

co_fftplan *co_fft_plan(int n);
void        co_fft_free(co_fftplan *pl);
co_fftplan *co_fft_cached(co_plancache *pc, int n);
void        co_plancache_clear(co_plancache *pc);
void        co_fft_c2c(const co_fftplan *pl, const co_cplx *in, int in_stride, co_cplx *out, int isign);
void        co_fft2_r2c(co_plancache *pc, int n1, int n2, const double *a, co_cplx *A);
void        co_fft2_c2r(co_plancache *pc, int n1, int n2, co_cplx *A, double *a, double scale);

/* ---- element division: t_eldiv, m_gridfunc.f90:116-139 ---- */
typedef struct {
    int  mx, my;
    int *el;              /* (npot) */
    int *row1st, *rowlst; /* (my), 1-based ix, empty row: first=mx, last=0 */
    int  ixmin, ixmax, iymin, iymax;
} co_eldiv;

/* ---- influence coefficients: t_inflcf, m_inflcf.f90:49-106 ---- */
typedef struct {
    int     cf_mx, cf_my;
    double *cf;           /* cf(-mx:mx-1, -my:my-1, 3, 3), x fastest, stored x G */
    double  dx, dy, dq;
    int     nt_cpl;
    double  ga, ga_inv;
    int     use_3bl, use_flxz;
    double  flx_3bl, flx_z;
    int     fft_ok[3][3];
    int     fft_mx, fft_my;
    co_cplx *fft_cf[3][3];
    long    fft_len;
    /* statistics (not in the reference) */
    long    n_cfft;       /* number of coefficient transforms performed */
} co_inflcf;

static inline double *co_cf_ptr(const co_inflcf *c, int ik, int jk)
{   /* block (ik,jk), 1-based; returns pointer to element (ix=-mx, iy=-my) */
    return c->cf + (long) ((jk - 1) * 3 + (ik - 1)) * (4L * c->cf_mx * c->cf_my);
}
#define CO_CF(c, blk, ix, iy) ((blk)[((long)((iy) + (c)->cf_my)) * (2 * (c)->cf_mx) + ((ix) + (c)->cf_mx)])

/* ---- material: t_material subset, m_hierarch_data.f90 ---- */
typedef struct {
    double gg[2], poiss[2];
    double ga, nu, ak;        /* combined */
} co_mater;

/* ---- work counters ---- */
typedef struct {
    long n_prod;          /* single-block FFT products (fft_VecAijPj calls) */
    long n_rowsum;        /* AijPj calls */
    double alg_bytes;     /* algorithmic bytes of products, SURVEY 8(d) */
    double alg_flops;
} co_stats;

/* ---- context: per-thread scratch (replaces the reference's threadprivate descriptors) ---- */
typedef struct {
    co_plancache pc;
    co_stats     st;
    int          fullbox;     /* 1: always use the full-grid box (GPU convention); 0: reference bbox rule */
} co_ctx;

co_ctx *co_ctx_new(void);
void    co_ctx_free(co_ctx *cx);

/* ---- gridfunc (co_gridfunc.c) ---- */
void   co_eldiv_init(co_eldiv *e, int mx, int my);
void   co_eldiv_free(co_eldiv *e);
void   co_areas(co_eldiv *e);

/* ---- coefficients (co_inflcf.c) ---- */
void   co_combin_mater(co_mater *m);
void   co_inflcf_init(co_inflcf *c, int mx, int my, double dx, double dy);
void   co_inflcf_free(co_inflcf *c);
void   co_inflcf_mater(co_inflcf *c, const co_mater *m);
void   co_elascf_pcwcns(double akv, double nuv, int mx, int my, double dx, double dy, double xshft, double yshft,
                        co_inflcf *cs);
/* sgencr for elastic materials (M=0), C=2: cs, cv, csv (+ ms allocated). is_roll: T=2,3 */
void   co_sgencr(const co_mater *m, int mx, int my, double dx, double dy, int is_roll, double chi, double dq,
                 co_inflcf *cs, co_inflcf *cv, co_inflcf *csv, co_inflcf *ms);

/* ---- influence product (co_aijpj.c) ---- */
int    co_opt_fft_size(int n);
double co_aijpj(int ii, int ik, const double *p, const co_eldiv *pel, int jkarg, const co_inflcf *c);
void   co_vecaijpj(co_ctx *cx, const co_eldiv *igs, int iigs, double *u, int ikarg, const double *p,
                   const co_eldiv *pel, int jkarg, co_inflcf *c);
void   co_vecaijpj_direct(const co_eldiv *igs, int iigs, double *u, int ikarg, const double *p,
                   const co_eldiv *pel, int jkarg, const co_inflcf *c);
void   co_fft_makeprec(co_ctx *cx, int ik, co_inflcf *c, int jk, co_inflcf *m);

/* ---- normal problem (co_norm.c) ---- */
typedef struct {
    int    maxgs, maxin, maxnr, maxout;
    double eps;
} co_solv;

typedef struct {
    int    itcg;       /* total CG iterations */
    int    itnorm;     /* NORM iterations (-1: maxin reached) */
    int    diverged;   /* normcg hit its abort_run condition */
} co_norm_info;

int    co_normcg(co_ctx *cx, int ic_norm, int npot, double dxdy, int use_fftprec, int maxcg, double eps,
                 co_inflcf *cs, co_inflcf *ms, const double *hstot, double *pen, double fntrue,
                 co_eldiv *igs, double *ps, int *itcg, double *err);
void   co_snorm(co_ctx *cx, int ic_norm, int mx, int my, double dxdy, const co_solv *solv, const double *hs,
                co_inflcf *cs, co_inflcf *ms, double *pen, double *fntrue, co_eldiv *igs, double *ps,
                co_norm_info *info);

/* ---- solver inputs (co_sdis.c) ---- */
void   co_grid_coords(int mx, int my, double xl, double yl, double dx, double dy, double *x, double *y);
void   co_set_norm_rhs(int ibase, int iplan, int npot, const double *x, const double *y, int nn,
                       const double *prmudf, const double *prmpln, double *hs_n);
void   co_eldiv0(int ic_norm, int mx, int my, double dx, double dy, int ibase, const double *prmudf,
                 const co_mater *m, double fntrue, double *pen, const double *hs_n, co_eldiv *igs);

/* ---- tangential problem and the case driver (co_tang.c) ---- */
#define CO_MAXNR 64
typedef struct {
    /* inputs */
    int    mx, my;
    double xl, yl, dx, dy;
    double gg[2], poiss[2];
    int    ibase, nn;
    const double *prmudf;
    int    tang, norm, force3;          /* T, N, F digits */
    double pen, fn, cksi, ceta, cphi, fxrel, fyrel, fstat, fkin;
    int    maxgs, maxin, maxnr, maxout;
    double eps;
    int    fullbox;
    double chi, dq, facphi;             /* rolling direction, step, spin offset factor (<= 0: default 1/6) */
    int    gausei;                      /* G digit */
    double omegah, omegas;              /* relaxation factors for G = 2, 3 (G = 0, 4: defaults of stang) */
    /* outputs */
    int    *el;                         /* [npot] */
    double *ps;                         /* [3][npot] */
    double *ss;                         /* [3][npot] shift/slip (may be NULL) */
    double pen_out, fn_out, fx_out, fy_out;
    int    itnorm, ittang, itcg_norm, itgs_tang;
    int    nr_n, nr_itcg[CO_MAXNR];     /* one entry per tangential solver call of the Newton-Raphson process */
    double nr_cksi[CO_MAXNR], nr_ceta[CO_MAXNR], nr_fx[CO_MAXNR], nr_fy[CO_MAXNR];
    long   n_prod;
    /* sequences of cases and Hertzian input (appended; zero = not used) */
    int    iestim;                      /* I digit: 0 new initial estimate, 1 regularised previous solution, 2, 3 */
    const int    *el_in;                /* [npot] previous element division (I >= 1) */
    const double *ps_in;                /* [3][npot] previous tractions (I >= 1) */
    const double *pv_in;                /* [3][npot] tractions of the previous time instance (P = 0/1), NULL = zero */
    int    ipotcn;                      /* -1 / -3: 3D Hertzian geometry (curvatures / semi-axes given); else grid as given */
    double hz_a1, hz_b1, hz_aa, hz_bb, hz_scale;
    int    itout;                       /* out: outer (Panagiotopoulos) iterations */
    double gd[8];                       /* G = 5: fdecay, betath, kdowfb, d_ifc, d_lin, d_cns, d_slp, pow_s (as in the .inp record) */
    int    gd_fallback;                 /* out: GDsteady stagnated, SteadyGS was used (m_solvpt.f90:474-484) */
    double pan_dif[16], pan_difid[16];  /* out: dif / difid of the convergence test of panprc (m_scontc.f90:510-513), per outer iteration */
} co_case;
/* parameters of GDsteady after solv_input (m_sinput.f90:649-680) */
typedef struct { int gd_meth, kdown, kdowfb; double fdecay, betath, d_ifc, d_lin, d_cns, d_slp, pow_s; } co_gdparams;
void   co_gdparams_set(const double gd[8], co_gdparams *sp);
int    co_gdsteady(co_ctx *cx, int mx, int my, int maxgd, double eps, const double *ws, co_inflcf *cs, const double *mus,
                   co_eldiv *igs, double *ps, double *ss, const co_gdparams *sp, double *err, int *lstagn);
void   co_ellip_kebd(double mc, double *K, double *E, double *B, double *D);
void   co_hertz3d(double e_star, int ipotcn, double *a1, double *b1, double *aa, double *bb, int ic_norm, double *pen,
                  double *fn, double *cp, double *rho);

void   co_tangcg(co_ctx *cx, int npot, int maxcg, double eps, const double *ws, co_inflcf *cs, co_inflcf *ms,
                 const double *mu, co_eldiv *igs, double *ps, double *ss, int *itcg, double *err);
int    co_contac(co_case *c);

/* ---- steady rolling (co_steady.c) ---- */
void   co_plstrc(int *el, const double coef[2][2], double eps, double omegah, double omegas, double pr[3], double mus,
                 double s[2]);
void   co_stdygs(co_ctx *cx, int mx, int my, const double *ws, co_inflcf *cs, const double *mus, co_eldiv *igs, double *ps,
                 double *ss, int k, double eps, int maxgs, double omegah, double omegas, int *info, int *itgs_out, double *err);
void   co_sxbnd_facdt(int mx, int my, const co_eldiv *igs, const double *x, double dx, double dq, double *facdt);
/* leading-edge administration: t_leadedge, m_leadedge.f90:30-70 (1-based position index j) */
typedef struct { int npos, *jbnd, *ixbnd, *ii2j; double *xbnd, *facdx, *ubnd, *facdt; } co_leadedge;
void   co_sxbnd(int mx, int my, int is_roll, int use_ledg, const co_eldiv *igs, const double *x, double dx, double dq,
                co_leadedge *lg);
void   co_leadedge_free(co_leadedge *lg);
void   co_subnd(co_ctx *cx, int mx, int my, const double *p, const co_eldiv *pel, const co_inflcf *c, co_leadedge *lg);
void   co_cnvxgs(co_ctx *cx, int mx, int my, int is_ssrol, const double *ws, co_inflcf *cs, co_inflcf *csv, co_leadedge *lg,
                 const double *mus, co_eldiv *igs, double *ps, double *ss, int k, const int *iel, double eps, int maxgs,
                 double omegah, double omegas, int *info, int *itgs_out, double *err);

/* ---- subsurface stresses (co_subsurf.c) ---- */
void   co_stres1_pcwcns(double dx, double dy, double gg, double v[3][3][4], double vnu[3][3][4], const double xw[3],
                        const double xp[2]);
void   co_sym3_eigenvalues(const double s[3][3], double ev[3]);
void   co_sstres_derived(double gg, double poiss, int neg, double vr[3][4], double out[18]);
void   co_sstres_point(int mx, int my, double dx, double dy, const double *x, const double *y, const double gg[2],
                       const double poiss[2], const double *ps, const double xw[3], double out[18]);
void   co_sstres_inflcf(int mx, int my, double dx, double dy, const double gg[2], const double poiss[2], double zw,
                        co_inflcf ck[4]);
void   co_subsurf_block_fft(co_ctx *cx, int mx, int my, double dx, double dy, const double gg[2], const double poiss[2],
                            const int *el, const double *ps, int nz, const double *z, int use_fft, double *table);

#ifdef __cplusplus
}
#endif
#endif
