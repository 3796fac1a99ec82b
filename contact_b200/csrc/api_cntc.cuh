// api_cntc.cuh -- the reference's cntc_* / subs_* C-ABI for module-3 contact problems, backed by the B200 engine.
//
// State model as in the reference (/root/reference/src/m_caddon_data.f90:89-225): result elements ire in [1,999]
// are created lazily on first touch, each holds contact problems icp in [1,9]; a case is set-calls ->
// cntc_calculate -> get-calls; state persists between cases.  Units handling follows cntc_select_units
// (m_global_data.f90:463-526).  Control digits outside the B200 hot-path scope are rejected with an error code
// (never abort): see DESIGN.md "scope".
#pragma once
#include <thread>
#include <atomic>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <memory>
#include "api_lowlevel.cuh"

namespace cb200 {

// magic numbers of /root/reference/src/caddon_flags.inc:13-178 (interface constants)
enum {
    CNTC_if_units = 1933, CNTC_un_cntc = 1934, CNTC_un_spck = 1935, CNTC_un_si = 1936, CNTC_un_imper = 1937,
    CNTC_ic_config = 1967, CNTC_ic_pvtime = 1970, CNTC_ic_bound = 1971, CNTC_ic_tang = 1972, CNTC_ic_norm = 1973,
    CNTC_ic_force = 1974, CNTC_ic_frclaw = 1976, CNTC_ic_discns = 1977, CNTC_ic_inflcf = 1978, CNTC_ic_mater = 1979,
    CNTC_ic_exrhs = 1980, CNTC_ic_xflow = 1981, CNTC_ic_heat = 1982, CNTC_ic_iestim = 1983, CNTC_ic_output = 1984,
    CNTC_ic_flow = 1985, CNTC_ic_return = 1986, CNTC_ic_matfil = 1987, CNTC_ic_sens = 1988, CNTC_ic_ifmeth = 1989,
    CNTC_ic_ifvari = 1990, CNTC_ic_sbsout = 1991, CNTC_ic_sbsfil = 1992, CNTC_ic_npomax = 1993,
    CNTC_if_idebug = 2000, CNTC_if_licdbg = 2001, CNTC_if_wrtinp = 2002, CNTC_if_openmp = 2003, CNTC_if_timers = 2004,
    CNTC_if_ncase = 2005,
    CNTC_fld_h = 1, CNTC_fld_mu = 2, CNTC_fld_px = 3, CNTC_fld_py = 4, CNTC_fld_pn = 5, CNTC_fld_ux = 7, CNTC_fld_uy = 8,
    CNTC_fld_un = 9, CNTC_fld_taucrt = 11, CNTC_fld_uplsx = 12, CNTC_fld_uplsy = 13, CNTC_fld_sx = 15, CNTC_fld_sy = 16,
    CNTC_fld_temp1 = 20, CNTC_fld_temp2 = 21, CNTC_fld_wx = 22, CNTC_fld_wy = 23,
    CNTC_err_allow = -12, CNTC_err_norm = -27, CNTC_err_tang = -28, CNTC_err_icp = -31, CNTC_err_discr = -34,
    CNTC_err_input = -39, CNTC_err_other = -99
};

struct Scaling { int units = CNTC_un_cntc; double len = 1, area = 1, forc = 1, veloc = 1, angle = 1, body = 1; };

struct Problem {
    // control digits (t_ic, m_hierarch_data.f90:134-409), defaults of ic_init for the library
    int pvtime = 2, bound = 0, tang = 0, norm = 0, force3 = 0, stress = 0;
    int frclaw = 0, discns = 2, gencr = 2, mater = 0, rznorm = 2, rztang = 0;
    int gausei = 0, iestim = 0, matfil = 0, output = 0, flow = 0, ret = 1, sens = 0, wrtinp = 0;
    Scaling scl;
    Material mat = { { 82000.0, 82000.0 }, { 0.28, 0.28 }, 0, 0, 0 };
    // potential contact (t_potcon)
    int ipotcn = 1, mx = 1, my = 1;
    double xl = 0, yl = 0, dx = 1, dy = 1, xh = 0, yh = 0, xc1 = 0, yc1 = 0, xcm = 0, ycm = 0;
    // hertz
    double hz_aa = 0, hz_bb = 0, hz_a1 = 0, hz_b1 = 0, hz_aob = 0, hz_scale = 1;
    // geometry
    int ibase = 1, iplan = 1, nn = 0;
    std::vector<double> prmudf = std::vector<double>(10, 0.0);
    // kinematics
    double pen = 0, fntrue = 0, cksi = 0, ceta = 0, cphi = 0, fxrel = 0, fyrel = 0, chi = 0, dq = 1, veloc = 1, dt = 1;
    // friction (L = 0)
    double fstat = 0.3, fkin = 0.3;
    // solver settings (solv_init, m_hierarch_data.f90:1722-1751)
    int maxgs = 999, maxin = 20, maxnr = 25, maxout = 1;
    double eps = 1e-5, omegah = 0.9, omegas = 1.0, dq_eff = 1.0;
    // GDsteady (G=5), defaults of solv_init (m_hierarch_data.f90:1722-1751): fdecay, betath, d_ifc, d_lin, d_cns, d_slp, pow_s
    double gd_fdecay = (double) 0.90f, gd_betath = (double) 0.10f, gd_difc = 1.0, gd_dlin = 1.0, gd_dcns = 1.0, gd_dslp = 1.0, gd_pows = (double) 0.90f;
    int gd_meth = 1, gd_kdown = 3, gd_kdowfb = 1, gd_fallback = 0, gd_ntrial = 0;
    // results
    int ncase = 0, itnorm = 0, ittang = 0, itcg = 0, ncon = 0, nadh = 0, nslip = 0, status = 0, itout = 0;
    double pan_dif[16] = { 0 }, pan_difid[16] = { 0 };
    double mxtrue = 0, mytrue = 0, elen = 0, frpow = 0;      // soutpt: moments about x and y, elastic energy, frictional power
    std::vector<int> el;
    std::vector<double> ps, us, hs, ss;  // [3][npot]
    std::vector<double> pv;              // [3][npot] tractions of the previous time instance (set_prev_data)
    double hz_cp = 0, hz_rho = 0;
    int itgs = 0;
    std::vector<int> nr_itcg;            // TangCG iterations per solver call of the Newton-Raphson process
    double fcntc[3] = { 0, 0, 0 }, mztrue = 0;
    double sens_nr[2][2] = { { 0, 0 }, { 0, 0 } };   // d(fx, fy)/d(cksi, ceta) of the Newton-Raphson process (m_solvpt.f90:51-378)
    int exrhs_len = 0, heat_meth = 0;                 // cntc_setextrarigidslip / cntc_settemperaturedata were called (stored, not served)
    double t_wall = 0, t_cpu = 0;
    bool solved = false;
    // subsurface blocks
    struct SubsBlock { int isubs = 0; std::vector<double> x, y, z; std::vector<double> table; int nx = 0, ny = 0, nz = 0; };
    std::map<int, SubsBlock> subs;
};

struct ResultElement { int imodul = 3; std::map<int, std::unique_ptr<Problem>> cps; };

struct Registry {
    std::mutex mu;
    std::map<int, std::unique_ptr<ResultElement>> res;
    int idebug = 1;
};

inline Registry &registry() { static Registry r; return r; }

// cntc_activate (m_caddon_data.f90:89-225): lazily create; error codes -101.. for invalid ids
inline Problem *activate(int ire, int icp, int *ierror)
{
    Registry &R = registry();
    if (ierror) *ierror = 0;
    if (ire < 1 || ire > 999) { if (ierror) *ierror = -101; last_error() = "invalid result element id"; return nullptr; }
    if (icp < 1 || icp > 9) { if (ierror) *ierror = CNTC_err_icp; last_error() = "invalid contact problem id (module 1, icp=-1, is outside the hot-path scope)"; return nullptr; }
    std::lock_guard<std::mutex> lk(R.mu);
    auto &re = R.res[ire];
    if (!re) re.reset(new ResultElement());
    auto &cp = re->cps[icp];
    if (!cp) cp.reset(new Problem());
    return cp.get();
}

inline void select_units(Scaling &s, int iunits)
{   // m_global_data.f90:463-526
    if (iunits == CNTC_un_cntc || iunits == CNTC_un_imper) { s = Scaling(); s.units = iunits; }
    else if (iunits == CNTC_un_spck) { s.units = iunits; s.len = 1e3; s.area = 1e6; s.forc = 1; s.veloc = 1e3; s.angle = 1; s.body = -1; }
    else if (iunits == CNTC_un_si) { s.units = iunits; s.len = 1e3; s.area = 1e6; s.forc = 1; s.veloc = 1e3; s.angle = 1; s.body = 1; }
}

inline void potcon_fill(Problem &p)
{   // m_hierarch_data.f90:1801-1859
    const int t = p.ipotcn;
    if (t == 2) { p.dx = (p.xh - p.xl) / p.mx; p.dy = (p.yh - p.yl) / p.my; }
    else if (t == 4) { p.dx = (p.xcm - p.xc1) / std::max(1, p.mx - 1); p.dy = (p.ycm - p.yc1) / std::max(1, p.my - 1); }
    if (t == 3 || t == 4) { p.xl = p.xc1 - 0.5 * p.dx; p.yl = p.yc1 - 0.5 * p.dy; }
    else { p.xc1 = p.xl + 0.5 * p.dx; p.yc1 = p.yl + 0.5 * p.dy; }
    if (t == 1 || t == 3 || t == 4) { p.xh = p.xl + p.mx * p.dx; p.yh = p.yl + p.my * p.dy; }
    if (t >= 1 && t <= 3) { p.xcm = p.xc1 + (p.mx - 1) * p.dx; p.ycm = p.yc1 + (p.my - 1) * p.dy; }
}

// ---- solver inputs on the host: set_norm_rhs (m_sdis.f90:321-494), eldiv0 (m_sdis.f90:818-1007) ----
inline void undeformed_distance(const Problem &p, std::vector<double> &h)
{
    const int npot = p.mx * p.my;
    h.assign(npot, 0.0);
    const double *b = p.prmudf.data();
    for (int iy = 0; iy < p.my; iy++)
        for (int ix = 0; ix < p.mx; ix++) {
            const double x = p.xc1 + ix * p.dx, y = p.yc1 + iy * p.dy;
            const int ii = iy * p.mx + ix;
            double v = 0.0;
            if (p.ibase == 1) {
                v = b[0] * x * x + b[1] * x * y + b[2] * y * y + b[3] * x + b[4] * y + b[5];
            } else if (p.ibase == 2) {
                const double xm = b[1], rm = b[2], y1 = b[3], dy1 = b[4], yn = y1 + (p.nn - 1) * dy1;
                int mleft; double yleft;
                if (y < y1) { yleft = y1; mleft = 1; }
                else if (y >= yn) { yleft = yn - dy1; mleft = p.nn - 1; }
                else { mleft = (int) ((y - y1) / dy1) + 1; yleft = y1 + (mleft - 1) * dy1; }
                mleft += 5;
                const double rc = (b[mleft] - b[mleft - 1]) / dy1;
                v = b[mleft - 1] + rc * (y - yleft);
                v = v + (x - xm) * (x - xm) / (2.0 * rm);
            } else if (p.ibase == 3) {
                v = b[0] * sin(b[1] * (x - b[2])) - b[3] * sin(b[4] * (x - b[5])) + x * x / b[6] + y * y / b[7];
            } else if (p.ibase == 9) {
                v = b[ii];
            }
            h[ii] = v;
        }
}

inline void initial_eldiv(Problem &p, const std::vector<double> &h, std::vector<int> &el, double &pen)
{
    const int npot = p.mx * p.my;
    const double facpen = (double) 0.60f, reltol = (double) 0.01f;      // REAL(4) literals, m_sdis.f90:832
    const double dxdy = p.dx * p.dy, fnscal = p.fntrue / p.mat.ga;
    const double pi = 3.14159265358979323846;
    double hsmin = h[0];
    for (int i = 1; i < npot; i++) if (h[i] < hsmin) hsmin = h[i];
    double pentru = pen - hsmin;
    auto count_below = [&](double thr) { int c = 0; for (int i = 0; i < npot; i++) if (h[i] - hsmin < thr) c++; return c; };
    if (p.norm == 1 && p.my == 1) {
        double rm = 1.0;
        if (p.ibase == 1) rm = 0.5 / std::max(1e-6, p.prmudf[0]);
        else if (p.ibase == 2) rm = p.prmudf[2];
        else if (p.ibase == 3) rm = 0.5 * p.prmudf[6];
        const double cdy = 0.35 - 0.05 * (log(p.dy) - log(200.0));
        pentru = pow(fnscal / p.dy * (1.0 - p.mat.nu) / 0.6 / cdy / pow(rm, (double) 0.1f), (double) 0.926f);
        pen = pentru + hsmin;
    } else if (p.norm == 1) {
        const double fac = 0.75 * sqrt(pi) * (1.0 - p.mat.nu);
        double penmin = 0.0, fnmin = 0.0, penmax = fac * fnscal / sqrt(dxdy);
        double fnmax = penmax * sqrt(dxdy * count_below(facpen * penmax)) / fac;
        for (int iter = 0; iter < 50 && fabs(fnmax - fnmin) > reltol * fnscal; iter++) {
            const double penmid = 0.5 * (penmax + penmin);
            const double fnmid = penmid * sqrt(dxdy * count_below(facpen * penmid)) / fac;
            if (fnmid < fnscal) { penmin = penmid; fnmin = fnmid; } else { penmax = penmid; fnmax = fnmid; }
        }
        pentru = penmax;
        pen = pentru + hsmin;
    }
    el.resize(npot);
    for (int i = 0; i < npot; i++) el[i] = (h[i] - hsmin < facpen * pentru) ? 1 : 0;
    for (int i = 0; i < npot; i++) if (h[i] > (double) 1e29f) el[i] = 0;          // m_sdis.f90:748-750
}

inline int count_at_boundary(const Problem &p)
{   // eldiv_count_atbnd(igs, 0), m_gridfunc.f90:437-490
    const int nx = p.mx, ny = p.my;
    int cnt = 0;
    const int iystep = (nx >= 3) ? 1 : std::max(1, ny - 1);
    for (int iy = 1; iy <= ny; iy += iystep) {
        const int ixstep = (ny >= 3 && (iy == 1 || iy == ny)) ? 1 : std::max(1, nx - 1);
        for (int ix = 1; ix <= nx; ix += ixstep) if (p.el[(iy - 1) * nx + ix - 1] >= 1) cnt++;
    }
    return cnt;
}

// check_case / scope check: returns 0 or an error code
inline int check_scope(const Problem &p)
{
    if (p.tang == 2 && fabs(p.chi) > 0.01) { last_error() = "transient rolling (T=2): only CHI = 0 is served by the B200 path"; return CNTC_err_other; }
    if (p.tang != 0 && p.frclaw != 0) { last_error() = "L-digit: only Coulomb friction (L=0)"; return CNTC_err_other; }
    if (p.tang == 3 && p.gausei == 2 && fabs(p.chi) > 0.01) { last_error() = "ConvexGS in steady rolling: only CHI = 0 is served by the B200 path"; return CNTC_err_other; }
    if (p.tang == 3 && p.gausei == 5 && fabs(p.chi) > 0.01) { last_error() = "GDsteady: only CHI = 0 is served by the B200 path"; return CNTC_err_other; }
    if (p.tang == 3 && fabs(p.chi) > 0.01 && fabs(p.chi - 3.14159265358979323846) <= 0.01) { last_error() = "CHI = pi (rolling in -x) is not served by the B200 path yet"; return CNTC_err_other; }
    if (p.mater != 0) { last_error() = "M-digit: only the elastic half-space (M=0) is in the hot-path scope"; return CNTC_err_other; }
    if (p.gencr != 2 && p.gencr != 1) { last_error() = "C-digit: only piecewise-constant analytical coefficients (C=2)"; return CNTC_err_other; }
    if (p.bound != 0) { last_error() = "B-digit: only the full normal problem (B=0)"; return CNTC_err_other; }
    if (!((p.ipotcn >= 1 && p.ipotcn <= 4) || p.ipotcn == -1 || p.ipotcn == -3)) { last_error() = "IPOTCN: only 1..4 and the 3D Hertzian options -1, -3 are served by the B200 path"; return CNTC_err_other; }
    if (p.iplan != 1) { last_error() = "IPLAN: only the unrestricted planform"; return CNTC_err_other; }
    if (p.ibase != 1 && p.ibase != 2 && p.ibase != 3 && p.ibase != 9) { last_error() = "invalid IBASE"; return CNTC_err_input; }
    if (p.ibase == 9 && (int) p.prmudf.size() < p.mx * p.my) { last_error() = "IBASE=9 needs npot values"; return CNTC_err_input; }
    if (p.tang != 0 && (p.rztang == 9 || p.exrhs_len > 0)) { last_error() = "E-digit: extra rigid slip (cntc_setextrarigidslip) is not served by the B200 path"; return CNTC_err_other; }
    if (p.heat_meth >= 1) { last_error() = "H-digit: temperature calculation is outside the hot-path scope"; return CNTC_err_other; }
    return 0;
}

// complete elliptic integrals K, E and the associate integrals B = (E - mc K)/m, D = (K - E)/m by the arithmetic-
// geometric mean (the reference uses Fukushima's series, m_hertz.f90:509-...; same functions, own evaluation).
// D stays finite for m -> 0: (K - E)/K = (1/2) sum 2^n c_n^2 with c_0^2 = m, evaluated divided by m term by term.
inline void ellip_kebd(double mc, double &K, double &E, double &B, double &D)
{
    const double pi = 3.14159265358979323846, m = 1.0 - mc;
    double a = 1.0, b = sqrt(mc);
    double t = (1.0 - b) / (4.0 * (1.0 + b));        // c_1^2 / m
    double sum = 1.0 + 2.0 * t, pw = 2.0;
    double an = 0.5 * (a + b), bn = sqrt(a * b);
    double c2 = t * m;                                  // c_1^2
    a = an; b = bn;
    for (int it = 0; it < 40 && fabs(a - b) > 1e-17 * a; it++) {
        an = 0.5 * (a + b); bn = sqrt(a * b);
        const double cn = 0.5 * (a - b);               // c_{n+1} = (a_n - b_n)/2
        pw *= 2.0;
        t = (m > 0.0) ? cn * cn / m : 0.0;
        if (m <= 1e-300) t = 0.0;
        sum += pw * t;
        c2 = cn * cn;
        a = an; b = bn;
    }
    (void) c2;
    K = pi / (2.0 * a);
    D = 0.5 * K * sum;
    B = K - D;
    E = B + mc * D;
}

// hzcalc3d (m_hertz.f90:267-385): 3D Hertzian point contact.  ipotcn -1: curvatures given, -3: semi-axes given;
// ic_norm 0: approach given, 1: normal force given
inline void hertz3d(double e_star, int ipotcn, double &a1, double &b1, double &aa, double &bb, int ic_norm, double &pen,
                    double &fn, double &cp, double &rho)
{
    const double pi = 3.14159265358979323846;
    double K, E, B, D;
    auto eli = [&](double k) { ellip_kebd(1.0 - k * k, K, E, B, D); return (E - B) / B; };
    if (ipotcn == -1) {
        const bool zbla = b1 <= a1;
        const double y = zbla ? b1 / a1 : a1 / b1;
        double xl = 0.0, xr = 1.0, elr = eli(xr), ell = eli(xl), x = 0.5;      // bisection on the modulus, :386-440
        while (fabs(xr - xl) > 1e-9) {
            x = 0.5 * (xl + xr);
            const double elx = eli(x);
            if ((y - elr) * (y - elx) <= 0.0) { xl = x; ell = elx; } else { xr = x; elr = elx; }
        }
        (void) ell;
        eli(x);
        const double g = sqrt(std::max(1e-40, 1.0 - x * x)), sg = sqrt(g);
        rho = 2.0 / (a1 + b1);
        if (ic_norm == 1) { cp = pow(3.0 * std::max(0.0, fn) * rho * E / (4.0 * pi * e_star * sg), 1.0 / 3.0); pen = 2.0 * (cp * sg) * (cp * sg) * K / (rho * E); }
        else { cp = sqrt(std::max(0.0, pen) * rho * E / (2.0 * K * sg * sg)); fn = 4.0 * pi * cp * cp * cp * e_star * sg / (3.0 * rho * E); }
        if (zbla) { aa = cp * sg; bb = cp / sg; } else { aa = cp / sg; bb = cp * sg; }
    } else {
        const bool zbla = bb <= aa;
        const double g = zbla ? bb / aa : aa / bb, sg = sqrt(g), k = sqrt(1.0 - g * g);
        const double y = eli(k);
        cp = sqrt(aa * bb);
        if (ic_norm == 1) { rho = 4.0 * pi * cp * cp * cp * e_star * sg / (3.0 * fn * E); pen = 2.0 * (cp * sg) * (cp * sg) * K / (rho * E); }
        else {
            pen = std::max(pen, 1e-9);
            rho = 2.0 * (cp * sg) * (cp * sg) * K / (pen * E);
            fn = 4.0 * pi * cp * cp * cp * e_star * sg / (3.0 * rho * E);
        }
        const double apb = 2.0 / rho, ama = apb / (y + 1.0), ami = y * ama;
        if (zbla) { a1 = ami; b1 = ama; } else { a1 = ama; b1 = ami; }
    }
}

struct Problem;
// hzsol + potcon_hertz (m_hertz.f90:29-91, m_hierarch_data.f90:1889-1959) for IPOTCN = -1, -3: quadratic geometry from
// the Hertz solution, potential contact = scale * contact ellipse
inline int hertz_setup(Problem &p);

// check_roll_stepsize (m_sdis.f90:125-204): SteadyGS forces chi = 0 and dq = dx; shifts use chi = 0, dq = 1
inline void roll_stepsize(const Problem &p, double &chi, double &dq)
{
    if (p.tang == 2 || p.tang == 3) {
        chi = p.chi; dq = p.dq;
        if (p.tang == 3 && p.gausei != 2) { if (fabs(chi) > 0.01 && fabs(chi - 3.14159265358979323846) > 0.01) chi = 0.0; dq = p.dx; }
    } else { chi = 0.0; dq = 1.0; }
}


inline int hertz_setup(Problem &p)
{
    if (p.ipotcn != -1 && p.ipotcn != -3) { last_error() = "IPOTCN: only the 3D Hertzian options -1 (curvatures) and -3 (semi-axes) are served by the B200 path"; return CNTC_err_other; }
    const double e_star = p.mat.ga / (1.0 - p.mat.nu);
    hertz3d(e_star, p.ipotcn, p.hz_a1, p.hz_b1, p.hz_aa, p.hz_bb, p.norm, p.pen, p.fntrue, p.hz_cp, p.hz_rho);
    p.ibase = 1; p.iplan = 1;
    p.prmudf.assign(10, 0.0);
    p.prmudf[0] = p.hz_a1; p.prmudf[2] = p.hz_b1;
    p.xl = -p.hz_scale * std::max(1e-9, p.hz_aa); p.yl = -p.hz_scale * std::max(1e-9, p.hz_bb);
    p.xh = -p.xl; p.yh = -p.yl;
    p.dx = (p.xh - p.xl) / p.mx; p.dy = (p.yh - p.yl) / p.my;           // potcon_fill with ipotcn = 2
    p.xc1 = p.xl + 0.5 * p.dx; p.yc1 = p.yl + 0.5 * p.dy;
    p.xcm = p.xc1 + (p.mx - 1) * p.dx; p.ycm = p.yc1 + (p.my - 1) * p.dy;
    return 0;
}

// contac (m_scontc.f90:37-216) for a batch of problems: host set-up, ONE device launch per coefficient class, gather
// wall-clock split of the last calculate_batch call (s): [0] host set-up of the cases, [1] coefficient transforms (cached),
// [2] device allocation + uploads, [3] solver kernel(s), [4] output products + downloads, [5] total
// split of [4]: [6] us products incl. their buffers, [7] downloads, [8] host post-processing, [9] device frees
inline double *batch_timing() { static double t[CB_MAX_DEVICES][10] = { { 0 } }; return t[current_device()]; }

// independent per-case host work (copies into the problems' own vectors, force sums) over a few host threads
template <class F> inline void host_parallel_for(int n, F fn)
{
    const int nt = std::max(1, std::min({ n / 16, (int) std::thread::hardware_concurrency(), 16 }));
    if (nt <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
    std::vector<std::thread> th;
    int dev = -1;
    cudaGetDevice(&dev);                             // new threads start on device 0: keep the workers on the caller's device
    for (int t = 0; t < nt; t++)
        th.emplace_back([=, &fn] {
            if (dev >= 0) cudaSetDevice(dev);
            for (int i = (int) ((long) n * t / nt); i < (int) ((long) n * (t + 1) / nt); i++) fn(i);
        });
    for (auto &x : th) x.join();
}

// Device and pinned-host work space of calculate_batch, kept between calls and grown on demand: a sweep that arrives in
// chunks (result elements are limited to 1..999) pays for cudaMalloc / cudaFree / cudaHostAlloc once.  The device part is
// cleared at every call, so a case never sees data of an earlier one.  Guarded by `mu`: one batch at a time per process.
struct BatchPool {
    std::mutex mu;
    double *d_buf = nullptr, *d_us = nullptr, *d_pb = nullptr, *h_fld = nullptr, *h_us = nullptr;
    int *d_el = nullptr, *d_next = nullptr, *h_el = nullptr;
    ContactCase *d_cases = nullptr;
    size_t c_buf = 0, c_us = 0, c_pb = 0, c_el = 0, c_cases = 0, c_hfld = 0, c_hus = 0, c_hel = 0;
};
inline BatchPool &batch_pool() { static BatchPool p[CB_MAX_DEVICES]; return p[current_device()]; }
template <class T> inline bool pool_dev(T *&ptr, size_t &cap, size_t need)
{
    if (need <= cap) return true;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    if (cudaMalloc(&ptr, need * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return false; }
    cap = need;
    return true;
}
template <class T> inline bool pool_pinned(T *&ptr, size_t &cap, size_t need)
{
    if (need <= cap) return true;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr; cap = 0;
    if (cudaMallocHost(&ptr, need * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return false; }
    cap = need;
    return true;
}

inline void calculate_batch(const std::vector<Problem *> &probs, std::vector<int> &ierr)
{
    BatchPool &BP = batch_pool();
    std::lock_guard<std::mutex> pool_lock(BP.mu);
    double *bt = batch_timing();
    for (int k = 0; k < 10; k++) bt[k] = 0.0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const size_t nb = probs.size();
    ierr.assign(nb, 0);
    std::map<CoefSet *, std::vector<size_t>> groups;
    std::vector<std::vector<double>> hs(nb);
    std::vector<std::vector<int>> el0(nb);
    std::vector<double> pen0(nb);
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t k = 0; k < nb; k++) {
        Problem &p = *probs[k];
        p.ncase++;
        if ((ierr[k] = check_scope(p))) continue;
        combine_material(p.mat);
        if (p.ncase <= 1) p.pvtime = 2;                                   // check_case, m_scontc.f90:268-274
        if (p.ipotcn < 0 && (ierr[k] = hertz_setup(p))) continue;         // contac, m_scontc.f90:81-91
    }
    // the O(npot) part of the set-up is independent per case: over host threads
    host_parallel_for((int) nb, [&](int kk) {
        const size_t k = (size_t) kk;
        if (ierr[k]) return;
        Problem &p = *probs[k];
        undeformed_distance(p, hs[k]);
        const int npot = p.mx * p.my;
        const bool have_prev = p.solved && (int) p.el.size() == npot && (int) p.ps.size() == 3 * npot;
        // set_prev_data (m_sdis.f90:208-317): P = 0 full sequence, 1 sequence for the normal part, 2 initiation of contact
        if (p.pvtime == 0 && have_prev) p.pv = p.ps;
        else if (p.pvtime == 1 && have_prev) { p.pv.assign(3 * (size_t) npot, 0.0); std::copy(p.ps.begin() + 2 * (size_t) npot, p.ps.end(), p.pv.begin() + 2 * (size_t) npot); }
        else if (p.pvtime != 3 || (int) p.pv.size() != 3 * npot) p.pv.assign(3 * (size_t) npot, 0.0);
        // init_curr_data (m_sdis.f90:587-768)
        double pen = p.pen;
        const int iestim = have_prev ? p.iestim : 0;
        if (iestim == 0) {
            p.ps.assign(3 * (size_t) npot, 0.0);
            if (p.ipotcn < 0 && p.bound == 0) {                           // Hertzian solution as initial estimate, :640-700
                const double pi = 3.14159265358979323846, pnmax = 3.0 * p.fntrue / (2.0 * pi * p.hz_aa * p.hz_bb);
                el0[k].assign(npot, 0);
                for (int iy = 0; iy < p.my; iy++) for (int ix = 0; ix < p.mx; ix++) {
                    const double x = p.xc1 + ix * p.dx, y = p.yc1 + iy * p.dy;
                    const double f = 1.0 - (x / p.hz_aa) * (x / p.hz_aa) - (y / p.hz_bb) * (y / p.hz_bb);
                    const double v = pnmax * sqrt(std::max(0.0, f));
                    p.ps[2 * (size_t) npot + iy * p.mx + ix] = v;
                    el0[k][iy * p.mx + ix] = v > 1e-20 ? 1 : 0;
                }
            } else
                initial_eldiv(p, hs[k], el0[k], pen);
        } else {
            el0[k] = p.el;                                                // I >= 1: keep the element division
            if (iestim == 2 || p.tang == 0) std::fill(p.ps.begin(), p.ps.begin() + 2 * (size_t) npot, 0.0);
            if (iestim == 1) {                                            // regularise the tractions, :722-735
                for (int i = 0; i < npot; i++) {
                    double &px = p.ps[i], &py = p.ps[npot + i], &pn = p.ps[2 * (size_t) npot + i];
                    if (el0[k][i] <= 0) { px = py = pn = 0.0; }
                    else if (el0[k][i] == 2) { const double pa = std::max(1e-10, sqrt(px * px + py * py)); px = p.fstat * pn * px / pa; py = p.fstat * pn * py / pa; }
                }
            }
            if (iestim == 2) for (int i = 0; i < npot; i++) if (el0[k][i] >= 1) el0[k][i] = 1;
            for (int i = 0; i < npot; i++) if (hs[k][i] > (double) 1e29f) el0[k][i] = 0;
        }
        pen0[k] = pen;
    });
    for (size_t k = 0; k < nb; k++) {
        if (ierr[k]) continue;
        Problem &p = *probs[k];
        if (p.ret > 1) { ierr[k] = 0; continue; }                        // R=2,3: checks only
        CoefSet *cs = nullptr;
        double chi_e, dq_e;
        roll_stepsize(p, chi_e, dq_e);
        // latency mode: a single case (the usual call pattern of a multibody code, and of case sequences) would occupy one
        // of 148 SMs on the batched path; with a grid of some size it runs faster spread over the whole GPU
        const int whole = (nb == 1 && p.tang != 3 && !(p.tang != 0 && p.gausei == 2) && p.mx * p.my >= 1024 && p.mx >= 16 && p.my >= 16) ? 1 : 0;
        int rc = get_coefset(p.mx, p.my, p.dx, p.dy, p.mat, p.tang >= 2 ? 1 : 0, chi_e, dq_e, 0, &cs, whole);
        if (rc) { ierr[k] = rc; continue; }
        // grids beyond one CTA's shared memory go to the whole-GPU path, which serves T = 0 and T = 1 (TangCG)
        groups[cs].push_back(k);
    }
    bt[0] = secs(t0, now());
    for (auto &g : groups) {
        auto tg0 = now();
        CoefSet &cs = *g.first;
        // Longest case first: the kernel hands the cases of a group to the CTAs through one queue in the order of this list, and a
        // contact case costs ~ (contact elements) x (Gauss-Seidel sweeps), anything from 0.1 to 1 s on one SM, so with more cases
        // than SMs the order decides how long the last CTAs run alone.  A-priori key (nothing of an earlier solve is used): solver
        // class first (Gauss-Seidel / tangential / normal only), then the load (normal force, or approach when that is prescribed).
        if (g.second.size() > (size_t) launch_blocks((int) g.second.size())) {
            auto cost = [&](size_t k) {
                const Problem &q = *probs[k];
                const double cls = q.tang == 3 ? 2.0 : (q.tang != 0 ? 1.0 : 0.0);
                return std::make_pair(cls, q.norm == 1 ? q.fntrue : pen0[k]);
            };
            std::stable_sort(g.second.begin(), g.second.end(), [&](size_t a, size_t b) { return cost(a) > cost(b); });
        }
        const std::vector<size_t> &ks = g.second;
        const int n = (int) ks.size(), npot = cs.mx * cs.my;
        const ConvPlan &P = cs.hp.p;
        auto fail = [&](int code) { for (size_t k : ks) ierr[k] = code; };
        bool any_tang = false;
        for (size_t k : ks) any_tang = any_tang || probs[k]->tang != 0;
        // coefficient transforms needed by this group
        int rc = build_prec(cs, 0, 3);
        if (!rc) rc = build_chat(cs, SET_CS, 3, 3, 0);
        if (!rc) rc = build_chat(cs, SET_MS, 3, 3, 0);
        if (!rc && any_tang) {
            for (int ik = 1; ik <= 2 && !rc; ik++) { rc = build_prec(cs, 0, ik); if (!rc) rc = build_chat(cs, SET_MS, ik, ik, 0); }
            for (int ik = 1; ik <= 2 && !rc; ik++) for (int jk = 1; jk <= 2 && !rc; jk++) rc = build_chat(cs, SET_CS, ik, jk, 0);
        }
        if (!rc && cs.nt_cpl) {
            for (int t = 1; t <= 2 && !rc; t++) { rc = build_chat(cs, SET_CS, 3, t, 0); if (!rc && any_tang) rc = build_chat(cs, SET_CS, t, 3, 0); }
            if (cs.key.is_roll) for (int t = 1; t <= 2 && !rc; t++) rc = build_chat(cs, SET_CV, t, 3, 0);
        }
        bool any_trans = false;                                   // transient rolling: A'_tt p'_t with the shifted coefficients
        for (size_t k : ks) any_trans = any_trans || probs[k]->tang == 2;
        if (!rc && any_trans) for (int ik = 1; ik <= 2 && !rc; ik++) for (int jk = 1; jk <= 2 && !rc; jk++) rc = build_chat(cs, SET_CV, ik, jk, 0);
        if (!rc && cs.key.is_roll)                                  // ConvexGS in steady rolling works with csv = cs - cv
            for (int ik = 1; ik <= 2 && !rc; ik++) for (int jk = 1; jk <= 2 && !rc; jk++) rc = build_chat(cs, SET_CSV, ik, jk, 0);
        if (!rc) rc = build_levels(cs, 0, any_tang);            // after the full-size transforms and preconditioners exist
        if (rc) { fail(rc); continue; }
        cudaDeviceSynchronize();
        auto tg1 = now();
        bt[1] += secs(tg0, tg1);
        // device buffers: per case hs_n(1) hst(2) ps(3) ss(2) work(9) twork(24) pv(3) = 44 n doubles, el n ints
        // (+ 16 n of GDsteady work space when a case of the group asks for G = 5)
        bool any_gd = false;
        for (size_t k : ks) any_gd = any_gd || (probs[k]->tang == 3 && probs[k]->gausei == 5);
        const size_t per = (size_t) (any_gd ? 60 : 44) * npot;
        // pinned staging is bounded (48 MB of case records): page-locking hundreds of MB costs more than the transfers, and
        // far more when several ranks of one host do it at once; the cases move in groups of `gcap`
        const int gcap = std::max(1, std::min(n, (int) ((size_t) (48u << 20) / (sizeof(double) * 6 * npot))));
        if (!pool_dev(BP.d_buf, BP.c_buf, per * n) || !pool_dev(BP.d_el, BP.c_el, (size_t) n * npot) ||
            !pool_dev(BP.d_cases, BP.c_cases, (size_t) n) || (!BP.d_next && cudaMalloc(&BP.d_next, sizeof(int)) != cudaSuccess) ||
            !pool_dev(BP.d_us, BP.c_us, (size_t) 3 * npot * n) || !pool_dev(BP.d_pb, BP.c_pb, (size_t) 3 * npot * n) ||
            !pool_pinned(BP.h_fld, BP.c_hfld, (size_t) 6 * npot * gcap) || !pool_pinned(BP.h_us, BP.c_hus, (size_t) 3 * npot * gcap) ||
            !pool_pinned(BP.h_el, BP.c_hel, (size_t) npot * gcap)) {
            last_error() = "device allocation failed"; fail(CNTC_err_other);
            continue;
        }
        double *d_buf = BP.d_buf; int *d_el = BP.d_el, *d_next = BP.d_next; ContactCase *d_cases = BP.d_cases;
        cudaMemsetAsync(d_buf, 0, sizeof(double) * per * n, 0);
        std::vector<ContactCase> hc(n);
        const size_t nblk = (size_t) 4 * cs.mx * cs.my;
        double c00[2] = { 0, 0 };
        cudaMemcpy(&c00[0], cs.d_cf[SET_CS] + 0 * nblk + (size_t) cs.my * 2 * cs.mx + cs.mx, sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(&c00[1], cs.d_cf[SET_CS] + 4 * nblk + (size_t) cs.my * 2 * cs.mx + cs.mx, sizeof(double), cudaMemcpyDeviceToHost);
        for (int g0 = 0; g0 < n; g0 += gcap) {
        const int gm = std::min(gcap, n - g0);
        host_parallel_for(gm, [&](int j) {                           // per-case records and pinned staging: independent
            const int i = g0 + j;
            Problem &p = *probs[ks[i]];
            double *base = d_buf + per * i;
            ContactCase &c = hc[i];
            memset(&c, 0, sizeof(c));
            NormCase &nc = c.nrm;
            nc.hs = base; c.hst = base + npot; c.ps = base + 3 * (size_t) npot; c.ss = base + 6 * (size_t) npot;
            nc.work = base + 8 * (size_t) npot; c.twork = base + 17 * (size_t) npot;
            nc.pn = c.ps + 2 * (size_t) npot; nc.el = d_el + (size_t) i * npot;
            nc.chatA = cs.d_chat[SET_CS][2][2]; nc.chatM = cs.d_chat[SET_MS][2][2];
            nc.chatA31 = cs.nt_cpl ? cs.d_chat[SET_CS][2][0] : nullptr; nc.chatA32 = cs.nt_cpl ? cs.d_chat[SET_CS][2][1] : nullptr;
            nc.ptx = cs.nt_cpl ? c.ps : nullptr; nc.pty = cs.nt_cpl ? c.ps + npot : nullptr;
            nc.cf33 = cs.d_cf[SET_CS] + 8 * nblk; nc.cmx = cs.mx; nc.cmy = cs.my; nc.ga_inv = cs.ga_inv;
            nc.ic_norm = p.norm; nc.maxgs = p.maxgs; nc.maxin = p.maxin; nc.eps = p.eps; nc.dxdy = p.dx * p.dy;
            nc.pen = pen0[ks[i]]; nc.fntrue = p.fntrue;
            nc.lev = cs.d_lev; nc.nlx = cs.nlx; nc.nly = cs.nly; nc.stage_bytes = cs.stage_bytes();
            c.tang = p.tang; c.force3 = p.force3; c.maxnr = p.maxnr; c.maxout = p.maxout;
            c.cksi = p.cksi; c.ceta = p.ceta; c.fxrel = p.fxrel; c.fyrel = p.fyrel; c.fstat = p.fstat;
            if (!p.solved || p.iestim == 0 || p.iestim == 2) { if (p.force3 >= 1) c.cksi = 1e-6; if (p.force3 == 2) c.ceta = 0.0; }   // m_sdis.f90:760-762
            bool pv_nonzero = false;
            for (double v : p.pv) if (v != 0.0) { pv_nonzero = true; break; }
            c.pv = (pv_nonzero && (p.tang == 1 || p.tang == 2)) ? base + 41 * (size_t) npot : nullptr;
            if (c.pv) cudaMemcpy(base + 41 * (size_t) npot, p.pv.data(), sizeof(double) * 3 * npot, cudaMemcpyHostToDevice);
            double chi_e, dq_e;
            roll_stepsize(p, chi_e, dq_e);
            const bool is_roll = cs.key.is_roll != 0;
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { c.chatA[a][b] = cs.d_chat[SET_CS][a][b]; c.chatV[a][b] = cs.d_chat[is_roll ? SET_CV : SET_CS][a][b]; }
            for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) c.chatSV[a][b] = is_roll ? cs.d_chat[SET_CSV][a][b] : nullptr;
            if (is_roll) { c.cfv11 = cs.d_cf[SET_CSV] + 0 * nblk; c.cfv12 = cs.d_cf[SET_CSV] + 3 * nblk; c.cfv22 = cs.d_cf[SET_CSV] + 4 * nblk; }
            c.cf13 = cs.nt_cpl ? cs.d_cf[SET_CS] + 6 * nblk : nullptr; c.cf23 = cs.nt_cpl ? cs.d_cf[SET_CS] + 7 * nblk : nullptr;
            c.cf12 = cs.d_cf[SET_CS] + 3 * nblk; c.dq = dq_e; c.dx = p.dx; c.gausei = p.gausei; c.omegah = p.omegah; c.omegas = p.omegas;
            c.chatM11 = cs.d_chat[SET_MS][0][0]; c.chatM22 = cs.d_chat[SET_MS][1][1];
            if (p.tang == 3 && p.gausei == 5) {                          // cntc_setsolverflags, contact_addon.f90:1220-1250
                c.gwork = base + 44 * (size_t) npot;
                GdParams &sp = c.gd;
                sp.fdecay = p.gd_fdecay; sp.betath = p.gd_betath; sp.kdowfb = p.gd_kdowfb;
                sp.d_ifc = std::max(0.01, p.gd_difc); sp.d_lin = p.gd_dlin; sp.d_cns = std::max(0.01, p.gd_dcns);
                sp.d_slp = std::max(0.01, p.gd_dslp); sp.pow_s = std::max(0.01, std::min(10.0, p.gd_pows));
                if (sp.d_lin * (sp.d_cns - sp.d_ifc) < 0.0) { sp.d_lin = 0.0; sp.d_cns = sp.d_ifc; }
                sp.gd_meth = p.gd_meth; sp.kdown = p.gd_kdown;
            }
            c.cf11 = cs.d_cf[SET_CS] + 0 * nblk; c.cf22 = cs.d_cf[SET_CS] + 4 * nblk;
            c.c11 = c00[0]; c.c22 = c00[1]; c.ga = cs.ga;
            // host inputs: hs_n, hst (set_tang_rhs, m_sdis.f90:498-583; shifts: dq = 1), ps
            double *stage = BP.h_fld + (size_t) j * 6 * npot;       // pinned staging [gcap][6 npot], one strided upload per group
            std::fill(stage, stage + 6 * (size_t) npot, 0.0);
            std::copy(hs[ks[i]].begin(), hs[ks[i]].end(), stage);
            // rolling: spin pole shifted by facphi*dq along the rolling direction (facphi = 1/6, m_sinput.f90:789-793)
            const double dq = dq_e, facphi = 1.0 / 6.0;
            const double xofs = is_roll ? cos(chi_e) * dq * facphi : 0.0, yofs = is_roll ? sin(chi_e) * dq * facphi : 0.0;
            for (int iy = 0; iy < p.my; iy++) for (int ix = 0; ix < p.mx; ix++) {
                const int ii = iy * p.mx + ix;
                const double x = p.xc1 + ix * p.dx, y = p.yc1 + iy * p.dy;
                double wx = -(y + yofs) * p.cphi, wy = (x + xofs) * p.cphi;
                if (p.force3 == 0) wx += p.cksi;
                if (p.force3 <= 1) wy += p.ceta;
                stage[npot + ii] = -dq * wx; stage[2 * (size_t) npot + ii] = -dq * wy;
            }
            std::copy(p.ps.begin(), p.ps.end(), stage + 3 * (size_t) npot);
            std::copy(el0[ks[i]].begin(), el0[ks[i]].end(), BP.h_el + (size_t) j * npot);
        });
        cudaMemcpy2DAsync(d_buf + per * g0, sizeof(double) * per, BP.h_fld, sizeof(double) * 6 * npot, sizeof(double) * 6 * npot, gm, cudaMemcpyHostToDevice, 0);
        cudaMemcpyAsync(d_el + (size_t) g0 * npot, BP.h_el, sizeof(int) * (size_t) npot * gm, cudaMemcpyHostToDevice, 0);
        cudaStreamSynchronize(0);                                   // the staging is refilled by the next group
        }
        cudaMemcpy(d_cases, hc.data(), sizeof(ContactCase) * n, cudaMemcpyHostToDevice);
        cudaMemset(d_next, 0, sizeof(int));
        auto tg2 = now();
        bt[2] += secs(tg1, tg2);
        NormBatch &NB = norm_batch();                               // events around the solver kernel(s): cb200_snorm_kernel_ms
        if (!NB.ev0) { cudaEventCreate(&NB.ev0); cudaEventCreate(&NB.ev1); }
        cudaEventRecord(NB.ev0, 0);
        cudaError_t le = cudaSuccess;
        if (cs.hp.fits) {
            // the Gauss-Seidel sweep arrays need room behind the FFT layout on small grids: only when such a case is present
            bool any_gs = false;
            for (size_t k : ks) any_gs = any_gs || probs[k]->tang == 3 || (probs[k]->tang != 0 && probs[k]->gausei == 2);
            const size_t smem = (size_t) P.smem_bytes + (any_gs ? steady_extra_smem(P) : 0);
            if (smem > (size_t) kSmemMax) {
                last_error() = "grid shape needs more shared memory than one CTA has for the Gauss-Seidel sweep (thin or 2-D grid)";
                fail(CNTC_err_discr);
                continue;
            }
            k_contac_batch<<<launch_blocks(n), CB_THREADS, smem>>>(P, d_cases, n, d_next);
            le = cudaGetLastError();
            engine().launches++;
        } else {
            LargeCtx X;
            X.L = cs.lp; X.T = cs.d_T; X.gpart = cs.d_gpart; X.prof = nullptr;
            // Gauss-Seidel cases (T = 3 with G != 5, G = 2, or the fall-back of a stagnating GDsteady): row arrays of the sweep in
            // the phase buffers of the product when they fit in front of the reduction scratch, behind them otherwise
            const size_t gs_fixed = steady_fixed_bytes(cs.mx, cs.my);
            X.gs_off = gs_fixed + 1024 <= (size_t) cs.lp.smem_bytes ? 0 : cs.lp.smem_bytes;
            const size_t lsmem = (size_t) cs.lp.smem_bytes + (X.gs_off ? gs_fixed : 0);
            if (lsmem > (size_t) kSmemMax) {
                last_error() = "grid rows too long for the shared-memory row arrays of the Gauss-Seidel sweep";
                fail(CNTC_err_discr);
                continue;
            }
            for (int i = 0; i < n; i++) {                              // one cooperative whole-GPU launch per case
                ContactCase *cp = d_cases + i;
                void *args[] = { (void *) &X, (void *) &cp };
                const cudaError_t ce = cudaLaunchCooperativeKernel((void *) k_lg_contac, dim3(engine().num_sms), dim3(CB_THREADS), args, lsmem, 0);
                if (ce != cudaSuccess && le == cudaSuccess) le = ce;
                engine().launches++;
            }
        }
        cudaEventRecord(NB.ev1, 0);
        cudaError_t e = cudaDeviceSynchronize();
        auto tg3 = now();
        bt[3] += secs(tg2, tg3);
        if (le != cudaSuccess) { last_error() = std::string("solver kernel launch: ") + cudaGetErrorString(le); fail(CNTC_err_other); }
        else if (e != cudaSuccess) { last_error() = std::string("k_contac_batch: ") + cudaGetErrorString(e); fail(CNTC_err_other); }
        else {
            cudaMemcpy(hc.data(), d_cases, sizeof(ContactCase) * n, cudaMemcpyDeviceToHost);
            // soutpt (m_soutpt.f90:378-398): us = A ps on the contact area, all directions
            double *d_us = BP.d_us, *d_pb = BP.d_pb;
            cudaMemsetAsync(d_us, 0, sizeof(double) * 3 * (size_t) npot * n, 0);
            // tractions of all cases [n][3][npot] (the case records are `per` doubles apart)
            cudaMemcpy2DAsync(d_pb, sizeof(double) * 3 * npot, d_buf + 3 * (size_t) npot, sizeof(double) * per, sizeof(double) * 3 * npot, n,
                              cudaMemcpyDeviceToDevice, 0);
            rc = vecaijpj_dev(cs, SET_CS, n, -8, any_tang ? -3 : 3, any_tang ? -3 : 3, d_pb, d_el, d_us, 0);
            if (rc) { fail(rc); continue; }
            auto to1 = now();
            bt[6] += secs(tg3, to1);
            // per group of cases: one strided download of ps (3 npot) + ss (2 npot), one of us and one of the element divisions
            for (int g0 = 0; g0 < n; g0 += gcap) {
                const int gm = std::min(gcap, n - g0);
                cudaMemcpyAsync(BP.h_us, d_us + (size_t) g0 * 3 * npot, sizeof(double) * 3 * (size_t) npot * gm, cudaMemcpyDeviceToHost, 0);
                cudaMemcpy2DAsync(BP.h_fld, sizeof(double) * 5 * npot, d_buf + per * g0 + 3 * (size_t) npot, sizeof(double) * per, sizeof(double) * 5 * npot, gm,
                                  cudaMemcpyDeviceToHost, 0);
                cudaMemcpyAsync(BP.h_el, d_el + (size_t) g0 * npot, sizeof(int) * (size_t) npot * gm, cudaMemcpyDeviceToHost, 0);
                cudaStreamSynchronize(0);
                host_parallel_for(gm, [&](int j) {
                    Problem &p = *probs[ks[g0 + j]];
                    const double *f = BP.h_fld + (size_t) j * 5 * npot;
                    p.el.assign(BP.h_el + (size_t) j * npot, BP.h_el + (size_t) (j + 1) * npot);
                    p.ps.assign(f, f + 3 * (size_t) npot);
                    p.ss.assign(3 * (size_t) npot, 0.0);
                    if (p.tang != 0) std::copy(f + 3 * (size_t) npot, f + 5 * (size_t) npot, p.ss.begin());
                    p.us.assign(BP.h_us + (size_t) j * 3 * npot, BP.h_us + (size_t) (j + 1) * 3 * npot);
                });
            }
            auto to2 = now();
            bt[7] += secs(to1, to2);
            std::atomic<int> unserved(0);
            host_parallel_for(n, [&](int i) {
                Problem &p = *probs[ks[i]];
                const ContactCase &c = hc[i];
                if (p.tang >= 2) p.dq_eff = c.dq;
                p.hs.assign(3 * (size_t) npot, 0.0);
                std::copy(hs[ks[i]].begin(), hs[ks[i]].end(), p.hs.begin() + 2 * (size_t) npot);
                p.pen = c.nrm.pen; p.fntrue = c.nrm.fntrue; p.itcg = c.nrm.itcg; p.itnorm = c.nrm.itnorm; p.ncon = c.nrm.ncon;
                p.status = c.nrm.status; p.ittang = c.ittang; p.itgs = c.itgs; p.itout = c.itout; p.nadh = c.nadh; p.nslip = c.nslip;
                for (int k = 0; k < 16; k++) { p.pan_dif[k] = c.pan_dif[k]; p.pan_difid[k] = c.pan_difid[k]; }
                p.nr_itcg.assign(c.nr_itcg, c.nr_itcg + std::min(c.nr_n, (int) CB_MAXNR_LOG));
                p.gd_fallback = c.gd_fallback; p.gd_ntrial = c.gd_ntrial;
                for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) p.sens_nr[a][b] = c.sens[a][b];
                const double muscal = p.fstat;
                double sx = 0, sy = 0, mz = 0;
                for (int iy = 0; iy < p.my; iy++) for (int ix = 0; ix < p.mx; ix++) {
                    const int ii = iy * p.mx + ix;
                    const double x = p.xc1 + ix * p.dx, y = p.yc1 + iy * p.dy;
                    sx += p.ps[ii]; sy += p.ps[npot + ii]; mz += -p.ps[ii] * y + p.ps[npot + ii] * x;
                }
                const double dxdy = p.dx * p.dy;
                {   // m_soutpt.f90:438-457: moments of the pressures about the x and y axes, elastic energy 0.5e-3 dxdy (us, ps) on C,
                    // frictional power dxdy (ps, ss)_t / (1e3 dt) with the shift in the slip area only (:470-500)
                    double smx = 0, smy = 0, se = 0, sf = 0;
                    for (int iy = 0; iy < p.my; iy++) for (int ix = 0; ix < p.mx; ix++) {
                        const int ii = iy * p.mx + ix;
                        const double x = p.xc1 + ix * p.dx, y = p.yc1 + iy * p.dy, pn = p.ps[2 * (size_t) npot + ii];
                        smx += pn * y; smy += pn * x;
                        if (p.el[ii] >= 1) for (int k = 0; k < 3; k++) se += p.us[(size_t) k * npot + ii] * p.ps[(size_t) k * npot + ii];
                        if (p.el[ii] == 2) sf += p.ps[ii] * p.ss[ii] + p.ps[npot + ii] * p.ss[npot + ii];
                    }
                    p.mxtrue = dxdy * smx; p.mytrue = -dxdy * smy; p.elen = 0.5 * 1e-3 * dxdy * se;
                    const double dt = p.tang >= 2 ? (p.tang >= 2 && p.dq_eff > 0 ? p.dq_eff : p.dq) / std::max(1e-30, p.veloc) : p.dt;
                    p.frpow = p.tang != 0 ? dxdy * sf / (1e3 * std::max(1e-30, dt)) : 0.0;
                }
                if (p.tang != 0) {                                       // m_soutpt.f90:424-450
                    if (p.force3 == 0) p.fxrel = dxdy * sx / (p.fntrue * muscal + 1e-20);
                    if (p.force3 <= 1) p.fyrel = dxdy * sy / (p.fntrue * muscal + 1e-20);
                    if (p.force3 >= 1) p.cksi = c.cksi;
                    if (p.force3 >= 2) p.ceta = c.ceta;
                    p.fcntc[0] = p.fxrel * (p.fntrue * muscal + 1e-20); p.fcntc[1] = p.fyrel * (p.fntrue * muscal + 1e-20);
                    p.mztrue = dxdy * mz;
                    if (fabs(p.mztrue) < 0.5 * p.eps * (muscal * p.fntrue + 1e-20)) p.mztrue = 0.0;
                } else { p.fcntc[0] = p.fcntc[1] = 0.0; p.mztrue = 0.0; }
                p.fcntc[2] = p.fntrue;
                p.solved = true;
                if (c.tstatus & 1) { unserved = 1; ierr[ks[i]] = CNTC_err_other; }
                else if (p.itnorm < 0 || (p.status & 1)) ierr[ks[i]] = CNTC_err_norm;
                else if (p.ittang < 0) ierr[ks[i]] = CNTC_err_tang;
                else ierr[ks[i]] = count_at_boundary(p);              // contact_addon.f90:3885-3891
            });
            if (unserved) last_error() = "TANG: the case needs a solver that the B200 path does not serve (a Gauss-Seidel solver -- also as the fall-back of a stagnating GDsteady -- on a grid beyond one CTA)";
            bt[8] += secs(to2, now());
        }
        auto tf0 = now();                                          // (buffers stay in the pool)
        bt[9] += secs(tf0, now());
        bt[4] += secs(tg3, now());
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    bt[5] = dt;
    for (size_t k = 0; k < nb; k++) { probs[k]->t_wall += dt / (double) nb; probs[k]->t_cpu += dt / (double) nb; }
}

}  // namespace cb200

extern "C" {

void cntc_initializefirst(int *ifcver, int *ierror, int *ioutput, const char *, const char *, const char *, int *, int *, int *)
{
    (void) ioutput;
    if (ifcver) *ifcver = 2400;                 // library interface version reported by the reference family
    if (ierror) *ierror = 0;
}

void cntc_initializefirst_new(int *ifcver, int *ierror, int *ioutput, const char *a, const char *b, const char *c, int *la, int *lb, int *lc)
{ cntc_initializefirst(ifcver, ierror, ioutput, a, b, c, la, lb, lc); }

void cntc_initialize(int *ire, int *imodul, int *ifcver, int *ierror, const char *, int *)
{
    if (ifcver) *ifcver = 2400;
    if (ierror) *ierror = 0;
    if (imodul && *imodul != 3) { if (ierror) *ierror = CNTC_err_other; last_error() = "only module 3 (contact on a given grid) is in the hot-path scope"; return; }
    Registry &R = registry();
    if (*ire < 1 || *ire > 999) { if (ierror) *ierror = -101; return; }
    std::lock_guard<std::mutex> lk(R.mu);
    auto &re = R.res[*ire];
    if (!re) re.reset(new ResultElement());
}

void cntc_setglobalflags(int *lenflg, int *params, int *values)
{
    for (int i = 0; i < *lenflg; i++) if (params[i] == CNTC_if_idebug) registry().idebug = values[i];
}

void cntc_setflags(int *ire, int *icp, int *lenflg, int *params, int *values)
{
    int ierr; Problem *p = activate(*ire, *icp, &ierr);
    if (!p) return;
    auto clampi = [](int v, int lo, int hi) { return std::max(lo, std::min(hi, v)); };
    for (int i = 0; i < *lenflg; i++) {
        const int c = params[i], v = values[i];
        if (c == 0) continue;
        else if (c == CNTC_if_units) select_units(p->scl, v);
        else if (c == CNTC_ic_pvtime) p->pvtime = clampi(v, 0, 3);
        else if (c == CNTC_ic_bound) p->bound = clampi(v, 0, 6);
        else if (c == CNTC_ic_tang) p->tang = clampi(v, 0, 3);
        else if (c == CNTC_ic_norm) p->norm = clampi(v, 0, 1);
        else if (c == CNTC_ic_force) p->force3 = clampi(v, 0, 2);
        else if (c == CNTC_ic_sens) p->sens = clampi(v, 0, 3);
        else if (c == CNTC_ic_inflcf) { if (v >= 0 && v <= 4) p->gencr = v; }
        else if (c == CNTC_ic_exrhs) p->rztang = clampi(v, 0, 3);
        else if (c == CNTC_ic_iestim) p->iestim = clampi(v, 0, 3);
        else if (c == CNTC_ic_matfil) p->matfil = (v < 0 || v > 2) ? 0 : v;
        else if (c == CNTC_ic_output) p->output = clampi(v, 0, 5);
        else if (c == CNTC_ic_flow) p->flow = clampi(v, 0, 9);
        else if (c == CNTC_ic_return) p->ret = clampi(v, 0, 3);
        else if (c == CNTC_if_wrtinp) p->wrtinp = v;
        else if (c == CNTC_if_ncase) p->ncase = std::max(0, v - 1);
        else if (c == CNTC_if_idebug) registry().idebug = v;
        /* other codes (config, discns, npomax, ifmeth, sbsout, ... ) concern module 1 or file output: accepted, ignored */
    }
}

void cntc_getflags(int *ire, int *icp, int *lenflg, int *params, int *values)
{
    int ierr; Problem *p = activate(*ire, *icp, &ierr);
    if (!p) return;
    for (int i = 0; i < *lenflg; i++) {
        const int c = params[i];
        int v = 0;
        if (c == CNTC_if_units) v = p->scl.units; else if (c == CNTC_ic_pvtime) v = p->pvtime;
        else if (c == CNTC_ic_bound) v = p->bound; else if (c == CNTC_ic_tang) v = p->tang;
        else if (c == CNTC_ic_norm) v = p->norm; else if (c == CNTC_ic_force) v = p->force3;
        else if (c == CNTC_ic_frclaw) v = p->frclaw; else if (c == CNTC_ic_discns) v = p->discns;
        else if (c == CNTC_ic_inflcf) v = p->gencr; else if (c == CNTC_ic_mater) v = p->mater;
        else if (c == CNTC_ic_exrhs) v = p->rztang; else if (c == CNTC_ic_iestim) v = p->iestim;
        else if (c == CNTC_ic_output) v = p->output; else if (c == CNTC_ic_flow) v = p->flow;
        else if (c == CNTC_ic_return) v = p->ret; else if (c == CNTC_ic_matfil) v = p->matfil;
        else if (c == CNTC_ic_sens) v = p->sens; else if (c == CNTC_if_ncase) v = p->ncase;
        else if (c == CNTC_if_wrtinp) v = p->wrtinp; else if (c == CNTC_if_idebug) v = registry().idebug;
        values[i] = v;
    }
}

void cntc_setmetadata(int *ire, int *icp, int *, int *, double *) { int e; activate(*ire, *icp, &e); }

void cntc_setsolverflags(int *ire, int *icp, int *gdigit, int *nints, int *iparam, int *nreals, double *rparam)
{
    int ierr; Problem *p = activate(*ire, *icp, &ierr);
    if (!p) return;
    static const int ni[7] = { 4, 0, 5, 5, 5, 5, 1 }, nr[7] = { 1, 0, 4, 4, 2, 8, 1 };
    const int g = *gdigit;
    if (g < 0 || g > 6 || *nints != ni[g] || *nreals != nr[g]) { last_error() = "cntc_setsolverflags: wrong G-digit or parameter count"; return; }
    if (g >= 0 && g <= 5) p->gausei = g;
    if (g != 1 && g != 6) {
        p->eps = std::max(1e-20, rparam[0]);
        p->maxgs = std::max(1, iparam[0]); p->maxin = std::max(1, iparam[1]);
        p->maxnr = std::max(1, iparam[2]); p->maxout = std::max(1, iparam[3]);
        if (g == 2 || g == 3) { p->omegah = std::max(1e-20, rparam[1]); p->omegas = std::max(1e-20, rparam[2]); }   // contact_addon.f90:1203-1208
        if (g == 5) {                                                                                                // :1220-1250
            p->gd_fdecay = rparam[1]; p->gd_betath = rparam[2]; p->gd_kdowfb = iparam[4];
            p->gd_difc = rparam[3]; p->gd_dlin = rparam[4]; p->gd_dcns = rparam[5]; p->gd_dslp = rparam[6]; p->gd_pows = rparam[7];
            if (p->gd_fdecay >= 0.999) p->gd_meth = 1;                                   // E_trl
            else if (p->gd_fdecay <= 0.001) { p->gd_meth = 2; p->gd_kdown = std::max(1, (int) nearbyint(-p->gd_fdecay)); }   // E_down(k)
            else p->gd_meth = 3;                                                         // E_keep(f)
        }
    }
}

void cntc_setmaterialparameters(int *ire, int *icp, int *mdigit, int *nparam, double *rparam)
{
    int ierr; Problem *p = activate(*ire, *icp, &ierr);
    if (!p) return;
    static const int np[8] = { 4, 8, 8, 7, 8, 7, 4, 4 };
    if (*mdigit < 0 || *mdigit > 7 || *nparam != np[*mdigit]) { last_error() = "cntc_setmaterialparameters: wrong M-digit or parameter count"; return; }
    p->mater = *mdigit;
    p->mat.poiss[0] = rparam[0]; p->mat.poiss[1] = rparam[1];
    p->mat.gg[0] = rparam[2] * p->scl.forc / p->scl.area;
    p->mat.gg[1] = rparam[3] * p->scl.forc / p->scl.area;
}

void cntc_settimestep(int *ire, int *icp, double *dt) { int e; Problem *p = activate(*ire, *icp, &e); if (p) p->dt = *dt; }

void cntc_setreferencevelocity(int *ire, int *icp, double *veloc)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) p->veloc = *veloc * p->scl.veloc; }

void cntc_setrollingstepsize(int *ire, int *icp, double *chi, double *dq)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) { p->chi = *chi * p->scl.angle; p->dq = *dq * p->scl.len; } }

void cntc_setfrictionmethod(int *ire, int *icp, int *imeth, int *nparam, double *params)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    p->frclaw = *imeth;
    if (*imeth == 0 && *nparam >= 2) { p->fstat = params[0]; p->fkin = params[1]; }
}

void cntc_sethertzcontact(int *ire, int *icp, int *ipotcn, int *nparam, double *prm)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const int t = *ipotcn;
    if (t < -6 || t > -1 || *nparam != (t == -6 ? 6 : 5)) { last_error() = "cntc_sethertzcontact: wrong IPOTCN or parameter count"; return; }
    p->ipotcn = t;
    p->mx = std::max(1, (int) lround(prm[0])); p->my = std::max(1, (int) lround(prm[1]));
    if (t == -3) { p->hz_aa = std::max(1e-6, prm[2]) * p->scl.len; p->hz_bb = std::max(1e-6, prm[3]) * p->scl.len; }
    else if (t == -2) { p->hz_a1 = std::max(1e-12, prm[2]) / p->scl.len; p->hz_aob = std::max(1e-6, prm[3]); }
    else if (t == -1) { p->hz_a1 = std::max(1e-12, prm[2]) / p->scl.len; p->hz_b1 = std::max(1e-12, prm[3]) / p->scl.len; }
    p->hz_scale = std::max(1e-6, prm[4]);
}

void cntc_setpotcontact(int *ire, int *icp, int *ipotcn, int *nparam, double *prm)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const int t = *ipotcn;
    if (t < 1 || t > 4 || *nparam != 6) { last_error() = "cntc_setpotcontact: wrong IPOTCN or parameter count"; return; }
    const double L = p->scl.len;
    p->ipotcn = t;
    p->mx = std::max(1, (int) lround(prm[0])); p->my = std::max(1, (int) lround(prm[1]));
    if (t == 1) { p->xl = prm[2] * L; p->yl = prm[3] * L; p->dx = std::max(1e-12, prm[4]) * L; p->dy = std::max(1e-12, prm[5]) * L; }
    else if (t == 2) { p->xl = prm[2] * L; p->yl = prm[3] * L; p->xh = prm[4] * L; p->yh = prm[5] * L; }
    else if (t == 3) { p->xc1 = prm[2] * L; p->yc1 = prm[3] * L; p->dx = std::max(1e-12, prm[4]) * L; p->dy = std::max(1e-12, prm[5]) * L; }
    else { p->xc1 = prm[2] * L; p->yc1 = prm[3] * L; p->xcm = prm[4] * L; p->ycm = prm[5] * L; }
    potcon_fill(*p);
}

void cntc_setpenetration(int *ire, int *icp, double *pen)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) { p->norm = 0; p->pen = *pen * p->scl.len; } }

void cntc_setnormalforce(int *ire, int *icp, double *fn)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) { p->norm = 1; p->fntrue = *fn; } }

void cntc_setundeformeddistc(int *ire, int *icp, int *ibase, int *nparam, double *prm)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const double L = p->scl.len;
    const int b = *ibase, npot = p->mx * p->my;
    int nn = 0, need;
    if (b == 2) nn = (int) lround(prm[0]);
    if (b == 1) need = 6; else if (b == 2) need = 5 + nn; else if (b == 3) need = 8; else if (b == 9) need = npot;
    else { last_error() = "cntc_setundeformeddistc: invalid IBASE"; return; }
    if (*nparam != need) { last_error() = "cntc_setundeformeddistc: wrong parameter count"; return; }
    p->ibase = b;
    if (b == 1) {
        p->prmudf.assign(10, 0.0);
        for (int i = 0; i < 3; i++) p->prmudf[i] = prm[i] / L;
        p->prmudf[3] = prm[3]; p->prmudf[4] = prm[4]; p->prmudf[5] = prm[5] * L;
    } else if (b == 2) {
        p->prmudf.assign(5 + nn, 0.0);
        p->nn = nn;
        p->prmudf[0] = nn;
        for (int i = 1; i < 5 + nn; i++) p->prmudf[i] = prm[i] * L;
    } else if (b == 3) {
        p->prmudf.assign(10, 0.0);
        p->prmudf[0] = prm[0] * L; p->prmudf[1] = prm[1] / L; p->prmudf[2] = prm[2] * L; p->prmudf[3] = prm[3] * L;
        p->prmudf[4] = prm[4] / L; p->prmudf[5] = prm[5] * L; p->prmudf[6] = prm[6] * L; p->prmudf[7] = prm[7] * L;
    } else {
        p->prmudf.assign(npot + 10, 0.0);
        for (int i = 0; i < npot; i++) p->prmudf[i] = prm[i] * L;
    }
}

void cntc_setcreepages(int *ire, int *icp, double *vx, double *vy, double *phi)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    // contact_addon.f90:2491-2530: shifts [length] for T=1, creepages [-] for T=2,3; sign by body convention
    const bool shift = (p->tang == 1);
    p->cksi = *vx * (shift ? p->scl.len : 1.0);
    p->ceta = *vy * (shift ? p->scl.len : 1.0);
    p->cphi = shift ? *phi : *phi / p->scl.len;
}

void cntc_settangentialforces(int *ire, int *icp, double *fx, double *fy)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    p->fxrel = *fx; p->fyrel = *fy;                       // relative to fstat*fn, contact_addon.f90:2597-2650
    const double fabs_ = sqrt(p->fxrel * p->fxrel + p->fyrel * p->fyrel);
    if (fabs_ > 1.0) { p->fxrel /= fabs_; p->fyrel /= fabs_; }
}

// initial element division and approach estimate of eldiv0 (m_sdis.f90:818-1007) for kernel-level callers
int cb200_eldiv0(int mx, int my, double dx, double dy, double gg1, double gg2, double poiss1, double poiss2,
                 int ibase, const double *prmudf, int ic_norm, double fn, double pen_in, const double *h, int *el,
                 double *pen_out)
{
    Problem p;
    p.mx = mx; p.my = my; p.dx = dx; p.dy = dy; p.ibase = ibase; p.norm = ic_norm; p.fntrue = fn;
    p.mat.gg[0] = gg1; p.mat.gg[1] = gg2; p.mat.poiss[0] = poiss1; p.mat.poiss[1] = poiss2;
    combine_material(p.mat);
    if (prmudf) p.prmudf.assign(prmudf, prmudf + (ibase == 9 ? mx * my : 8));
    std::vector<double> hv(h, h + (size_t) mx * my);
    std::vector<int> e;
    double pen = pen_in;
    initial_eldiv(p, hv, e, pen);
    std::copy(e.begin(), e.end(), el);
    *pen_out = pen;
    return 0;
}

// iteration counters of the last case: out[0..6] = itnorm, itcg (NormCG), ittang, itgs (tangential solver iterations),
// ncon, number of tangential solver calls nr_n, outer iterations; nr_itcg[0..nr_n) = iterations per solver call (at most lenarr)
int cb200_batch_timing(double *out) { for (int k = 0; k < 6; k++) out[k] = batch_timing()[k]; return 0; }
int cb200_batch_timing_output(double *out) { for (int k = 0; k < 4; k++) out[k] = batch_timing()[6 + k]; return 0; }

int cb200_get_iterations(int ire, int icp, int *out, int lenarr, int *nr_itcg)
{
    int e; Problem *p = activate(ire, icp, &e);
    if (!p) return e;
    out[0] = p->itnorm; out[1] = p->itcg; out[2] = p->ittang; out[3] = p->itgs; out[4] = p->ncon; out[5] = (int) p->nr_itcg.size(); out[6] = p->itout;
    out[7] = p->gd_fallback > 0 ? -p->gd_ntrial - 1 : p->gd_ntrial;
    for (int i = 0; i < lenarr && i < (int) p->nr_itcg.size(); i++) nr_itcg[i] = p->nr_itcg[i];
    return 0;
}

// Replace the stored solution of a problem (element division and tractions [3][npot]: x, y, n) -- the state that the next case
// of a sequence starts from (I and P digits).  The reference keeps this state inside gd and offers no setter; the parity tests use
// it to run every stage of a sequence from the oracle's previous stage, so that differences cannot accumulate over the stages.
int cb200_set_state(int ire, int icp, int npot, const int *el, const double *ps)
{
    int e; Problem *p = activate(ire, icp, &e);
    if (!p) return e;
    if (npot != p->mx * p->my) { last_error() = "cb200_set_state: npot does not match the grid of the problem"; return CNTC_err_input; }
    p->el.assign(el, el + npot);
    p->ps.assign(ps, ps + 3 * (size_t) npot);
    p->solved = true;
    return 0;
}

// soutpt scalars that the reference writes to its .out file only (m_soutpt.f90:424-500): out = Fn, Fx, Fy, Mx, My, Mz, elastic
// energy [J], frictional power [W], largest pressure, deformed distance is returned by cb200_get_deformed_distance
int cb200_get_soutpt(int ire, int icp, int lenarr, double *out)
{
    int e; Problem *p = activate(ire, icp, &e);
    if (!p) return e;
    double pmax = 0.0;
    const size_t npot = (size_t) p->mx * p->my;
    if (p->ps.size() == 3 * npot) for (size_t i = 0; i < npot; i++) pmax = std::max(pmax, p->ps[2 * npot + i]);
    const double v[9] = { p->fcntc[2], p->fcntc[0], p->fcntc[1], p->mxtrue, p->mytrue, p->mztrue, p->elen, p->frpow, pmax };
    for (int k = 0; k < lenarr && k < 9; k++) out[k] = v[k];
    return 0;
}

// deformed distance hs - pen + us_n of all elements (m_soutpt.f90:459-468; the reference stores it in ss(:,n))
int cb200_get_deformed_distance(int ire, int icp, int lenarr, double *out)
{
    int e; Problem *p = activate(ire, icp, &e);
    if (!p) return e;
    const size_t npot = (size_t) p->mx * p->my;
    if (p->hs.size() != 3 * npot || p->us.size() != 3 * npot) { last_error() = "no solution available (run cntc_calculate first)"; return CNTC_err_other; }
    for (size_t i = 0; i < npot && (int) i < lenarr; i++) out[i] = p->hs[2 * npot + i] - p->pen + p->us[2 * npot + i];
    return 0;
}

int cb200_get_outer_history(int ire, int icp, int lenarr, double *dif, double *difid)
{
    int e; Problem *p = activate(ire, icp, &e);
    if (!p) return e;
    const int n = std::min(16, p->itout);
    for (int k = 0; k < n && k < lenarr; k++) { dif[k] = p->pan_dif[k]; difid[k] = p->pan_difid[k]; }
    return n;
}

void cntc_calculate(int *ire, int *icp, int *ierror)
{
    Problem *p = activate(*ire, *icp, ierror);
    if (!p) return;
    std::vector<Problem *> v(1, p);
    std::vector<int> ie;
    int rc = engine_init();
    if (rc) { *ierror = rc; return; }
    calculate_batch(v, ie);
    *ierror = ie[0];
}

} // extern "C" (helpers of the multi-device fan-out follow)
inline std::vector<int> &fanout_list() { static std::vector<int> f; return f; }
inline bool &fanout_env_done() { static bool b = false; return b; }
inline std::vector<int> fanout_devices()
{
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!fanout_env_done()) {
        fanout_env_done() = true;
        const char *e = getenv("CONTACT_B200_DEVICES");
        int ndev = 0;
        if (e && *e && cudaGetDeviceCount(&ndev) == cudaSuccess) {
            std::vector<int> &F = fanout_list();
            std::string s(e);
            if (s == "all") { for (int d = 0; d < ndev && d < CB_MAX_DEVICES; d++) F.push_back(d); }
            else if (s.find(',') == std::string::npos) { const int n = atoi(e); for (int d = 0; d < n && d < ndev && d < CB_MAX_DEVICES; d++) F.push_back(d); }
            else {
                size_t a = 0;
                while (a < s.size()) {
                    size_t b = s.find(',', a); if (b == std::string::npos) b = s.size();
                    const int d = atoi(s.substr(a, b - a).c_str());
                    if (d >= 0 && d < ndev && d < CB_MAX_DEVICES) F.push_back(d);
                    a = b + 1;
                }
            }
        }
    }
    return fanout_list();
}
extern "C" {

void cntc_calculate_batch(int *nre, int *ire, int *icp, int *ierror)
{
    std::vector<Problem *> v;
    std::vector<int> pos;
    for (int k = 0; k < *nre; k++) {
        Problem *p = activate(ire[k], *icp, &ierror[k]);
        if (p) { v.push_back(p); pos.push_back(k); }
    }
    int rc = engine_init();
    if (rc) { for (int k = 0; k < *nre; k++) ierror[k] = rc; return; }
    std::vector<int> ie;
    const std::vector<int> devs = fanout_devices();
    const size_t nd = std::min(devs.size(), v.size() / 8);          // a device is worth its set-up from a handful of cases on
    if (nd <= 1) calculate_batch(v, ie);
    else {
        // The batched case scheduler across the GPUs of one box: contiguous shards of the case list, one host thread per device,
        // no inter-GPU traffic (every device builds and caches the coefficient transforms of the classes it meets); the results
        // land in the callers' Problem records, i.e. the "final gather" is the join.
        int dev0 = 0;
        cudaGetDevice(&dev0);
        ie.assign(v.size(), 0);
        std::vector<std::string> errs(nd);
        std::vector<std::thread> th;
        for (size_t d = 0; d < nd; d++)
            th.emplace_back([&, d] {
                const size_t lo = v.size() * d / nd, hi = v.size() * (d + 1) / nd;
                std::vector<Problem *> part(v.begin() + lo, v.begin() + hi);
                std::vector<int> pe;
                int r = cudaSetDevice(devs[d]) == cudaSuccess ? engine_init() : CNTC_err_other;
                if (r) pe.assign(part.size(), r);
                else calculate_batch(part, pe);
                for (size_t i = 0; i < part.size(); i++) ie[lo + i] = pe[i];
                errs[d] = last_error();                              // thread-local: hand the message to the caller's thread
            });
        for (auto &t : th) t.join();
        cudaSetDevice(dev0);
        for (size_t d = 0; d < nd; d++) if (!errs[d].empty()) last_error() = errs[d];
    }
    for (size_t i = 0; i < v.size(); i++) ierror[pos[i]] = ie[i];
}

/* Devices that cntc_calculate_batch spreads a batch over: n <= 0 or 1 = the calling thread's current device only (default);
 * n > 1 = devices 0 .. n-1 (or the list devs[0..n-1] when given).  Also settable by the environment variable
 * CONTACT_B200_DEVICES = "all" | n | "0,2,3" read at the first batch.  Returns the number of devices in use. */
int cb200_set_devices(int n, const int *devs)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { last_error() = "contact_addon_b200: no CUDA device available"; return -99; }
    std::vector<int> &F = fanout_list();
    F.clear();
    for (int k = 0; k < n; k++) {
        const int d = devs ? devs[k] : k;
        if (d < 0 || d >= ndev || d >= CB_MAX_DEVICES) { F.clear(); last_error() = "cb200_set_devices: no such device"; return CNTC_err_input; }
        F.push_back(d);
    }
    fanout_env_done() = true;
    return F.empty() ? 1 : (int) F.size();
}

void cntc_getnumelements(int *ire, int *icp, int *mx, int *my)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) { *mx = p->mx; *my = p->my; } }

void cntc_getgriddiscretization(int *ire, int *icp, double *dx, double *dy)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) { *dx = p->dx / p->scl.len; *dy = p->dy / p->scl.len; } }

void cntc_getpotcontact(int *ire, int *icp, int *lenarr, double *v)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const double L = p->scl.len;
    const double out[6] = { (double) p->mx, (double) p->my, p->xc1 / L, p->yc1 / L, p->dx / L, p->dy / L };   // contact_addon.f90:5144-5187
    for (int i = 0; i < *lenarr && i < 6; i++) v[i] = out[i];
}

void cntc_getpenetration(int *ire, int *icp, double *pen)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) *pen = p->pen / p->scl.len; }

void cntc_getcreepages(int *ire, int *icp, double *vx, double *vy, double *phi)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const bool shift = (p->tang == 1);
    *vx = p->cksi / (shift ? p->scl.len : 1.0);
    *vy = p->ceta / (shift ? p->scl.len : 1.0);
    *phi = shift ? p->cphi : p->cphi * p->scl.len;
}

void cntc_getcontactforces(int *ire, int *icp, double *fn, double *tx, double *ty, double *mz)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    *fn = p->fcntc[2]; *tx = p->scl.body * p->fcntc[0]; *ty = p->scl.body * p->fcntc[1];
    *mz = p->scl.body * p->mztrue / p->scl.len;
}

void cntc_getcontactpatchareas(int *ire, int *icp, double *carea, double *harea, double *sarea)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    int nadh = 0, nslip = 0, nplast = 0;
    for (int v : p->el) { if (v == 1) nadh++; else if (v == 2) nslip++; else if (v == 3) nplast++; }
    const double dxdy = p->dx * p->dy;
    *carea = (double) (nadh + nslip + nplast) * dxdy / p->scl.area;
    *harea = (double) nadh * dxdy / p->scl.area;
    *sarea = (double) nslip * dxdy / p->scl.area;
}

void cntc_getelementdivision(int *ire, int *icp, int *lenarr, int *eldiv)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    for (int i = 0; i < *lenarr && i < (int) p->el.size(); i++) eldiv[i] = p->el[i];
}

void cntc_getmaximumpressure(int *ire, int *icp, double *pnmax)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const int npot = p->mx * p->my;
    double m = 0.0;
    if ((int) p->ps.size() == 3 * npot) for (int i = 0; i < npot; i++) m = std::max(m, fabs(p->ps[2 * (size_t) npot + i]));
    *pnmax = m * p->scl.area;
}

void cntc_getmaximumtraction(int *ire, int *icp, double *ptmax)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const int npot = p->mx * p->my;
    double m = 0.0;
    if ((int) p->ps.size() == 3 * npot)
        for (int i = 0; i < npot; i++) m = std::max(m, p->ps[i] * p->ps[i] + p->ps[npot + i] * p->ps[npot + i]);
    *ptmax = sqrt(m) * p->scl.area;
}

void cntc_getfielddata(int *ire, int *icp, int *ifld, int *lenarr, double *fld)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const int npot = p->mx * p->my;
    const std::vector<double> *src = nullptr; int col = 0; double scl = 1.0;
    const Scaling &s = p->scl;
    switch (*ifld) {
    case CNTC_fld_h:  src = &p->hs; col = 2; scl = 1.0 / s.len; break;
    case CNTC_fld_px: src = &p->ps; col = 0; scl = s.area * s.body; break;
    case CNTC_fld_py: src = &p->ps; col = 1; scl = s.area * s.body; break;
    case CNTC_fld_pn: src = &p->ps; col = 2; scl = s.area; break;
    case CNTC_fld_ux: src = &p->us; col = 0; scl = 1.0 / s.len; break;
    case CNTC_fld_uy: src = &p->us; col = 1; scl = 1.0 / s.len; break;
    case CNTC_fld_un: src = &p->us; col = 2; scl = 1.0 / s.len; break;
    case CNTC_fld_sx: src = &p->ss; col = 0; scl = s.body / (p->tang >= 2 ? p->dq_eff : 1.0); break;
    case CNTC_fld_sy: src = &p->ss; col = 1; scl = s.body / (p->tang >= 2 ? p->dq_eff : 1.0); break;
    default: break;
    }
    for (int i = 0; i < *lenarr && i < npot; i++) {
        if (*ifld == CNTC_fld_mu) fld[i] = p->fstat;
        else if (src && (int) src->size() == 3 * npot) fld[i] = scl * (*src)[(size_t) col * npot + i];
        else fld[i] = 0.0;
    }
}

void cntc_gettractions(int *ire, int *icp, int *lenarr, double *pn, double *px, double *py)
{
    int f;
    f = CNTC_fld_pn; cntc_getfielddata(ire, icp, &f, lenarr, pn);
    f = CNTC_fld_px; cntc_getfielddata(ire, icp, &f, lenarr, px);
    f = CNTC_fld_py; cntc_getfielddata(ire, icp, &f, lenarr, py);
}

void cntc_getmicroslip(int *ire, int *icp, int *lenarr, double *sx, double *sy)
{
    int f;
    f = CNTC_fld_sx; cntc_getfielddata(ire, icp, &f, lenarr, sx);
    f = CNTC_fld_sy; cntc_getfielddata(ire, icp, &f, lenarr, sy);
}

void cntc_getdisplacements(int *ire, int *icp, int *lenarr, double *un, double *ux, double *uy)
{
    int f;
    f = CNTC_fld_un; cntc_getfielddata(ire, icp, &f, lenarr, un);
    f = CNTC_fld_ux; cntc_getfielddata(ire, icp, &f, lenarr, ux);
    f = CNTC_fld_uy; cntc_getfielddata(ire, icp, &f, lenarr, uy);
}

// ---- the remaining entry points of matlab_intfc/contact_addon.h:7-148, so that a caller linked against the reference's
//      library resolves every symbol.  Module-3 getters are served from the problem's data; the entry points of the
//      wheel/rail module (category "m=1 only" in contact_addon.f90) log an error and return, which is what the reference
//      does when they are called on a module-3 result element (cntc_activate with the wrong module) ----
static void module1_only(const char *name)
{ last_error() = std::string(name) + ": available for module 1 (wheel/rail contact) only, which is outside the hot-path scope of this library"; }

void cntc_getparameters(int *ire, int *icp, int *itask, int *lenarr, double *values)
{   // contact_addon.f90:4164-4251: 1 kinematic constants used by plot3d, 2 material, 3 friction
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const int n = *lenarr;
    auto put = [&](int k, double v) { if (n >= k) values[k - 1] = v; };
    if (*itask == 1) {
        double chi, dq; roll_stepsize(*p, chi, dq);
        put(1, p->veloc / p->scl.veloc); put(2, chi / p->scl.angle); put(3, (p->solved && p->tang >= 2 ? p->dq_eff : dq) / p->scl.len);
        put(4, 0.0); put(5, 0.0); put(6, 0.0);                       // spin centre offsets, tau_c0: not used on this path
    } else if (*itask == 2) {
        Material m = p->mat; combine_material(m);
        put(1, m.gg[0] * p->scl.area); put(2, m.gg[1] * p->scl.area); put(3, m.ga * p->scl.area);
        put(4, m.poiss[0]); put(5, m.poiss[1]); put(6, m.nu); put(7, m.ak);
        for (int k = 8; k <= 22; k++) put(k, 0.0);                    // flexibilities, interfacial layer, damping: M = 0 only
    } else if (*itask == 3) {
        put(1, 0.0); put(2, 1.0); put(3, 0.0); put(4, 0.0); put(5, 0.0); put(6, p->fstat); put(7, p->fkin);   // L = 0: one set (fstat, fkin)
    }
}

void cntc_getreferencevelocity(int *ire, int *icp, double *veloc)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) *veloc = p->veloc / p->scl.veloc; }

void cntc_gethertzcontact(int *ire, int *icp, int *lenarr, double *values)
{   // contact_addon.f90:5021-5063
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const int n = *lenarr;
    const double L = p->scl.len;
    const double v[10] = { p->hz_a1 * L, p->hz_b1 * L, p->hz_aa / L, p->hz_bb / L, p->hz_rho / L, p->hz_cp / L, p->hz_scale,
                           p->hz_bb / L, p->hz_bb / L, p->hz_bb > 0.0 ? p->hz_aa / p->hz_bb : 0.0 };
    for (int k = 0; k < 10 && k < n; k++) values[k] = v[k];
}

void cntc_getmaximumtemperature(int *ire, int *icp, double *t1max, double *t2max)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) { *t1max = 0.0; *t2max = 0.0; } }      // H = 0: no temperature calculation

void cntc_getsensitivities(int *ire, int *icp, int *lenout, int *lenin, double *sens)
{   // contact_addon.f90:5919-6040: sens(lenout, lenin), outputs fn, fx, fy, mz; inputs pen, cksi, ceta, cphi; zeros where not
    // computed.  Filled here: d(fx, fy)/d(cksi, ceta) from the Newton-Raphson process on the creepages (F = 1, 2).
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    const int no = *lenout, ni = *lenin;
    for (int k = 0; k < no * ni; k++) sens[k] = 0.0;
    const double fin = (p->tang == 1) ? p->scl.len : 1.0;             // shifts: cksi, ceta are distances
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++)
        if (a + 1 < no && b + 1 < ni) sens[(a + 1) + (b + 1) * no] = p->sens_nr[a][b] * fin;
}

void cntc_resetcalculationtime(int *ire, int *icp)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) { p->t_cpu = 0.0; p->t_wall = 0.0; } }

void cntc_setextrarigidslip(int *ire, int *icp, int *lenarr, double *, double *)
{   // E = 9 (m_sinput.f90, extra term of the rigid slip per element): accepted, refused at cntc_calculate when E-digit is set
    int e; Problem *p = activate(*ire, *icp, &e); if (p) p->exrhs_len = *lenarr;
}

void cntc_settemperaturedata(int *ire, int *icp, int *imeth, int *, double *)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) p->heat_meth = *imeth; }                 // H-digit: stored, H >= 1 is refused

void cntc_readinpfile(int *ire, int *, const char *, int *, int *ierror)
{   // contact_addon.f90:729-...: the .inp reader of this library is contact_b200/inp.py (parse_inp / run_inp), which drives
    // these same entry points; there is no second reader in C++
    (void) ire;
    if (ierror) *ierror = CNTC_err_other;
    last_error() = "cntc_readinpfile: use contact_b200.inp.run_inp (the .inp reader of this library drives the cntc_* entry points from Python)";
}

void cntc_setverticalforce(int *, double *) { module1_only("cntc_setverticalforce"); }
void cntc_setprofileinputfname(int *, const char *, int *, int *, int *, int *, double *) { module1_only("cntc_setprofileinputfname"); }
void cntc_setprofileinputvalues(int *, int *, double *, int *, int *, int *, double *) { module1_only("cntc_setprofileinputvalues"); }
void cntc_settrackdimensions(int *, int *, int *, double *) { module1_only("cntc_settrackdimensions"); }
void cntc_setwheelsetdimensions(int *, int *, int *, double *) { module1_only("cntc_setwheelsetdimensions"); }
void cntc_setwheelsetposition(int *, int *, int *, double *) { module1_only("cntc_setwheelsetposition"); }
void cntc_setwheelsetvelocity(int *, int *, int *, double *) { module1_only("cntc_setwheelsetvelocity"); }
void cntc_setwheelsetflexibility(int *, int *, int *, double *) { module1_only("cntc_setwheelsetflexibility"); }
void cntc_getprofilevalues(int *, int *, int *, int *, int *, double *, int *, double *) { module1_only("cntc_getprofilevalues"); }
void cntc_getprofilevalues_new(int *, int *, int *, int *, int *, double *, int *, double *) { module1_only("cntc_getprofilevalues_new"); }
void cntc_getwheelsetposition(int *, int *, double *) { module1_only("cntc_getwheelsetposition"); }
void cntc_getwheelsetvelocity(int *, int *, double *) { module1_only("cntc_getwheelsetvelocity"); }
void cntc_getnumcontactpatches(int *, int *npatch) { if (npatch) *npatch = 0; module1_only("cntc_getnumcontactpatches"); }
void cntc_getcontactlocation(int *, int *, int *, double *) { module1_only("cntc_getcontactlocation"); }
void cntc_getglobalforces(int *, int *, int *, double *) { module1_only("cntc_getglobalforces"); }

void cntc_getcalculationtime(int *ire, int *icp, double *tcpu, double *twall)
{ int e; Problem *p = activate(*ire, *icp, &e); if (p) { *tcpu = p->t_cpu; *twall = p->t_wall; } }

void subs_addblock(int *ire, int *icp, int *iblk, int *isubs, int *nx, int *ny, int *nz, double *xparam, double *yparam, double *zparam)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    // contact_addon.f90:3330-3515.  isubs 1,2,3: zparam = [NZ, ZL, DZ]; 5,6,7,9: zparam = z(1:npz);
    // 2,6: x/yparam = [IL, INC, IH]; 3,7: index lists; 9: coordinate lists.  iblk = 0 clears all blocks.
    const int is = *isubs;
    for (auto it = p->subs.lower_bound(*iblk); it != p->subs.end();) it = p->subs.erase(it);
    if (*iblk <= 0) return;
    if (is < 1 || is == 4 || is == 8 || is > 9) { last_error() = "subs_addblock: ISUBS does not exist"; return; }
    Problem::SubsBlock &b = p->subs[*iblk];
    b.isubs = is;
    const double L = p->scl.len;
    b.x.clear(); b.y.clear(); b.z.clear(); b.table.clear();
    if (is <= 3) { const int n = (int) lround(zparam[0]); for (int k = 0; k < n; k++) b.z.push_back((zparam[1] + k * zparam[2]) * L); }
    else for (int k = 0; k < *nz; k++) b.z.push_back(zparam[k] * L);
    if (is == 9) { for (int i = 0; i < *nx; i++) b.x.push_back(xparam[i] * L); for (int j = 0; j < *ny; j++) b.y.push_back(yparam[j] * L); }
    else if (is == 2 || is == 6) {
        for (int i = (int) lround(xparam[0]); i <= (int) lround(xparam[2]); i += std::max(1, (int) lround(xparam[1]))) b.x.push_back(i);
        for (int j = (int) lround(yparam[0]); j <= (int) lround(yparam[2]); j += std::max(1, (int) lround(yparam[1]))) b.y.push_back(j);
    } else if (is == 3 || is == 7) {
        for (int i = 0; i < *nx; i++) b.x.push_back(lround(xparam[i]));
        for (int j = 0; j < *ny; j++) b.y.push_back(lround(yparam[j]));
    }
    b.nx = b.ny = 0; b.nz = (int) b.z.size();
}

void subs_calculate(int *ire, int *icp, int *ierror)
{
    Problem *pp = activate(*ire, *icp, ierror);
    if (!pp) return;
    Problem &p = *pp;
    int rc = engine_init();
    if (rc) { *ierror = rc; return; }
    const int npot = p.mx * p.my;
    if ((int) p.ps.size() != 3 * npot) { last_error() = "subs_calculate: no tractions available (run cntc_calculate first)"; *ierror = CNTC_err_other; return; }
    combine_material(p.mat);
    for (auto &kv : p.subs) {
        Problem::SubsBlock &b = kv.second;
        if (b.isubs == 9) {
            b.nx = (int) b.x.size(); b.ny = (int) b.y.size();
            const int np = b.nx * b.ny * b.nz;
            std::vector<double> xyz((size_t) 3 * np), t18((size_t) 18 * np);
            int ip = 0;
            for (int k = 0; k < b.nz; k++) for (int j = 0; j < b.ny; j++) for (int i = 0; i < b.nx; i++, ip++) { xyz[3 * ip] = b.x[i]; xyz[3 * ip + 1] = b.y[j]; xyz[3 * ip + 2] = b.z[k]; }
            rc = cb200_subsurf_points(p.mx, p.my, p.xc1, p.yc1, p.dx, p.dy, p.mat.gg[0], p.mat.gg[1], p.mat.poiss[0], p.mat.poiss[1],
                                      p.ps.data(), np, xyz.data(), t18.data());
            if (rc) { *ierror = rc; return; }
            b.table.assign((size_t) 21 * np, 0.0);
            for (int q = 0; q < np; q++) { for (int c = 0; c < 3; c++) b.table[(size_t) q * 21 + c] = xyz[3 * q + c]; for (int c = 0; c < 18; c++) b.table[(size_t) q * 21 + 3 + c] = t18[(size_t) q * 18 + c]; }
        } else {
            CoefSet *cs = nullptr;
            rc = get_coefset(p.mx, p.my, p.dx, p.dy, p.mat, 0, 0.0, 1.0, 0, &cs);
            if (rc) { *ierror = rc; return; }
            std::vector<double> tbl((size_t) b.nz * npot * 18);
            rc = cb200_subsurf_batch(handle_of(cs), 1, b.nz, b.z.data(), p.mat.gg[0], p.mat.gg[1], p.mat.poiss[0], p.mat.poiss[1], p.ps.data(), tbl.data());
            if (rc) { *ierror = rc; return; }
            std::vector<int> ixs, iys;
            if (b.isubs == 1 || b.isubs == 5) { for (int i = 1; i <= p.mx; i++) ixs.push_back(i); for (int j = 1; j <= p.my; j++) iys.push_back(j); }
            else { for (double v : b.x) if (v >= 1 && v <= p.mx) ixs.push_back((int) v); for (double v : b.y) if (v >= 1 && v <= p.my) iys.push_back((int) v); }
            b.nx = (int) ixs.size(); b.ny = (int) iys.size();
            const int np = b.nx * b.ny * b.nz;
            b.table.assign((size_t) 21 * np, 0.0);
            int q = 0;
            for (int k = 0; k < b.nz; k++) for (int j : iys) for (int i : ixs) {
                const int ii = (i - 1) + (j - 1) * p.mx;
                double *row = &b.table[(size_t) q * 21];
                row[0] = p.xc1 + (i - 1) * p.dx; row[1] = p.yc1 + (j - 1) * p.dy; row[2] = b.z[k];
                for (int c = 0; c < 18; c++) row[3 + c] = tbl[((size_t) k * npot + ii) * 18 + c];
                q++;
            }
        }
    }
    *ierror = 0;
}

void subs_getblocksize(int *ire, int *icp, int *iblk, int *nx, int *ny, int *nz)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    auto it = p->subs.find(*iblk);
    if (it == p->subs.end()) { *nx = *ny = *nz = 0; return; }
    Problem::SubsBlock &b = it->second;
    if (b.nx == 0 && b.ny == 0) {          // before the calculation: sizes implied by the specification
        if (b.isubs == 1 || b.isubs == 5) { *nx = p->mx; *ny = p->my; } else { *nx = (int) b.x.size(); *ny = (int) b.y.size(); }
    } else { *nx = b.nx; *ny = b.ny; }
    *nz = b.nz;
}

void subs_getresults(int *ire, int *icp, int *iblk, int *lenarr, int *ncol, int *icol, double *values)
{
    int e; Problem *p = activate(*ire, *icp, &e);
    if (!p) return;
    auto it = p->subs.find(*iblk);
    if (it == p->subs.end() || it->second.table.empty()) { last_error() = "subs_getresults: no results for this block"; return; }
    const Problem::SubsBlock &b = it->second;
    const int np = b.nx * b.ny * b.nz, n = std::min(np, *lenarr);
    for (int jc = 0; jc < *ncol; jc++) {
        double *col = values + (size_t) jc * *lenarr;
        const int c = icol[jc];
        if (c <= 0 || c > 21) { for (int i = 0; i < *lenarr; i++) col[i] = -999.0; continue; }     // contact_addon.f90:6150-6152
        const double f = (c <= 6) ? 1.0 / p->scl.len : p->scl.area;
        for (int i = 0; i < n; i++) col[i] = b.table[(size_t) i * 21 + (c - 1)] * f;
    }
}

void cntc_finalize(int *ire)
{
    Registry &R = registry();
    std::lock_guard<std::mutex> lk(R.mu);
    R.res.erase(*ire);
}

void cntc_finalizelast(void)
{
    Registry &R = registry();
    std::lock_guard<std::mutex> lk(R.mu);
    R.res.clear();
    BatchPool &BP = batch_pool();                              // release the work space of calculate_batch
    std::lock_guard<std::mutex> pl(BP.mu);
    cudaFree(BP.d_buf); cudaFree(BP.d_us); cudaFree(BP.d_pb); cudaFree(BP.d_el); cudaFree(BP.d_next); cudaFree(BP.d_cases);
    cudaFreeHost(BP.h_fld); cudaFreeHost(BP.h_us); cudaFreeHost(BP.h_el);
    BP.d_buf = BP.d_us = BP.d_pb = BP.h_fld = BP.h_us = nullptr; BP.d_el = BP.d_next = BP.h_el = nullptr; BP.d_cases = nullptr;
    BP.c_buf = BP.c_us = BP.c_pb = BP.c_el = BP.c_cases = BP.c_hfld = BP.c_hus = BP.c_hel = 0;
}

}  // extern "C"
