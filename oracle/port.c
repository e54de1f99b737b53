/* oracle/port.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C CPU restatement of the reference's algorithm for the orbit-integration hot path.  Each
 * function cites the reference lines it follows (paths relative to /root/reference/src/gala/).
 * It is pinned against (a) the reference's own C++ compiled unmodified (oracle/_ref/libgala_ref.so)
 * and (b) the reference's known-answer tests / golden vectors in tests/test_oracle_cpu.py.
 * It needs no reference sources, so it can be rebuilt on the GPU box.
 *
 * Compiled twice: REAL = double (the restatement proper) and REAL = long double
 * (-DPORT_LONG_DOUBLE=1, exported with the same names from libgala_port_ld.so): an extended-
 * precision "truth" run of the same algorithm that arbitrates which FP64 build is closer.
 * The ABI always carries double arrays.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gala_b200.h"

#if PORT_LONG_DOUBLE
typedef long double REAL;
#define SQRT sqrtl
#define LOG logl
#define POW powl
#define SIN sinl
#define COS cosl
#define FABS fabsl
#define ATAN2 atan2l
#define ATAN atanl
#define EXP expl
#define LGAMMA lgammal
#define TGAMMA tgammal
#define CEIL ceill
#else
typedef double REAL;
#define SQRT sqrt
#define LOG log
#define POW pow
#define SIN sin
#define COS cos
#define FABS fabs
#define ATAN2 atan2
#define ATAN atan
#define EXP exp
#define LGAMMA lgamma
#define TGAMMA tgamma
#define CEIL ceil
#endif
#define PI_R ((REAL)3.14159265358979323846264338327950288L)

/* ------------------------------------------------------------------------------------------------
 * potentials: value / gradient (accumulating) / density at one point.
 * potential/potential/builtin/builtin_potentials.cpp, line ranges per function.
 * ---------------------------------------------------------------------------------------------- */
static REAL norm3(const REAL *q) { return SQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]); }

/* Kepler :56-85 */
static REAL kepler_value(const double *p, const REAL *q) { return -(REAL)p[0] * p[1] / norm3(q); }
static void kepler_grad(const double *p, const REAL *q, REAL *g) {
    REAL fac = (REAL)p[0] * p[1] / POW(norm3(q), 3);
    g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * q[2];
}
static REAL kepler_density(const double *p, const REAL *q) {
    return (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] == 0) ? INFINITY : 0;
}
/* Isochrone :128-160 */
static REAL isochrone_value(const double *p, const REAL *q) {
    REAL r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    return -(REAL)p[0] * p[1] / (SQRT(r2 + (REAL)p[2] * p[2]) + p[2]);
}
static void isochrone_grad(const double *p, const REAL *q, REAL *g) {
    REAL s = SQRT((q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) + (REAL)p[2] * p[2]);
    REAL denom = s * (s + p[2]) * (s + p[2]);
    REAL fac = (REAL)p[0] * p[1] / denom;
    g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * q[2];
}
static REAL isochrone_density(const double *p, const REAL *q) {
    REAL b = p[2], r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], a = SQRT(b * b + r2);
    return p[1] * (3 * (b + a) * a * a - r2 * (b + 3 * a)) / (4 * PI_R * POW(b + a, 3) * a * a * a);
}
/* Hernquist :211-244 */
static REAL hernquist_value(const double *p, const REAL *q) { return -(REAL)p[0] * p[1] / (norm3(q) + p[2]); }
static void hernquist_grad(const double *p, const REAL *q, REAL *g) {
    REAL r = norm3(q);
    REAL fac = (REAL)p[0] * p[1] / ((r + p[2]) * (r + p[2]) * r);
    g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * q[2];
}
static REAL hernquist_density(const double *p, const REAL *q) {
    REAL r = norm3(q);
    REAL rho0 = p[1] / (2 * PI_R * p[2] * p[2] * p[2]);
    return rho0 / ((r / p[2]) * POW(1 + r / p[2], 3));
}
/* Plummer :292-320 */
static REAL plummer_value(const double *p, const REAL *q) {
    return -(REAL)p[0] * p[1] / SQRT((q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) + (REAL)p[2] * p[2]);
}
static void plummer_grad(const double *p, const REAL *q, REAL *g) {
    REAL R2b = (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) + (REAL)p[2] * p[2];
    REAL fac = (REAL)p[0] * p[1] / SQRT(R2b) / R2b;
    g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * q[2];
}
static REAL plummer_density(const double *p, const REAL *q) {
    REAL r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    return 3 * p[1] / (4 * PI_R * p[2] * p[2] * p[2]) * POW(1 + r2 / ((REAL)p[2] * p[2]), -2.5);
}
/* Jaffe :365-393 */
static REAL jaffe_value(const double *p, const REAL *q) { return -(REAL)p[0] * p[1] / p[2] * LOG(1 + p[2] / norm3(q)); }
static void jaffe_grad(const double *p, const REAL *q, REAL *g) {
    REAL r = norm3(q);
    REAL fac = (REAL)p[0] * p[1] / p[2] * (p[2] / (r * (p[2] + r))) / r;
    g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * q[2];
}
static REAL jaffe_density(const double *p, const REAL *q) {
    REAL r = norm3(q);
    REAL rho0 = p[1] / (4 * PI_R * p[2] * p[2] * p[2]);
    return rho0 / (POW(r / p[2], 2) * POW(1 + r / p[2], 2));
}
/* NFW: spherical :819-864, flattened :925-962, triaxial :1028-1072 */
static REAL nfw_u(int type, const double *p, const REAL *q) {
    if (type == GB_POT_NFW_SPHERICAL) return norm3(q) / p[2];
    if (type == GB_POT_NFW_FLATTENED) return SQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] / ((REAL)p[5] * p[5])) / p[2];
    return SQRT(q[0] * q[0] / ((REAL)p[3] * p[3]) + q[1] * q[1] / ((REAL)p[4] * p[4]) + q[2] * q[2] / ((REAL)p[5] * p[5])) / p[2];
}
static REAL nfw_value(int type, const double *p, const REAL *q) {
    REAL v_h2 = -(REAL)p[0] * p[1] / p[2];
    REAL u = nfw_u(type, p, q);
    return (u == 0) ? v_h2 : v_h2 * LOG(1 + u) / u;
}
static void nfw_grad(int type, const double *p, const REAL *q, REAL *g) {
    REAL v_h2 = (REAL)p[0] * p[1] / p[2];
    REAL u = nfw_u(type, p, q);
    REAL fac = v_h2 / (u * u * u) / ((REAL)p[2] * p[2]) * (LOG(1 + u) - u / (1 + u));
    if (type == GB_POT_NFW_SPHERICAL) {
        g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * q[2];
    } else if (type == GB_POT_NFW_FLATTENED) {
        g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * q[2] / ((REAL)p[5] * p[5]);
    } else {
        g[0] += fac * q[0] / ((REAL)p[3] * p[3]); g[1] += fac * q[1] / ((REAL)p[4] * p[4]);
        g[2] += fac * q[2] / ((REAL)p[5] * p[5]);
    }
}
static REAL nfw_density(int type, const double *p, const REAL *q) {
    if (type != GB_POT_NFW_SPHERICAL) return NAN;     /* nan_density: cybuiltin.pyx:303-321 */
    REAL v_h2 = (REAL)p[0] * p[1] / p[2];
    REAL r = norm3(q);
    REAL rho0 = v_h2 / (4 * PI_R * p[0] * p[2] * p[2]);
    return rho0 / ((r / p[2]) * POW(1 + r / p[2], 2));
}
/* Miyamoto-Nagai :1287-1332 */
static REAL mn_value(REAL G, REAL m, REAL a, REAL b, const REAL *q) {
    REAL zd = a + SQRT(q[2] * q[2] + b * b);
    return -G * m / SQRT(q[0] * q[0] + q[1] * q[1] + zd * zd);
}
static void mn_grad(REAL G, REAL m, REAL a, REAL b, const REAL *q, REAL *g) {
    REAL sqrtz = SQRT(q[2] * q[2] + b * b);
    REAL zd = a + sqrtz;
    REAL fac = G * m * POW(q[0] * q[0] + q[1] * q[1] + zd * zd, -1.5);
    g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * q[2] * (1. + a / sqrtz);
}
static REAL mn_density(REAL M, REAL a, REAL b, const REAL *q) {
    REAL R2 = q[0] * q[0] + q[1] * q[1];
    REAL s = SQRT(q[2] * q[2] + b * b);
    REAL numer = (b * b * M / (4 * PI_R)) * (a * R2 + (a + 3 * s) * (a + s) * (a + s));
    REAL denom = POW(R2 + (a + s) * (a + s), 2.5) * s * s * s;
    return numer / denom;
}
/* Long-Murali bar :1681-1811 (density: Laplacian of the closed-form potential) */
static void lmbar_xyz(const double *p, const REAL *q, REAL *x, REAL *y, REAL *ca, REAL *sa) {
    *ca = COS((REAL)p[5]); *sa = SIN((REAL)p[5]);
    *x = q[0] * *ca + q[1] * *sa;
    *y = -q[0] * *sa + q[1] * *ca;
}
static REAL lmbar_value(const double *p, const REAL *q) {
    REAL x, y, ca, sa, z = q[2], a = p[2], b = p[3], c = p[4];
    lmbar_xyz(p, q, &x, &y, &ca, &sa);
    REAL Tm = SQRT((a - x) * (a - x) + y * y + POW(b + SQRT(c * c + z * z), 2));
    REAL Tp = SQRT((a + x) * (a + x) + y * y + POW(b + SQRT(c * c + z * z), 2));
    return (REAL)p[0] * p[1] / (2 * a) * LOG((x - a + Tm) / (x + a + Tp));
}
static void lmbar_grad(const double *p, const REAL *q, REAL *g) {
    REAL x, y, ca, sa, z = q[2], a = p[2], b = p[3], c = p[4];
    lmbar_xyz(p, q, &x, &y, &ca, &sa);
    REAL bcz = b + SQRT(c * c + z * z);
    REAL Tm = SQRT((a - x) * (a - x) + y * y + bcz * bcz);
    REAL Tp = SQRT((a + x) * (a + x) + y * y + bcz * bcz);
    REAL fac1 = (REAL)p[0] * p[1] / (2 * Tm * Tp);
    REAL fac2 = 1 / (y * y + bcz * bcz);
    REAL fac3 = Tp + Tm - (4 * x * x) / (Tp + Tm);
    REAL gx = 4 * fac1 * x / (Tp + Tm);
    REAL gy = fac1 * y * fac2 * fac3;
    REAL gz = fac1 * z * fac2 * fac3 * bcz / SQRT(z * z + c * c);
    g[0] += (gx * ca - gy * sa); g[1] += (gx * sa + gy * ca); g[2] += gz;
}
static REAL lmbar_density(const double *p, const REAL *q) {
    /* central-difference Laplacian of lmbar_value in extended steps is avoided: analytic second
     * derivatives of Phi = K ln((x-a+Tm)/(x+a+Tp)) */
    REAL x, y, ca, sa, z = q[2], a = p[2], b = p[3], c = p[4];
    lmbar_xyz(p, q, &x, &y, &ca, &sa);
    REAL zc = SQRT(c * c + z * z), B = b + zc, s2 = y * y + B * B;
    REAL Tm = SQRT((a - x) * (a - x) + s2), Tp = SQRT((a + x) * (a + x) + s2);
    REAL um = x - a + Tm, up = x + a + Tp;
    REAL Pxx = (a - x) / (Tm * Tm * Tm) + (a + x) / (Tp * Tp * Tp);
    REAL fm = 1 / (Tm * um), fp = 1 / (Tp * up);
    REAL hm = fm / (Tm * Tm) + fm * fm, hp = fp / (Tp * Tp) + fp * fp;
    REAL Pyy = (fm - y * y * hm) - (fp - y * y * hp);
    REAL PBB = (fm - B * B * hm) - (fp - B * B * hp);
    REAL PB = B * (fm - fp);
    REAL Bp = z / zc, Bpp = c * c / (zc * zc * zc);
    return p[1] / (8 * PI_R * a) * (Pxx + Pyy + PBB * Bp * Bp + PB * Bpp);
}

/* ---- SCF (potential/scf/src/bfe.cpp:64-259, bfe_helper.cpp:14-90) with the special functions the
 * reference takes from GSL restated by their textbook recurrences (GSL is a system dependency of
 * the reference, absent here; conventions: gsl_sf_legendre_Plm carries the Condon-Shortley phase). */
static REAL gegen(int n, REAL lam, REAL x) {
    if (n < 0) return 0;
    if (n == 0) return 1;
    REAL a = 1, b = 2 * lam * x;
    for (int k = 2; k <= n; k++) { REAL c = (2 * (k + lam - 1) * x * b - (k + 2 * lam - 2) * a) / k; a = b; b = c; }
    return b;
}
static REAL plm(int l, int m, REAL x) {
    if (m < 0 || m > l) return 0;
    REAL pmm = 1;
    if (m > 0) { REAL s = SQRT((1 - x) * (1 + x)), f = 1; for (int i = 1; i <= m; i++) { pmm *= -f * s; f += 2; } }
    if (l == m) return pmm;
    REAL pmmp1 = x * (2 * m + 1) * pmm;
    if (l == m + 1) return pmmp1;
    REAL pll = 0;
    for (int ll = m + 2; ll <= l; ll++) { pll = (x * (2 * ll - 1) * pmmp1 - (ll + m - 1) * pmm) / (ll - m); pmm = pmmp1; pmmp1 = pll; }
    return pll;
}
static REAL fact_ratio(int l, int m) { REAL r = 1; for (int k = l - m + 1; k <= l + m; k++) r /= k; return r; } /* (l-m)!/(l+m)! */
static REAL sphplm(int l, int m, REAL x) { return SQRT((2 * l + 1) / (4 * PI_R) * fact_ratio(l, m)) * plm(l, m, x); }
#define SQRT_FOURPI ((REAL)3.544907701811031L)
static REAL phi_nl(REAL s, int n, int l) {   /* bfe_helper.cpp:24-26 */
    return -SQRT_FOURPI * POW(s, l) * POW(1 + s, -2 * l - 1) * gegen(n, 2 * l + 1.5, (s - 1) / (s + 1));
}
static REAL rho_nl(REAL s, int n, int l) {   /* bfe_helper.cpp:14-19 */
    REAL Knl = 0.5 * n * (n + 4 * l + 3) + (l + 1) * (2 * l + 1);
    REAL RR = Knl / (2 * PI_R) * POW(s, l) / (s * POW(1 + s, 2 * l + 3)) * gegen(n, 2 * l + 1.5, (s - 1) / (s + 1));
    return SQRT_FOURPI * RR;
}
static void sph_grad_phi_nlm(REAL s, REAL X, int n, int l, int m, REAL *sg) {   /* bfe_helper.cpp:30-90 */
    REAL sintheta = SQRT(1 - X * X);
    REAL Phi_nl = phi_nl(s, n, l), Ylm = sphplm(l, m, X), Plm = (m <= l) ? plm(l, m, X) : 0, dPhinl_dr;
    if (n == 0)
        dPhinl_dr = SQRT_FOURPI * POW(s, -1 + l) * POW(1 + s, -3 - 2 * l) * (1 + s) * (l * (-1 + s) + s);
    else
        dPhinl_dr = (SQRT_FOURPI * POW(s, -1 + l) * POW(1 + s, -3 - 2 * l) *
                     (-2 * (3 + 4 * l) * s * gegen(-1 + n, 2.5 + 2 * l, (-1 + s) / (1 + s)) +
                      (1 + s) * (l * (-1 + s) + s) * gegen(n, 1.5 + 2 * l, (-1 + s) / (1 + s))));
    dPhinl_dr *= Ylm;
    REAL dY = 0;
    if (l != 0) {
        REAL Pl1m = (m <= l - 1) ? plm(l - 1, m, X) : 0;
        REAL A = SQRT((REAL)(2 * l + 1)) / SQRT_FOURPI * SQRT(fact_ratio(l, m));
        dY = A / sintheta * (l * X * Plm - (l + m) * Pl1m);
    }
    sg[0] = dPhinl_dr;
    sg[1] = dY * Phi_nl / s;
    sg[2] = ((m == 0) ? 0 : (REAL)m) * Ylm * Phi_nl;
}
/* p = [G, nmax, lmax, m, r_s, S..., T...]  (bfe.cpp:229-258) */
static void scf_unpack(const double *p, int *nmax, int *lmax, const double **S, const double **T) {
    *nmax = (int)p[1]; *lmax = (int)p[2];
    int nc = (*nmax + 1) * (*lmax + 1) * (*lmax + 1);
    *S = p + 5; *T = p + 5 + nc;
}
static REAL scf_value(const double *p, const REAL *q) {      /* scf_potential_helper, bfe.cpp:64-112 */
    int nmax, lmax; const double *S, *T; scf_unpack(p, &nmax, &lmax, &S, &T);
    REAL r = norm3(q), s = r / p[4], X = q[2] / r, phi = ATAN2(q[1], q[0]), val = 0;
    for (int n = 0; n <= nmax; n++) for (int l = 0; l <= lmax; l++) for (int m = 0; m <= l; m++) {
        int i = m + (lmax + 1) * (l + (lmax + 1) * n);
        if (S[i] == 0. && T[i] == 0.) continue;
        val += phi_nl(s, n, l) * sphplm(l, m, X) * (S[i] * COS(m * phi) + T[i] * SIN(m * phi));
    }
    return val * p[0] * p[3] / p[4];
}
static REAL scf_density(const double *p, const REAL *q) {    /* scf_density_helper, bfe.cpp:15-62 */
    int nmax, lmax; const double *S, *T; scf_unpack(p, &nmax, &lmax, &S, &T);
    REAL r = norm3(q), s = r / p[4], X = q[2] / r, phi = ATAN2(q[1], q[0]), val = 0;
    for (int n = 0; n <= nmax; n++) for (int l = 0; l <= lmax; l++) for (int m = 0; m <= l; m++) {
        int i = m + (lmax + 1) * (l + (lmax + 1) * n);
        if (S[i] == 0. && T[i] == 0.) continue;
        val += rho_nl(s, n, l) * sphplm(l, m, X) * (S[i] * COS(m * phi) + T[i] * SIN(m * phi));
    }
    return val * p[3] / ((REAL)p[4] * p[4] * p[4]);
}
static void scf_grad(const double *p, const REAL *q, REAL *g) {   /* scf_gradient_helper, bfe.cpp:114-189 */
    int nmax, lmax; const double *S, *T; scf_unpack(p, &nmax, &lmax, &S, &T);
    REAL r = norm3(q), s = r / p[4], X = q[2] / r, phi = ATAN2(q[1], q[0]);
    REAL sintheta = SQRT(1 - X * X), cosphi = COS(phi), sinphi = SIN(phi);
    REAL t2[3] = {0, 0, 0}, sg[3];
    for (int n = 0; n <= nmax; n++) for (int l = 0; l <= lmax; l++) for (int m = 0; m <= l; m++) {
        int i = m + (lmax + 1) * (l + (lmax + 1) * n);
        if (S[i] == 0. && T[i] == 0.) continue;
        REAL cm = COS(m * phi), sm = SIN(m * phi);
        REAL tmp = S[i] * cm + T[i] * sm;
        sph_grad_phi_nlm(s, X, n, l, m, sg);
        t2[0] += sg[0] * tmp;
        t2[1] += sg[1] * tmp;
        t2[2] += sg[2] * (T[i] * cm - S[i] * sm) / (s * sintheta);
    }
    REAL gx = sintheta * cosphi * t2[0] + X * cosphi * t2[1] - sinphi * t2[2];
    REAL gy = sintheta * sinphi * t2[0] + X * sinphi * t2[1] + cosphi * t2[2];
    REAL gz = X * t2[0] - sintheta * t2[1];
    REAL sc = (REAL)p[0] * p[3] / ((REAL)p[4] * p[4]);
    g[0] += gx * sc; g[1] += gy * sc; g[2] += gz * sc;
}

/* ---- Multipole (potential/potential/builtin/multipole.cpp:30-301), term by term like the reference:
 * p = [G, lmax, num_coeff, inner, m, r_s, S00, T00, S10, T10, S11, T11, ...] (:247-262) */
static REAL mp_phi_l(REAL s, int l, int inner) { return inner ? POW(s, l) : POW(s, -(l + 1)); }   /* :50-56 */
static void mp_sph_grad_phi_lm(REAL s, REAL X, int l, int m, int inner, REAL *sg) {             /* :69-138 */
    REAL sintheta = SQRT(1 - X * X);
    REAL Ylm = sphplm(l, m, X), Plm = (m <= l) ? plm(l, m, X) : 0, Phi_l = mp_phi_l(s, l, inner), dPhil_dr, dY = 0;
    if (inner) dPhil_dr = l * POW(s, l - 1) * Ylm; else dPhil_dr = -(l + 1) * POW(s, -l - 2) * Ylm;
    if (l != 0) {
        REAL Pl1m = (m <= l - 1) ? plm(l - 1, m, X) : 0;
        REAL A = SQRT((REAL)(2 * l + 1)) / SQRT_FOURPI * SQRT(fact_ratio(l, m));
        dY = (sintheta != 0) ? A / sintheta * (l * X * Plm - (l + m) * Pl1m) : 0;
    }
    if (s > 0) { sg[0] = dPhil_dr; sg[1] = dY * Phi_l / s; sg[2] = ((m == 0) ? 0 : (REAL)m) * Ylm * Phi_l; }
    else { sg[0] = sg[1] = sg[2] = 0; }
}
static REAL mp_value(const double *p, const REAL *q) {       /* mp_potential_helper :185-226 */
    int lmax = (int)p[1], inner = (int)p[3];
    REAL r = norm3(q), s = r / p[5], X = q[2] / r, phi = ATAN2(q[1], q[0]), val = 0;
    int i = 0;
    for (int l = 0; l <= lmax; l++) for (int m = 0; m <= l; m++, i++) {
        double S = p[6 + 2 * i], T = p[7 + 2 * i];
        if (S == 0. && T == 0.) continue;
        val += mp_phi_l(s, l, inner) * sphplm(l, m, X) * (S * COS(m * phi) + T * SIN(m * phi));
    }
    if (r == 0 && inner) val = 0;
    return val * p[0] * p[4] / p[5];
}
static void mp_grad(const double *p, const REAL *q, REAL *g) {    /* mp_gradient_helper :228-301 */
    int lmax = (int)p[1], inner = (int)p[3];
    REAL r = norm3(q), s = r / p[5], X = q[2] / r, phi = ATAN2(q[1], q[0]);
    REAL sintheta = SQRT(1 - X * X), cosphi = COS(phi), sinphi = SIN(phi);
    REAL t2[3] = {0, 0, 0}, sg[3];
    int i = 0;
    for (int l = 0; l <= lmax; l++) for (int m = 0; m <= l; m++, i++) {
        double S = p[6 + 2 * i], T = p[7 + 2 * i];
        if (S == 0. && T == 0.) continue;
        REAL cm = COS(m * phi), sm = SIN(m * phi), tmp = S * cm + T * sm;
        mp_sph_grad_phi_lm(s, X, l, m, inner, sg);
        t2[0] += sg[0] * tmp;
        t2[1] += sg[1] * tmp;
        if (sintheta != 0) t2[2] += sg[2] * (T * cm - S * sm) / (s * sintheta); else t2[2] = 0;
    }
    REAL gx = sintheta * cosphi * t2[0] + X * cosphi * t2[1] - sinphi * t2[2];
    REAL gy = sintheta * sinphi * t2[0] + X * sinphi * t2[1] + cosphi * t2[2];
    REAL gz = X * t2[0] - sintheta * t2[1];
    REAL sc = (REAL)p[0] * p[4] / ((REAL)p[5] * p[5]);
    g[0] += gx * sc; g[1] += gy * sc; g[2] += gz * sc;
}

/* ---- remaining analytic builtins ------------------------------------------------- */
/* Stone :662-719  [G, m, r_c, r_h] */
static REAL stone_value(const double *p, const REAL *q) {
    REAL r = norm3(q), u_c = r / p[2], u_h = r / p[3];
    REAL fac = 2 * (REAL)p[0] * p[1] / PI_R / ((REAL)p[3] - p[2]);
    if (r == 0) return -fac * 0.5 * LOG((REAL)p[3] * p[3] / ((REAL)p[2] * p[2]));
    return -fac * (ATAN(u_h) / u_h - ATAN(u_c) / u_c + 0.5 * LOG((r * r + (REAL)p[3] * p[3]) / (r * r + (REAL)p[2] * p[2])));
}
static void stone_grad(const double *p, const REAL *q, REAL *g) {
    REAL r = norm3(q), u_c = r / p[2], u_h = r / p[3];
    REAL fac = 2 * (REAL)p[0] * p[1] / (PI_R * r * r * r) / ((REAL)p[2] - p[3]);
    REAL d = fac * (p[2] * ATAN(u_c) - p[3] * ATAN(u_h));
    g[0] += d * q[0]; g[1] += d * q[1]; g[2] += d * q[2];
}
static REAL stone_density(const double *p, const REAL *q) {
    REAL r = norm3(q);
    REAL rho = (REAL)p[1] * ((REAL)p[2] + p[3]) / (2 * PI_R * PI_R * p[2] * p[2] * p[3] * p[3]);
    REAL u_c = r / p[2], u_t = r / p[3];
    return rho / ((1 + u_c * u_c) * (1 + u_t * u_t));
}
/* Burkert :2238-2279  [G, rho, r0] */
static REAL burkert_value(const double *p, const REAL *q) {
    REAL x = norm3(q) / p[2];
    return -PI_R * p[0] * p[1] * p[2] * p[2] *
           (PI_R - 2 * (1 + 1 / x) * ATAN(x) + 2 * (1 + 1 / x) * LOG(1 + x) - (1 - 1 / x) * LOG(1 + x * x));
}
static void burkert_grad(const double *p, const REAL *q, REAL *g) {
    REAL r = norm3(q), x = r / p[2];
    REAL d = -PI_R * p[0] * p[1] * p[2] / (x * x) * (2 * ATAN(x) - 2 * LOG(1 + x) - LOG(1 + x * x));
    g[0] += d * q[0] / r; g[1] += d * q[1] / r; g[2] += d * q[2] / r;
}
static REAL burkert_density(const double *p, const REAL *q) {
    REAL x = norm3(q) / p[2];
    return p[1] / ((1 + x) * (1 + x * x));
}
/* Satoh :1147-1186  [G, m, a, b] */
static REAL satoh_S2(const double *p, const REAL *q) {
    return (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) + p[2] * (p[2] + 2 * SQRT(q[2] * q[2] + (REAL)p[3] * p[3]));
}
static REAL satoh_value(const double *p, const REAL *q) { return -(REAL)p[0] * p[1] / SQRT(satoh_S2(p, q)); }
static void satoh_grad(const double *p, const REAL *q, REAL *g) {
    REAL S2 = satoh_S2(p, q), dS = (REAL)p[0] * p[1] / S2, S = SQRT(S2);
    g[0] += dS * q[0] / S; g[1] += dS * q[1] / S;
    g[2] += dS / S * q[2] * (1 + p[2] / SQRT(q[2] * q[2] + (REAL)p[3] * p[3]));
}
static REAL satoh_density(const double *p, const REAL *q) {
    REAL z2b2 = q[2] * q[2] + (REAL)p[3] * p[3], xyz2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    REAL S2 = xyz2 + p[2] * (p[2] + 2 * SQRT(z2b2));
    REAL A = (REAL)p[1] * p[2] * p[3] * p[3] / (4 * PI_R * S2 * SQRT(S2) * z2b2);
    return A * (1 / SQRT(z2b2) + 3 / p[2] * (1 - xyz2 / S2));
}
/* Kuzmin :1235-1283  [G, m, a] */
static REAL kuzmin_value(const double *p, const REAL *q) {
    REAL az = p[2] + FABS(q[2]);
    return -(REAL)p[0] * p[1] / SQRT(q[0] * q[0] + q[1] * q[1] + az * az);
}
static void kuzmin_grad(const double *p, const REAL *q, REAL *g) {
    REAL az = p[2] + FABS(q[2]);
    REAL fac = (REAL)p[0] * p[1] * POW(q[0] * q[0] + q[1] * q[1] + az * az, -1.5);
    REAL zs = (q[2] > 0) ? 1 : ((q[2] < 0) ? -1 : 0);
    g[0] += fac * q[0]; g[1] += fac * q[1]; g[2] += fac * zs * az;
}
static REAL kuzmin_density(const double *p, const REAL *q) {
    if (q[2] != 0) return 0;
    return (REAL)p[1] * p[2] / (2 * PI_R) * POW(q[0] * q[0] + q[1] * q[1] + (REAL)p[2] * p[2], -1.5);
}
/* Logarithmic :1560-1625  [G, v_c, r_h, q1, q2, q3, phi] */
static REAL log_value(const double *p, const REAL *q) {
    REAL cp = COS(p[6]), sp = SIN(p[6]);
    REAL x = q[0] * cp + q[1] * sp, y = -q[0] * sp + q[1] * cp, z = q[2];
    return 0.5 * p[1] * p[1] * LOG((REAL)p[2] * p[2] + x * x / ((REAL)p[3] * p[3]) + y * y / ((REAL)p[4] * p[4]) + z * z / ((REAL)p[5] * p[5]));
}
static void log_grad(const double *p, const REAL *q, REAL *g) {
    REAL cp = COS(p[6]), sp = SIN(p[6]);
    REAL x = q[0] * cp + q[1] * sp, y = -q[0] * sp + q[1] * cp, z = q[2];
    REAL fac = (REAL)p[1] * p[1] / ((REAL)p[2] * p[2] + x * x / ((REAL)p[3] * p[3]) + y * y / ((REAL)p[4] * p[4]) + z * z / ((REAL)p[5] * p[5]));
    REAL ax = fac * x / ((REAL)p[3] * p[3]), ay = fac * y / ((REAL)p[4] * p[4]), az = fac * z / ((REAL)p[5] * p[5]);
    g[0] += ax * cp - ay * sp; g[1] += ax * sp + ay * cp; g[2] += az;
}
static REAL log_density(const double *p, const REAL *q) {   /* ignores phi, as coded :1579-1601 */
    REAL q1s = (REAL)p[3] * p[3], q2s = (REAL)p[4] * p[4], q3s = (REAL)p[5] * p[5];
    REAL t2 = q1s * q2s, t3 = t2 * (q[2] * q[2]), t5 = q1s * q3s, t6 = t5 * (q[1] * q[1]);
    REAL t7 = q2s * q3s, t8 = t7 * (q[0] * q[0]), t9 = (REAL)p[2] * p[2] * t2 * q3s;
    REAL t10 = t6 + t8 + t9, t11 = t3 + t9, den = t10 + t3;
    return (REAL)p[1] * p[1] * (t2 * (t10 - t3) + t5 * (t11 - t6 + t8) + t7 * (t11 + t6 - t8)) / (den * den) / (4 * PI_R * p[0]);
}
/* Lee & Suto :1446-1558  [G, v_c, r_s, a, b, c] */
static REAL ls_vh2(const double *p, REAL *eb, REAL *ec) {
    REAL ba = (REAL)p[4] / p[3], ca = (REAL)p[5] / p[3];
    *eb = 1 - ba * ba; *ec = 1 - ca * ca;
    REAL ln2 = LOG((REAL)2);
    return (REAL)p[1] * p[1] / (ln2 - 0.5 + (ln2 - 0.75) * *eb + (ln2 - 0.75) * *ec);
}
static REAL ls_value(const double *p, const REAL *q) {
    REAL eb, ec, phi0 = ls_vh2(p, &eb, &ec);
    REAL x = q[0], y = q[1], z = q[2], r = SQRT(x * x + y * y + z * z), u = r / p[2];
    if (u == 0) return phi0;
    REAL l1u = LOG(1 + u);
    REAL F1 = -l1u / u;
    REAL F2 = -1 / (REAL)3 + (2 * u * u - 3 * u + 6) / (6 * u * u) + (1 / u - POW(u, -3)) * l1u;
    REAL F3 = (u * u - 3 * u - 6) / (2 * u * u * (1 + u)) + 3 * POW(u, -3) * l1u;
    REAL c2 = z * z / (r * r), s2 = 1 - c2, sp2 = y * y / (x * x + y * y);
    return phi0 * (F1 + (eb + ec) / 2 * F2 + (eb * s2 * sp2 + ec * c2) / 2 * F3);
}
static void ls_grad(const double *p, const REAL *q, REAL *g) {
    REAL eb, ec, vh2 = ls_vh2(p, &eb, &ec), rs = p[2];
    REAL x = q[0], y = q[1], z = q[2];
    REAL r2 = x * x + y * y + z * z, r = SQRT(r2), r4 = r2 * r2;
    REAL x0 = r + rs, x1 = x0 * x0, x2 = vh2 / (12 * r4 * r2 * r * x1), x10 = LOG(x0 / rs);
    REAL x13 = r * 3 * rs, x15 = x13 - r2, x16 = x15 + 6 * (rs * rs);
    REAL x17 = 6 * rs * x0 * (r * x16 - x0 * x10 * 6 * (rs * rs));
    REAL x20 = x0 * r2, x21 = 2 * r * x0, x7 = eb * y * y + ec * z * z;
    REAL x22 = -12 * r4 * r * rs * x0 + 12 * r4 * rs * x1 * x10
             + 3 * rs * x7 * (x16 * r2 - 18 * x1 * x10 * (rs * rs) + x20 * (2 * r - 3 * rs) + x21 * (x15 + 9 * (rs * rs)))
             - x20 * (eb + ec) * (-6 * r * rs * (r2 - (rs * rs)) + 6 * rs * x0 * x10 * (r2 - 3 * (rs * rs))
                                  + x20 * (-4 * r + 3 * rs) + x21 * (-x13 + 2 * r2 + 6 * (rs * rs)));
    g[0] += x2 * x * (x17 * x7 + x22);
    g[1] += x2 * y * (x17 * (x7 - r2 * eb) + x22);
    g[2] += x2 * z * (x17 * (x7 - r2 * ec) + x22);
}
static REAL ls_density(const double *p, const REAL *q) {
    REAL ba2 = (REAL)p[4] * p[4] / ((REAL)p[3] * p[3]), ca2 = (REAL)p[5] * p[5] / ((REAL)p[3] * p[3]);
    REAL eb, ec, vh2 = ls_vh2(p, &eb, &ec);
    REAL u = SQRT(q[0] * q[0] + q[1] * q[1] / ba2 + q[2] * q[2] / ca2) / p[2];
    return vh2 / (u * (1 + u) * (1 + u)) / (4 * PI_R * p[2] * p[2] * p[0]);
}
/* PowerLawCutoff :465-554  [G, m, alpha, r_c].  gsl_sf_gamma_inc_P (GSL, absent) restated from the
 * published series / continued fraction of the regularised incomplete gamma function. */
static REAL gamma_inc_P(REAL a, REAL x) {
    if (!(x > 0)) return 0;
    REAL lead = EXP(a * LOG(x) - x - LGAMMA(a));
    if (x < a + 1) {
        REAL ap = a, del = 1 / a, sum = del;
        for (int n = 0; n < 500; n++) { ap += 1; del *= x / ap; sum += del; if (FABS(del) < FABS(sum) * 1e-19) break; }
        return sum * lead;
    }
    REAL tiny = 1e-300, b = x + 1 - a, c = 1 / tiny, d = 1 / b, h = d;
    for (int i = 1; i < 500; i++) {
        REAL an = -(REAL)i * ((REAL)i - a);
        b += 2;
        d = an * d + b; if (FABS(d) < tiny) d = tiny;
        c = b + an / c; if (FABS(c) < tiny) c = tiny;
        d = 1 / d;
        REAL del = d * c; h *= del;
        if (FABS(del - 1) < 1e-19) break;
    }
    return 1 - lead * h;
}
static REAL safe_gamma_inc_r(REAL a, REAL x) {
    if (a > 0) return gamma_inc_P(a, x) * TGAMMA(a);
    int N = (int)CEIL(-a);
    REAL A = 1, B = 0;
    for (int n = 0; n < N; n++) {
        A = A * (a + n);
        REAL tmp = 1;
        for (int m = N - 1; m > n; m--) tmp = tmp * (a + m);
        B = B + POW(x, a + n) * EXP(-x) * tmp;
    }
    return (B + gamma_inc_P(a + N, x) * TGAMMA(a + N)) / A;
}
static REAL plc_value(const double *p, const REAL *q) {
    REAL r = norm3(q);
    if (r == 0) return -INFINITY;
    REAL t0 = (REAL)p[2] / 2, t1 = -t0, t2 = t1 + 1.5, t3 = r * r, t4 = t3 / ((REAL)p[3] * p[3]), t5 = (REAL)p[0] * p[1];
    REAL t6 = t5 * safe_gamma_inc_r(t2, t4) / (SQRT(t3) * TGAMMA(t1 + 2.5));
    REAL phi_r = t0 * t6 - 3.0 / 2.0 * t6 + t5 * safe_gamma_inc_r(t1 + 1, t4) / (p[3] * TGAMMA(t2));
    REAL phi_inf = 0;
    if (t2 > 0) phi_inf = t5 * TGAMMA(t1 + 1) / (p[3] * TGAMMA(t2));
    return phi_r - phi_inf;
}
static void plc_grad(const double *p, const REAL *q, REAL *g) {
    REAL r = norm3(q);
    REAL d = (REAL)p[0] * p[1] / (r * r * r) * gamma_inc_P(0.5 * (3 - (REAL)p[2]), r * r / ((REAL)p[3] * p[3]));
    g[0] += d * q[0]; g[1] += d * q[1]; g[2] += d * q[2];
}
static REAL plc_density(const double *p, const REAL *q) {
    REAL r = norm3(q);
    REAL A = p[1] / (2 * PI_R) * POW(p[3], (REAL)p[2] - 3) / TGAMMA(0.5 * (3 - (REAL)p[2]));
    return A * POW(r, -(REAL)p[2]) * EXP(-r * r / ((REAL)p[3] * p[3]));
}

/* ---- per-component dispatch ------------------------------------------------------------------- */
static void comp_grad(int type, const double *p, const REAL *q, REAL *g) {
    switch (type) {
        case GB_POT_HERNQUIST: hernquist_grad(p, q, g); break;
        case GB_POT_NFW_SPHERICAL: case GB_POT_NFW_FLATTENED: case GB_POT_NFW_TRIAXIAL: nfw_grad(type, p, q, g); break;
        case GB_POT_MIYAMOTONAGAI: mn_grad(p[0], p[1], p[2], p[3], q, g); break;
        case GB_POT_MN3: for (int i = 0; i < 3; i++) mn_grad(p[0], p[1 + 3 * i], p[2 + 3 * i], p[3 + 3 * i], q, g); break; /* :1404-1414 */
        case GB_POT_LONGMURALIBAR: lmbar_grad(p, q, g); break;
        case GB_POT_SCF: scf_grad(p, q, g); break;
        case GB_POT_MULTIPOLE: mp_grad(p, q, g); break;
        case GB_POT_KEPLER: kepler_grad(p, q, g); break;
        case GB_POT_PLUMMER: plummer_grad(p, q, g); break;
        case GB_POT_ISOCHRONE: isochrone_grad(p, q, g); break;
        case GB_POT_JAFFE: jaffe_grad(p, q, g); break;
        case GB_POT_STONE: stone_grad(p, q, g); break;
        case GB_POT_BURKERT: burkert_grad(p, q, g); break;
        case GB_POT_SATOH: satoh_grad(p, q, g); break;
        case GB_POT_KUZMIN: kuzmin_grad(p, q, g); break;
        case GB_POT_LOGARITHMIC: log_grad(p, q, g); break;
        case GB_POT_LEESUTO: ls_grad(p, q, g); break;
        case GB_POT_POWERLAWCUTOFF: plc_grad(p, q, g); break;
        default: break;
    }
}
static REAL comp_value(int type, const double *p, const REAL *q) {
    switch (type) {
        case GB_POT_HERNQUIST: return hernquist_value(p, q);
        case GB_POT_NFW_SPHERICAL: case GB_POT_NFW_FLATTENED: case GB_POT_NFW_TRIAXIAL: return nfw_value(type, p, q);
        case GB_POT_MIYAMOTONAGAI: return mn_value(p[0], p[1], p[2], p[3], q);
        case GB_POT_MN3: { REAL v = 0; for (int i = 0; i < 3; i++) v += mn_value(p[0], p[1 + 3 * i], p[2 + 3 * i], p[3 + 3 * i], q); return v; }
        case GB_POT_LONGMURALIBAR: return lmbar_value(p, q);
        case GB_POT_SCF: return scf_value(p, q);
        case GB_POT_MULTIPOLE: return mp_value(p, q);
        case GB_POT_KEPLER: return kepler_value(p, q);
        case GB_POT_PLUMMER: return plummer_value(p, q);
        case GB_POT_ISOCHRONE: return isochrone_value(p, q);
        case GB_POT_JAFFE: return jaffe_value(p, q);
        case GB_POT_STONE: return stone_value(p, q);
        case GB_POT_BURKERT: return burkert_value(p, q);
        case GB_POT_SATOH: return satoh_value(p, q);
        case GB_POT_KUZMIN: return kuzmin_value(p, q);
        case GB_POT_LOGARITHMIC: return log_value(p, q);
        case GB_POT_LEESUTO: return ls_value(p, q);
        case GB_POT_POWERLAWCUTOFF: return plc_value(p, q);
        default: return 0;
    }
}
static REAL comp_density(int type, const double *p, const REAL *q) {
    switch (type) {
        case GB_POT_HERNQUIST: return hernquist_density(p, q);
        case GB_POT_NFW_SPHERICAL: case GB_POT_NFW_FLATTENED: case GB_POT_NFW_TRIAXIAL: return nfw_density(type, p, q);
        case GB_POT_MIYAMOTONAGAI: return mn_density(p[1], p[2], p[3], q);
        case GB_POT_MN3: { REAL v = 0; for (int i = 0; i < 3; i++) v += mn_density(p[1 + 3 * i], p[2 + 3 * i], p[3 + 3 * i], q); return v; }
        case GB_POT_LONGMURALIBAR: return lmbar_density(p, q);
        case GB_POT_SCF: return scf_density(p, q);
        case GB_POT_MULTIPOLE: return 0;      /* mp_density :404-420 returns 0 */
        case GB_POT_KEPLER: return kepler_density(p, q);
        case GB_POT_PLUMMER: return plummer_density(p, q);
        case GB_POT_ISOCHRONE: return isochrone_density(p, q);
        case GB_POT_JAFFE: return jaffe_density(p, q);
        case GB_POT_STONE: return stone_density(p, q);
        case GB_POT_BURKERT: return burkert_density(p, q);
        case GB_POT_SATOH: return satoh_density(p, q);
        case GB_POT_KUZMIN: return kuzmin_density(p, q);
        case GB_POT_LOGARITHMIC: return log_density(p, q);
        case GB_POT_LEESUTO: return ls_density(p, q);
        case GB_POT_POWERLAWCUTOFF: return plc_density(p, q);
        default: return 0;
    }
}

/* shift + rotate into the component frame (cpotential.cpp:134-167) */
static void to_comp(const gb_component *c, const REAL *q, REAL *o) {
    REAL s[3] = {q[0] - c->q0[0], q[1] - c->q0[1], q[2] - c->q0[2]};
    for (int k = 0; k < 3; k++) o[k] = c->R[3 * k] * s[0] + c->R[3 * k + 1] * s[1] + c->R[3 * k + 2] * s[2];
}
/* c_gradient for one point (cpotential.cpp:214-287) */
static void pot_gradient(const gb_potential *P, REAL t, const REAL *q, REAL *g) {
    g[0] = g[1] = g[2] = 0;
    for (int i = 0; i < P->n_components; i++) {
        const gb_component *c = &P->comp[i];
        if (!c->do_shift_rotate) { comp_grad(c->type_id, c->params, q, g); continue; }
        REAL qq[3], tg[3] = {0, 0, 0};
        to_comp(c, q, qq);
        comp_grad(c->type_id, c->params, qq, tg);
        for (int k = 0; k < 3; k++) g[k] += c->R[k] * tg[0] + c->R[3 + k] * tg[1] + c->R[6 + k] * tg[2];
    }
}
/* c_potential / c_density (cpotential.cpp:170-211) */
static REAL pot_value(const gb_potential *P, REAL t, const REAL *q) {
    REAL v = 0;
    for (int i = 0; i < P->n_components; i++) {
        const gb_component *c = &P->comp[i];
        REAL qq[3] = {q[0], q[1], q[2]};
        if (c->do_shift_rotate) to_comp(c, q, qq);
        v = v + comp_value(c->type_id, c->params, qq);
    }
    return v;
}
static REAL pot_density(const gb_potential *P, REAL t, const REAL *q) {
    REAL v = 0;
    for (int i = 0; i < P->n_components; i++) {
        const gb_component *c = &P->comp[i];
        REAL qq[3] = {q[0], q[1], q[2]};
        if (c->do_shift_rotate) to_comp(c, q, qq);
        v = v + comp_density(c->type_id, c->params, qq);
    }
    return v;
}
/* frame energy (frame/builtin/builtin_frames.cpp:8-16, 73-92) */
static REAL frame_energy(const gb_frame *fr, const REAL *w) {
    if (!fr || fr->type_id == GB_FRAME_STATIC) return 0.5 * (w[3] * w[3] + w[4] * w[4] + w[5] * w[5]);
    REAL E = 0.5 * w[3] * w[3] + 0.5 * w[4] * w[4] + 0.5 * w[5] * w[5];
    REAL Lx = w[1] * w[5] - w[2] * w[4], Ly = -w[0] * w[5] + w[2] * w[3], Lz = w[0] * w[4] - w[1] * w[3];
    return E - (fr->omega[0] * Lx + fr->omega[1] * Ly + fr->omega[2] * Lz);
}
/* hamiltonian_gradient (hamiltonian/src/chamiltonian.cpp:21-57; frame terms builtin_frames.cpp:18-24,94-112) */
static void ham_rhs(const gb_potential *P, const gb_frame *fr, REAL t, const REAL *w, REAL *f) {
    REAL g[3];
    pot_gradient(P, t, w, g);
    if (!fr || fr->type_id == GB_FRAME_STATIC) {
        f[0] = w[3]; f[1] = w[4]; f[2] = w[5];
        f[3] = -g[0]; f[4] = -g[1]; f[5] = -g[2];
        return;
    }
    const double *om = fr->omega;
    REAL Cx = om[1] * w[2] - om[2] * w[1], Cy = -om[0] * w[2] + om[2] * w[0], Cz = om[0] * w[1] - om[1] * w[0];
    f[0] = w[3] - Cx; f[1] = w[4] - Cy; f[2] = w[5] - Cz;
    Cx = om[1] * w[5] - om[2] * w[4]; Cy = -om[0] * w[5] + om[2] * w[3]; Cz = om[0] * w[4] - om[1] * w[3];
    f[3] = -(g[0] + Cx); f[4] = -(g[1] + Cy); f[5] = -(g[2] + Cz);
}

/* ------------------------------------------------------------------------------------------------
 * exported evaluation functions (q is (3,N), w is (6,N))
 * ---------------------------------------------------------------------------------------------- */
int port_gradient(const gb_potential *P, const double *q, double t, size_t N, double *grad) {
    for (size_t i = 0; i < N; i++) {
        REAL qq[3] = {q[i], q[N + i], q[2 * N + i]}, g[3];
        pot_gradient(P, t, qq, g);
        grad[i] = (double)g[0]; grad[N + i] = (double)g[1]; grad[2 * N + i] = (double)g[2];
    }
    return 0;
}
int port_energy(const gb_potential *P, const double *q, double t, size_t N, double *out) {
    for (size_t i = 0; i < N; i++) { REAL qq[3] = {q[i], q[N + i], q[2 * N + i]}; out[i] = (double)pot_value(P, t, qq); }
    return 0;
}
int port_density(const gb_potential *P, const double *q, double t, size_t N, double *out) {
    for (size_t i = 0; i < N; i++) { REAL qq[3] = {q[i], q[N + i], q[2 * N + i]}; out[i] = (double)pot_density(P, t, qq); }
    return 0;
}
int port_hamiltonian_energy(const gb_potential *P, const gb_frame *fr, const double *w, double t, size_t N, double *out) {
    for (size_t i = 0; i < N; i++) {
        REAL ww[6]; for (int k = 0; k < 6; k++) ww[k] = w[k * N + i];
        out[i] = (double)(pot_value(P, t, ww) + frame_energy(fr, ww));
    }
    return 0;
}
int port_hamiltonian_gradient(const gb_potential *P, const gb_frame *fr, const double *w, double t, size_t N, double *f) {
    for (size_t i = 0; i < N; i++) {
        REAL ww[6], ff[6]; for (int k = 0; k < 6; k++) ww[k] = w[k * N + i];
        ham_rhs(P, fr, t, ww, ff);
        for (int k = 0; k < 6; k++) f[k * N + i] = (double)ff[k];
    }
    return 0;
}
/* c_d2_dr2 (cpotential.cpp:346-371) */
double port_d2_dr2(const gb_potential *P, double t, const double *q3) {
    REAL q[3] = {q3[0], q3[1], q3[2]}, e[3], h = 1E-2, r2 = 0, d;
    for (int j = 0; j < 3; j++) r2 = r2 + q[j] * q[j];
    REAL r = SQRT(r2);
    for (int j = 0; j < 3; j++) e[j] = q[j] + h * q[j] / r;
    d = pot_value(P, t, e);
    d = d - 2. * pot_value(P, t, q);
    for (int j = 0; j < 3; j++) e[j] = q[j] - h * q[j] / r;
    d = d + pot_value(P, t, e);
    return (double)(d / (h * h));
}

/* ------------------------------------------------------------------------------------------------
 * Leapfrog (integrate/cyintegrators/leapfrog.pyx:24-51, 54-121), orbit by orbit (orbits are
 * independent; the reference's loop order -- all orbits per step -- gives the same numbers).
 * ---------------------------------------------------------------------------------------------- */
int port_leapfrog(const gb_potential *P, const double *w0, size_t N, const double *t, int ntimes, int save_all, double *out) {
    REAL dt = (REAL)t[1] - (REAL)t[0];
    size_t TS = (size_t)ntimes * N;
    for (size_t i = 0; i < N; i++) {
        REAL x[3], v[3], h[3], g[3];
        for (int k = 0; k < 3; k++) { x[k] = w0[k * N + i]; v[k] = w0[(3 + k) * N + i]; }
        if (save_all) for (int k = 0; k < 3; k++) { out[k * TS + i] = (double)x[k]; out[(3 + k) * TS + i] = (double)v[k]; }
        pot_gradient(P, t[0], x, g);
        for (int k = 0; k < 3; k++) h[k] = v[k] - g[k] * dt / 2.;
        for (int j = 1; j < ntimes; j++) {
            for (int k = 0; k < 3; k++) x[k] = x[k] + h[k] * dt;
            pot_gradient(P, t[j], x, g);
            for (int k = 0; k < 3; k++) { v[k] = h[k] - g[k] * dt / 2.; h[k] = h[k] - g[k] * dt; }
            if (save_all) for (int k = 0; k < 3; k++) { out[k * TS + (size_t)j * N + i] = (double)x[k]; out[(3 + k) * TS + (size_t)j * N + i] = (double)v[k]; }
        }
        if (!save_all) for (int k = 0; k < 3; k++) { out[k * N + i] = (double)x[k]; out[(3 + k) * N + i] = (double)v[k]; }
    }
    return 0;
}

/* Ruth4 (integrate/cyintegrators/ruth4.pyx:24-35, 65-78; rotating frame: the Python integrator,
 * integrate/pyintegrators/ruth4.py:106-124 with F = Hamiltonian._gradient, chamiltonian.pyx:88-99) */
int port_ruth4(const gb_potential *P, const gb_frame *fr, const double *w0, size_t N, const double *t, int ntimes, int save_all, double *out) {
    const double two_13 = pow(2., 1. / 3.);     /* coefficients are doubles in the reference */
    const double cs[4] = {1. / (2. * (2. - two_13)), (1. - two_13) / (2. * (2. - two_13)), (1. - two_13) / (2. * (2. - two_13)), 1. / (2. * (2. - two_13))};
    const double ds[4] = {0., 1. / (2. - two_13), -two_13 / (2. - two_13), 1. / (2. - two_13)};
    const int rot = fr && fr->type_id == GB_FRAME_ROTATING_3D;
    REAL dt = (REAL)t[1] - (REAL)t[0];
    size_t TS = (size_t)ntimes * N;
    for (size_t i = 0; i < N; i++) {
        REAL w[6], g[3], f[6];
        for (int k = 0; k < 6; k++) w[k] = w0[k * N + i];
        if (save_all) for (int k = 0; k < 6; k++) out[k * TS + i] = (double)w[k];
        for (int j = 1; j < ntimes; j++) {
            for (int s = 0; s < 4; s++) {
                if (!rot) {
                    pot_gradient(P, t[j], w, g);
                    for (int k = 0; k < 3; k++) { w[3 + k] = w[3 + k] - ds[s] * g[k] * dt; w[k] = w[k] + cs[s] * w[3 + k] * dt; }
                } else {
                    ham_rhs(P, fr, t[j], w, f);
                    for (int k = 0; k < 3; k++) { w[3 + k] = w[3 + k] + ds[s] * f[3 + k] * dt; w[k] = w[k] + cs[s] * w[3 + k] * dt; }
                }
            }
            if (save_all) for (int k = 0; k < 6; k++) out[k * TS + (size_t)j * N + i] = (double)w[k];
        }
        if (!save_all) for (int k = 0; k < 6; k++) out[k * N + i] = (double)w[k];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * DOP853 for one orbit (n = 6): dopcor + hinit + dense output
 * (integrate/cyintegrators/dopri/dop853.cpp:18-650, 869-904; defaults :673-788).
 * Coefficients: Hairer, Norsett & Wanner, DOP853 (published; same values as dop853.cpp:128-294).
 * ---------------------------------------------------------------------------------------------- */
#include "port_dop853_coeffs.h"

typedef struct { double atol, rtol; long nmax, nstiff; double hmax, uround, h0; } dop_args;

static REAL sgn(REAL a, REAL b) { return (b < 0) ? -FABS(a) : FABS(a); }
static REAL mn_(REAL a, REAL b) { return a < b ? a : b; }
static REAL mx_(REAL a, REAL b) { return a > b ? a : b; }
#define N6 6
#define RHS(tt, ww, ff) ham_rhs(P, fr, (tt), (ww), (ff))

static int dop853_one(const gb_potential *P, const gb_frame *fr, const dop_args *a, REAL x, REAL xend, REAL *y, REAL h,
                      const double *tout, int ntout, double *out, size_t out_stride_t, size_t out_stride_k,
                      int *nstep_, int *naccpt_, int *nrejct_, int *nfcn_) {
    REAL k1[N6], k2[N6], k3[N6], k4[N6], k5[N6], k6[N6], k7[N6], k8[N6], k9[N6], k10[N6], yy1[N6];
    REAL rc1[N6], rc2[N6], rc3[N6], rc4[N6], rc5[N6], rc6[N6], rc7[N6], rc8[N6];
    const int dense = out != NULL && ntout > 0;
    const REAL safe = 0.9, fac1 = 0.333, fac2 = 6.0, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    const REAL posneg = sgn(1.0, xend - x), atoli = a->atol, rtoli = a->rtol;
    REAL hmax = FABS(a->hmax == 0.0 ? (xend - x) : (REAL)a->hmax);
    REAL facold = 1.0E-4, hlamb = 0.0, hnew, err, err2, deno, fac, fac11;
    int last = 0, reject = 0, out_idx = 0, code = 0;
    int nstep = 0, naccpt = 0, nrejct = 0, nfcn = 0;
    RHS(x, y, k1);
    if (h == 0.0) {   /* hinit :18-86 */
        REAL dnf = 0, dny = 0, sk, sqr, der2 = 0, der12, h1;
        for (int i = 0; i < N6; i++) { sk = atoli + rtoli * FABS(y[i]); sqr = k1[i] / sk; dnf += sqr * sqr; sqr = y[i] / sk; dny += sqr * sqr; }
        h = (dnf <= 1.0E-10 || dny <= 1.0E-10) ? 1.0E-6 : SQRT(dny / dnf) * 0.01;
        h = mn_(h, hmax); h = sgn(h, posneg);
        for (int i = 0; i < N6; i++) k3[i] = y[i] + h * k1[i];
        RHS(x + h, k3, k2);
        for (int i = 0; i < N6; i++) { sk = atoli + rtoli * FABS(y[i]); sqr = (k2[i] - k1[i]) / sk; der2 += sqr * sqr; }
        der2 = SQRT(der2) / h;
        der12 = mx_(FABS(der2), SQRT(dnf));
        h1 = (der12 <= 1.0E-15) ? mx_(1.0E-6, FABS(h) * 1.0E-3) : POW(0.01 / der12, 1.0 / 8.0);
        h = mn_(100.0 * FABS(h), mn_(h1, hmax));
        h = sgn(h, posneg);
    }
    nfcn += 2;
    for (;;) {
        if (nstep > a->nmax) { code = -2; break; }
        if (0.1 * FABS(h) <= FABS(x) * a->uround) { code = -3; break; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = 1; }
        nstep++;
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * a21 * k1[i];
        RHS(x + c2 * h, yy1, k2);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a31 * k1[i] + a32 * k2[i]);
        RHS(x + c3 * h, yy1, k3);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a41 * k1[i] + a43 * k3[i]);
        RHS(x + c4 * h, yy1, k4);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a51 * k1[i] + a53 * k3[i] + a54 * k4[i]);
        RHS(x + c5 * h, yy1, k5);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a61 * k1[i] + a64 * k4[i] + a65 * k5[i]);
        RHS(x + c6 * h, yy1, k6);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a71 * k1[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
        RHS(x + c7 * h, yy1, k7);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a81 * k1[i] + a84 * k4[i] + a85 * k5[i] + a86 * k6[i] + a87 * k7[i]);
        RHS(x + c8 * h, yy1, k8);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a91 * k1[i] + a94 * k4[i] + a95 * k5[i] + a96 * k6[i] + a97 * k7[i] + a98 * k8[i]);
        RHS(x + c9 * h, yy1, k9);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a101 * k1[i] + a104 * k4[i] + a105 * k5[i] + a106 * k6[i] + a107 * k7[i] + a108 * k8[i] + a109 * k9[i]);
        RHS(x + c10 * h, yy1, k10);
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a111 * k1[i] + a114 * k4[i] + a115 * k5[i] + a116 * k6[i] + a117 * k7[i] + a118 * k8[i] + a119 * k9[i] + a1110 * k10[i]);
        RHS(x + c11 * h, yy1, k2);
        REAL xph = x + h;
        for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a121 * k1[i] + a124 * k4[i] + a125 * k5[i] + a126 * k6[i] + a127 * k7[i] + a128 * k8[i] + a129 * k9[i] + a1210 * k10[i] + a1211 * k2[i]);
        RHS(xph, yy1, k3);
        nfcn += 11;
        for (int i = 0; i < N6; i++) {
            k4[i] = b1 * k1[i] + b6 * k6[i] + b7 * k7[i] + b8 * k8[i] + b9 * k9[i] + b10 * k10[i] + b11 * k2[i] + b12 * k3[i];
            k5[i] = y[i] + h * k4[i];
        }
        err = 0; err2 = 0;
        for (int i = 0; i < N6; i++) {
            REAL sk = atoli + rtoli * mx_(FABS(y[i]), FABS(k5[i]));
            REAL erri = k4[i] - bhh1 * k1[i] - bhh2 * k9[i] - bhh3 * k3[i];
            REAL sqr = erri / sk; err2 += sqr * sqr;
            erri = er1 * k1[i] + er6 * k6[i] + er7 * k7[i] + er8 * k8[i] + er9 * k9[i] + er10 * k10[i] + er11 * k2[i] + er12 * k3[i];
            sqr = erri / sk; err += sqr * sqr;
        }
        deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = FABS(h) * err * SQRT(1.0 / (deno * (REAL)N6));
        fac11 = POW(err, (REAL)(1.0 / 8.0));
        fac = fac11;                                   /* beta = 0: pow(facold, 0) = 1 */
        fac = mx_(facc2, mn_(facc1, fac / safe));
        hnew = h / fac;
        if (err <= 1.0) {
            facold = mx_(err, 1.0E-4);
            naccpt++;
            RHS(xph, k5, k4);
            nfcn++;
            if (!(naccpt % a->nstiff)) {               /* :460-485 as coded: a hit returns -4 at once */
                REAL stnum = 0, stden = 0, sqr;
                for (int i = 0; i < N6; i++) { sqr = k4[i] - k3[i]; stnum += sqr * sqr; sqr = k5[i] - yy1[i]; stden += sqr * sqr; }
                if (stden > 0.0) hlamb = h * SQRT(stnum / stden);
                if (hlamb > 6.1) { code = -4; break; }
            }
            if (dense) {
                for (int i = 0; i < N6; i++) {
                    rc1[i] = y[i];
                    REAL ydiff = k5[i] - y[i]; rc2[i] = ydiff;
                    REAL bspl = h * k1[i] - ydiff; rc3[i] = bspl;
                    rc4[i] = ydiff - h * k4[i] - bspl;
                    rc5[i] = d41 * k1[i] + d46 * k6[i] + d47 * k7[i] + d48 * k8[i] + d49 * k9[i] + d410 * k10[i] + d411 * k2[i] + d412 * k3[i];
                    rc6[i] = d51 * k1[i] + d56 * k6[i] + d57 * k7[i] + d58 * k8[i] + d59 * k9[i] + d510 * k10[i] + d511 * k2[i] + d512 * k3[i];
                    rc7[i] = d61 * k1[i] + d66 * k6[i] + d67 * k7[i] + d68 * k8[i] + d69 * k9[i] + d610 * k10[i] + d611 * k2[i] + d612 * k3[i];
                    rc8[i] = d71 * k1[i] + d76 * k6[i] + d77 * k7[i] + d78 * k8[i] + d79 * k9[i] + d710 * k10[i] + d711 * k2[i] + d712 * k3[i];
                }
                for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a141 * k1[i] + a147 * k7[i] + a148 * k8[i] + a149 * k9[i] + a1410 * k10[i] + a1411 * k2[i] + a1412 * k3[i] + a1413 * k4[i]);
                RHS(x + c14 * h, yy1, k10);
                for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a151 * k1[i] + a156 * k6[i] + a157 * k7[i] + a158 * k8[i] + a1511 * k2[i] + a1512 * k3[i] + a1513 * k4[i] + a1514 * k10[i]);
                RHS(x + c15 * h, yy1, k2);
                for (int i = 0; i < N6; i++) yy1[i] = y[i] + h * (a161 * k1[i] + a166 * k6[i] + a167 * k7[i] + a168 * k8[i] + a169 * k9[i] + a1613 * k4[i] + a1614 * k10[i] + a1615 * k2[i]);
                RHS(x + c16 * h, yy1, k3);
                nfcn += 3;
                for (int i = 0; i < N6; i++) {
                    rc5[i] = h * (rc5[i] + d413 * k4[i] + d414 * k10[i] + d415 * k2[i] + d416 * k3[i]);
                    rc6[i] = h * (rc6[i] + d513 * k4[i] + d514 * k10[i] + d515 * k2[i] + d516 * k3[i]);
                    rc7[i] = h * (rc7[i] + d613 * k4[i] + d614 * k10[i] + d615 * k2[i] + d616 * k3[i]);
                    rc8[i] = h * (rc8[i] + d713 * k4[i] + d714 * k10[i] + d715 * k2[i] + d716 * k3[i]);
                }
                REAL x0 = x, x1 = x0 + h;
                while (out_idx < ntout) {
                    REAL to = tout[out_idx];
                    if ((x0 <= to && to <= x1) || (x1 <= to && to <= x0)) {
                        REAL s = (to - x0) / h, s1 = 1.0 - s;
                        for (int i = 0; i < N6; i++)
                            out[(size_t)out_idx * out_stride_t + i * out_stride_k] = (double)(rc1[i] + s * (rc2[i] + s1 * (rc3[i] + s * (rc4[i] + s1 * (rc5[i] + s * (rc6[i] + s1 * (rc7[i] + s * rc8[i])))))));
                        out_idx++;
                    } else break;
                }
            }
            for (int i = 0; i < N6; i++) { k1[i] = k4[i]; y[i] = k5[i]; }
            x = xph;
            if (last) { code = 1; break; }
            if (FABS(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * mn_(FABS(hnew), FABS(h));
            reject = 0;
        } else {
            hnew = h / mn_(facc1, fac11 / safe);
            reject = 1;
            if (naccpt >= 1) nrejct = nrejct + 1;
            last = 0;
        }
        h = hnew;
    }
    if (dense) for (; out_idx < ntout; out_idx++) for (int i = 0; i < N6; i++) out[(size_t)out_idx * out_stride_t + i * out_stride_k] = NAN;
    if (nstep_) *nstep_ = nstep; if (naccpt_) *naccpt_ = naccpt; if (nrejct_) *nrejct_ = nrejct; if (nfcn_) *nfcn_ = nfcn;
    return code;
}

static int dop_defaults(dop_args *a, double atol, double rtol, long nmax, double dt_max, long nstiff, double uround, double h0) {
    a->atol = atol; a->rtol = rtol;
    if (!nmax) nmax = 1000000; else if (nmax < 0) return -1;       /* dop853.cpp:673-680 */
    a->nmax = nmax;
    if (!nstiff) nstiff = 1000; else if (nstiff < 0) nstiff = nmax + 10;   /* :691-694 */
    a->nstiff = nstiff;
    a->uround = (uround == 0.0) ? 2.3E-16 : uround;                /* :751-752 */
    a->hmax = dt_max; a->h0 = h0;
    return 0;
}

/* dop853_integrate_hamiltonian with nbatch = 1 (integrate/cyintegrators/dop853.pyx:196-250;
 * dop853_helper :90-193 passes uround = eps, h = t[1]-t[0]).  `nbatch` is accepted for interface
 * compatibility with the reference driver and must be 1. */
int port_dop853(const gb_potential *P, const gb_frame *fr, const double *w0, size_t N, const double *t, int ntimes,
                double atol, double rtol, long nmax, double dt_max, long nstiff, int save_all, int nbatch,
                double *w_out, int32_t *status) {
    dop_args a;
    if (nbatch != 1) return -12;
    if (dop_defaults(&a, atol, rtol, nmax, dt_max, nstiff, 2.220446049250313e-16, t[1] - t[0])) return -1;
    int worst = 0;
    size_t TS = (size_t)ntimes * N;
    for (size_t i = 0; i < N; i++) {
        REAL y[6]; for (int k = 0; k < 6; k++) y[k] = w0[k * N + i];
        int code = dop853_one(P, fr, &a, t[0], t[ntimes - 1], y, a.h0, t, save_all ? ntimes : 0,
                              save_all ? w_out + i : NULL, N, TS, NULL, NULL, NULL, NULL);
        if (!save_all) for (int k = 0; k < 6; k++) w_out[k * N + i] = (double)y[k];
        if (status) status[i] = code;
        if (code < worst) worst = code;
    }
    return worst;
}

/* dop853_step settings (dop853.pyx:27-75): uround default, nstiff = 1, no dense output; rows (Np,6).
 * group must be 1 (each particle its own n = 6 system). */
int port_dop853_step_rows(const gb_potential *P, const gb_frame *fr, double *rows, size_t Np, double t1, double t2,
                          double dt0, double atol, double rtol, long nmax, int group, int32_t *status) {
    dop_args a;
    if (!group) return -12;
    if (dop_defaults(&a, atol, rtol, nmax, 0.0, 1, 0.0, dt0)) return -1;
    int worst = 0;
    for (size_t i = 0; i < Np; i++) {
        REAL y[6]; for (int k = 0; k < 6; k++) y[k] = rows[6 * i + k];
        int code = dop853_one(P, fr, &a, t1, t2, y, dt0, NULL, 0, NULL, 0, 0, NULL, NULL, NULL, NULL);
        for (int k = 0; k < 6; k++) rows[6 * i + k] = (double)y[k];
        if (status) status[i] = code;
        if (code < worst) worst = code;
    }
    return worst;
}

const char *port_build_flags(void) {
#if PORT_LONG_DOUBLE
    return "gcc -std=c11 -O2 -ffp-contract=off, REAL = long double (oracle/port.c)";
#else
    return "gcc -std=c11 -O2 -ffp-contract=off, REAL = double (oracle/port.c)";
#endif
}
