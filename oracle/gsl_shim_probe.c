/* oracle/gsl_shim_probe.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 * Exposes the special functions of oracle/gsl_shim/ (the stand-ins for the GSL calls the reference's SCF /
 * multipole / PowerLawCutoff sources make; GSL is a system library of the reference, absent here) so that
 * tests/test_oracle_cpu.py can pin them against scipy.special over the range the hot path uses.
 * Call sites in the reference: potential/scf/src/bfe_helper.cpp:17-74, potential/potential/builtin/
 * multipole.cpp:40-111, builtin_potentials.cpp (powerlawcutoff). */
#include "gsl/gsl_sf_gegenbauer.h"
#include "gsl/gsl_sf_legendre.h"
#include "gsl/gsl_sf_gamma.h"

void shim_gegenpoly_n(const int *n, const double *lambda, const double *x, int count, double *out) {
    for (int i = 0; i < count; i++) out[i] = gsl_sf_gegenpoly_n(n[i], lambda[i], x[i]);
}
void shim_legendre_Plm(const int *l, const int *m, const double *x, int count, double *out) {
    for (int i = 0; i < count; i++) out[i] = gsl_sf_legendre_Plm(l[i], m[i], x[i]);
}
void shim_legendre_sphPlm(const int *l, const int *m, const double *x, int count, double *out) {
    for (int i = 0; i < count; i++) out[i] = gsl_sf_legendre_sphPlm(l[i], m[i], x[i]);
}
void shim_gamma(const double *x, int count, double *out) {
    for (int i = 0; i < count; i++) out[i] = gsl_sf_gamma(x[i]);
}

#include "gsl/gsl_spline.h"
/* y[i] = spline of kind (0 linear, 1 cspline, 2 akima, 3 steffen) through (xk, yk)[nk] evaluated at t[i]; returns 0 / -1 */
int shim_spline_eval(int kind, const double *xk, const double *yk, int nk, const double *t, int count, double *out) {
    gsl_spline *s = gsl_spline_alloc(&gb_shim_interp_types[kind], (size_t)nk);
    if (!s) return -1;
    if (gsl_spline_init(s, xk, yk, (size_t)nk) != GSL_SUCCESS) { gsl_spline_free(s); return -1; }
    gsl_interp_accel *a = gsl_interp_accel_alloc();
    for (int i = 0; i < count; i++) out[i] = gsl_spline_eval(s, t[i], a);
    gsl_interp_accel_free(a); gsl_spline_free(s);
    return 0;
}
