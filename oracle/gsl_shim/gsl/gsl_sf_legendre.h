/* oracle/gsl_shim: TEST INFRASTRUCTURE.  gsl_sf_legendre_Plm(l,m,x) = associated Legendre P_l^m(x)
 * INCLUDING the Condon-Shortley phase (-1)^m (GSL's and scipy.special.lpmv's convention; SURVEY.md
 * Appendix A shows the reference's Fortran golden vectors require it), and
 * gsl_sf_legendre_sphPlm(l,m,x) = sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_l^m(x).
 * Standard recurrences (Numerical Recipes 6.8 / GSL specfunc/legendre_poly.c):
 *   P_m^m = (-1)^m (2m-1)!! (1-x^2)^{m/2};  P_{m+1}^m = x (2m+1) P_m^m;
 *   (l-m) P_l^m = x (2l-1) P_{l-1}^m - (l+m-1) P_{l-2}^m.
 * Call sites: potential/scf/src/bfe_helper.cpp:21,28,42,46,66. */
#ifndef GB_SHIM_GSL_SF_LEGENDRE_H
#define GB_SHIM_GSL_SF_LEGENDRE_H
#include <math.h>
static inline double gsl_sf_legendre_Plm(int l, int m, double x) {
    if (m < 0 || m > l) return 0.;
    double pmm = 1.;
    if (m > 0) {
        double somx2 = sqrt((1. - x) * (1. + x));
        double fact = 1.;
        for (int i = 1; i <= m; i++) { pmm *= -fact * somx2; fact += 2.; }
    }
    if (l == m) return pmm;
    double pmmp1 = x * (2 * m + 1) * pmm;
    if (l == m + 1) return pmmp1;
    double pll = 0.;
    for (int ll = m + 2; ll <= l; ll++) {
        pll = (x * (2 * ll - 1) * pmmp1 - (ll + m - 1) * pmm) / (ll - m);
        pmm = pmmp1; pmmp1 = pll;
    }
    return pll;
}
static inline double gsl_sf_legendre_sphPlm(int l, int m, double x) {
    double ratio = 1.;                       /* (l-m)!/(l+m)! */
    for (int k = l - m + 1; k <= l + m; k++) ratio /= (double)k;
    return sqrt((2. * l + 1.) / (4. * M_PI) * ratio) * gsl_sf_legendre_Plm(l, m, x);
}
#endif
