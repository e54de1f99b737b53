/* oracle/gsl_shim: TEST INFRASTRUCTURE.  gsl_sf_gegenpoly_n(n, lambda, x) = C_n^{(lambda)}(x), by the
 * published three-term recurrence (Abramowitz & Stegun 22.7.3), which is also what GSL's
 * specfunc/gegenbauer.c does for n >= 4:
 *   C_0 = 1, C_1 = 2 lambda x,
 *   k C_k = 2 (k + lambda - 1) x C_{k-1} - (k + 2 lambda - 2) C_{k-2}.
 * Call sites: potential/scf/src/bfe_helper.cpp:17,25,56-57. */
#ifndef GB_SHIM_GSL_SF_GEGENBAUER_H
#define GB_SHIM_GSL_SF_GEGENBAUER_H
static inline double gsl_sf_gegenpoly_n(int n, double lambda, double x) {
    if (n < 0) return 0.;
    if (n == 0) return 1.;
    double gkm2 = 1., gkm1 = 2. * lambda * x;
    for (int k = 2; k <= n; k++) {
        double gk = (2. * (k + lambda - 1.) * x * gkm1 - (k + 2. * lambda - 2.) * gkm2) / k;
        gkm2 = gkm1; gkm1 = gk;
    }
    return gkm1;
}
#endif
