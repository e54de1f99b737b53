/* oracle/gsl_shim: TEST INFRASTRUCTURE.  GSL (GNU Scientific Library, system package, unpinned in the
 * reference: setup.py:89-103 requires >=1.16) is absent from this image.  gsl_sf_gamma is called by the
 * reference only with positive integer arguments l-m+1, l+m+1 (potential/scf/src/bfe_helper.cpp:72-74),
 * where Gamma(k) = (k-1)! exactly; tgamma() is used for any other argument. */
#ifndef GB_SHIM_GSL_SF_GAMMA_H
#define GB_SHIM_GSL_SF_GAMMA_H
#include <math.h>
static inline double gsl_sf_gamma(double x) {
    if (x >= 1. && x <= 171. && x == floor(x)) {
        double f = 1.;
        for (int k = 2; k < (int)x; k++) f *= (double)k;
        return f;
    }
    return tgamma(x);
}
/* gsl_sf_gamma_inc_P(a, x): regularised lower incomplete gamma function, called by
 * powerlawcutoff_* (potential/potential/builtin/builtin_potentials.cpp:467-548) with a > 0.  Restated
 * from the published power series (x < a+1) and Lentz continued fraction for Q = 1-P (otherwise),
 * iterated in long double; pinned against scipy.special.gammainc in tests/test_oracle_cpu.py. */
static inline double gsl_sf_gamma_inc_P(double a_, double x_) {
    long double a = a_, x = x_;
    if (!(x > 0)) return 0.;
    long double lead = expl(a * logl(x) - x - lgammal(a));
    if (x < a + 1) {
        long double ap = a, del = 1 / a, sum = del;
        for (int n = 0; n < 1000; n++) { ap += 1; del *= x / ap; sum += del; if (fabsl(del) < fabsl(sum) * 1e-20L) break; }
        return (double)(sum * lead);
    }
    long double tiny = 1e-4000L, b = x + 1 - a, c = 1 / tiny, d = 1 / b, h = d;
    for (int i = 1; i < 1000; i++) {
        long double an = -(long double)i * ((long double)i - a);
        b += 2;
        d = an * d + b; if (fabsl(d) < tiny) d = tiny;
        c = b + an / c; if (fabsl(c) < tiny) c = tiny;
        d = 1 / d;
        long double del = d * c; h *= del;
        if (fabsl(del - 1) < 1e-20L) break;
    }
    return (double)(1 - lead * h);
}
#endif
