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
#endif
