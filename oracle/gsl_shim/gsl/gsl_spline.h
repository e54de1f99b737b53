/* oracle/gsl_shim: TEST INFRASTRUCTURE.  Empty stand-in for <gsl/gsl_spline.h>: bfe.cpp includes it
 * (potential/scf/src/bfe.cpp:11) but uses no spline symbol. */
#ifndef GB_SHIM_GSL_SPLINE_H
#define GB_SHIM_GSL_SPLINE_H
#endif
