/* oracle/gsl_shim: TEST INFRASTRUCTURE.  Minimal stand-in for <gsl/gsl_math.h>; the reference's SCF
 * sources (potential/scf/src/bfe.cpp:9-12) include it only for M_PI-class constants. */
#ifndef GB_SHIM_GSL_MATH_H
#define GB_SHIM_GSL_MATH_H
#include <math.h>
#endif
