/* oracle/gsl_shim: TEST INFRASTRUCTURE.  See gsl_spline.h (same stand-in). */
#include "gsl_spline.h"
