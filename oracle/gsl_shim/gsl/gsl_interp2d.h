/* oracle/gsl_shim: TEST INFRASTRUCTURE.  Declarations-only stand-in for <gsl/gsl_interp2d.h>:
 * potential/potential/builtin/multipole.cpp includes it for the CylSpline functions in the same
 * file (:640-830), which the oracle never calls (CylSpline is out of scope).
 * Calling any of these aborts. */
#ifndef GB_SHIM_GSL_INTERP2D_H
#define GB_SHIM_GSL_INTERP2D_H
#include <stdlib.h>
#include <stddef.h>
typedef struct { int unused; } gsl_interp2d_type;
typedef struct { int unused; } gsl_interp_accel;
static const gsl_interp2d_type gb_shim_bicubic = {0};
static const gsl_interp2d_type *const gsl_interp2d_bicubic = &gb_shim_bicubic;
static inline gsl_interp_accel *gsl_interp_accel_alloc(void) { abort(); return NULL; }
static inline void gsl_interp_accel_free(gsl_interp_accel *a) { (void)a; abort(); }
#endif
