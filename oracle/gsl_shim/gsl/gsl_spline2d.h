/* oracle/gsl_shim: TEST INFRASTRUCTURE.  Declarations-only stand-in for <gsl/gsl_spline2d.h>; see
 * gsl_interp2d.h.  Calling any of these aborts. */
#ifndef GB_SHIM_GSL_SPLINE2D_H
#define GB_SHIM_GSL_SPLINE2D_H
#include "gsl_interp2d.h"
typedef struct { int unused; } gsl_spline2d;
static inline gsl_spline2d *gsl_spline2d_alloc(const gsl_interp2d_type *T, size_t nx, size_t ny) { (void)T; (void)nx; (void)ny; abort(); return NULL; }
static inline int gsl_spline2d_init(gsl_spline2d *s, const double *x, const double *y, const double *z, size_t nx, size_t ny) { (void)s; (void)x; (void)y; (void)z; (void)nx; (void)ny; abort(); return 0; }
static inline void gsl_spline2d_free(gsl_spline2d *s) { (void)s; abort(); }
#define GB_SHIM_EVAL(name) static inline double name(const gsl_spline2d *s, double x, double y, gsl_interp_accel *xa, gsl_interp_accel *ya) { (void)s; (void)x; (void)y; (void)xa; (void)ya; abort(); return 0; }
GB_SHIM_EVAL(gsl_spline2d_eval)
GB_SHIM_EVAL(gsl_spline2d_eval_deriv_x)
GB_SHIM_EVAL(gsl_spline2d_eval_deriv_y)
GB_SHIM_EVAL(gsl_spline2d_eval_deriv_xx)
GB_SHIM_EVAL(gsl_spline2d_eval_deriv_yy)
#endif
