/* oracle/gsl_shim: TEST INFRASTRUCTURE.  Stand-in for <gsl/gsl_spline.h> / <gsl/gsl_interp.h>.
 * bfe.cpp includes it (potential/scf/src/bfe.cpp:11) but uses no spline symbol; builtin_potentials.cpp
 * compiled with USE_GSL == 1 (needed for powerlawcutoff_*, :465-660) also compiles the spherical-spline
 * potentials (:1950-2230), which are OUT OF SCOPE here: the symbols they reference are declared so the
 * file compiles, and abort if ever called. */
#ifndef GB_SHIM_GSL_SPLINE_H
#define GB_SHIM_GSL_SPLINE_H
#include <stdio.h>
#include <stdlib.h>
typedef struct { int dummy; } gsl_interp_type;
typedef struct { int dummy; } gsl_interp;
typedef struct { int dummy; } gsl_interp_accel;
typedef struct { int dummy; } gsl_spline;
static inline double gb_shim_no_spline(void) {
    fprintf(stderr, "oracle/gsl_shim: GSL 1-D splines are not implemented (out of scope)\n"); abort(); return 0.;
}
static inline double gsl_spline_eval(const gsl_spline *, double, gsl_interp_accel *) { return gb_shim_no_spline(); }
static inline double gsl_spline_eval_deriv(const gsl_spline *, double, gsl_interp_accel *) { return gb_shim_no_spline(); }
static inline double gsl_spline_eval_deriv2(const gsl_spline *, double, gsl_interp_accel *) { return gb_shim_no_spline(); }
static inline double gsl_spline_eval_integ(const gsl_spline *, double, double, gsl_interp_accel *) { return gb_shim_no_spline(); }
#endif
