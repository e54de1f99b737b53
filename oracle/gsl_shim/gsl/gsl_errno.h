/* oracle/gsl_shim: TEST INFRASTRUCTURE.  Stand-in for <gsl/gsl_errno.h> (time_interp.h:8 includes it for GSL_SUCCESS). */
#ifndef GB_SHIM_GSL_ERRNO_H
#define GB_SHIM_GSL_ERRNO_H
#ifndef GSL_SUCCESS
#define GSL_SUCCESS 0
#endif
#endif
