// oracle/ref_timeinterp.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Builds the reference's TimeInterpState for one GB_POT_TIMEINTERP component out of the reference's OWN functions
// (potential/potential/builtin/time_interp.cpp, time_interp_wrapper.cpp; compiled in place by oracle/Makefile with
// USE_GSL == 1 against oracle/gsl_shim's spline stand-in) and restates only what the Cython wrapper does around them:
// TimeInterpolatedWrapper.__init__, potential/potential/builtin/cytimeinterp.pyx:157-300.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "potential/src/cpotential.h"
#include "potential/builtin/time_interp.h"
#include "potential/builtin/time_interp_wrapper.h"

// params = [G, wrapped_type, method, n_knots, n_wpar, n_origin, n_R, t_knots[n], wpar[n][n_wpar], origin[n_origin][3],
// R[n_R][9]] (include/gala_b200.h, GB_POT_TIMEINTERP).  Every wrapped C parameter is handed over as its own
// one-element interpolated parameter -- what the reference itself does for classes with a parameter transform
// (time_interpolated.py:300-330 "_c{j}" entries); time_interp_init_param turns constant ones into constants.
extern "C" void *gb_ref_ti_build(const double *params, int n_params, CPotential *wrapped) {
    if (n_params < 7) return NULL;
    const int method = (int)params[2], n = (int)params[3], nwp = (int)params[4], no = (int)params[5], nR = (int)params[6];
    const gsl_interp_type *T = method == 0 ? gsl_interp_linear : method == 1 ? gsl_interp_cspline
                             : method == 2 ? gsl_interp_akima : gsl_interp_steffen;
    std::vector<double> tk(params + 7, params + 7 + n);
    const double *wv = params + 7 + n, *ov = wv + (size_t)n * nwp, *Rv = ov + (size_t)3 * no;
    TimeInterpState *st = time_interp_alloc(1 + nwp, 3, T);                    // cytimeinterp.pyx:181
    if (!st) return NULL;
    st->t_min = tk[0]; st->t_max = tk[n - 1];                                   // :188-189
    double G = params[0];
    if (time_interp_init_constant_param(&st->params[0], &G, 1) != 0) { time_interp_free(st); return NULL; }   // :196-201
    std::vector<double> col(n);
    for (int k = 0; k < nwp; k++) {
        for (int i = 0; i < n; i++) col[i] = wv[(size_t)i * nwp + k];
        if (time_interp_init_param(&st->params[1 + k], tk.data(), col.data(), n, 1, T) != 0) { time_interp_free(st); return NULL; }   // :226-230
    }
    int rc;
    std::vector<double> oflat(ov, ov + 3 * no);
    if (no == 1) rc = time_interp_init_constant_param(&st->origin, oflat.data(), 3);          // :243-246
    else rc = time_interp_init_param(&st->origin, tk.data(), oflat.data(), n, 3, T);           // :248-253
    if (rc != 0) { time_interp_free(st); return NULL; }
    std::vector<double> Rflat(Rv, Rv + 9 * nR);
    if (nR == 1) time_interp_init_constant_rotation(&st->rotation, Rflat.data());             // :268-271
    else time_interp_init_rotation(&st->rotation, tk.data(), Rflat.data(), n, T);             // :273-278
    st->wrapped_potential = (void *)wrapped;                                                   // :283
    return st;
}
extern "C" void gb_ref_ti_free(void *state) { if (state) time_interp_free((TimeInterpState *)state); }
// function pointers + state of the wrapper's own CPotential (cytimeinterp.pyx:292-299)
extern "C" void gb_ref_ti_hook(CPotential *cp, int i, void *state) {
    cp->value[i] = (energyfunc)time_interp_value;
    cp->gradient[i] = (gradientfunc)time_interp_gradient;
    cp->density[i] = (densityfunc)time_interp_density;
    cp->hessian[i] = (hessianfunc)time_interp_hessian;
    cp->state[i] = state;
}
