// timeinterp.cuh -- TimeInterpolatedPotential on the device.
//
// Reference: potential/potential/builtin/time_interpolated.py:29 (host class), time_interp.cpp:181-440 (one GSL
// spline per parameter element, per origin component and per axis-angle component of the rotation), and
// time_interp_wrapper.cpp:91-323 (value / gradient / density = interpolate everything at t, shift-rotate q with the
// interpolated origin and rotation, evaluate the WRAPPED potential with the interpolated parameter vector, rotate
// the gradient back; NaN outside [t_min, t_max], :103-106,148-155).
//
// Here the host (capi.cu:ti_pack) turns every interpolated quantity into one table of cubic pieces in the form GSL
// evaluates, y = y_i + dx (b_i + dx (c_i + dx d_i)) on [x_i, x_{i+1}) (linear, natural cspline, non-periodic Akima
// and Steffen all fit it; a constant is a table with b = c = d = 0), so the device needs one code path:
//   ext block of the component:  knots[n] | table[nel][n-1][4] (y, b, c, d) | constR[9]
//   nel = n_wpar + 1 (G first, like CPotentialWrapper._params) + 3 (origin) + 4 (rotation axis x, y, z, angle)
// A thread evaluates at ITS time: in the fixed-step kernels every lane passes the same t[j] (the loads are warp-
// uniform), in DOP853 every lane has its own stage time.
#pragma once

#define GB_TI_MAXPAR 16          // wrapped parameter vector incl. G (MN3: 13)

struct TiView {
    int wtype, n, nwp, rot_const;
    double tmin, tmax;
    const double* knots;
    const double* tab;
    const double* constR;
};
// small parameters of the component in the constant bank: [G, wrapped_type, method, n_knots, n_wpar, rot_const, t_min, t_max]
GB_DEV TiView ti_view(const double* p, const double* e) {
    TiView v;
    v.wtype = (int)p[1]; v.n = (int)p[3]; v.nwp = (int)p[4]; v.rot_const = (int)p[5];
    v.tmin = p[6]; v.tmax = p[7];
    v.knots = e;
    v.tab = e + v.n;
    v.constR = v.tab + (size_t)(v.nwp + 1 + 3 + 4) * (v.n - 1) * 4;
    return v;
}
// interval i with knots[i] <= t < knots[i+1]; t == knots[n-1] belongs to the last one (gsl_interp_bsearch)
GB_DEV int ti_interval(const TiView& v, double t) {
    int lo = 0, hi = v.n - 1;
    while (hi > lo + 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(v.knots + mid) > t) hi = mid; else lo = mid;
    }
    return lo;
}
GB_DEV double ti_piece(const TiView& v, int el, int i, double dx) {
    const double* c = v.tab + ((size_t)el * (v.n - 1) + i) * 4;
    const double2 yb = __ldg(reinterpret_cast<const double2*>(c)), cd = __ldg(reinterpret_cast<const double2*>(c) + 1);
    return yb.x + dx * (yb.y + dx * (cd.x + dx * cd.y));
}
// everything the wrapped potential needs at time t.  Returns false outside [t_min, t_max].
GB_DEV bool ti_state(const TiView& v, double t, double (&wp)[GB_TI_MAXPAR], double (&o)[3], double (&R)[9]) {
    if (t < v.tmin || t > v.tmax || !(t == t)) return false;
    const int i = ti_interval(v, t);
    const double dx = t - __ldg(v.knots + i);
    const int npar = v.nwp + 1;
    for (int k = 0; k < npar; k++) wp[k] = ti_piece(v, k, i, dx);
    for (int k = 0; k < 3; k++) o[k] = ti_piece(v, npar + k, i, dx);
    if (v.rot_const) {
        for (int k = 0; k < 9; k++) R[k] = __ldg(v.constR + k);
    } else {
        // axis_angle_to_rotation_matrix (time_interp.cpp:516-536) of the four interpolated components
        const double x = ti_piece(v, npar + 3, i, dx), y = ti_piece(v, npar + 4, i, dx), z = ti_piece(v, npar + 5, i, dx);
        const double ang = ti_piece(v, npar + 6, i, dx);
        const double c = cos(ang), s = sin(ang), C = 1.0 - c;
        R[0] = x * x * C + c;     R[1] = x * y * C - z * s; R[2] = x * z * C + y * s;
        R[3] = y * x * C + z * s; R[4] = y * y * C + c;     R[5] = y * z * C - x * s;
        R[6] = z * x * C - y * s; R[7] = z * y * C + x * s; R[8] = z * z * C + c;
    }
    return true;
}
// shift to the interpolated origin and rotate: apply_shift_rotate (cpotential.cpp:151-167)
GB_DEV void ti_to_body(const double (&o)[3], const double (&R)[9], double x, double y, double z, double& X, double& Y, double& Z) {
    const double sx = x - o[0], sy = y - o[1], sz = z - o[2];
    X = R[0] * sx + R[1] * sy + R[2] * sz;
    Y = R[3] * sx + R[4] * sy + R[5] * sz;
    Z = R[6] * sx + R[7] * sy + R[8] * sz;
}
