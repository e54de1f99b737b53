// hessian.cuh -- Hessian of the potential by forward-mode automatic differentiation of the gradient.
//
// Reference: c_hessian (potential/potential/src/cpotential.cpp:290-314) sums per-component `*_hessian`
// functions that are sympy-generated expression dumps (builtin_potentials.cpp:41-54, 90-125, ... 1813+).
// Here nothing is transcribed: every gradient formula of potentials.cuh is written once more over a
// generic scalar T, and the Hessian row i is the derivative part of g_i evaluated on dual numbers
// (value + 3 partials).  H_ij = d g_i / d q_j, exact to rounding.  Like the reference, components with a
// rotation are refused (cpotential.cpp:307-309, core.py "NotImplementedError"); a shifted origin is applied.
// Potentials the reference has NO Hessian for (Burkert, LeeSuto, Kuzmin: null/absent entries in
// cybuiltin.pyx:226-262,336-345) get the true one here.
#pragma once

struct Dual3 {
    double v, d[3];
};
GB_DEV Dual3 mk(double v, double a, double b, double c) { Dual3 r; r.v = v; r.d[0] = a; r.d[1] = b; r.d[2] = c; return r; }
GB_DEV Dual3 operator+(const Dual3& a, const Dual3& b) { return mk(a.v + b.v, a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2]); }
GB_DEV Dual3 operator-(const Dual3& a, const Dual3& b) { return mk(a.v - b.v, a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]); }
GB_DEV Dual3 operator-(const Dual3& a) { return mk(-a.v, -a.d[0], -a.d[1], -a.d[2]); }
GB_DEV Dual3 operator*(const Dual3& a, const Dual3& b) {
    return mk(a.v * b.v, a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2]);
}
GB_DEV Dual3 operator/(const Dual3& a, const Dual3& b) {
    const double q = a.v / b.v, ib = 1. / b.v;
    return mk(q, (a.d[0] - q * b.d[0]) * ib, (a.d[1] - q * b.d[1]) * ib, (a.d[2] - q * b.d[2]) * ib);
}
GB_DEV Dual3 operator+(const Dual3& a, double s) { return mk(a.v + s, a.d[0], a.d[1], a.d[2]); }
GB_DEV Dual3 operator+(double s, const Dual3& a) { return a + s; }
GB_DEV Dual3 operator-(const Dual3& a, double s) { return mk(a.v - s, a.d[0], a.d[1], a.d[2]); }
GB_DEV Dual3 operator-(double s, const Dual3& a) { return mk(s - a.v, -a.d[0], -a.d[1], -a.d[2]); }
GB_DEV Dual3 operator*(const Dual3& a, double s) { return mk(a.v * s, a.d[0] * s, a.d[1] * s, a.d[2] * s); }
GB_DEV Dual3 operator*(double s, const Dual3& a) { return a * s; }
GB_DEV Dual3 operator/(const Dual3& a, double s) { return a * (1. / s); }
GB_DEV Dual3 operator/(double s, const Dual3& a) { const double q = s / a.v, f = -q / a.v; return mk(q, f * a.d[0], f * a.d[1], f * a.d[2]); }
GB_DEV Dual3 gb_chain(const Dual3& a, double f, double df) { return mk(f, df * a.d[0], df * a.d[1], df * a.d[2]); }
GB_DEV Dual3 gb_sqrt(const Dual3& a) { const double s = sqrt(a.v); return gb_chain(a, s, 0.5 / s); }
GB_DEV Dual3 gb_ln(const Dual3& a) { return gb_chain(a, log(a.v), 1. / a.v); }
GB_DEV Dual3 gb_atan(const Dual3& a) { return gb_chain(a, atan(a.v), 1. / (1. + a.v * a.v)); }
GB_DEV Dual3 gb_abs(const Dual3& a) { return (a.v < 0.) ? -a : a; }
GB_DEV Dual3 gb_powr(const Dual3& a, double e) { const double f = pow(a.v, e); return gb_chain(a, f, e * f / a.v); }
GB_DEV double gb_powr(double a, double e) { return pow(a, e); }
GB_DEV double gb_val(double a) { return a; }
GB_DEV double gb_val(const Dual3& a) { return a.v; }
// regularised incomplete gamma function P(a, x) and d/dx = x^(a-1) e^-x / Gamma(a)
GB_DEV double gb_gamma_inc_P_t(double a, double x) { return gb_gamma_inc_P(a, x); }
GB_DEV Dual3 gb_gamma_inc_P_t(double a, const Dual3& x) {
    return gb_chain(x, gb_gamma_inc_P(a, x.v), exp((a - 1.) * log(x.v) - x.v - lgamma(a)));
}

// Gradient of ONE simple component over a generic scalar (accumulating, like PotXxx::gradient).
template <class T>
GB_DEV void gb_grad_t(int type, const double* p, const T& x, const T& y, const T& z, T& gx, T& gy, T& gz) {
    const T r2 = x * x + y * y + z * z;
    switch (type) {
        case GB_POT_KEPLER: {
            const T f = (p[0] * p[1]) * gb_powr(r2, -1.5);
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z;
        } break;
        case GB_POT_ISOCHRONE: {
            const T s = gb_sqrt(r2 + p[2] * p[2]);
            const T f = (p[0] * p[1]) / (s * (s + p[2]) * (s + p[2]));
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z;
        } break;
        case GB_POT_HERNQUIST: {
            const T r = gb_sqrt(r2);
            const T f = (p[0] * p[1]) / ((r + p[2]) * (r + p[2]) * r);
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z;
        } break;
        case GB_POT_PLUMMER: {
            const T f = (p[0] * p[1]) * gb_powr(r2 + p[2] * p[2], -1.5);
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z;
        } break;
        case GB_POT_JAFFE: {
            const T r = gb_sqrt(r2);
            const T f = (p[0] * p[1]) / (r2 * (r + p[2]));
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z;
        } break;
        case GB_POT_NFW_SPHERICAL: case GB_POT_NFW_FLATTENED: case GB_POT_NFW_TRIAXIAL: {
            const double ia2 = (type == GB_POT_NFW_TRIAXIAL) ? 1. / (p[3] * p[3]) : 1.;
            const double ib2 = (type == GB_POT_NFW_TRIAXIAL) ? 1. / (p[4] * p[4]) : 1.;
            const double ic2 = (type == GB_POT_NFW_SPHERICAL) ? 1. : 1. / (p[5] * p[5]);
            const T u = gb_sqrt(x * x * ia2 + y * y * ib2 + z * z * ic2) / p[2];
            const T f = (p[0] * p[1] / p[2]) / (u * u * u) / (p[2] * p[2]) * (gb_ln(1. + u) - u / (1. + u));
            gx = gx + f * x * ia2; gy = gy + f * y * ib2; gz = gz + f * z * ic2;
        } break;
        case GB_POT_MIYAMOTONAGAI: case GB_POT_MN3: {
            const int nd = (type == GB_POT_MN3) ? 3 : 1;
            for (int i = 0; i < nd; i++) {
                const double m = p[1 + 3 * i], a = p[2 + 3 * i], b = p[3 + 3 * i];
                const T sz = gb_sqrt(z * z + b * b);
                const T zd = a + sz;
                const T f = (p[0] * m) * gb_powr(x * x + y * y + zd * zd, -1.5);
                gx = gx + f * x; gy = gy + f * y; gz = gz + f * z * (1. + a / sz);
            }
        } break;
        case GB_POT_SATOH: {
            const T zb = gb_sqrt(z * z + p[3] * p[3]);
            const T S2 = r2 + p[2] * (p[2] + 2. * zb);
            const T f = (p[0] * p[1]) * gb_powr(S2, -1.5);
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z * (1. + p[2] / zb);
        } break;
        case GB_POT_KUZMIN: {
            const T az = p[2] + gb_abs(z);
            const T f = (p[0] * p[1]) * gb_powr(x * x + y * y + az * az, -1.5);
            const double zs = (gb_val(z) > 0) ? 1. : ((gb_val(z) < 0) ? -1. : 0.);
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * zs * az;
        } break;
        case GB_POT_STONE: {
            const T r = gb_sqrt(r2);
            const T f = (2 * p[0] * p[1] / GB_PI / (p[2] - p[3])) / (r2 * r) * (p[2] * gb_atan(r / p[2]) - p[3] * gb_atan(r / p[3]));
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z;
        } break;
        case GB_POT_BURKERT: {
            const T r = gb_sqrt(r2);
            const T u = r / p[2];
            const T dphi = (-GB_PI * p[0] * p[1] * p[2]) / (u * u) * (2. * gb_atan(u) - 2. * gb_ln(1. + u) - gb_ln(1. + u * u));
            const T f = dphi / r;
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z;
        } break;
        case GB_POT_POWERLAWCUTOFF: {
            const T f = (p[0] * p[1]) * gb_powr(r2, -1.5) * gb_gamma_inc_P_t(0.5 * (3 - p[2]), r2 / (p[3] * p[3]));
            gx = gx + f * x; gy = gy + f * y; gz = gz + f * z;
        } break;
        case GB_POT_LOGARITHMIC: {
            const double sp = sin(p[6]), cp = cos(p[6]);
            const T X = x * cp + y * sp, Y = y * cp - x * sp;
            const double i1 = 1. / (p[3] * p[3]), i2 = 1. / (p[4] * p[4]), i3 = 1. / (p[5] * p[5]);
            const T f = (p[1] * p[1]) / (p[2] * p[2] + X * X * i1 + Y * Y * i2 + z * z * i3);
            const T ax = f * X * i1, ay = f * Y * i2;
            gx = gx + (ax * cp - ay * sp); gy = gy + (ax * sp + ay * cp); gz = gz + f * z * i3;
        } break;
        case GB_POT_LONGMURALIBAR: {
            const double sa = sin(p[5]), ca = cos(p[5]);
            const T X = x * ca + y * sa, Y = y * ca - x * sa;
            const double a = p[2], b = p[3], c = p[4];
            const T zc = gb_sqrt(z * z + c * c);
            const T bcz = b + zc;
            const T S = Y * Y + bcz * bcz;
            const T Tm = gb_sqrt((a - X) * (a - X) + S), Tp = gb_sqrt((a + X) * (a + X) + S);
            const T f1 = (p[0] * p[1]) / (2. * Tm * Tp);
            const T f3 = Tp + Tm - (4. * X * X) / (Tp + Tm);
            const T GX = 4. * f1 * X / (Tp + Tm);
            const T GY = f1 * Y * f3 / S;
            gx = gx + (GX * ca - GY * sa); gy = gy + (GX * sa + GY * ca);
            gz = gz + f1 * z * f3 / S * bcz / zc;
        } break;
        case GB_POT_LEESUTO: {
            const double ba = p[4] / p[3], ca = p[5] / p[3];
            const double eb = 1 - ba * ba, ec = 1 - ca * ca, ln2 = 0.6931471805599453, rs = p[2];
            const double vh2 = p[1] * p[1] / (ln2 - 0.5 + (ln2 - 0.75) * eb + (ln2 - 0.75) * ec);
            const T r = gb_sqrt(r2), r4 = r2 * r2;
            const T x0 = r + rs, x1 = x0 * x0;
            const T x2 = vh2 / (12. * r4 * r2 * r * x1);
            const T x10 = gb_ln(x0 / rs);
            const T x13 = r * (3. * rs), x15 = x13 - r2, x16 = x15 + 6. * (rs * rs);
            const T x17 = (6. * rs) * x0 * (r * x16 - x0 * x10 * (6. * (rs * rs)));
            const T x20 = x0 * r2, x21 = 2. * r * x0;
            const T x7 = eb * y * y + ec * z * z;
            const T x22 = -12. * r4 * r * rs * x0 + 12. * r4 * rs * x1 * x10 +
                          (3. * rs) * x7 * (x16 * r2 - 18. * x1 * x10 * (rs * rs) + x20 * (2. * r - 3. * rs) + x21 * (x15 + 9. * (rs * rs))) -
                          x20 * (eb + ec) * (-6. * r * rs * (r2 - (rs * rs)) + (6. * rs) * x0 * x10 * (r2 - 3. * (rs * rs)) +
                                             x20 * (-4. * r + 3. * rs) + x21 * (2. * r2 - x13 + 6. * (rs * rs)));
            gx = gx + x2 * x * (x17 * x7 + x22);
            gy = gy + x2 * y * (x17 * (x7 - r2 * eb) + x22);
            gz = gz + x2 * z * (x17 * (x7 - r2 * ec) + x22);
        } break;
        default: break;   // Null; SCF / multipole are refused by the host (the reference has no Hessian for them either)
    }
}

// hess (3,3,N): H[i][j][n] = d^2 Phi / dq_i dq_j at point n (CPotentialWrapper.hessian, cpotential.pyx:164-182,
// returned by PotentialBase.hessian as (n_dim, n_dim, N), core.py:535-600)
// A TimeInterpolated component (time_interp_hessian, time_interp_wrapper.cpp:254-318): the wrapped potential's
// Hessian at the interpolated parameters in body coordinates X = R (q - o), turned back as R^T H R -- here the dual
// parts simply ride through the two linear maps; NaN outside the knots.
__global__ void k_eval_hessian(const __grid_constant__ DevPot P, const double* __restrict__ q, double t, size_t N,
                               double* __restrict__ hess) {
    const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    Dual3 g[3] = {mk(0, 0, 0, 0), mk(0, 0, 0, 0), mk(0, 0, 0, 0)};
    for (int i = 0; i < P.n; i++) {
        const DevComp& c = P.c[i];
        if (c.type == GB_POT_TIMEINTERP) {
            const TiView v = ti_view(&P.par[c.poff], P.ext + c.eoff);
            double wp[GB_TI_MAXPAR], o[3], R[9];
            if (!ti_state(v, t, wp, o, R)) {
                for (int k = 0; k < 3; k++) g[k] = mk(CUDART_NAN, CUDART_NAN, CUDART_NAN, CUDART_NAN);
                continue;
            }
            const Dual3 sx = mk(q[n] - o[0], 1, 0, 0), sy = mk(q[N + n] - o[1], 0, 1, 0), sz = mk(q[2 * N + n] - o[2], 0, 0, 1);
            const Dual3 X = R[0] * sx + R[1] * sy + R[2] * sz, Y = R[3] * sx + R[4] * sy + R[5] * sz,
                        Z = R[6] * sx + R[7] * sy + R[8] * sz;
            Dual3 b[3] = {mk(0, 0, 0, 0), mk(0, 0, 0, 0), mk(0, 0, 0, 0)};
            gb_grad_t<Dual3>(v.wtype, wp, X, Y, Z, b[0], b[1], b[2]);
            g[0] = g[0] + (R[0] * b[0] + R[3] * b[1] + R[6] * b[2]);
            g[1] = g[1] + (R[1] * b[0] + R[4] * b[1] + R[7] * b[2]);
            g[2] = g[2] + (R[2] * b[0] + R[5] * b[1] + R[8] * b[2]);
            continue;
        }
        const double sx = c.shift ? c.q0[0] : 0., sy = c.shift ? c.q0[1] : 0., sz = c.shift ? c.q0[2] : 0.;
        const Dual3 x = mk(q[n] - sx, 1, 0, 0), y = mk(q[N + n] - sy, 0, 1, 0), z = mk(q[2 * N + n] - sz, 0, 0, 1);
        gb_grad_t<Dual3>(c.type, &P.par[c.poff], x, y, z, g[0], g[1], g[2]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) hess[(size_t)(i * 3 + j) * N + n] = g[i].d[j];
}
