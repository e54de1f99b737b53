/* oracle/gsl_shim: TEST INFRASTRUCTURE.  Stand-in for <gsl/gsl_spline.h> / <gsl/gsl_interp.h>.
 *
 * GSL (GNU Scientific Library; a system dependency of the reference, unpinned, >= 1.16, setup.py:89-103) is absent
 * from this image.  The reference's TimeInterpolatedPotential (potential/potential/builtin/time_interp.cpp:181-320,
 * 415-440) needs gsl_spline_alloc / _init / _eval / _free and gsl_interp_accel_alloc / _free for the four
 * interpolation types it offers (time_interpolated.py:44-60): linear, cspline (natural cubic spline), akima
 * (non-periodic) and steffen.  They are restated here from the published algorithms, in the evaluation form GSL
 * documents (y = y_i + dx (b_i + dx (c_i + dx d_i)) on the interval [x_i, x_{i+1}) found by bisection, the last
 * interval closed on the right):
 *   cspline  natural cubic spline: tridiagonal system for the second-derivative coefficients c_i with
 *            c_0 = c_{n-1} = 0 (de Boor; GSL interpolation/cspline.c)
 *   akima    Akima (1970) with the end extension m_{-1} = 2 m_0 - m_1, m_{-2} = 2 m_{-1} - m_0 (and mirrored at the
 *            right end) and slope t_i = m_i when both weights vanish (GSL interpolation/akima.c)
 *   steffen  Steffen (1990) monotone cubic: y'_i = (sign s_{i-1} + sign s_i) min(|s_{i-1}|, |s_i|, |p_i| / 2),
 *            p_i = (s_{i-1} h_i + s_i h_{i-1}) / (h_{i-1} + h_i); end slopes = the end secants (GSL steffen.c)
 * Pinned in tests/test_oracle_cpu.py against scipy.interpolate (CubicSpline(bc_type="natural"), Akima1DInterpolator,
 * np.interp) and, for steffen, against an independent numpy statement of the 1990 formulas + its monotonicity.
 *
 * The spherical-spline potentials (builtin_potentials.cpp:1950-2230) also reference gsl_spline_eval_deriv / _integ;
 * they are OUT OF SCOPE: declared so the file compiles, abort if ever called. */
#ifndef GB_SHIM_GSL_SPLINE_H
#define GB_SHIM_GSL_SPLINE_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef GSL_SUCCESS
#define GSL_SUCCESS 0
#endif

typedef struct { int kind; } gsl_interp_type;      /* 0 linear, 1 cspline, 2 akima, 3 steffen */
typedef struct { int dummy; } gsl_interp;
typedef struct { size_t cache; } gsl_interp_accel;
typedef struct {
    int kind;
    size_t size;
    double *x, *y;            /* copies of the knots (gsl_spline keeps its own, interpolation/spline.c) */
    double *b, *c, *d;        /* per interval: y = y_i + dx (b_i + dx (c_i + dx d_i)) */
} gsl_spline;

static const gsl_interp_type gb_shim_interp_types[4] = {{0}, {1}, {2}, {3}};
#define gsl_interp_linear  (&gb_shim_interp_types[0])
#define gsl_interp_cspline (&gb_shim_interp_types[1])
#define gsl_interp_akima   (&gb_shim_interp_types[2])
#define gsl_interp_steffen (&gb_shim_interp_types[3])

static inline gsl_interp_accel *gsl_interp_accel_alloc(void) { return (gsl_interp_accel *)calloc(1, sizeof(gsl_interp_accel)); }
static inline void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

static inline gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size) {
    size_t min_size = (T->kind == 0) ? 2 : (T->kind == 1) ? 3 : (T->kind == 2) ? 5 : 3;      /* GSL's min_size per type */
    if (size < min_size) return NULL;
    gsl_spline *s = (gsl_spline *)calloc(1, sizeof(gsl_spline));
    if (!s) return NULL;
    s->kind = T->kind; s->size = size;
    s->x = (double *)malloc(size * sizeof(double)); s->y = (double *)malloc(size * sizeof(double));
    s->b = (double *)calloc(size, sizeof(double)); s->c = (double *)calloc(size, sizeof(double));
    s->d = (double *)calloc(size, sizeof(double));
    return s;
}
static inline void gsl_spline_free(gsl_spline *s) {
    if (!s) return;
    free(s->x); free(s->y); free(s->b); free(s->c); free(s->d); free(s);
}

static inline int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t size) {
    if (!s || size != s->size) return -1;
    const size_t n = size;
    memcpy(s->x, xa, n * sizeof(double)); memcpy(s->y, ya, n * sizeof(double));
    for (size_t i = 0; i + 1 < n; i++) if (!(xa[i + 1] > xa[i])) return -1;      /* x must be strictly increasing */
    if (s->kind == 0) {
        for (size_t i = 0; i + 1 < n; i++) { s->b[i] = (ya[i + 1] - ya[i]) / (xa[i + 1] - xa[i]); s->c[i] = 0.; s->d[i] = 0.; }
    } else if (s->kind == 1) {
        /* natural cubic spline: cc[0] = cc[n-1] = 0; interior from the symmetric tridiagonal system
         * h_{i-1} cc_{i-1} + 2 (h_{i-1} + h_i) cc_i + h_i cc_{i+1} = 3 (dy_i / h_i - dy_{i-1} / h_{i-1}) */
        double *cc = (double *)calloc(n, sizeof(double));
        const size_t m = n - 2;                                    /* unknowns cc[1..n-2] */
        if (m >= 1) {
            double *diag = (double *)malloc(m * sizeof(double)), *off = (double *)malloc(m * sizeof(double));
            double *g = (double *)malloc(m * sizeof(double));
            for (size_t i = 0; i < m; i++) {
                const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
                const double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
                off[i] = h_ip1; diag[i] = 2.0 * (h_ip1 + h_i);
                g[i] = 3.0 * (yd_ip1 / h_ip1 - yd_i / h_i);
            }
            /* Thomas algorithm (the system is symmetric positive definite and diagonally dominant) */
            for (size_t i = 1; i < m; i++) {
                const double w = off[i - 1] / diag[i - 1];
                diag[i] -= w * off[i - 1]; g[i] -= w * g[i - 1];
            }
            cc[m] = g[m - 1] / diag[m - 1];
            for (size_t i = m - 1; i-- > 0;) cc[i + 1] = (g[i] - off[i] * cc[i + 2]) / diag[i];
            free(diag); free(off); free(g);
        }
        for (size_t i = 0; i + 1 < n; i++) {
            const double dx = xa[i + 1] - xa[i], dy = ya[i + 1] - ya[i];
            s->b[i] = dy / dx - dx * (cc[i + 1] + 2.0 * cc[i]) / 3.0;
            s->c[i] = cc[i];
            s->d[i] = (cc[i + 1] - cc[i]) / (3.0 * dx);
        }
        free(cc);
    } else if (s->kind == 2) {
        /* Akima: m[i] = secant of interval i, i = 0..n-2, extended by two on each side */
        double *mm = (double *)malloc((n + 3) * sizeof(double));
        double *m = mm + 2;
        for (size_t i = 0; i + 1 < n; i++) m[i] = (ya[i + 1] - ya[i]) / (xa[i + 1] - xa[i]);
        m[-2] = 3.0 * m[0] - 2.0 * m[1];
        m[-1] = 2.0 * m[0] - m[1];
        m[n - 1] = 2.0 * m[n - 2] - m[n - 3];
        m[n] = 3.0 * m[n - 2] - 2.0 * m[n - 3];
        for (size_t ii = 0; ii + 1 < n; ii++) {
            const long i = (long)ii;
            const double NE = fabs(m[i + 1] - m[i]) + fabs(m[i - 1] - m[i - 2]);
            if (NE == 0.0) { s->b[ii] = m[i]; s->c[ii] = 0.0; s->d[ii] = 0.0; continue; }
            const double h_i = xa[ii + 1] - xa[ii];
            const double NE_next = fabs(m[i + 2] - m[i + 1]) + fabs(m[i] - m[i - 1]);
            const double alpha_i = fabs(m[i - 1] - m[i - 2]) / NE;
            double tL_ip1;
            if (NE_next == 0.0) tL_ip1 = m[i];
            else { const double alpha_ip1 = fabs(m[i] - m[i - 1]) / NE_next; tL_ip1 = (1.0 - alpha_ip1) * m[i] + alpha_ip1 * m[i + 1]; }
            s->b[ii] = (1.0 - alpha_i) * m[i - 1] + alpha_i * m[i];
            s->c[ii] = (3.0 * m[i] - 2.0 * s->b[ii] - tL_ip1) / h_i;
            s->d[ii] = (s->b[ii] + tL_ip1 - 2.0 * m[i]) / (h_i * h_i);
        }
        free(mm);
    } else {
        /* Steffen (1990): slopes yp[i] at the knots */
        double *yp = (double *)malloc(n * sizeof(double));
        yp[0] = (ya[1] - ya[0]) / (xa[1] - xa[0]);
        yp[n - 1] = (ya[n - 1] - ya[n - 2]) / (xa[n - 1] - xa[n - 2]);
        for (size_t i = 1; i + 1 < n; i++) {
            const double hi = xa[i + 1] - xa[i], him1 = xa[i] - xa[i - 1];
            const double si = (ya[i + 1] - ya[i]) / hi, sim1 = (ya[i] - ya[i - 1]) / him1;
            const double pi = (sim1 * hi + si * him1) / (him1 + hi);
            const double sg = (sim1 > 0 ? 1. : (sim1 < 0 ? -1. : 0.)) + (si > 0 ? 1. : (si < 0 ? -1. : 0.));
            yp[i] = sg * fmin(fabs(sim1), fmin(fabs(si), 0.5 * fabs(pi)));
        }
        for (size_t i = 0; i + 1 < n; i++) {
            const double hi = xa[i + 1] - xa[i], si = (ya[i + 1] - ya[i]) / hi;
            s->d[i] = (yp[i] + yp[i + 1] - 2.0 * si) / (hi * hi);
            s->c[i] = (3.0 * si - 2.0 * yp[i] - yp[i + 1]) / hi;
            s->b[i] = yp[i];
        }
        free(yp);
    }
    return GSL_SUCCESS;
}

/* interval index i with x[i] <= t < x[i+1]; t == x[n-1] belongs to the last interval (gsl_interp_bsearch) */
static inline size_t gb_shim_bsearch(const double *x, double t, size_t n) {
    size_t lo = 0, hi = n - 1;
    while (hi > lo + 1) { const size_t mid = (lo + hi) / 2; if (x[mid] > t) hi = mid; else lo = mid; }
    return lo;
}
static inline double gsl_spline_eval(const gsl_spline *s, double t, gsl_interp_accel *a) {
    (void)a;
    if (t < s->x[0] || t > s->x[s->size - 1]) return NAN;      /* GSL_EDOM: gsl_spline_eval returns NaN out of range */
    const size_t i = gb_shim_bsearch(s->x, t, s->size);
    const double dx = t - s->x[i];
    return s->y[i] + dx * (s->b[i] + dx * (s->c[i] + dx * s->d[i]));
}

static inline double gb_shim_no_spline(void) {
    fprintf(stderr, "oracle/gsl_shim: gsl_spline_eval_deriv / _deriv2 / _integ are not implemented (spline potentials are out of scope)\n"); abort(); return 0.;
}
static inline double gsl_spline_eval_deriv(const gsl_spline *s, double t, gsl_interp_accel *a) { (void)s; (void)t; (void)a; return gb_shim_no_spline(); }
static inline double gsl_spline_eval_deriv2(const gsl_spline *s, double t, gsl_interp_accel *a) { (void)s; (void)t; (void)a; return gb_shim_no_spline(); }
static inline double gsl_spline_eval_integ(const gsl_spline *s, double t0, double t1, gsl_interp_accel *a) { (void)s; (void)t0; (void)t1; (void)a; return gb_shim_no_spline(); }
#endif
