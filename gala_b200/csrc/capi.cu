// capi.cu -- the C ABI declared in include/gala_b200.h.
//
// Host side of the drop-in boundary: turns a gb_potential (the flat mirror of the reference's
// CPotential, potential/potential/src/cpotential.h:10-36) into the constant-bank DevPot, resolves
// the composite signature, stages HOST buffers through device memory when asked to, launches the
// kernels on the caller's stream and maps failures to the error codes of the header.
// No CPU fallback exists: without a CUDA device every compute entry point returns -10.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/gala_b200.h"
#include "kernels.h"

namespace {

thread_local std::string g_err;
std::atomic<long> g_launches{0};

int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
    return fail(-10, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__, #call); } while (0)
#define RET_IF(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

// expected parameter counts ([G, ...]) of each type for the compile-time signatures
int expected_npar(int type) {
    switch (type) {
        case GB_POT_NULL: return 1;
        case GB_POT_HERNQUIST: return 3;
        case GB_POT_NFW_SPHERICAL: case GB_POT_NFW_FLATTENED: case GB_POT_NFW_TRIAXIAL: return 6;
        case GB_POT_MIYAMOTONAGAI: return 4;
        case GB_POT_MN3: return 13;
        case GB_POT_LONGMURALIBAR: return 6;
        case GB_POT_KEPLER: return 2;
        case GB_POT_PLUMMER: case GB_POT_ISOCHRONE: case GB_POT_JAFFE: return 3;
        case GB_POT_STONE: case GB_POT_SATOH: case GB_POT_POWERLAWCUTOFF: return 4;
        case GB_POT_BURKERT: case GB_POT_KUZMIN: return 3;
        case GB_POT_LOGARITHMIC: return 7;
        case GB_POT_LEESUTO: return 6;
        default: return -1;
    }
}
int min_npar(int type) {
    switch (type) {
        case GB_POT_NULL: return 0;
        case GB_POT_MN3: return 10;
        case GB_POT_NFW_SPHERICAL: return 3;
        case GB_POT_SCF: return 5;
        case GB_POT_MULTIPOLE: return 6;
        case GB_POT_TIMEINTERP: return 7;
        default: return expected_npar(type);
    }
}

#define GB_SCF_NMAX_CONST 10
#define GB_SCF_LMAX_CONST 6
static_assert(GB_CEXT == 2 * (GB_SCF_NMAX_CONST + 1) * ((GB_SCF_LMAX_CONST + 1) * (GB_SCF_LMAX_CONST + 2) / 2), "cext layout");

void ext_cache_unpin(double* d);
struct Resolved {
    DevPot P;
    std::vector<double> ext;   // host copy of large parameter blocks (SCF coefficients)
    double* d_ext = nullptr;   // device copy (owned by the table cache below, pinned there while this object lives)
    Resolved() = default;
    Resolved(const Resolved&) = delete;
    Resolved& operator=(const Resolved&) = delete;
    ~Resolved() { if (d_ext) ext_cache_unpin(d_ext); }
};

// Device copies of the large parameter blocks, kept across calls: a cudaMalloc + cudaFree pair per call cost
// 7-30 ms inside a process that holds tens of GB of torch allocations (measured on BovyMWPotential2014:
// 40-66 ms per call around a 34 ms kernel).  Keyed on (device, contents); at most 16 unpinned entries, the
// oldest unpinned one is evicted.  An entry is pinned from resolve() until the call that resolved it has
// launched its kernels (Resolved's destructor), so a concurrent caller cannot free a table between another
// thread's lookup and its launch; after the launch cudaFree's implicit device synchronisation protects it.
struct ExtEntry { int dev; std::vector<double> host; double* d; int pins; };
std::mutex g_ext_mu;
std::vector<ExtEntry> g_ext_cache;
void ext_cache_clear() {            // gb_release_scratch
    std::lock_guard<std::mutex> g(g_ext_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    std::vector<ExtEntry> keep;
    for (auto& e : g_ext_cache) {
        if (e.pins > 0) { keep.push_back(std::move(e)); continue; }     // in use by a call in flight on another thread
        cudaSetDevice(e.dev); cudaFree(e.d);
    }
    g_ext_cache.swap(keep);
    cudaSetDevice(cur);
}
cudaError_t ext_cache_get(const std::vector<double>& ext, double** out) {
    std::vector<ExtEntry>& cache = g_ext_cache;
    std::lock_guard<std::mutex> g(g_ext_mu);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    for (size_t i = 0; i < cache.size(); i++)
        if (cache[i].dev == dev && cache[i].host.size() == ext.size() &&
            memcmp(cache[i].host.data(), ext.data(), ext.size() * sizeof(double)) == 0) {
            if (i + 1 != cache.size()) { ExtEntry t = std::move(cache[i]); cache.erase(cache.begin() + i); cache.push_back(std::move(t)); }
            cache.back().pins++;
            *out = cache.back().d;
            return cudaSuccess;
        }
    if (cache.size() >= 16)
        for (size_t i = 0; i < cache.size(); i++)
            if (cache[i].pins == 0) {
                cudaSetDevice(cache[i].dev); cudaFree(cache[i].d); cudaSetDevice(dev);
                cache.erase(cache.begin() + i);
                break;
            }
    double* d = nullptr;
    if ((e = cudaMalloc(&d, ext.size() * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d, ext.data(), ext.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) { cudaFree(d); return e; }
    cache.push_back(ExtEntry{dev, ext, d, 1});
    *out = d;
    return cudaSuccess;
}
void ext_cache_unpin(double* d) {
    std::lock_guard<std::mutex> g(g_ext_mu);
    for (auto& e : g_ext_cache) if (e.d == d) { if (e.pins > 0) e.pins--; return; }
}

// SCF coefficients for the device (scf.cuh): (S,T) pairs in [l][m<=l][n] order with the spherical-
// harmonic normalisation sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) (what gsl_sf_legendre_sphPlm and the
// gamma-function ratio of reference bfe_helper.cpp:21,42,72-74 apply per term) folded in.
// Input layout: params = [G, nmax, lmax, m, r_s, S[n][l][m]..., T[n][l][m]...] (bfe.cpp:229-258,
// index i = m + (lmax+1)(l + (lmax+1) n), bfe.cpp:160).
void scf_pack(const double* params, int nmax, int lmax, std::vector<double>& ext) {
    const int L1 = lmax + 1, ncoef = (nmax + 1) * L1 * L1;
    const double* S = params + 5;
    const double* T = params + 5 + ncoef;
    for (int l = 0; l <= lmax; l++)
        for (int m = 0; m <= l; m++) {
            double ratio = 1.;                        // (l-m)!/(l+m)!
            for (int k = l - m + 1; k <= l + m; k++) ratio /= (double)k;
            const double nlm = sqrt((2. * l + 1.) / (4. * M_PI) * ratio);
            for (int n = 0; n <= nmax; n++) {
                const int i = m + L1 * (l + L1 * n);
                ext.push_back(S[i] * nlm);
                ext.push_back(T[i] * nlm);
            }
        }
}

// Multipole coefficients for the device (multipole.cuh): (S,T) pairs in the reference's (l, m<=l)
// order with the spherical-harmonic normalisation folded in (gsl_sf_legendre_sphPlm and the factor A
// of multipole.cpp:106-111).  Input: params = [G, lmax, num_coeff, inner, m, r_s, S00, T00, S10, ...].
void mp_pack(const double* params, int lmax, std::vector<double>& ext) {
    int i = 0;
    for (int l = 0; l <= lmax; l++)
        for (int m = 0; m <= l; m++, i++) {
            double ratio = 1.;
            for (int k = l - m + 1; k <= l + m; k++) ratio /= (double)k;
            const double nlm = sqrt((2. * l + 1.) / (4. * M_PI) * ratio);
            ext.push_back(params[6 + 2 * i] * nlm);
            ext.push_back(params[7 + 2 * i] * nlm);
        }
}

// PowerLawCutoff, fast build: Chebyshev fit of F(s) = gamma*(a, s^2) = P(a, s^2) / s^(2a) (Tricomi's entire
// incomplete gamma function) on GB_PLC_NINT equal intervals of s in [0, GB_PLC_SMAX], degree GB_PLC_DEG.
// F(s) = e^-x sum_n x^n / Gamma(a+n+1), x = s^2: all terms positive, summed in long double.
void plc_fit(double a, std::vector<double>& ext);
// The fit costs 1.4 ms of long-double arithmetic (more on a loaded host) and depends on `a` only: the last few are
// kept, so repeated calls with the same bulge (every integrate_orbit of one potential) skip it.
void plc_pack(double a, std::vector<double>& ext) {
    static std::mutex mu;
    static std::vector<std::pair<double, std::vector<double>>> memo;
    std::lock_guard<std::mutex> g(mu);
    for (const auto& e : memo)
        if (e.first == a) { ext.insert(ext.end(), e.second.begin(), e.second.end()); return; }
    std::vector<double> tab;
    plc_fit(a, tab);
    if (memo.size() >= 8) memo.erase(memo.begin());
    memo.emplace_back(a, tab);
    ext.insert(ext.end(), tab.begin(), tab.end());
}
void plc_fit(double a, std::vector<double>& ext) {
    const int D = GB_PLC_DEG + 1;
    const long double h = (long double)GB_PLC_SMAX / GB_PLC_NINT, pi = 3.14159265358979323846264338327950288L;
    auto F = [&](long double sx) {
        const long double x = sx * sx;
        long double term = 1.0L / tgammal((long double)a + 1.0L), sum = term;
        for (int n = 0; n < 2000; n++) {
            term *= x / ((long double)a + n + 1.0L);
            sum += term;
            if (term < sum * 1e-22L) break;
        }
        return expl(-x) * sum;
    };
    for (int k = 0; k < GB_PLC_NINT; k++) {
        const long double mid = (k + 0.5L) * h;
        long double v[GB_PLC_DEG + 1];
        for (int j = 0; j < D; j++) v[j] = F(mid + 0.5L * h * cosl(pi * (2 * j + 1) / (2.0L * D)));
        for (int i = 0; i < D; i++) {
            long double c = 0;
            for (int j = 0; j < D; j++) c += v[j] * cosl(pi * i * (2 * j + 1) / (2.0L * D));
            ext.push_back((double)(2.0L * c / D));
        }
    }
}

// ---- TimeInterpolatedPotential (device side: csrc/timeinterp.cuh) -----------------------------
// One table of cubic pieces y = y_i + dx (b_i + dx (c_i + dx d_i)) per interpolated element, in the form
// gsl_spline_eval evaluates (time_interp.cpp:415-440 calls it per element).  The four interpolation types of
// time_interpolated.py:44-60, from their published definitions:
//   linear; cspline = natural cubic spline (second-derivative coefficients from the tridiagonal system, zero at
//   both ends); akima = Akima (1970), non-periodic end extension, slope m_i when both weights vanish; steffen =
//   Steffen (1990) monotone cubic with the end secants as end slopes.
int ti_spline(int method, const std::vector<double>& x, const std::vector<double>& y, double* tab /* [n-1][4] */) {
    const size_t n = x.size();
    static const size_t min_size[4] = {2, 3, 5, 3};                 // GSL's minimum number of points per type
    if (method < 0 || method > 3) return fail(-12, "TimeInterpolated: unknown interpolation method");
    if (n < min_size[method]) return fail(-12, "TimeInterpolated: too few time knots for this interpolation method");
    std::vector<double> b(n, 0.), c(n, 0.), d(n, 0.);
    if (method == 0) {
        for (size_t i = 0; i + 1 < n; i++) b[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    } else if (method == 1) {
        std::vector<double> cc(n, 0.);
        const size_t m = n - 2;
        std::vector<double> diag(m), off(m), g(m);
        for (size_t i = 0; i < m; i++) {
            const double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
            off[i] = h_ip1; diag[i] = 2.0 * (h_ip1 + h_i);
            g[i] = 3.0 * ((y[i + 2] - y[i + 1]) / h_ip1 - (y[i + 1] - y[i]) / h_i);
        }
        for (size_t i = 1; i < m; i++) {
            const double w = off[i - 1] / diag[i - 1];
            diag[i] -= w * off[i - 1]; g[i] -= w * g[i - 1];
        }
        cc[m] = g[m - 1] / diag[m - 1];
        for (size_t i = m - 1; i-- > 0;) cc[i + 1] = (g[i] - off[i] * cc[i + 2]) / diag[i];
        for (size_t i = 0; i + 1 < n; i++) {
            const double dx = x[i + 1] - x[i];
            b[i] = (y[i + 1] - y[i]) / dx - dx * (cc[i + 1] + 2.0 * cc[i]) / 3.0;
            c[i] = cc[i];
            d[i] = (cc[i + 1] - cc[i]) / (3.0 * dx);
        }
    } else if (method == 2) {
        std::vector<double> mm(n + 3);
        double* m = mm.data() + 2;
        for (size_t i = 0; i + 1 < n; i++) m[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
        m[-2] = 3.0 * m[0] - 2.0 * m[1];
        m[-1] = 2.0 * m[0] - m[1];
        m[n - 1] = 2.0 * m[n - 2] - m[n - 3];
        m[n] = 3.0 * m[n - 2] - 2.0 * m[n - 3];
        for (long i = 0; i + 1 < (long)n; i++) {
            const double NE = fabs(m[i + 1] - m[i]) + fabs(m[i - 1] - m[i - 2]);
            if (NE == 0.0) { b[i] = m[i]; continue; }
            const double h_i = x[i + 1] - x[i];
            const double NE_next = fabs(m[i + 2] - m[i + 1]) + fabs(m[i] - m[i - 1]);
            const double alpha_i = fabs(m[i - 1] - m[i - 2]) / NE;
            double tL = m[i];
            if (NE_next != 0.0) { const double al = fabs(m[i] - m[i - 1]) / NE_next; tL = (1.0 - al) * m[i] + al * m[i + 1]; }
            b[i] = (1.0 - alpha_i) * m[i - 1] + alpha_i * m[i];
            c[i] = (3.0 * m[i] - 2.0 * b[i] - tL) / h_i;
            d[i] = (b[i] + tL - 2.0 * m[i]) / (h_i * h_i);
        }
    } else {
        std::vector<double> yp(n);
        yp[0] = (y[1] - y[0]) / (x[1] - x[0]);
        yp[n - 1] = (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]);
        auto sgn = [](double v) { return v > 0 ? 1. : (v < 0 ? -1. : 0.); };
        for (size_t i = 1; i + 1 < n; i++) {
            const double hi = x[i + 1] - x[i], hm = x[i] - x[i - 1];
            const double si = (y[i + 1] - y[i]) / hi, sm = (y[i] - y[i - 1]) / hm;
            const double pi = (sm * hi + si * hm) / (hm + hi);
            yp[i] = (sgn(sm) + sgn(si)) * fmin(fabs(sm), fmin(fabs(si), 0.5 * fabs(pi)));
        }
        for (size_t i = 0; i + 1 < n; i++) {
            const double hi = x[i + 1] - x[i], si = (y[i + 1] - y[i]) / hi;
            d[i] = (yp[i] + yp[i + 1] - 2.0 * si) / (hi * hi);
            c[i] = (3.0 * si - 2.0 * yp[i] - yp[i + 1]) / hi;
            b[i] = yp[i];
        }
    }
    for (size_t i = 0; i + 1 < n; i++) { tab[4 * i] = y[i]; tab[4 * i + 1] = b[i]; tab[4 * i + 2] = c[i]; tab[4 * i + 3] = d[i]; }
    return 0;
}
// a constant element: every piece is (value, 0, 0, 0).  The reference decides "constant" the same way: all knot
// values within 1e-15 of the first (time_interp.cpp:196-207,340-351).
void ti_const(double v, size_t n, double* tab) {
    for (size_t i = 0; i + 1 < n; i++) { tab[4 * i] = v; tab[4 * i + 1] = 0.; tab[4 * i + 2] = 0.; tab[4 * i + 3] = 0.; }
}
// rotation_matrix_to_axis_angle (time_interp.cpp:447-498)
void ti_axis_angle(const double* M, double* axis, double* angle) {
    const double trace = M[0] + M[4] + M[8];
    *angle = acos((trace - 1.0) / 2.0);
    if (fabs(*angle) < 1e-15) { axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0; *angle = 0.0; }
    else if (fabs(*angle - M_PI) < 1e-15) {
        const double xx = (M[0] + 1.0) / 2.0, yy = (M[4] + 1.0) / 2.0, zz = (M[8] + 1.0) / 2.0;
        const double xy = M[1] / 2.0, xz = M[2] / 2.0, yz = M[5] / 2.0;
        if (xx > yy && xx > zz) { axis[0] = sqrt(xx); axis[1] = xy / axis[0]; axis[2] = xz / axis[0]; }
        else if (yy > zz) { axis[1] = sqrt(yy); axis[0] = xy / axis[1]; axis[2] = yz / axis[1]; }
        else { axis[2] = sqrt(zz); axis[0] = xz / axis[2]; axis[1] = yz / axis[2]; }
    } else {
        const double sa = sin(*angle);
        axis[0] = (M[7] - M[5]) / (2.0 * sa); axis[1] = (M[2] - M[6]) / (2.0 * sa); axis[2] = (M[3] - M[1]) / (2.0 * sa);
        const double norm = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
        if (norm > 1e-15) { axis[0] /= norm; axis[1] /= norm; axis[2] /= norm; }
    }
}
// params -> small parameters [G, wtype, method, n, nwp, rot_const, t_min, t_max] + ext block (timeinterp.cuh layout)
int ti_pack(const gb_component& c, double* small, std::vector<double>& ext) {
    const double* p = c.params;
    if (c.n_params < 7) return fail(-12, "TimeInterpolated: parameter vector too short");
    const int wtype = (int)p[1], method = (int)p[2], n = (int)p[3], nwp = (int)p[4], no = (int)p[5], nR = (int)p[6];
    if (wtype <= GB_POT_NULL || wtype >= GB_POT_TIMEINTERP || wtype == GB_POT_SCF || wtype == GB_POT_MULTIPOLE)
        return fail(-11, "TimeInterpolated: the wrapped potential must be one of the analytic builtin types");
    if (n < 2 || nwp < 0 || nwp + 1 > 16 || (no != 1 && no != n) || (nR != 1 && nR != n))
        return fail(-12, "TimeInterpolated: bad knot / parameter / origin / rotation counts");
    if (nwp + 1 < min_npar(wtype)) return fail(-12, "TimeInterpolated: too few parameters for the wrapped type");
    if (c.n_params != 7 + n + n * nwp + 3 * no + 9 * nR) return fail(-12, "TimeInterpolated: parameter vector length does not match its header");
    if (c.do_shift_rotate) return fail(-12, "TimeInterpolated: origin and R travel inside the parameter vector, not in q0 / R");
    const double* tk = p + 7;
    const double* wv = tk + n;
    const double* ov = wv + (size_t)n * nwp;
    const double* Rv = ov + (size_t)3 * no;
    std::vector<double> x(tk, tk + n), y(n);
    for (int i = 0; i + 1 < n; i++) if (!(x[i + 1] > x[i])) return fail(-12, "TimeInterpolated: time knots must be strictly increasing");
    const int nel = nwp + 1 + 3 + 4;
    const size_t base = ext.size();
    ext.insert(ext.end(), x.begin(), x.end());
    ext.resize(base + n + (size_t)nel * (n - 1) * 4 + 9, 0.);
    double* tab = ext.data() + base + n;
    auto element = [&](int el, const double* vals, int stride, int count) -> int {
        // vals[k * stride], k < count (count == 1: constant)
        bool constant = count == 1;
        if (!constant) { constant = true; for (int k = 1; k < count; k++) if (fabs(vals[(size_t)k * stride] - vals[0]) > 1e-15) { constant = false; break; } }
        double* tb = tab + (size_t)el * (n - 1) * 4;
        if (constant) { ti_const(vals[0], n, tb); return 0; }
        for (int k = 0; k < n; k++) y[k] = vals[(size_t)k * stride];
        return ti_spline(method, x, y, tb);
    };
    RET_IF(element(0, p, 0, 1));                                                    // G: always constant (cytimeinterp.pyx:196-201)
    for (int k = 0; k < nwp; k++) RET_IF(element(1 + k, wv + k, nwp, n));
    for (int k = 0; k < 3; k++) RET_IF(element(nwp + 1 + k, ov + k, 3, no));
    // rotation: constant matrix, or splines through the axis-angle components (time_interp.cpp:325-392)
    bool rconst = nR == 1;
    if (!rconst) { rconst = true; for (int i = 1; i < nR && rconst; i++) for (int j = 0; j < 9; j++) if (fabs(Rv[i * 9 + j] - Rv[j]) > 1e-15) { rconst = false; break; } }
    double* constR = tab + (size_t)nel * (n - 1) * 4;
    for (int j = 0; j < 9; j++) constR[j] = Rv[j];
    if (!rconst) {
        std::vector<double> aa((size_t)4 * n);
        for (int i = 0; i < n; i++) ti_axis_angle(Rv + 9 * i, &aa[4 * i], &aa[4 * i + 3]);
        for (int k = 0; k < 4; k++) RET_IF(element(nwp + 4 + k, aa.data() + k, 4, n));
    } else {
        for (int k = 0; k < 4; k++) ti_const(0., n, tab + (size_t)(nwp + 4 + k) * (n - 1) * 4);
    }
    small[0] = p[0]; small[1] = wtype; small[2] = method; small[3] = n; small[4] = nwp; small[5] = rconst ? 1. : 0.;
    small[6] = x[0]; small[7] = x[n - 1];
    return 0;
}

bool sig_matches(const gb_potential* pot, std::initializer_list<int> types) {
    if ((size_t)pot->n_components != types.size()) return false;
    int i = 0;
    for (int t : types) {
        const gb_component& c = pot->comp[i++];
        if (c.type_id != t || c.do_shift_rotate || c.n_params != expected_npar(t)) return false;
        // the compile-time MN3 evaluates sqrt(z^2+b^2) once for the three discs (potentials.cuh)
        if (t == GB_POT_MN3 && !(c.params[3] == c.params[6] && c.params[6] == c.params[9])) return false;
    }
    return true;
}

// Build the DevPot from the spec: the equivalent of CPotentialWrapper.init +
// CCompositePotentialWrapper.__init__ (cpotential.pyx:57-102, ccompositepotential.pyx:27-70).
int resolve(const gb_potential* pot, Resolved& r, cudaStream_t stream, bool allow_time_dep = false) {
    if (!pot || !pot->comp) return fail(-12, "null potential spec");
    if (pot->n_dim != 3) return fail(-11, "only n_dim = 3 potentials are supported");
    if (pot->n_components < 1 || pot->n_components > GB_MAXC)
        return fail(-11, "number of potential components must be in 1.." + std::to_string(GB_MAXC));
    DevPot& P = r.P;
    memset(&P, 0, sizeof(P));
    P.n = pot->n_components;
    int off = 0, doff = 0;
    for (int i = 0; i < P.n; i++) {
        const gb_component& c = pot->comp[i];
        if (c.type_id < 0 || c.type_id >= GB_POT_NTYPES) return fail(-11, "unknown potential type id");
        if (c.type_id == GB_POT_SCF) {
#if !GB_HAVE_SCF
            return fail(-11, "SCF potential is not available in this build");
#endif
        }
        if (c.n_params < min_npar(c.type_id) || (c.n_params > 0 && !c.params))
            return fail(-12, "component " + std::to_string(i) + ": too few parameters for its type");
        if (c.type_id == GB_POT_POWERLAWCUTOFF && !(c.params[2] < 3.))
            return fail(-12, "PowerLawCutoff: alpha must be < 3 (gsl_sf_gamma_inc_P needs a > 0)");
        DevComp& d = P.c[i];
        d.type = c.type_id;
        d.shift = c.do_shift_rotate ? 1 : 0;
        d.poff = off;
        d.eoff = (int)r.ext.size();
        int nsmall = c.n_params;
        if (c.type_id == GB_POT_SCF) {
            nsmall = 5;
            const int nmax = (int)c.params[1], lmax = (int)c.params[2];
            if (!(c.params[1] >= 0.) || !(c.params[2] >= 0.) || (double)nmax != c.params[1] || (double)lmax != c.params[2])
                return fail(-12, "SCF: nmax and lmax must be non-negative integers");
            const int ncoef = (nmax + 1) * (lmax + 1) * (lmax + 1);
            if (c.n_params < 5 + 2 * ncoef) return fail(-12, "SCF: parameter vector shorter than 5 + 2*(nmax+1)(lmax+1)^2");
            if (lmax > 15 || nmax > 63) return fail(-11, "SCF: lmax <= 15 and nmax <= 63 supported");
            scf_pack(c.params, nmax, lmax, r.ext);
        }
        if (c.type_id == GB_POT_MULTIPOLE) {
            nsmall = 6;
            const int lmax = (int)c.params[1], ncoef = (int)c.params[2];
            if (lmax < 0 || lmax > GB_MP_LMAX) return fail(-11, "Multipole: 0 <= lmax <= 15 supported");
            if (ncoef != (lmax + 1) * (lmax + 2) / 2 || c.n_params < 6 + 2 * ncoef)
                return fail(-12, "Multipole: num_coeff must be (lmax+1)(lmax+2)/2 and the parameter vector 6 + 2*num_coeff long");
            if (r.ext.size() & 1) r.ext.push_back(0.);     // keep (S,T) pairs 16-byte aligned
            d.eoff = (int)r.ext.size();
            mp_pack(c.params, lmax, r.ext);
        }
        if (c.type_id == GB_POT_POWERLAWCUTOFF && !getenv("GB_PLC_NO_TABLE")) {
            if (r.ext.size() & 1) r.ext.push_back(0.);
            d.eoff = (int)r.ext.size();
            plc_pack(0.5 * (3. - c.params[2]), r.ext);
        }
        double ti_small[8];
        if (c.type_id == GB_POT_TIMEINTERP) {
            nsmall = 8;
            // block = knots[n] | tables: the tables are read with 16-byte loads, so eoff + n must be even
            if (c.n_params < 7) return fail(-12, "TimeInterpolated: parameter vector too short");
            d.eoff = (int)r.ext.size();
            if (((size_t)d.eoff + (size_t)(int)c.params[3]) & 1) { r.ext.push_back(0.); d.eoff = (int)r.ext.size(); }
            RET_IF(ti_pack(c, ti_small, r.ext));
            P.time_dep = 1;
        }
        d.npar = nsmall;
        if (off + nsmall > GB_MAXP) return fail(-11, "too many potential parameters for the constant bank");
        for (int k = 0; k < nsmall; k++) P.par[off + k] = (c.type_id == GB_POT_TIMEINTERP) ? ti_small[k] : c.params[k];
        off += nsmall;
        d.doff = doff;
        if (doff + gb_nderived(c.type_id) > GB_MAXD) return fail(-11, "too many potential components for the constant bank");
        gb_derive(c.type_id, c.params, &P.drv[doff]);
        if (c.type_id == GB_POT_POWERLAWCUTOFF && !getenv("GB_PLC_NO_TABLE")) P.drv[doff + 5] = (double)d.eoff;
        doff += gb_nderived(c.type_id);
        for (int k = 0; k < 3; k++) d.q0[k] = c.q0[k];
        for (int k = 0; k < 9; k++) d.R[k] = c.R[k];
    }
    // signature resolution
    P.sig = SIG_GENERIC;
    if (sig_matches(pot, {GB_POT_NFW_SPHERICAL})) P.sig = SIG_NFW;
    else if (sig_matches(pot, {GB_POT_HERNQUIST})) P.sig = SIG_HERNQUIST;
    else if (sig_matches(pot, {GB_POT_MN3, GB_POT_HERNQUIST, GB_POT_HERNQUIST, GB_POT_NFW_SPHERICAL})) P.sig = SIG_MW2022;
    else if (sig_matches(pot, {GB_POT_LONGMURALIBAR, GB_POT_MN3, GB_POT_HERNQUIST, GB_POT_HERNQUIST, GB_POT_NFW_SPHERICAL})) P.sig = SIG_BAR_MW2022;
    else if (sig_matches(pot, {GB_POT_MN3, GB_POT_HERNQUIST, GB_POT_HERNQUIST, GB_POT_NFW_SPHERICAL, GB_POT_LONGMURALIBAR})) P.sig = SIG_MW2022_BAR;
    else if (sig_matches(pot, {GB_POT_MIYAMOTONAGAI, GB_POT_HERNQUIST, GB_POT_HERNQUIST, GB_POT_NFW_SPHERICAL})) P.sig = SIG_MW_V1;
    else if (sig_matches(pot, {GB_POT_MIYAMOTONAGAI, GB_POT_HERNQUIST, GB_POT_LOGARITHMIC})) P.sig = SIG_LM10;
    else if (sig_matches(pot, {GB_POT_MIYAMOTONAGAI, GB_POT_POWERLAWCUTOFF, GB_POT_NFW_SPHERICAL})) P.sig = SIG_BOVY2014;
    else if (pot->n_components == 1 && pot->comp[0].type_id == GB_POT_SCF && !pot->comp[0].do_shift_rotate) P.sig = SIG_SCF;
    if (getenv("GB_FORCE_GENERIC")) P.sig = SIG_GENERIC;
    if (P.sig == SIG_GENERIC && !getenv("GB_FORCE_GENERIC_HEAVY")) {
        bool heavy = false;
        for (int i = 0; i < P.n; i++) heavy |= (P.c[i].type == GB_POT_SCF || P.c[i].type == GB_POT_MULTIPOLE);
        if (!heavy) P.sig = P.time_dep ? SIG_GENERIC_TI : SIG_GENERIC_LIGHT;
    }
    if (P.time_dep && P.sig != SIG_GENERIC && P.sig != SIG_GENERIC_TI) P.sig = SIG_GENERIC_TI;   // GB_FORCE_GENERIC_HEAVY etc.
    if (P.sig == SIG_SCF && !getenv("GB_SCF_NO_CONST")) {
        // small expansions also go to the constant bank in the fixed (10,6) layout of scf_fast_gradient
        const double* p = pot->comp[0].params;
        const int nmax = (int)p[1], lmax = (int)p[2];
        if (nmax <= GB_SCF_NMAX_CONST && lmax <= GB_SCF_LMAX_CONST) {
            size_t k = 0;                                   // r.ext is [l][m][n] with this nmax
            for (int l = 0; l <= lmax; l++)
                for (int m = 0; m <= l; m++)
                    for (int n = 0; n <= nmax; n++, k += 2) {
                        const int i = 2 * (((l * (l + 1)) / 2 + m) * (GB_SCF_NMAX_CONST + 1) + n);
                        P.cext[i] = r.ext[k]; P.cext[i + 1] = r.ext[k + 1];
                    }
            P.cext_ok = 1;
        }
    }

    if (P.time_dep && !allow_time_dep)
        return fail(-11, "this entry point does not evaluate time-dependent (TimeInterpolated) potentials");
    if (!r.ext.empty()) {
        CU(ext_cache_get(r.ext, &r.d_ext));
        P.ext = r.d_ext;
    }
    return 0;
}

int resolve_frame(const gb_frame* fr, DevFrame& F) {
    memset(&F, 0, sizeof(F));
    if (!fr) { F.type = GB_FRAME_STATIC; return 0; }
    if (fr->type_id != GB_FRAME_STATIC && fr->type_id != GB_FRAME_ROTATING_3D)
        return fail(-11, "unsupported frame type");
    F.type = fr->type_id;
    for (int k = 0; k < 3; k++) F.om[k] = fr->omega[k];
    return 0;
}

// Device staging for GB_MEM_HOST calls.  A small grow-only cache per (device, slot) avoids paying
// cudaMalloc/cudaFree on every call of a time-stepping loop written on the Python side.  Each device has its
// own slots and its own mutex, so the per-device worker threads of a multi-device call never share a buffer.
struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
};
constexpr int NSLOT = 16;
constexpr int MAXDEV = 64;
std::mutex g_scratch_mu[MAXDEV];
Scratch g_scratch[MAXDEV][NSLOT];

cudaError_t scratch_get(int slot, size_t bytes, void** out) {
    int dev;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAXDEV) return cudaErrorInvalidDevice;
    Scratch& s = g_scratch[dev][slot];
    if (s.ptr && s.cap < bytes) {
        cudaFree(s.ptr);
        s.ptr = nullptr; s.cap = 0;
    }
    if (!s.ptr) {
        if (bytes == 0) bytes = 8;
        e = cudaMalloc(&s.ptr, bytes);
        if (e != cudaSuccess) { s.ptr = nullptr; return e; }
        s.cap = bytes;
    }
    *out = s.ptr;
    return cudaSuccess;
}

struct Ctx {
    cudaStream_t stream = nullptr;
    bool host = true;
    bool strict = false;
    int block = 0;
    int dev = 0;         // the device this call runs on
    int nsm = 148;
    int prev_dev = -1;
    std::unique_lock<std::mutex> lock;   // held while the call uses the device's scratch slots
    void lock_scratch() { if (!lock.owns_lock()) lock = std::unique_lock<std::mutex>(g_scratch_mu[dev]); }
    ~Ctx() { if (lock.owns_lock()) lock.unlock(); if (prev_dev >= 0) cudaSetDevice(prev_dev); }
};

int sm_count(int dev) {
    static std::mutex mu;
    static int tab[MAXDEV] = {0};
    std::lock_guard<std::mutex> g(mu);
    if (!tab[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        tab[dev] = n;
    }
    return tab[dev];
}

int open_ctx(const gb_launch* opt, Ctx& c) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(-10, std::string("no CUDA device available (this engine has no CPU fallback): ") +
                             (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    CU(cudaGetDevice(&c.dev));
    if (opt && opt->n_devices > 0)
        return fail(-12, "this entry point runs on one device (gb_launch.n_devices must be 0)");
    if (opt) {
        c.stream = (cudaStream_t)opt->stream;
        c.host = opt->mem == GB_MEM_HOST;
        c.strict = opt->strict_math != 0;
        c.block = opt->block_threads;
        if (opt->device >= 0 && c.dev != opt->device) {
            if (opt->device >= ndev) return fail(-12, "gb_launch.device is not a CUDA device ordinal of this process");
            c.prev_dev = c.dev; CU(cudaSetDevice(opt->device)); c.dev = opt->device;
        }
    }
    if (c.dev >= MAXDEV) return fail(-10, "device ordinal >= 64");
    c.nsm = sm_count(c.dev);
    if (c.host) c.lock_scratch();
    return 0;
}

// Fixed-step / evaluation kernels, thread = orbit: 128-thread CTAs unless that leaves SMs idle -- C1's 10^4
// orbits are 79 CTAs of 128 on 148 SMs -- in which case the CTA shrinks (64, then 32 threads) until there are
// at least two CTAs per SM or one warp per CTA.
int pick_block(const Ctx& c, size_t N) {
    if (c.block > 0) return c.block;
    int block = 128;
    while (block > 32 && (N + block - 1) / block < (size_t)2 * c.nsm) block >>= 1;
    return block;
}

// input staging: returns a device pointer for `p` (copying when the call is HOST-staged).  The 2-D forms
// move `rows` rows of n doubles between a host array whose rows are `pitch` doubles apart (one device's
// slice [lo, lo+n) of a (rows, N) array: p already points at column lo, pitch = N) and a dense (rows, n)
// device array.
int stage_in(Ctx& c, int slot, const void* p, size_t bytes, const void** dptr) {
    if (!c.host) { *dptr = p; return 0; }
    void* d; CU(scratch_get(slot, bytes, &d));
    if (bytes) CU(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, c.stream));
    *dptr = d;
    return 0;
}
int stage_in_2d(Ctx& c, int slot, const double* p, size_t rows, size_t n, size_t pitch, const void** dptr) {
    if (!c.host || pitch == n) return stage_in(c, slot, p, rows * n * sizeof(double), dptr);
    void* d; CU(scratch_get(slot, rows * n * sizeof(double), &d));
    if (rows && n)
        CU(cudaMemcpy2DAsync(d, n * sizeof(double), p, pitch * sizeof(double), n * sizeof(double), rows,
                             cudaMemcpyHostToDevice, c.stream));
    *dptr = d;
    return 0;
}
int stage_out_alloc(Ctx& c, int slot, void* p, size_t bytes, void** dptr) {
    if (!c.host) { *dptr = p; return 0; }
    CU(scratch_get(slot, bytes, dptr));
    return 0;
}
int stage_out_copy(Ctx& c, void* host, const void* dev, size_t bytes) {
    if (!c.host || !host || bytes == 0) return 0;
    CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
    return 0;
}
int stage_out_copy_2d(Ctx& c, double* host, const void* dev, size_t rows, size_t n, size_t pitch) {
    if (!c.host || !host || rows * n == 0) return 0;
    if (pitch == n) return stage_out_copy(c, host, dev, rows * n * sizeof(double));
    CU(cudaMemcpy2DAsync(host, pitch * sizeof(double), dev, n * sizeof(double), n * sizeof(double), rows,
                         cudaMemcpyDeviceToHost, c.stream));
    return 0;
}
int finish(Ctx& c) {
    // HOST-staged calls return results in caller memory, so they must complete before returning.
    // DEVICE calls stay asynchronous on the caller's stream (errors surface at the caller's sync).
    if (c.host) CU(cudaStreamSynchronize(c.stream));
    return 0;
}

// ---- several devices inside one HOST-mode call (gb_launch.n_devices) ------------------------
// The orbit index is cut into contiguous slices, one per device; one host thread per device runs the ordinary
// single-device path on its slice (that path already copies straight between the caller's arrays and device
// memory with 2-D copies, so nothing is gathered afterwards).  Pinned caller arrays make those copies
// asynchronous per device; with pageable arrays the driver stages them, still concurrently across threads.
// No collective, no peer access: orbits never interact.
bool multi_device(const gb_launch* opt) { return opt && opt->n_devices > 0; }

int check_devices(const gb_launch* opt) {
    if (opt->mem != GB_MEM_HOST) return fail(-12, "gb_launch.n_devices > 0 needs GB_MEM_HOST buffers (a device pointer belongs to one device)");
    if (!opt->devices) return fail(-12, "gb_launch.devices is null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(-10, std::string("no CUDA device available (this engine has no CPU fallback): ") +
                             (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (opt->n_devices > MAXDEV) return fail(-12, "at most 64 devices per call");
    for (int k = 0; k < opt->n_devices; k++) {
        if (opt->devices[k] < 0 || opt->devices[k] >= ndev) return fail(-12, "gb_launch.devices holds an ordinal that is not a CUDA device of this process");
        // GB_ALLOW_DUP_DEVICES (test hook): the slices of one device then run one after the other (they share the
        // device's scratch mutex), which lets a single-GPU box exercise the slicing / pitched-copy logic
        for (int j = 0; j < k; j++)
            if (opt->devices[j] == opt->devices[k] && !getenv("GB_ALLOW_DUP_DEVICES"))
                return fail(-12, "gb_launch.devices holds a device twice");
    }
    return 0;
}

// body(sub_opt, k, nd) -> rc, run on one thread per device; returns the most negative rc with its message.
template <class Body>
int run_on_devices(const gb_launch* opt, Body body) {
    int rc0 = check_devices(opt);
    if (rc0) return rc0;
    const int nd = opt->n_devices;
    std::vector<int> rc(nd, 0);
    std::vector<std::string> msg(nd);
    std::vector<std::thread> th;
    th.reserve(nd);
    for (int k = 0; k < nd; k++) {
        gb_launch sub = *opt;
        sub.n_devices = 0; sub.devices = nullptr; sub.device = opt->devices[k];
        sub.stream = nullptr;            // the caller's stream belongs to one device; each device uses its default stream
        th.emplace_back([&rc, &msg, &body, sub, k, nd]() {
            rc[k] = body(&sub, k, nd);
            if (rc[k]) msg[k] = g_err;   // g_err is thread-local: hand the text to the calling thread
        });
    }
    for (auto& t : th) t.join();
    int worst = 0, who = -1;
    for (int k = 0; k < nd; k++) if (rc[k] < worst) { worst = rc[k]; who = k; }
    if (who >= 0) return fail(worst, "device " + std::to_string(opt->devices[who]) + ": " + msg[who]);
    return 0;
}

// contiguous slice k of nd of range(N): sizes differ by at most one (the rule of gala_b200/dist.py:shard_bounds)
void slice_of(size_t N, int k, int nd, size_t* lo, size_t* n) {
    const size_t base = N / nd, rem = N % nd;
    *lo = (size_t)k * base + ((size_t)k < rem ? (size_t)k : rem);
    *n = base + ((size_t)k < rem ? 1 : 0);
}

#define KCALL(c, fn, ...) ((c).strict ? gbk_strict::fn(__VA_ARGS__) : gbk_fast::fn(__VA_ARGS__))

// Ruth4 coefficients exactly as the reference computes them (ruth4.pyx:65-78; numpy float pow ==
// C pow on the host).
void ruth4_coeffs(double* cs, double* ds) {
    const double two_13 = pow(2., 1. / 3.);
    cs[0] = 1. / (2. * (2. - two_13));
    cs[1] = (1. - two_13) / (2. * (2. - two_13));
    cs[2] = (1. - two_13) / (2. * (2. - two_13));
    cs[3] = 1. / (2. * (2. - two_13));
    ds[0] = 0.;
    ds[1] = 1. / (2. - two_13);
    ds[2] = -two_13 / (2. - two_13);
    ds[3] = 1. / (2. - two_13);
}

enum EvalKind { EV_GRAD, EV_ENERGY, EV_DENSITY };

int eval_impl(EvalKind kind, const gb_potential* pot, const double* q, double t, size_t N, size_t pitch, double* out,
              const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (N && (!q || !out)) return fail(-12, "null data pointer");
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    const int block = pick_block(c, N);
    const void* dq; RET_IF(stage_in_2d(c, 0, q, 3, N, pitch, &dq));
    const size_t orows = kind == EV_GRAD ? 3 : 1;
    void* dout; RET_IF(stage_out_alloc(c, 1, out, orows * N * sizeof(double), &dout));
    cudaError_t e;
    if (kind == EV_GRAD) e = KCALL(c, eval_gradient, r.P, (const double*)dq, t, N, (double*)dout, block, c.stream);
    else if (kind == EV_ENERGY) e = KCALL(c, eval_energy, r.P, (const double*)dq, t, N, (double*)dout, block, c.stream);
    else e = KCALL(c, eval_density, r.P, (const double*)dq, t, N, (double*)dout, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "evaluation kernel launch");
    if (N) g_launches++;
    RET_IF(stage_out_copy_2d(c, out, dout, orows, N, pitch));
    return finish(c);
}

int eval_common(EvalKind kind, const gb_potential* pot, const double* q, double t, size_t N, double* out,
                const gb_launch* opt) {
    if (!multi_device(opt)) return eval_impl(kind, pot, q, t, N, N, out, opt);
    if (N && (!q || !out)) return fail(-12, "null data pointer");
    return run_on_devices(opt, [=](const gb_launch* sub, int k, int nd) {
        size_t lo, n; slice_of(N, k, nd, &lo, &n);
        return n ? eval_impl(kind, pot, q + lo, t, n, N, out + lo, sub) : 0;
    });
}

}  // namespace

extern "C" {

const char* gb_last_error(void) { return g_err.c_str(); }
const char* gb_version(void) { return "gala_b200 0.2 (sm_100a)"; }
long gb_launch_count(void) { return g_launches.load(); }
int gb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int gb_gradient(const gb_potential* pot, const double* q, double t, size_t N, double* grad, const gb_launch* opt) {
    return eval_common(EV_GRAD, pot, q, t, N, grad, opt);
}
int gb_energy(const gb_potential* pot, const double* q, double t, size_t N, double* out, const gb_launch* opt) {
    return eval_common(EV_ENERGY, pot, q, t, N, out, opt);
}
int gb_density(const gb_potential* pot, const double* q, double t, size_t N, double* out, const gb_launch* opt) {
    return eval_common(EV_DENSITY, pot, q, t, N, out, opt);
}

int gb_hessian(const gb_potential* pot, const double* q, double t, size_t N, double* hess, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (N && (!q || !hess)) return fail(-12, "null data pointer");
    if (!pot) return fail(-12, "null potential");
    for (int i = 0; i < pot->n_components; i++) {
        const gb_component& cc = pot->comp[i];
        if (cc.type_id == GB_POT_SCF || cc.type_id == GB_POT_MULTIPOLE)
            return fail(-11, "no Hessian for basis-function expansions (the reference has none either)");
        if (cc.do_shift_rotate) {
            static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            for (int k = 0; k < 9; k++)
                if (cc.R[k] != I3[k])
                    return fail(-14, "Computing Hessian matrices for rotated potentials is currently not supported.");
        }
    }
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    const int block = pick_block(c, N);
    const void* dq; RET_IF(stage_in(c, 0, q, 3 * N * sizeof(double), &dq));
    void* dout; RET_IF(stage_out_alloc(c, 1, hess, 9 * N * sizeof(double), &dout));
    cudaError_t e = KCALL(c, eval_hessian, r.P, (const double*)dq, t, N, (double*)dout, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "hessian kernel launch");
    if (N) g_launches++;
    RET_IF(stage_out_copy(c, hess, dout, 9 * N * sizeof(double)));
    (void)t;
    return finish(c);
}

int gb_release_scratch(void) {
    int cur = 0, ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
    cudaGetDevice(&cur);
    for (int dev = 0; dev < ndev && dev < MAXDEV; dev++) {
        std::lock_guard<std::mutex> g(g_scratch_mu[dev]);
        bool any = false;
        for (int k = 0; k < NSLOT; k++) any |= g_scratch[dev][k].ptr != nullptr;
        if (!any) continue;
        cudaSetDevice(dev);
        cudaDeviceSynchronize();
        for (int k = 0; k < NSLOT; k++) {
            Scratch& sc = g_scratch[dev][k];
            if (sc.ptr) cudaFree(sc.ptr);
            sc.ptr = nullptr; sc.cap = 0;
        }
    }
    cudaSetDevice(cur);
    ext_cache_clear();
    return 0;
}

int gb_math_probe(int which, const double* x, size_t N, double* y, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (which < 0 || which > 3) return fail(-12, "gb_math_probe: which must be 0..3");
    if (N && (!x || !y)) return fail(-12, "null data pointer");
    const void* dx; RET_IF(stage_in(c, 0, x, N * sizeof(double), &dx));
    void* dy; RET_IF(stage_out_alloc(c, 1, y, N * sizeof(double), &dy));
    cudaError_t e = KCALL(c, math_probe, which, (const double*)dx, N, (double*)dy, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "math_probe launch");
    if (N) g_launches++;
    RET_IF(stage_out_copy(c, y, dy, N * sizeof(double)));
    return finish(c);
}

static int ham_eval_impl(bool grad, const gb_potential* pot, const gb_frame* fr, const double* w, double t, size_t N,
                         size_t pitch, double* out, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (N && (!w || !out)) return fail(-12, "null data pointer");
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    DevFrame F; RET_IF(resolve_frame(fr, F));
    const int block = pick_block(c, N);
    const size_t orows = grad ? 6 : 1;
    const void* dw; RET_IF(stage_in_2d(c, 0, w, 6, N, pitch, &dw));
    void* dout; RET_IF(stage_out_alloc(c, 1, out, orows * N * sizeof(double), &dout));
    cudaError_t e = grad ? KCALL(c, ham_gradient, r.P, F, (const double*)dw, t, N, (double*)dout, block, c.stream)
                         : KCALL(c, ham_energy, r.P, F, (const double*)dw, t, N, (double*)dout, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, grad ? "ham_gradient launch" : "ham_energy launch");
    if (N) g_launches++;
    RET_IF(stage_out_copy_2d(c, out, dout, orows, N, pitch));
    return finish(c);
}
static int ham_eval(bool grad, const gb_potential* pot, const gb_frame* fr, const double* w, double t, size_t N,
                    double* out, const gb_launch* opt) {
    if (!multi_device(opt)) return ham_eval_impl(grad, pot, fr, w, t, N, N, out, opt);
    if (N && (!w || !out)) return fail(-12, "null data pointer");
    return run_on_devices(opt, [=](const gb_launch* sub, int k, int nd) {
        size_t lo, n; slice_of(N, k, nd, &lo, &n);
        return n ? ham_eval_impl(grad, pot, fr, w + lo, t, n, N, out + lo, sub) : 0;
    });
}

int gb_hamiltonian_energy(const gb_potential* pot, const gb_frame* fr, const double* w, double t, size_t N,
                          double* out, const gb_launch* opt) {
    return ham_eval(false, pot, fr, w, t, N, out, opt);
}

int gb_hamiltonian_gradient(const gb_potential* pot, const gb_frame* fr, const double* w, double t, size_t N,
                            double* f, const gb_launch* opt) {
    return ham_eval(true, pot, fr, w, t, N, f, opt);
}

// Two side streams per device for the chunked HOST pipeline (created once, never destroyed).
// The streams and events are shared by every call on a device, so a call that uses them holds S->mu.
struct SideStreams { std::mutex mu; cudaStream_t s[2] = {nullptr, nullptr}; cudaEvent_t ev[3] = {nullptr, nullptr, nullptr}; };
static int side_streams(SideStreams** out) {
    static std::mutex mu;
    static SideStreams tab[MAXDEV];
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev >= MAXDEV) return fail(-10, "device ordinal >= 64");
    std::lock_guard<std::mutex> g(mu);
    SideStreams& S = tab[dev];
    if (!S.s[0]) {
        for (int k = 0; k < 2; k++) CU(cudaStreamCreateWithFlags(&S.s[k], cudaStreamNonBlocking));
        for (int k = 0; k < 3; k++) CU(cudaEventCreateWithFlags(&S.ev[k], cudaEventDisableTiming));
    }
    *out = &S;
    return 0;
}

// dt < 0 sentinel is not usable (backward integrations have dt < 0), so the kernels take a flag: dt_from_t != 0
// => the kernel itself forms dt = t[1] - t[0] from the device copy of the grid (DEVICE-mode calls: reading two
// doubles back to the host would cost a stream synchronisation and break the "DEVICE calls stay asynchronous"
// contract of the header; the subtraction is the same IEEE operation on either side).
static cudaError_t launch_fixed(Ctx& c, bool is_ruth4, const DevPot& P, const DevFrame& F, const double* dw0, size_t n,
                                const double* dt_dev, int ntimes, double dt, int dt_from_t, int save_all, double* dout,
                                const void* ti_tab, int block, cudaStream_t st) {
    if (!is_ruth4)
        return KCALL(c, leapfrog, P, dw0, n, dt_dev, ntimes, dt, dt_from_t, save_all, dout, ti_tab, block, st);
    double cs[4], ds[4];
    ruth4_coeffs(cs, ds);
    return KCALL(c, ruth4, P, F, dw0, n, dt_dev, ntimes, dt, dt_from_t, cs, ds, save_all, dout, ti_tab, block, st);
}

// Time-dependent composites in the fixed-step integrators: the state of every TimeInterpolated component at every
// time of the grid is tabulated once on the device (k_ti_table) -- every lane evaluates at the same t[j].  The table
// is a stream-ordered allocation (a few hundred KB), so DEVICE-mode calls stay asynchronous.
struct TiTable {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~TiTable() { if (p) cudaFreeAsync(p, s); }
};
static int make_ti_table(Ctx& c, const DevPot& P, const double* dt_dev, int ntimes, TiTable& tab) {
    if (!P.time_dep) return 0;
    const size_t nb = KCALL(c, ti_table_bytes, P, ntimes);
    tab.s = c.stream;
    {   // keep freed table memory in the device's default pool instead of returning it to the OS at every synchronise
        static std::mutex mu;
        static bool done[MAXDEV] = {false};
        std::lock_guard<std::mutex> g(mu);
        if (!done[c.dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, c.dev) == cudaSuccess) {
                uint64_t keep = (uint64_t)64 << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            done[c.dev] = true;
        }
    }
    CU(cudaMallocAsync(&tab.p, nb ? nb : 8, c.stream));
    cudaError_t e = KCALL(c, ti_table, P, dt_dev, ntimes, tab.p, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "TimeInterpolated state table launch");
    g_launches++;
    return 0;
}

// One device's share of a fixed-step call: orbits [0, N) of arrays whose rows are `pitch` doubles apart on the
// host side (pitch == N for a whole-array call; pitch = the full orbit count when w0 / w_out point at one
// device's slice of the caller's (6, N_total) / (6, ntimes, N_total) arrays).
static int fixed_step_impl(bool is_ruth4, const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N,
                           size_t pitch, const double* t, int ntimes, int save_all, double* w_out, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (ntimes < 2) return fail(-12, "the time grid needs at least 2 entries");
    if (!t || (N && (!w0 || !w_out))) return fail(-12, "null data pointer");
    DevFrame F; RET_IF(resolve_frame(fr, F));
    if (!is_ruth4 && F.type != GB_FRAME_STATIC)
        return fail(-13, "Leapfrog integration is currently only supported for StaticFrame");
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    // HOST: dt from the caller's grid; DEVICE: the kernel reads it from the device grid (no synchronisation here)
    const double dt = c.host ? t[1] - t[0] : 0.0;
    const int dt_from_t = c.host ? 0 : 1;
    const void* dtg; RET_IF(stage_in(c, 2, t, (size_t)ntimes * sizeof(double), &dtg));
    const double* dt_dev = (const double*)dtg;
    const size_t rows = save_all ? (size_t)ntimes : 1;       // output rows per phase-space component
    TiTable ti; RET_IF(make_ti_table(c, r.P, dt_dev, ntimes, ti));     // freed (stream-ordered) when the call returns

    // HOST buffers, many orbits: orbit-index chunks pipelined over two streams, so the H2D copy of
    // chunk k+1 and the D2H copy of chunk k-1 overlap the kernel of chunk k (the (6,N) layout makes a
    // chunk of orbits a 2-D copy: 6 [x ntimes] rows of nb doubles with a pitch of N doubles).
    const size_t kPipeMin = 1 << 16;
    if (c.host && N >= kPipeMin && !getenv("GB_NO_PIPELINE")) {
        SideStreams* S; RET_IF(side_streams(&S));
        std::lock_guard<std::mutex> side_lock(S->mu);
        size_t nb = (N + 7) / 8;                                   // ~8 chunks
        const size_t cap = ((size_t)256 << 20) / (rows * 6 * sizeof(double));   // <= 256 MB of output per chunk
        if (nb > cap) nb = cap;
        if (nb < 16384) nb = 16384;
        nb = (nb + 127) & ~(size_t)127;
        const int block = pick_block(c, nb);
        void *din[2], *dou[2];
        for (int k = 0; k < 2; k++) {
            CU(scratch_get(8 + k, 6 * nb * sizeof(double), &din[k]));
            CU(scratch_get(10 + k, rows * 6 * nb * sizeof(double), &dou[k]));
        }
        CU(cudaEventRecord(S->ev[2], c.stream));                   // t grid (and SCF coefficients) staged
        for (int k = 0; k < 2; k++) CU(cudaStreamWaitEvent(S->s[k], S->ev[2], 0));
        int k = 0;
        for (size_t a0 = 0; a0 < N; a0 += nb, k ^= 1) {
            const size_t n = (N - a0 < nb) ? N - a0 : nb;
            cudaStream_t st = S->s[k];
            CU(cudaMemcpy2DAsync(din[k], n * sizeof(double), w0 + a0, pitch * sizeof(double), n * sizeof(double), 6,
                                 cudaMemcpyHostToDevice, st));
            cudaError_t e = launch_fixed(c, is_ruth4, r.P, F, (const double*)din[k], n, dt_dev, ntimes, dt, dt_from_t,
                                         save_all, (double*)dou[k], ti.p, block, st);
            if (e != cudaSuccess) return cuda_fail(e, "integrator kernel launch");
            g_launches++;
            CU(cudaMemcpy2DAsync(w_out + a0, pitch * sizeof(double), dou[k], n * sizeof(double), n * sizeof(double),
                                 6 * rows, cudaMemcpyDeviceToHost, st));
        }
        for (int q = 0; q < 2; q++) {
            CU(cudaEventRecord(S->ev[q], S->s[q]));
            CU(cudaStreamWaitEvent(c.stream, S->ev[q], 0));
        }
        return finish(c);       // synchronises c.stream, which now depends on both side streams
    }

    const int block = pick_block(c, N);
    const void* dw0; RET_IF(stage_in_2d(c, 0, w0, 6, N, pitch, &dw0));
    void* dout; RET_IF(stage_out_alloc(c, 1, w_out, rows * 6 * N * sizeof(double), &dout));
    cudaError_t e = launch_fixed(c, is_ruth4, r.P, F, (const double*)dw0, N, dt_dev, ntimes, dt, dt_from_t, save_all,
                                 (double*)dout, ti.p, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "integrator kernel launch");
    if (N) g_launches++;
    RET_IF(stage_out_copy_2d(c, w_out, dout, 6 * rows, N, pitch));
    return finish(c);
}

static int fixed_step_common(bool is_ruth4, const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N,
                             const double* t, int ntimes, int save_all, double* w_out, const gb_launch* opt) {
    if (!multi_device(opt)) return fixed_step_impl(is_ruth4, pot, fr, w0, N, N, t, ntimes, save_all, w_out, opt);
    if (ntimes < 2) return fail(-12, "the time grid needs at least 2 entries");
    if (!t || (N && (!w0 || !w_out))) return fail(-12, "null data pointer");
    return run_on_devices(opt, [=](const gb_launch* sub, int k, int nd) {
        size_t lo, n; slice_of(N, k, nd, &lo, &n);
        // an empty slice still validates the potential / frame on its device (same error codes as one device)
        return fixed_step_impl(is_ruth4, pot, fr, w0 + lo, n, N, t, ntimes, save_all, w_out + lo, sub);
    });
}

int gb_leapfrog(const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N, const double* t,
                int ntimes, int save_all, double* w_out, const gb_launch* opt) {
    return fixed_step_common(false, pot, fr, w0, N, t, ntimes, save_all, w_out, opt);
}
int gb_ruth4(const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N, const double* t, int ntimes,
             int save_all, double* w_out, const gb_launch* opt) {
    return fixed_step_common(true, pot, fr, w0, N, t, ntimes, save_all, w_out, opt);
}

// ---- trajectory reductions (csrc/extrema.cuh) ------------------------------------------------
int gb_orbit_extrema(const gb_potential* pot, const gb_frame* fr, const double* w, const double* t, int ntimes,
                     size_t N, int with_energy, double* stats, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (ntimes < 1 || !t || (N && (!w || !stats))) return fail(-12, "null data pointer / empty time grid");
    DevFrame F; RET_IF(resolve_frame(fr, F));
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    const int block = pick_block(c, N);
    const void *dw, *dtg;
    RET_IF(stage_in(c, 0, w, 6 * (size_t)ntimes * N * sizeof(double), &dw));
    RET_IF(stage_in(c, 2, t, (size_t)ntimes * sizeof(double), &dtg));
    void* dst; RET_IF(stage_out_alloc(c, 1, stats, GB_EXT_NSTAT * N * sizeof(double), &dst));
    cudaError_t e = with_energy
        ? KCALL(c, trajectory_extrema_e1, r.P, F, (const double*)dw, (const double*)dtg, ntimes, N, (double*)dst, block, c.stream)
        : KCALL(c, trajectory_extrema_e0, r.P, F, (const double*)dw, (const double*)dtg, ntimes, N, (double*)dst, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "trajectory_extrema launch");
    if (N) g_launches++;
    RET_IF(stage_out_copy(c, stats, dst, GB_EXT_NSTAT * N * sizeof(double)));
    return finish(c);
}

int gb_orbit_extrema_list(const double* w, const double* t, int ntimes, size_t N, int kind, int kmax, double* vals,
                          double* times, int32_t* counts, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (ntimes < 1 || !t || (N && (!w || !counts))) return fail(-12, "null data pointer / empty time grid");
    if (kind < 0 || kind > 2) return fail(-12, "kind must be 0 (pericentres), 1 (apocentres) or 2 (z heights)");
    if (kmax < 0 || (kmax > 0 && N && (!vals || !times))) return fail(-12, "kmax > 0 needs the vals / times arrays");
    const int block = pick_block(c, N);
    const void *dw, *dtg;
    RET_IF(stage_in(c, 0, w, 6 * (size_t)ntimes * N * sizeof(double), &dw));
    RET_IF(stage_in(c, 2, t, (size_t)ntimes * sizeof(double), &dtg));
    void *dv, *dt2, *dc;
    const size_t lb = (size_t)kmax * N * sizeof(double);
    RET_IF(stage_out_alloc(c, 1, vals, lb, &dv));
    RET_IF(stage_out_alloc(c, 4, times, lb, &dt2));
    RET_IF(stage_out_alloc(c, 5, counts, N * sizeof(int32_t), &dc));
    cudaError_t e = KCALL(c, trajectory_extrema_list, (const double*)dw, (const double*)dtg, ntimes, N, kind, kmax,
                          (double*)dv, (double*)dt2, (int32_t*)dc, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "trajectory_extrema_list launch");
    if (N) g_launches++;
    RET_IF(stage_out_copy(c, vals, dv, lb));
    RET_IF(stage_out_copy(c, times, dt2, lb));
    RET_IF(stage_out_copy(c, counts, dc, N * sizeof(int32_t)));
    return finish(c);
}

static int integrate_extrema_impl(const gb_potential* pot, const gb_frame* fr, int scheme, const double* w0, size_t N,
                                  size_t pitch, const double* t, int ntimes, int with_energy, double* w_final,
                                  double* stats, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (scheme != 0 && scheme != 1) return fail(-12, "scheme must be 0 (leapfrog) or 1 (ruth4)");
    if (ntimes < 2) return fail(-12, "the time grid needs at least 2 entries");
    if (!t || (N && (!w0 || !stats))) return fail(-12, "null data pointer");
    DevFrame F; RET_IF(resolve_frame(fr, F));
    if (scheme == 0 && F.type != GB_FRAME_STATIC)
        return fail(-13, "Leapfrog integration is currently only supported for StaticFrame");
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    const double dt = c.host ? t[1] - t[0] : 0.0;
    const int block = pick_block(c, N);
    double cs[4], ds[4];
    ruth4_coeffs(cs, ds);
    const void *dw0, *dtg;
    RET_IF(stage_in_2d(c, 0, w0, 6, N, pitch, &dw0));
    RET_IF(stage_in(c, 2, t, (size_t)ntimes * sizeof(double), &dtg));
    void *dst, *dfin = nullptr;
    RET_IF(stage_out_alloc(c, 1, stats, GB_EXT_NSTAT * N * sizeof(double), &dst));
    if (w_final) RET_IF(stage_out_alloc(c, 4, w_final, 6 * N * sizeof(double), &dfin));
    cudaError_t e = with_energy
        ? KCALL(c, integrate_extrema_e1, r.P, F, scheme, cs, ds, (const double*)dw0, N, (const double*)dtg, ntimes, dt,
                c.host ? 0 : 1, (double*)dfin, (double*)dst, block, c.stream)
        : KCALL(c, integrate_extrema_e0, r.P, F, scheme, cs, ds, (const double*)dw0, N, (const double*)dtg, ntimes, dt,
                c.host ? 0 : 1, (double*)dfin, (double*)dst, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "integrate_extrema launch");
    if (N) g_launches++;
    RET_IF(stage_out_copy_2d(c, stats, dst, GB_EXT_NSTAT, N, pitch));
    if (w_final) RET_IF(stage_out_copy_2d(c, w_final, dfin, 6, N, pitch));
    return finish(c);
}

int gb_integrate_extrema(const gb_potential* pot, const gb_frame* fr, int scheme, const double* w0, size_t N,
                         const double* t, int ntimes, int with_energy, double* w_final, double* stats,
                         const gb_launch* opt) {
    if (!multi_device(opt))
        return integrate_extrema_impl(pot, fr, scheme, w0, N, N, t, ntimes, with_energy, w_final, stats, opt);
    if (ntimes < 2) return fail(-12, "the time grid needs at least 2 entries");
    if (!t || (N && (!w0 || !stats))) return fail(-12, "null data pointer");
    return run_on_devices(opt, [=](const gb_launch* sub, int k, int nd) {
        size_t lo, n; slice_of(N, k, nd, &lo, &n);
        return integrate_extrema_impl(pot, fr, scheme, w0 + lo, n, N, t, ntimes, with_energy,
                                      w_final ? w_final + lo : nullptr, stats + lo, sub);
    });
}

// dop853() front-end defaults (dopri/dop853.cpp:673-788)
static int dop853_defaults(Dop853Args& a, double atol, double rtol, long nmax, double dt_max, long nstiff,
                           double uround, double h0) {
    a.atol = atol; a.rtol = rtol;
    if (!nmax) nmax = 1000000;
    else if (nmax < 0) return fail(-1, "dop853: wrong input, nmax < 0");
    a.nmax = nmax;
    if (!nstiff) nstiff = 1000;
    else if (nstiff < 0) nstiff = nmax + 10;
    a.nstiff = nstiff;
    if (uround == 0.0) uround = 2.3E-16;
    else if (uround <= 1.0E-35 || uround >= 1.0) return fail(-1, "dop853: bad uround");
    a.uround = uround;
    a.hmax = dt_max;
    a.h0 = h0;
    return 0;
}

static int worst_status(Ctx& c, const int32_t* dstatus, size_t N, int32_t* host_status, int* worst) {
    // the per-orbit codes decide the return value, so they are always brought to the host
    std::vector<int32_t> tmp;
    int32_t* hs = host_status;
    if (!c.host || !hs) { tmp.resize(N); hs = tmp.data(); }
    if (!c.host || !host_status) {
        CU(cudaMemcpyAsync(hs, dstatus, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    }
    CU(cudaStreamSynchronize(c.stream));
    int w = 0;
    for (size_t i = 0; i < N; i++) if (hs[i] < w) w = hs[i];
    *worst = w;
    return 0;
}

// The default stream-ordered pool returns freed memory to the driver at the next synchronisation
// (release threshold 0), so a time-stepping loop of calls would re-map its scratch every call
// (measured: 1.5 s per call for 14.5 GB).  Keep up to 6 GB cached per device.
static int pool_keep() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    CU(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(mu);
    if (dev < 64 && !done[dev]) {
        cudaMemPool_t pool;
        CU(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long thr = 6ull << 30;
        CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        done[dev] = true;
    }
    return 0;
}

// Stream-ordered temporary device memory (freed with cudaFreeAsync on the same stream).
struct AsyncBuf {
    void* p_cached = nullptr;    // alternatively a pointer owned by the scratch cache (not freed here)
    void* get() const { return p ? p : p_cached; }
    void* p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) { s = stream; return cudaMallocAsync(&p, bytes ? bytes : 8, s); }
    ~AsyncBuf() { if (p) cudaFreeAsync(p, s); }
};

// One device's share of a gb_dop853 call; `pitch` as in fixed_step_impl.
static int dop853_impl(const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N, size_t pitch,
                       const double* t, int ntimes, double atol, double rtol, long nmax, double dt_max, long nstiff,
                       int save_all, double* w_out, int32_t* status, const gb_dop853_stats* stats, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (ntimes < 2) return fail(-12, "the time grid needs at least 2 entries");
    if (!t || (N && (!w0 || !w_out))) return fail(-12, "null data pointer");
    if (N >= 0xffffffffull) return fail(-12, "at most 2^32-2 orbits per call");
    RET_IF(pool_keep());
    DevFrame F; RET_IF(resolve_frame(fr, F));
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    const int block = c.block > 0 ? c.block : 128;     // 4 warps per CTA, step-synchronised (dop853.cuh)
    double two[2];
    if (c.host) { two[0] = t[0]; two[1] = t[1]; }
    else {
        CU(cudaMemcpyAsync(two, t, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        CU(cudaStreamSynchronize(c.stream));
    }
    Dop853Args a;
    // dop853_helper passes uround = np.finfo(float).eps and h = t[1]-t[0] (dop853.pyx:157-182)
    RET_IF(dop853_defaults(a, atol, rtol, nmax, dt_max, nstiff, 2.220446049250313e-16, two[1] - two[0]));
    const void* dw0; RET_IF(stage_in_2d(c, 0, w0, 6, N, pitch, &dw0));
    const void* dtg; RET_IF(stage_in(c, 2, t, (size_t)ntimes * sizeof(double), &dtg));
    const size_t orows = (save_all ? (size_t)ntimes : 1) * 6;
    void* dout; RET_IF(stage_out_alloc(c, 1, w_out, orows * N * sizeof(double), &dout));
    // status + optional stats
    void* dstat;
    if (c.host || !status) {
        c.lock_scratch();
        CU(scratch_get(3, N * sizeof(int32_t), &dstat));
    } else dstat = status;
    int32_t* dst[4] = {nullptr, nullptr, nullptr, nullptr};
    int32_t* hst[4] = {nullptr, nullptr, nullptr, nullptr};
    if (stats) { hst[0] = stats->nstep; hst[1] = stats->naccpt; hst[2] = stats->nrejct; hst[3] = stats->nfcn; }
    for (int k = 0; k < 4; k++) {
        if (!hst[k]) continue;
        if (c.host) { void* d; CU(scratch_get(4 + k, N * sizeof(int32_t), &d)); dst[k] = (int32_t*)d; }
        else dst[k] = hst[k];
    }

    // Orbit queue: chunks of contiguous orbit indices; inside a chunk the queue order is ascending
    // dynamical time (most steps first), see dop853.cuh.  Dense output needs an orbit-major scratch of
    // chunk x ntimes x 48 bytes, so the chunk is sized to a fraction of the free device memory.
    size_t chunk = N;
    if (save_all && N) {
        const size_t per_orbit = (size_t)ntimes * 6 * sizeof(double);
        c.lock_scratch();
        size_t free_b = 0, total_b = 0;
        // cudaMemGetInfo costs about a millisecond: only asked when the cached scratch cannot hold the whole call
        if (g_scratch[c.dev][14].cap < N * per_orbit || getenv("GB_D8_SCRATCH_MB")) CU(cudaMemGetInfo(&free_b, &total_b));
        else free_b = g_scratch[c.dev][14].cap;          // => budget >= cap: one chunk
        // As many orbits per persistent launch as memory allows: a launch cannot end before its longest
        // orbit does (~1000 sequential steps), so the work per resident lane must be several times
        // that, i.e. >> 38k orbits per launch.  Up to half of the free memory (counting what the
        // scratch cache already holds); the buffer stays cached until gb_release_scratch().
        c.lock_scratch();
        size_t budget = (free_b + g_scratch[c.dev][14].cap) / 2;
        if (const char* e = getenv("GB_D8_SCRATCH_MB")) budget = (size_t)atoll(e) << 20;
        chunk = budget / per_orbit;
        if (chunk < 64) chunk = 64;
        chunk &= ~(size_t)31;
        if (chunk > N) chunk = N;
    }
    const bool sorted = N >= 2048 && !getenv("GB_D8_NOSORT");
    // More than one chunk: chunks alternate between two side streams (each with its own queue
    // counter, sort buffers and scratch), so the tail of one persistent kernel -- a few long orbits
    // on a mostly idle GPU -- overlaps the start of the next chunk.
    // (measured: slower than one stream with a larger chunk -- each persistent grid fills the GPU, so the
    // second launch only becomes resident as the first one drains; kept as an opt-in experiment.)
    const int nstreams = (N > chunk && getenv("GB_D8_TWO_STREAMS")) ? 2 : 1;
    if (nstreams == 2 && save_all) {
        chunk = (chunk / 2) & ~(size_t)31;
        if (chunk < 64) chunk = 64;
    }
    std::unique_lock<std::mutex> side_lock;
    SideStreams* S = nullptr;
    if (nstreams == 2) { RET_IF(side_streams(&S)); side_lock = std::unique_lock<std::mutex>(S->mu); }
    AsyncBuf queue[2], keys_in[2], keys_out[2], idx_in[2], perm[2], temp[2], scratch[2];
    size_t temp_bytes = 0;
    cudaStream_t st[2] = {c.stream, c.stream};
    if (N) {
        if (sorted) CU(gb_sort_pairs_bytes(chunk, &temp_bytes));
        for (int k = 0; k < nstreams; k++) {
            CU(queue[k].alloc(sizeof(unsigned long long), c.stream));
            if (sorted) {
                CU(keys_in[k].alloc(chunk * 4, c.stream)); CU(keys_out[k].alloc(chunk * 4, c.stream));
                CU(idx_in[k].alloc(chunk * 4, c.stream)); CU(perm[k].alloc(chunk * 4, c.stream));
                CU(temp[k].alloc(temp_bytes, c.stream));
            }
            if (save_all) {
                const size_t sb = chunk * (size_t)ntimes * 6 * sizeof(double);
                if (nstreams == 1) CU(scratch_get(14, sb, &scratch[k].p_cached)); else CU(scratch[k].alloc(sb, c.stream));
            }
        }
        if (nstreams == 2) {
            CU(cudaEventRecord(S->ev[2], c.stream));       // inputs staged, buffers allocated
            for (int k = 0; k < 2; k++) { st[k] = S->s[k]; CU(cudaStreamWaitEvent(st[k], S->ev[2], 0)); }
        }
    }
    // GB_D8_TIMING=1 (diagnostic): CUDA-event time of the sort, the persistent kernel and the transpose of every
    // chunk, in place (ncu's per-kernel durations are serialised and cold-cache), printed to stderr
    const bool timing = getenv("GB_D8_TIMING") != nullptr;
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (timing) for (auto& e : tev) cudaEventCreate(&e);
    int k = 0;
    for (size_t orb0 = 0; orb0 < N; orb0 += chunk, k = (k + 1) % nstreams) {
        const size_t nc = (N - orb0 < chunk) ? N - orb0 : chunk;
        if (timing) cudaEventRecord(tev[0], st[k]);
        CU(cudaMemsetAsync(queue[k].p, 0, sizeof(unsigned long long), st[k]));
        if (sorted) {
            cudaError_t e = KCALL(c, dyn_time_keys, r.P, (const double*)dw0, N, two[0], orb0, nc, (float*)keys_in[k].p,
                                  (uint32_t*)idx_in[k].p, st[k]);
            if (e != cudaSuccess) return cuda_fail(e, "dyn_time_keys launch");
            g_launches++;
            CU(gb_sort_pairs((const float*)keys_in[k].p, (float*)keys_out[k].p, (const uint32_t*)idx_in[k].p,
                             (uint32_t*)perm[k].p, nc, temp[k].p, temp_bytes, st[k]));
        }
        double* kout = save_all ? (double*)scratch[k].get() : (double*)dout;
        const uint32_t* pp = sorted ? (const uint32_t*)perm[k].p : nullptr;
        if (timing) cudaEventRecord(tev[1], st[k]);
        cudaError_t e = (F.type == GB_FRAME_STATIC)
            ? KCALL(c, dop853_static, r.P, F, (const double*)dw0, N, (const double*)dtg, ntimes, a, save_all, pp,
                    (unsigned long long*)queue[k].p, orb0, nc, kout, (int32_t*)dstat, dst[0], dst[1], dst[2], dst[3],
                    block, st[k])
            : KCALL(c, dop853_rotating, r.P, F, (const double*)dw0, N, (const double*)dtg, ntimes, a, save_all, pp,
                    (unsigned long long*)queue[k].p, orb0, nc, kout, (int32_t*)dstat, dst[0], dst[1], dst[2], dst[3],
                    block, st[k]);
        if (e != cudaSuccess) return cuda_fail(e, "dop853 kernel launch");
        g_launches++;
        if (timing) cudaEventRecord(tev[2], st[k]);
        if (save_all) {
            e = KCALL(c, dop853_transpose, (const double*)scratch[k].get(), orb0, nc, ntimes, N, (double*)dout, st[k]);
            if (e != cudaSuccess) return cuda_fail(e, "dop853 transpose launch");
            g_launches++;
        }
        if (timing) {
            cudaEventRecord(tev[3], st[k]);
            cudaEventSynchronize(tev[3]);
            float a = 0, b = 0, c2 = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b, tev[1], tev[2]); cudaEventElapsedTime(&c2, tev[2], tev[3]);
            fprintf(stderr, "[gb_dop853 timing] chunk at %zu (%zu orbits): sort %.3f ms, kernel %.3f ms, transpose %.3f ms\n", orb0, nc, a, b, c2);
        }
    }
    if (timing) for (auto& e : tev) cudaEventDestroy(e);
    if (nstreams == 2) {
        for (int q = 0; q < 2; q++) {
            CU(cudaEventRecord(S->ev[q], S->s[q]));
            CU(cudaStreamWaitEvent(c.stream, S->ev[q], 0));
        }
    }
    RET_IF(stage_out_copy_2d(c, w_out, dout, orows, N, pitch));
    if (c.host) {
        if (status) RET_IF(stage_out_copy(c, status, dstat, N * sizeof(int32_t)));
        for (int k = 0; k < 4; k++) if (hst[k]) RET_IF(stage_out_copy(c, hst[k], dst[k], N * sizeof(int32_t)));
    }
    int worst = 0;
    if (N) RET_IF(worst_status(c, (const int32_t*)dstat, N, c.host ? status : nullptr, &worst));
    RET_IF(finish(c));
    if (worst < 0) return fail(worst, "Integration failed with code " + std::to_string(worst));
    return 0;
}

int gb_dop853(const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N, const double* t, int ntimes,
              double atol, double rtol, long nmax, double dt_max, long nstiff, int save_all, double* w_out,
              int32_t* status, const gb_dop853_stats* stats, const gb_launch* opt) {
    if (!multi_device(opt))
        return dop853_impl(pot, fr, w0, N, N, t, ntimes, atol, rtol, nmax, dt_max, nstiff, save_all, w_out, status, stats, opt);
    if (ntimes < 2) return fail(-12, "the time grid needs at least 2 entries");
    if (!t || (N && (!w0 || !w_out))) return fail(-12, "null data pointer");
    return run_on_devices(opt, [=](const gb_launch* sub, int k, int nd) {
        size_t lo, n; slice_of(N, k, nd, &lo, &n);
        gb_dop853_stats st = {nullptr, nullptr, nullptr, nullptr};
        if (stats) {
            st.nstep = stats->nstep ? stats->nstep + lo : nullptr; st.naccpt = stats->naccpt ? stats->naccpt + lo : nullptr;
            st.nrejct = stats->nrejct ? stats->nrejct + lo : nullptr; st.nfcn = stats->nfcn ? stats->nfcn + lo : nullptr;
        }
        return dop853_impl(pot, fr, w0 + lo, n, N, t, ntimes, atol, rtol, nmax, dt_max, nstiff, save_all, w_out + lo,
                           status ? status + lo : nullptr, stats ? &st : nullptr, sub);
    });
}

struct DevTmp {                      // stream-ordered device temporary with optional H2D fill
    void* p = nullptr; cudaStream_t s = nullptr;
    cudaError_t put(const void* host, size_t bytes, cudaStream_t st) {
        s = st;
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 8, s);
        if (e != cudaSuccess) { p = nullptr; return e; }
        if (host && bytes) e = cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, s);
        return e;
    }
    ~DevTmp() { if (p) cudaFreeAsync(p, s); }
};

int gb_stream_release(const gb_potential* pot, double G, const double* prog_w, const double* prog_t,
                      const double* prog_m, int ntimes, const int32_t* prog_idx, const double* sign,
                      const double* draws, int ncols, size_t Np, int df_kind, int flags, double* stream_w0,
                      const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (ntimes < 1 || !prog_w || !prog_t || !prog_m) return fail(-12, "null progenitor arrays");
    if (df_kind < 0 || df_kind > 3) return fail(-12, "unknown stream DF kind");
    static const int need[4] = {4, 0, 3, 6};
    if (ncols < need[df_kind]) return fail(-12, "too few random deviates per particle for this DF");
    if (Np && (!prog_idx || !sign || !stream_w0 || (need[df_kind] && !draws))) return fail(-12, "null data pointer");
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    const int block = pick_block(c, Np);
    const void *dpw, *dpt, *dpm, *dpi, *dsg, *dnr;
    RET_IF(stage_in(c, 0, prog_w, (size_t)ntimes * 6 * sizeof(double), &dpw));
    RET_IF(stage_in(c, 2, prog_t, (size_t)ntimes * sizeof(double), &dpt));
    RET_IF(stage_in(c, 3, prog_m, (size_t)ntimes * sizeof(double), &dpm));
    RET_IF(stage_in(c, 4, prog_idx, Np * sizeof(int32_t), &dpi));
    RET_IF(stage_in(c, 5, sign, Np * sizeof(double), &dsg));
    RET_IF(stage_in(c, 6, draws, draws ? Np * (size_t)ncols * sizeof(double) : 0, &dnr));
    void* dout; RET_IF(stage_out_alloc(c, 1, stream_w0, Np * 6 * sizeof(double), &dout));
    cudaError_t e = KCALL(c, fardal_release, r.P, G, (const double*)dpw, (const double*)dpt, (const double*)dpm, ntimes,
                          (const int32_t*)dpi, (const double*)dsg, (const double*)dnr, ncols, Np, df_kind, flags,
                          (double*)dout, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "stream release launch");
    if (Np) g_launches++;
    RET_IF(stage_out_copy(c, stream_w0, dout, Np * 6 * sizeof(double)));
    return finish(c);
}

int gb_fardal_release(const gb_potential* pot, double G, const double* prog_w, const double* prog_t,
                      const double* prog_m, int ntimes, const int32_t* prog_idx, const double* sign,
                      const double* normals, size_t Np, int gala_modified, double* stream_w0,
                      const gb_launch* opt) {
    return gb_stream_release(pot, G, prog_w, prog_t, prog_m, ntimes, prog_idx, sign, normals, 4, Np, 0, gala_modified,
                             stream_w0, opt);
}

// Mock-stream particles over several devices: rows are dealt in groups of GB_DEAL consecutive particles,
// group g to device g mod nd.  The caller's rows are ordered by release time (mockstream_generator.py:300-330),
// i.e. by remaining work, so a contiguous split would give the first device all the long integrations; dealing
// small groups round-robin gives every device the same mix (the C-ABI form of gala_b200/dist.py:deal_by_work).
// A device's share is a 2-D copy: its groups are GB_DEAL rows wide and nd * GB_DEAL rows apart.
#define GB_DEAL 128
struct Deal {
    int k = 0, nd = 1;
    size_t Np = 0;          // particles of the whole call
    size_t ngroups() const { return (Np + GB_DEAL - 1) / GB_DEAL; }
    size_t my_groups() const { const size_t g = ngroups(); return g > (size_t)k ? (g - k + nd - 1) / nd : 0; }
    bool has_tail() const { const size_t g = ngroups(); return my_groups() && (g - 1) % nd == (size_t)k && Np % GB_DEAL; }
    size_t count() const {                                   // particles of this device
        if (nd == 1) return Np;
        const size_t mg = my_groups();
        if (!mg) return 0;
        return has_tail() ? (mg - 1) * GB_DEAL + Np % GB_DEAL : mg * GB_DEAL;
    }
};
// copy this device's rows (elem bytes each) between the caller's array and a dense device array
static int deal_copy(Ctx& c, const Deal& D, bool to_device, void* host, void* dev, size_t elem) {
    const size_t n = D.count();
    if (!n || !host) return 0;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (D.nd == 1) {
        CU(cudaMemcpyAsync(to_device ? dev : host, to_device ? host : dev, n * elem, kind, c.stream));
        return 0;
    }
    const size_t w = GB_DEAL * elem, hp = (size_t)D.nd * w;
    char* h0 = (char*)host + (size_t)D.k * w;
    const size_t full = D.has_tail() ? D.my_groups() - 1 : D.my_groups();
    if (full) {
        if (to_device) CU(cudaMemcpy2DAsync(dev, w, h0, hp, w, full, kind, c.stream));
        else CU(cudaMemcpy2DAsync(h0, hp, dev, w, w, full, kind, c.stream));
    }
    if (D.has_tail()) {
        const size_t tb = (D.Np % GB_DEAL) * elem;
        char* ht = h0 + full * hp; char* dt_ = (char*)dev + full * w;
        CU(cudaMemcpyAsync(to_device ? (void*)dt_ : (void*)ht, to_device ? (void*)ht : (void*)dt_, tb, kind, c.stream));
    }
    return 0;
}

static int mock_dop853_impl(const gb_potential* pot, const gb_frame* fr, const double* stream_w0, const double* t1,
                            const Deal& D, double tfinal, double dt0, double atol, double rtol, long nmax,
                            double* stream_w, int32_t* status, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    const size_t Np = D.count();
    if (D.Np && (!stream_w0 || !t1 || !stream_w)) return fail(-12, "null data pointer");
    DevFrame F; RET_IF(resolve_frame(fr, F));
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    const int block = c.block > 0 ? c.block : 64;
    Dop853Args a;
    // dop853_step (dop853.pyx:45-69): uround 0 -> 2.3e-16, hmax 0, nstiff hard-coded to 1
    RET_IF(dop853_defaults(a, atol, rtol, nmax, 0.0, 1, 0.0, dt0));
    const void *dw0 = stream_w0, *dt1 = t1;
    void *dout = stream_w, *dstat = status;
    if (c.host) {
        void* p;
        CU(scratch_get(0, Np * 6 * sizeof(double), &p)); dw0 = p;
        RET_IF(deal_copy(c, D, true, (void*)stream_w0, p, 6 * sizeof(double)));
        CU(scratch_get(2, Np * sizeof(double), &p)); dt1 = p;
        RET_IF(deal_copy(c, D, true, (void*)t1, p, sizeof(double)));
        CU(scratch_get(1, Np * 6 * sizeof(double), &dout));
    }
    if (c.host || !status) { c.lock_scratch(); CU(scratch_get(3, Np * sizeof(int32_t), &dstat)); }
    cudaError_t e = KCALL(c, mock_dop853, r.P, F, (const double*)dw0, (const double*)dt1, Np, tfinal, a,
                          (double*)dout, (int32_t*)dstat, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "mock_dop853 launch");
    if (Np) g_launches++;
    if (c.host) {
        RET_IF(deal_copy(c, D, false, stream_w, dout, 6 * sizeof(double)));
        RET_IF(deal_copy(c, D, false, status, dstat, sizeof(int32_t)));
    }
    int worst = 0;
    if (Np) RET_IF(worst_status(c, (const int32_t*)dstat, Np, nullptr, &worst));
    RET_IF(finish(c));
    if (worst < 0) return fail(worst, "Integration failed with code " + std::to_string(worst));
    return 0;
}

int gb_mockstream_dop853(const gb_potential* pot, const gb_frame* fr, const double* stream_w0, const double* t1,
                         size_t Np, double tfinal, double dt0, double atol, double rtol, long nmax,
                         double* stream_w, int32_t* status, const gb_launch* opt) {
    Deal D; D.Np = Np;
    if (!multi_device(opt))
        return mock_dop853_impl(pot, fr, stream_w0, t1, D, tfinal, dt0, atol, rtol, nmax, stream_w, status, opt);
    return run_on_devices(opt, [=](const gb_launch* sub, int k, int nd) {
        Deal Dk = D; Dk.k = k; Dk.nd = nd;
        return mock_dop853_impl(pot, fr, stream_w0, t1, Dk, tfinal, dt0, atol, rtol, nmax, stream_w, status, sub);
    });
}

int gb_mockstream_dop853_animate(const gb_potential* pot, const gb_frame* fr, const double* w0_rows,
                                 const int32_t* release_idx, size_t Np, const double* t, int ntimes, double atol,
                                 double rtol, long nmax, int output_every, double* snapshots, double* final_w,
                                 int32_t* status, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (!c.host) return fail(-12, "gb_mockstream_dop853_animate takes host buffers");
    if (ntimes < 2 || !t) return fail(-12, "the time grid needs at least 2 entries");
    if (output_every < 1) return fail(-12, "output_every must be >= 1");
    if (Np && (!w0_rows || !release_idx || !snapshots || !final_w)) return fail(-12, "null data pointer");
    RET_IF(pool_keep());
    DevFrame F; RET_IF(resolve_frame(fr, F));
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    int nout = (ntimes - 1) / output_every + 1;
    if ((ntimes - 1) % output_every != 0) nout += 1;                  // mockstream.pyx:360-363
    Dop853Args a;
    RET_IF(dop853_defaults(a, atol, rtol, nmax, 0.0, 1, 0.0, t[1] - t[0]));   // dop853_step's settings
    DevTmp dw0, dri, dtg, dsn, dou, dst;
    CU(dw0.put(w0_rows, Np * 6 * sizeof(double), c.stream));
    CU(dri.put(release_idx, Np * sizeof(int32_t), c.stream));
    CU(dtg.put(t, (size_t)ntimes * sizeof(double), c.stream));
    const size_t sb = (size_t)nout * Np * 6 * sizeof(double);
    CU(dsn.put(nullptr, sb, c.stream));
    CU(dou.put(nullptr, Np * 6 * sizeof(double), c.stream));
    CU(dst.put(nullptr, (Np ? Np : 1) * sizeof(int32_t), c.stream));
    cudaError_t e = KCALL(c, mock_dop853_animate, r.P, F, (const double*)dw0.p, (const int32_t*)dri.p, Np,
                          (const double*)dtg.p, ntimes, a, output_every, (double*)dsn.p, (double*)dou.p,
                          (int32_t*)dst.p, c.block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "mock_dop853_animate launch");
    if (Np) g_launches++;
    std::vector<int32_t> hs(Np);
    if (Np) {
        CU(cudaMemcpyAsync(snapshots, dsn.p, sb, cudaMemcpyDeviceToHost, c.stream));
        CU(cudaMemcpyAsync(final_w, dou.p, Np * 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        CU(cudaMemcpyAsync(hs.data(), dst.p, Np * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    }
    CU(cudaStreamSynchronize(c.stream));
    int worst = 0;
    for (size_t i = 0; i < Np; i++) { if (hs[i] < worst) worst = hs[i]; if (status) status[i] = hs[i]; }
    if (worst < 0) return fail(worst, "Integration failed with code " + std::to_string(worst));
    return 0;
}

static int mock_leapfrog_impl(const gb_potential* pot, const double* stream_w0, const double* t1, const Deal& D,
                              double tfinal, double dt, double* stream_w, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    const size_t Np = D.count();
    if (D.Np && (!stream_w0 || !t1 || !stream_w)) return fail(-12, "null data pointer");
    if (dt == 0.0) return fail(-12, "dt must be non-zero");
    Resolved r; RET_IF(resolve(pot, r, c.stream, true));
    const int block = pick_block(c, Np);
    const void *dw0 = stream_w0, *dt1 = t1;
    void* dout = stream_w;
    if (c.host) {
        void* p;
        CU(scratch_get(0, Np * 6 * sizeof(double), &p)); dw0 = p;
        RET_IF(deal_copy(c, D, true, (void*)stream_w0, p, 6 * sizeof(double)));
        CU(scratch_get(2, Np * sizeof(double), &p)); dt1 = p;
        RET_IF(deal_copy(c, D, true, (void*)t1, p, sizeof(double)));
        CU(scratch_get(1, Np * 6 * sizeof(double), &dout));
    }
    cudaError_t e = KCALL(c, mock_leapfrog, r.P, (const double*)dw0, (const double*)dt1, Np, tfinal, dt,
                          (double*)dout, block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "mock_leapfrog launch");
    if (Np) g_launches++;
    if (c.host) RET_IF(deal_copy(c, D, false, stream_w, dout, 6 * sizeof(double)));
    return finish(c);
}

int gb_mockstream_leapfrog(const gb_potential* pot, const double* stream_w0, const double* t1, size_t Np,
                           double tfinal, double dt, double* stream_w, const gb_launch* opt) {
    Deal D; D.Np = Np;
    if (!multi_device(opt)) return mock_leapfrog_impl(pot, stream_w0, t1, D, tfinal, dt, stream_w, opt);
    return run_on_devices(opt, [=](const gb_launch* sub, int k, int nd) {
        Deal Dk = D; Dk.k = k; Dk.nd = nd;
        return mock_leapfrog_impl(pot, stream_w0, t1, Dk, tfinal, dt, stream_w, sub);
    });
}

// ---- massive bodies ------------------------------------------------------------------
// Flattens the per-body potentials into DevBodies (gb_device.cuh).  A body is a force source unless its
// potential is all-Null (CPotential::null, set by CPotentialWrapper.init when every component is a
// NullWrapper; cpotential.cpp:397-399 skips those).
static int resolve_bodies(const gb_bodies* bodies, DevBodies& B) {
    memset(&B, 0, sizeof(B));
    if (!bodies || bodies->n_bodies < 1 || bodies->n_bodies > GB_MAXB)
        return fail(-11, "the N-body kernels carry 1.." + std::to_string(GB_MAXB) + " bodies per system");
    B.nb = bodies->n_bodies;
    int nc = 0, off = 0;
    for (int b = 0; b < B.nb; b++) {
        const gb_potential* bp = bodies->body_pot ? &bodies->body_pot[b] : nullptr;
        B.cbeg[b] = nc;
        int massive = 0;
        if (bp) for (int i = 0; i < bp->n_components; i++) {
            const gb_component& c = bp->comp[i];
            if (c.type_id == GB_POT_NULL) continue;
            if (c.type_id < 0 || c.type_id >= GB_POT_NTYPES || c.type_id == GB_POT_SCF || c.type_id == GB_POT_MULTIPOLE)
                return fail(-11, "body potentials must be analytic builtin types");
            if (c.n_params < min_npar(c.type_id) || !c.params) return fail(-12, "body potential: too few parameters");
            if (nc >= GB_MAXBC || off + c.n_params > GB_MAXBP) return fail(-11, "too many body-potential components/parameters");
            B.type[nc] = c.type_id; B.poff[nc] = off;
            for (int k = 0; k < c.n_params; k++) B.par[off + k] = c.params[k];
            off += c.n_params;
            for (int k = 0; k < 9; k++) B.R[nc][k] = c.R[k];
            nc++; massive = 1;
        }
        B.null_[b] = !massive;
    }
    B.cbeg[B.nb] = nc; B.nc = nc;
    return 0;
}


int gb_nbody_leapfrog(const gb_potential* pot, const gb_bodies* bodies, const double* body_w0, int ngroups,
                      const int32_t* group, const double* w0_rows, const double* t1, size_t Np, double t0,
                      double tfinal, int nsteps, double dt, int scheme, double* out_particles, double* out_bodies,
                      size_t body_writer, double* traj, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (!c.host) return fail(-12, "gb_nbody_leapfrog takes host buffers");
    if (scheme != 0 && scheme != 1) return fail(-12, "scheme must be 0 (leapfrog) or 1 (ruth4)");
    double cs[4], ds[4];
    ruth4_coeffs(cs, ds);
    if (!body_w0 || ngroups < 1 || (Np && !w0_rows)) return fail(-12, "null data pointer");
    if (dt == 0.0) return fail(-12, "dt must be non-zero");
    RET_IF(pool_keep());
    Resolved r; RET_IF(resolve(pot, r, c.stream));
    DevBodies B; RET_IF(resolve_bodies(bodies, B));
    const size_t nb = B.nb, ntot = nb + Np;
    if (body_writer >= (Np ? Np : 1)) return fail(-12, "body_writer out of range");
    DevTmp dbw, dgrp, dw0, dt1, dop, dob, dtr;
    CU(dbw.put(body_w0, (size_t)ngroups * nb * 6 * sizeof(double), c.stream));
    if (group) CU(dgrp.put(group, Np * sizeof(int32_t), c.stream));
    CU(dw0.put(w0_rows, Np * 6 * sizeof(double), c.stream));
    if (t1) CU(dt1.put(t1, Np * sizeof(double), c.stream));
    if (out_particles) CU(dop.put(nullptr, Np * 6 * sizeof(double), c.stream));
    if (out_bodies) CU(dob.put(nullptr, nb * 6 * sizeof(double), c.stream));
    const size_t tb = traj ? (size_t)(nsteps + 1) * ntot * 6 * sizeof(double) : 0;
    if (traj) { if (t1) return fail(-12, "trajectories need a common start time"); CU(dtr.put(nullptr, tb, c.stream)); }
    cudaError_t e = KCALL(c, nbody_leapfrog, r.P, B, scheme, cs, ds, (const double*)dbw.p, (const int32_t*)dgrp.p, (const double*)dw0.p,
                          (const double*)dt1.p, Np, t0, tfinal, nsteps, dt, (double*)dop.p, (double*)dob.p, body_writer,
                          (double*)dtr.p, ntot, c.block, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "nbody_leapfrog launch");
    g_launches++;
    if (out_particles && Np) CU(cudaMemcpyAsync(out_particles, dop.p, Np * 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    if (out_bodies) CU(cudaMemcpyAsync(out_bodies, dob.p, nb * 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    if (traj) CU(cudaMemcpyAsync(traj, dtr.p, tb, cudaMemcpyDeviceToHost, c.stream));
    return finish(c);
}

int gb_nbody_dop853(const gb_potential* pot, const gb_bodies* bodies, const double* body_w0, int ngroups,
                    const int32_t* group, const double* w0_rows, const double* t1, size_t Np,
                    const double* tgrid, int ntimes, double tfinal, double dt0, double atol, double rtol, long nmax,
                    double dt_max, int step_mode, double* out_particles, double* out_bodies, size_t body_writer,
                    double* traj, int32_t* status, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (!c.host) return fail(-12, "gb_nbody_dop853 takes host buffers");
    if (!body_w0 || ngroups < 1 || (Np && !w0_rows)) return fail(-12, "null data pointer");
    if (traj && (!tgrid || ntimes < 2 || t1)) return fail(-12, "dense output needs a time grid and a common start time");
    RET_IF(pool_keep());
    Resolved r; RET_IF(resolve(pot, r, c.stream));
    DevBodies B; RET_IF(resolve_bodies(bodies, B));
    const size_t nb = B.nb, ntot = nb + Np;
    if (body_writer >= (Np ? Np : 1)) return fail(-12, "body_writer out of range");
    Dop853Args a;
    // step_mode 0: dop853_helper (dop853.pyx:157-182: uround = eps, nstiff = -1 from direct_nbody_dop853, nbody.pyx:106);
    // step_mode 1: dop853_step (dop853.pyx:45-69: uround 0 -> 2.3e-16, hmax 0, nstiff hard-coded to 1)
    if (step_mode == 0) RET_IF(dop853_defaults(a, atol, rtol, nmax, dt_max, -1, 2.220446049250313e-16, dt0));
    else RET_IF(dop853_defaults(a, atol, rtol, nmax, 0.0, 1, 0.0, dt0));
    const size_t nthreads = Np ? Np : 1;
    DevTmp dbw, dgrp, dw0, dt1, dtg, dop, dob, dtr, dst;
    CU(dbw.put(body_w0, (size_t)ngroups * nb * 6 * sizeof(double), c.stream));
    if (group) CU(dgrp.put(group, Np * sizeof(int32_t), c.stream));
    CU(dw0.put(w0_rows, Np * 6 * sizeof(double), c.stream));
    if (t1) CU(dt1.put(t1, Np * sizeof(double), c.stream));
    if (tgrid) CU(dtg.put(tgrid, (size_t)ntimes * sizeof(double), c.stream));
    if (out_particles) CU(dop.put(nullptr, Np * 6 * sizeof(double), c.stream));
    if (out_bodies) CU(dob.put(nullptr, nb * 6 * sizeof(double), c.stream));
    const size_t tb = traj ? (size_t)ntimes * ntot * 6 * sizeof(double) : 0;
    if (traj) { CU(dtr.put(nullptr, tb, c.stream)); CU(cudaMemsetAsync(dtr.p, 0xff, tb, c.stream)); }   // NaN where an orbit failed
    CU(dst.put(nullptr, nthreads * sizeof(int32_t), c.stream));
    const double t0 = tgrid ? tgrid[0] : 0.0;
    cudaError_t e = KCALL(c, nbody_dop853, r.P, B, a, (const double*)dbw.p, (const int32_t*)dgrp.p, (const double*)dw0.p,
                          (const double*)dt1.p, Np, (const double*)dtg.p, ntimes, t0, tfinal, (double*)dop.p, (double*)dob.p,
                          body_writer, (double*)dtr.p, ntot, (int32_t*)dst.p, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "nbody_dop853 launch");
    g_launches++;
    if (out_particles && Np) CU(cudaMemcpyAsync(out_particles, dop.p, Np * 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    if (out_bodies) CU(cudaMemcpyAsync(out_bodies, dob.p, nb * 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    if (traj) CU(cudaMemcpyAsync(traj, dtr.p, tb, cudaMemcpyDeviceToHost, c.stream));
    std::vector<int32_t> hs(nthreads);
    CU(cudaMemcpyAsync(hs.data(), dst.p, nthreads * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    int worst = 0;
    for (size_t i = 0; i < nthreads; i++) { if (hs[i] < worst) worst = hs[i]; if (status) status[i] = hs[i]; }
    if (worst < 0) return fail(worst, "Integration failed with code " + std::to_string(worst));
    return 0;
}

int gb_nbody_dop853_animate(const gb_potential* pot, const gb_bodies* bodies, const double* body_w0,
                            const double* w0_rows, const int32_t* release_idx, size_t Np, const double* t, int ntimes,
                            double atol, double rtol, long nmax, int output_every, double* snapshots,
                            double* out_particles, double* out_bodies, int32_t* status, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (!c.host) return fail(-12, "gb_nbody_dop853_animate takes host buffers");
    if (ntimes < 2 || !t || !body_w0 || !snapshots) return fail(-12, "null data pointer / short time grid");
    if (output_every < 1) return fail(-12, "output_every must be >= 1");
    if (Np && (!w0_rows || !release_idx)) return fail(-12, "null data pointer");
    RET_IF(pool_keep());
    Resolved r; RET_IF(resolve(pot, r, c.stream));
    DevBodies B; RET_IF(resolve_bodies(bodies, B));
    const size_t nb = B.nb, ntot = nb + Np;
    int nout = (ntimes - 1) / output_every + 1;
    if ((ntimes - 1) % output_every != 0) nout += 1;
    Dop853Args a;
    RET_IF(dop853_defaults(a, atol, rtol, nmax, 0.0, 1, 0.0, t[1] - t[0]));      // dop853_step's settings
    // the lane of the last particle writes the bodies' end state (the reference's w[:nbodies] after the loop)
    DevTmp dba, dw0, dri, dtg, dsn, dop, dob, dst;
    CU(dba.put(nullptr, (size_t)ntimes * nb * 6 * sizeof(double), c.stream));
    CU(cudaMemcpyAsync(dba.p, body_w0, nb * 6 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    CU(dw0.put(w0_rows, Np * 6 * sizeof(double), c.stream));
    CU(dri.put(release_idx, Np * sizeof(int32_t), c.stream));
    CU(dtg.put(t, (size_t)ntimes * sizeof(double), c.stream));
    const size_t sb = (size_t)nout * ntot * 6 * sizeof(double);
    CU(dsn.put(nullptr, sb, c.stream));
    CU(dop.put(nullptr, Np * 6 * sizeof(double), c.stream));
    CU(dob.put(nullptr, nb * 6 * sizeof(double), c.stream));
    const size_t nst = Np + 1;
    CU(dst.put(nullptr, nst * sizeof(int32_t), c.stream));
    cudaError_t e = KCALL(c, nbody_dop853_march, r.P, B, a, (double*)dba.p, nullptr, nullptr, Np, 0, (const double*)dtg.p,
                          ntimes, output_every, (double*)dsn.p, nullptr, (double*)dob.p, 0, (int32_t*)dst.p + Np, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "nbody_dop853_march (bodies) launch");
    g_launches++;
    if (Np) {
        e = KCALL(c, nbody_dop853_march, r.P, B, a, (double*)dba.p, (const double*)dw0.p, (const int32_t*)dri.p, Np, 1,
                  (const double*)dtg.p, ntimes, output_every, (double*)dsn.p, (double*)dop.p, nullptr, 0,
                  (int32_t*)dst.p, c.stream);
        if (e != cudaSuccess) return cuda_fail(e, "nbody_dop853_march (particles) launch");
        g_launches++;
    }
    std::vector<int32_t> hs(nst);
    CU(cudaMemcpyAsync(snapshots, dsn.p, sb, cudaMemcpyDeviceToHost, c.stream));
    if (Np && out_particles) CU(cudaMemcpyAsync(out_particles, dop.p, Np * 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    if (out_bodies) CU(cudaMemcpyAsync(out_bodies, dob.p, nb * 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaMemcpyAsync(hs.data(), dst.p, nst * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    int worst = 0;
    for (size_t i = 0; i < nst; i++) { if (hs[i] < worst) worst = hs[i]; if (status && i < Np) status[i] = hs[i]; }
    if (worst < 0) return fail(worst, "Integration failed with code " + std::to_string(worst));
    return 0;
}

int gb_lyapunov_max(const gb_potential* pot, const gb_frame* fr, const double* w0_rows, const double* d0_vec, size_t N,
                    const double* t, int n_steps, double d0, int n_steps_per_pullback, int noffset_orbits, double atol,
                    double rtol, long nmax, double* LEs_raw, double* traj, int32_t* status, const gb_launch* opt) {
    Ctx c; RET_IF(open_ctx(opt, c));
    if (!c.host) return fail(-12, "gb_lyapunov_max takes host buffers");
    if (n_steps < 2 || !t) return fail(-12, "the time grid needs at least 2 entries");
    if (noffset_orbits < 1 || noffset_orbits > GB_MAXB) return fail(-12, "1..16 offset orbits per parent orbit");
    if (n_steps_per_pullback < 1 || !(d0 > 0.)) return fail(-12, "n_steps_per_pullback >= 1 and d0 > 0 required");
    if (N && (!w0_rows || !d0_vec || !LEs_raw)) return fail(-12, "null data pointer");
    RET_IF(pool_keep());
    DevFrame F; RET_IF(resolve_frame(fr, F));
    Resolved r; RET_IF(resolve(pot, r, c.stream));
    Dop853Args a;
    RET_IF(dop853_defaults(a, atol, rtol, nmax, 0.0, 1, 0.0, t[1] - t[0]));      // dop853_step (dop853.pyx:45-69)
    const int norb = 1 + noffset_orbits, niter = n_steps / n_steps_per_pullback;
    DevTmp dw0, dd0, dtg, dle, dtr, dst;
    CU(dw0.put(w0_rows, N * 6 * sizeof(double), c.stream));
    CU(dd0.put(d0_vec, N * (size_t)noffset_orbits * 6 * sizeof(double), c.stream));
    CU(dtg.put(t, (size_t)n_steps * sizeof(double), c.stream));
    const size_t lb = N * (size_t)niter * noffset_orbits * sizeof(double);
    CU(dle.put(nullptr, lb, c.stream));
    CU(cudaMemsetAsync(dle.p, 0, lb ? lb : 8, c.stream));                       // LEs = np.zeros (:44)
    const size_t tb = traj ? N * (size_t)n_steps * norb * 6 * sizeof(double) : 0;
    if (traj) { CU(dtr.put(nullptr, tb, c.stream)); CU(cudaMemsetAsync(dtr.p, 0, tb, c.stream)); }
    CU(dst.put(nullptr, (N ? N : 1) * sizeof(int32_t), c.stream));
    cudaError_t e = KCALL(c, lyapunov, r.P, F, a, (const double*)dw0.p, (const double*)dd0.p, N, (const double*)dtg.p,
                          n_steps, d0, n_steps_per_pullback, noffset_orbits, (double*)dle.p, (double*)dtr.p,
                          (int32_t*)dst.p, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "lyapunov launch");
    if (N) g_launches++;
    std::vector<int32_t> hs(N);
    if (N) {
        if (lb) CU(cudaMemcpyAsync(LEs_raw, dle.p, lb, cudaMemcpyDeviceToHost, c.stream));
        if (traj) CU(cudaMemcpyAsync(traj, dtr.p, tb, cudaMemcpyDeviceToHost, c.stream));
        CU(cudaMemcpyAsync(hs.data(), dst.p, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    }
    CU(cudaStreamSynchronize(c.stream));
    int worst = 0;
    for (size_t i = 0; i < N; i++) { if (hs[i] < worst) worst = hs[i]; if (status) status[i] = hs[i]; }
    if (worst < 0) return fail(worst, "Integration failed with code " + std::to_string(worst));
    return 0;
}

// ---- the partition of a multi-device call, for callers and tests (host-only arithmetic) -----------------
int gb_shard_bounds(size_t N, int k, int nd, size_t* lo, size_t* n) {
    if (nd < 1 || k < 0 || k >= nd || !lo || !n) return fail(-12, "gb_shard_bounds: need 0 <= k < nd and output pointers");
    slice_of(N, k, nd, lo, n);
    return 0;
}
long gb_deal_count(size_t Np, int k, int nd) {
    if (nd < 1 || k < 0 || k >= nd) return fail(-12, "gb_deal_count: need 0 <= k < nd");
    Deal D; D.Np = Np; D.k = k; D.nd = nd;
    return (long)D.count();
}

}  // extern "C"
