// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// C-ABI driver that links the reference's OWN C++ objects (compiled in place from
// /root/reference/src/gala by oracle/Makefile into oracle/_ref/) and restates only the
// Cython time loops that cannot be imported here (they import astropy at module load).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the resulting library.
//
// Each function cites the reference lines it follows (paths relative to src/gala/).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "potential/src/cpotential.h"
#include "potential/builtin/builtin_potentials.h"
#include "frame/src/cframe.h"
#include "frame/builtin/builtin_frames.h"
#include "hamiltonian/src/chamiltonian.h"
#include "dopri/dop853.h"
#if GB_REF_HAVE_SCF
#include "scf/src/bfe.h"
#include "potential/builtin/multipole.h"
#endif

#include "gala_b200.h"

extern void Fwrapper_T(unsigned full_ndim, double t, double *w, double *f, CPotential *p,
                       CFrameType *fr, unsigned norbits, unsigned na, void *args);
extern void Fwrapper(unsigned full_ndim, double t, double *w, double *f, CPotential *p,
                     CFrameType *fr, unsigned norbits, unsigned na, void *args);
extern double six_norm(double *x);
extern void Fwrapper_direct_nbody(unsigned full_ndim, double t, double *w, double *f,
                                  CPotential *p, CFrameType *fr, unsigned norbits,
                                  unsigned nbody, void *args);

namespace {

// Owns a CPotential built from a gb_potential spec.  Follows CPotentialWrapper.init
// (potential/potential/cpotential.pyx:57-102) and the *Wrapper.__init__ bodies
// (potential/potential/builtin/cybuiltin.pyx:126-376): one function-pointer set per type.
struct RefPotential;
#if GB_REF_HAVE_SCF
// one TimeInterpolated component: the reference's TimeInterpState (opaque here) + the wrapped CPotential it points to
struct RefTimeInterp {
    void *state = nullptr;
    RefPotential *wrapped = nullptr;
    ~RefTimeInterp();
};
#endif
struct RefPotential {
    CPotential *cp = nullptr;
    std::vector<std::vector<double>> pars, q0, R;
#if GB_REF_HAVE_SCF
    std::vector<RefTimeInterp *> ti;
#endif
    ~RefPotential();
};

#if GB_REF_HAVE_SCF
// scf_value / scf_density take 4 arguments in the reference (scf/src/bfe.h); the Cython
// wrapper casts them to energyfunc (scf/bfe_class.pyx).  Adapt explicitly here.
double scf_value5(double t, double *pars, double *q, int n_dim, void *) { return scf_value(t, pars, q, n_dim); }
double scf_density5(double t, double *pars, double *q, int n_dim, void *) { return scf_density(t, pars, q, n_dim); }
// same for the multipole functions (builtin/multipole.h:3-5; cybuiltin.pyx:385-388)
double mp_potential5(double t, double *pars, double *q, int n_dim, void *) { return mp_potential(t, pars, q, n_dim); }
double mp_density5(double t, double *pars, double *q, int n_dim, void *) { return mp_density(t, pars, q, n_dim); }
#endif

#if GB_REF_HAVE_SCF
}  // namespace
// oracle/ref_timeinterp.cpp (compiled with USE_GSL == 1): restates TimeInterpolatedWrapper.__init__
// (potential/potential/builtin/cytimeinterp.pyx:73-300) on the reference's own time_interp_* functions
extern "C" void *gb_ref_ti_build(const double *params, int n_params, CPotential *wrapped);
extern "C" void gb_ref_ti_free(void *state);
extern "C" void gb_ref_ti_hook(CPotential *cp, int i, void *state);
namespace {
RefTimeInterp::~RefTimeInterp() { if (state) gb_ref_ti_free(state); delete wrapped; }
RefPotential::~RefPotential() { if (cp) free_cpotential(cp); for (auto *t : ti) delete t; }
#else
RefPotential::~RefPotential() { if (cp) free_cpotential(cp); }
#endif

bool fill_component(CPotential *cp, int i, int type_id) {
    switch (type_id) {
    case GB_POT_NULL:
        cp->value[i] = null_value; cp->density[i] = null_density;
        cp->gradient[i] = null_gradient; cp->hessian[i] = null_hessian; return true;
    case GB_POT_HERNQUIST:
        cp->value[i] = hernquist_value; cp->density[i] = hernquist_density;
        cp->gradient[i] = hernquist_gradient; cp->hessian[i] = hernquist_hessian; return true;
    case GB_POT_NFW_SPHERICAL:
        cp->value[i] = sphericalnfw_value; cp->density[i] = sphericalnfw_density;
        cp->gradient[i] = sphericalnfw_gradient; cp->hessian[i] = sphericalnfw_hessian; return true;
    case GB_POT_NFW_FLATTENED:   // FlattenedNFWWrapper has no density (cybuiltin.pyx:303-311 uses nan_density)
        cp->value[i] = flattenednfw_value; cp->density[i] = nan_density;
        cp->gradient[i] = flattenednfw_gradient; cp->hessian[i] = flattenednfw_hessian; return true;
    case GB_POT_NFW_TRIAXIAL:
        cp->value[i] = triaxialnfw_value; cp->density[i] = nan_density;
        cp->gradient[i] = triaxialnfw_gradient; cp->hessian[i] = triaxialnfw_hessian; return true;
    case GB_POT_MIYAMOTONAGAI:
        cp->value[i] = miyamotonagai_value; cp->density[i] = miyamotonagai_density;
        cp->gradient[i] = miyamotonagai_gradient; cp->hessian[i] = miyamotonagai_hessian; return true;
    case GB_POT_MN3:
        cp->value[i] = mn3_value; cp->density[i] = mn3_density;
        cp->gradient[i] = mn3_gradient; cp->hessian[i] = mn3_hessian; return true;
    case GB_POT_LONGMURALIBAR:
        cp->value[i] = longmuralibar_value; cp->density[i] = longmuralibar_density;
        cp->gradient[i] = longmuralibar_gradient; cp->hessian[i] = longmuralibar_hessian; return true;
    case GB_POT_KEPLER:
        cp->value[i] = kepler_value; cp->density[i] = kepler_density;
        cp->gradient[i] = kepler_gradient; cp->hessian[i] = kepler_hessian; return true;
    case GB_POT_PLUMMER:
        cp->value[i] = plummer_value; cp->density[i] = plummer_density;
        cp->gradient[i] = plummer_gradient; cp->hessian[i] = plummer_hessian; return true;
    case GB_POT_ISOCHRONE:
        cp->value[i] = isochrone_value; cp->density[i] = isochrone_density;
        cp->gradient[i] = isochrone_gradient; cp->hessian[i] = isochrone_hessian; return true;
    case GB_POT_JAFFE:
        cp->value[i] = jaffe_value; cp->density[i] = jaffe_density;
        cp->gradient[i] = jaffe_gradient; cp->hessian[i] = jaffe_hessian; return true;
    case GB_POT_STONE:
        cp->value[i] = stone_value;      // the header declares stone_density as void (builtin_potentials.h:95); the definition returns double
        cp->density[i] = reinterpret_cast<densityfunc>(stone_density);
        cp->gradient[i] = stone_gradient; cp->hessian[i] = stone_hessian; return true;
    case GB_POT_BURKERT:         // BurkertWrapper sets no hessian (cybuiltin.pyx:226-234)
        cp->value[i] = burkert_value; cp->density[i] = burkert_density;
        cp->gradient[i] = burkert_gradient; cp->hessian[i] = null_hessian; return true;
    case GB_POT_SATOH:
        cp->value[i] = satoh_value; cp->density[i] = satoh_density;
        cp->gradient[i] = satoh_gradient; cp->hessian[i] = satoh_hessian; return true;
    case GB_POT_KUZMIN:
        cp->value[i] = kuzmin_value; cp->density[i] = kuzmin_density;
        cp->gradient[i] = kuzmin_gradient; cp->hessian[i] = null_hessian; return true;
    case GB_POT_LOGARITHMIC:
        cp->value[i] = logarithmic_value; cp->density[i] = logarithmic_density;
        cp->gradient[i] = logarithmic_gradient; cp->hessian[i] = logarithmic_hessian; return true;
    case GB_POT_LEESUTO:
        cp->value[i] = leesuto_value; cp->density[i] = leesuto_density;
        cp->gradient[i] = leesuto_gradient; cp->hessian[i] = null_hessian; return true;
#if GB_REF_HAVE_SCF
    case GB_POT_POWERLAWCUTOFF:  // compiled only with USE_GSL == 1 (builtin_potentials.cpp:465)
        cp->value[i] = powerlawcutoff_value; cp->density[i] = powerlawcutoff_density;
        cp->gradient[i] = powerlawcutoff_gradient; cp->hessian[i] = powerlawcutoff_hessian; return true;
    case GB_POT_SCF:
        cp->value[i] = scf_value5; cp->density[i] = scf_density5;
        cp->gradient[i] = scf_gradient; cp->hessian[i] = null_hessian; return true;
    case GB_POT_MULTIPOLE:
        cp->value[i] = mp_potential5; cp->density[i] = mp_density5;
        cp->gradient[i] = mp_gradient; cp->hessian[i] = null_hessian; return true;
#endif
    default: return false;
    }
}

bool build(const gb_potential *spec, RefPotential &out) {
    int nc = spec->n_components;
    out.cp = allocate_cpotential(nc);
    out.cp->n_dim = spec->n_dim;
    out.pars.resize(nc); out.q0.resize(nc); out.R.resize(nc);
    int all_null = 1;
    for (int i = 0; i < nc; i++) {
        const gb_component &c = spec->comp[i];
#if GB_REF_HAVE_SCF
        if (c.type_id == GB_POT_TIMEINTERP) {
            // wrapped potential: a one-component CPotential of the wrapped type with the first knot's parameters
            // (cytimeinterp.pyx:283 keeps wrapped_potential.cpotential only for its function pointers and state)
            const int wtype = (int)c.params[1], n = (int)c.params[3], nwp = (int)c.params[4];
            std::vector<double> wp(1 + nwp);
            wp[0] = c.params[0];
            for (int k = 0; k < nwp; k++) wp[1 + k] = c.params[7 + n + k];
            gb_component wc;
            memset(&wc, 0, sizeof(wc));
            wc.type_id = wtype; wc.n_params = 1 + nwp; wc.params = wp.data(); wc.R[0] = wc.R[4] = wc.R[8] = 1.;
            gb_potential wspec = {1, 3, &wc};
            RefTimeInterp *t = new RefTimeInterp;
            t->wrapped = new RefPotential;
            out.ti.push_back(t);
            if (!build(&wspec, *t->wrapped)) return false;
            t->state = gb_ref_ti_build(c.params, c.n_params, t->wrapped->cp);
            if (!t->state) return false;
            all_null = 0;
            // CPotentialWrapper.init([0.0], zeros, eye) (cytimeinterp.pyx:285-290): G placeholder, no shift/rotate
            out.pars[i].assign(1, 0.); out.q0[i].assign(3, 0.); out.R[i].assign(9, 0.);
            out.R[i][0] = out.R[i][4] = out.R[i][8] = 1.;
            out.cp->n_params[i] = 1; out.cp->parameters[i] = out.pars[i].data();
            out.cp->q0[i] = out.q0[i].data(); out.cp->R[i] = out.R[i].data();
            out.cp->do_shift_rotate[i] = 0;
            gb_ref_ti_hook(out.cp, i, t->state);
            continue;
        }
#endif
        if (!fill_component(out.cp, i, c.type_id)) return false;
        if (c.type_id != GB_POT_NULL) all_null = 0;
        out.pars[i].assign(c.params, c.params + c.n_params);
        if (out.pars[i].empty()) out.pars[i].push_back(0.);
        out.q0[i].assign(c.q0, c.q0 + 3);
        out.R[i].assign(c.R, c.R + 9);
        out.cp->n_params[i] = c.n_params;
        out.cp->parameters[i] = out.pars[i].data();
        out.cp->q0[i] = out.q0[i].data();
        out.cp->R[i] = out.R[i].data();
        out.cp->state[i] = NULL;
        out.cp->do_shift_rotate[i] = c.do_shift_rotate;
    }
    out.cp->null = all_null;
    return true;
}

// Frame wrappers: potential/frame/builtin/frames.pyx:37-50 (static), :90-106 (rotating 3D).
struct RefFrame {
    CFrameType cf;
    double omega[3];
};
void build_frame(const gb_frame *fr, RefFrame &out) {
    memset(&out.cf, 0, sizeof(out.cf));
    if (!fr || fr->type_id == GB_FRAME_STATIC) {
        out.cf.energy = (energyfunc)static_frame_hamiltonian;
        out.cf.gradient = (gradientfunc)static_frame_gradient;
        out.cf.hessian = (hessianfunc)static_frame_hessian;
        out.cf.n_params = 0;
        out.cf.parameters = NULL;
    } else {
        for (int k = 0; k < 3; k++) out.omega[k] = fr->omega[k];
        out.cf.energy = (energyfunc)constant_rotating_frame_3d_hamiltonian;
        out.cf.gradient = (gradientfunc)constant_rotating_frame_3d_gradient;
        out.cf.hessian = (hessianfunc)constant_rotating_frame_3d_hessian;
        out.cf.n_params = 3;
        out.cf.parameters = out.omega;
    }
}

}  // namespace

extern "C" {

// CPotentialWrapper.gradient (cpotential.pyx:144-162): one c_gradient(N) call.
int ref_gradient(const gb_potential *spec, const double *q, double t, size_t N, double *grad) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    c_gradient(rp.cp, N, t, const_cast<double *>(q), grad);
    return 0;
}

// CPotentialWrapper.energy / density (cpotential.pyx:104-142): per-point c_potential /
// c_density on an AoS copy of each point.  q here is (3,N) like everywhere else in the ABI.
int ref_energy(const gb_potential *spec, const double *q, double t, size_t N, double *out) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    for (size_t i = 0; i < N; i++) {
        double p[3] = {q[i], q[N + i], q[2 * N + i]};
        out[i] = c_potential(rp.cp, t, p);
    }
    return 0;
}
int ref_density(const gb_potential *spec, const double *q, double t, size_t N, double *out) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    for (size_t i = 0; i < N; i++) {
        double p[3] = {q[i], q[N + i], q[2 * N + i]};
        out[i] = c_density(rp.cp, t, p);
    }
    return 0;
}

// Hamiltonian.energy (hamiltonian/chamiltonian.pyx:107-128): potential energy via
// c_potential + frame energy (frame_hamiltonian), per point.
int ref_hamiltonian_energy(const gb_potential *spec, const gb_frame *fr, const double *w, double t,
                           size_t N, double *out) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefFrame rf; build_frame(fr, rf);
    for (size_t i = 0; i < N; i++) {
        double p[6];
        for (int k = 0; k < 6; k++) p[k] = w[k * N + i];
        out[i] = c_potential(rp.cp, t, p) + frame_hamiltonian(&rf.cf, t, p, 3);
    }
    return 0;
}

int ref_hamiltonian_gradient(const gb_potential *spec, const gb_frame *fr, const double *w, double t,
                             size_t N, double *f) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefFrame rf; build_frame(fr, rf);
    hamiltonian_gradient_T(rp.cp, &rf.cf, N, t, const_cast<double *>(w), f);
    return 0;
}

// leapfrog_integrate_hamiltonian (integrate/cyintegrators/leapfrog.pyx:54-121) with
// c_init_velocity (:24-32) and c_leapfrog_step (:35-51).
int ref_leapfrog(const gb_potential *spec, const double *w0, size_t N, const double *t, int ntimes,
                 int save_all, double *w_out) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    const size_t n = N;
    const double dt = t[1] - t[0];
    std::vector<double> tmp_w(w0, w0 + 6 * n), v12(3 * n, 0.), grad(3 * n, 0.);
    double *x = tmp_w.data(), *v = tmp_w.data() + 3 * n;
    if (save_all)
        for (int k = 0; k < 6; k++) memcpy(w_out + (size_t)k * ntimes * n, w0 + k * n, n * sizeof(double));

    c_gradient(rp.cp, n, t[0], x, grad.data());
    for (int k = 0; k < 3; k++)
        for (size_t i = 0; i < n; i++) v12[i + k * n] = v[i + k * n] - grad[i + k * n] * dt / 2.;

    for (int j = 1; j < ntimes; j++) {
        for (int k = 0; k < 3; k++)
            for (size_t i = 0; i < n; i++) x[i + k * n] = x[i + k * n] + v12[i + k * n] * dt;
        c_gradient(rp.cp, n, t[j], x, grad.data());
        for (int k = 0; k < 3; k++)
            for (size_t i = 0; i < n; i++) {
                v[i + k * n] = v12[i + k * n] - grad[i + k * n] * dt / 2.;
                v12[i + k * n] = v12[i + k * n] - grad[i + k * n] * dt;
            }
        if (save_all)
            for (int k = 0; k < 6; k++)
                memcpy(w_out + ((size_t)k * ntimes + j) * n, tmp_w.data() + k * n, n * sizeof(double));
    }
    if (!save_all) memcpy(w_out, tmp_w.data(), 6 * n * sizeof(double));
    return 0;
}

// Ruth4 coefficients: integrate/cyintegrators/ruth4.pyx:65-78.
static void ruth4_coeffs(double *cs, double *ds) {
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

// ruth4_integrate_hamiltonian (integrate/cyintegrators/ruth4.pyx:37-113, step :24-35) for the
// static frame.  For the rotating frame the reference only runs through the Python
// Ruth4Integrator (integrate/pyintegrators/ruth4.py:106-124) with
// F = Hamiltonian._gradient (hamiltonian/chamiltonian.pyx:88-99), i.e. F = hamiltonian_gradient_T;
// the integrator uses only a = F[3:]:  p += d_j*a*dt ; q += c_j*p*dt.
int ref_ruth4(const gb_potential *spec, const gb_frame *fr, const double *w0, size_t N, const double *t,
              int ntimes, int save_all, double *w_out) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefFrame rf; build_frame(fr, rf);
    const bool rotating = fr && fr->type_id == GB_FRAME_ROTATING_3D;
    const size_t n = N;
    const double dt = t[1] - t[0];
    double cs[4], ds[4];
    ruth4_coeffs(cs, ds);
    std::vector<double> w(w0, w0 + 6 * n), grad(3 * n, 0.), F(6 * n, 0.);
    if (save_all)
        for (int k = 0; k < 6; k++) memcpy(w_out + (size_t)k * ntimes * n, w0 + k * n, n * sizeof(double));
    for (int j = 1; j < ntimes; j++) {
        for (int s = 0; s < 4; s++) {
            if (!rotating) {
                c_gradient(rp.cp, n, t[j], w.data(), grad.data());
                for (int k = 0; k < 3; k++)
                    for (size_t i = 0; i < n; i++) {
                        w[(3 + k) * n + i] = w[(3 + k) * n + i] - ds[s] * grad[k * n + i] * dt;
                        w[k * n + i] = w[k * n + i] + cs[s] * w[(3 + k) * n + i] * dt;
                    }
            } else {
                // pyintegrators/ruth4.py:110-122: a_i = F(t, w)[ndim:]; p = p + d*a_i*dt; q = q + c*p*dt
                // step() is called with times[ii] (ruth4.py:141), the same t[j] the Cython loop uses.
                hamiltonian_gradient_T(rp.cp, &rf.cf, n, t[j], w.data(), F.data());
                for (int k = 0; k < 3; k++)
                    for (size_t i = 0; i < n; i++) {
                        w[(3 + k) * n + i] = w[(3 + k) * n + i] + ds[s] * F[(3 + k) * n + i] * dt;
                        w[k * n + i] = w[k * n + i] + cs[s] * w[(3 + k) * n + i] * dt;
                    }
            }
        }
        if (save_all)
            for (int k = 0; k < 6; k++)
                memcpy(w_out + ((size_t)k * ntimes + j) * n, w.data() + k * n, n * sizeof(double));
    }
    if (!save_all) memcpy(w_out, w.data(), 6 * n * sizeof(double));
    return 0;
}

// The reference's dopcor keeps nstep / naccpt / nrejct / nfcn as locals and its header only DECLARES
// nfcnRead() ... nrejctRead() (dopri/dop853.h:252-255; no definition exists in dop853.cpp), so the step
// statistics cannot be read from the unmodified sources.  What CAN be observed from outside is every call of
// the right-hand side: this wrapper is handed to dop853() in place of Fwrapper_T, counts, and forwards.
// With nbatch = 1 the count is dopcor's nfcn of that orbit: 2 + 11 nstep + naccpt (+ 3 naccpt with dense output).
static thread_local long g_fcn_calls = 0;
static void counting_Fwrapper_T(unsigned full_ndim, double t, double *w, double *f, CPotential *p, CFrameType *fr,
                                unsigned norbits, unsigned na, void *args) {
    g_fcn_calls++;
    Fwrapper_T(full_ndim, t, w, f, p, fr, norbits, na, args);
}

int ref_dop853_nfcn(const gb_potential *spec, const gb_frame *fr, const double *w0, size_t N, const double *t,
                    int ntimes, double atol, double rtol, long nmax, double dt_max, long nstiff, int save_all,
                    int nbatch, double *w_out, int32_t *status, int32_t *nfcn);

// dop853_integrate_hamiltonian (integrate/cyintegrators/dop853.pyx:196-250) calling
// dop853_helper (:90-193) per batch of `nbatch` orbits.  status[i] = dop853 return code of the
// batch orbit i belongs to.  Returns the most negative code, or 0.
int ref_dop853(const gb_potential *spec, const gb_frame *fr, const double *w0, size_t N, const double *t,
               int ntimes, double atol, double rtol, long nmax, double dt_max, long nstiff, int save_all,
               int nbatch, double *w_out, int32_t *status) {
    return ref_dop853_nfcn(spec, fr, w0, N, t, ntimes, atol, rtol, nmax, dt_max, nstiff, save_all, nbatch, w_out,
                           status, NULL);
}

// nfcn (may be NULL): per orbit, the number of right-hand-side calls dop853() made for the batch it belongs to.
int ref_dop853_nfcn(const gb_potential *spec, const gb_frame *fr, const double *w0, size_t N, const double *t,
                    int ntimes, double atol, double rtol, long nmax, double dt_max, long nstiff, int save_all,
                    int nbatch, double *w_out, int32_t *status, int32_t *nfcn) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefFrame rf; build_frame(fr, rf);
    if (ntimes < 1) return -12;
    int worst = 0;
    for (size_t i0 = 0; i0 < N; i0 += nbatch) {
        size_t nb = (i0 + nbatch <= N) ? nbatch : N - i0;
        unsigned size = 6 * nb;
        std::vector<double> w(size), out((size_t)ntimes * size);
        for (int k = 0; k < 6; k++)
            for (size_t i = 0; i < nb; i++) w[k * nb + i] = w0[k * N + i0 + i];
        Dop853DenseState *state = save_all ? dop853_dense_state_alloc(size, size) : NULL;
        double rt = rtol, at = atol;
        g_fcn_calls = 0;
        int res = dop853(size, nfcn ? (FcnEqDiff)counting_Fwrapper_T : (FcnEqDiff)Fwrapper_T, rp.cp, &rf.cf, nb, 0, NULL,
                         t[0], w.data(),
                         t[ntimes - 1], &rt, &at, 0, NULL, 0, NULL,
                         2.220446049250313e-16,  // np.finfo(float).eps (dop853.pyx:165)
                         0.0, 0.0, 0.0, 0.0, dt_max, t[1] - t[0], nmax, 1, nstiff,
                         save_all ? size : 0, NULL, 0, state, const_cast<double *>(t), ntimes,
                         save_all ? out.data() : NULL);
        if (state) dop853_dense_state_free(state, size);
        if (res < worst) worst = res;
        for (size_t i = 0; i < nb; i++) if (status) status[i0 + i] = res;
        for (size_t i = 0; i < nb; i++) if (nfcn) nfcn[i0 + i] = (int32_t)g_fcn_calls;
        if (save_all) {
            // wres[:, :, i:j] = wbatchout.transpose(1,0,2)   (dop853.pyx:242-243)
            for (int j = 0; j < ntimes; j++)
                for (int k = 0; k < 6; k++)
                    for (size_t i = 0; i < nb; i++)
                        w_out[((size_t)k * ntimes + j) * N + i0 + i] = out[(size_t)j * size + k * nb + i];
        } else {
            for (int k = 0; k < 6; k++)
                for (size_t i = 0; i < nb; i++) w_out[k * N + i0 + i] = w[k * nb + i];
        }
    }
    return worst;
}

// One (Np,6)-row-per-particle integration with dop853_step's settings
// (integrate/cyintegrators/dop853.pyx:27-75): uround=0 (-> 2.3e-16), nstiff hard-coded 1,
// no dense output, F = Fwrapper_direct_nbody with nbody=0 bodies carrying mass, i.e. the
// no-self-gravity case of mockstream_dop853 (dynamics/mockstream/mockstream.pyx:259-283).
// group=1 integrates each particle as its own n=6 system (the per-lane definition used by the
// GPU engine); group=0 integrates ALL given particles as one coupled system of 6*Np.
int ref_dop853_step_rows(const gb_potential *spec, const gb_frame *fr, double *w_rows, size_t Np, double t1,
                         double t2, double dt0, double atol, double rtol, long nmax, int group,
                         int32_t *status) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefFrame rf; build_frame(fr, rf);
    int worst = 0;
    double rt = rtol, at = atol;
    if (!group) {
        int res = dop853(6 * Np, (FcnEqDiff)Fwrapper_direct_nbody, rp.cp, &rf.cf, Np, 0, NULL, t1, w_rows, t2,
                         &rt, &at, 0, NULL, 0, NULL, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, dt0, nmax, 0, 1, 0, NULL, 0,
                         NULL, NULL, 0, NULL);
        for (size_t i = 0; i < Np; i++) if (status) status[i] = res;
        return res < 0 ? res : 0;
    }
    for (size_t i = 0; i < Np; i++) {
        int res = dop853(6, (FcnEqDiff)Fwrapper_direct_nbody, rp.cp, &rf.cf, 1, 0, NULL, t1, w_rows + 6 * i, t2,
                         &rt, &at, 0, NULL, 0, NULL, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, dt0, nmax, 0, 1, 0, NULL, 0,
                         NULL, NULL, 0, NULL);
        if (status) status[i] = res;
        if (res < worst) worst = res;
    }
    return worst;
}

// dop853_lyapunov_max (dynamics/lyapunov/dop853_lyapunov.pyx:22-118): parent orbit w0 + noff offset orbits
// w0 + d0_vec[i] (already of length d0) advanced interval by interval with dop853_step's call
// (integrate/cyintegrators/dop853.pyx:45-69) over F = Fwrapper, pulled back every `pullback` intervals.
// LEs_raw (niter, noff) = ln(|d1|/d0); all_w (n_steps, 1+noff, 6) or NULL.  Returns the dop853 code.
int ref_lyapunov(const gb_potential *spec, const gb_frame *fr, const double *w0, const double *d0_vec,
                 const double *t, int n_steps, double d0, int pullback, int noff, double atol, double rtol,
                 long nmax, double *LEs_raw, double *all_w) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefFrame rf; build_frame(fr, rf);
    const int norb = 1 + noff;
    std::vector<double> w(6 * norb), d1(6);
    for (int k = 0; k < 6; k++) w[k] = w0[k];
    for (int i = 1; i < norb; i++)
        for (int k = 0; k < 6; k++) w[6 * i + k] = w0[k] + d0_vec[(i - 1) * 6 + k];
    if (all_w) memcpy(all_w, w.data(), 6 * norb * sizeof(double));
    const double dt0 = t[1] - t[0];
    int jiter = 0;
    for (int j = 1; j < n_steps; j++) {
        double rt = rtol, at = atol;
        int res = dop853(6 * norb, (FcnEqDiff)Fwrapper, rp.cp, &rf.cf, norb, 0, NULL, t[j - 1], w.data(), t[j],
                         &rt, &at, 0, NULL, 0, NULL, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, dt0, nmax, 0, 1, 0, NULL, 0,
                         NULL, NULL, 0, NULL);
        if (res < 0) return res;
        if (all_w) memcpy(all_w + (size_t)j * 6 * norb, w.data(), 6 * norb * sizeof(double));
        if ((j % pullback) == 0) {
            for (int i = 1; i < norb; i++) {
                for (int k = 0; k < 6; k++) d1[k] = w[6 * i + k] - w[k];
                const double mag = six_norm(d1.data());
                LEs_raw[jiter * noff + (i - 1)] = log(mag / d0);
                for (int k = 0; k < 6; k++) w[6 * i + k] = w[k] + d0 * d1[k] / mag;
            }
            jiter++;
        }
    }
    return 1;
}

// CPotentialWrapper.hessian (potential/potential/cpotential.pyx:164-182): c_hessian per point; q (3,N) SoA
// in, hess (3,3,N) out (the layout PotentialBase.hessian returns, core.py:535-600).
int ref_hessian(const gb_potential *spec, const double *q, double t, size_t N, double *hess) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    for (size_t n = 0; n < N; n++) {
        double qq[3] = {q[n], q[N + n], q[2 * N + n]}, h[9];
        c_hessian(rp.cp, t, qq, h);
        for (int k = 0; k < 9; k++) hess[(size_t)k * N + n] = h[k];
    }
    return 0;
}

// c_d2_dr2 (potential/potential/src/cpotential.cpp:346-371) exposed for the release tests.
double ref_d2_dr2(const gb_potential *spec, double t, const double *q3) {
    RefPotential rp; if (!build(spec, rp)) return NAN;
    double eps[3], q[3] = {q3[0], q3[1], q3[2]};
    return c_d2_dr2(rp.cp, t, q, eps);
}

// ---- massive bodies -----------------------------------------------------------------------------------
// System rows (n, 6): the first `nbodies` rows are bodies with their own potentials body_specs[b]
// (a Null-type spec = massless), the rest are test particles (Null potentials, as
// _setup_particle_potentials does, dynamics/mockstream/mockstream.pyx:49-77).
struct RefBodies {
    std::vector<RefPotential> pots;       // one per row of the system
    std::vector<CPotential *> ptrs;
    bool build_all(const gb_potential *body_specs, int nbodies, size_t n, const gb_potential *null_spec) {
        pots.resize(n); ptrs.resize(n);
        for (size_t i = 0; i < n; i++) {
            if (!build(i < (size_t)nbodies ? &body_specs[i] : null_spec, pots[i])) return false;
            ptrs[i] = pots[i].cp;
        }
        return true;
    }
};

// leapfrog_integrate_nbody (integrate/cyintegrators/leapfrog.pyx:161-257) == the per-group loop of
// mockstream_leapfrog (dynamics/mockstream/mockstream.pyx:556-590): c_init_velocity_nbody for every row,
// then nsteps of c_leapfrog_step_nbody over the rows IN ORDER, in place.  `nsrc` is the `nbody` argument
// handed to c_nbody_gradient_symplectic (sources = rows j < nsrc).  traj: (nsteps+1, n, 6) or NULL.
int ref_nbody_leapfrog(const gb_potential *spec, const gb_potential *body_specs, int nbodies,
                       const gb_potential *null_spec, double *w_rows, size_t n, int nsrc, double t0, int nsteps,
                       double dt, double *traj) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefBodies rb; if (!rb.build_all(body_specs, nbodies, n, null_spec)) return -11;
    std::vector<double> v12(3 * n, 0.);
    double grad[3];
    if (traj) memcpy(traj, w_rows, n * 6 * sizeof(double));
    for (size_t i = 0; i < n; i++) {
        double *x = w_rows + 6 * i, *v = x + 3;
        grad[0] = grad[1] = grad[2] = 0.;
        c_gradient(rp.cp, 1, t0, x, grad);                                               // leapfrog.pyx:134
        c_nbody_gradient_symplectic(rb.ptrs.data(), t0, x, w_rows, nsrc, (int)i, 3, grad);
        for (int k = 0; k < 3; k++) v12[3 * i + k] = v[k] - grad[k] * dt / 2.;
    }
    for (int j = 0; j < nsteps; j++) {
        const double tj = t0 + (j + 1) * dt;
        for (size_t i = 0; i < n; i++) {
            double *x = w_rows + 6 * i, *v = x + 3;
            grad[0] = grad[1] = grad[2] = 0.;
            for (int k = 0; k < 3; k++) x[k] = x[k] + v12[3 * i + k] * dt;                 // leapfrog.pyx:147-148
            c_gradient(rp.cp, 1, tj, x, grad);
            c_nbody_gradient_symplectic(rb.ptrs.data(), tj, x, w_rows, nsrc, (int)i, 3, grad);
            for (int k = 0; k < 3; k++) {
                v[k] = v12[3 * i + k] - grad[k] * dt / 2.;
                v12[3 * i + k] = v12[3 * i + k] - grad[k] * dt;
            }
        }
        if (traj) memcpy(traj + (size_t)(j + 1) * n * 6, w_rows, n * 6 * sizeof(double));
    }
    return 0;
}

// ruth4_integrate_nbody (integrate/cyintegrators/ruth4.pyx:139-241) with c_ruth4_step_nbody (:116-136).
int ref_nbody_ruth4(const gb_potential *spec, const gb_potential *body_specs, int nbodies,
                    const gb_potential *null_spec, double *w_rows, size_t n, int nsrc, double t0, int nsteps,
                    double dt, double *traj) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefBodies rb; if (!rb.build_all(body_specs, nbodies, n, null_spec)) return -11;
    double cs[4], ds[4], grad[3];
    ruth4_coeffs(cs, ds);
    if (traj) memcpy(traj, w_rows, n * 6 * sizeof(double));
    for (int j = 0; j < nsteps; j++) {
        const double tj = t0 + (j + 1) * dt;
        for (size_t i = 0; i < n; i++) {
            double *w = w_rows + 6 * i;
            for (int q = 0; q < 4; q++) {
                grad[0] = grad[1] = grad[2] = 0.;
                c_gradient(rp.cp, 1, tj, w, grad);
                c_nbody_gradient_symplectic(rb.ptrs.data(), tj, w, w_rows, nsrc, (int)i, 3, grad);
                for (int k = 0; k < 3; k++) {
                    w[3 + k] = w[3 + k] - ds[q] * grad[k] * dt;
                    w[k] = w[k] + cs[q] * w[3 + k] * dt;
                }
            }
        }
        if (traj) memcpy(traj + (size_t)(j + 1) * n * 6, w_rows, n * 6 * sizeof(double));
    }
    return 0;
}

// dop853 over the whole system with Fwrapper_direct_nbody (dop853.cpp:990-1006).
// mode 0: dop853_helper's call (dop853.pyx:157-182) as used by direct_nbody_dop853 (nbody.pyx:96-110,
//         nstiff = -1) with dense output at tgrid -> traj (ntimes, n, 6) when traj != NULL;
// mode 1: dop853_step's call (dop853.pyx:45-69) from t1 to t2, final state only.
int ref_nbody_dop853(const gb_potential *spec, const gb_potential *body_specs, int nbodies,
                     const gb_potential *null_spec, double *w_rows, size_t n, const double *tgrid, int ntimes,
                     double t1, double t2, double dt0, double atol, double rtol, long nmax, double dt_max, int mode,
                     double *traj) {
    RefPotential rp; if (!build(spec, rp)) return -11;
    RefFrame rf; build_frame(NULL, rf);
    RefBodies rb; if (!rb.build_all(body_specs, nbodies, n, null_spec)) return -11;
    double rt = rtol, at = atol;
    const unsigned size = 6 * n;
    void *args = (void *)rb.ptrs.data();
    if (mode == 1)
        return dop853(size, (FcnEqDiff)Fwrapper_direct_nbody, rp.cp, &rf.cf, n, nbodies, args, t1, w_rows, t2,
                      &rt, &at, 0, NULL, 0, NULL, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, dt0, nmax, 0, 1, 0, NULL, 0,
                      NULL, NULL, 0, NULL);
    Dop853DenseState *state = traj ? dop853_dense_state_alloc(size, size) : NULL;
    int res = dop853(size, (FcnEqDiff)Fwrapper_direct_nbody, rp.cp, &rf.cf, n, nbodies, args, tgrid[0], w_rows,
                     tgrid[ntimes - 1], &rt, &at, 0, NULL, 0, NULL, 2.220446049250313e-16, 0.0, 0.0, 0.0, 0.0,
                     dt_max, tgrid[1] - tgrid[0], nmax, 1, -1, traj ? size : 0, NULL, 0, state,
                     const_cast<double *>(tgrid), ntimes, traj);
    if (state) dop853_dense_state_free(state, size);
    return res;
}

const char *ref_build_flags(void) { return GB_REF_FLAGS; }

}  // extern "C"
