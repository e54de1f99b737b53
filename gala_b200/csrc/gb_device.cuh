// gb_device.cuh -- device-side potential description and small math helpers.
//
// The potential arrives in a kernel as ONE __grid_constant__ struct (DevPot): it lives in the
// constant bank, every thread of a warp reads the same word, so parameter loads are uniform
// constant-cache operands of the FP64 instructions rather than register-file traffic.
// It replaces struct _CPotential (reference potential/potential/src/cpotential.h:10-36):
// type ids instead of host function pointers, all parameter vectors packed into one array.
#pragma once
#include <stdint.h>
#include "../../include/gala_b200.h"

#define GB_HAVE_SCF 1   // flipped to 1 when scf.cuh carries the recurrence implementation
#define GB_MP_LMAX 15  // largest multipole order the device evaluates
#define GB_MAXC 8      // components per composite held in the constant bank
#define GB_MAXP 120    // packed doubles of "small" parameters ([G, ...] of every component)
#define GB_CEXT 616    // 2 x 11 x 28 doubles: (S,T) pairs of SCF(nmax=10,lmax=6), [l][m][n]
#define GB_PLC_NINT 64   // PowerLawCutoff: Chebyshev fit of gamma*(a, s^2) on [0, GB_PLC_SMAX], s = r/r_c
#define GB_PLC_DEG 9
#define GB_PLC_SMAX 6.4
#define GB_MAXD 40     // packed doubles of host-derived constants (G*m, b^2, 1/r_s, ...) for the fast build

struct DevComp {
    int32_t type;      // gb_pot_type
    int32_t shift;     // do_shift_rotate
    int32_t poff;      // offset of this component's [G, ...] in DevPot::par
    int32_t npar;
    int32_t eoff;      // offset of this component's large-parameter block in DevPot::ext
    int32_t doff;      // offset of this component's derived constants in DevPot::drv
    double q0[3];
    double R[9];
};

struct DevPot {
    int32_t n;                 // n_components
    int32_t sig;               // GbSig the host resolved (informational on device)
    DevComp c[GB_MAXC];
    double par[GB_MAXP];
    double drv[GB_MAXD];       // per component: gb_nderived(type) doubles, see gb_derive() in capi.cu
    int32_t cext_ok;           // 1: cext holds the SCF coefficients of component 0 in the (10,6) padded layout
    int32_t time_dep;          // 1: some component is a TimeInterpolated wrapper (the kernels then pass real times)
    double cext[GB_CEXT];      // constant-bank copy of a small `ext` block (scf.cuh: scf_fast_gradient)
    const double* ext;         // device-global parameters of "large" components (SCF coefficients)
};

// Massive bodies carried by every lane of the N-body kernels (nbody.cuh): each body owns a small potential
// (its components are evaluated about the body's current position: c_nbody_acceleration /
// c_nbody_gradient_symplectic set do_shift_rotate = 1 and q0 = &w[body], cpotential.cpp:389-442).
#define GB_MAXB 16     // bodies (massive or massless) carried per system
#define GB_MAXBC 16    // potential components over all bodies
#define GB_MAXBP 96    // packed parameters over all body components
#define GB_ND_MAX (6 * (GB_MAXB + 1))   // largest DOP853 system of a lane: all bodies + its particle
struct DevBodies {
    int32_t nb;                   // bodies (massive or Null) at the front of the system
    int32_t nc;
    int32_t null_[GB_MAXB];       // CPotential::null of body b (skipped as a force source)
    int32_t cbeg[GB_MAXB + 1];    // components of body b: [cbeg[b], cbeg[b+1])
    int32_t type[GB_MAXBC];
    int32_t poff[GB_MAXBC];
    double R[GB_MAXBC][9];
    double par[GB_MAXBP];
};

struct DevFrame {
    int32_t type;              // gb_frame_type
    int32_t _pad;
    double om[3];
};

// Compile-time composite signatures the host can resolve a gb_potential to.  Everything else
// runs through SIG_GENERIC (a warp-uniform switch per component).
enum GbSig {
    SIG_GENERIC = 0,
    SIG_NFW,            // [SphericalNFW]
    SIG_HERNQUIST,      // [Hernquist]
    SIG_MW2022,         // [MN3, Hernquist, Hernquist, SphericalNFW]  (special.py:221-271 order)
    SIG_BAR_MW2022,     // [LongMuraliBar, MN3, Hernquist, Hernquist, SphericalNFW]
    SIG_MW2022_BAR,     // [MN3, Hernquist, Hernquist, SphericalNFW, LongMuraliBar]
    SIG_SCF,            // [SCF]
    SIG_MW_V1,          // [MiyamotoNagai, Hernquist, Hernquist, SphericalNFW]   MilkyWayPotential v1 (special.py:88-124)
    SIG_LM10,           // [MiyamotoNagai, Hernquist, Logarithmic]               LM10Potential (special.py:26-87)
    SIG_BOVY2014,       // [MiyamotoNagai, PowerLawCutoff, SphericalNFW]         BovyMWPotential2014 (special.py:274-347)
    SIG_GENERIC_LIGHT,  // any list of analytic components (no SCF / multipole): the generic loop at their register budget
    SIG_GENERIC_TI,     // analytic components of which at least one is a TimeInterpolated wrapper: the light loop + the
                        // interpolation code and real step times (kept out of SIG_GENERIC_LIGHT: it cost that loop 10 %)
    SIG_COUNT
};

// Number of host-derived constants per potential type (filled by capi.cu:gb_derive, read by the
// fast build's accum() functions in potentials.cuh; the strict build ignores them):
//   Hernquist/Kepler/Jaffe/Kuzmin [G m] ; Satoh [G m, b^2] ; NFW flattened/triaxial [G m, 1/r_s, 1/a^2, 1/b^2, 1/c^2] ; NFW spherical [G m, 1/r_s] ; MiyamotoNagai/Plummer/Isochrone [G m, b^2] ;
//   MN3 [G m1, G m2, G m3, b1^2, b2^2, b3^2] ; LongMuraliBar [G m, sin(alpha), cos(alpha), c^2] ;
//   SCF [G m / r_s^2, 1 / r_s] ; PowerLawCutoff [G m, lgamma(a), 1/r_c^2, 1/r_c, 2a, ext offset of the gamma* fit], a = (3-alpha)/2 ;
//   Logarithmic [v_c^2, r_h^2, 1/q1^2, 1/q2^2, 1/q3^2, sin(phi), cos(phi)].
constexpr int gb_nderived(int type) {
    return (type == GB_POT_HERNQUIST || type == GB_POT_KEPLER || type == GB_POT_JAFFE || type == GB_POT_KUZMIN) ? 1
         : (type == GB_POT_NFW_SPHERICAL || type == GB_POT_MIYAMOTONAGAI || type == GB_POT_PLUMMER ||
            type == GB_POT_ISOCHRONE || type == GB_POT_SATOH) ? 2
         : (type == GB_POT_NFW_FLATTENED || type == GB_POT_NFW_TRIAXIAL) ? 5
         : (type == GB_POT_MN3) ? 6
         : (type == GB_POT_LONGMURALIBAR) ? 4
         : (type == GB_POT_SCF) ? 2
         : (type == GB_POT_POWERLAWCUTOFF) ? 6
         : (type == GB_POT_LOGARITHMIC) ? 7
         : 0;
}

#ifdef __CUDACC__
#define GB_HD __host__ __device__ inline
#else
#define GB_HD inline
#endif
#include <math.h>
// Derived constants of one component for the fast build (layout: gb_nderived() above; consumers: the accum()
// functions of potentials.cuh).  Pure functions of the parameter vector: evaluated once per call on the host
// (capi.cu:resolve), and per evaluation on the device for a TimeInterpolated component, whose parameters change with t.
GB_HD void gb_derive(int type, const double* p, double* d) {
    switch (type) {
        case GB_POT_HERNQUIST: case GB_POT_KEPLER: case GB_POT_JAFFE: case GB_POT_KUZMIN:
            d[0] = p[0] * p[1]; break;
        case GB_POT_SATOH:
            d[0] = p[0] * p[1]; d[1] = p[3] * p[3]; break;
        case GB_POT_NFW_FLATTENED:      // a = b = 1 (flattenednfw_* ignore them, builtin_potentials.cpp:925-962)
            d[0] = p[0] * p[1]; d[1] = 1. / p[2]; d[2] = 1.; d[3] = 1.; d[4] = 1. / (p[5] * p[5]); break;
        case GB_POT_NFW_TRIAXIAL:
            d[0] = p[0] * p[1]; d[1] = 1. / p[2]; d[2] = 1. / (p[3] * p[3]); d[3] = 1. / (p[4] * p[4]);
            d[4] = 1. / (p[5] * p[5]); break;
        case GB_POT_NFW_SPHERICAL:
            d[0] = p[0] * p[1]; d[1] = 1. / p[2]; break;
        case GB_POT_MIYAMOTONAGAI:
            d[0] = p[0] * p[1]; d[1] = p[3] * p[3]; break;
        case GB_POT_PLUMMER: case GB_POT_ISOCHRONE:
            d[0] = p[0] * p[1]; d[1] = p[2] * p[2]; break;
        case GB_POT_MN3:
            for (int i = 0; i < 3; i++) { d[i] = p[0] * p[1 + 3 * i]; d[3 + i] = p[3 + 3 * i] * p[3 + 3 * i]; }
            break;
        case GB_POT_LONGMURALIBAR:
            d[0] = p[0] * p[1]; d[1] = sin(p[5]); d[2] = cos(p[5]); d[3] = p[4] * p[4]; break;
        case GB_POT_SCF:
            d[0] = p[0] * p[3] / (p[4] * p[4]); d[1] = 1. / p[4]; break;
        case GB_POT_LOGARITHMIC:
            d[0] = p[1] * p[1]; d[1] = p[2] * p[2]; d[2] = 1. / (p[3] * p[3]); d[3] = 1. / (p[4] * p[4]);
            d[4] = 1. / (p[5] * p[5]); d[5] = sin(p[6]); d[6] = cos(p[6]); break;
        case GB_POT_POWERLAWCUTOFF:
            d[0] = p[0] * p[1]; d[1] = lgamma(0.5 * (3. - p[2])); d[2] = 1. / (p[3] * p[3]); d[3] = 1. / p[3];
            d[4] = 3. - p[2]; d[5] = -1.; break;     // d[5] = offset of the fit in ext (resolve()); < 0: no fit, use the series
        default: break;
    }
}


#ifdef __CUDACC__
#define GB_DEV __device__ __forceinline__

// ---- math helpers -------------------------------------------------------------------------
// GB_STRICT=1: IEEE sqrt/div and libm pow/log in the reference's operation order, compiled with
// -fmad=false.  GB_STRICT=0: same formulas, FMA contraction allowed, x^-1.5 via rsqrt,
// reciprocals shared where the reference divides repeatedly.
#ifndef GB_STRICT
#define GB_STRICT 0
#endif

#include "fastmath.cuh"
#if GB_STRICT
GB_DEV double gb_pow_m1p5(double x) { return pow(x, -1.5); }
#endif

GB_DEV double gb_norm3(double x, double y, double z) { return sqrt(x * x + y * y + z * z); }

#endif  // __CUDACC__
