// composite.cuh -- composite-potential dispatch.
//
// Replaces c_gradient / c_potential / c_density (reference potential/potential/src/
// cpotential.cpp:170-287), which loop over components calling host function pointers.
// Two forms:
//   Composite<SIG_GENERIC>  a loop with a warp-uniform switch on the component type id; handles
//                           any component list incl. per-component shift/rotate (cpotential.cpp:
//                           246-277, apply_shift_rotate_N + apply_rotate_T transpose).
//   Composite<SIG_xxx>      a compile-time component list (variadic template): no loop, no switch,
//                           parameter offsets are immediates into the constant bank.
#pragma once
#include "potentials.cuh"
#include "scf.cuh"
#include "multipole.cuh"
#include "timeinterp.cuh"

template <int T> struct PotOf;
template <> struct PotOf<GB_POT_NULL>          { using type = PotNull;          static constexpr int NP = 1; };
template <> struct PotOf<GB_POT_HERNQUIST>     { using type = PotHernquist;     static constexpr int NP = 3; };
template <> struct PotOf<GB_POT_NFW_SPHERICAL> { using type = PotNFWSpherical;  static constexpr int NP = 6; };
template <> struct PotOf<GB_POT_NFW_FLATTENED> { using type = PotNFWFlattened;  static constexpr int NP = 6; };
template <> struct PotOf<GB_POT_NFW_TRIAXIAL>  { using type = PotNFWTriaxial;   static constexpr int NP = 6; };
template <> struct PotOf<GB_POT_MIYAMOTONAGAI> { using type = PotMiyamotoNagai; static constexpr int NP = 4; };
template <> struct PotOf<GB_POT_MN3>           { using type = PotMN3;           static constexpr int NP = 13; };
template <> struct PotOf<GB_POT_LONGMURALIBAR> { using type = PotLongMuraliBar; static constexpr int NP = 6; };
template <> struct PotOf<GB_POT_KEPLER>        { using type = PotKepler;        static constexpr int NP = 2; };
template <> struct PotOf<GB_POT_PLUMMER>       { using type = PotPlummer;       static constexpr int NP = 3; };
template <> struct PotOf<GB_POT_ISOCHRONE>     { using type = PotIsochrone;     static constexpr int NP = 3; };
template <> struct PotOf<GB_POT_JAFFE>         { using type = PotJaffe;         static constexpr int NP = 3; };
template <> struct PotOf<GB_POT_STONE>         { using type = PotStone;         static constexpr int NP = 4; };
template <> struct PotOf<GB_POT_BURKERT>       { using type = PotBurkert;       static constexpr int NP = 3; };
template <> struct PotOf<GB_POT_SATOH>         { using type = PotSatoh;         static constexpr int NP = 4; };
template <> struct PotOf<GB_POT_KUZMIN>        { using type = PotKuzmin;        static constexpr int NP = 3; };
template <> struct PotOf<GB_POT_LOGARITHMIC>   { using type = PotLogarithmic;   static constexpr int NP = 7; };
template <> struct PotOf<GB_POT_LEESUTO>       { using type = PotLeeSuto;       static constexpr int NP = 6; };
template <> struct PotOf<GB_POT_POWERLAWCUTOFF>{ using type = PotPowerLawCutoff;static constexpr int NP = 4; };

// ---- runtime (generic) component dispatch -----------------------------------------------------
#define GB_FOR_EACH_SIMPLE_TYPE(X) \
    X(GB_POT_NULL) X(GB_POT_HERNQUIST) X(GB_POT_NFW_SPHERICAL) X(GB_POT_NFW_FLATTENED) X(GB_POT_NFW_TRIAXIAL) \
    X(GB_POT_MIYAMOTONAGAI) X(GB_POT_MN3) X(GB_POT_LONGMURALIBAR) X(GB_POT_KEPLER) X(GB_POT_PLUMMER) \
    X(GB_POT_ISOCHRONE) X(GB_POT_JAFFE) X(GB_POT_STONE) X(GB_POT_BURKERT) X(GB_POT_SATOH) X(GB_POT_KUZMIN) \
    X(GB_POT_LOGARITHMIC) X(GB_POT_LEESUTO) X(GB_POT_POWERLAWCUTOFF)

// HEAVY = false leaves the basis-function expansions (SCF, multipole: 255 registers, 1.7 KB of stack) out of
// the switch: a kernel instantiated that way keeps the register budget of the analytic potentials.
template <bool HEAVY = true>
GB_DEV void gb_comp_gradient(int type, const double* p, const double* e, double x, double y, double z,
                             double& gx, double& gy, double& gz) {
    switch (type) {
#define X(T) case T: PotOf<T>::type::gradient(p, x, y, z, gx, gy, gz); break;
        GB_FOR_EACH_SIMPLE_TYPE(X)
#undef X
        case GB_POT_SCF: if constexpr (HEAVY) PotSCF::gradient(p, e, x, y, z, gx, gy, gz); break;
        case GB_POT_MULTIPOLE: if constexpr (HEAVY) PotMultipole::gradient(p, e, x, y, z, gx, gy, gz); break;
        default: break;
    }
}
#if !GB_STRICT
template <bool HEAVY = true, class Ctx>
GB_DEV void gb_comp_accum(int type, const double* p, const double* d, const double* e, Ctx& c) {
    switch (type) {
#define X(T) case T: PotOf<T>::type::accum(p, d, c); break;
        GB_FOR_EACH_SIMPLE_TYPE(X)
#undef X
        case GB_POT_SCF: if constexpr (HEAVY) PotSCF::gradient(p, e, c.x, c.y, c.z, c.gx, c.gy, c.gz); break;
        case GB_POT_MULTIPOLE: if constexpr (HEAVY) PotMultipole::gradient(p, e, c.x, c.y, c.z, c.gx, c.gy, c.gz); break;
        default: break;
    }
}
#endif
template <bool HEAVY = true>
GB_DEV double gb_comp_value(int type, const double* p, const double* e, double x, double y, double z) {
    switch (type) {
#define X(T) case T: return PotOf<T>::type::value(p, x, y, z);
        GB_FOR_EACH_SIMPLE_TYPE(X)
#undef X
        case GB_POT_SCF: if constexpr (HEAVY) return PotSCF::value(p, e, x, y, z); else return 0.;
        case GB_POT_MULTIPOLE: if constexpr (HEAVY) return PotMultipole::value(p, e, x, y, z); else return 0.;
        default: return 0.;
    }
}
template <bool HEAVY = true>
GB_DEV double gb_comp_density(int type, const double* p, const double* e, double x, double y, double z) {
    switch (type) {
#define X(T) case T: return PotOf<T>::type::density(p, x, y, z);
        GB_FOR_EACH_SIMPLE_TYPE(X)
#undef X
        case GB_POT_SCF: if constexpr (HEAVY) return PotSCF::density(p, e, x, y, z); else return 0.;
        case GB_POT_MULTIPOLE: if constexpr (HEAVY) return PotMultipole::density(p, e, x, y, z); else return 0.;
        default: return 0.;
    }
}

// shift to the component origin and rotate (cpotential.cpp:151-167 + apply_rotate_T :113-131)
GB_DEV void gb_shift_rotate(const DevComp& c, double x, double y, double z, double& X, double& Y, double& Z) {
    const double sx = x - c.q0[0], sy = y - c.q0[1], sz = z - c.q0[2];
    X = c.R[0] * sx + c.R[1] * sy + c.R[2] * sz;
    Y = c.R[3] * sx + c.R[4] * sy + c.R[5] * sz;
    Z = c.R[6] * sx + c.R[7] * sy + c.R[8] * sz;
}

// ---- TimeInterpolated component (timeinterp.cuh; time_interp_wrapper.cpp:91-250): the wrapped analytic potential
// evaluated with the parameter vector, origin and rotation interpolated at time t; NaN outside the knot range.
GB_DEV void ti_gradient(const double* p, const double* e, double t, double x, double y, double z,
                        double& gx, double& gy, double& gz) {
    const TiView v = ti_view(p, e);
    double wp[GB_TI_MAXPAR], o[3], R[9];
    if (!ti_state(v, t, wp, o, R)) { gx = CUDART_NAN; gy = CUDART_NAN; gz = CUDART_NAN; return; }   // :148-155
    double X, Y, Z, ax = 0., ay = 0., az = 0.;
    ti_to_body(o, R, x, y, z, X, Y, Z);
#if GB_STRICT
    gb_comp_gradient<false>(v.wtype, wp, nullptr, X, Y, Z, ax, ay, az);
#else
    // fast build: the wrapped potential's fused accumulation, its derived constants formed from the interpolated parameters
    double dl[8];
    gb_derive(v.wtype, wp, dl);
    FastCtx<GB_USE_ALL> c2(X, Y, Z);
    gb_comp_accum<false>(v.wtype, wp, dl, nullptr, c2);
    c2.finish(ax, ay, az);
#endif
    // grad += R^T grad_body (time_interp_wrapper.cpp:186-201)
    gx += R[0] * ax + R[3] * ay + R[6] * az;
    gy += R[1] * ax + R[4] * ay + R[7] * az;
    gz += R[2] * ax + R[5] * ay + R[8] * az;
}
// ---- per-step state table (fixed-step integrators): every lane of leapfrog / Ruth4 evaluates the TimeInterpolated
// components at the SAME time t[j], so the interval search, the cubics, the axis-angle -> matrix conversion and
// (fast build) the derived constants are formed ONCE per step and component by a small pre-pass kernel
// (kernels.cu: k_ti_table, one thread per (step, component)) and read back by every lane with warp-uniform loads --
// the same functions on the same inputs as the per-lane path of ti_gradient (which DOP853, with its per-lane stage
// times, keeps), so the results are bit-identical to it.  Row j of the table holds one state per TimeInterpolated
// component, in component order (ti_count(P) states).
struct TiState {
    double wp[GB_TI_MAXPAR];
    double o[3], R[9];
    double dl[8];
    int ok, wtype;
};
GB_HD int ti_count(const DevPot& P) {
    int n = 0;
    for (int i = 0; i < P.n; i++) n += (P.c[i].type == GB_POT_TIMEINTERP);
    return n;
}
GB_DEV void ti_fill_state(const DevPot& P, int i, double t, TiState& s) {
    const DevComp& c = P.c[i];
    const TiView v = ti_view(&P.par[c.poff], P.ext + c.eoff);
    double wp[GB_TI_MAXPAR], o[3], R[9], dl[8];
    for (int k = 0; k < GB_TI_MAXPAR; k++) wp[k] = 0.;
    for (int k = 0; k < 8; k++) dl[k] = 0.;
    for (int k = 0; k < 3; k++) o[k] = 0.;
    for (int k = 0; k < 9; k++) R[k] = 0.;
    const bool ok = ti_state(v, t, wp, o, R);
#if !GB_STRICT
    if (ok) gb_derive(v.wtype, wp, dl);
#endif
    for (int k = 0; k < GB_TI_MAXPAR; k++) s.wp[k] = wp[k];
    for (int k = 0; k < 3; k++) s.o[k] = o[k];
    for (int k = 0; k < 9; k++) s.R[k] = R[k];
    for (int k = 0; k < 8; k++) s.dl[k] = dl[k];
    s.ok = ok ? 1 : 0;
    s.wtype = v.wtype;
}
GB_DEV void ti_gradient_state(const TiState& s, double x, double y, double z, double& gx, double& gy, double& gz) {
    if (!s.ok) { gx = CUDART_NAN; gy = CUDART_NAN; gz = CUDART_NAN; return; }
    double X, Y, Z, ax = 0., ay = 0., az = 0.;
    ti_to_body(s.o, s.R, x, y, z, X, Y, Z);
#if GB_STRICT
    gb_comp_gradient<false>(s.wtype, s.wp, nullptr, X, Y, Z, ax, ay, az);
#else
    FastCtx<GB_USE_ALL> c2(X, Y, Z);
    gb_comp_accum<false>(s.wtype, s.wp, s.dl, nullptr, c2);
    c2.finish(ax, ay, az);
#endif
    gx += s.R[0] * ax + s.R[3] * ay + s.R[6] * az;
    gy += s.R[1] * ax + s.R[4] * ay + s.R[7] * az;
    gz += s.R[2] * ax + s.R[5] * ay + s.R[8] * az;
}
template <int WHAT>     // 0 value, 1 density
GB_DEV double ti_scalar(const double* p, const double* e, double t, double x, double y, double z) {
    const TiView v = ti_view(p, e);
    double wp[GB_TI_MAXPAR], o[3], R[9];
    if (!ti_state(v, t, wp, o, R)) return CUDART_NAN;
    double X, Y, Z;
    ti_to_body(o, R, x, y, z, X, Y, Z);
    return WHAT == 0 ? gb_comp_value<false>(v.wtype, wp, nullptr, X, Y, Z) : gb_comp_density<false>(v.wtype, wp, nullptr, X, Y, Z);
}

template <int SIG> struct Composite;

template <bool HEAVY, bool TI> struct CompositeGeneric {
    // the adaptive integrator evaluates the gradient at 15-17 sites per step: ONE out-of-line copy of this
    // loop-and-switch per kernel instead of 17 inlined ones (hamiltonian.cuh: ham_rhs)
    static constexpr bool kOutOfLineInRhs = true;
    static constexpr bool kTimeDependent = TI;        // may hold a TimeInterpolated component: the kernels pass real times
    // __launch_bounds__ of k_leapfrog: the analytic-only loop is capped at 128 registers (4 CTAs of 128 per SM);
    // uncapped it grew from 124 to 140 registers as more fast accumulators were inlined (48 -> 59 ms for MW2022)
    static constexpr int kFixedStepMaxThreads = HEAVY ? 256 : 128, kFixedStepMinBlocks = HEAVY ? 1 : 4;
    GB_DEV static void gradient(const DevPot& P, double t, double x, double y, double z,
                                double& gx, double& gy, double& gz) {
        gradient_t<false>(P, t, nullptr, x, y, z, gx, gy, gz);
    }
    // the fixed-step integrators: `row` = the TimeInterpolated components' states at this step's time (k_ti_table)
    GB_DEV static void gradient_row(const DevPot& P, const TiState* __restrict__ row, double x, double y, double z,
                                    double& gx, double& gy, double& gz) {
        gradient_t<true>(P, 0., row, x, y, z, gx, gy, gz);
    }
    template <bool CTA_STATE>
    GB_DEV static void gradient_t(const DevPot& P, double t, const TiState* __restrict__ row, double x, double y, double z,
                                  double& gx, double& gy, double& gz) {
        [[maybe_unused]] int k_ti = 0;          // rank of the next TimeInterpolated component = its slot in `row`
#if GB_STRICT
        gx = 0.; gy = 0.; gz = 0.;
        for (int i = 0; i < P.n; i++) {
            const DevComp& c = P.c[i];
            const double* p = &P.par[c.poff];
            const double* e = P.ext + c.eoff;
            if constexpr (TI) {
                if (c.type == GB_POT_TIMEINTERP) {
                    if constexpr (CTA_STATE) ti_gradient_state(row[k_ti++], x, y, z, gx, gy, gz);
                    else ti_gradient(p, e, t, x, y, z, gx, gy, gz);
                    continue;
                }
            }
            if (!c.shift) {
                gb_comp_gradient<HEAVY>(c.type, p, e, x, y, z, gx, gy, gz);
            } else {
                double X, Y, Z, ax = 0., ay = 0., az = 0.;
                gb_shift_rotate(c, x, y, z, X, Y, Z);
                gb_comp_gradient<HEAVY>(c.type, p, e, X, Y, Z, ax, ay, az);
                // rotate back with R^T and accumulate (cpotential.cpp:263-277)
                gx += c.R[0] * ax + c.R[3] * ay + c.R[6] * az;
                gy += c.R[1] * ax + c.R[4] * ay + c.R[7] * az;
                gz += c.R[2] * ax + c.R[5] * ay + c.R[8] * az;
            }
        }
#else
        // fast build: unshifted components share one evaluation context (r, 1/r computed once);
        // a shifted/rotated component gets its own context in its own coordinates.
        FastCtx<GB_USE_ALL> ctx(x, y, z);
        ctx.ext = P.ext;
        for (int i = 0; i < P.n; i++) {
            const DevComp& c = P.c[i];
            const double* p = &P.par[c.poff];
            const double* d = &P.drv[c.doff];
            const double* e = P.ext + c.eoff;
            if constexpr (TI) {
                if (c.type == GB_POT_TIMEINTERP) {
                    if constexpr (CTA_STATE) ti_gradient_state(row[k_ti++], x, y, z, ctx.gx, ctx.gy, ctx.gz);
                    else ti_gradient(p, e, t, x, y, z, ctx.gx, ctx.gy, ctx.gz);
                    continue;
                }
            }
            if (!c.shift) {
                gb_comp_accum<HEAVY>(c.type, p, d, e, ctx);
            } else {
                double X, Y, Z, ax, ay, az;
                gb_shift_rotate(c, x, y, z, X, Y, Z);
                FastCtx<GB_USE_ALL> c2(X, Y, Z);
                c2.ext = P.ext;
                gb_comp_accum<HEAVY>(c.type, p, d, e, c2);
                c2.finish(ax, ay, az);
                ctx.gx += c.R[0] * ax + c.R[3] * ay + c.R[6] * az;
                ctx.gy += c.R[1] * ax + c.R[4] * ay + c.R[7] * az;
                ctx.gz += c.R[2] * ax + c.R[5] * ay + c.R[8] * az;
            }
        }
        ctx.finish(gx, gy, gz);
#endif
    }
    GB_DEV static double value(const DevPot& P, double t, double x, double y, double z) {
        double v = 0.;
        for (int i = 0; i < P.n; i++) {
            const DevComp& c = P.c[i];
            double X = x, Y = y, Z = z;
            if constexpr (TI) { if (c.type == GB_POT_TIMEINTERP) { v = v + ti_scalar<0>(&P.par[c.poff], P.ext + c.eoff, t, x, y, z); continue; } }
            if (c.shift) gb_shift_rotate(c, x, y, z, X, Y, Z);
            v = v + gb_comp_value<HEAVY>(c.type, &P.par[c.poff], P.ext + c.eoff, X, Y, Z);
        }
        return v;
    }
    GB_DEV static double density(const DevPot& P, double t, double x, double y, double z) {
        double v = 0.;
        for (int i = 0; i < P.n; i++) {
            const DevComp& c = P.c[i];
            double X = x, Y = y, Z = z;
            if constexpr (TI) { if (c.type == GB_POT_TIMEINTERP) { v = v + ti_scalar<1>(&P.par[c.poff], P.ext + c.eoff, t, x, y, z); continue; } }
            if (c.shift) gb_shift_rotate(c, x, y, z, X, Y, Z);
            v = v + gb_comp_density<HEAVY>(c.type, &P.par[c.poff], P.ext + c.eoff, X, Y, Z);
        }
        return v;
    }
};
template <> struct Composite<SIG_GENERIC> : CompositeGeneric<true, true> {};
template <> struct Composite<SIG_GENERIC_LIGHT> : CompositeGeneric<false, false> {};
template <> struct Composite<SIG_GENERIC_TI> : CompositeGeneric<false, true> {};

// ---- compile-time component lists ------------------------------------------------------------
template <int OFF, int DOFF, int... Ts> struct SeqImpl;
template <int OFF, int DOFF> struct SeqImpl<OFF, DOFF> {
#if !GB_STRICT
    static constexpr int USE = 0;
    template <class Ctx> GB_DEV static void accum(const DevPot&, Ctx&) {}
#endif
    GB_DEV static void gradient(const DevPot&, double, double, double, double&, double&, double&) {}
    GB_DEV static double value(const DevPot&, double, double, double, double v) { return v; }
    GB_DEV static double density(const DevPot&, double, double, double, double v) { return v; }
};
template <int OFF, int DOFF, int T0, int... Ts> struct SeqImpl<OFF, DOFF, T0, Ts...> {
    using Pt = typename PotOf<T0>::type;
    using Next = SeqImpl<OFF + PotOf<T0>::NP, DOFF + gb_nderived(T0), Ts...>;
#if !GB_STRICT
    static constexpr int USE = Pt::USE | Next::USE;
    template <class Ctx> GB_DEV static void accum(const DevPot& P, Ctx& c) {
        if constexpr (T0 == GB_POT_MN3) PotMN3::template accum_mn3<true>(&P.par[OFF], &P.drv[DOFF], c);   // host verified b1==b2==b3
        else Pt::accum(&P.par[OFF], &P.drv[DOFF], c);
        Next::accum(P, c);
    }
#endif
    GB_DEV static void gradient(const DevPot& P, double x, double y, double z, double& gx, double& gy, double& gz) {
        Pt::gradient(&P.par[OFF], x, y, z, gx, gy, gz);
        Next::gradient(P, x, y, z, gx, gy, gz);
    }
    GB_DEV static double value(const DevPot& P, double x, double y, double z, double v) {
        return Next::value(P, x, y, z, v + Pt::value(&P.par[OFF], x, y, z));
    }
    GB_DEV static double density(const DevPot& P, double x, double y, double z, double v) {
        return Next::density(P, x, y, z, v + Pt::density(&P.par[OFF], x, y, z));
    }
};
template <int... Ts> struct Seq {
    static constexpr bool kOutOfLineInRhs = false;
    static constexpr bool kTimeDependent = false;     // compile-time lists of static analytic potentials ignore t
    static constexpr int kFixedStepMaxThreads = 256, kFixedStepMinBlocks = 1;
    GB_DEV static void gradient(const DevPot& P, double t, double x, double y, double z,
                                double& gx, double& gy, double& gz) {
#if GB_STRICT
        gx = 0.; gy = 0.; gz = 0.;
        SeqImpl<0, 0, Ts...>::gradient(P, x, y, z, gx, gy, gz);
#else
        FastCtx<SeqImpl<0, 0, Ts...>::USE> c(x, y, z);
        c.ext = P.ext;
        SeqImpl<0, 0, Ts...>::accum(P, c);
        c.finish(gx, gy, gz);
#endif
    }
    GB_DEV static double value(const DevPot& P, double t, double x, double y, double z) {
        return SeqImpl<0, 0, Ts...>::value(P, x, y, z, 0.);
    }
    GB_DEV static double density(const DevPot& P, double t, double x, double y, double z) {
        return SeqImpl<0, 0, Ts...>::density(P, x, y, z, 0.);
    }
};

template <> struct Composite<SIG_NFW>        : Seq<GB_POT_NFW_SPHERICAL> {};
template <> struct Composite<SIG_HERNQUIST>  : Seq<GB_POT_HERNQUIST> {};
template <> struct Composite<SIG_MW2022>     : Seq<GB_POT_MN3, GB_POT_HERNQUIST, GB_POT_HERNQUIST, GB_POT_NFW_SPHERICAL> {};
template <> struct Composite<SIG_BAR_MW2022> : Seq<GB_POT_LONGMURALIBAR, GB_POT_MN3, GB_POT_HERNQUIST, GB_POT_HERNQUIST, GB_POT_NFW_SPHERICAL> {};
template <> struct Composite<SIG_MW2022_BAR> : Seq<GB_POT_MN3, GB_POT_HERNQUIST, GB_POT_HERNQUIST, GB_POT_NFW_SPHERICAL, GB_POT_LONGMURALIBAR> {};
template <> struct Composite<SIG_MW_V1>      : Seq<GB_POT_MIYAMOTONAGAI, GB_POT_HERNQUIST, GB_POT_HERNQUIST, GB_POT_NFW_SPHERICAL> {};
template <> struct Composite<SIG_LM10>       : Seq<GB_POT_MIYAMOTONAGAI, GB_POT_HERNQUIST, GB_POT_LOGARITHMIC> {};
template <> struct Composite<SIG_BOVY2014>   : Seq<GB_POT_MIYAMOTONAGAI, GB_POT_POWERLAWCUTOFF, GB_POT_NFW_SPHERICAL> {};
template <> struct Composite<SIG_SCF> {
    static constexpr bool kOutOfLineInRhs = true;
    static constexpr bool kTimeDependent = false;
#ifndef GB_SCF_MINBLOCKS
#define GB_SCF_MAXTHREADS 256
#define GB_SCF_MINBLOCKS 1
#endif
    static constexpr int kFixedStepMaxThreads = GB_SCF_MAXTHREADS, kFixedStepMinBlocks = GB_SCF_MINBLOCKS;
    GB_DEV static void gradient(const DevPot& P, double t, double x, double y, double z, double& gx, double& gy, double& gz) {
#if !GB_STRICT
        if (P.cext_ok) { scf_fast_gradient(P, &P.drv[0], x, y, z, gx, gy, gz); return; }   // warp-uniform
#endif
        gx = 0.; gy = 0.; gz = 0.;
        PotSCF::gradient(&P.par[0], P.ext, x, y, z, gx, gy, gz);
    }
    GB_DEV static double value(const DevPot& P, double t, double x, double y, double z) { return PotSCF::value(&P.par[0], P.ext, x, y, z); }
    GB_DEV static double density(const DevPot& P, double t, double x, double y, double z) { return PotSCF::density(&P.par[0], P.ext, x, y, z); }
};
