// potentials.cuh -- analytic potentials as __device__ functions (gradient / value / density).
//
// Each struct restates one family of the reference's
// potential/potential/builtin/builtin_potentials.cpp (line ranges cited per struct) for ONE
// point held in registers: thread = orbit, so the reference's double6ptr/gradientv SIMD wrapper
// (potential/src/vectorization.h:7-78) has no equivalent here.
// `p` points at the component's [G, ...] parameter vector inside the constant-bank DevPot.
// gradient() ACCUMULATES into (gx,gy,gz) exactly like the reference (grad[k] = grad[k] + ...),
// so the floating-point summation order over components is the reference's.
#pragma once

#define GB_PI 3.14159265358979323846

#if !GB_STRICT
// ---- fast build: fused gradient accumulation ------------------------------------------------------
// Every potential adds its contribution to ONE evaluation context instead of to (gx,gy,gz):
//   spherical terms   grad = F(r) q            -> Fs (direct) or Sir (still to be multiplied by 1/r)
//   axisymmetric disc grad = (Fd x, Fd y, Fdz z)
//   anything else     -> (gx, gy, gz)
// so r, 1/r and the final multiplications by (x,y,z) are done once per evaluation, not once per
// component, and the seeds/refinements of fastmath.cuh replace the library div/sqrt/log.  Same
// formulas as the reference; only the association of the sums and the rounding of the primitives
// differ ("two math modes").  `d` points at the component's host-derived
// constants (gb_nderived()).
enum { GB_USE_SIR = 1, GB_USE_FS = 2, GB_USE_DISC = 4, GB_USE_GEN = 8, GB_USE_ALL = 15 };

template <int USE> struct FastCtx {
    double x, y, z, R2, z2, r2, ir, r;
    double Sir, Fs, Fd, Fdz, gx, gy, gz;
    const double* ext = nullptr;    // DevPot::ext (device-global tables: the PowerLawCutoff fit); set by the composites
    GB_DEV FastCtx(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {
        R2 = fma(y, y, x * x);
        z2 = z * z;
        r2 = R2 + z2;
        ir = gb_rsqrt(r2);      // dead-code-eliminated when no component reads it
        r = r2 * ir;
        Sir = 0.; Fs = 0.; Fd = 0.; Fdz = 0.; gx = 0.; gy = 0.; gz = 0.;
    }
    GB_DEV void finish(double& ox, double& oy, double& oz) const {
        double s = 0.;
        // generic loop (USE == GB_USE_ALL): no component may have touched Sir -- a lone Plummer / MiyamotoNagai / ...
        // evaluated at r = 0, where ir is NaN (0 * inf in the rsqrt refinement) while the reference's value is finite
        if (USE == GB_USE_ALL) s = (Sir != 0.) ? fma(Sir, ir, Fs) : Fs;
        else if (USE & GB_USE_SIR) s = (USE & GB_USE_FS) ? fma(Sir, ir, Fs) : Sir * ir;
        else if (USE & GB_USE_FS) s = Fs;
        constexpr bool SPH = (USE & (GB_USE_SIR | GB_USE_FS)) != 0;
        double fxy, fz;
        if (USE & GB_USE_DISC) { fxy = SPH ? s + Fd : Fd; fz = SPH ? s + Fdz : Fdz; }
        else { fxy = s; fz = s; }
        if (USE & GB_USE_GEN) { ox = fma(fxy, x, gx); oy = fma(fxy, y, gy); oz = fma(fz, z, gz); }
        else { ox = fxy * x; oy = fxy * y; oz = fz * z; }
    }
};
#define GB_ACCUM_VIA_GRADIENT                                                                      \
    static constexpr int USE = GB_USE_GEN;                                                         \
    template <class Ctx> GB_DEV static void accum(const double* p, const double*, Ctx& c) {        \
        gradient(p, c.x, c.y, c.z, c.gx, c.gy, c.gz);                                              \
    }
#else
#define GB_ACCUM_VIA_GRADIENT
#endif

// ---- Null (builtin_potentials.cpp:18-21) -----------------------------------------------------
struct PotNull {
    GB_DEV static void gradient(const double*, double, double, double, double&, double&, double&) {}
    GB_DEV static double value(const double*, double, double, double) { return 0.; }
    GB_DEV static double density(const double*, double, double, double) { return 0.; }
#if !GB_STRICT
    static constexpr int USE = 0;
    template <class Ctx> GB_DEV static void accum(const double*, const double*, Ctx&) {}
#endif
};

// ---- Kepler (builtin_potentials.cpp:56-85): [G, m] -------------------------------------------
struct PotKepler {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double r = gb_norm3(x, y, z);
#if GB_STRICT
        const double fac = p[0] * p[1] / pow(r, 3.);
#else
        const double fac = p[0] * p[1] / (r * r * r);
#endif
        gx = gx + fac * x; gy = gy + fac * y; gz = gz + fac * z;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) { return -p[0] * p[1] / gb_norm3(x, y, z); }
    GB_DEV static double density(const double*, double x, double y, double z) {
        return (x * x + y * y + z * z == 0.) ? CUDART_INF : 0.;
    }
#if !GB_STRICT
    static constexpr int USE = GB_USE_SIR;
    template <class Ctx> GB_DEV static void accum(const double*, const double* d, Ctx& c) {
        c.Sir = fma(d[0], c.ir * c.ir, c.Sir);                  // G m / r^3
    }
#endif
};

// ---- Isochrone (builtin_potentials.cpp:128-160): [G, m, b] -----------------------------------
struct PotIsochrone {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double s = sqrt((x * x + y * y + z * z) + p[2] * p[2]);
        const double denom = s * (s + p[2]) * (s + p[2]);
        const double fac = p[0] * p[1] / denom;
        gx = gx + fac * x; gy = gy + fac * y; gz = gz + fac * z;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        const double r2 = x * x + y * y + z * z;
        return -p[0] * p[1] / (sqrt(r2 + p[2] * p[2]) + p[2]);
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double b = p[2];
        const double r2 = x * x + y * y + z * z;
        const double a = sqrt(b * b + r2);
        return p[1] * (3 * (b + a) * a * a - r2 * (b + 3 * a)) / (4 * GB_PI * pow(b + a, 3.) * a * a * a);
    }
#if !GB_STRICT
    static constexpr int USE = GB_USE_FS;
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        const double s2 = c.r2 + d[1];
        const double is = gb_rsqrt(s2);
        const double isb = gb_rcp(fma(s2, is, p[2]));           // 1 / (s + b)
        c.Fs = fma(d[0] * is, isb * isb, c.Fs);                  // G m / (s (s+b)^2)
    }
#endif
};

// ---- Hernquist (builtin_potentials.cpp:211-244): [G, m, c] -----------------------------------
struct PotHernquist {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double r = gb_norm3(x, y, z);
        const double fac = p[0] * p[1] / ((r + p[2]) * (r + p[2]) * r);
        gx = gx + fac * x; gy = gy + fac * y; gz = gz + fac * z;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        return -p[0] * p[1] / (gb_norm3(x, y, z) + p[2]);
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double r = gb_norm3(x, y, z);
        const double rho0 = p[1] / (2 * GB_PI * p[2] * p[2] * p[2]);
        return rho0 / ((r / p[2]) * pow(1 + r / p[2], 3.));
    }
#if !GB_STRICT
    static constexpr int USE = GB_USE_SIR;
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        const double ia = gb_rcp(c.r + p[2]);
        c.Sir = fma(d[0], ia * ia, c.Sir);                      // G m / ((r+c)^2 r)
    }
#endif
};

// ---- Plummer (builtin_potentials.cpp:292-320): [G, m, b] -------------------------------------
struct PotPlummer {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double R2b = (x * x + y * y + z * z) + p[2] * p[2];
        const double fac = p[0] * p[1] / sqrt(R2b) / R2b;
        gx = gx + fac * x; gy = gy + fac * y; gz = gz + fac * z;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        return -p[0] * p[1] / sqrt((x * x + y * y + z * z) + p[2] * p[2]);
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double r2 = x * x + y * y + z * z;
        return 3 * p[1] / (4 * GB_PI * p[2] * p[2] * p[2]) * pow(1 + r2 / (p[2] * p[2]), -2.5);
    }
#if !GB_STRICT
    static constexpr int USE = GB_USE_FS;
    template <class Ctx> GB_DEV static void accum(const double*, const double* d, Ctx& c) {
        c.Fs = fma(d[0], gb_pow_m1p5(c.r2 + d[1]), c.Fs);       // G m (r^2+b^2)^-3/2
    }
#endif
};

// ---- Jaffe (builtin_potentials.cpp:365-393): [G, m, c] ---------------------------------------
struct PotJaffe {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double r = gb_norm3(x, y, z);
        const double fac = p[0] * p[1] / p[2] * (p[2] / (r * (p[2] + r))) / r;
        gx = gx + fac * x; gy = gy + fac * y; gz = gz + fac * z;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        const double r = gb_norm3(x, y, z);
        return -p[0] * p[1] / p[2] * log(1 + p[2] / r);
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double r = gb_norm3(x, y, z);
        const double rho0 = p[1] / (4 * GB_PI * p[2] * p[2] * p[2]);
        const double u = r / p[2];
        return rho0 / ((u * u) * ((1 + u) * (1 + u)));
    }
#if !GB_STRICT
    static constexpr int USE = GB_USE_SIR;
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        c.Sir = fma(d[0] * c.ir, gb_rcp(p[2] + c.r), c.Sir);    // G m / (r^2 (c+r))
    }
#endif
};

// ---- NFW family (builtin_potentials.cpp:819-864 spherical, 925-962 flattened, 1028-1072
//      triaxial): [G, m, r_s, a, b, c] ------------------------------------------------------------
GB_DEV double gb_nfw_fac(const double* p, double u) {
    const double v_h2 = p[0] * p[1] / p[2];
    return v_h2 / (u * u * u) / (p[2] * p[2]) * (log(1 + u) - u / (1 + u));
}
GB_DEV double gb_nfw_value(const double* p, double u) {
    const double v_h2 = -p[0] * p[1] / p[2];
    return (u == 0) ? v_h2 : v_h2 * log(1 + u) / u;
}
struct PotNFWSpherical {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double u = gb_norm3(x, y, z) / p[2];
        const double fac = gb_nfw_fac(p, u);
        gx = gx + fac * x; gy = gy + fac * y; gz = gz + fac * z;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        return gb_nfw_value(p, gb_norm3(x, y, z) / p[2]);
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double v_h2 = p[0] * p[1] / p[2];
        const double r = gb_norm3(x, y, z);
        const double rho0 = v_h2 / (4 * GB_PI * p[0] * p[2] * p[2]);
        const double u = r / p[2];
        return rho0 / (u * ((1 + u) * (1 + u)));
    }
#if !GB_STRICT
    static constexpr int USE = GB_USE_SIR;
    // G m [ln(1+u) - u/(1+u)] / r^3, u = r/r_s (same expression as gb_nfw_fac, v_h2/(u^3 r_s^2) = G m/r^3)
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        const double L = gb_log(fma(c.r, d[1], 1.0));
        const double h = fma(-c.r, gb_rcp(c.r + p[2]), L);
        c.Sir = fma(d[0], (c.ir * c.ir) * h, c.Sir);
    }
#endif
};
#if !GB_STRICT
// flattened / triaxial NFW, fast build: d = [G m, 1/r_s, 1/a^2, 1/b^2, 1/c^2]; with m the ellipsoidal radius,
// grad_i = G m [ln(1+m/r_s) - m/(m+r_s)] / m^3 * x_i / a_i^2 (the same expression as gb_nfw_fac)
template <class Ctx> GB_DEV void gb_nfw_ellipsoidal_accum(const double* p, const double* d, Ctx& c) {
    const double xi = c.x * d[2], yi = c.y * d[3], zi = c.z * d[4];
    const double m2 = fma(c.x, xi, fma(c.y, yi, c.z * zi));
    const double im = gb_rsqrt(m2);
    const double m = m2 * im;
    const double L = gb_log(fma(m, d[1], 1.0));
    const double h = fma(-m, gb_rcp(m + p[2]), L);
    const double f = (d[0] * h) * (im * (im * im));
    c.gx = fma(f, xi, c.gx); c.gy = fma(f, yi, c.gy); c.gz = fma(f, zi, c.gz);
}
#endif
struct PotNFWFlattened {
    GB_DEV static double u_of(const double* p, double x, double y, double z) {
        return sqrt(x * x + y * y + z * z / (p[5] * p[5])) / p[2];
    }
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double fac = gb_nfw_fac(p, u_of(p, x, y, z));
        gx = gx + fac * x; gy = gy + fac * y; gz = gz + fac * z / (p[5] * p[5]);
    }
    GB_DEV static double value(const double* p, double x, double y, double z) { return gb_nfw_value(p, u_of(p, x, y, z)); }
    GB_DEV static double density(const double*, double, double, double) { return CUDART_NAN; }  // nan_density (cybuiltin.pyx:303-311)
    #if !GB_STRICT
    static constexpr int USE = GB_USE_GEN;
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) { gb_nfw_ellipsoidal_accum(p, d, c); }
#endif
};
struct PotNFWTriaxial {
    GB_DEV static double u_of(const double* p, double x, double y, double z) {
        return sqrt(x * x / (p[3] * p[3]) + y * y / (p[4] * p[4]) + z * z / (p[5] * p[5])) / p[2];
    }
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double fac = gb_nfw_fac(p, u_of(p, x, y, z));
        gx = gx + fac * x / (p[3] * p[3]); gy = gy + fac * y / (p[4] * p[4]); gz = gz + fac * z / (p[5] * p[5]);
    }
    GB_DEV static double value(const double* p, double x, double y, double z) { return gb_nfw_value(p, u_of(p, x, y, z)); }
    GB_DEV static double density(const double*, double, double, double) { return CUDART_NAN; }
    #if !GB_STRICT
    static constexpr int USE = GB_USE_GEN;
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) { gb_nfw_ellipsoidal_accum(p, d, c); }
#endif
};

// ---- Miyamoto-Nagai (builtin_potentials.cpp:1287-1332): [G, m, a, b] --------------------------
GB_DEV void gb_mn_gradient(double G, double m, double a, double b, double x, double y, double z,
                           double& gx, double& gy, double& gz) {
    const double sqrtz = sqrt(z * z + b * b);
    const double zd = a + sqrtz;
    const double fac = G * m * gb_pow_m1p5(x * x + y * y + zd * zd);
    gx = gx + fac * x;
    gy = gy + fac * y;
    gz = gz + fac * z * (1. + a / sqrtz);
}
GB_DEV double gb_mn_value(double G, double m, double a, double b, double x, double y, double z) {
    const double zd = (a + sqrt(z * z + b * b));
    return -G * m / sqrt(x * x + y * y + zd * zd);
}
GB_DEV double gb_mn_density(double M, double a, double b, double x, double y, double z) {
    const double R2 = x * x + y * y;
    const double sqrt_zb = sqrt(z * z + b * b);
    const double numer = (b * b * M / (4 * GB_PI)) * (a * R2 + (a + 3 * sqrt_zb) * (a + sqrt_zb) * (a + sqrt_zb));
    const double denom = pow(R2 + (a + sqrt_zb) * (a + sqrt_zb), 2.5) * sqrt_zb * sqrt_zb * sqrt_zb;
    return numer / denom;
}
#if !GB_STRICT
// one Miyamoto-Nagai disc given sq = sqrt(z^2+b^2) and isq = 1/sq: 12 FP64 instructions
template <class Ctx> GB_DEV void gb_mn_accum(double Gm, double a, double sq, double isq, Ctx& c) {
    const double zd = a + sq;
    const double f = gb_pow_m1p5(fma(zd, zd, c.R2));
    c.Fd = fma(Gm, f, c.Fd);
    c.Fdz = fma(Gm, f * fma(a, isq, 1.0), c.Fdz);
}
#endif
struct PotMiyamotoNagai {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        gb_mn_gradient(p[0], p[1], p[2], p[3], x, y, z, gx, gy, gz);
    }
    GB_DEV static double value(const double* p, double x, double y, double z) { return gb_mn_value(p[0], p[1], p[2], p[3], x, y, z); }
    GB_DEV static double density(const double* p, double x, double y, double z) { return gb_mn_density(p[1], p[2], p[3], x, y, z); }
#if !GB_STRICT
    static constexpr int USE = GB_USE_DISC;
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        const double s2 = c.z2 + d[1];
        const double isq = gb_rsqrt(s2);
        gb_mn_accum(d[0], p[2], s2 * isq, isq, c);
    }
#endif
};

// ---- MN3 exponential disk (builtin_potentials.cpp:1390-1441): [G, m1,a1,b1, m2,a2,b2, m3,a3,b3, ...]
struct PotMN3 {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
#pragma unroll
        for (int i = 0; i < 3; i++) gb_mn_gradient(p[0], p[1 + 3 * i], p[2 + 3 * i], p[3 + 3 * i], x, y, z, gx, gy, gz);
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        double val = 0.;
#pragma unroll
        for (int i = 0; i < 3; i++) val += gb_mn_value(p[0], p[1 + 3 * i], p[2 + 3 * i], p[3 + 3 * i], x, y, z);
        return val;
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        double val = 0.;
#pragma unroll
        for (int i = 0; i < 3; i++) val += gb_mn_density(p[1 + 3 * i], p[2 + 3 * i], p[3 + 3 * i], x, y, z);
        return val;
    }
#if !GB_STRICT
    static constexpr int USE = GB_USE_DISC;
    // SHARED_B: the three discs have the same b (always true for MN3ExponentialDiskPotential,
    // builtin/core.py:658-664; the host verifies it before resolving a compile-time signature), so
    // sqrt(z^2+b^2) is evaluated once.
    template <bool SHARED_B, class Ctx> GB_DEV static void accum_mn3(const double* p, const double* d, Ctx& c) {
        double sq = 0., isq = 0.;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (i == 0 || !SHARED_B) {
                const double s2 = c.z2 + d[3 + i];
                isq = gb_rsqrt(s2);
                sq = s2 * isq;
            }
            gb_mn_accum(d[i], p[2 + 3 * i], sq, isq, c);
        }
    }
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) { accum_mn3<false>(p, d, c); }
#endif
};

// ---- Long & Murali bar (builtin_potentials.cpp:1681-1739): [G, m, a, b, c, alpha] --------------
// The reference evaluates sin/cos(alpha) six times per call; they are functions of a parameter,
// so they are computed once per call here (same libm-class result each time, same products).
struct PotLongMuraliBar {
    GB_DEV static void gradient(const double* p, double qx, double qy, double qz, double& gx_, double& gy_, double& gz_) {
        double sa, ca;
        sincos(p[5], &sa, &ca);
        const double x = qx * ca + qy * sa;
        const double y = -qx * sa + qy * ca;
        const double z = qz;
        const double a = p[2], b = p[3], c = p[4];
        const double bcz = b + sqrt(c * c + z * z);
        const double Tm = sqrt((a - x) * (a - x) + y * y + bcz * bcz);
        const double Tp = sqrt((a + x) * (a + x) + y * y + bcz * bcz);
        const double fac1 = p[0] * p[1] / (2 * Tm * Tp);
        const double fac2 = 1 / (y * y + bcz * bcz);
        const double fac3 = Tp + Tm - (4 * x * x) / (Tp + Tm);
        const double gx = 4 * fac1 * x / (Tp + Tm);
        const double gy = fac1 * y * fac2 * fac3;
        const double gz = fac1 * z * fac2 * fac3 * bcz / sqrt(z * z + c * c);
        gx_ = gx_ + (gx * ca - gy * sa);
        gy_ = gy_ + (gx * sa + gy * ca);
        gz_ = gz_ + gz;
    }
    GB_DEV static double value(const double* p, double qx, double qy, double qz) {
        double sa, ca;
        sincos(p[5], &sa, &ca);
        const double x = qx * ca + qy * sa;
        const double y = -qx * sa + qy * ca;
        const double z = qz;
        const double a = p[2], b = p[3], c = p[4];
        const double bcz = b + sqrt(c * c + z * z);
        const double Tm = sqrt((a - x) * (a - x) + y * y + bcz * bcz);
        const double Tp = sqrt((a + x) * (a + x) + y * y + bcz * bcz);
        return p[0] * p[1] / (2 * a) * log((x - a + Tm) / (x + a + Tp));
    }
#if !GB_STRICT
    static constexpr int USE = GB_USE_GEN;
    // same closed form as gradient() above with sin/cos(alpha) and c^2 from the host-derived block,
    // 3 rsqrt + 2 rcp seeds instead of 5 sqrt + 5 div
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        const double sa = d[1], ca = d[2];
        const double x = fma(c.y, sa, c.x * ca);
        const double y = fma(c.y, ca, -(c.x * sa));
        const double a = p[2], b = p[3];
        const double zc2 = c.z2 + d[3];
        const double izc = gb_rsqrt(zc2);
        const double bcz = fma(zc2, izc, b);
        const double S = fma(bcz, bcz, y * y);
        const double am = a - x, ap = a + x;
        const double Tm2 = fma(am, am, S), Tp2 = fma(ap, ap, S);
        const double iTm = gb_rsqrt(Tm2), iTp = gb_rsqrt(Tp2);
        const double T = fma(Tm2, iTm, Tp2 * iTp);               // Tp + Tm
        const double iT = gb_rcp(T);
        const double fac1 = (0.5 * d[0]) * (iTm * iTp);
        const double fac2 = gb_rcp(S);
        const double x4 = 4. * x;
        const double fac3 = fma(-(x4 * x), iT, T);
        const double gx = fac1 * (x4 * iT);
        const double f123 = fac1 * (fac2 * fac3);
        const double gy = f123 * y;
        c.gx += fma(gx, ca, -(gy * sa));
        c.gy += fma(gx, sa, gy * ca);
        c.gz = fma(f123 * c.z, bcz * izc, c.gz);
    }
#endif
    // Density: the reference uses a sympy-generated expression (builtin_potentials.cpp:1741-1811);
    // here it is derived independently as rho = Laplacian(Phi)/(4 pi G) from the closed-form second
    // derivatives of Phi = GM/(2a) * ln((x-a+Tm)/(x+a+Tp)); see DESIGN.md.
    GB_DEV static double density(const double* p, double qx, double qy, double qz) {
        double sa, ca;
        sincos(p[5], &sa, &ca);
        const double x = qx * ca + qy * sa;
        const double y = -qx * sa + qy * ca;
        const double z = qz;
        const double a = p[2], b = p[3], c = p[4];
        const double zc = sqrt(c * c + z * z);
        const double B = b + zc;                    // bcz
        const double s2 = y * y + B * B;            // y^2 + bcz^2
        const double Tm = sqrt((a - x) * (a - x) + s2);
        const double Tp = sqrt((a + x) * (a + x) + s2);
        // Phi = K [ln(um) - ln(up)], um = x-a+Tm, up = x+a+Tp.
        // d/dx ln(um) = 1/Tm ; d/dx ln(up) = 1/Tp  => Phi_xx = K[(a-x)/Tm^3 + (a+x)/Tp^3]
        // For w in {y, B}: d ln(u)/dw = w/(T u); second derivative below.
        const double um = x - a + Tm, up = x + a + Tp;
        const double Pxx = (a - x) / (Tm * Tm * Tm) + (a + x) / (Tp * Tp * Tp);
        // f(T,u) = 1/(T u);  d2 ln u / dw2 = f - w^2 (1/(T^3 u) + 1/(T^2 u^2))
        const double fm = 1. / (Tm * um), fp = 1. / (Tp * up);
        const double hm = fm / (Tm * Tm) + fm * fm, hp = fp / (Tp * Tp) + fp * fp;
        const double Pyy = (fm - y * y * hm) - (fp - y * y * hp);
        const double PBB = (fm - B * B * hm) - (fp - B * B * hp);
        const double PB = B * (fm - fp);
        // z enters through B(z): B' = z/zc, B'' = c^2/zc^3
        const double Bp = z / zc, Bpp = c * c / (zc * zc * zc);
        const double Pzz = PBB * Bp * Bp + PB * Bpp;
        return p[1] / (8. * GB_PI * a) * (Pxx + Pyy + Pzz);
    }
};

// ================================================================================================
// Remaining analytic builtins.  All of them accumulate through the generic
// (gx,gy,gz) path in the fast build; strict mode mirrors the reference's statements.
// ================================================================================================

// ---- Stone & Ostriker 2015 (builtin_potentials.cpp:662-719): [G, m, r_c, r_h] ------------------
struct PotStone {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double r = gb_norm3(x, y, z);
        const double u_c = r / p[2];
        const double u_h = r / p[3];
        const double fac = 2 * p[0] * p[1] / (GB_PI * r * r * r) / (p[2] - p[3]);
        const double dphi_dr = fac * (p[2] * atan(u_c) - p[3] * atan(u_h));
        gx = gx + dphi_dr * x; gy = gy + dphi_dr * y; gz = gz + dphi_dr * z;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        const double r = gb_norm3(x, y, z);
        const double u_c = r / p[2];
        const double u_h = r / p[3];
        const double fac = 2 * p[0] * p[1] / GB_PI / (p[3] - p[2]);
        if (r == 0) return -fac * 0.5 * log(p[3] * p[3] / (p[2] * p[2]));
        return -fac * (atan(u_h) / u_h - atan(u_c) / u_c + 0.5 * log((r * r + p[3] * p[3]) / (r * r + p[2] * p[2])));
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double r = gb_norm3(x, y, z);
        const double rho = p[1] * (p[2] + p[3]) / (2 * GB_PI * GB_PI * p[2] * p[2] * p[3] * p[3]);
        const double u_c = r / p[2];
        const double u_t = r / p[3];
        return rho / ((1 + u_c * u_c) * (1 + u_t * u_t));
    }
    GB_ACCUM_VIA_GRADIENT
};

// ---- Burkert (builtin_potentials.cpp:2238-2279): [G, rho, r0] ----------------------------------
struct PotBurkert {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double r = gb_norm3(x, y, z);
        const double u = r / p[2];
        const double dphi_dr = -GB_PI * p[0] * p[1] * p[2] / (u * u) * (2 * atan(u) - 2 * log(1 + u) - log(1 + u * u));
        gx = gx + dphi_dr * x / r; gy = gy + dphi_dr * y / r; gz = gz + dphi_dr * z / r;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        const double r = gb_norm3(x, y, z);
        const double u = r / p[2];
        return -GB_PI * p[0] * p[1] * p[2] * p[2] *
               (GB_PI - 2 * (1 + 1 / u) * atan(u) + 2 * (1 + 1 / u) * log(1 + u) - (1 - 1 / u) * log(1 + u * u));
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double u = gb_norm3(x, y, z) / p[2];
        return p[1] / ((1 + u) * (1 + u * u));
    }
    GB_ACCUM_VIA_GRADIENT
};

// ---- Satoh (builtin_potentials.cpp:1147-1186): [G, m, a, b] ------------------------------------
struct PotSatoh {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double zb = sqrt(z * z + p[3] * p[3]);
        const double S2 = (x * x + y * y + z * z) + p[2] * (p[2] + 2 * zb);
        const double dPhi_dS = p[0] * p[1] / S2;
        const double S = sqrt(S2);
        gx = gx + dPhi_dS * x / S;
        gy = gy + dPhi_dS * y / S;
        gz = gz + dPhi_dS / S * z * (1 + p[2] / zb);
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        const double S2 = (x * x + y * y + z * z) + p[2] * (p[2] + 2 * sqrt(z * z + p[3] * p[3]));
        return -p[0] * p[1] / sqrt(S2);
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double z2b2 = z * z + p[3] * p[3];
        const double xyz2 = x * x + y * y + z * z;
        const double S2 = xyz2 + p[2] * (p[2] + 2 * sqrt(z2b2));
        const double A = p[1] * p[2] * p[3] * p[3] / (4 * GB_PI * S2 * sqrt(S2) * z2b2);
        return A * (1 / sqrt(z2b2) + 3 / p[2] * (1 - xyz2 / S2));
    }
    #if !GB_STRICT
    static constexpr int USE = GB_USE_DISC;
    // d = [G m, b^2]: grad = G m S^-3 (x, y, z (1 + a / sqrt(z^2+b^2))), S^2 = r^2 + a (a + 2 sqrt(z^2+b^2))
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        const double s2 = c.z2 + d[1];
        const double isq = gb_rsqrt(s2);
        const double S2 = fma(p[2], fma(2., s2 * isq, p[2]), c.r2);
        const double f = d[0] * gb_pow_m1p5(S2);
        c.Fd += f;
        c.Fdz = fma(f, fma(p[2], isq, 1.0), c.Fdz);
    }
#endif
};

// ---- Kuzmin disc (builtin_potentials.cpp:1235-1283): [G, m, a] ---------------------------------
struct PotKuzmin {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double az = p[2] + fabs(z);
        const double S2 = x * x + y * y + az * az;
        const double fac = p[0] * p[1] * gb_pow_m1p5(S2);
        const double zsign = (z > 0) ? 1. : ((z < 0) ? -1. : 0.);
        gx = gx + fac * x; gy = gy + fac * y; gz = gz + fac * zsign * az;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        const double az = p[2] + fabs(z);
        return -p[0] * p[1] / sqrt(x * x + y * y + az * az);
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        if (z != 0.) return 0.;
        return p[1] * p[2] / (2 * GB_PI) * pow(x * x + y * y + p[2] * p[2], -1.5);
    }
    #if !GB_STRICT
    static constexpr int USE = GB_USE_DISC | GB_USE_GEN;
    // d = [G m]: grad = G m S^-3 (x, y, sign(z) (a + |z|)), S^2 = R^2 + (a + |z|)^2
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        const double az = p[2] + fabs(c.z);
        const double f = d[0] * gb_pow_m1p5(fma(az, az, c.R2));
        const double zs = (c.z > 0) ? 1. : ((c.z < 0) ? -1. : 0.);
        c.Fd += f;
        c.gz = fma(f * zs, az, c.gz);          // the z-component is not proportional to z: general accumulator
    }
#endif
};

// ---- Logarithmic, triaxial, rotated by phi about z (builtin_potentials.cpp:1560-1625):
//      [G, v_c, r_h, q1, q2, q3, phi] ------------------------------------------------------------
struct PotLogarithmic {
    GB_DEV static void gradient(const double* p, double qx, double qy, double qz, double& gx, double& gy, double& gz) {
        double sp, cp;
        sincos(p[6], &sp, &cp);
        const double x = qx * cp + qy * sp;
        const double y = -qx * sp + qy * cp;
        const double z = qz;
        const double fac = p[1] * p[1] / (p[2] * p[2] + x * x / (p[3] * p[3]) + y * y / (p[4] * p[4]) + z * z / (p[5] * p[5]));
        const double ax = fac * x / (p[3] * p[3]);
        const double ay = fac * y / (p[4] * p[4]);
        const double az = fac * z / (p[5] * p[5]);
        gx = gx + (ax * cp - ay * sp);
        gy = gy + (ax * sp + ay * cp);
        gz = gz + az;
    }
    GB_DEV static double value(const double* p, double qx, double qy, double qz) {
        double sp, cp;
        sincos(p[6], &sp, &cp);
        const double x = qx * cp + qy * sp;
        const double y = -qx * sp + qy * cp;
        return 0.5 * p[1] * p[1] * log(p[2] * p[2] + x * x / (p[3] * p[3]) + y * y / (p[4] * p[4]) + qz * qz / (p[5] * p[5]));
    }
    // the reference's density ignores phi (it uses q directly, :1579-1601); reproduced as coded
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double q1s = p[3] * p[3], q2s = p[4] * p[4], q3s = p[5] * p[5];
        const double t2 = q1s * q2s;
        const double t3 = t2 * (z * z);
        const double t5 = q1s * q3s;
        const double t6 = t5 * (y * y);
        const double t7 = q2s * q3s;
        const double t8 = t7 * (x * x);
        const double t9 = p[2] * p[2] * t2 * q3s;
        const double t10 = t6 + t8 + t9;
        const double t11 = t3 + t9;
        const double den = t10 + t3;
        return p[1] * p[1] * (t2 * (t10 - t3) + t5 * (t11 - t6 + t8) + t7 * (t11 + t6 - t8)) / (den * den) / (4 * GB_PI * p[0]);
    }
    #if !GB_STRICT
    static constexpr int USE = GB_USE_GEN;
    // d = [v_c^2, r_h^2, 1/q1^2, 1/q2^2, 1/q3^2, sin(phi), cos(phi)]: one reciprocal instead of seven divisions
    template <class Ctx> GB_DEV static void accum(const double*, const double* d, Ctx& c) {
        const double sp = d[5], cp = d[6];
        const double X = fma(c.y, sp, c.x * cp), Y = fma(c.y, cp, -(c.x * sp));
        const double Xi = X * d[2], Yi = Y * d[3], zi = c.z * d[4];
        const double f = d[0] * gb_rcp(fma(X, Xi, fma(Y, Yi, fma(c.z, zi, d[1]))));
        const double ax = f * Xi, ay = f * Yi;
        c.gx += fma(ax, cp, -(ay * sp));
        c.gy += fma(ax, sp, ay * cp);
        c.gz = fma(f, zi, c.gz);
    }
#endif
};

// ---- Lee & Suto 2003 triaxial NFW (builtin_potentials.cpp:1446-1558): [G, v_c, r_s, a, b, c] ----
struct PotLeeSuto {
    GB_DEV static double vh2(const double* p, double& e_b2, double& e_c2) {
        const double ba = p[4] / p[3], ca = p[5] / p[3];
        e_b2 = 1 - ba * ba;
        e_c2 = 1 - ca * ca;
        const double ln2 = 0.6931471805599453;
        return p[1] * p[1] / (ln2 - 0.5 + (ln2 - 0.75) * e_b2 + (ln2 - 0.75) * e_c2);
    }
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        double e_b2, e_c2;
        const double v_h2 = vh2(p, e_b2, e_c2);
        const double rs = p[2];
        const double r2 = x * x + y * y + z * z;
        const double r = sqrt(r2);
        const double r4 = r2 * r2;
        const double x0 = r + rs;
        const double x1 = x0 * x0;
        const double x2 = v_h2 / (12. * r4 * r2 * r * x1);
        const double x10 = log(x0 / rs);
        const double x13 = r * 3. * rs;
        const double x15 = x13 - r2;
        const double x16 = x15 + 6. * (rs * rs);
        const double x17 = 6. * rs * x0 * (r * x16 - x0 * x10 * 6. * (rs * rs));
        const double x20 = x0 * r2;
        const double x21 = 2. * r * x0;
        const double x7 = e_b2 * y * y + e_c2 * z * z;
        const double x22 = -12. * r4 * r * rs * x0 + 12. * r4 * rs * x1 * x10 +
                           3. * rs * x7 * (x16 * r2 - 18. * x1 * x10 * (rs * rs) + x20 * (2. * r - 3. * rs) + x21 * (x15 + 9. * (rs * rs))) -
                           x20 * (e_b2 + e_c2) * (-6. * r * rs * (r2 - (rs * rs)) + 6. * rs * x0 * x10 * (r2 - 3. * (rs * rs)) +
                                                  x20 * (-4. * r + 3. * rs) + x21 * (-x13 + 2. * r2 + 6. * (rs * rs)));
        gx = gx + x2 * x * (x17 * x7 + x22);
        gy = gy + x2 * y * (x17 * (x7 - r2 * e_b2) + x22);
        gz = gz + x2 * z * (x17 * (x7 - r2 * e_c2) + x22);
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        double e_b2, e_c2;
        const double phi0 = vh2(p, e_b2, e_c2);
        const double r = sqrt(x * x + y * y + z * z);
        const double u = r / p[2];
        if (u == 0) return phi0;
        const double l1u = log(1 + u);
        const double F1 = -l1u / u;
        const double F2 = -1 / 3. + (2 * u * u - 3 * u + 6) / (6 * u * u) + (1 / u - pow(u, -3.)) * l1u;
        const double F3 = (u * u - 3 * u - 6) / (2 * u * u * (1 + u)) + 3 * pow(u, -3.) * l1u;
        const double costh2 = z * z / (r * r);
        const double sinth2 = 1 - costh2;
        const double sinph2 = y * y / (x * x + y * y);
        return phi0 * (F1 + (e_b2 + e_c2) / 2. * F2 + (e_b2 * sinth2 * sinph2 + e_c2 * costh2) / 2. * F3);
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double b_a2 = p[4] * p[4] / (p[3] * p[3]);
        const double c_a2 = p[5] * p[5] / (p[3] * p[3]);
        double e_b2, e_c2;
        const double v_h2 = vh2(p, e_b2, e_c2);
        const double u = sqrt(x * x + y * y / b_a2 + z * z / c_a2) / p[2];
        return v_h2 / (u * (1 + u) * (1 + u)) / (4. * GB_PI * p[2] * p[2] * p[0]);
    }
    GB_ACCUM_VIA_GRADIENT
};

// ---- Power law with exponential cutoff (builtin_potentials.cpp:465-554): [G, m, alpha, r_c] -----
// The reference calls GSL (gsl_sf_gamma_inc_P, gsl_sf_gamma; system library, absent here).  The
// regularised lower incomplete gamma function P(a,x) is evaluated from its published series
// (x < a+1) and Lentz continued fraction (otherwise), both to double-precision convergence.
GB_DEV double gb_gamma_inc_P(double a, double x) {
    if (!(x > 0.)) return 0.;
    const double lead = exp(a * log(x) - x - lgamma(a));
    if (x < a + 1.) {
        double ap = a, del = 1. / a, sum = del;
        for (int n = 0; n < 500; n++) {
            ap += 1.;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return sum * lead;
    }
    const double tiny = 1e-300;
    double b = x + 1. - a, c = 1. / tiny, d = 1. / b, h = d;
    for (int i = 1; i < 500; i++) {
        const double an = -(double)i * ((double)i - a);
        b += 2.;
        d = an * d + b; if (fabs(d) < tiny) d = tiny;
        c = b + an / c; if (fabs(c) < tiny) c = tiny;
        d = 1. / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.) < 1e-16) break;
    }
    return 1. - lead * h;
}
// lower incomplete gamma for any real a (safe_gamma_inc, :467-489): recurrence down from a+N > 0
GB_DEV double gb_safe_gamma_inc(double a, double x) {
    if (a > 0) return gb_gamma_inc_P(a, x) * tgamma(a);
    const int N = (int)ceil(-a);
    double A = 1., B = 0.;
    for (int n = 0; n < N; n++) {
        A = A * (a + n);
        double tmp = 1.;
        for (int m = N - 1; m > n; m--) tmp = tmp * (a + m);
        B = B + pow(x, a + n) * exp(-x) * tmp;
    }
    return (B + gb_gamma_inc_P(a + N, x) * tgamma(a + N)) / A;
}
#if !GB_STRICT
// fast build: lgamma(a) comes from the host (a is a parameter), reciprocals from the seed+refine primitive,
// and P = 1 to within 3e-17 once x > 40 (Q(a,x) < x^(a-1) e^-x / Gamma(a), a <= 1.5)
GB_DEV double gb_gamma_inc_P_fast(double a, double x, double lgam_a) {
    if (!(x > 0.)) return 0.;
    if (x > 40.) return 1.;
    const double lead = exp(a * log(x) - x - lgam_a);
    if (x < a + 1.) {
        double ap = a, del = gb_rcp(a), sum = del;
        for (int n = 0; n < 200; n++) {
            ap += 1.;
            del *= x * gb_rcp(ap);
            sum += del;
            if (del < sum * 1e-17) break;
        }
        return sum * lead;
    }
    const double tiny = 1e-300;
    double b = x + 1. - a, c = 1. / tiny, d = gb_rcp(b), h = d;
    for (int i = 1; i < 200; i++) {
        const double an = -(double)i * ((double)i - a);
        b += 2.;
        d = fma(an, d, b); if (fabs(d) < tiny) d = tiny;
        c = fma(an, gb_rcp(c), b); if (fabs(c) < tiny) c = tiny;
        d = gb_rcp(d);
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.) < 2e-15) break;     // the reciprocal seeds carry 2 ulp: del never settles on exactly 1
    }
    return fma(-lead, h, 1.);
}
#endif
struct PotPowerLawCutoff {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        const double r = gb_norm3(x, y, z);
        const double dPhi_dr = p[0] * p[1] / (r * r * r) * gb_gamma_inc_P(0.5 * (3 - p[2]), r * r / (p[3] * p[3]));
        gx = gx + dPhi_dr * x; gy = gy + dPhi_dr * y; gz = gz + dPhi_dr * z;
    }
    GB_DEV static double value(const double* p, double x, double y, double z) {
        const double r = gb_norm3(x, y, z);
        if (r == 0.) return -CUDART_INF;
        const double t0 = p[2] / 2.0, t1 = -t0, t2 = t1 + 1.5;
        const double t3 = r * r;
        const double t4 = t3 / (p[3] * p[3]);
        const double t5 = p[0] * p[1];
        const double t6 = t5 * gb_safe_gamma_inc(t2, t4) / (sqrt(t3) * tgamma(t1 + 2.5));
        const double phi_r = t0 * t6 - 3.0 / 2.0 * t6 + t5 * gb_safe_gamma_inc(t1 + 1, t4) / (p[3] * tgamma(t2));
        double phi_inf = 0.0;
        if (t2 > 0) phi_inf = t5 * tgamma(t1 + 1) / (p[3] * tgamma(t2));
        return phi_r - phi_inf;
    }
    GB_DEV static double density(const double* p, double x, double y, double z) {
        const double r = gb_norm3(x, y, z);
        const double A = p[1] / (2 * GB_PI) * pow(p[3], p[2] - 3) / tgamma(0.5 * (3 - p[2]));
        return A * pow(r, -p[2]) * exp(-r * r / (p[3] * p[3]));
    }
    #if !GB_STRICT
    static constexpr int USE = GB_USE_SIR;
    // d = [G m, lgamma(a), 1/r_c^2, 1/r_c, 2a, offset of the fit in DevPot::ext], a = (3-alpha)/2.
    // P(a, x) = x^a gamma*(a, x) with Tricomi's entire function gamma*; the host fits F(s) = gamma*(a, s^2),
    // s = r/r_c, by Chebyshev polynomials of degree GB_PLC_DEG on GB_PLC_NINT intervals of [0, GB_PLC_SMAX]
    // (capi.cu:plc_pack; 2e-16 relative), so one evaluation is a 10-term Clenshaw sum + exp(2a ln s)
    // instead of a 30-term series or continued fraction.  Beyond s = 6.4 (x = 41) P = 1 to 3e-17.
    template <class Ctx> GB_DEV static void accum(const double* p, const double* d, Ctx& c) {
        double P;
        const double sr = c.r * d[3];
        if (c.ext == nullptr || d[5] < 0.) {
            P = gb_gamma_inc_P_fast(0.5 * (3 - p[2]), c.r2 * d[2], d[1]);
        } else if (sr >= GB_PLC_SMAX) {
            P = 1.;
        } else {
            const double* tab = c.ext + (int)d[5];
            const double u = sr * (GB_PLC_NINT / GB_PLC_SMAX);
            int k = (int)u;
            k = k < GB_PLC_NINT ? k : GB_PLC_NINT - 1;
            const double t = 2. * (u - (double)k) - 1.;
            const double* co = tab + k * (GB_PLC_DEG + 1);
            double b1 = 0., b2 = 0.;
#pragma unroll
            for (int i = GB_PLC_DEG; i >= 1; i--) { const double b0 = fma(2. * t, b1, co[i] - b2); b2 = b1; b1 = b0; }
            const double F = fma(t, b1, 0.5 * co[0] - b2);
            P = (sr > 0.) ? exp(d[4] * log(sr)) * F : 0.;
        }
        c.Sir = fma(d[0] * P, c.ir * c.ir, c.Sir);              // G m P(a, r^2/r_c^2) / r^3
    }
#endif
};
