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

// ---- Null (builtin_potentials.cpp:18-21) -----------------------------------------------------
struct PotNull {
    GB_DEV static void gradient(const double*, double, double, double, double&, double&, double&) {}
    GB_DEV static double value(const double*, double, double, double) { return 0.; }
    GB_DEV static double density(const double*, double, double, double) { return 0.; }
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
};
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
struct PotMiyamotoNagai {
    GB_DEV static void gradient(const double* p, double x, double y, double z, double& gx, double& gy, double& gz) {
        gb_mn_gradient(p[0], p[1], p[2], p[3], x, y, z, gx, gy, gz);
    }
    GB_DEV static double value(const double* p, double x, double y, double z) { return gb_mn_value(p[0], p[1], p[2], p[3], x, y, z); }
    GB_DEV static double density(const double* p, double x, double y, double z) { return gb_mn_density(p[1], p[2], p[3], x, y, z); }
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
