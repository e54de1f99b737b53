// multipole.cuh -- inner / outer multipole expansion on the device.
//
// Replaces mp_gradient / mp_potential / mp_density (reference potential/potential/builtin/
// multipole.cpp:30-139 helpers, :143-301 high-level, parameter layout :247-262:
// [G, lmax, num_coeff, inner, m, r_s, (S_lm, T_lm) interleaved in (l, m<=l) order]).
//   Phi = G M / r_s * sum_lm R_l(s) Y_lm(theta) (S_lm cos m phi + T_lm sin m phi),  s = r / r_s
//   R_l = s^l (inner) or s^-(l+1) (outer)
// Same angular machinery as scf.cuh (one Legendre recurrence per point instead of the reference's
// per-term gsl_sf_legendre_* calls; cos/sin(m phi) from powers of (x + i y)/R); the spherical-
// harmonic normalisation is folded into the coefficients by capi.cu:mp_pack, which stores (S,T)
// pairs in the reference's own (l, m) order.  Reference quirks kept: on the z axis (sin theta = 0)
// the theta and phi components are zero (:117-121, :281-283); the inner potential is 0 at r = 0
// (:222); the density is identically 0 (:404-420, "BUG HERE" note in the reference).
#pragma once

// WHAT: 0 = gradient (out[0..2] accumulate), 1 = value (out[0])
template <int LM, int WHAT>
__device__ __noinline__ void mp_eval(const double* __restrict__ p, const double* __restrict__ e,
                                     double x, double y, double z, double* __restrict__ out) {
    const double G = p[0];
    const int lmax = (int)p[1];
    const int inner = (int)p[3];
    const double M = p[4], rs = p[5];
    const double r = sqrt(x * x + y * y + z * z);
    const double s = r / rs;
    const double X = z / r;
    const double sintheta = sqrt(1. - X * X);
    const double Rc = sqrt(x * x + y * y);
    const double cphi = (Rc > 0.) ? x / Rc : 1., sphi = (Rc > 0.) ? y / Rc : 0.;

    double cm[LM + 1], sm[LM + 1];
    cm[0] = 1.; sm[0] = 0.;
    for (int m = 1; m <= lmax; m++) {
        cm[m] = cm[m - 1] * cphi - sm[m - 1] * sphi;
        sm[m] = sm[m - 1] * cphi + cm[m - 1] * sphi;
    }
    double Pl[LM + 1], Pm1[LM + 1], Pm2[LM + 1];
    for (int m = 0; m <= lmax; m++) { Pl[m] = 0.; Pm1[m] = 0.; Pm2[m] = 0.; }

    // radial: R_l and dR_l/ds as running products
    const double is = 1. / s;
    double Rl = inner ? 1. : is;                  // s^0 or s^-1
    double gr = 0., gt = 0., gp = 0., val = 0.;
    double pmm = 1.;
    const double2* __restrict__ co = reinterpret_cast<const double2*>(e);
    int i = 0;
    for (int l = 0; l <= lmax; l++) {
        for (int m = 0; m <= lmax; m++) { Pm2[m] = Pm1[m]; Pm1[m] = Pl[m]; }
        if (l > 0) pmm *= -(2. * l - 1.) * sintheta;
        for (int m = 0; m <= l; m++) {
            if (m < l - 1)       Pl[m] = (X * (2. * l - 1.) * Pm1[m] - (l + m - 1.) * Pm2[m]) / (double)(l - m);
            else if (m == l - 1) Pl[m] = X * (2. * m + 1.) * Pm1[m];
            else                 Pl[m] = pmm;
        }
        // dR_l/ds: l s^(l-1) (inner; 0 for l = 0) or -(l+1) s^(-l-2) (outer)
        const double dRl = inner ? ((l == 0) ? 0. : l * Rl * is) : -(l + 1.) * Rl * is;
        for (int m = 0; m <= l; m++, i++) {
            const double2 st = __ldg(co + i);
            const double CS = st.x * cm[m] + st.y * sm[m];
            if (WHAT == 1) { val += Rl * Pl[m] * CS; continue; }
            gr += dRl * Pl[m] * CS;
            if (l > 0) gt += (l * X * Pl[m] - (l + m) * Pm1[m]) * Rl * CS;
            if (m > 0) gp += m * Pl[m] * Rl * (st.y * cm[m] - st.x * sm[m]);
        }
        Rl = inner ? Rl * s : Rl * is;
    }
    if (WHAT == 1) {
        if (r == 0. && inner) val = 0.;
        out[0] = val * G * M / rs;
        return;
    }
    if (sintheta != 0.) { gt = gt / (sintheta * s); gp = gp / (s * sintheta); }
    else { gt = 0.; gp = 0.; }
    if (!(s > 0.)) { gr = 0.; gt = 0.; gp = 0.; }
    const double gx = sintheta * cphi * gr + X * cphi * gt - sphi * gp;
    const double gy = sintheta * sphi * gr + X * sphi * gt + cphi * gp;
    const double gz = X * gr - sintheta * gt;
    const double sc = G * M / (rs * rs);
    out[0] += gx * sc; out[1] += gy * sc; out[2] += gz * sc;
}

struct PotMultipole {
    GB_DEV static void gradient(const double* p, const double* e, double x, double y, double z,
                                double& gx, double& gy, double& gz) {
        double o[3] = {0., 0., 0.};
        mp_eval<GB_MP_LMAX, 0>(p, e, x, y, z, o);
        gx = gx + o[0]; gy = gy + o[1]; gz = gz + o[2];
    }
    GB_DEV static double value(const double* p, const double* e, double x, double y, double z) {
        double o[1];
        mp_eval<GB_MP_LMAX, 1>(p, e, x, y, z, o);
        return o[0];
    }
    GB_DEV static double density(const double*, const double*, double, double, double) { return 0.; }
};
