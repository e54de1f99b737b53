// scf.cuh -- Hernquist-Ostriker SCF basis-function expansion on the device.
//
// Replaces scf_gradient / scf_value / scf_density (reference potential/scf/src/bfe.cpp:15-259) and
// their per-term helpers (bfe_helper.cpp:14-90).  The reference re-evaluates, for EVERY (n,l,m) term,
// two or three GSL Gegenbauer polynomials, a spherical-harmonic Legendre function, two plain Legendre
// functions, two gamma functions and two pow() calls (it carries a TODO about recurrences at
// bfe.cpp:87-88).  Here every special function is produced once per point by its three-term
// recurrence and shared across terms:
//   * cos(m phi), sin(m phi): powers of (x + i y)/R -- no atan2, no sincos;
//   * P_l^m(cos theta) with the Condon-Shortley phase (GSL's and the reference's convention,
//     SURVEY.md appendix A): standard upward recurrence in l;
//   * C_n^{(2l+3/2)}(xi) and C_{n-1}^{(2l+5/2)}(xi), xi = (s-1)/(s+1): upward recurrence in n;
//   * s^l (1+s)^{-2l-1}: a running product in l.
// The spherical-harmonic normalisation sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) is folded into the
// coefficients on the host (capi.cu: scf_pack), which also stores them as (S,T) pairs in [l][m][n]
// order so the inner loop over n reads one 16-byte pair per (n,l,m) with a warp-uniform address.
//
//   sums per (l,m):  A = sum_n Phi_nl S,  B = sum_n Phi_nl T,  A' = sum_n dPhi_nl/ds S,  B' likewise
//   grad_r     = sum_lm P_lm (cos A' + sin B')
//   grad_theta = sum_lm (l X P_lm - (l+m) P_{l-1,m}) / sin(theta) / s * (cos A + sin B)
//   grad_phi   = sum_lm m P_lm / (s sin(theta)) * (cos B - sin A)
// then the spherical -> Cartesian transform of bfe.cpp:181-187 and the G M / r_s^2 scale.
//
// Two instantiations: NM=10, LM=6 (everything in registers: BASELINE config C5) and NM=63, LM=15
// (runtime bounds, local-memory arrays).  Both are __noinline__: one SCF evaluation is > 2000 FP64
// operations, a call costs nothing next to it, and the generic composite would otherwise inline
// this body into all 15 RHS evaluations of the DOP853 kernel.
#pragma once

#define GB_SQRT_FOURPI 3.544907701811031

struct ScfHeader {   // p = [G, nmax, lmax, m, r_s]; e = packed (S,T) pairs, [l][m][n], m <= l
    double G, M, rs;
    int nmax, lmax;
};
GB_DEV ScfHeader scf_header(const double* p) {
    ScfHeader h;
    h.G = p[0]; h.nmax = (int)p[1]; h.lmax = (int)p[2]; h.M = p[3]; h.rs = p[4];
    return h;
}

// WHAT: 0 = gradient, 1 = value, 2 = density.  out[0..2] accumulate (gradient) or out[0] (scalar).
template <int NM, int LM, int WHAT>
__device__ __noinline__ void scf_eval(const double* __restrict__ p, const double* __restrict__ e,
                                      double x, double y, double z, double* __restrict__ out) {
    constexpr int UN = (NM <= 10) ? NM + 1 : 1;   // full unroll only for the register-resident variant
    constexpr int UM = (LM <= 6) ? LM + 1 : 1;
    const ScfHeader h = scf_header(p);
    const int nmax = h.nmax, lmax = h.lmax;
    const double r2 = x * x + y * y + z * z;
    const double r = sqrt(r2);
    const double s = r / h.rs;
    const double X = z / r;                      // cos(theta)
    const double R2 = x * x + y * y;
    const double sintheta = sqrt(1. - X * X);
    // cos(phi), sin(phi); on the z-axis phi = atan2(0,0) = 0 in the reference
    const double Rc = sqrt(R2);
    const double cphi = (Rc > 0.) ? x / Rc : 1., sphi = (Rc > 0.) ? y / Rc : 0.;

    // cos(m phi), sin(m phi)
    double cm[LM + 1], sm[LM + 1];
    cm[0] = 1.; sm[0] = 0.;
#pragma unroll UM
    for (int m = 1; m <= LM; m++) {
        if (m <= lmax) {
            cm[m] = cm[m - 1] * cphi - sm[m - 1] * sphi;
            sm[m] = sm[m - 1] * cphi + cm[m - 1] * sphi;
        }
    }

    const double xi = (s - 1.) / (s + 1.);
    const double ops = 1. + s;
    const double rfac = s / (ops * ops);         // ratio of s^l (1+s)^(-2l-1) between consecutive l
    double radial = 1. / ops;                    // s^0 (1+s)^(-1)

    double gr = 0., gt = 0., gp = 0., val = 0.;

    // Legendre P_l^m for the current l (Pl[m]) and l-1 (Pm1[m]) and l-2 (Pm2[m]), all m at once
    double Pl[LM + 1], Pm1[LM + 1], Pm2[LM + 1];
#pragma unroll UM
    for (int m = 0; m <= LM; m++) { Pl[m] = 0.; Pm1[m] = 0.; Pm2[m] = 0.; }

    size_t eoff = 0;                             // running offset into the packed coefficients
    double pmm = 1.;                             // P_l^l by the diagonal recurrence
#pragma unroll 1
    for (int l = 0; l <= lmax; l++) {
        // ---- Legendre row l -----------------------------------------------------------------------
#pragma unroll UM
        for (int m = 0; m <= LM; m++) { Pm2[m] = Pm1[m]; Pm1[m] = Pl[m]; }
        if (l > 0) pmm *= -(2. * l - 1.) * sintheta;         // P_l^l = (-1)^l (2l-1)!! sin^l
#pragma unroll UM
        for (int m = 0; m <= LM; m++) {
            if (m < l - 1)       Pl[m] = (X * (2. * l - 1.) * Pm1[m] - (l + m - 1.) * Pm2[m]) / (double)(l - m);
            else if (m == l - 1) Pl[m] = X * (2. * m + 1.) * Pm1[m];  // P_{m+1}^m = x (2m+1) P_m^m
            else if (m == l)     Pl[m] = pmm;
        }

        // ---- radial functions for this l: Phi_nl, dPhi_nl/ds (or rho_nl), n = 0..nmax ----------------
        const double lam = 2. * l + 1.5;
        double Phi[NM + 1], dPhi[NM + 1];
        {
            // C_n^{(lam)}(xi) and C_{n-1}^{(lam+1)}(xi)
            double ca = 1., cb = 2. * lam * xi;              // C_0, C_1 of family lam
            double da = 0., db = 1.;                          // C_{-1}, C_0 of family lam+1
            const double pre = -GB_SQRT_FOURPI * radial;                       // -sqrt(4pi) s^l (1+s)^(-2l-1)
            // sqrt(4pi) s^(l-1) (1+s)^(-3-2l) = sqrt(4pi) * radial / (s (1+s)^2)
            const double dpre = GB_SQRT_FOURPI * radial / (s * ops * ops);
            const double poly = ops * (l * (s - 1.) + s);
            const double dcoef = -2. * (3. + 4. * l) * s;
#pragma unroll UN
            for (int n = 0; n <= NM; n++) {
                if (n <= nmax) {
                    double Cn, Dn;                            // C_n^{(lam)}, C_{n-1}^{(lam+1)}
                    if (n == 0) { Cn = ca; Dn = da; }
                    else if (n == 1) { Cn = cb; Dn = db; }
                    else {
                        Cn = (2. * (n + lam - 1.) * xi * cb - (n + 2. * lam - 2.) * ca) / n;
                        ca = cb; cb = Cn;
                        const int k = n - 1; const double lam1 = lam + 1.;
                        Dn = (k == 1) ? 2. * lam1 * xi
                                      : (2. * (k + lam1 - 1.) * xi * db - (k + 2. * lam1 - 2.) * da) / k;
                        da = db; db = Dn;
                    }
                    if (WHAT == 2) {
                        const double Knl = 0.5 * n * (n + 4. * l + 3.) + (l + 1.) * (2. * l + 1.);
                        // rho_nl = sqrt(4pi) Knl/(2pi) s^l / (s (1+s)^(2l+3)) C_n  = Knl/(2pi) * dpre-like
                        Phi[n] = GB_SQRT_FOURPI * Knl / (2. * GB_PI) * radial / (s * ops * ops) * Cn;
                    } else {
                        Phi[n] = pre * Cn;
                        if (WHAT == 0) dPhi[n] = dpre * (dcoef * Dn + poly * Cn);
                    }
                }
            }
        }

        // ---- angular sums for this l -------------------------------------------------------------------
#pragma unroll UM
        for (int m = 0; m <= LM; m++) {
            if (m <= l) {
                double A = 0., B = 0., Ap = 0., Bp = 0.;
                const double2* __restrict__ co = reinterpret_cast<const double2*>(e) + eoff;
#pragma unroll UN
                for (int n = 0; n <= NM; n++) {
                    if (n <= nmax) {
                        const double2 st = __ldg(co + n);
                        A = fma(Phi[n], st.x, A); B = fma(Phi[n], st.y, B);
                        if (WHAT == 0) { Ap = fma(dPhi[n], st.x, Ap); Bp = fma(dPhi[n], st.y, Bp); }
                    }
                }
                eoff += nmax + 1;
                const double CS = cm[m] * A + sm[m] * B;
                if (WHAT != 0) {
                    val += Pl[m] * CS;
                } else {
                    gr += Pl[m] * (cm[m] * Ap + sm[m] * Bp);
                    if (l > 0) gt += (l * X * Pl[m] - (l + m) * Pm1[m]) * CS;
                    if (m > 0) gp += m * Pl[m] * (cm[m] * B - sm[m] * A);
                }
            }
        }
        radial *= rfac;
    }

    if (WHAT == 1) { out[0] = val * h.G * h.M / h.rs; return; }
    if (WHAT == 2) { out[0] = val * h.M / (h.rs * h.rs * h.rs); return; }
    // common factors of the theta and phi components (bfe_helper.cpp:76-87, bfe.cpp:168-170)
    gt = gt / (sintheta * s);
    gp = gp / (s * sintheta);
    const double gx = sintheta * cphi * gr + X * cphi * gt - sphi * gp;
    const double gy = sintheta * sphi * gr + X * sphi * gt + cphi * gp;
    const double gz = X * gr - sintheta * gt;
    const double sc = h.G * h.M / (h.rs * h.rs);
    out[0] += gx * sc; out[1] += gy * sc; out[2] += gz * sc;
}

#if !GB_STRICT
// ------------------------------------------------------------------------------------------------
// Fast-build gradient for nmax <= 10, lmax <= 6 (BASELINE config C5) with everything resolved at
// compile time.  Differences from scf_eval<10,6,0> above (same series, same recurrences):
//   * the (S,T) pairs come from DevPot::cext -- the kernel-parameter constant bank -- in a fixed
//     [l][m][n] layout padded to (10,6), so each of the 4 x 308 coefficient FMAs takes its coefficient
//     as a constant-bank operand: no load instruction, no address arithmetic;
//   * l, m, n loops are fully unrolled, so every recurrence coefficient ((2l-1)/(l-m), 2(n+lam-1)/n ...)
//     is a literal: the first version spent ~1700 of its 4600 FP64 instructions per evaluation on ~150
//     divisions by small integers (profiles/ncu_r1_leapfrog_scf_v0.txt);
//   * the radial prefactors are applied to the (l) partial sums, not to each Phi_nl;
//   * n is the outer loop of each l: one recurrence step feeds 4 (l+1) independent FMAs.
// ~2300 FP64 instructions per evaluation.
// ------------------------------------------------------------------------------------------------
#define GB_SCF_NM 10
#define GB_SCF_LM 6
GB_DEV void scf_fast_gradient(const DevPot& P, const double* __restrict__ d, double x, double y, double z,
                              double& ogx, double& ogy, double& ogz) {
    constexpr int NM = GB_SCF_NM, LM = GB_SCF_LM;
    const double* __restrict__ co = P.cext;
    const double R2 = fma(y, y, x * x);
    const double r2 = fma(z, z, R2);
    const double ir = gb_rsqrt(r2);
    const double r = r2 * ir;
    const double s = r * d[1];                     // r / r_s
    const double X = z * ir;                       // cos(theta)
    const bool offaxis = R2 > 0.;
    const double iR = gb_rsqrt(offaxis ? R2 : 1.);
    const double cphi = offaxis ? x * iR : 1., sphi = offaxis ? y * iR : 0.;
    const double sintheta = offaxis ? (R2 * iR) * ir : 0.;

    double cm[LM + 1], sm[LM + 1];
    cm[0] = 1.; sm[0] = 0.;
#pragma unroll
    for (int m = 1; m <= LM; m++) {
        cm[m] = fma(cm[m - 1], cphi, -(sm[m - 1] * sphi));
        sm[m] = fma(sm[m - 1], cphi, cm[m - 1] * sphi);
    }
    const double ops = 1. + s;
    const double iops = gb_rcp(ops);
    const double is = gb_rcp(s);
    const double xi = (s - 1.) * iops;
    const double rfac = s * (iops * iops);         // ratio of s^l (1+s)^(-2l-1) between consecutive l
    double radial = iops;                          // s^0 (1+s)^(-1)
    const double dscale = is * (iops * iops);      // dPhi prefactor relative to radial: 1/(s (1+s)^2)

    double gr = 0., gt = 0., gp = 0.;
    double Pl[LM + 1], Pm1[LM + 1], Pm2[LM + 1];
#pragma unroll
    for (int m = 0; m <= LM; m++) { Pl[m] = 0.; Pm1[m] = 0.; Pm2[m] = 0.; }
    double pmm = 1.;
#pragma unroll
    for (int l = 0; l <= LM; l++) {
        // ---- Legendre row l (Condon-Shortley phase) ----------------------------------------------
#pragma unroll
        for (int m = 0; m <= LM; m++) { Pm2[m] = Pm1[m]; Pm1[m] = Pl[m]; }
        if (l > 0) pmm *= -(2. * l - 1.) * sintheta;
#pragma unroll
        for (int m = 0; m <= l; m++) {
            if (m < l - 1) {
                const double a = (2. * l - 1.) / (double)(l - m), b = (l + m - 1.) / (double)(l - m);   // literals
                Pl[m] = fma(a, X * Pm1[m], -(b * Pm2[m]));
            } else if (m == l - 1) {
                Pl[m] = (X * (2. * m + 1.)) * Pm1[m];
            } else {
                Pl[m] = pmm;
            }
        }
        // ---- Gegenbauer recurrences in n, accumulating the four sums of every m <= l -----------------
        const double lam = 2. * l + 1.5, lam1 = lam + 1.;
        double A0[LM + 1], B0[LM + 1], AD[LM + 1], BD[LM + 1];
#pragma unroll
        for (int m = 0; m <= LM; m++) { A0[m] = 0.; B0[m] = 0.; AD[m] = 0.; BD[m] = 0.; }
        double ca = 0., cb = 0., da = 0., db = 0.;
#pragma unroll
        for (int n = 0; n <= NM; n++) {
            double Cn, Dn;
            if (n == 0) { Cn = 1.; Dn = 0.; }
            else if (n == 1) { Cn = (2. * lam) * xi; Dn = 1.; }
            else {
                const double an = 2. * (n + lam - 1.) / n, bn = (n + 2. * lam - 2.) / n;               // literals
                Cn = fma(an * xi, cb, -(bn * ca));
                const int k = n - 1;
                if (k == 1) Dn = (2. * lam1) * xi;
                else {
                    const double ak = 2. * (k + lam1 - 1.) / k, bk = (k + 2. * lam1 - 2.) / k;
                    Dn = fma(ak * xi, db, -(bk * da));
                }
            }
            ca = cb; cb = Cn; da = db; db = Dn;
#pragma unroll
            for (int m = 0; m <= l; m++) {
                const int i = 2 * (((l * (l + 1)) / 2 + m) * (NM + 1) + n);
                const double S = co[i], T = co[i + 1];
                A0[m] = fma(Cn, S, A0[m]); B0[m] = fma(Cn, T, B0[m]);
                if (n > 0) { AD[m] = fma(Dn, S, AD[m]); BD[m] = fma(Dn, T, BD[m]); }
            }
        }
        // ---- angular sums of this l; radial prefactors applied once per l -------------------------------
        const double poly = ops * fma((double)l, s - 1., s);
        const double dcoef = (-2. * (3. + 4. * l)) * s;
        double gr_l = 0., gt_l = 0., gp_l = 0.;
#pragma unroll
        for (int m = 0; m <= l; m++) {
            const double CS0 = fma(cm[m], A0[m], sm[m] * B0[m]);
            const double CSD = fma(cm[m], AD[m], sm[m] * BD[m]);
            gr_l = fma(Pl[m], fma(dcoef, CSD, poly * CS0), gr_l);
            if (l > 0) gt_l = fma(fma((double)l * X, Pl[m], -((double)(l + m) * Pm1[m])), CS0, gt_l);
            if (m > 0) gp_l = fma((double)m * Pl[m], fma(cm[m], B0[m], -(sm[m] * A0[m])), gp_l);
        }
        const double pre = -GB_SQRT_FOURPI * radial;          // Phi_nl  = pre * C_n
        const double dpre = (GB_SQRT_FOURPI * radial) * dscale;   // dPhi_nl = dpre * (dcoef D_n + poly C_n)
        gr = fma(dpre, gr_l, gr);
        gt = fma(pre, gt_l, gt);
        gp = fma(pre, gp_l, gp);
        radial *= rfac;
    }
    // common factors of the theta and phi components (bfe_helper.cpp:76-87, bfe.cpp:168-170)
    const double ist = gb_rcp(sintheta * s);
    gt *= ist; gp *= ist;
    const double gx = fma(sintheta * cphi, gr, fma(X * cphi, gt, -(sphi * gp)));
    const double gy = fma(sintheta * sphi, gr, fma(X * sphi, gt, cphi * gp));
    const double gz = fma(X, gr, -(sintheta * gt));
    ogx = gx * d[0]; ogy = gy * d[0]; ogz = gz * d[0];
}
#endif

struct PotSCF {
    GB_DEV static void gradient(const double* p, const double* e, double x, double y, double z,
                                double& gx, double& gy, double& gz) {
        double o[3] = {0., 0., 0.};
        if ((int)p[1] <= 10 && (int)p[2] <= 6) scf_eval<10, 6, 0>(p, e, x, y, z, o);
        else scf_eval<63, 15, 0>(p, e, x, y, z, o);
        gx = gx + o[0]; gy = gy + o[1]; gz = gz + o[2];
    }
    GB_DEV static double value(const double* p, const double* e, double x, double y, double z) {
        double o[1];
        if ((int)p[1] <= 10 && (int)p[2] <= 6) scf_eval<10, 6, 1>(p, e, x, y, z, o);
        else scf_eval<63, 15, 1>(p, e, x, y, z, o);
        return o[0];
    }
    GB_DEV static double density(const double* p, const double* e, double x, double y, double z) {
        double o[1];
        if ((int)p[1] <= 10 && (int)p[2] <= 6) scf_eval<10, 6, 2>(p, e, x, y, z, o);
        else scf_eval<63, 15, 2>(p, e, x, y, z, o);
        return o[0];
    }
};
