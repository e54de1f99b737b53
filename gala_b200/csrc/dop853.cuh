// dop853.cuh -- adaptive Dormand-Prince 8(5,3), one orbit (n = 6) per thread, per-lane step control.
//
// Restates dopcor/hinit/contd8 of the reference (integrate/cyintegrators/dopri/dop853.cpp:18-650,
// 869-904) and the driver settings of dop853_helper / dop853_step
// (integrate/cyintegrators/dop853.pyx:27-75,90-193).  Differences that are deliberate and
// documented in DESIGN.md:
//   * the error norm and the step size h are PER ORBIT (the reference shares them across a batch of
//     nbatch orbits, dop853.pyx:228-245; this kernel is the reference's nbatch=1);
//   * every lane keeps its own output cursor into the caller's time grid (dense output,
//     dop853.cpp:584-612).
// The stiffness test reproduces the reference as coded (dop853.cpp:460-485): a hit (hlamb > 6.1)
// ends the orbit with -4 immediately because of the missing braces at :474-480.
#pragma once
#include "dop853_coeffs.cuh"

struct Dop853Stats {
    int32_t* status;
    int32_t* nstep;
    int32_t* naccpt;
    int32_t* nrejct;
    int32_t* nfcn;
};

GB_DEV double gb_sign(double a, double b) { return (b < 0.0) ? -fabs(a) : fabs(a); }
GB_DEV double gb_min(double a, double b) { return (a < b) ? a : b; }
GB_DEV double gb_max(double a, double b) { return (a > b) ? a : b; }

// err^(1/8): libm pow in strict mode (dop853.cpp:447); three correctly-rounded square roots otherwise.
GB_DEV double gb_pow_eighth(double err) {
#if GB_STRICT
    return pow(err, 1.0 / 8.0);
#else
    return sqrt(sqrt(sqrt(err)));
#endif
}

// RHS functor interface: void operator()(double t, const double (&w)[6], double (&f)[6]) const
//
// OUT interface (DENSE only): void operator()(int idx, const double (&v)[6]) const
//
// Returns the dop853 code: 1 ok, -2 nmax exceeded, -3 step too small, -4 stiff.
template <bool DENSE, class RHS, class OUT>
GB_DEV int dop853_integrate(const RHS& rhs, const OUT& emit, const Dop853Args& a, double x, double xend,
                            double (&y)[6], double h, const double* __restrict__ tout, int ntout,
                            int& out_idx, int& nstep_, int& naccpt_, int& nrejct_, int& nfcn_) {
    using namespace dp8;
    constexpr int n = 6;
    double k1[n], k2[n], k3[n], k4[n], k5[n], k6[n], k7[n], k8[n], k9[n], k10[n], yy1[n];
    double rc1[n], rc2[n], rc3[n], rc4[n], rc5[n], rc6[n], rc7[n], rc8[n];

    const double safe = 0.9, fac1 = 0.333, fac2 = 6.0;     // dop853.cpp:755-768 defaults
    const double facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    const double posneg = gb_sign(1.0, xend - x);
    const double atoli = a.atol, rtoli = a.rtol;
    double hmax = (a.hmax == 0.0) ? (xend - x) : a.hmax;   // dop853.cpp:787-788
    hmax = fabs(hmax);
    double facold = 1.0E-4;
    double hlamb = 0.0;
    int last = 0, reject = 0;
    int nstep = 0, naccpt = 0, nrejct = 0, nfcn = 0;
    double hnew;

    rhs(x, y, k1);
    if (h == 0.0) {
        // hinit (dop853.cpp:18-86), iord = 8
        double dnf = 0.0, dny = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            const double sk = atoli + rtoli * fabs(y[i]);
            double sqr = k1[i] / sk; dnf += sqr * sqr;
            sqr = y[i] / sk; dny += sqr * sqr;
        }
        double hh = ((dnf <= 1.0E-10) || (dny <= 1.0E-10)) ? 1.0E-6 : sqrt(dny / dnf) * 0.01;
        hh = gb_min(hh, hmax);
        hh = gb_sign(hh, posneg);
#pragma unroll
        for (int i = 0; i < n; i++) k3[i] = y[i] + hh * k1[i];
        rhs(x + hh, k3, k2);
        double der2 = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            const double sk = atoli + rtoli * fabs(y[i]);
            const double sqr = (k2[i] - k1[i]) / sk; der2 += sqr * sqr;
        }
        der2 = sqrt(der2) / hh;
        const double der12 = gb_max(fabs(der2), sqrt(dnf));
        const double h1 = (der12 <= 1.0E-15) ? gb_max(1.0E-6, fabs(hh) * 1.0E-3) : pow(0.01 / der12, 1.0 / 8.0);
        hh = gb_min(100.0 * fabs(hh), gb_min(h1, hmax));
        h = gb_sign(hh, posneg);
    }
    nfcn += 2;

    int code = 0;
    while (true) {
        if (nstep > a.nmax) { code = -2; break; }
        if (0.1 * fabs(h) <= fabs(x) * a.uround) { code = -3; break; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = 1; }
        nstep++;

        // the twelve stages (dop853.cpp:369-409)
#pragma unroll
        for (int i = 0; i < n; i++) yy1[i] = y[i] + h * a21 * k1[i];
        rhs(x + c2 * h, yy1, k2);
#pragma unroll
        for (int i = 0; i < n; i++) yy1[i] = y[i] + h * (a31 * k1[i] + a32 * k2[i]);
        rhs(x + c3 * h, yy1, k3);
#pragma unroll
        for (int i = 0; i < n; i++) yy1[i] = y[i] + h * (a41 * k1[i] + a43 * k3[i]);
        rhs(x + c4 * h, yy1, k4);
#pragma unroll
        for (int i = 0; i < n; i++) yy1[i] = y[i] + h * (a51 * k1[i] + a53 * k3[i] + a54 * k4[i]);
        rhs(x + c5 * h, yy1, k5);
#pragma unroll
        for (int i = 0; i < n; i++) yy1[i] = y[i] + h * (a61 * k1[i] + a64 * k4[i] + a65 * k5[i]);
        rhs(x + c6 * h, yy1, k6);
#pragma unroll
        for (int i = 0; i < n; i++) yy1[i] = y[i] + h * (a71 * k1[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
        rhs(x + c7 * h, yy1, k7);
#pragma unroll
        for (int i = 0; i < n; i++)
            yy1[i] = y[i] + h * (a81 * k1[i] + a84 * k4[i] + a85 * k5[i] + a86 * k6[i] + a87 * k7[i]);
        rhs(x + c8 * h, yy1, k8);
#pragma unroll
        for (int i = 0; i < n; i++)
            yy1[i] = y[i] + h * (a91 * k1[i] + a94 * k4[i] + a95 * k5[i] + a96 * k6[i] + a97 * k7[i] + a98 * k8[i]);
        rhs(x + c9 * h, yy1, k9);
#pragma unroll
        for (int i = 0; i < n; i++)
            yy1[i] = y[i] + h * (a101 * k1[i] + a104 * k4[i] + a105 * k5[i] + a106 * k6[i] + a107 * k7[i] +
                                 a108 * k8[i] + a109 * k9[i]);
        rhs(x + c10 * h, yy1, k10);
#pragma unroll
        for (int i = 0; i < n; i++)
            yy1[i] = y[i] + h * (a111 * k1[i] + a114 * k4[i] + a115 * k5[i] + a116 * k6[i] + a117 * k7[i] +
                                 a118 * k8[i] + a119 * k9[i] + a1110 * k10[i]);
        rhs(x + c11 * h, yy1, k2);
        const double xph = x + h;
#pragma unroll
        for (int i = 0; i < n; i++)
            yy1[i] = y[i] + h * (a121 * k1[i] + a124 * k4[i] + a125 * k5[i] + a126 * k6[i] + a127 * k7[i] +
                                 a128 * k8[i] + a129 * k9[i] + a1210 * k10[i] + a1211 * k2[i]);
        rhs(xph, yy1, k3);
        nfcn += 11;
#pragma unroll
        for (int i = 0; i < n; i++) {
            k4[i] = b1 * k1[i] + b6 * k6[i] + b7 * k7[i] + b8 * k8[i] + b9 * k9[i] + b10 * k10[i] + b11 * k2[i] +
                    b12 * k3[i];
            k5[i] = y[i] + h * k4[i];
        }

        // error estimation (dop853.cpp:416-444), scalar tolerances, norm over this orbit's 6 components
        double err = 0.0, err2 = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            const double sk = atoli + rtoli * gb_max(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - bhh1 * k1[i] - bhh2 * k9[i] - bhh3 * k3[i];
            double sqr = erri / sk;
            err2 += sqr * sqr;
            erri = er1 * k1[i] + er6 * k6[i] + er7 * k7[i] + er8 * k8[i] + er9 * k9[i] + er10 * k10[i] +
                   er11 * k2[i] + er12 * k3[i];
            sqr = erri / sk;
            err += sqr * sqr;
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / (deno * (double)n));

        // step-size controller (dop853.cpp:446-452), beta = 0 => pow(facold, beta) == 1
        const double fac11 = gb_pow_eighth(err);
        double fac = fac11;
        fac = gb_max(facc2, gb_min(facc1, fac / safe));
        hnew = h / fac;

        if (err <= 1.0) {
            // accepted
            facold = gb_max(err, 1.0E-4);
            naccpt++;
            rhs(xph, k5, k4);
            nfcn++;

            // stiffness detection as coded in the reference (dop853.cpp:460-485)
            if (!(naccpt % a.nstiff)) {
                double stnum = 0.0, stden = 0.0;
#pragma unroll
                for (int i = 0; i < n; i++) {
                    double sqr = k4[i] - k3[i]; stnum += sqr * sqr;
                    sqr = k5[i] - yy1[i]; stden += sqr * sqr;
                }
                if (stden > 0.0) hlamb = h * sqrt(stnum / stden);
                if (hlamb > 6.1) { code = -4; break; }
            }

            if (DENSE) {
                // dense-output preparation (dop853.cpp:492-582)
#pragma unroll
                for (int i = 0; i < n; i++) {
                    rc1[i] = y[i];
                    const double ydiff = k5[i] - y[i];
                    rc2[i] = ydiff;
                    const double bspl = h * k1[i] - ydiff;
                    rc3[i] = bspl;
                    rc4[i] = ydiff - h * k4[i] - bspl;
                    rc5[i] = d41 * k1[i] + d46 * k6[i] + d47 * k7[i] + d48 * k8[i] + d49 * k9[i] + d410 * k10[i] +
                             d411 * k2[i] + d412 * k3[i];
                    rc6[i] = d51 * k1[i] + d56 * k6[i] + d57 * k7[i] + d58 * k8[i] + d59 * k9[i] + d510 * k10[i] +
                             d511 * k2[i] + d512 * k3[i];
                    rc7[i] = d61 * k1[i] + d66 * k6[i] + d67 * k7[i] + d68 * k8[i] + d69 * k9[i] + d610 * k10[i] +
                             d611 * k2[i] + d612 * k3[i];
                    rc8[i] = d71 * k1[i] + d76 * k6[i] + d77 * k7[i] + d78 * k8[i] + d79 * k9[i] + d710 * k10[i] +
                             d711 * k2[i] + d712 * k3[i];
                }
#pragma unroll
                for (int i = 0; i < n; i++)
                    yy1[i] = y[i] + h * (a141 * k1[i] + a147 * k7[i] + a148 * k8[i] + a149 * k9[i] + a1410 * k10[i] +
                                         a1411 * k2[i] + a1412 * k3[i] + a1413 * k4[i]);
                rhs(x + c14 * h, yy1, k10);
#pragma unroll
                for (int i = 0; i < n; i++)
                    yy1[i] = y[i] + h * (a151 * k1[i] + a156 * k6[i] + a157 * k7[i] + a158 * k8[i] + a1511 * k2[i] +
                                         a1512 * k3[i] + a1513 * k4[i] + a1514 * k10[i]);
                rhs(x + c15 * h, yy1, k2);
#pragma unroll
                for (int i = 0; i < n; i++)
                    yy1[i] = y[i] + h * (a161 * k1[i] + a166 * k6[i] + a167 * k7[i] + a168 * k8[i] + a169 * k9[i] +
                                         a1613 * k4[i] + a1614 * k10[i] + a1615 * k2[i]);
                rhs(x + c16 * h, yy1, k3);
                nfcn += 3;
#pragma unroll
                for (int i = 0; i < n; i++) {
                    rc5[i] = h * (rc5[i] + d413 * k4[i] + d414 * k10[i] + d415 * k2[i] + d416 * k3[i]);
                    rc6[i] = h * (rc6[i] + d513 * k4[i] + d514 * k10[i] + d515 * k2[i] + d516 * k3[i]);
                    rc7[i] = h * (rc7[i] + d613 * k4[i] + d614 * k10[i] + d615 * k2[i] + d616 * k3[i]);
                    rc8[i] = h * (rc8[i] + d713 * k4[i] + d714 * k10[i] + d715 * k2[i] + d716 * k3[i]);
                }
                // fill every requested time inside [x, x+h] (dop853.cpp:584-612; contd8 :869-904)
                const double x0 = x, x1 = x0 + h;
                while (out_idx < ntout) {
                    const double t_out = tout[out_idx];
                    if ((x0 <= t_out && t_out <= x1) || (x1 <= t_out && t_out <= x0)) {
                        const double s = (t_out - x0) / h;
                        const double s1 = 1.0 - s;
                        double v[n];
#pragma unroll
                        for (int i = 0; i < n; i++)
                            v[i] = rc1[i] + s * (rc2[i] + s1 * (rc3[i] + s * (rc4[i] + s1 * (rc5[i] + s * (rc6[i] + s1 * (rc7[i] + s * rc8[i]))))));
                        emit(out_idx, v);
                        out_idx++;
                    } else {
                        break;
                    }
                }
            }

#pragma unroll
            for (int i = 0; i < n; i++) { k1[i] = k4[i]; y[i] = k5[i]; }
            x = xph;
            if (last) { code = 1; break; }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * gb_min(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            // rejected (dop853.cpp:638-645)
            hnew = h / gb_min(facc1, fac11 / safe);
            reject = 1;
            if (naccpt >= 1) nrejct = nrejct + 1;
            last = 0;
        }
        h = hnew;
    }
    nstep_ = nstep; naccpt_ = naccpt; nrejct_ = nrejct; nfcn_ = nfcn;
    return code;
}

// ------------------------------------------------------------------------------------------------
// dop853_integrate_hamiltonian (integrate/cyintegrators/dop853.pyx:196-250): (6,N) in,
// (6,ntimes,N) dense output or (6,N) final state.
// ------------------------------------------------------------------------------------------------
template <class C, bool ROT, bool DENSE>
__global__ void __launch_bounds__(128)
k_dop853(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F, const __grid_constant__ Dop853Args a,
         const double* __restrict__ w0, size_t N, const double* __restrict__ t, int ntimes,
         double* __restrict__ out, Dop853Stats st) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double y[6];
#pragma unroll
    for (int k = 0; k < 6; k++) y[k] = w0[k * N + i];
    const size_t TS = (size_t)ntimes * N;
    auto rhs = [&](double tt, const double (&w)[6], double (&f)[6]) { ham_rhs<C, ROT>(P, F, tt, w, f); };
    auto emit = [&](int idx, const double (&v)[6]) {
        double* o = out + (size_t)idx * N + i;
#pragma unroll
        for (int k = 0; k < 6; k++) __stcs(o + k * TS, v[k]);
    };
    int out_idx = 0, nstep, naccpt, nrejct, nfcn;
    const int code = dop853_integrate<DENSE>(rhs, emit, a, t[0], t[ntimes - 1], y, a.h0, t, ntimes, out_idx,
                                             nstep, naccpt, nrejct, nfcn);
    if (DENSE) {
        // a failed orbit leaves its remaining rows undefined in the reference (np.empty); use NaN
        const double nan = CUDART_NAN;
        for (int j = out_idx; j < ntimes; j++) {
            double* o = out + (size_t)j * N + i;
#pragma unroll
            for (int k = 0; k < 6; k++) o[k * TS] = nan;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) out[k * N + i] = y[k];
    }
    if (st.status) st.status[i] = code;
    if (st.nstep) st.nstep[i] = nstep;
    if (st.naccpt) st.naccpt[i] = naccpt;
    if (st.nrejct) st.nrejct[i] = nrejct;
    if (st.nfcn) st.nfcn[i] = nfcn;
}
