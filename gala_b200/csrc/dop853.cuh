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
    // three square roots as x * rsqrt(x) (2 ulp each, no slow path); err == 0 stays 0
    if (!(err > 0.0)) return 0.0;
    const double s1 = err * gb_rsqrt(err);
    const double s2 = s1 * gb_rsqrt(s1);
    return s2 * gb_rsqrt(s2);
#endif
}

// RHS functor interface: void operator()(double t, const double (&w)[6], double (&f)[6]) const
//
// OUT interface (DENSE only), one of
//   immediate: void operator()(int idx, const double (&v)[6]) const -- step() evaluates every requested time
//              inside the accepted step itself (a per-lane loop) and hands the sample over;
//   deferred:  static constexpr bool kDeferred = true; st(row, v) / ld(row) / stash(x0, h, ih) -- step() only
//              deposits the eight dense-output coefficient vectors (row = 6 * k + component, k = 0..7 for
//              rcont1..rcont8) and the step's (x, h); the caller evaluates the samples afterwards, outside the
//              divergent accept branch (k_dop853_dyn: the whole warp shares the pending samples of its lanes).
//
// Dop853Lane holds everything dopcor keeps between attempted steps (dop853.cpp:333-366), so one
// lane can be (re)initialised with a new orbit at any iteration of a persistent warp loop.
//   init(): dop853() front-end + the first RHS evaluation + hinit (dop853.cpp:18-86, 361-366)
//   step(): one attempted step of the dopcor loop (dop853.cpp:367-650); returns 0 to continue or
//           the dop853 code: 1 ok, -2 nmax exceeded, -3 step too small, -4 stiff.
#ifndef GB_D8_UNROLL_MAX
#define GB_D8_UNROLL_MAX 12
#endif
template <class T, class = void> struct gb_is_deferred { static constexpr bool value = false; };
template <class T> struct gb_is_deferred<T, decltype((void)T::kDeferred)> { static constexpr bool value = T::kDeferred; };
template <bool DENSE, int NDIM = 6>
struct Dop853Lane {
    static constexpr int n = NDIM;
    // loops over the n components: fully unrolled (state in registers, the compiler spills what does not
    // fit) up to n = 12 = one massive body + one particle; rolled (state in local memory, small code) beyond
    static constexpr int GB_NU = (NDIM <= GB_D8_UNROLL_MAX) ? NDIM : 1;   // 6 for one orbit; 6 (nb + 1) when a lane carries nb massive bodies (nbody.cuh)
    double y[n], k1[n];
    int nrun = NDIM;   // number of components in use (n > GB_D8_UNROLL_MAX: a system of nrun <= NDIM equations)
    double x, xend, h, posneg, hmax, facold, hlamb;
    int last, reject, nstep, naccpt, nrejct, nfcn, out_idx;

    template <class RHS>
    GB_DEV void init(const RHS& rhs, const Dop853Args& a, double x0, double xend_, double h0) {
        const int nn = (NDIM <= GB_D8_UNROLL_MAX) ? NDIM : nrun;
        x = x0; xend = xend_; h = h0;
        posneg = gb_sign(1.0, xend - x);
        hmax = (a.hmax == 0.0) ? (xend - x) : a.hmax;   // dop853.cpp:787-788
        hmax = fabs(hmax);
        facold = 1.0E-4;
        hlamb = 0.0;
        last = 0; reject = 0;
        nstep = 0; naccpt = 0; nrejct = 0; nfcn = 0; out_idx = 0;
        const double atoli = a.atol, rtoli = a.rtol;
        rhs(x, y, k1);
        if (h == 0.0) {
            // hinit (dop853.cpp:18-86), iord = 8
            double k2[n], k3[n];
            double dnf = 0.0, dny = 0.0;
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) {
                const double sk = atoli + rtoli * fabs(y[i]);
                double sqr = k1[i] / sk; dnf += sqr * sqr;
                sqr = y[i] / sk; dny += sqr * sqr;
            }
            double hh = ((dnf <= 1.0E-10) || (dny <= 1.0E-10)) ? 1.0E-6 : sqrt(dny / dnf) * 0.01;
            hh = gb_min(hh, hmax);
            hh = gb_sign(hh, posneg);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) k3[i] = y[i] + hh * k1[i];
            rhs(x + hh, k3, k2);
            double der2 = 0.0;
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) {
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
    }

    template <class RHS, class OUT>
    GB_DEV int step(const RHS& rhs, const OUT& emit, const Dop853Args& a, const double* __restrict__ tout, int ntout) {
        using namespace dp8;
        const int nn = (NDIM <= GB_D8_UNROLL_MAX) ? NDIM : nrun;
        double k2[n], k3[n], k4[n], k5[n], k6[n], k7[n], k8[n], k9[n], k10[n], yy1[n];
        double rc1[n], rc2[n], rc3[n], rc4[n], rc5[n], rc6[n], rc7[n], rc8[n];
        const double safe = 0.9, fac1 = 0.333, fac2 = 6.0;     // dop853.cpp:755-768 defaults
        const double facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
        const double atoli = a.atol, rtoli = a.rtol;
        double hnew;
        {
            if (nstep > a.nmax) return -2;
            if (0.1 * fabs(h) <= fabs(x) * a.uround) return -3;
            if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = 1; }
            nstep++;

            // the twelve stages (dop853.cpp:369-409)
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) yy1[i] = y[i] + h * a21 * k1[i];
            rhs(x + c2 * h, yy1, k2);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) yy1[i] = y[i] + h * (a31 * k1[i] + a32 * k2[i]);
            rhs(x + c3 * h, yy1, k3);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) yy1[i] = y[i] + h * (a41 * k1[i] + a43 * k3[i]);
            rhs(x + c4 * h, yy1, k4);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) yy1[i] = y[i] + h * (a51 * k1[i] + a53 * k3[i] + a54 * k4[i]);
            rhs(x + c5 * h, yy1, k5);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) yy1[i] = y[i] + h * (a61 * k1[i] + a64 * k4[i] + a65 * k5[i]);
            rhs(x + c6 * h, yy1, k6);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) yy1[i] = y[i] + h * (a71 * k1[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
            rhs(x + c7 * h, yy1, k7);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++)
                yy1[i] = y[i] + h * (a81 * k1[i] + a84 * k4[i] + a85 * k5[i] + a86 * k6[i] + a87 * k7[i]);
            rhs(x + c8 * h, yy1, k8);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++)
                yy1[i] = y[i] + h * (a91 * k1[i] + a94 * k4[i] + a95 * k5[i] + a96 * k6[i] + a97 * k7[i] + a98 * k8[i]);
            rhs(x + c9 * h, yy1, k9);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++)
                yy1[i] = y[i] + h * (a101 * k1[i] + a104 * k4[i] + a105 * k5[i] + a106 * k6[i] + a107 * k7[i] +
                                     a108 * k8[i] + a109 * k9[i]);
            rhs(x + c10 * h, yy1, k10);
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++)
                yy1[i] = y[i] + h * (a111 * k1[i] + a114 * k4[i] + a115 * k5[i] + a116 * k6[i] + a117 * k7[i] +
                                     a118 * k8[i] + a119 * k9[i] + a1110 * k10[i]);
            rhs(x + c11 * h, yy1, k2);
            const double xph = x + h;
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++)
                yy1[i] = y[i] + h * (a121 * k1[i] + a124 * k4[i] + a125 * k5[i] + a126 * k6[i] + a127 * k7[i] +
                                     a128 * k8[i] + a129 * k9[i] + a1210 * k10[i] + a1211 * k2[i]);
            rhs(xph, yy1, k3);
            nfcn += 11;
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) {
                k4[i] = b1 * k1[i] + b6 * k6[i] + b7 * k7[i] + b8 * k8[i] + b9 * k9[i] + b10 * k10[i] + b11 * k2[i] +
                        b12 * k3[i];
                k5[i] = y[i] + h * k4[i];
            }

            // error estimation (dop853.cpp:416-444), scalar tolerances, norm over this orbit's 6 components
            double err = 0.0, err2 = 0.0;
#pragma unroll(GB_NU)
            for (int i = 0; i < nn; i++) {
                const double sk = atoli + rtoli * gb_max(fabs(y[i]), fabs(k5[i]));
                double erri = k4[i] - bhh1 * k1[i] - bhh2 * k9[i] - bhh3 * k3[i];
#if GB_STRICT
                double sqr = erri / sk;
                err2 += sqr * sqr;
                erri = er1 * k1[i] + er6 * k6[i] + er7 * k7[i] + er8 * k8[i] + er9 * k9[i] + er10 * k10[i] +
                       er11 * k2[i] + er12 * k3[i];
                sqr = erri / sk;
#else
                const double isk = gb_rcp(sk);          // one reciprocal for the two error components
                double sqr = erri * isk;
                err2 += sqr * sqr;
                erri = er1 * k1[i] + er6 * k6[i] + er7 * k7[i] + er8 * k8[i] + er9 * k9[i] + er10 * k10[i] +
                       er11 * k2[i] + er12 * k3[i];
                sqr = erri * isk;
#endif
                err += sqr * sqr;
            }
            double deno = err + 0.01 * err2;
            if (deno <= 0.0) deno = 1.0;
#if GB_STRICT
            err = fabs(h) * err * sqrt(1.0 / (deno * (double)nn));
#else
            err = fabs(h) * err * gb_rsqrt(deno * (double)nn);
#endif

            // step-size controller (dop853.cpp:446-452), beta = 0 => pow(facold, beta) == 1
            const double fac11 = gb_pow_eighth(err);
            double fac = fac11;
            fac = gb_max(facc2, gb_min(facc1, fac / safe));
#if GB_STRICT
            hnew = h / fac;
#else
            hnew = h * gb_rcp(fac);
#endif

            if (err <= 1.0) {
                // accepted
                facold = gb_max(err, 1.0E-4);
                naccpt++;
                rhs(xph, k5, k4);
                nfcn++;

                // stiffness detection as coded in the reference (dop853.cpp:460-485)
                if (!(naccpt % a.nstiff)) {
                    double stnum = 0.0, stden = 0.0;
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++) {
                        double sqr = k4[i] - k3[i]; stnum += sqr * sqr;
                        sqr = k5[i] - yy1[i]; stden += sqr * sqr;
                    }
                    if (stden > 0.0) hlamb = h * sqrt(stnum / stden);
                    if (hlamb > 6.1) return -4;
                }

                if constexpr (DENSE && gb_is_deferred<OUT>::value) {
                    // dense-output preparation (dop853.cpp:492-582), deferred form.  Register pressure peaks here
                    // (y, k1, the new y and k1, k6..k9 for the three extra stages, the stage results, the gradient's
                    // temporaries): at 168 registers the compiler spilled ~50 values per accepted step to local
                    // memory, which no longer fits L1 next to the 160 KB of shared memory (long-scoreboard stalls,
                    // profiles/ncu_r2_dop853.txt).  So the OUT object also serves as explicit spill space: rcont5..8
                    // go there as partial sums before the extra stages, k6..k9 are parked in the rows that rcont1..4
                    // take only at the very end, and the extra stages read them back.  Same expressions in the same
                    // order as the immediate form below.
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++) {
                        emit.st(24 + i, d41 * k1[i] + d46 * k6[i] + d47 * k7[i] + d48 * k8[i] + d49 * k9[i] + d410 * k10[i] +
                                        d411 * k2[i] + d412 * k3[i]);
                        emit.st(30 + i, d51 * k1[i] + d56 * k6[i] + d57 * k7[i] + d58 * k8[i] + d59 * k9[i] + d510 * k10[i] +
                                        d511 * k2[i] + d512 * k3[i]);
                        emit.st(36 + i, d61 * k1[i] + d66 * k6[i] + d67 * k7[i] + d68 * k8[i] + d69 * k9[i] + d610 * k10[i] +
                                        d611 * k2[i] + d612 * k3[i]);
                        emit.st(42 + i, d71 * k1[i] + d76 * k6[i] + d77 * k7[i] + d78 * k8[i] + d79 * k9[i] + d710 * k10[i] +
                                        d711 * k2[i] + d712 * k3[i]);
                        emit.st(i, k6[i]); emit.st(6 + i, k7[i]); emit.st(12 + i, k8[i]); emit.st(18 + i, k9[i]);
                    }
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++)
                        yy1[i] = y[i] + h * (a141 * k1[i] + a147 * emit.ld(6 + i) + a148 * emit.ld(12 + i) + a149 * emit.ld(18 + i) +
                                             a1410 * k10[i] + a1411 * k2[i] + a1412 * k3[i] + a1413 * k4[i]);
                    rhs(x + c14 * h, yy1, k10);
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++)
                        yy1[i] = y[i] + h * (a151 * k1[i] + a156 * emit.ld(i) + a157 * emit.ld(6 + i) + a158 * emit.ld(12 + i) +
                                             a1511 * k2[i] + a1512 * k3[i] + a1513 * k4[i] + a1514 * k10[i]);
                    rhs(x + c15 * h, yy1, k2);
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++)
                        yy1[i] = y[i] + h * (a161 * k1[i] + a166 * emit.ld(i) + a167 * emit.ld(6 + i) + a168 * emit.ld(12 + i) +
                                             a169 * emit.ld(18 + i) + a1613 * k4[i] + a1614 * k10[i] + a1615 * k2[i]);
                    rhs(x + c16 * h, yy1, k3);
                    nfcn += 3;
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++) {
                        emit.st(24 + i, h * (emit.ld(24 + i) + d413 * k4[i] + d414 * k10[i] + d415 * k2[i] + d416 * k3[i]));
                        emit.st(30 + i, h * (emit.ld(30 + i) + d513 * k4[i] + d514 * k10[i] + d515 * k2[i] + d516 * k3[i]));
                        emit.st(36 + i, h * (emit.ld(36 + i) + d613 * k4[i] + d614 * k10[i] + d615 * k2[i] + d616 * k3[i]));
                        emit.st(42 + i, h * (emit.ld(42 + i) + d713 * k4[i] + d714 * k10[i] + d715 * k2[i] + d716 * k3[i]));
                        // rcont1..4 last: their rows held k6..k9 until here
                        emit.st(i, y[i]);
                        const double ydiff = k5[i] - y[i];
                        emit.st(6 + i, ydiff);
                        const double bspl = h * k1[i] - ydiff;
                        emit.st(12 + i, bspl);
                        emit.st(18 + i, ydiff - h * k4[i] - bspl);
                    }
#if GB_STRICT
                    emit.stash(x, h, 0.0);
#else
                    emit.stash(x, h, 1.0 / h);            // one division per step, not one per output sample
#endif
                } else if constexpr (DENSE) {
                    // dense-output preparation (dop853.cpp:492-582)
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++) {
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
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++)
                        yy1[i] = y[i] + h * (a141 * k1[i] + a147 * k7[i] + a148 * k8[i] + a149 * k9[i] + a1410 * k10[i] +
                                             a1411 * k2[i] + a1412 * k3[i] + a1413 * k4[i]);
                    rhs(x + c14 * h, yy1, k10);
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++)
                        yy1[i] = y[i] + h * (a151 * k1[i] + a156 * k6[i] + a157 * k7[i] + a158 * k8[i] + a1511 * k2[i] +
                                             a1512 * k3[i] + a1513 * k4[i] + a1514 * k10[i]);
                    rhs(x + c15 * h, yy1, k2);
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++)
                        yy1[i] = y[i] + h * (a161 * k1[i] + a166 * k6[i] + a167 * k7[i] + a168 * k8[i] + a169 * k9[i] +
                                             a1613 * k4[i] + a1614 * k10[i] + a1615 * k2[i]);
                    rhs(x + c16 * h, yy1, k3);
                    nfcn += 3;
#pragma unroll(GB_NU)
                    for (int i = 0; i < nn; i++) {
                        rc5[i] = h * (rc5[i] + d413 * k4[i] + d414 * k10[i] + d415 * k2[i] + d416 * k3[i]);
                        rc6[i] = h * (rc6[i] + d513 * k4[i] + d514 * k10[i] + d515 * k2[i] + d516 * k3[i]);
                        rc7[i] = h * (rc7[i] + d613 * k4[i] + d614 * k10[i] + d615 * k2[i] + d616 * k3[i]);
                        rc8[i] = h * (rc8[i] + d713 * k4[i] + d714 * k10[i] + d715 * k2[i] + d716 * k3[i]);
                    }
                    // fill every requested time inside [x, x+h] (dop853.cpp:584-612; contd8 :869-904)
                    const double x0 = x, x1 = x0 + h;
#if !GB_STRICT
                    const double ih = 1.0 / h;            // one division per step, not one per output sample
#endif
                    while (out_idx < ntout) {
                        const double t_out = tout[out_idx];
                        if ((x0 <= t_out && t_out <= x1) || (x1 <= t_out && t_out <= x0)) {
#if GB_STRICT
                            const double s = (t_out - x0) / h;
#else
                            const double s = (t_out - x0) * ih;
#endif
                            const double s1 = 1.0 - s;
                            double v[n];
#pragma unroll(GB_NU)
                            for (int i = 0; i < nn; i++)
                                v[i] = rc1[i] + s * (rc2[i] + s1 * (rc3[i] + s * (rc4[i] + s1 * (rc5[i] + s * (rc6[i] + s1 * (rc7[i] + s * rc8[i]))))));
                            emit(out_idx, v);
                            out_idx++;
                        } else {
                            break;
                        }
                    }
                }

#pragma unroll(GB_NU)
                for (int i = 0; i < nn; i++) { k1[i] = k4[i]; y[i] = k5[i]; }
                x = xph;
                if (last) return 1;
                if (fabs(hnew) > hmax) hnew = posneg * hmax;
                if (reject) hnew = posneg * gb_min(fabs(hnew), fabs(h));
                reject = 0;
            } else {
                // rejected (dop853.cpp:638-645)
#if GB_STRICT
                hnew = h / gb_min(facc1, fac11 / safe);
#else
                hnew = h * gb_rcp(gb_min(facc1, fac11 / safe));
#endif
                reject = 1;
                if (naccpt >= 1) nrejct = nrejct + 1;
                last = 0;
            }
            h = hnew;
        }
        return 0;
    }
};

// Thread-per-orbit driver (used by the mock-stream kernel and as the reference semantics of the
// persistent kernel): run one lane to completion.
template <bool DENSE, int NDIM = 6, class RHS, class OUT>
GB_DEV int dop853_integrate(const RHS& rhs, const OUT& emit, const Dop853Args& a, double x, double xend,
                            double (&y)[NDIM], double h, const double* __restrict__ tout, int ntout,
                            int& out_idx, int& nstep_, int& naccpt_, int& nrejct_, int& nfcn_, int nrun = NDIM) {
    Dop853Lane<DENSE, NDIM> L;
    constexpr int NU = (NDIM <= GB_D8_UNROLL_MAX) ? NDIM : 1;
    L.nrun = nrun;
#pragma unroll(NU)
    for (int i = 0; i < NDIM; i++) L.y[i] = y[i];
    L.init(rhs, a, x, xend, h);
    int code;
    do { code = L.step(rhs, emit, a, tout, ntout); } while (code == 0);
#pragma unroll(NU)
    for (int i = 0; i < NDIM; i++) y[i] = L.y[i];
    out_idx = L.out_idx; nstep_ = L.nstep; naccpt_ = L.naccpt; nrejct_ = L.nrejct; nfcn_ = L.nfcn;
    return code;
}

// ------------------------------------------------------------------------------------------------
// dop853_integrate_hamiltonian (integrate/cyintegrators/dop853.pyx:196-250): (6,N) in, (6,ntimes,N)
// dense output or (6,N) final state -- as a persistent lane-refill kernel.
//
// Per-orbit step control makes the work per orbit vary by more than 10x (an orbit at 4 kpc takes
// ~10x the steps of one at 50 kpc), so a fixed thread = orbit mapping leaves most lanes of a warp
// idle most of the time (first version: 9.2 of 32 lanes active on average,
// profiles/ncu_r1_dop853_mw2022_v0.txt).  Here warps are persistent: whenever a lane's orbit ends
// the lane takes the next orbit index from a global queue, so every lane has work until the queue
// is empty.  The queue order is `perm` (orbits sorted by dynamical time on the device, k_dyn_time
// + radix sort): lanes of a warp then work on orbits with similar step sizes, which keeps the
// per-step dense-output loop (samples of the caller's grid inside [x, x+h]) balanced too.
//
// Dense output goes to an ORBIT-MAJOR scratch array [orbit - orb0][ntimes][6]: one sample is 48 contiguous
// bytes and consecutive samples of a lane are adjacent, so the scattered per-lane stores still fill
// whole 32-byte sectors; k_transpose_dense then rewrites it as the caller's (6, ntimes, N) with
// fully coalesced reads and writes.  (Writing (6,ntimes,N) directly from per-lane cursors cost 8x
// the algorithmic DRAM traffic in the first version.)
// ------------------------------------------------------------------------------------------------
// Dense output is evaluated by the WARP, not by the lane that owns the step (round 2): in the first version every
// lane looped over the caller's times inside its own accepted step, so a warp iterated max-over-lanes times
// with mean-over-lanes useful work (22.8 of 32 lanes active per instruction over the whole kernel,
// profiles/ncu_r1_dop853_mw2022_v2_benchsize.txt) and held rcont1..8 (48 doubles) in registers across three RHS
// evaluations (230-255 registers, 8 warps per SM).  Now step() deposits the eight coefficient vectors in the
// warp's shared-memory block ([row][lane], conflict-free for the writer and for readers of distinct or equal
// owners), and after the step -- outside the accept branch, all 32 lanes converged -- the warp pools its lanes'
// pending samples: sample s of the pool is evaluated by lane s mod 32 from the owner's column, so the loop runs
// ceil(total / 32) times with every lane busy, and consecutive lanes write consecutive 48-byte records of the
// same orbit row.  The numbers are the owner's: same coefficients, same (t - x0) / h, same Horner order.
#ifndef GB_D8_DENSE_MINB
#define GB_D8_DENSE_MINB 3       // CTAs of GB_D8_MAXT threads per SM the dense kernel is compiled for (168 registers)
#endif
#ifndef GB_D8_MAXT
#define GB_D8_MAXT 128           // largest CTA of k_dop853_dyn (A/B builds: 192 x 2, 384 x 1 keep 12 warps per SM)
#endif
#define GB_WD_RC 0               // rcont[48][32]
#define GB_WD_X0 1536            // x0[32]
#define GB_WD_H 1568             // h[32]
#define GB_WD_IH 1600            // 1/h[32] (fast build)
#define GB_WD_ROW 1632           // row pointer[32] (as 64-bit words)
#define GB_WD_BEG 1664           // int out_idx[32]
#define GB_WD_PRE 1680           // int exclusive prefix of the sample counts[32]
#define GB_WD_DOUBLES 1696       // 13,568 bytes per warp

struct WarpDenseOut {
    static constexpr bool kDeferred = true;
    double* col;                 // this lane's column of the warp block
    int* pending;                // set by stash(): this lane completed an accepted step in this iteration
    GB_DEV void st(int row, double v) const { col[row * 32] = v; }
    GB_DEV double ld(int row) const { return col[row * 32]; }
    GB_DEV void stash(double x0, double h, double ih) const {
        col[GB_WD_X0] = x0; col[GB_WD_H] = h; col[GB_WD_IH] = ih;
        *pending = 1;
    }
};

// All 32 lanes of the warp call this after every attempted step.  pending: this lane deposited a step.
// Returns the number of samples of THIS lane that were written (its out_idx advances by that much).
GB_DEV int warp_dense_flush(double* wb, unsigned lane, int pending, int out_idx, double* srow,
                            const double* __restrict__ tout, int ntout, double inv_spacing) {
    int cnt = 0;
    if (pending && out_idx < ntout) {
        // Every requested time inside [x, x+h], scanning from the cursor and stopping at the first miss
        // (dop853.cpp:584-612).  The caller's grid is monotonic (integrate/timespec.py), so that run of hits is an
        // interval: its end is guessed from the mean grid spacing and corrected with the reference's own inclusion
        // test on the actual grid values -- the same count as the linear scan, in 2-3 loads instead of up to
        // max-over-lanes (measured: 17 dependent L1 loads per warp-step for a mean of 3.6 samples per lane,
        // profiles/ncu_r2_dop853_source_regions.txt).
        const double x0 = wb[GB_WD_X0 + lane], x1 = x0 + wb[GB_WD_H + lane];
        auto inside = [&](double t_out) { return (x0 <= t_out && t_out <= x1) || (x1 <= t_out && t_out <= x0); };
        const double tf = tout[out_idx];
        if (inside(tf)) {
            const double guess = fmin(fabs(x1 - tf) * inv_spacing, (double)(ntout - out_idx));
            int e = out_idx + 1 + (int)guess;             // exclusive end of the run, to be corrected
            if (e > ntout) e = ntout;
            while (e < ntout && inside(tout[e])) e++;
            while (e > out_idx + 1 && !inside(tout[e - 1])) e--;
            cnt = e - out_idx;
        }
    }
    if (!__any_sync(0xffffffffu, cnt > 0)) return 0;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int* beg = reinterpret_cast<int*>(wb + GB_WD_BEG);
    int* pre = reinterpret_cast<int*>(wb + GB_WD_PRE);
    beg[lane] = out_idx;
    pre[lane] = incl - cnt;
    reinterpret_cast<double**>(wb + GB_WD_ROW)[lane] = srow;
    __syncwarp();
    for (int s0 = 0; s0 < total; s0 += 32) {
        const int s = s0 + (int)lane;
        if (s < total) {
            // owner = the last lane whose exclusive prefix is <= s (lanes without samples share their successor's prefix)
            int o = 0;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) if (pre[o + d] <= s) o += d;
            const int idx = beg[o] + (s - pre[o]);
            const double t_out = tout[idx];
#if GB_STRICT
            const double sg = (t_out - wb[GB_WD_X0 + o]) / wb[GB_WD_H + o];
#else
            const double sg = (t_out - wb[GB_WD_X0 + o]) * wb[GB_WD_IH + o];
#endif
            const double s1 = 1.0 - sg;
            const double* rc = wb + GB_WD_RC + o;
            double v[6];
#pragma unroll
            for (int i = 0; i < 6; i++)
                v[i] = rc[i * 32] + sg * (rc[(6 + i) * 32] + s1 * (rc[(12 + i) * 32] + sg * (rc[(18 + i) * 32] + s1 * (rc[(24 + i) * 32] +
                       sg * (rc[(30 + i) * 32] + s1 * (rc[(36 + i) * 32] + sg * rc[(42 + i) * 32]))))));
            double2* op = reinterpret_cast<double2*>(reinterpret_cast<double**>(wb + GB_WD_ROW)[o] + (size_t)idx * 6);
            __stcs(op, make_double2(v[0], v[1])); __stcs(op + 1, make_double2(v[2], v[3])); __stcs(op + 2, make_double2(v[4], v[5]));
        }
    }
    __syncwarp();            // the block is rewritten by the next step
    return cnt;
}

// Register budget (A/B on B200, 303,104 MW2022 orbits): round 1, coefficients in registers: 255 registers /
// 2 CTAs of 128 per SM 34.1 ms, capped at 168 / 3 CTAs 34.8 ms, at 128 / 4 CTAs 38.9 ms (profiles/c2_ab_r1.txt);
// the final-state kernel gains from 3 CTAs per SM (14.5 -> 13.3 ms).  Round 2 (coefficients in shared memory):
// see profiles/c2_ab_r2.txt.
template <class C, bool ROT, bool DENSE>
__global__ void __launch_bounds__(GB_D8_MAXT, DENSE ? GB_D8_DENSE_MINB : (GB_D8_MAXT == 128 ? 3 : GB_D8_DENSE_MINB))
k_dop853_dyn(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F, const __grid_constant__ Dop853Args a,
             const double* __restrict__ w0, size_t N, const double* __restrict__ t, int ntimes,
             const uint32_t* __restrict__ perm, unsigned long long* __restrict__ queue,
             size_t orb0, size_t nslots, double* __restrict__ out, Dop853Stats st, int block_sync) {
    // This launch integrates the orbits [orb0, orb0 + nslots); perm (length nslots, global orbit
    // indices of that range in queue order) may be null = natural order.
    extern __shared__ double gb_d8_smem[];       // DENSE: one GB_WD_DOUBLES block per warp
    auto rhs = [&](double tt, const double (&w)[6], double (&f)[6]) { ham_rhs<C, ROT>(P, F, tt, w, f); };
    Dop853Lane<DENSE> L;
    const unsigned lane = threadIdx.x & 31u;
    const double t0 = t[0], tend = t[ntimes - 1];
    // mean samples per unit time of the caller's grid (first guess of warp_dense_flush); 0 for a degenerate grid
    const double inv_spacing = (tend != t0) ? (double)(ntimes - 1) / fabs(tend - t0) : 0.0;
    bool active = false, drained = false;
    size_t orb = 0;          // orbit index (column of w0 / of the caller's output)
    double* srow = nullptr;  // DENSE: this orbit's [ntimes][6] block of the scratch array
    double* wb = gb_d8_smem + (size_t)(threadIdx.x >> 5) * GB_WD_DOUBLES;
    int pending = 0;
    WarpDenseOut emit{wb + lane, &pending};
    while (true) {
        if (!drained) {
            const unsigned need = __ballot_sync(0xffffffffu, !active);
            if (need) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(queue, (unsigned long long)__popc(need));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + __popc(need) >= nslots) drained = true;       // warp-uniform
                if (!active) {
                    const unsigned long long k = base + __popc(need & ((1u << lane) - 1u));
                    if (k < nslots) {
                        orb = perm ? (size_t)perm[k] : orb0 + (size_t)k;
#pragma unroll
                        for (int c = 0; c < 6; c++) L.y[c] = w0[c * N + orb];
                        if (DENSE) srow = out + (orb - orb0) * (size_t)ntimes * 6;
                        L.init(rhs, a, t0, tend, a.h0);
                        active = true;
                    }
                }
            }
        }
        // block_sync: the warps of a CTA start every attempted step together, so they walk the same
        // ~90 KB of straight-line code at the same time and share its instruction-cache misses
        // (the L1.5 I-cache is 32 KB; unsynchronised warps each stream the whole loop body from L2).
        if (block_sync) { if (!__syncthreads_or(active)) break; }
        else if (!__any_sync(0xffffffffu, active)) break;
        int code = 0;
        pending = 0;
        if (active) code = L.step(rhs, emit, a, t, ntimes);
        if (DENSE) {
            __syncwarp();
            L.out_idx += warp_dense_flush(wb, lane, pending, L.out_idx, srow, t, ntimes, inv_spacing);
        }
        if (active && code != 0) {
            if (DENSE) {
                // rows the integration did not reach.  code 1 (success): only the last requested time can be missing,
                // when x + (xend - x) rounds an ulp short of xend -- the state there IS the final state (the reference
                // leaves that row of its np.empty output unwritten).  code < 0: the reference leaves the remaining rows
                // undefined as well; NaN here, and the per-orbit status says why.
                const double nan = CUDART_NAN;
                for (int j = L.out_idx; j < ntimes; j++)
#pragma unroll
                    for (int c = 0; c < 6; c++) srow[(size_t)j * 6 + c] = (code == 1) ? L.y[c] : nan;
            } else {
#pragma unroll
                for (int c = 0; c < 6; c++) out[c * N + orb] = L.y[c];
            }
            if (st.status) st.status[orb] = code;
            if (st.nstep) st.nstep[orb] = L.nstep;
            if (st.naccpt) st.naccpt[orb] = L.naccpt;
            if (st.nrejct) st.nrejct[orb] = L.nrejct;
            if (st.nfcn) st.nfcn[orb] = L.nfcn;
            active = false;
        }
    }
}

// scratch [nslots][ntimes][6] (row k = orbit orb0 + k) -> the caller's out (6, ntimes, N).
// One block = 32 orbits x TT times: reads are contiguous runs of TT*6 doubles per orbit; each warp
// store is one 256-byte segment out[c][j][orb0+k0 .. +31].
template <int TT>
__global__ void __launch_bounds__(256)
k_transpose_dense(const double* __restrict__ scratch, size_t orb0, size_t nslots,
                  int ntimes, size_t N, double* __restrict__ out) {
    __shared__ double tile[32][TT * 6 + 1];
    const size_t k0 = (size_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * TT;
    const int nt = min(TT, ntimes - j0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s = warp; s < 32; s += 8) {
        const size_t k = k0 + s;
        if (k >= nslots) break;
        const double* src = scratch + (k * ntimes + j0) * 6;
        for (int e = lane; e < nt * 6; e += 32) tile[s][e] = __ldcs(src + e);
    }
    __syncthreads();
    const size_t k = k0 + lane;
    if (k >= nslots) return;
    const size_t orb = orb0 + k;
    const size_t TS = (size_t)ntimes * N;
    for (int e = warp; e < nt * 6; e += 8) {
        const int jj = e / 6, c = e - jj * 6;
        __stcs(out + c * TS + (size_t)(j0 + jj) * N + orb, tile[lane][e]);
    }
}

// sort key of the orbit queue: dynamical time sqrt(r / |grad Phi|) at the initial position, as a
// float (monotone bit pattern for positive floats, so it radix-sorts as an unsigned integer).
template <class C>
__global__ void k_dyn_time(const __grid_constant__ DevPot P, const double* __restrict__ w0, size_t N, double t0,
                           size_t orb0, size_t n, float* __restrict__ key, uint32_t* __restrict__ idx) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t i = orb0 + k;
    const double x = w0[i], y = w0[N + i], z = w0[2 * N + i];
    double gx, gy, gz;
    C::gradient(P, t0, x, y, z, gx, gy, gz);
    const double r2 = x * x + y * y + z * z, g2 = gx * gx + gy * gy + gz * gz;
    const float v = (float)sqrt(sqrt(r2 / g2));
    key[k] = (v == v) ? v : 0.f;      // NaN (r = 0, Null potential) sorts first
    idx[k] = (uint32_t)i;
}
