// fastmath.cuh -- branch-free FP64 primitives of the `fast` build (GB_STRICT=0).
//
// FP64 division, sqrt, rsqrt and log have no hardware instruction on sm_100a: the CUDA math
// library expands each into a MUFU seed (MUFU.RCP64H / MUFU.RSQ64H, ~20 good bits) followed by
// DFMA refinement PLUS range checks, a slow-path CALL and BSSY/BSYNC reconvergence.  In the first
// version of k_leapfrog<MW2022> those expansions were 212 FP64 instructions and ~200 integer/
// branch instructions per orbit-step for 146 source-level flops (profiles/ncu_r1_leapfrog_mw2022_v0.txt).
// The hot path never sees zero, subnormal, infinite or negative arguments for these calls
// (radii of bound orbits in kpc), so the fast build uses the seed + a fixed refinement only.
// Out-of-domain inputs give NaN/Inf exactly where the reference's formula gives NaN/Inf
// (r = 0 for a cuspy potential), never a silently wrong finite number.
//
// Accuracy (measured on the device by tests/test_gpu_fastmath.py against numpy/IEEE):
//   gb_rcp, gb_rsqrt, gb_pow_m1p5: <= 2 ulp; gb_log: <= 3 ulp.
// The strict build does not include this file's fast bodies: it keeps IEEE div/sqrt and libm.
#pragma once

#if !GB_STRICT

GB_DEV double gb_rcp_seed(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
GB_DEV double gb_rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// 1/x: seed y (relative error e ~ 2^-20), then the cubic step y (1 + e + e^2): 3 DFMA.
GB_DEV double gb_rcp(double x) {
    const double y = gb_rcp_seed(x);
    const double e = fma(-x, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);
}

// x^-1/2: seed y, e = 1 - x y^2, y (1 + e/2 + 3 e^2/8): 5 FP64 instructions.
GB_DEV double gb_rsqrt(double x) {
    const double y = gb_rsqrt_seed(x);
    const double t = y * y;
    const double e = fma(-x, t, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double ey = e * y;
    return fma(p, ey, y);
}

// x^-3/2 directly from the rsqrt seed: y^3 (1 + 3e/2 + 15 e^2/8): 6 FP64 instructions
// (one fewer than rsqrt + cube, and one rounding fewer).
GB_DEV double gb_pow_m1p5(double x) {
    const double y = gb_rsqrt_seed(x);
    const double t = y * y;
    const double e = fma(-x, t, 1.0);
    const double y3 = t * y;
    const double p = fma(e, 1.875, 1.5);
    const double pe = p * e;
    return fma(y3, pe, y3);
}

// natural log of a positive normal double.  Argument reduction to m in [sqrt(1/2), sqrt(2)) with
// integer operations on the high word (ALU pipe), log m = 2s + s R(s^2), s = (m-1)/(m+1), with the
// classic degree-7 minimax R of fdlibm's e_log.c (|error| < 2^-58.45 on this interval).
// 17 FP64 instructions + 1 MUFU + 1 I2F.
GB_DEV double gb_log(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    hi += 0x3ff00000 - 0x3fe6a09e;
    const int ex = (hi >> 20) - 0x3ff;
    hi = (hi & 0x000fffff) + 0x3fe6a09e;
    const double m = __hiloint2double(hi, lo);
    const double f = m - 1.0;
    const double s = f * gb_rcp(m + 1.0);
    const double z = s * s;
    double R = 1.479819860511658591e-01;
    R = fma(R, z, 1.531383769920937332e-01);
    R = fma(R, z, 1.818357216161805012e-01);
    R = fma(R, z, 2.222219843214978396e-01);
    R = fma(R, z, 2.857142874366239149e-01);
    R = fma(R, z, 3.999999999940941908e-01);
    R = fma(R, z, 6.666666666666735130e-01);
    R = R * z;
    const double lm = fma(s, R, s + s);
    return fma((double)ex, 6.93147180559945286227e-01, lm);
}

#endif  // !GB_STRICT
