// lyapunov.cuh -- maximum Lyapunov exponent by the offset-orbit method, one lane per parent orbit.
//
// Reference: dop853_lyapunov_max / dop853_lyapunov_max_dont_save (dynamics/lyapunov/dop853_lyapunov.pyx:
// 22-118, 120-192).  A lane integrates the parent orbit and its `noff` offset orbits as ONE DOP853 system of
// 6 (1 + noff) equations (one step size, like the reference: Fwrapper over norbits, dop853.cpp:966-976),
// restarted on every interval of the time grid (dop853_step, dop853.pyx:27-75); every `pullback` intervals the
// separation of each offset orbit is recorded as ln(|d1| / d0) and the offset orbit is pulled back to distance
// d0 along d1 (:88-100).  The reference handles ONE parent orbit per call; N lanes make a chaos map.
#pragma once

template <class C, bool ROT, int NDIM>
struct OrbitsRhs {
    const DevPot& P;
    const DevFrame& F;
    int npts;
    __device__ __noinline__ void operator()(double tt, const double (&w)[NDIM], double (&f)[NDIM]) const {
        for (int i = 0; i < npts; i++) {
            double wi[6], fi[6];
#pragma unroll
            for (int k = 0; k < 6; k++) wi[k] = w[6 * i + k];
            ham_rhs<C, ROT>(P, F, tt, wi, fi);
#pragma unroll
            for (int k = 0; k < 6; k++) f[6 * i + k] = fi[k];
        }
    }
};

// w0 (N,6) rows; d0_vec (N, noff, 6) already scaled to length d0; t (n_steps); LEs (N, niter, noff) raw
// ln(|d1|/d0) with niter = n_steps / pullback; traj (N, n_steps, 1+noff, 6) or null.
template <class C, bool ROT>
__global__ void __launch_bounds__(64)
k_lyapunov(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F, const __grid_constant__ Dop853Args a,
           const double* __restrict__ w0, const double* __restrict__ d0_vec, size_t N, const double* __restrict__ t,
           int n_steps, double d0, int pullback, int noff, double* __restrict__ LEs, double* __restrict__ traj,
           int32_t* __restrict__ status) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    constexpr int NDIM = GB_ND_MAX;
    const int norb = 1 + noff, nrun = 6 * norb, niter = n_steps / pullback;
    double y[NDIM];
    for (int i = 0; i < NDIM; i++) y[i] = 0.;
    for (int k = 0; k < 6; k++) y[k] = w0[p * 6 + k];
    for (int i = 1; i < norb; i++)
        for (int k = 0; k < 6; k++) y[6 * i + k] = w0[p * 6 + k] + d0_vec[(p * noff + (i - 1)) * 6 + k];
    OrbitsRhs<C, ROT, NDIM> rhs{P, F, norb};
    auto emit = [&](int, const double (&)[NDIM]) {};
    double* tr = traj ? traj + p * (size_t)n_steps * nrun : nullptr;
    if (tr) for (int i = 0; i < nrun; i++) tr[i] = y[i];
    int jiter = 0, code = 1;
    for (int j = 1; j < n_steps; j++) {
        int out_idx = 0, nstep, naccpt, nrejct, nfcn;
        code = dop853_integrate<false, NDIM>(rhs, emit, a, t[j - 1], t[j], y, a.h0, nullptr, 0, out_idx, nstep, naccpt,
                                             nrejct, nfcn, nrun);
        if (code < 0) break;
        if (tr) for (int i = 0; i < nrun; i++) tr[(size_t)j * nrun + i] = y[i];
        if ((j % pullback) == 0) {
            for (int i = 1; i < norb; i++) {
                double d1[6], norm = 0.;
                for (int k = 0; k < 6; k++) { d1[k] = y[6 * i + k] - y[k]; norm = norm + d1[k] * d1[k]; }
                const double mag = sqrt(norm);
                if (jiter < niter) LEs[(p * niter + jiter) * noff + (i - 1)] = log(mag / d0);
                for (int k = 0; k < 6; k++) y[6 * i + k] = y[k] + d0 * d1[k] / mag;
            }
            jiter++;
        }
    }
    if (status) status[p] = code;
}
