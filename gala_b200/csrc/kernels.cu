// kernels.cu -- sm_100a FP64 kernels of the orbit-integration hot path.
//
// One thread = one orbit; phase-space state lives in registers for the whole time loop; global
// memory is touched only for the initial load, the (optional) per-step trajectory store and the
// final state.  All arrays are structure-of-arrays (6,N) / (6,ntimes,N) like the reference
// (integrate/cyintegrators/leapfrog.pyx:57-59,92), so a warp's load/store of one phase-space
// component is one contiguous 256-byte segment.
//
// Compiled twice (see kernels.h): GB_NS = gbk_fast | gbk_strict.
#include <math_constants.h>
#include <stdlib.h>
#include "kernels.h"

#ifndef GB_NS
#error "GB_NS must be defined (gbk_fast or gbk_strict)"
#endif

// Everything below -- device functions included -- lives in the per-build namespace, so the
// strict and the fast translation units never share a (weak) template symbol.
namespace GB_NS {
#include "composite.cuh"
#include "hamiltonian.cuh"
#include "dop853.cuh"
#include "mockstream.cuh"
#if GB_PART == 4
#include "hessian.cuh"
#endif
#if GB_PART == 5 || GB_PART == 6
#include "nbody.cuh"
#endif
#if GB_PART == 6
#include "lyapunov.cuh"
#endif
#if GB_PART == 7 || GB_PART == 8
#include "extrema.cuh"
#endif

#if GB_PART == 1
// ------------------------------------------------------------------------------------------------
// batched evaluation kernels (CPotentialWrapper.gradient/energy/density, cpotential.pyx:104-162;
// Hamiltonian.energy/gradient, hamiltonian/src/chamiltonian.cpp:7-57)
// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void k_eval_gradient(const __grid_constant__ DevPot P, const double* __restrict__ q, double t,
                                size_t N, double* __restrict__ g) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double gx, gy, gz;
    C::gradient(P, t, q[i], q[N + i], q[2 * N + i], gx, gy, gz);
    g[i] = gx; g[N + i] = gy; g[2 * N + i] = gz;
}
template <class C, int WHAT>
__global__ void k_eval_scalar(const __grid_constant__ DevPot P, const double* __restrict__ q, double t,
                              size_t N, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    out[i] = (WHAT == 0) ? C::value(P, t, q[i], q[N + i], q[2 * N + i])
                         : C::density(P, t, q[i], q[N + i], q[2 * N + i]);
}

// diagnostic: the math primitives of this build, one per `which` (tests/test_gpu_fastmath.py)
__global__ void k_math_probe(int which, const double* __restrict__ x, size_t N, double* __restrict__ y) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double v = x[i];
#if GB_STRICT
    y[i] = which == 0 ? 1.0 / v : which == 1 ? 1.0 / sqrt(v) : which == 2 ? gb_pow_m1p5(v) : log(v);
#else
    y[i] = which == 0 ? gb_rcp(v) : which == 1 ? gb_rsqrt(v) : which == 2 ? gb_pow_m1p5(v) : gb_log(v);
#endif
}

template <class C>
__global__ void k_ham_energy(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F,
                             const double* __restrict__ w, double t, size_t N, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double x = w[i], y = w[N + i], z = w[2 * N + i];
    out[i] = C::value(P, t, x, y, z) + frame_energy(F, x, y, z, w[3 * N + i], w[4 * N + i], w[5 * N + i]);
}
template <class C, bool ROT>
__global__ void k_ham_gradient(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F,
                               const double* __restrict__ w, double t, size_t N, double* __restrict__ f) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double ww[6], ff[6];
#pragma unroll
    for (int k = 0; k < 6; k++) ww[k] = w[k * N + i];
    ham_rhs<C, ROT>(P, F, t, ww, ff);
#pragma unroll
    for (int k = 0; k < 6; k++) f[k * N + i] = ff[k];
}

// ------------------------------------------------------------------------------------------------
// Leapfrog (leapfrog_integrate_hamiltonian, integrate/cyintegrators/leapfrog.pyx:54-121;
// c_init_velocity :24-32; c_leapfrog_step :35-51).  Stored state at step j is (x, v) with the
// synchronised velocity; the half-step velocity v12 is carried in registers.
// Trajectory rows are written with streaming stores: each warp store is 256 contiguous bytes of
// out[k][j][i0..i0+31], never re-read by the kernel, so it should not occupy L2.
// ------------------------------------------------------------------------------------------------
template <class C, bool SAVE>
__global__ void __launch_bounds__(C::kFixedStepMaxThreads, C::kFixedStepMinBlocks)
k_leapfrog(const __grid_constant__ DevPot P, const double* __restrict__ w0, size_t N,
           const double* __restrict__ t, int ntimes, double dt, int dt_from_t, double* __restrict__ out,
           const TiState* __restrict__ ti_tab) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (dt_from_t) dt = t[1] - t[0];       // DEVICE-mode calls: the grid lives on the device (capi.cu:launch_fixed)
    double x = w0[i], y = w0[N + i], z = w0[2 * N + i];
    double vx = w0[3 * N + i], vy = w0[4 * N + i], vz = w0[5 * N + i];
    const size_t TS = (size_t)ntimes * N;  // stride between phase-space components when SAVE
    if (SAVE) {
        __stcs(out + i, x); __stcs(out + TS + i, y); __stcs(out + 2 * TS + i, z);
        __stcs(out + 3 * TS + i, vx); __stcs(out + 4 * TS + i, vy); __stcs(out + 5 * TS + i, vz);
    }
    double gx, gy, gz;
    // time-dependent composites: row j of ti_tab = the TimeInterpolated components at t[j] (k_ti_table below)
    const int nti = C::kTimeDependent ? ti_count(P) : 0;
    if constexpr (C::kTimeDependent) C::gradient_row(P, ti_tab, x, y, z, gx, gy, gz);
    else C::gradient(P, t[0], x, y, z, gx, gy, gz);
    double hx = vx - gx * dt / 2., hy = vy - gy * dt / 2., hz = vz - gz * dt / 2.;
    // The synchronised velocity is only formed where it is read: every step when SAVE, else on the
    // last step (same expression, same operands as c_leapfrog_step, leapfrog.pyx:35-51).
#pragma unroll 1
    for (int j = 1; j < ntimes; j++) {
        x = x + hx * dt; y = y + hy * dt; z = z + hz * dt;
        if constexpr (C::kTimeDependent) C::gradient_row(P, ti_tab + (size_t)j * nti, x, y, z, gx, gy, gz);
        else C::gradient(P, 0., x, y, z, gx, gy, gz);                             // c_leapfrog_step(..., t[j], ...) leapfrog.pyx:106
        if (SAVE || j == ntimes - 1) { vx = hx - gx * dt / 2.; vy = hy - gy * dt / 2.; vz = hz - gz * dt / 2.; }
        hx = hx - gx * dt; hy = hy - gy * dt; hz = hz - gz * dt;
        if (SAVE) {
            double* o = out + (size_t)j * N + i;
            __stcs(o, x); __stcs(o + TS, y); __stcs(o + 2 * TS, z);
            __stcs(o + 3 * TS, vx); __stcs(o + 4 * TS, vy); __stcs(o + 5 * TS, vz);
        }
    }
    if (!SAVE) {
        out[i] = x; out[N + i] = y; out[2 * N + i] = z;
        out[3 * N + i] = vx; out[4 * N + i] = vy; out[5 * N + i] = vz;
    }
}

// ------------------------------------------------------------------------------------------------
// Ruth4 (ruth4_integrate_hamiltonian, integrate/cyintegrators/ruth4.pyx:37-113; step :24-35).
// ROT=true reproduces the reference's Python Ruth4Integrator in a ConstantRotatingFrame
// (integrate/pyintegrators/ruth4.py:106-124 with F = Hamiltonian._gradient,
// hamiltonian/chamiltonian.pyx:88-99): only F[3:] = -(grad + Omega x p) is used.
// Sub-stage 0 has d_0 = 0: its gradient is multiplied by zero in the reference, so it is not
// evaluated here (3 gradient evaluations per step).
// ------------------------------------------------------------------------------------------------
struct Ruth4Coef { double c[4], d[4]; };

template <class C, bool ROT, bool SAVE>
__global__ void __launch_bounds__(C::kFixedStepMaxThreads, C::kFixedStepMinBlocks)
k_ruth4(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F, const __grid_constant__ Ruth4Coef K,
        const double* __restrict__ w0, size_t N, const double* __restrict__ t, int ntimes, double dt, int dt_from_t,
        double* __restrict__ out, const TiState* __restrict__ ti_tab) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (dt_from_t) dt = t[1] - t[0];
    double x = w0[i], y = w0[N + i], z = w0[2 * N + i];
    double vx = w0[3 * N + i], vy = w0[4 * N + i], vz = w0[5 * N + i];
    const size_t TS = (size_t)ntimes * N;
    if (SAVE) {
        __stcs(out + i, x); __stcs(out + TS + i, y); __stcs(out + 2 * TS + i, z);
        __stcs(out + 3 * TS + i, vx); __stcs(out + 4 * TS + i, vy); __stcs(out + 5 * TS + i, vz);
    }
    const int nti = C::kTimeDependent ? ti_count(P) : 0;
    for (int j = 1; j < ntimes; j++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (s > 0) {
                double gx, gy, gz;
                if constexpr (C::kTimeDependent) C::gradient_row(P, ti_tab + (size_t)j * nti, x, y, z, gx, gy, gz);
                else C::gradient(P, 0., x, y, z, gx, gy, gz);                        // c_ruth4_step(..., t[j], ...) ruth4.pyx:100
                if (!ROT) {
                    vx = vx - K.d[s] * gx * dt; vy = vy - K.d[s] * gy * dt; vz = vz - K.d[s] * gz * dt;
                } else {
                    const double Cx = F.om[1] * vz - F.om[2] * vy;
                    const double Cy = -F.om[0] * vz + F.om[2] * vx;
                    const double Cz = F.om[0] * vy - F.om[1] * vx;
                    const double ax = -(gx + Cx), ay = -(gy + Cy), az = -(gz + Cz);
                    vx = vx + K.d[s] * ax * dt; vy = vy + K.d[s] * ay * dt; vz = vz + K.d[s] * az * dt;
                }
            }
            x = x + K.c[s] * vx * dt; y = y + K.c[s] * vy * dt; z = z + K.c[s] * vz * dt;
        }
        if (SAVE) {
            double* o = out + (size_t)j * N + i;
            __stcs(o, x); __stcs(o + TS, y); __stcs(o + 2 * TS, z);
            __stcs(o + 3 * TS, vx); __stcs(o + 4 * TS, vy); __stcs(o + 5 * TS, vz);
        }
    }
    if (!SAVE) {
        out[i] = x; out[N + i] = y; out[2 * N + i] = z;
        out[3 * N + i] = vx; out[4 * N + i] = vy; out[5 * N + i] = vz;
    }
}

// State of every TimeInterpolated component at every time of the grid: thread = (step j, component i); row j holds
// the TimeInterpolated components only, in component order.
__global__ void k_ti_table(const __grid_constant__ DevPot P, const double* __restrict__ t, int ntimes, TiState* __restrict__ tab) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (size_t)ntimes * P.n) return;
    const int j = (int)(k / P.n), i = (int)(k % P.n);
    if (P.c[i].type != GB_POT_TIMEINTERP) return;
    int rank = 0;
    for (int q = 0; q < i; q++) rank += (P.c[q].type == GB_POT_TIMEINTERP);
    ti_fill_state(P, i, t[j], tab[(size_t)j * ti_count(P) + rank]);
}

#endif  // GB_PART == 1

// ------------------------------------------------------------------------------------------------
// host launchers.  GB_PART splits this file into translation units that compile in parallel:
//   1 = evaluation + leapfrog + ruth4, 2 = dop853 static frame, 3 = dop853 rotating frame,
//   4 = mock-stream kernels, 5 = N-body (massive bodies) kernels.
// ------------------------------------------------------------------------------------------------
static inline unsigned nblocks(size_t N, int block) { return (unsigned)((N + block - 1) / block); }

// the TimeInterpolated signature exists only where time reaches the potential (evaluation, fixed-step integrators,
// DOP853, mock streams without massive bodies, extrema: parts 1-4, 7, 8); capi.cu:resolve refuses time-dependent
// potentials for the other entry points (massive bodies, Lyapunov)
#if GB_PART == 1 || GB_PART == 2 || GB_PART == 3 || GB_PART == 4 || GB_PART == 7 || GB_PART == 8
#define GB_SIG_CASE_TI(CALL) case SIG_GENERIC_TI: { using C = Composite<SIG_GENERIC_TI>; CALL; } break;
#else
#define GB_SIG_CASE_TI(CALL)
#endif
#define GB_SIG_SWITCH(sig, CALL)                                  \
    switch (sig) {                                                \
        case SIG_NFW:        { using C = Composite<SIG_NFW>;        CALL; } break; \
        case SIG_HERNQUIST:  { using C = Composite<SIG_HERNQUIST>;  CALL; } break; \
        case SIG_MW2022:     { using C = Composite<SIG_MW2022>;     CALL; } break; \
        case SIG_BAR_MW2022: { using C = Composite<SIG_BAR_MW2022>; CALL; } break; \
        case SIG_MW2022_BAR: { using C = Composite<SIG_MW2022_BAR>; CALL; } break; \
        case SIG_SCF:        { using C = Composite<SIG_SCF>;        CALL; } break; \
        case SIG_MW_V1:      { using C = Composite<SIG_MW_V1>;      CALL; } break; \
        case SIG_LM10:       { using C = Composite<SIG_LM10>;       CALL; } break; \
        case SIG_BOVY2014:   { using C = Composite<SIG_BOVY2014>;   CALL; } break; \
        case SIG_GENERIC_LIGHT: { using C = Composite<SIG_GENERIC_LIGHT>; CALL; } break; \
        GB_SIG_CASE_TI(CALL) \
        default:             { using C = Composite<SIG_GENERIC>;    CALL; } break; \
    }

#if GB_PART == 1
cudaError_t eval_gradient(const DevPot& P, const double* q, double t, size_t N, double* g, int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    GB_SIG_SWITCH(P.sig, (k_eval_gradient<C><<<nblocks(N, block), block, 0, s>>>(P, q, t, N, g)));
    return cudaGetLastError();
}
cudaError_t math_probe(int which, const double* x, size_t N, double* y, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    k_math_probe<<<nblocks(N, 128), 128, 0, s>>>(which, x, N, y);
    return cudaGetLastError();
}
cudaError_t eval_energy(const DevPot& P, const double* q, double t, size_t N, double* out, int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    GB_SIG_SWITCH(P.sig, (k_eval_scalar<C, 0><<<nblocks(N, block), block, 0, s>>>(P, q, t, N, out)));
    return cudaGetLastError();
}
cudaError_t eval_density(const DevPot& P, const double* q, double t, size_t N, double* out, int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    GB_SIG_SWITCH(P.sig, (k_eval_scalar<C, 1><<<nblocks(N, block), block, 0, s>>>(P, q, t, N, out)));
    return cudaGetLastError();
}
cudaError_t ham_energy(const DevPot& P, const DevFrame& F, const double* w, double t, size_t N, double* out,
                       int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    GB_SIG_SWITCH(P.sig, (k_ham_energy<C><<<nblocks(N, block), block, 0, s>>>(P, F, w, t, N, out)));
    return cudaGetLastError();
}
cudaError_t ham_gradient(const DevPot& P, const DevFrame& F, const double* w, double t, size_t N, double* f,
                         int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    if (F.type == GB_FRAME_STATIC) {
        GB_SIG_SWITCH(P.sig, (k_ham_gradient<C, false><<<nblocks(N, block), block, 0, s>>>(P, F, w, t, N, f)));
    } else {
        GB_SIG_SWITCH(P.sig, (k_ham_gradient<C, true><<<nblocks(N, block), block, 0, s>>>(P, F, w, t, N, f)));
    }
    return cudaGetLastError();
}

size_t ti_table_bytes(const DevPot& P, int ntimes) { return (size_t)ntimes * ti_count(P) * sizeof(TiState); }

cudaError_t ti_table(const DevPot& P, const double* t, int ntimes, void* tab, cudaStream_t s) {
    const size_t n = (size_t)ntimes * P.n;
    if (n == 0) return cudaSuccess;
    k_ti_table<<<nblocks(n, 128), 128, 0, s>>>(P, t, ntimes, (TiState*)tab);
    return cudaGetLastError();
}

cudaError_t leapfrog(const DevPot& P, const double* w0, size_t N, const double* t, int ntimes, double dt,
                     int dt_from_t, int save_all, double* out, const void* ti_tab, int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    if (save_all) {
        GB_SIG_SWITCH(P.sig, (k_leapfrog<C, true><<<nblocks(N, block < C::kFixedStepMaxThreads ? block : C::kFixedStepMaxThreads), block < C::kFixedStepMaxThreads ? block : C::kFixedStepMaxThreads, 0, s>>>(P, w0, N, t, ntimes, dt, dt_from_t, out, (const TiState*)ti_tab)));
    } else {
        GB_SIG_SWITCH(P.sig, (k_leapfrog<C, false><<<nblocks(N, block < C::kFixedStepMaxThreads ? block : C::kFixedStepMaxThreads), block < C::kFixedStepMaxThreads ? block : C::kFixedStepMaxThreads, 0, s>>>(P, w0, N, t, ntimes, dt, dt_from_t, out, (const TiState*)ti_tab)));
    }
    return cudaGetLastError();
}

cudaError_t ruth4(const DevPot& P, const DevFrame& F, const double* w0, size_t N, const double* t, int ntimes,
                  double dt, int dt_from_t, const double* cs, const double* ds, int save_all, double* out,
                  const void* ti_tab, int block,
                  cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    Ruth4Coef K;
    for (int k = 0; k < 4; k++) { K.c[k] = cs[k]; K.d[k] = ds[k]; }
    const bool rot = F.type != GB_FRAME_STATIC;
#define GB_R4(ROT, SAVE) GB_SIG_SWITCH(P.sig, (k_ruth4<C, ROT, SAVE><<<nblocks(N, block < C::kFixedStepMaxThreads ? block : C::kFixedStepMaxThreads), block < C::kFixedStepMaxThreads ? block : C::kFixedStepMaxThreads, 0, s>>>(P, F, K, w0, N, t, ntimes, dt, dt_from_t, out, (const TiState*)ti_tab)))
    if (rot) { if (save_all) { GB_R4(true, true); } else { GB_R4(true, false); } }
    else     { if (save_all) { GB_R4(false, true); } else { GB_R4(false, false); } }
#undef GB_R4
    return cudaGetLastError();
}

#endif  // GB_PART == 1

#if GB_PART == 2 || GB_PART == 3
#if GB_PART == 2
#define GB_D8_NAME dop853_static
#define GB_D8_ROT false
#else
#define GB_D8_NAME dop853_rotating
#define GB_D8_ROT true
#endif
// Persistent launch: one CTA slot per resident block (occupancy x SM count), never more CTAs than
// there are 64-orbit groups.  `out` is the (6,N) final-state array, or for save_all the orbit-major
// scratch [nslots][ntimes][6] that dop853_transpose() turns into the caller's layout.
template <class C, bool DENSE>
static cudaError_t launch_d8(const DevPot& P, const DevFrame& F, const double* w0, size_t N, const double* t,
                             int ntimes, const Dop853Args& a, const uint32_t* perm, unsigned long long* queue,
                             size_t orb0, size_t nslots, double* out, const Dop853Stats& st, int block, int nsm,
                             int block_sync, cudaStream_t s) {
    auto kern = k_dop853_dyn<C, GB_D8_ROT, DENSE>;
    // DENSE: one block of dense-output coefficients per warp in shared memory (dop853.cuh: warp_dense_flush)
    const size_t smem = DENSE ? (size_t)(block / 32) * GB_WD_DOUBLES * sizeof(double) : 0;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem);
    if (e != cudaSuccess) return e;
    const size_t want = (nslots + block - 1) / block;
    const size_t cap = (size_t)(per_sm > 0 ? per_sm : 1) * nsm;
    kern<<<(unsigned)(want < cap ? want : cap), block, smem, s>>>(P, F, a, w0, N, t, ntimes, perm, queue, orb0, nslots,
                                                                   out, st, block_sync);
    return cudaGetLastError();
}
cudaError_t GB_D8_NAME(const DevPot& P, const DevFrame& F, const double* w0, size_t N, const double* t, int ntimes,
                       const Dop853Args& a, int save_all, const uint32_t* perm, unsigned long long* queue,
                       size_t orb0, size_t nslots, double* out, int32_t* status, int32_t* nstep,
                       int32_t* naccpt, int32_t* nrejct, int32_t* nfcn, int block, cudaStream_t s) {
    if (nslots == 0) return cudaSuccess;
    Dop853Stats st{status, nstep, naccpt, nrejct, nfcn};
    int dev = 0, nsm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (const char* eb = getenv("GB_D8_BLOCK")) block = atoi(eb);
    if (block > GB_D8_MAXT || block <= 0 || (block & 31)) block = 64;      // __launch_bounds__ of k_dop853_dyn
    const char* bs = getenv("GB_D8_BLOCKSYNC");
    const int block_sync = bs ? atoi(bs) : (block > 32);
    if (save_all) {
        GB_SIG_SWITCH(P.sig, (e = launch_d8<C, true>(P, F, w0, N, t, ntimes, a, perm, queue, orb0, nslots, out, st, block, nsm, block_sync, s)));
    } else {
        GB_SIG_SWITCH(P.sig, (e = launch_d8<C, false>(P, F, w0, N, t, ntimes, a, perm, queue, orb0, nslots, out, st, block, nsm, block_sync, s)));
    }
    return e;
}
#if GB_PART == 2
cudaError_t dop853_transpose(const double* scratch, size_t orb0, size_t nslots, int ntimes, size_t N, double* out,
                             cudaStream_t s) {
    if (nslots == 0) return cudaSuccess;
    constexpr int TT = 16;
    dim3 grid((unsigned)((nslots + 31) / 32), (unsigned)((ntimes + TT - 1) / TT));
    k_transpose_dense<TT><<<grid, 256, 0, s>>>(scratch, orb0, nslots, ntimes, N, out);
    return cudaGetLastError();
}
cudaError_t dyn_time_keys(const DevPot& P, const double* w0, size_t N, double t0, size_t orb0, size_t n, float* key,
                          uint32_t* idx, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    GB_SIG_SWITCH(P.sig, (k_dyn_time<C><<<nblocks(n, 256), 256, 0, s>>>(P, w0, N, t0, orb0, n, key, idx)));
    return cudaGetLastError();
}
#endif
#endif  // GB_PART == 2 || 3

#if GB_PART == 4
cudaError_t mock_dop853(const DevPot& P, const DevFrame& F, const double* w0_rows, const double* t1, size_t Np,
                        double tfinal, const Dop853Args& a, double* out_rows, int32_t* status, int block,
                        cudaStream_t s) {
    if (Np == 0) return cudaSuccess;
    const bool rot = F.type != GB_FRAME_STATIC;
    if (rot) {
        GB_SIG_SWITCH(P.sig, (k_mock_dop853<C, true><<<nblocks(Np, block), block, 0, s>>>(P, F, a, w0_rows, t1, Np, tfinal, out_rows, status)));
    } else {
        GB_SIG_SWITCH(P.sig, (k_mock_dop853<C, false><<<nblocks(Np, block), block, 0, s>>>(P, F, a, w0_rows, t1, Np, tfinal, out_rows, status)));
    }
    return cudaGetLastError();
}

cudaError_t mock_dop853_animate(const DevPot& P, const DevFrame& F, const double* w0_rows, const int32_t* ridx,
                                size_t Np, const double* t, int ntimes, const Dop853Args& a, int output_every,
                                double* snap, double* out_rows, int32_t* status, int block, cudaStream_t s) {
    if (Np == 0) return cudaSuccess;
    if (block > 128 || block <= 0) block = 64;
    const bool rot = F.type != GB_FRAME_STATIC;
    if (rot) {
        GB_SIG_SWITCH(P.sig, (k_mock_dop853_animate<C, true><<<nblocks(Np, block), block, 0, s>>>(P, F, a, w0_rows, ridx, Np, t, ntimes, output_every, snap, out_rows, status)));
    } else {
        GB_SIG_SWITCH(P.sig, (k_mock_dop853_animate<C, false><<<nblocks(Np, block), block, 0, s>>>(P, F, a, w0_rows, ridx, Np, t, ntimes, output_every, snap, out_rows, status)));
    }
    return cudaGetLastError();
}

cudaError_t mock_leapfrog(const DevPot& P, const double* w0_rows, const double* t1, size_t Np, double tfinal,
                          double dt, double* out_rows, int block, cudaStream_t s) {
    if (Np == 0) return cudaSuccess;
    GB_SIG_SWITCH(P.sig, (k_mock_leapfrog<C><<<nblocks(Np, block), block, 0, s>>>(P, w0_rows, t1, Np, tfinal, dt, out_rows)));
    return cudaGetLastError();
}

cudaError_t fardal_release(const DevPot& P, double G, const double* prog_w, const double* prog_t,
                           const double* prog_m, int ntimes, const int32_t* prog_idx, const double* sign,
                           const double* normals, int ncols, size_t Np, int kind, int gala_modified, double* out_rows,
                           int block, cudaStream_t s) {
    if (Np == 0) return cudaSuccess;
    GB_SIG_SWITCH(P.sig, (k_fardal_release<C><<<nblocks(Np, block), block, 0, s>>>(P, G, prog_w, prog_t, prog_m, ntimes, prog_idx, sign, normals, ncols, Np, kind, gala_modified, out_rows)));
    return cudaGetLastError();
}

cudaError_t eval_hessian(const DevPot& P, const double* q, double t, size_t N, double* hess, int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    k_eval_hessian<<<nblocks(N, block), block, 0, s>>>(P, q, t, N, hess);
    return cudaGetLastError();
}
#endif  // GB_PART == 4

#if GB_PART == 5 || GB_PART == 6
// Three composite forms are instantiated for the N-body kernels: the compile-time MW2022 list, the generic
// loop over analytic components, and the generic loop that also knows the basis-function expansions.
#define GB_SIG_SWITCH2(sig, CALL)                                 \
    switch (sig) {                                                \
        case SIG_MW2022:     { using C = Composite<SIG_MW2022>;     CALL; } break; \
        case SIG_GENERIC: case SIG_SCF: { using C = Composite<SIG_GENERIC>; CALL; } break; \
        default:             { using C = Composite<SIG_GENERIC_LIGHT>; CALL; } break; \
    }
#if GB_PART == 5
cudaError_t nbody_leapfrog(const DevPot& P, const DevBodies& B, int scheme, const double* cs, const double* ds,
                           const double* body_w0, const int32_t* group,
                           const double* w0, const double* t1, size_t Np, double t0, double tfinal, int nsteps_fixed,
                           double dt, double* out_p, double* out_b, size_t body_writer, double* traj, size_t ntot,
                           int block, cudaStream_t s) {
    const int hasp = Np > 0;
    const size_t nthreads = hasp ? Np : 1;
    if (block > 128 || block <= 0) block = 128;
    NbodyRuth4Coef rc;
    for (int k = 0; k < 4; k++) { rc.cs[k] = cs ? cs[k] : 0.; rc.ds[k] = ds ? ds[k] : 0.; }
    GB_SIG_SWITCH2(P.sig, (k_nbody_leapfrog<C><<<nblocks(nthreads, block), block, 0, s>>>(
        P, B, rc, scheme, body_w0, group, w0, t1, Np, hasp, t0, tfinal, nsteps_fixed, dt, out_p, out_b, body_writer, traj, ntot)));
    return cudaGetLastError();
}
#endif

template <class C, int NDIM>
static void launch_nbody_d8(const DevPot& P, const DevBodies& B, const Dop853Args& a, const double* body_w0,
                            const int32_t* group, const double* w0, const double* t1, size_t Np, int hasp,
                            const double* tgrid, int ntimes, double t0, double tfinal, double* out_p, double* out_b,
                            size_t body_writer, double* traj, size_t ntot, int32_t* status, cudaStream_t s) {
    const size_t nthreads = hasp ? Np : 1;
    const int block = 64;
    if (traj)
        k_nbody_dop853<C, NDIM, true><<<nblocks(nthreads, block), block, 0, s>>>(
            P, B, a, body_w0, group, w0, t1, Np, hasp, tgrid, ntimes, t0, tfinal, out_p, out_b, body_writer, traj, ntot, status);
    else
        k_nbody_dop853<C, NDIM, false><<<nblocks(nthreads, block), block, 0, s>>>(
            P, B, a, body_w0, group, w0, t1, Np, hasp, tgrid, ntimes, t0, tfinal, out_p, out_b, body_writer, traj, ntot, status);
}
// systems of 1-2 points (n = 6, 12: unrolled, state in registers) compile in part 5; larger ones run ONE
// kernel sized for GB_ND_MAX equations with a run-time count (rolled loops, state in local memory), part 6
#if GB_PART == 5
#define GB_ND_NAME nbody_dop853_small
#else
#define GB_ND_NAME nbody_dop853_big
#endif
cudaError_t GB_ND_NAME(const DevPot& P, const DevBodies& B, const Dop853Args& a, const double* body_w0,
                         const int32_t* group, const double* w0, const double* t1, size_t Np,
                         const double* tgrid, int ntimes, double t0, double tfinal, double* out_p, double* out_b,
                         size_t body_writer, double* traj, size_t ntot, int32_t* status, cudaStream_t s) {
    const int hasp = Np > 0;
    const int npts = B.nb + hasp;
#define GB_ND(N) GB_SIG_SWITCH2(P.sig, (launch_nbody_d8<C, N>(P, B, a, body_w0, group, w0, t1, Np, hasp, tgrid, ntimes, t0, tfinal, out_p, out_b, body_writer, traj, ntot, status, s)))
    switch (npts) {
#if GB_PART == 5
        case 1: GB_ND(6); break;
        case 2: GB_ND(12); break;
        default: return cudaErrorInvalidValue;
#else
        default: if (npts < 3 || npts > GB_MAXB + 1) return cudaErrorInvalidValue; GB_ND(GB_ND_MAX); break;
#endif
    }
#undef GB_ND
    return cudaGetLastError();
}
#if GB_PART == 6
cudaError_t nbody_dop853_march(const DevPot& P, const DevBodies& B, const Dop853Args& a, double* body_all,
                               const double* w0, const int32_t* ridx, size_t Np, int has_particle, const double* t,
                               int ntimes, int output_every, double* snap, double* out_p, double* out_b,
                               size_t body_writer, int32_t* status, cudaStream_t s) {
    const size_t nthreads = has_particle ? Np : 1;
    if (nthreads == 0) return cudaSuccess;
    const int block = 64;
    GB_SIG_SWITCH2(P.sig, (k_nbody_dop853_march<C><<<nblocks(nthreads, block), block, 0, s>>>(
        P, B, a, body_all, w0, ridx, Np, has_particle, t, ntimes, output_every, snap, out_p, out_b, body_writer, status)));
    return cudaGetLastError();
}
cudaError_t lyapunov(const DevPot& P, const DevFrame& F, const Dop853Args& a, const double* w0, const double* d0_vec,
                     size_t N, const double* t, int n_steps, double d0, int pullback, int noff, double* LEs, double* traj,
                     int32_t* status, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    const int block = 64;
    if (F.type != GB_FRAME_STATIC) {
        GB_SIG_SWITCH2(P.sig, (k_lyapunov<C, true><<<nblocks(N, block), block, 0, s>>>(P, F, a, w0, d0_vec, N, t, n_steps, d0, pullback, noff, LEs, traj, status)));
    } else {
        GB_SIG_SWITCH2(P.sig, (k_lyapunov<C, false><<<nblocks(N, block), block, 0, s>>>(P, F, a, w0, d0_vec, N, t, n_steps, d0, pullback, noff, LEs, traj, status)));
    }
    return cudaGetLastError();
}
#endif
#if GB_PART == 5
cudaError_t nbody_dop853(const DevPot& P, const DevBodies& B, const Dop853Args& a, const double* body_w0,
                         const int32_t* group, const double* w0, const double* t1, size_t Np,
                         const double* tgrid, int ntimes, double t0, double tfinal, double* out_p, double* out_b,
                         size_t body_writer, double* traj, size_t ntot, int32_t* status, cudaStream_t s) {
    return (B.nb + (Np > 0) <= 2)
        ? nbody_dop853_small(P, B, a, body_w0, group, w0, t1, Np, tgrid, ntimes, t0, tfinal, out_p, out_b, body_writer, traj, ntot, status, s)
        : nbody_dop853_big(P, B, a, body_w0, group, w0, t1, Np, tgrid, ntimes, t0, tfinal, out_p, out_b, body_writer, traj, ntot, status, s);
}
#endif
#endif  // GB_PART == 5 || 6

#if GB_PART == 7 || GB_PART == 8
// part 7 = the reductions without the Hamiltonian (ENERGY = false) + the extremum lists; part 8 = with it: two
// translation units so that the 11 signatures x 3 schemes compile side by side
#if GB_PART == 7
#define GB_EXT_ENERGY false
#define GB_EXT_SUFFIX(name) name##_e0
#else
#define GB_EXT_ENERGY true
#define GB_EXT_SUFFIX(name) name##_e1
#endif
cudaError_t GB_EXT_SUFFIX(trajectory_extrema)(const DevPot& P, const DevFrame& F, const double* w, const double* t, int ntimes,
                                              size_t N, double* stats, int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    GB_SIG_SWITCH(P.sig, (k_trajectory_extrema<C, GB_EXT_ENERGY><<<nblocks(N, block), block, 0, s>>>(P, F, w, t, ntimes, N, stats)));
    return cudaGetLastError();
}
#if GB_PART == 7
cudaError_t trajectory_extrema_list(const double* w, const double* t, int ntimes, size_t N, int kind, int kmax,
                                    double* vals, double* times, int32_t* counts, int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    k_trajectory_extrema_list<<<nblocks(N, block), block, 0, s>>>(w, t, ntimes, N, kind, kmax, vals, times, counts);
    return cudaGetLastError();
}
#endif
cudaError_t GB_EXT_SUFFIX(integrate_extrema)(const DevPot& P, const DevFrame& F, int scheme, const double* cs, const double* ds,
                                             const double* w0, size_t N, const double* t, int ntimes, double dt, int dt_from_t,
                                             double* wfin, double* stats, int block, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    Ruth4CoefE K;
    for (int k = 0; k < 4; k++) { K.c[k] = cs[k]; K.d[k] = ds[k]; }
#define GB_IE(SCH) GB_SIG_SWITCH(P.sig, (k_integrate_extrema<C, SCH, GB_EXT_ENERGY><<<nblocks(N, block < C::kFixedStepMaxThreads ? block : C::kFixedStepMaxThreads), block < C::kFixedStepMaxThreads ? block : C::kFixedStepMaxThreads, 0, s>>>(P, F, K, w0, N, t, ntimes, dt, dt_from_t, wfin, stats)))
    const int sch = scheme == 0 ? 0 : (F.type == GB_FRAME_STATIC ? 1 : 2);
    if (sch == 0) { GB_IE(0); } else if (sch == 1) { GB_IE(1); } else { GB_IE(2); }
#undef GB_IE
    return cudaGetLastError();
}
#endif  // GB_PART == 7 || 8

}  // namespace GB_NS
