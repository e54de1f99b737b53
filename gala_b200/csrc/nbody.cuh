// nbody.cuh -- test particles in an external potential PLUS a few massive bodies (direct N-body).
//
// Reference: c_nbody_acceleration / c_nbody_gradient_symplectic (potential/potential/src/cpotential.cpp:
// 389-442), Fwrapper_direct_nbody (integrate/cyintegrators/dopri/dop853.cpp:990-1006),
// leapfrog_integrate_nbody + c_init_velocity_nbody / c_leapfrog_step_nbody
// (integrate/cyintegrators/leapfrog.pyx:126-257), direct_nbody_dop853 (dynamics/nbody/nbody.pyx:30-115)
// and the massive-body branches of mockstream_dop853 / mockstream_leapfrog
// (dynamics/mockstream/mockstream.pyx:176-303, 442-620).
//
// Test particles do not act on anything, so they stay independent of each other: ONE LANE = ONE TEST
// PARTICLE + ITS OWN COPY OF THE nb MASSIVE BODIES (nb <= GB_MAXB).  Every lane advances the bodies with
// exactly the reference's arithmetic (including the leapfrog loop's sequential in-place update of the
// bodies, leapfrog.pyx:238-249) and then its particle, so no inter-thread exchange is needed and the
// result for a particle equals the reference's for a release group of one.  The redundant body update
// costs nb extra gradient evaluations per step per lane; it buys bit-identical bodies in every lane.
#pragma once

// gradient of body b's own potential at (x,y,z), the body sitting at (bx,by,bz): c_gradient with
// do_shift_rotate = 1 for every component (cpotential.cpp:246-277); returns f2 (zeroed first).
static __device__ __noinline__ void gb_body_gradient(const DevBodies& B, int b, double bx, double by, double bz,
                             double x, double y, double z, double& fx, double& fy, double& fz) {
    fx = 0.; fy = 0.; fz = 0.;
    for (int c = B.cbeg[b]; c < B.cbeg[b + 1]; c++) {
        const double* R = B.R[c];
        const double sx = x - bx, sy = y - by, sz = z - bz;
        const double X = R[0] * sx + R[1] * sy + R[2] * sz;
        const double Y = R[3] * sx + R[4] * sy + R[5] * sz;
        const double Z = R[6] * sx + R[7] * sy + R[8] * sz;
        double ax = 0., ay = 0., az = 0.;
        gb_comp_gradient<false>(B.type[c], &B.par[B.poff[c]], nullptr, X, Y, Z, ax, ay, az);      // body potentials are analytic only (capi.cu:resolve_bodies): the switch without SCF / multipole saves 2 KB of stack per lane
        fx += R[0] * ax + R[3] * ay + R[6] * az;
        fy += R[1] * ax + R[4] * ay + R[7] * az;
        fz += R[2] * ax + R[5] * ay + R[8] * az;
    }
}

// total gradient on point i of the system (i < nb: a body, i == nb: the lane's particle):
// c_gradient(external) then c_nbody_gradient_symplectic (sources j != i, non-Null), leapfrog.pyx:126-158
template <class C>
GB_DEV void gb_nbody_grad_symplectic(const DevPot& P, const DevBodies& B, double t, const double (*bq)[3], int i,
                                     double x, double y, double z, double& gx, double& gy, double& gz) {
    C::gradient(P, t, x, y, z, gx, gy, gz);
    for (int j = 0; j < B.nb; j++) {
        if (B.null_[j] || j == i) continue;
        double fx, fy, fz;
        gb_body_gradient(B, j, bq[j][0], bq[j][1], bq[j][2], x, y, z, fx, fy, fz);
        gx += fx; gy += fy; gz += fz;
    }
}

// Leapfrog.  Particle p starts at t1[p] (or t0 when t1 == null) from w0 row p with the bodies in the
// state body_w0[group[p]] and takes nsteps = int((tfinal - t1)/dt + 0.5) steps (mockstream.pyx:571) or
// `nsteps_fixed` (leapfrog_integrate_nbody).  traj != null: every step is stored as rows of
// (ntimes, ntot, 6) (leapfrog.pyx:249-252), the bodies by lane `body_writer`.
struct NbodyRuth4Coef { double cs[4], ds[4]; };   // ruth4.pyx:166-178, computed on the host

// scheme 0: leapfrog (above).  scheme 1: Ruth4 (c_ruth4_step_nbody, ruth4.pyx:116-136): no half-step
// velocity; every point takes its four sub-stages in one go while the others stay where they are.
template <class C>
__global__ void __launch_bounds__(128)
k_nbody_leapfrog(const __grid_constant__ DevPot P, const __grid_constant__ DevBodies B,
                 const __grid_constant__ NbodyRuth4Coef rc, int scheme,
                 const double* __restrict__ body_w0, const int32_t* __restrict__ group,
                 const double* __restrict__ w0, const double* __restrict__ t1, size_t Np, int has_particle,
                 double t0, double tfinal, int nsteps_fixed, double dt,
                 double* __restrict__ out_p, double* __restrict__ out_b, size_t body_writer,
                 double* __restrict__ traj, size_t ntot) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (has_particle ? Np : 1)) return;
    const int nb = B.nb;
    double bq[GB_MAXB][3], bv[GB_MAXB][3], bh[GB_MAXB][3];
    const double* bw = body_w0 + (group ? (size_t)group[p] : 0) * (size_t)nb * 6;
    for (int b = 0; b < nb; b++)
        for (int k = 0; k < 3; k++) { bq[b][k] = bw[b * 6 + k]; bv[b][k] = bw[b * 6 + 3 + k]; }
    double x = 0., y = 0., z = 0., vx = 0., vy = 0., vz = 0., hx = 0., hy = 0., hz = 0.;
    if (has_particle) {
        x = w0[p * 6]; y = w0[p * 6 + 1]; z = w0[p * 6 + 2];
        vx = w0[p * 6 + 3]; vy = w0[p * 6 + 4]; vz = w0[p * 6 + 5];
    }
    const double ts = t1 ? t1[p] : t0;
    const int n_steps = t1 ? (int)((tfinal - ts) / dt + 0.5) : nsteps_fixed;
    const bool wb = (p == body_writer);
    double gx, gy, gz;
    // c_init_velocity_nbody for the bodies in order, then the particle (mockstream.pyx:556-566)
    if (scheme == 0) for (int b = 0; b < nb; b++) {
        gb_nbody_grad_symplectic<C>(P, B, ts, bq, b, bq[b][0], bq[b][1], bq[b][2], gx, gy, gz);
        bh[b][0] = bv[b][0] - gx * dt / 2.; bh[b][1] = bv[b][1] - gy * dt / 2.; bh[b][2] = bv[b][2] - gz * dt / 2.;
    }
    if (scheme == 0 && has_particle) {
        gb_nbody_grad_symplectic<C>(P, B, ts, bq, nb, x, y, z, gx, gy, gz);
        hx = vx - gx * dt / 2.; hy = vy - gy * dt / 2.; hz = vz - gz * dt / 2.;
    }
    if (traj) {
        if (wb) for (int b = 0; b < nb; b++) for (int k = 0; k < 3; k++) { traj[b * 6 + k] = bq[b][k]; traj[b * 6 + 3 + k] = bv[b][k]; }
        if (has_particle) {
            double* o = traj + ((size_t)nb + p) * 6;
            o[0] = x; o[1] = y; o[2] = z; o[3] = vx; o[4] = vy; o[5] = vz;
        }
    }
    for (int j = 0; j < n_steps; j++) {
        const double tj = ts + (j + 1) * dt;
        if (scheme == 1) {
            for (int b = 0; b < nb; b++)
                for (int q = 0; q < 4; q++) {
                    gb_nbody_grad_symplectic<C>(P, B, tj, bq, b, bq[b][0], bq[b][1], bq[b][2], gx, gy, gz);
                    bv[b][0] = bv[b][0] - rc.ds[q] * gx * dt; bv[b][1] = bv[b][1] - rc.ds[q] * gy * dt; bv[b][2] = bv[b][2] - rc.ds[q] * gz * dt;
                    bq[b][0] = bq[b][0] + rc.cs[q] * bv[b][0] * dt; bq[b][1] = bq[b][1] + rc.cs[q] * bv[b][1] * dt; bq[b][2] = bq[b][2] + rc.cs[q] * bv[b][2] * dt;
                }
            if (has_particle)
                for (int q = 0; q < 4; q++) {
                    gb_nbody_grad_symplectic<C>(P, B, tj, bq, nb, x, y, z, gx, gy, gz);
                    vx = vx - rc.ds[q] * gx * dt; vy = vy - rc.ds[q] * gy * dt; vz = vz - rc.ds[q] * gz * dt;
                    x = x + rc.cs[q] * vx * dt; y = y + rc.cs[q] * vy * dt; z = z + rc.cs[q] * vz * dt;
                }
        }
        // bodies one after the other, each seeing the others where they are NOW (in-place update)
        if (scheme == 0) for (int b = 0; b < nb; b++) {
            bq[b][0] = bq[b][0] + bh[b][0] * dt; bq[b][1] = bq[b][1] + bh[b][1] * dt; bq[b][2] = bq[b][2] + bh[b][2] * dt;
            gb_nbody_grad_symplectic<C>(P, B, tj, bq, b, bq[b][0], bq[b][1], bq[b][2], gx, gy, gz);
            bv[b][0] = bh[b][0] - gx * dt / 2.; bv[b][1] = bh[b][1] - gy * dt / 2.; bv[b][2] = bh[b][2] - gz * dt / 2.;
            bh[b][0] = bh[b][0] - gx * dt; bh[b][1] = bh[b][1] - gy * dt; bh[b][2] = bh[b][2] - gz * dt;
        }
        if (scheme == 0 && has_particle) {
            x = x + hx * dt; y = y + hy * dt; z = z + hz * dt;
            gb_nbody_grad_symplectic<C>(P, B, tj, bq, nb, x, y, z, gx, gy, gz);
            vx = hx - gx * dt / 2.; vy = hy - gy * dt / 2.; vz = hz - gz * dt / 2.;
            hx = hx - gx * dt; hy = hy - gy * dt; hz = hz - gz * dt;
        }
        if (traj) {
            double* row = traj + (size_t)(j + 1) * ntot * 6;
            if (wb) for (int b = 0; b < nb; b++) for (int k = 0; k < 3; k++) { row[b * 6 + k] = bq[b][k]; row[b * 6 + 3 + k] = bv[b][k]; }
            if (has_particle) {
                double* o = row + ((size_t)nb + p) * 6;
                o[0] = x; o[1] = y; o[2] = z; o[3] = vx; o[4] = vy; o[5] = vz;
            }
        }
    }
    if (has_particle && out_p) {
        double* o = out_p + p * 6;
        o[0] = x; o[1] = y; o[2] = z; o[3] = vx; o[4] = vy; o[5] = vz;
    }
    if (wb && out_b)
        for (int b = 0; b < nb; b++) for (int k = 0; k < 3; k++) { out_b[b * 6 + k] = bq[b][k]; out_b[b * 6 + 3 + k] = bv[b][k]; }
}

// DOP853.  The lane's ODE system is [bodies..., particle] with n = NDIM = 6 (nb + has_particle) and the
// right-hand side of Fwrapper_direct_nbody (dop853.cpp:990-1006): hamiltonian_gradient per point (static
// frame), then c_nbody_acceleration (sources j < nb non-Null acting on every i != j, cpotential.cpp:389-415).
// One step size per lane: the reference shares it across a whole release group / the whole system
// (deviation 1).  DENSE: samples at the caller's times as rows of (ntimes, ntot, 6).
// Right-hand side.  n > 12: ONE out-of-line copy per kernel (the integrator calls it at 16 sites; the state
// lives in local memory anyway).  n <= 12: inlined so that the state can stay in registers, with the two
// expensive leaves -- the external gradient and the body gradient -- as out-of-line calls on scalars.
template <class C>
static __device__ __noinline__ void gb_ext_gradient_call(const DevPot& P, double t, double x, double y, double z,
                                                         double& gx, double& gy, double& gz) {
    C::gradient(P, t, x, y, z, gx, gy, gz);
}
template <class C, int NDIM>
struct NbodyRhsInline {
    const DevPot& P;
    const DevBodies& B;
    __device__ __forceinline__ void operator()(double tt, const double (&w)[NDIM], double (&f)[NDIM]) const {
        constexpr int npts = NDIM / 6;
#pragma unroll
        for (int i = 0; i < npts; i++) {
            double gx, gy, gz;
            gb_ext_gradient_call<C>(P, tt, w[6 * i], w[6 * i + 1], w[6 * i + 2], gx, gy, gz);
            f[6 * i] = w[6 * i + 3]; f[6 * i + 1] = w[6 * i + 4]; f[6 * i + 2] = w[6 * i + 5];
            f[6 * i + 3] = -gx; f[6 * i + 4] = -gy; f[6 * i + 5] = -gz;
        }
#pragma unroll
        for (int j = 0; j < npts; j++) {
            if (j >= B.nb || B.null_[j]) continue;
#pragma unroll
            for (int i = 0; i < npts; i++) {
                if (i == j) continue;
                double fx, fy, fz;
                gb_body_gradient(B, j, w[6 * j], w[6 * j + 1], w[6 * j + 2], w[6 * i], w[6 * i + 1], w[6 * i + 2], fx, fy, fz);
                f[6 * i + 3] += -fx; f[6 * i + 4] += -fy; f[6 * i + 5] += -fz;
            }
        }
    }
};
template <class C, int NDIM>
struct NbodyRhs {
    const DevPot& P;
    const DevBodies& B;
    int npts;                      // points in use (bodies + the particle); the arrays are sized for NDIM / 6
    __device__ __noinline__ void operator()(double tt, const double (&w)[NDIM], double (&f)[NDIM]) const {
        const int nb = B.nb;
        for (int i = 0; i < npts; i++) {
            double gx, gy, gz;
            C::gradient(P, tt, w[6 * i], w[6 * i + 1], w[6 * i + 2], gx, gy, gz);
            f[6 * i] = w[6 * i + 3]; f[6 * i + 1] = w[6 * i + 4]; f[6 * i + 2] = w[6 * i + 5];
            f[6 * i + 3] = -gx; f[6 * i + 4] = -gy; f[6 * i + 5] = -gz;
        }
        for (int j = 0; j < nb; j++) {
            if (B.null_[j]) continue;
            for (int i = 0; i < npts; i++) {
                if (i == j) continue;
                double fx, fy, fz;
                gb_body_gradient(B, j, w[6 * j], w[6 * j + 1], w[6 * j + 2], w[6 * i], w[6 * i + 1], w[6 * i + 2], fx, fy, fz);
                f[6 * i + 3] += -fx; f[6 * i + 4] += -fy; f[6 * i + 5] += -fz;
            }
        }
    }
};

template <class C, int NDIM, bool DENSE>
__global__ void __launch_bounds__(64)
k_nbody_dop853(const __grid_constant__ DevPot P, const __grid_constant__ DevBodies B, const __grid_constant__ Dop853Args a,
               const double* __restrict__ body_w0, const int32_t* __restrict__ group,
               const double* __restrict__ w0, const double* __restrict__ t1, size_t Np, int has_particle,
               const double* __restrict__ tgrid, int ntimes, double t0, double tfinal,
               double* __restrict__ out_p, double* __restrict__ out_b, size_t body_writer,
               double* __restrict__ traj, size_t ntot, int32_t* __restrict__ status) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (has_particle ? Np : 1)) return;
    const int nb = B.nb;
    constexpr bool FIXED = NDIM <= GB_D8_UNROLL_MAX;          // NDIM == 6 (nb + has_particle), state in registers
    const int nrun = FIXED ? NDIM : 6 * (nb + (has_particle ? 1 : 0));
    double y[NDIM];
    const double* bw = body_w0 + (group ? (size_t)group[p] : 0) * (size_t)nb * 6;
    // rows [bodies..., particle]
#pragma unroll(FIXED ? NDIM : 1)
    for (int i = 0; i < NDIM; i++) y[i] = (i < nb * 6) ? bw[i] : ((i < nrun) ? w0[p * 6 + (i - nb * 6)] : 0.);
    typename std::conditional<FIXED, NbodyRhsInline<C, NDIM>, NbodyRhs<C, NDIM>>::type rhs{P, B};
    if constexpr (!FIXED) rhs.npts = nrun / 6;
    const bool wb = (p == body_writer);
    auto emit = [&](int idx, const double (&v)[NDIM]) {
        double* row = traj + (size_t)idx * ntot * 6;
#pragma unroll(FIXED ? NDIM : 1)
        for (int i = 0; i < NDIM; i++) {
            if (i < nb * 6) { if (wb) row[i] = v[i]; }
            else if (i < nrun) row[((size_t)nb + p) * 6 + (i - nb * 6)] = v[i];
        }
    };
    int out_idx = 0, nstep, naccpt, nrejct, nfcn;
    const double ts = t1 ? t1[p] : t0;
    const int code = dop853_integrate<DENSE, NDIM>(rhs, emit, a, ts, tfinal, y, a.h0, tgrid, ntimes, out_idx, nstep,
                                                   naccpt, nrejct, nfcn, nrun);
#pragma unroll(FIXED ? NDIM : 1)
    for (int i = 0; i < NDIM; i++) {
        if (i < nb * 6) { if (wb && out_b) out_b[i] = y[i]; }
        else if (i < nrun && out_p) out_p[p * 6 + (i - nb * 6)] = y[i];
    }
    if (status) status[p] = code;
}


// mockstream_dop853_animate with massive bodies (mockstream.pyx:306-440): the system is marched over the
// caller's time grid interval by interval, every interval a fresh dop853_step call (initial step dt0), with
// snapshots every `output_every` intervals.  Two launches of this kernel:
//   has_particle = 0  one lane marches the bodies alone over the whole grid and writes their state at EVERY
//                     grid time to body_all (ntimes, nb, 6) and at the snapshot times to snap;
//   has_particle = 1  lane p starts at its release index ridx[p] from body_all[ridx[p]] and marches
//                     [bodies, its particle]; snapshot rows before the release are NaN.
// snap rows: (nout, nb + Np, 6).  The reference marches bodies and ALL released particles as one system with
// one step size; here a lane's system is the bodies + one particle (deviation 1).
template <class C>
__global__ void __launch_bounds__(64)
k_nbody_dop853_march(const __grid_constant__ DevPot P, const __grid_constant__ DevBodies B,
                     const __grid_constant__ Dop853Args a, double* __restrict__ body_all,
                     const double* __restrict__ w0, const int32_t* __restrict__ ridx, size_t Np, int has_particle,
                     const double* __restrict__ t, int ntimes, int output_every, double* __restrict__ snap,
                     double* __restrict__ out_p, double* __restrict__ out_b, size_t body_writer,
                     int32_t* __restrict__ status) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (has_particle ? Np : 1)) return;
    constexpr int NDIM = GB_ND_MAX;
    const int nb = B.nb;
    const size_t ntot = (size_t)nb + Np;
    const int k0 = has_particle ? ridx[p] : 0;
    const int nrun = 6 * (nb + (has_particle ? 1 : 0));
    const bool wb = has_particle ? (p == body_writer) : true;
    const double nan = CUDART_NAN;
    double y[NDIM];
    for (int i = 0; i < NDIM; i++) y[i] = 0.;
    const double* bw = body_all + (size_t)k0 * nb * 6;
    for (int i = 0; i < nb * 6; i++) y[i] = bw[i];
    if (has_particle) for (int k = 0; k < 6; k++) y[nb * 6 + k] = w0[p * 6 + k];
    NbodyRhs<C, NDIM> rhs{P, B, nrun / 6};
    auto emit = [&](int, const double (&)[NDIM]) {};
    auto store = [&](int jj, bool live) {
        double* row = snap + (size_t)jj * ntot * 6;
        if (!has_particle) for (int i = 0; i < nb * 6; i++) row[i] = y[i];
        else for (int k = 0; k < 6; k++) row[((size_t)nb + p) * 6 + k] = live ? y[nb * 6 + k] : nan;
    };
    int code = 1, j = 0;
    store(0, k0 == 0);
    for (int i = 1; i < ntimes; i++) {
        if (i > k0 && code > 0) {
            int out_idx = 0, nstep, naccpt, nrejct, nfcn;
            code = dop853_integrate<false, NDIM>(rhs, emit, a, t[i - 1], t[i], y, a.h0, nullptr, 0, out_idx, nstep, naccpt,
                                                 nrejct, nfcn, nrun);
            if (!has_particle) for (int q = 0; q < nb * 6; q++) body_all[(size_t)i * nb * 6 + q] = y[q];
        }
        if ((i % output_every) == 0 || i == ntimes - 1) { j++; store(j, i >= k0 && code > 0); }
    }
    if (has_particle && out_p) for (int k = 0; k < 6; k++) out_p[p * 6 + k] = y[nb * 6 + k];
    if (wb && out_b) for (int i = 0; i < nb * 6; i++) out_b[i] = y[i];
    if (status) status[p] = code;
}
