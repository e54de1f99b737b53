// mockstream.cuh -- mock-stream kernels: particle release (Fardal+15 DF) and the batched
// integration of all released particles to the final time, one particle per thread.
//
// Reference: dynamics/mockstream/df.pyx:61-106 (get_rj_vj_R, transform_from_sat), :363-456
// (FardalStreamDF._sample), dynamics/mockstream/_coord.pyx:23-48, c_d2_dr2
// (potential/potential/src/cpotential.cpp:346-371), mockstream_dop853 / mockstream_leapfrog
// (dynamics/mockstream/mockstream.pyx:176-303, 442-620) for the case without massive bodies:
// every stream particle is then an independent test particle, which is what makes the
// one-thread-per-particle batch valid.  Rows are AoS (Np,6) as in the reference.
#pragma once

// One thread per stream particle.  The random deviates of particle p are the row normals[p * ncols ..]
// drawn on the host in the reference's RNG order; sign = +1 trailing / -1 leading.  kind selects the DF:
//   0 Fardal+15      (df.pyx:363-456)  row [kx, z, vt, vz] ~ N(k_mean, k_disp); flag = gala_modified
//   1 Streakline     (df.pyx:242-318)  no deviates: released at +-rj with +-vj
//   2 LagrangeCloud  (df.pyx:460-552)  row [vx, vy, vz] ~ N(0, v_disp)
//   3 Chen+24        (df.pyx:556-702)  row [r, phi, theta, v, alpha, beta] ~ N(mean, cov), angles in degrees
template <class C>
__global__ void k_fardal_release(const __grid_constant__ DevPot P, double G, const double* __restrict__ prog_w,
                                 const double* __restrict__ prog_t, const double* __restrict__ prog_m, int ntimes,
                                 const int32_t* __restrict__ prog_idx, const double* __restrict__ sign,
                                 const double* __restrict__ normals, int ncols, size_t Np, int kind, int gala_modified,
                                 double* __restrict__ out) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Np) return;
    const int it = prog_idx[p];
    const double* pw = prog_w + (size_t)it * 6;
    const double px[3] = {pw[0], pw[1], pw[2]}, pv[3] = {pw[3], pw[4], pw[5]};
    const double t = prog_t[it], m = prog_m[it];

    // get_rj_vj_R (df.pyx:61-92)
    double val = 0.;
    val += px[0] * px[0]; val += px[1] * px[1]; val += px[2] * px[2];
    const double dist = sqrt(val);
    double L[3];
    L[0] = px[1] * pv[2] - px[2] * pv[1];
    L[1] = -px[0] * pv[2] + px[2] * pv[0];
    L[2] = px[0] * pv[1] - px[1] * pv[0];
    val = 0.;
    val += L[0] * L[0]; val += L[1] * L[1]; val += L[2] * L[2];
    const double Lnorm = sqrt(val);
    double R[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) { R[0][i] = px[i] / dist; R[2][i] = L[i] / Lnorm; }
    const double Om = Lnorm / (dist * dist);
    // c_d2_dr2 (cpotential.cpp:346-371), h = 1e-2
    double d2r;
    {
        const double h = 1E-2;
        double r2 = 0;
        r2 = r2 + px[0] * px[0]; r2 = r2 + px[1] * px[1]; r2 = r2 + px[2] * px[2];
        const double r = sqrt(r2);
        double e[3];
#pragma unroll
        for (int j = 0; j < 3; j++) e[j] = px[j] + h * px[j] / r;
        d2r = C::value(P, t, e[0], e[1], e[2]);
        d2r = d2r - 2. * C::value(P, t, px[0], px[1], px[2]);
#pragma unroll
        for (int j = 0; j < 3; j++) e[j] = px[j] - h * px[j] / r;
        d2r = d2r + C::value(P, t, e[0], e[1], e[2]);
        d2r = d2r / (h * h);
    }
    const double rj = pow(G * m / (Om * Om - d2r), 1 / 3.);
    const double vj = Om * rj;
    // R[1] = -(R[0] x R[2])
    R[1][0] = -(R[0][1] * R[2][2] - R[0][2] * R[2][1]);
    R[1][1] = -(-R[0][0] * R[2][2] + R[0][2] * R[2][0]);
    R[1][2] = -(R[0][0] * R[2][1] - R[0][1] * R[2][0]);

    // FardalStreamDF._sample body (df.pyx:405-454)
    const double sg = sign[p];
    const double srj = (sg < 0) ? -rj : rj, svj = (sg < 0) ? -vj : vj;
    const double* nr = normals + p * (size_t)ncols;
    double tx[3] = {0., 0., 0.}, tv[3] = {0., 0., 0.};
    if (kind == 0) {
        const double kx = nr[0];
        tx[0] = kx * srj; tx[2] = nr[1] * srj;
        tv[1] = nr[2] * svj; tv[2] = nr[3] * svj;
        if (gala_modified) tv[1] *= kx;
    } else if (kind == 1) {
        tx[0] = srj; tv[1] = svj;
    } else if (kind == 2) {
        tx[0] = srj;
        tv[0] = nr[0]; tv[1] = nr[1]; tv[2] = nr[2];
    } else {
        // ChenStreamDF (df.pyx:631-697): leading particles are rotated by pi in phi and alpha
        const double Dr = nr[0] * rj;
        const double Dv = nr[3] * sqrt(2 * G * m / Dr);
        double a1 = nr[1] * (GB_PI / 180), a4 = nr[4] * (GB_PI / 180);
        if (sg < 0) { a1 = a1 + GB_PI; a4 = a4 + GB_PI; }
        const double a2 = nr[2] * (GB_PI / 180), a5 = nr[5] * (GB_PI / 180);
        tx[0] = Dr * cos(a2) * cos(a1); tx[1] = Dr * cos(a2) * sin(a1); tx[2] = Dr * sin(a2);
        tv[0] = Dv * cos(a5) * cos(a4); tv[1] = Dv * cos(a5) * sin(a4); tv[2] = Dv * sin(a5);
    }
    // transform_from_sat (df.pyx:94-106): R^T . x + prog
    double* o = out + p * 6;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double ox = R[0][i] * tx[0] + R[1][i] * tx[1] + R[2][i] * tx[2];
        double ov = R[0][i] * tv[0] + R[1][i] * tv[1] + R[2][i] * tv[2];
        ox += px[i]; ov += pv[i];
        o[i] = ox; o[3 + i] = ov;
    }
}

// mockstream_dop853 without massive bodies (mockstream.pyx:259-283): particle p integrates from
// t1[p] to tfinal with dop853_step's settings (dop853.pyx:27-75; set by the host in `a`).
template <class C, bool ROT>
__global__ void __launch_bounds__(128)
k_mock_dop853(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F,
              const __grid_constant__ Dop853Args a, const double* __restrict__ w0, const double* __restrict__ t1,
              size_t Np, double tfinal, double* __restrict__ out, int32_t* __restrict__ status) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Np) return;
    double y[6];
#pragma unroll
    for (int k = 0; k < 6; k++) y[k] = w0[p * 6 + k];
    auto rhs = [&](double tt, const double (&w)[6], double (&f)[6]) { ham_rhs<C, ROT>(P, F, tt, w, f); };
    auto emit = [&](int, const double (&)[6]) {};
    int out_idx = 0, nstep, naccpt, nrejct, nfcn;
    const int code = dop853_integrate<false>(rhs, emit, a, t1[p], tfinal, y, a.h0, nullptr, 0, out_idx, nstep,
                                             naccpt, nrejct, nfcn);
#pragma unroll
    for (int k = 0; k < 6; k++) out[p * 6 + k] = y[k];
    if (status) status[p] = code;
}

// mockstream_dop853_animate without massive bodies (mockstream.pyx:306-440): the stream is marched over the
// caller's time grid one interval at a time -- every interval is a fresh dop853_step call (initial step dt0
// again, dop853.pyx:27-75) -- and the state of every released particle is stored every `output_every`
// intervals (and at the end).  ridx[p] = index of particle p's release time in t; before that its snapshot
// rows are NaN (the HDF5 fill value, mockstream.pyx:133-140).  snap rows: (nout, Np, 6).
template <class C, bool ROT>
__global__ void __launch_bounds__(128)
k_mock_dop853_animate(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F,
                      const __grid_constant__ Dop853Args a, const double* __restrict__ w0,
                      const int32_t* __restrict__ ridx, size_t Np, const double* __restrict__ t, int ntimes,
                      int output_every, double* __restrict__ snap, double* __restrict__ out,
                      int32_t* __restrict__ status) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Np) return;
    double y[6];
#pragma unroll
    for (int k = 0; k < 6; k++) y[k] = w0[p * 6 + k];
    auto rhs = [&](double tt, const double (&w)[6], double (&f)[6]) { ham_rhs<C, ROT>(P, F, tt, w, f); };
    auto emit = [&](int, const double (&)[6]) {};
    const int k0 = ridx[p];
    const double nan = CUDART_NAN;
    int code = 1, j = 0;
    auto store = [&](int jj, bool live) {
        double* o = snap + ((size_t)jj * Np + p) * 6;
#pragma unroll
        for (int k = 0; k < 6; k++) o[k] = live ? y[k] : nan;
    };
    store(0, k0 == 0);
    for (int i = 1; i < ntimes; i++) {
        if (i > k0 && code > 0) {
            int out_idx = 0, nstep, naccpt, nrejct, nfcn;
            code = dop853_integrate<false>(rhs, emit, a, t[i - 1], t[i], y, a.h0, nullptr, 0, out_idx, nstep, naccpt,
                                           nrejct, nfcn);
        }
        if ((i % output_every) == 0 || i == ntimes - 1) { j++; store(j, i >= k0 && code > 0); }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) out[p * 6 + k] = y[k];
    if (status) status[p] = code;
}

// mockstream_leapfrog without massive bodies (mockstream.pyx:556-590 with c_init_velocity_nbody /
// c_leapfrog_step_nbody, integrate/cyintegrators/leapfrog.pyx:126-158).
template <class C>
__global__ void __launch_bounds__(256)
k_mock_leapfrog(const __grid_constant__ DevPot P, const double* __restrict__ w0, const double* __restrict__ t1,
                size_t Np, double tfinal, double dt, double* __restrict__ out) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Np) return;
    double x = w0[p * 6], y = w0[p * 6 + 1], z = w0[p * 6 + 2];
    double vx = w0[p * 6 + 3], vy = w0[p * 6 + 4], vz = w0[p * 6 + 5];
    const double ts = t1[p];
    const int n_steps = (int)((tfinal - ts) / dt + 0.5);
    double gx, gy, gz;
    C::gradient(P, ts, x, y, z, gx, gy, gz);
    double hx = vx - gx * dt / 2., hy = vy - gy * dt / 2., hz = vz - gz * dt / 2.;
    for (int j = 0; j < n_steps; j++) {
        x = x + hx * dt; y = y + hy * dt; z = z + hz * dt;
        C::gradient(P, ts + (j + 1) * dt, x, y, z, gx, gy, gz);
        vx = hx - gx * dt / 2.; vy = hy - gy * dt / 2.; vz = hz - gz * dt / 2.;
        hx = hx - gx * dt; hy = hy - gy * dt; hz = hz - gz * dt;
    }
    double* o = out + p * 6;
    o[0] = x; o[1] = y; o[2] = z; o[3] = vx; o[4] = vy; o[5] = vz;
}
