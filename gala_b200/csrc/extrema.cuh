// extrema.cuh -- trajectory reductions on the device: Orbit.pericenter / Orbit.apocenter
// (reference dynamics/orbit.py:391-553) and the energy drift, so that a caller who only wants those numbers does
// not pull a (6, ntimes, N) trajectory over PCIe (C1 end to end is 30x below its device rate for that reason).
//
// The reference works per orbit on the saved samples r_j = |q(t_j)|:
//   ix = argrelmax(+-r)            interior samples strictly greater than both neighbours (orbit.py:402-403;
//                                  mode="wrap" only concerns the two edge samples, which :403 removes again)
//   for j in ix: polyfit(t[j-1:j+2], r[j-1:j+2], 2) -> vertex time -c1/(2 c2) and the value there (:413-422)
//   pericenter()/apocenter() reduce the refined values with func (default np.mean; np.min / np.max; or all).
// Here one thread walks one orbit's samples (either while integrating, k_integrate_extrema, or over a trajectory
// that already lies in device memory, k_trajectory_extrema) and keeps count / sum / min / max / first and last
// time of each kind.  The parabola through three points is formed in coordinates centred on t_j (the reference's
// polyfit solves the same 3x3 interpolation by least squares in absolute t; centred is the better-conditioned
// form of the same polynomial).
#pragma once

// rows of the statistics array (GB_EXT_NSTAT, N): enum gb_extrema_row of include/gala_b200.h

// the vertex of the parabola through (t_{j-1}, a), (t_j, b), (t_{j+1}, c), in coordinates centred on t_j
GB_DEV void ext_parabola(double tm2, double tm1, double tj, double a, double b, double c, double& tv, double& val) {
    const double da = tm2 - tm1, db = tj - tm1;
    const double sb = (c - b) / db, sa = (a - b) / da;
    const double c2 = (sb - sa) / (db - da);
    const double c1 = sb - c2 * db;
    const double tau = -c1 / (2. * c2);
    tv = tm1 + tau;
    val = b + tau * (c1 + c2 * tau);
}

// kinds: 0 pericentres (local minima of r), 1 apocentres (local maxima of r), 2 z-heights (local maxima of |z|,
// Orbit.zmax, dynamics/orbit.py:600-656)
struct ExtremaAcc {
    double rm2, rm1, zm2, zm1, tm2, tm1;      // the two previous samples of r and |z|
    int seen;
    double n[3], sum[3], mn[3], mx[3], tfirst[3], tlast[3];
    double zabs;
    GB_DEV void init() {
        seen = 0; rm2 = rm1 = zm2 = zm1 = tm2 = tm1 = 0.;
        for (int k = 0; k < 3; k++) { n[k] = 0.; sum[k] = 0.; mn[k] = CUDART_INF; mx[k] = -CUDART_INF; tfirst[k] = CUDART_NAN; tlast[k] = CUDART_NAN; }
        zabs = 0.;
    }
    GB_DEV void add(int k, double tv, double val) {
        if (n[k] == 0.) tfirst[k] = tv;
        tlast[k] = tv;
        n[k] += 1.; sum[k] += val; mn[k] = fmin(mn[k], val); mx[k] = fmax(mx[k], val);
    }
    // sample j: position (x, y, z) at time tj.  An extremum at sample j-1 is recognised when sample j arrives.
    // LIST (k_trajectory_extrema_list): hand every extremum of `kind` to `sink(value, time)` as well.
    template <class Sink>
    GB_DEV void push(double x, double y, double z, double tj, int kind, const Sink& sink) {
        const double r = sqrt(x * x + y * y + z * z), az = fabs(z);
        zabs = fmax(zabs, az);
        if (seen >= 2) {
            const bool is_max = (rm1 > rm2) && (rm1 > r);
            const bool is_min = (rm1 < rm2) && (rm1 < r);      // argrelmax(-r)
            double tv, val;
            if (is_max || is_min) {
                ext_parabola(tm2, tm1, tj, rm2, rm1, r, tv, val);
                const int k = is_max ? 1 : 0;
                add(k, tv, val);
                if (kind == k) sink(val, tv);
            }
            if ((zm1 > zm2) && (zm1 > az)) {
                ext_parabola(tm2, tm1, tj, zm2, zm1, az, tv, val);
                add(2, tv, val);
                if (kind == 2) sink(val, tv);
            }
        }
        rm2 = rm1; zm2 = zm1; tm2 = tm1; rm1 = r; zm1 = az; tm1 = tj; seen++;
    }
    GB_DEV void push(double x, double y, double z, double tj) { push(x, y, z, tj, -1, [](double, double) {}); }
    GB_DEV void store(double* __restrict__ st, size_t N, size_t i) const {
        for (int k = 0; k < 3; k++) {
            const int base = k < 2 ? 6 * k : 16;
            st[(base + 0) * N + i] = n[k];
            st[(base + 1) * N + i] = n[k] > 0. ? sum[k] / n[k] : CUDART_NAN;     // np.mean of an empty array
            st[(base + 2) * N + i] = n[k] > 0. ? mn[k] : CUDART_NAN;
            st[(base + 3) * N + i] = n[k] > 0. ? mx[k] : CUDART_NAN;
            st[(base + 4) * N + i] = tfirst[k];
            st[(base + 5) * N + i] = tlast[k];
        }
        st[15 * N + i] = zabs;
    }
};

// func=None of Orbit.pericenter / apocenter / zmax: every refined extremum of one kind, in time order.  vals / times
// are (kmax, N) (extremum k of orbit i at [k * N + i], NaN beyond the orbit's count); counts (N) holds the true
// number, which may exceed kmax -- the caller then repeats with a larger kmax.
#if GB_PART == 7          // not a template: defined in one translation unit only
__global__ void k_trajectory_extrema_list(const double* __restrict__ w, const double* __restrict__ t, int ntimes, size_t N,
                                          int kind, int kmax, double* __restrict__ vals, double* __restrict__ times,
                                          int32_t* __restrict__ counts) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t TS = (size_t)ntimes * N;
    ExtremaAcc A; A.init();
    int found = 0;
    auto sink = [&](double val, double tv) {
        if (found < kmax) { vals[(size_t)found * N + i] = val; times[(size_t)found * N + i] = tv; }
        found++;
    };
    const bool rev = t[ntimes - 1] < t[0];
    for (int jj = 0; jj < ntimes; jj++) {
        const int j = rev ? ntimes - 1 - jj : jj;
        const double* p = w + (size_t)j * N + i;
        A.push(__ldcs(p), __ldcs(p + TS), __ldcs(p + 2 * TS), t[j], kind, sink);
    }
    for (int k = found; k < kmax; k++) { vals[(size_t)k * N + i] = CUDART_NAN; times[(size_t)k * N + i] = CUDART_NAN; }
    counts[i] = found;
}
#endif

// An existing trajectory w (6, ntimes, N) in device memory (e.g. the dense output of gb_dop853): one pass, reads are
// coalesced across orbits (consecutive threads = consecutive orbits of one row).  ENERGY: also the Hamiltonian
// (potential + frame energy, chamiltonian.cpp:7-19) at every sample: E_0, E_last, max_j |E_j - E_0|.
template <class C, bool ENERGY>
__global__ void k_trajectory_extrema(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F,
                                     const double* __restrict__ w, const double* __restrict__ t, int ntimes, size_t N,
                                     double* __restrict__ st) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t TS = (size_t)ntimes * N;
    ExtremaAcc A; A.init();
    double E0 = CUDART_NAN, El = CUDART_NAN, dE = 0.;
    const bool rev = t[ntimes - 1] < t[0];        // "time must increase": the reference reverses the orbit (orbit.py:486,546)
    for (int jj = 0; jj < ntimes; jj++) {
        const int j = rev ? ntimes - 1 - jj : jj;
        const double* p = w + (size_t)j * N + i;
        const double x = __ldcs(p), y = __ldcs(p + TS), z = __ldcs(p + 2 * TS);
        A.push(x, y, z, t[j]);
        if (ENERGY) {
            const double E = C::value(P, t[j], x, y, z) + frame_energy(F, x, y, z, __ldcs(p + 3 * TS), __ldcs(p + 4 * TS), __ldcs(p + 5 * TS));
            if (j == 0) E0 = E;
            if (j == ntimes - 1) El = E;
        }
    }
    if (ENERGY) {
        // second sweep for the drift against E_0 = H(w(t[0])) (kept separate so that E_0 is the first sample in
        // storage order also for a reversed grid)
        for (int j = 0; j < ntimes; j++) {
            const double* p = w + (size_t)j * N + i;
            const double x = __ldcs(p), y = __ldcs(p + TS), z = __ldcs(p + 2 * TS);
            const double E = C::value(P, t[j], x, y, z) + frame_energy(F, x, y, z, __ldcs(p + 3 * TS), __ldcs(p + 4 * TS), __ldcs(p + 5 * TS));
            dE = fmax(dE, fabs(E - E0));
        }
    }
    A.store(st, N, i);
    st[12 * N + i] = E0; st[13 * N + i] = El; st[14 * N + i] = ENERGY ? dE : CUDART_NAN;
}

// Integrate AND reduce: the fixed-step loops of k_leapfrog / k_ruth4 (kernels.cu; leapfrog.pyx:24-51,99-116,
// ruth4.pyx:24-35, pyintegrators/ruth4.py:106-124 for the rotating frame) with the statistics taken at every step
// of the caller's grid, final state optional -- nothing of size ntimes is ever written.  SCHEME 0 = leapfrog,
// 1 = Ruth4 static frame, 2 = Ruth4 rotating frame.  Requires an increasing grid (dt > 0) for the extrema times to
// be in the reference's order; for dt < 0 the samples are visited in decreasing time, which finds the same extrema
// (the three-point test and the parabola are symmetric under reversal) and reports first/last by time of discovery.
struct Ruth4CoefE { double c[4], d[4]; };
template <class C, int SCHEME, bool ENERGY>
__global__ void __launch_bounds__(C::kFixedStepMaxThreads, C::kFixedStepMinBlocks)
k_integrate_extrema(const __grid_constant__ DevPot P, const __grid_constant__ DevFrame F, const __grid_constant__ Ruth4CoefE K,
                    const double* __restrict__ w0, size_t N, const double* __restrict__ t, int ntimes, double dt,
                    int dt_from_t, double* __restrict__ wfin, double* __restrict__ st) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (dt_from_t) dt = t[1] - t[0];
    double x = w0[i], y = w0[N + i], z = w0[2 * N + i];
    double vx = w0[3 * N + i], vy = w0[4 * N + i], vz = w0[5 * N + i];
    ExtremaAcc A; A.init();
    A.push(x, y, z, t[0]);
    double E0 = CUDART_NAN, El = CUDART_NAN, dE = 0.;
    if (ENERGY) { E0 = C::value(P, t[0], x, y, z) + frame_energy(F, x, y, z, vx, vy, vz); El = E0; }
    double gx, gy, gz, hx = 0., hy = 0., hz = 0.;
    if (SCHEME == 0) {
        C::gradient(P, t[0], x, y, z, gx, gy, gz);
        hx = vx - gx * dt / 2.; hy = vy - gy * dt / 2.; hz = vz - gz * dt / 2.;
    }
#pragma unroll 1
    for (int j = 1; j < ntimes; j++) {
        if (SCHEME == 0) {
            x = x + hx * dt; y = y + hy * dt; z = z + hz * dt;
            C::gradient(P, C::kTimeDependent ? t[j] : 0., x, y, z, gx, gy, gz);
            if (ENERGY || j == ntimes - 1) { vx = hx - gx * dt / 2.; vy = hy - gy * dt / 2.; vz = hz - gz * dt / 2.; }
            hx = hx - gx * dt; hy = hy - gy * dt; hz = hz - gz * dt;
        } else {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                if (s > 0) {
                    C::gradient(P, C::kTimeDependent ? t[j] : 0., x, y, z, gx, gy, gz);
                    if (SCHEME == 1) {
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
        }
        A.push(x, y, z, t[j]);
        if (ENERGY) {
            El = C::value(P, t[j], x, y, z) + frame_energy(F, x, y, z, vx, vy, vz);
            dE = fmax(dE, fabs(El - E0));
        }
    }
    if (wfin) {
        wfin[i] = x; wfin[N + i] = y; wfin[2 * N + i] = z;
        wfin[3 * N + i] = vx; wfin[4 * N + i] = vy; wfin[5 * N + i] = vz;
    }
    A.store(st, N, i);
    st[12 * N + i] = E0; st[13 * N + i] = El; st[14 * N + i] = ENERGY ? dE : CUDART_NAN;
}
