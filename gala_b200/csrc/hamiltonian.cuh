// hamiltonian.cuh -- frame terms and the Hamiltonian right-hand side, fused with the potential
// gradient (reference potential/hamiltonian/src/chamiltonian.cpp:7-57 and
// potential/frame/builtin/builtin_frames.cpp:8-24,73-112).
#pragma once

// frame energy: static 0.5*p^2 (builtin_frames.cpp:8-16); rotating 0.5 p^2 - Omega.L (:73-92)
GB_DEV double frame_energy(const DevFrame& F, double x, double y, double z, double px, double py, double pz) {
    if (F.type == GB_FRAME_STATIC) {
        double E = 0.;
        E += px * px; E += py * py; E += pz * pz;
        return 0.5 * E;
    }
    double E = 0.;
    E += 0.5 * px * px; E += 0.5 * py * py; E += 0.5 * pz * pz;
    const double Lx = y * pz - z * py;
    const double Ly = -x * pz + z * px;
    const double Lz = x * py - y * px;
    return E - (F.om[0] * Lx + F.om[1] * Ly + F.om[2] * Lz);
}

// f = [dH/dp ; -dH/dq]  (hamiltonian_gradient_T, chamiltonian.cpp:38-57).  Static frame: qdot = p
// (builtin_frames.cpp:18-24).  Rotating frame: qdot = p - Omega x q, pdot = -(grad + Omega x p)
// (builtin_frames.cpp:94-112).
// GB_RHS_NOINLINE (DOP853 translation units): one out-of-line copy of the potential gradient instead
// of 18 inlined ones keeps the adaptive kernel's code within the instruction cache.
template <class C>
__device__ __noinline__ void gradient_call(const DevPot& P, double t, double x, double y, double z,
                                           double& gx, double& gy, double& gz) {
    C::gradient(P, t, x, y, z, gx, gy, gz);
}

template <class C, bool ROT>
GB_DEV void ham_rhs(const DevPot& P, const DevFrame& F, double t, const double (&w)[6], double (&f)[6]) {
    double gx, gy, gz;
#ifdef GB_RHS_NOINLINE
    gradient_call<C>(P, t, w[0], w[1], w[2], gx, gy, gz);
#else
    if constexpr (C::kOutOfLineInRhs) gradient_call<C>(P, t, w[0], w[1], w[2], gx, gy, gz);
    else C::gradient(P, t, w[0], w[1], w[2], gx, gy, gz);
#endif
    if (!ROT) {
        f[0] = w[3]; f[1] = w[4]; f[2] = w[5];
        f[3] = -gx; f[4] = -gy; f[5] = -gz;
    } else {
        double Cx = F.om[1] * w[2] - F.om[2] * w[1];
        double Cy = -F.om[0] * w[2] + F.om[2] * w[0];
        double Cz = F.om[0] * w[1] - F.om[1] * w[0];
        f[0] = w[3] - Cx; f[1] = w[4] - Cy; f[2] = w[5] - Cz;
        Cx = F.om[1] * w[5] - F.om[2] * w[4];
        Cy = -F.om[0] * w[5] + F.om[2] * w[3];
        Cz = F.om[0] * w[4] - F.om[1] * w[3];
        f[3] = -(gx + Cx); f[4] = -(gy + Cy); f[5] = -(gz + Cz);
    }
}
