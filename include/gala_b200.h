/*
 * gala_b200.h -- C ABI of the B200 (sm_100a) orbit-integration engine.
 *
 * This is the drop-in boundary for ONE path of adrn/gala: many independent
 * test-particle orbits advanced through an analytic potential.  Every entry
 * point replaces one reference function (cited as path:line relative to the
 * reference's src/gala/).  Plain pointers and sizes only; no torch types.
 *
 * Layouts are the reference's own:
 *   w0      (6, N)          C-contiguous  (leapfrog.pyx:54-59)
 *   w_out   (6, ntimes, N)  if save_all, else (6, N)   (leapfrog.pyx:92,118-121)
 *   q       (3, N)          for gradient/energy/density (cpotential.pyx:144-162)
 *   stream  (Np, 6)         AoS rows for mock streams  (mockstream.pyx:176-184)
 *
 * Pointer location: every data pointer may be a HOST pointer or a DEVICE
 * pointer; gb_launch.mem says which (GB_MEM_HOST: the library stages through
 * device memory itself, copies are part of the call; GB_MEM_DEVICE: pointers
 * are device memory on the current CUDA device, no copies are made).
 * The potential / frame specs are always host structs; they are copied to the
 * device at call time and the library keeps no reference to them
 * (cf. cpotential.pyx:60-92 where the wrapper owns the arrays).
 *
 * Return value: 0 on success, <0 on failure; gb_last_error() gives the text.
 *   -1..-4   : DOP853 codes as in dopri/dop853.h:157-164 (worst over orbits)
 *   -10      : CUDA runtime error          -11 : unsupported potential/frame
 *   -12      : invalid argument (shape etc; the Python shim raises ValueError)
 *   -13      : integrator does not support this frame (TypeError in the
 *              reference: leapfrog.pyx:64-68, ruth4.pyx:49-52)
 *   -14      : not implemented in the reference either (Hessian of a rotated
 *              potential: NotImplementedError, potential/potential/core.py:572-575)
 * The N-body, snapshot and Lyapunov entry points (gb_nbody_*, gb_*_animate,
 * gb_lyapunov_max) take HOST buffers only.
 *
 * Several GPUs: gb_launch.n_devices / devices shard ONE HOST-buffer call over the listed devices inside the
 * library (gb_gradient / gb_energy / gb_density, gb_hamiltonian_*, gb_leapfrog, gb_ruth4, gb_dop853,
 * gb_integrate_extrema, gb_mockstream_dop853 / _leapfrog); the other entry points run on one device and
 * return -12 when handed a device list.  gb_shard_bounds / gb_deal_count tell the partition.
 */
#ifndef GALA_B200_H
#define GALA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Potential component type ids.  One per *Wrapper class of the reference
 * (potential/potential/builtin/cybuiltin.pyx:99-122); the parameter vector of
 * each is exactly CPotentialWrapper._params = [G, c_only..., params...]
 * (cpotential.pyx:281-316). */
enum gb_pot_type {
    GB_POT_NULL             = 0,  /* NullWrapper            cybuiltin.pyx:362  [G]                       */
    GB_POT_HERNQUIST        = 1,  /* HernquistWrapper       :165  [G, m, c]                              */
    GB_POT_NFW_SPHERICAL    = 2,  /* SphericalNFWWrapper    :292  [G, m, r_s, a, b, c] (a,b,c ignored)   */
    GB_POT_NFW_FLATTENED    = 3,  /* FlattenedNFWWrapper    :303  [G, m, r_s, a, b, c] (c used)          */
    GB_POT_NFW_TRIAXIAL     = 4,  /* TriaxialNFWWrapper     :313  [G, m, r_s, a, b, c]                   */
    GB_POT_MIYAMOTONAGAI    = 5,  /* MiyamotoNagaiWrapper   :264  [G, m, a, b]                           */
    GB_POT_MN3              = 6,  /* MN3ExponentialDiskWrapper :276 [G, m1,a1,b1, m2,a2,b2, m3,a3,b3, ...] */
    GB_POT_LONGMURALIBAR    = 7,  /* LongMuraliBarWrapper   :347  [G, m, a, b, c, alpha]                 */
    GB_POT_SCF              = 8,  /* SCFWrapper scf/bfe_class.pyx:38 [G, nmax, lmax, m, r_s, S.., T..]   */
    GB_POT_KEPLER           = 9,  /* KeplerWrapper          :141  [G, m]                                 */
    GB_POT_PLUMMER          = 10, /* PlummerWrapper         :177  [G, m, b]                              */
    GB_POT_ISOCHRONE        = 11, /* IsochroneWrapper       :153  [G, m, b]                              */
    GB_POT_JAFFE            = 12, /* JaffeWrapper           :189  [G, m, c]                              */
    GB_POT_MULTIPOLE        = 13, /* MultipoleWrapper       :378  [G, lmax, n_coeff, inner, m, r_s, (S_lm,T_lm)..]
                                     (builtin/multipole.cpp:247-262); lmax <= 15                        */
    GB_POT_STONE            = 14, /* StoneWrapper           :201  [G, m, r_c, r_h]                       */
    GB_POT_BURKERT          = 15, /* BurkertWrapper         :226  [G, rho, r0]                           */
    GB_POT_SATOH            = 16, /* SatohWrapper           :240  [G, m, a, b]                           */
    GB_POT_KUZMIN           = 17, /* KuzminWrapper          :252  [G, m, a]                              */
    GB_POT_LOGARITHMIC      = 18, /* LogarithmicWrapper     :324  [G, v_c, r_h, q1, q2, q3, phi]         */
    GB_POT_LEESUTO          = 19, /* LeeSutoTriaxialNFWWrapper :336 [G, v_c, r_s, a, b, c]               */
    GB_POT_POWERLAWCUTOFF   = 20, /* PowerLawCutoffWrapper  :213  [G, m, alpha, r_c] (alpha < 3)         */
    GB_POT_TIMEINTERP       = 21, /* TimeInterpolatedWrapper builtin/cytimeinterp.pyx:73 around one of the analytic types above:
                                     [G, wrapped_type, method, n_knots, n_wpar, n_origin, n_R,
                                      t_knots[n_knots], wpar[n_knots][n_wpar], origin[n_origin][3], R[n_R][9]]
                                     method 0 linear, 1 cspline, 2 akima, 3 steffen (time_interpolated.py:44-60);
                                     wpar rows = the wrapped potential's c_parameters (without G) at every knot, a
                                     constant parameter repeated; n_origin / n_R = 1 (constant) or n_knots.  The
                                     library builds the per-element splines and the axis-angle form of the rotation
                                     (time_interp.cpp:181-405); evaluation outside [t_knots[0], t_knots[-1]] gives NaN
                                     (time_interp_wrapper.cpp:103-106).  Supported by gb_gradient / gb_energy /
                                     gb_density / gb_hessian, the Hamiltonian entries, gb_leapfrog, gb_ruth4, gb_dop853,
                                     gb_integrate_extrema, gb_stream_release and the mock-stream entries without
                                     massive bodies (gb_mockstream_dop853[_animate], gb_mockstream_leapfrog); the
                                     N-body and Lyapunov entry points return -11 for it.                           */
    GB_POT_NTYPES
};

/* One component of a (composite) potential; mirrors struct _CPotential per
 * index i (potential/potential/src/cpotential.h:10-36) with a type id in the
 * place of the four host function pointers. */
typedef struct {
    int32_t type_id;          /* enum gb_pot_type                                   */
    int32_t n_params;         /* length of params (n_params[i])                     */
    int32_t do_shift_rotate;  /* do_shift_rotate[i]: 0 => q0 and R are ignored      */
    int32_t _pad;
    const double* params;     /* parameters[i]: [G, ...]                            */
    double q0[3];             /* q0[i]: origin                                      */
    double R[9];              /* R[i]: row-major 3x3 rotation                       */
} gb_component;

typedef struct {
    int32_t n_components;     /* n_components                                       */
    int32_t n_dim;            /* n_dim; only 3 is supported                         */
    const gb_component* comp;
} gb_potential;

/* Reference frame; mirrors CFrameType (potential/frame/src/cframe.h:7-17). */
enum gb_frame_type {
    GB_FRAME_STATIC      = 0, /* StaticFrameWrapper            frame/builtin/frames.pyx:37-50   */
    GB_FRAME_ROTATING_3D = 1  /* ConstantRotatingFrameWrapper3D frame/builtin/frames.pyx:90-106 */
};
typedef struct {
    int32_t type_id;
    int32_t _pad;
    double omega[3];          /* frame.c_parameters for the rotating frame          */
} gb_frame;

enum gb_mem { GB_MEM_HOST = 0, GB_MEM_DEVICE = 1 };

/* Launch options.  Zero-initialise for defaults. */
typedef struct {
    int32_t mem;              /* enum gb_mem for ALL data pointers of the call      */
    int32_t device;           /* CUDA device ordinal, -1 = current                  */
    void*   stream;           /* cudaStream_t to launch on, NULL = default stream   */
    int32_t strict_math;      /* 1 = strict-IEEE kernels (reference operation order,
                                 no FMA contraction, libm pow/log); 0 = fast kernels */
    int32_t block_threads;    /* 0 = library default                                */
    int32_t n_devices;        /* GB_MEM_HOST calls only: > 0 = shard the call over devices[0..n_devices-1]
                                 (orbit-index sharding; no collective: every device takes a
                                 contiguous slice of the orbit index -- block-cyclic groups of 128 particles for
                                 the mock-stream integrators, whose work per particle is triangular in release
                                 time -- copies its slice straight out of / into the caller's arrays and the call
                                 returns when all devices are done).  0 = the single `device` above.  The result
                                 is bit-identical to the single-device call: orbits never interact.            */
    int32_t _pad;
    const int32_t* devices;   /* CUDA device ordinals, n_devices entries                  */
} gb_launch;

/* Per-orbit DOP853 statistics (dopcor's nstep/naccpt/nrejct/nfcn,
 * dopri/dop853.cpp:115-118); each pointer may be NULL. */
typedef struct {
    int32_t* nstep;
    int32_t* naccpt;
    int32_t* nrejct;
    int32_t* nfcn;
} gb_dop853_stats;

/* ---- potential evaluation (cpotential.cpp:170-287; cpotential.pyx:104-162) ---- */
int gb_gradient(const gb_potential* pot, const double* q, double t, size_t N,
                double* grad /* (3,N) */, const gb_launch* opt);
int gb_energy  (const gb_potential* pot, const double* q, double t, size_t N,
                double* out /* (N) */, const gb_launch* opt);
int gb_density (const gb_potential* pot, const double* q, double t, size_t N,
                double* out /* (N) */, const gb_launch* opt);
/* CPotentialWrapper.hessian -> c_hessian (potential/potential/cpotential.pyx:164-182,
 * potential/potential/src/cpotential.cpp:290-314): hess (3,3,N), H[i][j][n] = d2Phi/dq_i dq_j.  Evaluated by
 * forward-mode differentiation of the gradient on the device (csrc/hessian.cuh).  -14 for a rotated
 * component (NotImplementedError in the reference, core.py:572-575); -11 for SCF / multipole. */
int gb_hessian(const gb_potential* pot, const double* q, double t, size_t N, double* hess, const gb_launch* opt);

/* Hamiltonian.energy: potential + frame energy, w is (6,N)
 * (hamiltonian/src/chamiltonian.cpp:7-19; frame/builtin/builtin_frames.cpp:8-16,73-92) */
int gb_hamiltonian_energy(const gb_potential* pot, const gb_frame* fr, const double* w,
                          double t, size_t N, double* out /* (N) */, const gb_launch* opt);
/* Hamiltonian gradient f = [dH/dp ; -dH/dq] for (6,N)
 * (hamiltonian/src/chamiltonian.cpp:38-57) */
int gb_hamiltonian_gradient(const gb_potential* pot, const gb_frame* fr, const double* w,
                            double t, size_t N, double* f /* (6,N) */, const gb_launch* opt);

/* ---- fixed-step integrators ---------------------------------------------------
 * leapfrog_integrate_hamiltonian (integrate/cyintegrators/leapfrog.pyx:54-121)
 * ruth4_integrate_hamiltonian    (integrate/cyintegrators/ruth4.pyx:37-113)
 * dt = t[1]-t[0]; step j uses time t[j].  save_all=1: w_out is (6,ntimes,N) with
 * w_out[:,0,:]=w0; save_all=0: w_out is (6,N), the state at t[ntimes-1].
 * gb_leapfrog requires a static frame (-13 otherwise).  gb_ruth4 accepts the
 * rotating frame with the semantics of the reference's Python Ruth4Integrator
 * driven by Hamiltonian._gradient (integrate/pyintegrators/ruth4.py:106-124,
 * hamiltonian/chamiltonian.pyx:88-99): p += d_j*(-Omega x p - grad)*dt; q += c_j*p*dt. */
int gb_leapfrog(const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N,
                const double* t, int ntimes, int save_all, double* w_out,
                const gb_launch* opt);
int gb_ruth4   (const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N,
                const double* t, int ntimes, int save_all, double* w_out,
                const gb_launch* opt);

/* ---- adaptive DOP853 -----------------------------------------------------------
 * dop853_integrate_hamiltonian (integrate/cyintegrators/dop853.pyx:196-250) with
 * per-orbit step control (the reference's nbatch=1).  Dense output
 * onto t[] when save_all=1 (dopri/dop853.cpp:584-612,869-904).
 * status: per-orbit dop853 return code (1 ok, -2 nmax, -3 step too small, -4 stiff),
 * may be NULL.  The function returns the most negative per-orbit code, or 0. */
int gb_dop853(const gb_potential* pot, const gb_frame* fr, const double* w0, size_t N,
              const double* t, int ntimes, double atol, double rtol, long nmax,
              double dt_max, long nstiff, int save_all, double* w_out,
              int32_t* status, const gb_dop853_stats* stats, const gb_launch* opt);

/* ---- mock streams ----------------------------------------------------------------
 * Particle release: FardalStreamDF._sample + BaseStreamDF.get_rj_vj_R +
 * transform_from_sat (dynamics/mockstream/df.pyx:61-106,363-456).  The normal
 * deviates are drawn on the host (numpy RNG, reference order) and passed in:
 * normals is (Np,4) rows [kx, z, vt, vz] already scaled to N(mean,disp).
 * prog_idx[p] = progenitor timestep of particle p; sign[p] = +1 trailing, -1 leading. */
int gb_fardal_release(const gb_potential* pot, double G,
                      const double* prog_w /* (ntimes,6) */, const double* prog_t,
                      const double* prog_m, int ntimes,
                      const int32_t* prog_idx, const double* sign, const double* normals,
                      size_t Np, int gala_modified,
                      double* stream_w0 /* (Np,6) */, const gb_launch* opt);

/* The same release step for every builtin stream DF (dynamics/mockstream/df.pyx): df_kind 0 = Fardal
 * (draws (Np,4), flags = gala_modified), 1 = Streakline (:242-318, no draws), 2 = LagrangeCloud
 * (:460-552, draws (Np,3) = N(0, v_disp) velocity offsets), 3 = Chen+24 (:556-702, draws (Np,6) =
 * multivariate_normal(mean, cov) rows [r, phi, theta, v, alpha, beta], angles in degrees). */
int gb_stream_release(const gb_potential* pot, double G,
                      const double* prog_w /* (ntimes,6) */, const double* prog_t,
                      const double* prog_m, int ntimes,
                      const int32_t* prog_idx, const double* sign, const double* draws, int ncols,
                      size_t Np, int df_kind, int flags,
                      double* stream_w0 /* (Np,6) */, const gb_launch* opt);

/* mockstream_dop853 (dynamics/mockstream/mockstream.pyx:176-303), no massive
 * bodies: every stream particle p is integrated from t1[p] to tfinal as its own
 * n=6 DOP853 system with dop853_step's settings (dop853.pyx:27-75: uround default,
 * initial step dt0, stiffness test after every accepted step). */
int gb_mockstream_dop853(const gb_potential* pot, const gb_frame* fr,
                         const double* stream_w0 /* (Np,6) */, const double* t1 /* (Np) */,
                         size_t Np, double tfinal, double dt0,
                         double atol, double rtol, long nmax,
                         double* stream_w /* (Np,6) */, int32_t* status,
                         const gb_launch* opt);
/* mockstream_leapfrog (mockstream.pyx:442-620), no massive bodies: particle p takes
 * n_steps = (int)((tfinal-t1[p])/dt+0.5) leapfrog steps of size dt. */
int gb_mockstream_leapfrog(const gb_potential* pot,
                           const double* stream_w0, const double* t1, size_t Np,
                           double tfinal, double dt,
                           double* stream_w, const gb_launch* opt);

/* mockstream_dop853_animate (dynamics/mockstream/mockstream.pyx:306-440), no massive bodies: the released
 * particles are marched over the time grid t interval by interval (one dop853_step call per interval, initial
 * step t[1]-t[0] each time) and stored every `output_every` intervals and at the end.  release_idx[p] = index
 * in t of particle p's release time.  snapshots: (nout, Np, 6) rows with
 * nout = (ntimes-1)/output_every + 1 (+1 when the last interval is not a multiple, :360-363); NaN before release. */
int gb_mockstream_dop853_animate(const gb_potential* pot, const gb_frame* fr,
                                 const double* w0_rows /* (Np,6) */, const int32_t* release_idx, size_t Np,
                                 const double* t, int ntimes, double atol, double rtol, long nmax,
                                 int output_every, double* snapshots, double* final_w /* (Np,6) */,
                                 int32_t* status, const gb_launch* opt);

/* ---- massive bodies (direct N-body) --------------------------------------------------
 * Replaces, for systems of a few massive bodies plus any number of TEST particles:
 *   leapfrog_integrate_nbody   integrate/cyintegrators/leapfrog.pyx:161-257
 *   direct_nbody_dop853        dynamics/nbody/nbody.pyx:30-115
 *   the massive-body branches of mockstream_leapfrog / mockstream_dop853
 *                              dynamics/mockstream/mockstream.pyx:442-620, 176-303
 * (c_nbody_gradient_symplectic / c_nbody_acceleration, potential/potential/src/cpotential.cpp:389-442).
 * body_pot[b] is body b's own potential (the reference's particle_potentials[b]; n_components == 0 or
 * all-Null components = a massless body), evaluated about the body's current position.  Every device
 * lane integrates [the bodies, ONE test particle]; test particles never act on anything.  Host buffers
 * only.  Rows are AoS: body_w0 (ngroups, n_bodies, 6) -- particle p starts with the bodies in state
 * body_w0[group[p]] (group == NULL: state 0) at time t1[p] (t1 == NULL: the common start time);
 * w0_rows (Np, 6); out_particles (Np, 6); out_bodies (n_bodies, 6) = the bodies as integrated by lane
 * `body_writer` (Np == 0: the single bodies-only lane); traj (ntimes, n_bodies + Np, 6) or NULL. */
#define GB_MAX_BODIES 16
typedef struct {
    int32_t n_bodies;              /* 1..GB_MAX_BODIES */
    int32_t _pad;
    const gb_potential* body_pot;  /* [n_bodies] */
} gb_bodies;

/* fixed step: particle p takes int((tfinal - t1[p])/dt + 0.5) steps (mockstream.pyx:571), or `nsteps`
 * when t1 == NULL.  scheme 0 = leapfrog; 1 = Ruth4 (ruth4_integrate_nbody,
 * integrate/cyintegrators/ruth4.pyx:116-241).  StaticFrame only, like the reference (leapfrog.pyx:172-176). */
int gb_nbody_leapfrog(const gb_potential* pot, const gb_bodies* bodies,
                      const double* body_w0, int ngroups, const int32_t* group,
                      const double* w0_rows, const double* t1, size_t Np,
                      double t0, double tfinal, int nsteps, double dt, int scheme,
                      double* out_particles, double* out_bodies, size_t body_writer,
                      double* traj, const gb_launch* opt);

/* DOP853: step_mode 0 = dop853_helper's settings with nstiff = -1 (direct_nbody_dop853), tgrid/ntimes
 * = the caller's output grid (dense output when traj != NULL); step_mode 1 = dop853_step's settings
 * (mockstream_dop853).  Systems of 1-2 points keep their state in registers; larger ones (up to
 * GB_MAX_BODIES + 1 points) run with their state in local memory.  status:
 * one dop853 code per lane (Np entries, or 1). */
int gb_nbody_dop853(const gb_potential* pot, const gb_bodies* bodies,
                    const double* body_w0, int ngroups, const int32_t* group,
                    const double* w0_rows, const double* t1, size_t Np,
                    const double* tgrid, int ntimes, double tfinal, double dt0,
                    double atol, double rtol, long nmax, double dt_max, int step_mode,
                    double* out_particles, double* out_bodies, size_t body_writer,
                    double* traj, int32_t* status, const gb_launch* opt);

/* mockstream_dop853_animate with massive bodies: the bodies are marched alone over t (their states at the
 * snapshot times fill snapshots[:, :n_bodies]); particle p then starts at release_idx[p] from the bodies' state
 * there and is marched as [bodies, particle].  body_w0 (n_bodies, 6) at t[0]; snapshots (nout, n_bodies + Np, 6);
 * out_bodies = the bodies' end state of the bodies-only march. */
int gb_nbody_dop853_animate(const gb_potential* pot, const gb_bodies* bodies, const double* body_w0,
                            const double* w0_rows, const int32_t* release_idx, size_t Np,
                            const double* t, int ntimes, double atol, double rtol, long nmax, int output_every,
                            double* snapshots, double* out_particles, double* out_bodies, int32_t* status,
                            const gb_launch* opt);

/* ---- chaos indicators ----------------------------------------------------------------
 * dop853_lyapunov_max(_dont_save) (dynamics/lyapunov/dop853_lyapunov.pyx:22-192) for N parent orbits at once
 * (the reference takes one per call): lane p integrates parent orbit p and its offset orbits
 * w0[p] + d0_vec[p][i] as one DOP853 system, interval by interval over t (dop853_step), and every
 * n_steps_per_pullback intervals stores ln(|d1|/d0) per offset orbit in LEs_raw (N, n_steps / pullback, noff)
 * and pulls the offset orbit back to distance d0.  traj: (N, n_steps, 1 + noff, 6) or NULL.  Host buffers. */
int gb_lyapunov_max(const gb_potential* pot, const gb_frame* fr,
                    const double* w0_rows /* (N,6) */, const double* d0_vec /* (N,noff,6) */, size_t N,
                    const double* t, int n_steps, double d0, int n_steps_per_pullback, int noffset_orbits,
                    double atol, double rtol, long nmax,
                    double* LEs_raw, double* traj, int32_t* status, const gb_launch* opt);

/* ---- trajectory reductions on the device -----------------------------------
 * Orbit.pericenter / Orbit.apocenter (dynamics/orbit.py:391-553) and the energy drift without moving a
 * (6, ntimes, N) trajectory over PCIe.  Per orbit, on the samples r_j = |q(t_j)| of the caller's grid: interior
 * samples strictly above (below) both neighbours are apocentres (pericentres) -- scipy.signal.argrelmax as called at
 * orbit.py:402-403 -- each refined to the vertex of the parabola through (t, r) at j-1, j, j+1 (orbit.py:413-422).
 * stats is (GB_EXT_NSTAT, N), rows as enumerated below; means / minima / maxima are NaN for an orbit without such an
 * extremum (np.mean of an empty array).  with_energy != 0 also evaluates the Hamiltonian (potential + frame energy)
 * at every sample: E(t[0]), E(t[-1]) and max_j |E_j - E(t[0])|. */
enum gb_extrema_row {
    GB_EXT_N_PERI = 0, GB_EXT_PERI_MEAN, GB_EXT_PERI_MIN, GB_EXT_PERI_MAX, GB_EXT_PERI_T_FIRST, GB_EXT_PERI_T_LAST,
    GB_EXT_N_APO = 6,  GB_EXT_APO_MEAN,  GB_EXT_APO_MIN,  GB_EXT_APO_MAX,  GB_EXT_APO_T_FIRST,  GB_EXT_APO_T_LAST,
    GB_EXT_E_FIRST = 12, GB_EXT_E_LAST, GB_EXT_DE_MAX, GB_EXT_ABS_Z_MAX,
    /* Orbit.zmax (dynamics/orbit.py:600-656): local maxima of |z|, refined the same way */
    GB_EXT_N_ZMAX = 16, GB_EXT_ZMAX_MEAN, GB_EXT_ZMAX_MIN, GB_EXT_ZMAX_MAX, GB_EXT_ZMAX_T_FIRST, GB_EXT_ZMAX_T_LAST,
    GB_EXT_NSTAT = 22
};
/* reduce a trajectory that already exists: w is (6, ntimes, N) (e.g. the dense output of gb_dop853 left in device
 * memory); a decreasing grid is walked backwards, like the reference reverses the orbit (orbit.py:486,546). */
int gb_orbit_extrema(const gb_potential* pot, const gb_frame* fr, const double* w, const double* t, int ntimes,
                     size_t N, int with_energy, double* stats /* (GB_EXT_NSTAT, N) */, const gb_launch* opt);
/* func=None of Orbit.pericenter / apocenter / zmax: every refined extremum of one kind (0 pericentres, 1 apocentres,
 * 2 z-heights) of every orbit, in increasing time.  vals / times are (kmax, N), NaN beyond an orbit's count;
 * counts (N) holds the true number per orbit (repeat with a larger kmax if any exceeds it). */
int gb_orbit_extrema_list(const double* w, const double* t, int ntimes, size_t N, int kind, int kmax,
                          double* vals, double* times, int32_t* counts, const gb_launch* opt);
/* integrate like gb_leapfrog (scheme 0) / gb_ruth4 (scheme 1; rotating frame with gb_ruth4's semantics) and reduce on
 * the fly: nothing of size ntimes is written.  w_final (6, N) may be NULL.  Shards over gb_launch.devices. */
int gb_integrate_extrema(const gb_potential* pot, const gb_frame* fr, int scheme, const double* w0, size_t N,
                         const double* t, int ntimes, int with_energy, double* w_final,
                         double* stats /* (GB_EXT_NSTAT, N) */, const gb_launch* opt);

/* ---- misc ------------------------------------------------------------------------ */
const char* gb_last_error(void);
int  gb_device_count(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
long gb_launch_count(void);
const char* gb_version(void);
/* FP64 DFMA throughput (TFLOP/s, 2 flops per FMA) measured on the current device: the roofline
 * denominator bench.py reports against (MEASURED_PEAKS.json carries no FP64 figure). */
double gb_fp64_peak_tflops(int reps);

/* The partition a multi-device call uses (no device needed): orbit slice [*lo, *lo + *n) of device k of nd for N
 * orbits (contiguous, sizes differ by at most one), and the number of mock-stream particles of device k when Np rows
 * are dealt in groups of 128 consecutive rows, group g to device g mod nd.  Return -12 for k outside [0, nd). */
int gb_shard_bounds(size_t N, int k, int nd, size_t* lo, size_t* n);
long gb_deal_count(size_t Np, int k, int nd);

/* Frees the device staging / scratch buffers the library keeps cached between calls (HOST-mode staging
 * of inputs and outputs; the orbit-major dense-output scratch of gb_dop853, up to half of the free
 * device memory; the device copies of SCF / multipole coefficient blocks and of the PowerLawCutoff
 * table, at most 16 small buffers).  Safe to call at any time from the thread that made the calls. */
int gb_release_scratch(void);

/* Diagnostic: y[i] = f(x[i]) with the math primitive the selected build (opt->strict_math) uses
 * inside the kernels: which = 0: 1/x, 1: x^-1/2, 2: x^-3/2, 3: ln x.  The fast build replaces the
 * CUDA library expansions by seed + fixed refinement (csrc/fastmath.cuh); this entry lets the test
 * suite measure their error in ulp on the device. */
int gb_math_probe(int which, const double* x, size_t N, double* y, const gb_launch* opt);

#ifdef __cplusplus
}
#endif
#endif /* GALA_B200_H */
