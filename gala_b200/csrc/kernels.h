// kernels.h -- host-callable launchers of the sm_100a kernels.  kernels.cu is compiled twice:
//   -DGB_STRICT=0 (namespace gbk_fast):   default nvcc FP contraction, rsqrt-based x^-1.5
//   -DGB_STRICT=1 -fmad=false (gbk_strict): reference operation order, IEEE div/sqrt, libm pow
// capi.cu picks one per call from gb_launch.strict_math.
#pragma once
#include <cuda_runtime.h>
#include "gb_device.cuh"

struct Dop853Args {
    double atol, rtol;
    long nmax;        // already defaulted (0 -> 1e6)
    long nstiff;      // already defaulted (0 -> 1000, <0 -> nmax+10)
    double hmax;      // 0 -> |xend-x| per orbit
    double uround;
    double h0;        // initial step (t[1]-t[0]); 0 -> hinit
};

#define GB_DECL_ND(NAME)                                                                                   \
    cudaError_t NAME(const DevPot& P, const DevBodies& B, const Dop853Args& a, const double* body_w0,     \
                     const int32_t* group, const double* w0, const double* t1, size_t Np,                  \
                     const double* tgrid, int ntimes, double t0, double tfinal, double* out_p,             \
                     double* out_b, size_t body_writer, double* traj, size_t ntot, int32_t* status,        \
                     cudaStream_t s);
#define GB_DECLARE_KERNEL_API(NS)                                                                          \
    namespace NS {                                                                                         \
    cudaError_t eval_gradient(const DevPot& P, const double* q, double t, size_t N, double* g,            \
                              int block, cudaStream_t s);                                                  \
    cudaError_t math_probe(int which, const double* x, size_t N, double* y, cudaStream_t s);              \
    cudaError_t eval_energy(const DevPot& P, const double* q, double t, size_t N, double* out,            \
                            int block, cudaStream_t s);                                                    \
    cudaError_t eval_density(const DevPot& P, const double* q, double t, size_t N, double* out,           \
                             int block, cudaStream_t s);                                                   \
    cudaError_t eval_hessian(const DevPot& P, const double* q, double t, size_t N, double* hess,         \
                             int block, cudaStream_t s);                                                   \
    cudaError_t ham_energy(const DevPot& P, const DevFrame& F, const double* w, double t, size_t N,       \
                           double* out, int block, cudaStream_t s);                                        \
    cudaError_t ham_gradient(const DevPot& P, const DevFrame& F, const double* w, double t, size_t N,     \
                             double* f, int block, cudaStream_t s);                                        \
    /* time-dependent composites: the TimeInterpolated components' state at every grid time (ti_tab) */    \
    size_t ti_table_bytes(const DevPot& P, int ntimes);                                                    \
    cudaError_t ti_table(const DevPot& P, const double* t, int ntimes, void* tab, cudaStream_t s);        \
    cudaError_t leapfrog(const DevPot& P, const double* w0, size_t N, const double* t, int ntimes,        \
                         double dt, int dt_from_t, int save_all, double* out, const void* ti_tab,          \
                         int block, cudaStream_t s);                                                       \
    cudaError_t ruth4(const DevPot& P, const DevFrame& F, const double* w0, size_t N, const double* t,    \
                      int ntimes, double dt, int dt_from_t, const double* cs, const double* ds,            \
                      int save_all, double* out, const void* ti_tab, int block, cudaStream_t s);           \
    cudaError_t dop853_static(const DevPot& P, const DevFrame& F, const double* w0, size_t N,             \
                              const double* t, int ntimes, const Dop853Args& a, int save_all,              \
                              const uint32_t* perm, unsigned long long* queue, size_t orb0, size_t nslots, \
                              double* out, int32_t* status, int32_t* nstep, int32_t* naccpt,               \
                              int32_t* nrejct, int32_t* nfcn, int block, cudaStream_t s);                  \
    cudaError_t dop853_rotating(const DevPot& P, const DevFrame& F, const double* w0, size_t N,           \
                                const double* t, int ntimes, const Dop853Args& a, int save_all,            \
                                const uint32_t* perm, unsigned long long* queue, size_t orb0,              \
                                size_t nslots, double* out, int32_t* status, int32_t* nstep,               \
                                int32_t* naccpt, int32_t* nrejct, int32_t* nfcn, int block,                \
                                cudaStream_t s);                                                           \
    cudaError_t dop853_transpose(const double* scratch, size_t orb0, size_t nslots, int ntimes, size_t N, \
                                 double* out, cudaStream_t s);                                             \
    cudaError_t dyn_time_keys(const DevPot& P, const double* w0, size_t N, double t0, size_t orb0,        \
                              size_t n, float* key, uint32_t* idx, cudaStream_t s);                        \
    cudaError_t mock_dop853(const DevPot& P, const DevFrame& F, const double* w0_rows, const double* t1,  \
                            size_t Np, double tfinal, const Dop853Args& a, double* out_rows,               \
                            int32_t* status, int block, cudaStream_t s);                                   \
    cudaError_t mock_dop853_animate(const DevPot& P, const DevFrame& F, const double* w0_rows,            \
                                    const int32_t* ridx, size_t Np, const double* t, int ntimes,           \
                                    const Dop853Args& a, int output_every, double* snap, double* out_rows, \
                                    int32_t* status, int block, cudaStream_t s);                           \
    cudaError_t mock_leapfrog(const DevPot& P, const double* w0_rows, const double* t1, size_t Np,        \
                              double tfinal, double dt, double* out_rows, int block, cudaStream_t s);      \
    cudaError_t nbody_leapfrog(const DevPot& P, const DevBodies& B, int scheme, const double* cs,          \
                               const double* ds, const double* body_w0, const int32_t* group,              \
                               const double* w0, const double* t1, size_t Np, double t0, double tfinal,   \
                               int nsteps_fixed, double dt, double* out_p, double* out_b,                  \
                               size_t body_writer, double* traj, size_t ntot, int block, cudaStream_t s);  \
    GB_DECL_ND(nbody_dop853_small) GB_DECL_ND(nbody_dop853_big)                                            \
    cudaError_t nbody_dop853(const DevPot& P, const DevBodies& B, const Dop853Args& a,                    \
                             const double* body_w0, const int32_t* group, const double* w0,                \
                             const double* t1, size_t Np, const double* tgrid, int ntimes, double t0,      \
                             double tfinal, double* out_p, double* out_b, size_t body_writer,              \
                             double* traj, size_t ntot, int32_t* status, cudaStream_t s);                  \
    cudaError_t nbody_dop853_march(const DevPot& P, const DevBodies& B, const Dop853Args& a,              \
                                   double* body_all, const double* w0, const int32_t* ridx, size_t Np,     \
                                   int has_particle, const double* t, int ntimes, int output_every,        \
                                   double* snap, double* out_p, double* out_b, size_t body_writer,         \
                                   int32_t* status, cudaStream_t s);                                       \
    cudaError_t lyapunov(const DevPot& P, const DevFrame& F, const Dop853Args& a, const double* w0,       \
                         const double* d0_vec, size_t N, const double* t, int n_steps, double d0,          \
                         int pullback, int noff, double* LEs, double* traj, int32_t* status,               \
                         cudaStream_t s);                                                                  \
    cudaError_t trajectory_extrema_e0(const DevPot& P, const DevFrame& F, const double* w, const double* t, \
                                      int ntimes, size_t N, double* stats, int block, cudaStream_t s);      \
    cudaError_t trajectory_extrema_e1(const DevPot& P, const DevFrame& F, const double* w, const double* t, \
                                      int ntimes, size_t N, double* stats, int block, cudaStream_t s);      \
    cudaError_t trajectory_extrema_list(const double* w, const double* t, int ntimes, size_t N, int kind,  \
                                        int kmax, double* vals, double* times, int32_t* counts, int block,  \
                                        cudaStream_t s);                                                    \
    cudaError_t integrate_extrema_e0(const DevPot& P, const DevFrame& F, int scheme, const double* cs,     \
                                     const double* ds, const double* w0, size_t N, const double* t,        \
                                     int ntimes, double dt, int dt_from_t, double* wfin, double* stats,     \
                                     int block, cudaStream_t s);                                            \
    cudaError_t integrate_extrema_e1(const DevPot& P, const DevFrame& F, int scheme, const double* cs,     \
                                     const double* ds, const double* w0, size_t N, const double* t,        \
                                     int ntimes, double dt, int dt_from_t, double* wfin, double* stats,     \
                                     int block, cudaStream_t s);                                            \
    cudaError_t fardal_release(const DevPot& P, double G, const double* prog_w, const double* prog_t,     \
                               const double* prog_m, int ntimes, const int32_t* prog_idx,                  \
                               const double* sign, const double* normals, int ncols, size_t Np, int kind, \
                               int gala_modified, double* out_rows, int block, cudaStream_t s);            \
    }

// sort.cu: radix sort of (float key, uint32 value) pairs with cub (plumbing for the orbit queue order)
cudaError_t gb_sort_pairs_bytes(size_t n, size_t* temp_bytes);
cudaError_t gb_sort_pairs(const float* keys_in, float* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                          size_t n, void* temp, size_t temp_bytes, cudaStream_t s);

GB_DECLARE_KERNEL_API(gbk_fast)
GB_DECLARE_KERNEL_API(gbk_strict)
