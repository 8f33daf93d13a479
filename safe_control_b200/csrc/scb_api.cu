// scb_api.cu -- the C ABI of libscb.so (include/scb.h): argument checks, launch dispatch,
// and the host-pointer staging context.  No torch, no exceptions, no global mutable state
// (only a thread-local copy of the last CUDA error code).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "scb_kernels.cuh"
#if defined(SCB_SINGLE_TU) || defined(SCB_MPC_PROFILE)
// one translation unit (debug / profiling builds): the MPC kernels are instantiated here instead of in scb_mpc_inst.cu
#include "scb_mpc_impl.cuh"
namespace scb {
SCB_MPC_INSTANTIATE(SCB_SINGLE_INTEGRATOR_2D)
SCB_MPC_INSTANTIATE(SCB_DYNAMIC_UNICYCLE_2D)
SCB_MPC_INSTANTIATE(SCB_KINEMATIC_BICYCLE_2D)
SCB_MPC_INSTANTIATE(SCB_QUAD_3D)
SCB_MPC_INSTANTIATE(SCB_DOUBLE_INTEGRATOR_2D)
SCB_MPC_INSTANTIATE(SCB_QUAD_2D)
SCB_MPC_INSTANTIATE(SCB_UNICYCLE_2D)
SCB_MPC_INSTANTIATE(SCB_KINEMATIC_BICYCLE_2D_C3BF)
SCB_MPC_INSTANTIATE(SCB_KINEMATIC_BICYCLE_2D_DPCBF)
SCB_MPC_INSTANTIATE(SCB_VTOL_2D)
SCB_MPC_INSTANTIATE(kMpcOdBase + SCB_DYNAMIC_UNICYCLE_2D)
SCB_MPC_INSTANTIATE(kMpcOdBase + SCB_KINEMATIC_BICYCLE_2D)
SCB_MPC_INSTANTIATE(kMpcOdBase + SCB_QUAD_2D)
SCB_MPC_INSTANTIATE(kMpcOdBase + SCB_VTOL_2D)
SCB_MPC_INSTANTIATE(kMpcSeBase + SCB_SINGLE_INTEGRATOR_2D)
SCB_MPC_INSTANTIATE(kMpcSeBase + SCB_DYNAMIC_UNICYCLE_2D)
SCB_MPC_INSTANTIATE(kMpcSeBase + SCB_DOUBLE_INTEGRATOR_2D)
}
#else
#include "scb_mpc_kernels.cuh"
#endif
#include "scb_track_kernels.cuh"
#include "scb_backup_kernels.cuh"

using namespace scb;

static thread_local int g_last_cuda = 0;

static int cuda_fail(cudaError_t e) {
  g_last_cuda = (int)e;
  return SCB_ERR_CUDA;
}
#define CK(x)                                  \
  do {                                         \
    cudaError_t e__ = (x);                     \
    if (e__ != cudaSuccess) return cuda_fail(e__); \
  } while (0)

// SM count of the CURRENT device (cached per device index; the value is immutable hardware data)
static int sm_count_of_current() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  if (dev >= 0 && dev < 64) cache[dev] = v;       // (benign race: every writer stores the same value)
  return v;
}

// ------------------------------------------------------------------------------ dispatch
// can the obstacle blocks be fetched with 1-D bulk copies?  (16-byte aligned source addresses and sizes)
static bool tma_ok(const double* OBS, long stride, int M) {
  const char* e = getenv("SCB_QP_TMA");
  if (e && e[0] == '0') return false;
  return OBS && stride > 0 && (stride & 1) == 0 && M > 0 && (M & 1) == 0 && (((uintptr_t)OBS) & 15) == 0;
}

template <int MODEL>
static int launch_cbfqp_m(const scb_params& p, const LaunchGeom& g, int N, int M, const double* X, const double* Uref,
                          const double* OBS, long stride, const int32_t* nobs, double* U, int32_t* status,
                          uint64_t* active, int words, cudaStream_t s, bool eager, const int32_t* skip) {
  // warp per agent (small batches), inputs in device memory: obstacle rows are loaded before nobs is known
  if (eager && g.lanes == 32 && g.rpl == 1) {
    cbfqp_kernel<MODEL, 32, 1, true><<<g.grid, kBlock, 0, s>>>(p, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, skip);
    return SCB_OK;
  }
  if (eager && g.lanes == 32 && g.rpl == 2) {
    cbfqp_kernel<MODEL, 32, 2, true><<<g.grid, kBlock, 0, s>>>(p, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, skip);
    return SCB_OK;
  }
  // bulk-async staged variant for the lane-group geometries (large batches): per-agent lists, 16-byte aligned blocks.
  // SCB_QP_TMA=0 forces the LDG kernel (A/B switch, read per call).
  if (g.lanes < 32 && tma_ok(OBS, stride, M)) {
    const int slot = tma_slot_doubles(M, g.lanes);
    const size_t smem = (size_t)(kBlock / 32) * (32 / g.lanes) * slot * sizeof(double);
#define GOT(L, R)                                                                                                   \
    if (g.lanes == L && g.rpl == R) {                                                                               \
      auto kern = cbfqp_tma_kernel<MODEL, L, R>;                                                                    \
      if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
        return SCB_ERR_TOO_LARGE;                                                                                   \
      kern<<<g.grid, kBlock, smem, s>>>(p, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, skip, slot); \
      return SCB_OK;                                                                                                \
    }
    GOT(8, 3) GOT(8, 4) GOT(8, 8) GOT(4, 5) GOT(4, 8)
#undef GOT
  }
#define GO(L, R)                                                                                           \
  if (g.lanes == L && g.rpl == R) {                                                                        \
    cbfqp_kernel<MODEL, L, R><<<g.grid, kBlock, 0, s>>>(p, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, skip); \
    return SCB_OK;                                                                                         \
  }
  GO(32, 1) GO(32, 2) GO(32, 4) GO(8, 3) GO(8, 4) GO(8, 8) GO(4, 5) GO(4, 8)
#undef GO
  return SCB_ERR_TOO_LARGE;
}

template <int MODEL, int NW>
static int launch_od_m(const scb_params& p, const LaunchGeom& g, int N, int M, const double* X, const double* Uref,
                       const double* OBS, long stride, const int32_t* nobs, double* U, double* omega, int32_t* sel,
                       int32_t* status, uint64_t* active, cudaStream_t s, const int32_t* skip) {
  if (g.lanes < 32 && tma_ok(OBS, stride, M)) {      // bulk-async staged variant (see odcbf_tma_kernel)
    const int slot = tma_slot_doubles(M, g.lanes);
    const size_t smem = (size_t)(kBlock / 32) * (32 / g.lanes) * slot * sizeof(double);
#define GOT(L, R)                                                                                                   \
    if (g.lanes == L && g.rpl == R) {                                                                               \
      auto kern = odcbf_tma_kernel<MODEL, NW, L, R>;                                                                \
      if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
        return SCB_ERR_TOO_LARGE;                                                                                   \
      kern<<<g.grid, kBlock, smem, s>>>(p, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active, skip, slot); \
      return SCB_OK;                                                                                                \
    }
    GOT(8, 3) GOT(8, 4) GOT(8, 8) GOT(4, 5) GOT(4, 8)
#undef GOT
  }
#define GO(L, R)                                                                                              \
  if (g.lanes == L && g.rpl == R) {                                                                           \
    odcbf_kernel<MODEL, NW, L, R><<<g.grid, kBlock, 0, s>>>(p, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active, skip); \
    return SCB_OK;                                                                                            \
  }
  GO(32, 1) GO(32, 2) GO(32, 4) GO(8, 3) GO(8, 4) GO(8, 8) GO(4, 5) GO(4, 8)
#undef GO
  return SCB_ERR_TOO_LARGE;
}

static int forced_lanes() {      // tuning override, read per call (no cached state)
  const char* e = getenv("SCB_QP_LANES");
  return e ? atoi(e) : 0;
}

static bool qp_model_ok(int m) {
  return m == SCB_SINGLE_INTEGRATOR_2D || m == SCB_DYNAMIC_UNICYCLE_2D || m == SCB_KINEMATIC_BICYCLE_2D ||
         m == SCB_KINEMATIC_BICYCLE_2D_C3BF || m == SCB_DOUBLE_INTEGRATOR_2D || m == SCB_QUAD_2D ||
         m == SCB_KINEMATIC_BICYCLE_2D_DPCBF || m == SCB_UNICYCLE_2D || m == SCB_MANIPULATOR_2D;
}

extern "C" {

int scb_last_cuda_error(void) { return g_last_cuda; }

int scb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int scb_limits(int* max_obs_qp, int* max_obs_mpc, int* max_horizon) {
  if (max_obs_qp) *max_obs_qp = 124;       // rows = M + 4 <= 128
  if (max_obs_mpc) *max_obs_mpc = kMpcMaxObs;
  if (max_horizon) *max_horizon = kMpcMaxH;
  return SCB_OK;
}

// ------------------------------------------------------------------------------ CBF-QP
int scb_cbfqp_rows(const scb_params* p, int N, int M, const double* X, const double* OBS, long stride,
                   const int32_t* nobs, double* A, double* b, void* stream) {
  if (!p || N < 0 || M < 0 || (N > 0 && M > 0 && (!X || !OBS || !A || !b))) return SCB_ERR_BAD_ARG;
  if (!qp_model_ok(p->model)) return p->model == SCB_QUAD_3D ? SCB_ERR_UNSUPPORTED : SCB_ERR_BAD_ARG;
  if (N == 0 || M == 0) return SCB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const long total = (long)N * M;
  long blocks = (total + 255) / 256;
  const long cap = (long)sm_count_of_current() * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
  switch (p->model) {
    case SCB_SINGLE_INTEGRATOR_2D: cbfqp_rows_kernel<SCB_SINGLE_INTEGRATOR_2D><<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
    case SCB_DYNAMIC_UNICYCLE_2D: cbfqp_rows_kernel<SCB_DYNAMIC_UNICYCLE_2D><<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
    case SCB_KINEMATIC_BICYCLE_2D: cbfqp_rows_kernel<SCB_KINEMATIC_BICYCLE_2D><<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
    case SCB_KINEMATIC_BICYCLE_2D_C3BF: cbfqp_rows_kernel<SCB_KINEMATIC_BICYCLE_2D_C3BF><<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
    case SCB_DOUBLE_INTEGRATOR_2D: cbfqp_rows_kernel<SCB_DOUBLE_INTEGRATOR_2D><<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
    case SCB_QUAD_2D: cbfqp_rows_kernel<SCB_QUAD_2D><<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
    case SCB_KINEMATIC_BICYCLE_2D_DPCBF: cbfqp_rows_kernel<SCB_KINEMATIC_BICYCLE_2D_DPCBF><<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
    case SCB_UNICYCLE_2D: cbfqp_rows_kernel<SCB_UNICYCLE_2D><<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
    case SCB_MANIPULATOR_2D: manip_rows_kernel<<<grid, 256, 0, s>>>(*p, N, M, X, OBS, stride, nobs, A, b); break;
  }
  CK(cudaGetLastError());
  return SCB_OK;
}

static int cbfqp_solve_impl(const scb_params* p, int N, int M, const double* X, const double* Uref, const double* OBS,
                            long stride, const int32_t* nobs, double* U, int32_t* status, uint64_t* active, void* stream,
                            bool eager, const int32_t* skip = nullptr);

int scb_cbfqp_solve(const scb_params* p, int N, int M, const double* X, const double* Uref, const double* OBS,
                    long stride, const int32_t* nobs, double* U, int32_t* status, uint64_t* active, void* stream) {
  return cbfqp_solve_impl(p, N, M, X, Uref, OBS, stride, nobs, U, status, active, stream, true);
}

// eager = false: the zero-copy host path (inputs are mapped host memory; rows beyond nobs must not cross PCIe)
static int cbfqp_solve_impl(const scb_params* p, int N, int M, const double* X, const double* Uref, const double* OBS,
                            long stride, const int32_t* nobs, double* U, int32_t* status, uint64_t* active, void* stream,
                            bool eager, const int32_t* skip) {
  if (!p || N < 0 || M < 0) return SCB_ERR_BAD_ARG;
  if (!qp_model_ok(p->model)) return p->model == SCB_QUAD_3D ? SCB_ERR_UNSUPPORTED : SCB_ERR_BAD_ARG;
  if (N == 0) return SCB_OK;
  if (!X || !Uref || !U || !status || (M > 0 && !OBS)) return SCB_ERR_BAD_ARG;
  int nx_m = 0, nu_m = 0;
  if (scb_model_dims(p->model, &nx_m, &nu_m) != SCB_OK) return SCB_ERR_BAD_ARG;
  const int words = scb_active_words(M, nu_m);
  cudaStream_t s = (cudaStream_t)stream;
  if (p->model == SCB_MANIPULATOR_2D) {                 // 3-input QP, warp per arm
    const int rows = M + 6, need = (rows + 31) / 32;
    const long blocks = ((long)N + 3) / 4, cap = (long)sm_count_of_current() * 16;
    const int grid = (int)(blocks < cap ? blocks : cap);
    if (need <= 1) manipqp_kernel<1><<<grid, kBlock, 0, s>>>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active, words);
    else if (need <= 2) manipqp_kernel<2><<<grid, kBlock, 0, s>>>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active, words);
    else if (need <= 4) manipqp_kernel<4><<<grid, kBlock, 0, s>>>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active, words);
    else return SCB_ERR_TOO_LARGE;
    CK(cudaGetLastError());
    return SCB_OK;
  }
  LaunchGeom g;
  if (!pick_geom(N, M + 2 * nu_m, sm_count_of_current(), g, forced_lanes())) return SCB_ERR_TOO_LARGE;
  int rc = SCB_ERR_BAD_ARG;
  switch (p->model) {
    case SCB_SINGLE_INTEGRATOR_2D: rc = launch_cbfqp_m<SCB_SINGLE_INTEGRATOR_2D>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, s, eager, skip); break;
    case SCB_DYNAMIC_UNICYCLE_2D: rc = launch_cbfqp_m<SCB_DYNAMIC_UNICYCLE_2D>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, s, eager, skip); break;
    case SCB_KINEMATIC_BICYCLE_2D: rc = launch_cbfqp_m<SCB_KINEMATIC_BICYCLE_2D>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, s, eager, skip); break;
    case SCB_KINEMATIC_BICYCLE_2D_C3BF: rc = launch_cbfqp_m<SCB_KINEMATIC_BICYCLE_2D_C3BF>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, s, eager, skip); break;
    case SCB_DOUBLE_INTEGRATOR_2D: rc = launch_cbfqp_m<SCB_DOUBLE_INTEGRATOR_2D>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, s, eager, skip); break;
    case SCB_QUAD_2D: rc = launch_cbfqp_m<SCB_QUAD_2D>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, s, eager, skip); break;
    case SCB_KINEMATIC_BICYCLE_2D_DPCBF: rc = launch_cbfqp_m<SCB_KINEMATIC_BICYCLE_2D_DPCBF>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, s, eager, skip); break;
    case SCB_UNICYCLE_2D: rc = launch_cbfqp_m<SCB_UNICYCLE_2D>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, status, active, words, s, eager, skip); break;
  }
  if (rc != SCB_OK) return rc;
  CK(cudaGetLastError());
  return SCB_OK;
}

// ------------------------------------------------------------------------------ optimal decay
static int odcbf_solve_impl(const scb_params* p, int N, int M, const double* X, const double* Uref, const double* OBS,
                            long stride, const int32_t* nobs, double* U, double* omega, int32_t* sel, int32_t* status,
                            uint64_t* active, void* stream, const int32_t* skip);

int scb_odcbf_solve(const scb_params* p, int N, int M, const double* X, const double* Uref, const double* OBS,
                    long stride, const int32_t* nobs, double* U, double* omega, int32_t* sel, int32_t* status,
                    uint64_t* active, void* stream) {
  return odcbf_solve_impl(p, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active, stream, nullptr);
}

static int odcbf_solve_impl(const scb_params* p, int N, int M, const double* X, const double* Uref, const double* OBS,
                            long stride, const int32_t* nobs, double* U, double* omega, int32_t* sel, int32_t* status,
                            uint64_t* active, void* stream, const int32_t* skip) {
  if (!p || N < 0 || M < 0) return SCB_ERR_BAD_ARG;
  if (p->model == SCB_SINGLE_INTEGRATOR_2D || p->model == SCB_QUAD_3D || p->model == SCB_DOUBLE_INTEGRATOR_2D || p->model == SCB_UNICYCLE_2D || p->model == SCB_MANIPULATOR_2D ||
      p->model == SCB_KINEMATIC_BICYCLE_2D_DPCBF)
    return SCB_ERR_UNSUPPORTED;                        // optimal_decay_cbf_qp.py:51-52 raises NotCompatibleError
  if (!qp_model_ok(p->model)) return SCB_ERR_BAD_ARG;
  if (N == 0) return SCB_OK;
  if (!X || !Uref || !U || !status || (M > 0 && !OBS)) return SCB_ERR_BAD_ARG;
  LaunchGeom g;
  if (!pick_geom(N, M > 1 ? M : 1, sm_count_of_current(), g)) return SCB_ERR_TOO_LARGE;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = SCB_ERR_BAD_ARG;
  switch (p->model) {
    case SCB_DYNAMIC_UNICYCLE_2D: rc = launch_od_m<SCB_DYNAMIC_UNICYCLE_2D, 2>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active, s, skip); break;
    case SCB_KINEMATIC_BICYCLE_2D: rc = launch_od_m<SCB_KINEMATIC_BICYCLE_2D, 2>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active, s, skip); break;
    case SCB_KINEMATIC_BICYCLE_2D_C3BF: rc = launch_od_m<SCB_KINEMATIC_BICYCLE_2D_C3BF, 1>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active, s, skip); break;
    case SCB_QUAD_2D: rc = launch_od_m<SCB_QUAD_2D, 2>(*p, g, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active, s, skip); break;
  }
  if (rc != SCB_OK) return rc;
  CK(cudaGetLastError());
  return SCB_OK;
}

// ------------------------------------------------------------------------------ MPC-CBF
size_t scb_mpccbf_workspace_bytes(int N) { return mpc_workspace_bytes(N); }

int scb_mpccbf_launch_count(const scb_params* p, int N, int M, int H, int with_workspace) {
  if (!p || N < 0 || M < 0 || H < 1) return SCB_ERR_BAD_ARG;
  if (N == 0) return 0;
  int n = 0;
  static int dummy;
  MpcIO io{};
  int rc = mpc_launch(*p, N, M, H, io, with_workspace ? (void*)&dummy : nullptr,
                      with_workspace ? mpc_workspace_bytes(N) : 0, nullptr, sm_count_of_current(), &n);
  return rc == SCB_OK ? n : rc;
}

int scb_mpccbf_solve(const scb_params* p, int N, int M, int H, const double* X, const double* Uref,
                     const double* goal, const double* u_prev, const int32_t* track, const double* OBS, long stride,
                     const int32_t* nobs, double* U, int32_t* status, double* pred_x, double* pred_u,
                     int32_t* iters, double* kkt, uint64_t* active, void* stream) {
  return scb_mpccbf_solve_ws(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status, pred_x, pred_u,
                             iters, kkt, active, nullptr, 0, stream);
}

int scb_mpccbf_solve_ws(const scb_params* p, int N, int M, int H, const double* X, const double* Uref,
                        const double* goal, const double* u_prev, const int32_t* track, const double* OBS, long stride,
                        const int32_t* nobs, double* U, int32_t* status, double* pred_x, double* pred_u,
                        int32_t* iters, double* kkt, uint64_t* active, void* workspace, size_t workspace_bytes,
                        void* stream) {
  if (!p || N < 0 || M < 0 || H < 1) return SCB_ERR_BAD_ARG;
  if (N == 0) return SCB_OK;
  if (!X || !goal || !u_prev || !U || !status || (M > 0 && !OBS)) return SCB_ERR_BAD_ARG;
  if (track && !Uref) return SCB_ERR_BAD_ARG;
  MpcIO io{X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status, pred_x, pred_u, iters, kkt, active, nullptr};
  int rc = mpc_launch(*p, N, M, H, io, workspace, workspace_bytes, (cudaStream_t)stream, sm_count_of_current());
  if (rc != SCB_OK) return rc;
  CK(cudaGetLastError());
  return SCB_OK;
}

// ------------------------------------------------------------------------------ closed loop
size_t scb_track_sizeof(void) { return sizeof(scb_track); }

int scb_select_obstacles(const scb_params* p, int N, int K, int M, const double* X, const double* yaw,
                         const double* SCENE, long sstride, double* OBS, int32_t* nobs, int32_t* idx, void* stream) {
  if (!p || N < 0 || K < 0 || M < 0) return SCB_ERR_BAD_ARG;
  if (p->model < 0 || p->model >= SCB_NUM_MODELS) return SCB_ERR_BAD_ARG;
  if (N == 0) return SCB_OK;
  if (!X || !nobs || (M > 0 && !OBS) || (K > 0 && !SCENE)) return SCB_ERR_BAD_ARG;
  if (K > kTrackMaxScene) return SCB_ERR_TOO_LARGE;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = track_grid(N, sm_count_of_current());
  const size_t smem = (size_t)(kTrackBlock / 32) * (size_t)(K > 0 ? K : 1) * sizeof(double);
#define SEL(MODEL) case MODEL: select_kernel<MODEL><<<grid, kTrackBlock, smem, s>>>(*p, N, K, M, X, yaw, SCENE, sstride, OBS, nobs, idx); break;
  switch (p->model) {
    SEL(SCB_SINGLE_INTEGRATOR_2D) SEL(SCB_DYNAMIC_UNICYCLE_2D) SEL(SCB_KINEMATIC_BICYCLE_2D)
    SEL(SCB_KINEMATIC_BICYCLE_2D_C3BF) SEL(SCB_QUAD_3D) SEL(SCB_KINEMATIC_BICYCLE_2D_DPCBF) SEL(SCB_DOUBLE_INTEGRATOR_2D) SEL(SCB_UNICYCLE_2D) SEL(SCB_QUAD_2D)
    default: return SCB_ERR_UNSUPPORTED;               // Manipulator2D: closed-loop laws not built
  }
#undef SEL
  CK(cudaGetLastError());
  return SCB_OK;
}

static int track_check(const scb_params* p, const scb_track* t) {
  if (!p || !t) return SCB_ERR_BAD_ARG;
  if (t->N < 0 || t->K < 0 || t->M < 0 || t->W < 1) return SCB_ERR_BAD_ARG;
  if (p->model < 0 || p->model >= SCB_NUM_MODELS) return SCB_ERR_BAD_ARG;
  if (t->controller < SCB_CTRL_CBF_QP || t->controller > SCB_CTRL_MPC_CBF) return SCB_ERR_BAD_ARG;
  if (t->K > kTrackMaxScene) return SCB_ERR_TOO_LARGE;
  if (t->N == 0) return SCB_OK;
  if (!t->X || !t->yaw || !t->sm || !t->wp_idx || !t->WP || !t->nwp || !t->goal || !t->has_goal || !t->u_att ||
      !t->ret || !t->done || !t->nsteps || !t->Uref || !t->nobs || !t->U || !t->status || (t->M > 0 && !t->OBS) ||
      (t->K > 0 && !t->SCENE))
    return SCB_ERR_BAD_ARG;
  if (t->controller == SCB_CTRL_MPC_CBF && (!t->u_prev || !t->track_flag || t->H < 1)) return SCB_ERR_BAD_ARG;
  if (p->model == SCB_MANIPULATOR_2D) return SCB_ERR_UNSUPPORTED;   // (the arm is not driven by LocalTrackingController)                        // loop laws not built yet
  if (t->controller != SCB_CTRL_MPC_CBF && p->model == SCB_QUAD_3D) return SCB_ERR_UNSUPPORTED;
  if (t->controller != SCB_CTRL_CBF_QP && p->model == SCB_KINEMATIC_BICYCLE_2D_DPCBF) return SCB_ERR_UNSUPPORTED;
  if (t->controller == SCB_CTRL_OPTIMAL_DECAY && (p->model == SCB_SINGLE_INTEGRATOR_2D || p->model == SCB_DOUBLE_INTEGRATOR_2D ||
                                                   p->model == SCB_UNICYCLE_2D))
    return SCB_ERR_UNSUPPORTED;
  return SCB_OK;
}

static int control_step_impl(const scb_params* p, const scb_track* t, cudaStream_t s) {
  const int smc = sm_count_of_current();
#define PRE(MODEL) case MODEL: launch_pre<MODEL>(*p, *t, s, smc); break;
  switch (p->model) {
    PRE(SCB_SINGLE_INTEGRATOR_2D) PRE(SCB_DYNAMIC_UNICYCLE_2D) PRE(SCB_KINEMATIC_BICYCLE_2D)
    PRE(SCB_KINEMATIC_BICYCLE_2D_C3BF) PRE(SCB_QUAD_3D) PRE(SCB_KINEMATIC_BICYCLE_2D_DPCBF) PRE(SCB_DOUBLE_INTEGRATOR_2D) PRE(SCB_UNICYCLE_2D) PRE(SCB_QUAD_2D)
  }
#undef PRE
  if (t->dynamic_obs && t->K > 0) dyn_obs_kernel<<<(t->K + 127) / 128, 128, 0, s>>>(t->SCENE, t->K, p->dt);
  int rc;
  const long stride = 7L * t->M;
  // agents that are done are skipped by the solve kernels (`done` for the QP controllers, track_flag < 0 for MPC): their
  // U / status / active stay those of their terminating step, exactly as in the fused single-launch path
  if (t->controller == SCB_CTRL_CBF_QP)
    rc = cbfqp_solve_impl(p, t->N, t->M, t->X, t->Uref, t->OBS, stride, t->nobs, t->U, t->status, t->active, s, true, t->done);
  else if (t->controller == SCB_CTRL_OPTIMAL_DECAY)
    rc = odcbf_solve_impl(p, t->N, t->M, t->X, t->Uref, t->OBS, stride, t->nobs, t->U, nullptr, nullptr, t->status,
                          t->active, s, t->done);
  else
    rc = scb_mpccbf_solve_ws(p, t->N, t->M, t->H, t->X, t->Uref, t->goal, t->u_prev, t->track_flag, t->OBS,
                             stride, t->nobs, t->U, t->status, nullptr, nullptr, t->mpc_iters, nullptr, nullptr, t->mpc_ws,
                             (size_t)t->mpc_ws_bytes, s);
  if (rc != SCB_OK) return rc;
#define POST(MODEL) case MODEL: launch_post<MODEL>(*p, *t, s, smc); break;
  switch (p->model) {
    POST(SCB_SINGLE_INTEGRATOR_2D) POST(SCB_DYNAMIC_UNICYCLE_2D) POST(SCB_KINEMATIC_BICYCLE_2D)
    POST(SCB_KINEMATIC_BICYCLE_2D_C3BF) POST(SCB_QUAD_3D) POST(SCB_KINEMATIC_BICYCLE_2D_DPCBF) POST(SCB_DOUBLE_INTEGRATOR_2D) POST(SCB_UNICYCLE_2D) POST(SCB_QUAD_2D)
  }
#undef POST
  CK(cudaGetLastError());
  return SCB_OK;
}

int scb_control_step(const scb_params* p, const scb_track* t, void* stream) {
  int rc = track_check(p, t);
  if (rc != SCB_OK || t->N == 0) return rc;
  return control_step_impl(p, t, (cudaStream_t)stream);
}

static bool fused_applicable(const scb_params* p, const scb_track* t) {
  const char* e = getenv("SCB_TRACK_FUSED");                 // tuning / A-B switch, read per call: "0" disables
  if (e && e[0] == '0') return false;
  if (t->K > kFusedMaxScene) return false;
  if (t->controller == SCB_CTRL_CBF_QP)
    return t->M + 4 <= 64 && (p->model == SCB_SINGLE_INTEGRATOR_2D || p->model == SCB_DYNAMIC_UNICYCLE_2D ||
                              p->model == SCB_KINEMATIC_BICYCLE_2D || p->model == SCB_KINEMATIC_BICYCLE_2D_C3BF ||
                              p->model == SCB_KINEMATIC_BICYCLE_2D_DPCBF || p->model == SCB_DOUBLE_INTEGRATOR_2D ||
                              p->model == SCB_UNICYCLE_2D || p->model == SCB_QUAD_2D);
  if (t->controller == SCB_CTRL_OPTIMAL_DECAY)
    return t->M <= 64 && (p->model == SCB_DYNAMIC_UNICYCLE_2D || p->model == SCB_KINEMATIC_BICYCLE_2D ||
                          p->model == SCB_KINEMATIC_BICYCLE_2D_C3BF);
  return false;
}

// fused single-launch path for the QP controllers (scb_track_kernels.cuh); false -> not applicable, use the 3-launch loop
static bool run_fused(const scb_params* p, const scb_track* t, int n_steps, cudaStream_t s) {
  if (!fused_applicable(p, t)) return false;
  if (t->controller == SCB_CTRL_CBF_QP) {
    switch (p->model) {
      case SCB_SINGLE_INTEGRATOR_2D: return launch_fused<SCB_SINGLE_INTEGRATOR_2D, SCB_CTRL_CBF_QP, 1>(*p, *t, n_steps, s);
      case SCB_DYNAMIC_UNICYCLE_2D: return launch_fused<SCB_DYNAMIC_UNICYCLE_2D, SCB_CTRL_CBF_QP, 1>(*p, *t, n_steps, s);
      case SCB_KINEMATIC_BICYCLE_2D: return launch_fused<SCB_KINEMATIC_BICYCLE_2D, SCB_CTRL_CBF_QP, 1>(*p, *t, n_steps, s);
      case SCB_KINEMATIC_BICYCLE_2D_C3BF: return launch_fused<SCB_KINEMATIC_BICYCLE_2D_C3BF, SCB_CTRL_CBF_QP, 1>(*p, *t, n_steps, s);
      case SCB_KINEMATIC_BICYCLE_2D_DPCBF: return launch_fused<SCB_KINEMATIC_BICYCLE_2D_DPCBF, SCB_CTRL_CBF_QP, 1>(*p, *t, n_steps, s);
      case SCB_DOUBLE_INTEGRATOR_2D: return launch_fused<SCB_DOUBLE_INTEGRATOR_2D, SCB_CTRL_CBF_QP, 1>(*p, *t, n_steps, s);
      case SCB_UNICYCLE_2D: return launch_fused<SCB_UNICYCLE_2D, SCB_CTRL_CBF_QP, 1>(*p, *t, n_steps, s);
      case SCB_QUAD_2D: return launch_fused<SCB_QUAD_2D, SCB_CTRL_CBF_QP, 1>(*p, *t, n_steps, s);
      default: return false;
    }
  }
  if (t->controller == SCB_CTRL_OPTIMAL_DECAY) {
    switch (p->model) {
      case SCB_DYNAMIC_UNICYCLE_2D: return launch_fused<SCB_DYNAMIC_UNICYCLE_2D, SCB_CTRL_OPTIMAL_DECAY, 2>(*p, *t, n_steps, s);
      case SCB_KINEMATIC_BICYCLE_2D: return launch_fused<SCB_KINEMATIC_BICYCLE_2D, SCB_CTRL_OPTIMAL_DECAY, 2>(*p, *t, n_steps, s);
      case SCB_KINEMATIC_BICYCLE_2D_C3BF: return launch_fused<SCB_KINEMATIC_BICYCLE_2D_C3BF, SCB_CTRL_OPTIMAL_DECAY, 1>(*p, *t, n_steps, s);
      default: return false;
    }
  }
  return false;
}

long scb_run_all_steps_launches(const scb_params* p, const scb_track* t, int n_steps) {
  if (track_check(p, t) != SCB_OK || t->N == 0 || n_steps <= 0) return 0;
  const long dyn = (t->dynamic_obs && t->K > 0) ? 1 : 0;
  if (fused_applicable(p, t)) return 1 + dyn;
  long solve = 1;
  if (t->controller == SCB_CTRL_MPC_CBF) {
    const int n = scb_mpccbf_launch_count(p, t->N, t->M, t->H, t->mpc_ws && t->mpc_ws_bytes >= mpc_workspace_bytes(t->N));
    if (n > 0) solve = n;
  }
  return (long)n_steps * (2 + solve + dyn);
}

int scb_run_all_steps(const scb_params* p, const scb_track* t, int n_steps, void* stream) {
  int rc = track_check(p, t);
  if (rc != SCB_OK || t->N == 0) return rc;
  if (n_steps < 0) return SCB_ERR_BAD_ARG;
  if (n_steps == 0) return SCB_OK;
  if (run_fused(p, t, n_steps, (cudaStream_t)stream)) {
    CK(cudaGetLastError());
    return SCB_OK;
  }
  for (int k = 0; k < n_steps; ++k) {
    rc = control_step_impl(p, t, (cudaStream_t)stream);
    if (rc != SCB_OK) return rc;
  }
  return SCB_OK;
}

#if defined(SCB_MPC_PROFILE)
// debug build only: per-phase cycle counters of warp 0 / block 0 (see scb_mpc.cuh)
int scb_debug_mpc_profile(long long* out, int reset) {
  if (out) CK(cudaMemcpyFromSymbol(out, g_mpc_prof, sizeof(long long) * 24));
  if (reset) { long long z[24] = {0}; CK(cudaMemcpyToSymbol(g_mpc_prof, z, sizeof(z))); }
  return SCB_OK;
}
#endif

// ------------------------------------------------------------------------------ measurement helper
// FP64 FMA throughput of the current device: 16 independent DFMA chains per thread, 2048 threads per SM.  This is the
// measured denominator of the MPC kernels' roofline (bench.py: roofline.bound = "fp64"); MEASURED_PEAKS.json has no FP64 entry.
__global__ void __launch_bounds__(512) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (double)(threadIdx.x + i);
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double sum = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) sum += acc[i];
  if (sum == 12345.678) out[0] = sum;          // never true: keeps the chains alive
}

extern "C" int scb_measure_fp64_peak(double* tflops, void* stream) {
  if (!tflops) return SCB_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  double* d = nullptr;
  CK(cudaMalloc((void**)&d, sizeof(double)));
  const int sms = sm_count_of_current(), blocks = sms * 4, iters = 2048;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {         // first repetition warms up clocks / instruction cache
    CK(cudaEventRecord(e0, s));
    fp64_peak_kernel<<<blocks, 512, 0, s>>>(d, iters, 0.999999, 1e-7);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 16.0 * iters * 512.0 * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  CK(cudaGetLastError());
  *tflops = best;
  return SCB_OK;
}

// Latency floor of a small launch on this device (bench.py: roofline.latency_floor): what a kernel of cfg2's geometry
// (256 CTAs x 128 threads) costs inside a CUDA graph when it does (a) nothing, (b) one and (c) two DEPENDENT cold-DRAM
// round trips per warp followed by a store.  (c) - (b) is the price of one dependent DRAM round trip; the cfg2 kernel
// has two (nobs / state, then rows are already in flight) plus ~700 dependent instructions.
__global__ void floor_empty_kernel(int* sink) { if (sink && threadIdx.x == 9999) *sink = 0; }
__global__ void floor_init_kernel(unsigned* buf, unsigned n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    buf[i] = (unsigned)(((unsigned long long)i * 2654435761ull + 12345ull) % n);
}
__global__ void floor_chase_kernel(const unsigned* __restrict__ buf, unsigned n, unsigned offset, int hops, unsigned* out) {
  const unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned idx = (offset + w * 1000003u) % n;
  for (int h = 0; h < hops; ++h) idx = buf[idx];
  if ((threadIdx.x & 31) == 0) out[w] = idx;
}

extern "C" int scb_measure_latency_floor(double* empty_us, double* one_trip_us, double* two_trip_us, void* stream) {
  if (!empty_us || !one_trip_us || !two_trip_us) return SCB_ERR_BAD_ARG;
  const unsigned n = 128u << 20;                       // 512 MB of uint32: 4x L2
  const int K = 200, grid = 256, block = 128;
  unsigned *buf = nullptr, *out = nullptr;
  CK(cudaMalloc((void**)&buf, (size_t)n * 4));
  CK(cudaMalloc((void**)&out, (size_t)grid * block / 32 * 4));
  cudaStream_t s = nullptr;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  (void)stream;
  floor_init_kernel<<<1024, 256, 0, s>>>(buf, n);
  CK(cudaStreamSynchronize(s));
  double res[3] = {0, 0, 0};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int mode = 0; mode < 3; ++mode) {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int k = 0; k < K; ++k) {
      if (mode == 0) floor_empty_kernel<<<grid, block, 0, s>>>(nullptr);
      else floor_chase_kernel<<<grid, block, 0, s>>>(buf, n, (unsigned)(k * 7919u * 4099u), mode, out);
    }
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, s)); CK(cudaStreamSynchronize(s));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ge, s)); CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
      float ms = 0.f; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    res[mode] = (double)best * 1e3 / K;
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s); cudaFree(buf); cudaFree(out);
  *empty_us = res[0]; *one_trip_us = res[1]; *two_trip_us = res[2];
  return SCB_OK;
}

// ------------------------------------------------------------------------------ host-pointer context
struct scb_ctx {
  int device;
  cudaStream_t stream;
  char* dbuf;
  size_t dcap;
  long launches;
};

int scb_ctx_create(scb_ctx** out, int device) {
  if (!out) return SCB_ERR_BAD_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return SCB_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return SCB_ERR_BAD_ARG;
  CK(cudaSetDevice(device));
  scb_ctx* c = new (std::nothrow) scb_ctx();
  if (!c) return SCB_ERR_ALLOC;
  c->device = device; c->dbuf = nullptr; c->dcap = 0; c->launches = 0;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return cuda_fail(e); }
  *out = c;
  return SCB_OK;
}

void scb_ctx_destroy(scb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->dbuf) cudaFree(c->dbuf);
  cudaStreamDestroy(c->stream);
  delete c;
}

long scb_ctx_launches(const scb_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"

// bump allocator over the context's device buffer (256-byte aligned pieces)
struct Carver {
  char* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    T* p = (T*)(base + off);
    off += (n * sizeof(T) + 255) & ~(size_t)255;
    return p;
  }
};
static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

static int ctx_reserve(scb_ctx* c, size_t bytes) {
  if (bytes <= c->dcap) return SCB_OK;
  if (c->dbuf) { cudaFree(c->dbuf); c->dbuf = nullptr; c->dcap = 0; }
  size_t want = bytes + bytes / 4;
  cudaError_t e = cudaMalloc((void**)&c->dbuf, want);
  if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? SCB_ERR_ALLOC : cuda_fail(e);
  c->dcap = want;
  return SCB_OK;
}

// ---- zero-copy ("mapped") variant of the host-pointer calls ---------------------------------------------
// When every caller buffer is page-locked host memory (cudaHostAlloc / cudaHostRegister, e.g. torch pinned tensors)
// it is already mapped into the device address space (UVA): the kernel then reads its inputs and writes its outputs
// over PCIe directly, which removes the 4 + 3 staged cudaMemcpyAsync (and their per-copy latency) from a call that
// moves ~1 MB.  Pageable buffers, or SCB_HOST_PATH=staged, take the staged path; SCB_HOST_PATH=mapped forces the
// mapped path whenever it is possible.  Above kMappedMaxBytes the DMA engines win and `auto` stages.
constexpr size_t kMappedMaxBytes = 32u << 20;

static int host_path_mode() {          // 0 auto, 1 staged, 2 mapped
  const char* e = getenv("SCB_HOST_PATH");
  if (!e) return 0;
  if (strcmp(e, "staged") == 0) return 1;
  if (strcmp(e, "mapped") == 0) return 2;
  return 0;
}

// device alias of a page-locked host pointer; false if `h` is not page-locked (NULL maps to NULL)
template <typename T>
static bool mapped_alias(T* h, T** d) {
  *d = nullptr;
  if (!h) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, (const void*)h) != cudaSuccess) { cudaGetLastError(); return false; }
  if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;
  *d = (T*)a.devicePointer;
  return true;
}

static bool use_mapped(size_t bytes) {
  const int mode = host_path_mode();
  if (mode == 1) return false;
  return mode == 2 || bytes <= kMappedMaxBytes;
}

#define H2D(dst, src, n, T) CK(cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(T), cudaMemcpyHostToDevice, c->stream))
#define D2H(dst, src, n, T) CK(cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(T), cudaMemcpyDeviceToHost, c->stream))

extern "C" {

int scb_cbfqp_solve_host(scb_ctx* c, const scb_params* p, int N, int M, const double* X, const double* Uref,
                         const double* OBS, long stride, const int32_t* nobs, double* U, int32_t* status,
                         uint64_t* active) {
  if (!c || !p || N < 0 || M < 0) return SCB_ERR_BAD_ARG;
  if (N == 0) return SCB_OK;
  if (!X || !Uref || !U || !status || (M > 0 && !OBS)) return SCB_ERR_BAD_ARG;
  if (!qp_model_ok(p->model)) return p->model == SCB_QUAD_3D ? SCB_ERR_UNSUPPORTED : SCB_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  int nx = 0, nu = 0;
  if (scb_model_dims(p->model, &nx, &nu) != SCB_OK) return SCB_ERR_BAD_ARG;     // sizes from the validated model id
  const int words = scb_active_words(M, nu);
  const size_t nobs_el = (stride == 0) ? (size_t)M * 7 : (size_t)N * (size_t)stride;
  size_t need = padded((size_t)N * nx * 8) + padded((size_t)N * nu * 8) * 2 + padded(nobs_el * 8) +
                padded((size_t)N * 4) * 2 + padded((size_t)N * words * 8);
  int rc;
  if (use_mapped(need)) {
    const double *mX, *mUr, *mO; const int32_t* mN; double* mU; int32_t* mS; uint64_t* mA;
    if (mapped_alias(X, &mX) && mapped_alias(Uref, &mUr) && mapped_alias(OBS, &mO) && mapped_alias(nobs, &mN) &&
        mapped_alias(U, &mU) && mapped_alias(status, &mS) && mapped_alias(active, &mA)) {
      rc = cbfqp_solve_impl(p, N, M, mX, mUr, mO, stride, mN, mU, mS, mA, c->stream, false);
      if (rc != SCB_OK) return rc;
      c->launches += 1;
      CK(cudaStreamSynchronize(c->stream));
      return SCB_OK;
    }
  }
  rc = ctx_reserve(c, need);
  if (rc != SCB_OK) return rc;
  Carver cv{c->dbuf, 0};
  double* dX = cv.take<double>((size_t)N * nx);
  double* dUr = cv.take<double>((size_t)N * nu);
  double* dU = cv.take<double>((size_t)N * nu);
  double* dO = cv.take<double>(nobs_el);
  int32_t* dN = cv.take<int32_t>(N);
  int32_t* dS = cv.take<int32_t>(N);
  uint64_t* dA = cv.take<uint64_t>((size_t)N * words);
  H2D(dX, X, (size_t)N * nx, double);
  H2D(dUr, Uref, (size_t)N * nu, double);
  if (nobs_el) H2D(dO, OBS, nobs_el, double);
  if (nobs) H2D(dN, nobs, N, int32_t);
  rc = scb_cbfqp_solve(p, N, M, dX, dUr, dO, stride, nobs ? dN : nullptr, dU, dS, active ? dA : nullptr, c->stream);
  if (rc != SCB_OK) return rc;
  c->launches += 1;
  D2H(U, dU, (size_t)N * nu, double);
  D2H(status, dS, N, int32_t);
  if (active) D2H(active, dA, (size_t)N * words, uint64_t);
  CK(cudaStreamSynchronize(c->stream));
  return SCB_OK;
}

int scb_odcbf_solve_host(scb_ctx* c, const scb_params* p, int N, int M, const double* X, const double* Uref,
                         const double* OBS, long stride, const int32_t* nobs, double* U, double* omega, int32_t* sel,
                         int32_t* status, uint64_t* active) {
  if (!c || !p || N < 0 || M < 0) return SCB_ERR_BAD_ARG;
  if (N == 0) return SCB_OK;
  if (!X || !Uref || !U || !status || (M > 0 && !OBS)) return SCB_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  const size_t nobs_el = (stride == 0) ? (size_t)M * 7 : (size_t)N * (size_t)stride;
  int nx = 0, nu_unused = 0;
  if (scb_model_dims(p->model, &nx, &nu_unused) != SCB_OK) return SCB_ERR_BAD_ARG;   // sizes from the validated model id
  size_t need = padded((size_t)N * nx * 8) + padded((size_t)N * 2 * 8) * 3 + padded(nobs_el * 8) +
                padded((size_t)N * 4) * 3 + padded((size_t)N * 8);
  int rc;
  if (use_mapped(need)) {
    const double *mX, *mUr, *mO; const int32_t* mN; double *mU, *mW; int32_t *mS, *mSel; uint64_t* mA;
    if (mapped_alias(X, &mX) && mapped_alias(Uref, &mUr) && mapped_alias(OBS, &mO) && mapped_alias(nobs, &mN) &&
        mapped_alias(U, &mU) && mapped_alias(omega, &mW) && mapped_alias(sel, &mSel) && mapped_alias(status, &mS) &&
        mapped_alias(active, &mA)) {
      rc = scb_odcbf_solve(p, N, M, mX, mUr, mO, stride, mN, mU, mW, mSel, mS, mA, c->stream);
      if (rc != SCB_OK) return rc;
      c->launches += 1;
      CK(cudaStreamSynchronize(c->stream));
      return SCB_OK;
    }
  }
  rc = ctx_reserve(c, need);
  if (rc != SCB_OK) return rc;
  Carver cv{c->dbuf, 0};
  double* dX = cv.take<double>((size_t)N * nx);
  double* dUr = cv.take<double>((size_t)N * 2);
  double* dU = cv.take<double>((size_t)N * 2);
  double* dW = cv.take<double>((size_t)N * 2);
  double* dO = cv.take<double>(nobs_el);
  int32_t* dN = cv.take<int32_t>(N);
  int32_t* dS = cv.take<int32_t>(N);
  int32_t* dSel = cv.take<int32_t>(N);
  uint64_t* dA = cv.take<uint64_t>(N);
  H2D(dX, X, (size_t)N * nx, double);
  H2D(dUr, Uref, (size_t)N * 2, double);
  if (nobs_el) H2D(dO, OBS, nobs_el, double);
  if (nobs) H2D(dN, nobs, N, int32_t);
  rc = scb_odcbf_solve(p, N, M, dX, dUr, dO, stride, nobs ? dN : nullptr, dU, omega ? dW : nullptr,
                       sel ? dSel : nullptr, dS, active ? dA : nullptr, c->stream);
  if (rc != SCB_OK) return rc;
  c->launches += 1;
  D2H(U, dU, (size_t)N * 2, double);
  if (omega) D2H(omega, dW, (size_t)N * 2, double);
  if (sel) D2H(sel, dSel, N, int32_t);
  D2H(status, dS, N, int32_t);
  if (active) D2H(active, dA, N, uint64_t);
  CK(cudaStreamSynchronize(c->stream));
  return SCB_OK;
}

int scb_mpccbf_solve_host(scb_ctx* c, const scb_params* p, int N, int M, int H, const double* X, const double* Uref,
                          const double* goal, const double* u_prev, const int32_t* track, const double* OBS,
                          long stride, const int32_t* nobs, double* U, int32_t* status, double* pred_x,
                          double* pred_u, int32_t* iters, double* kkt, uint64_t* active) {
  if (!c || !p || N < 0 || M < 0 || H < 1) return SCB_ERR_BAD_ARG;
  if (N == 0) return SCB_OK;
  if (!X || !goal || !u_prev || !U || !status || (M > 0 && !OBS) || (track && !Uref)) return SCB_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  int nx = 0, nu = 0;
  if (scb_model_dims(p->model, &nx, &nu) != SCB_OK) return SCB_ERR_BAD_ARG;      // sizes from the validated model id
  if (p->od_mpc) nu += 2;                                  // optimal decay: [u, omega1, omega2] per stage
  const int ng = (p->model == SCB_QUAD_3D) ? 3 : 2;
  const int aw = active ? scb_mpc_active_words(p, M, H) : 0;
  if (aw < 0) return aw;
  const size_t nobs_el = (stride == 0) ? (size_t)M * 7 : (size_t)N * (size_t)stride;
  size_t need = padded((size_t)N * aw * 8) + padded((size_t)N * nx * 8) + padded((size_t)N * nu * 8) * 3 + padded((size_t)N * ng * 8) +
                padded(nobs_el * 8) + padded((size_t)N * 4) * 4 + padded((size_t)N * 8) +
                padded((size_t)N * (H + 1) * nx * 8) + padded((size_t)N * H * nu * 8) + padded(mpc_workspace_bytes(N));
  int rc = ctx_reserve(c, need);
  if (rc != SCB_OK) return rc;
  Carver cv{c->dbuf, 0};
  double* dX = cv.take<double>((size_t)N * nx);
  double* dUr = cv.take<double>((size_t)N * nu);
  double* dUp = cv.take<double>((size_t)N * nu);
  double* dU = cv.take<double>((size_t)N * nu);
  double* dG = cv.take<double>((size_t)N * ng);
  double* dO = cv.take<double>(nobs_el);
  int32_t* dN = cv.take<int32_t>(N);
  int32_t* dT = cv.take<int32_t>(N);
  int32_t* dS = cv.take<int32_t>(N);
  int32_t* dI = cv.take<int32_t>(N);
  double* dK = cv.take<double>(N);
  double* dPx = cv.take<double>((size_t)N * (H + 1) * nx);
  double* dPu = cv.take<double>((size_t)N * H * nu);
  char* dWs = cv.take<char>(mpc_workspace_bytes(N));
  uint64_t* dAct = cv.take<uint64_t>((size_t)N * aw);
  H2D(dX, X, (size_t)N * nx, double);
  if (Uref) H2D(dUr, Uref, (size_t)N * nu, double);
  H2D(dUp, u_prev, (size_t)N * nu, double);
  H2D(dG, goal, (size_t)N * ng, double);
  if (nobs_el) H2D(dO, OBS, nobs_el, double);
  if (nobs) H2D(dN, nobs, N, int32_t);
  if (track) H2D(dT, track, N, int32_t);
  rc = scb_mpccbf_solve_ws(p, N, M, H, dX, Uref ? dUr : nullptr, dG, dUp, track ? dT : nullptr, dO, stride,
                           nobs ? dN : nullptr, dU, dS, pred_x ? dPx : nullptr, pred_u ? dPu : nullptr,
                           iters ? dI : nullptr, kkt ? dK : nullptr, active ? dAct : nullptr, dWs, mpc_workspace_bytes(N),
                           c->stream);
  if (rc != SCB_OK) return rc;
  c->launches += scb_mpccbf_launch_count(p, N, M, H, 1);
  D2H(U, dU, (size_t)N * nu, double);
  D2H(status, dS, N, int32_t);
  if (pred_x) D2H(pred_x, dPx, (size_t)N * (H + 1) * nx, double);
  if (pred_u) D2H(pred_u, dPu, (size_t)N * H * nu, double);
  if (iters) D2H(iters, dI, N, int32_t);
  if (kkt) D2H(kkt, dK, N, double);
  if (active) D2H(active, dAct, (size_t)N * aw, uint64_t);
  CK(cudaStreamSynchronize(c->stream));
  return SCB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------- Backup-CBF QP
static int backup_check(const scb_backup_params* p, int N, int K) {
  if (!p || N < 0 || K < 0) return SCB_ERR_BAD_ARG;
  if (p->n_backup < 1 || !(p->dt > 0.0) || !(p->a_max > 0.0) || !(p->q0 > 0.0) || !(p->q1 > 0.0)) return SCB_ERR_BAD_ARG;
  if (p->n_backup + 4 > 256) return SCB_ERR_TOO_LARGE;
  return SCB_OK;
}

extern "C" int scb_backupcbf_solve(const scb_backup_params* p, int N, int K, const double* X, const double* Uref,
                                   const double* MOV, long mov_stride, double* U, int32_t* status, int32_t* intervene,
                                   double* h_min, double* phi, double* rows, uint64_t* active, void* stream) {
  int rc = backup_check(p, N, K);
  if (rc != SCB_OK) return rc;
  if (N == 0) return SCB_OK;
  if (!X || !Uref || !U || !status || (K > 0 && !MOV) || mov_stride < 0 || (mov_stride > 0 && mov_stride < (long)K * kBkMov)) return SCB_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = p->n_backup, words = scb_backup_active_words(nb);
  int forced = 0;
  if (const char* e = getenv("SCB_BK_LANES")) forced = atoi(e);
  if (rows && h_min && !getenv("SCB_BK_FUSED")) {
    // two launches (the caller gave the [N, n_backup, 3] row buffer): rollout with 8 lanes per agent (5 carry the step
    // variants, 4 the barrier variants) -- a warp per agent when the batch cannot fill the device anyway -- then the QP
    // rollout geometry: 5 lanes per agent -- one per step variant, six agents per warp (2.46 ms per 65 536 agents against 3.26
    // with groups of 8 and 9.96 with a warp per agent; 0.46 / 0.49 / 0.50 ms at 2048) -- a warp per agent for a handful of agents
    const int rl = (forced == 5 || forced == 8 || forced == 32) ? forced : (N >= 256 ? 5 : 32);
    const int per_cta = (rl == 5) ? (kBkBlock / 32) * 6 : kBkBlock / rl;
    const unsigned rgrid = (unsigned)((N + per_cta - 1) / per_cta);
    if (rl == 5)      backup_rollout_kernel<5><<<rgrid, kBkBlock, 0, s>>>(*p, N, K, X, K > 0 ? MOV : nullptr, mov_stride, h_min, phi, rows);
    else if (rl == 8) backup_rollout_kernel<8><<<rgrid, kBkBlock, 0, s>>>(*p, N, K, X, K > 0 ? MOV : nullptr, mov_stride, h_min, phi, rows);
    else              backup_rollout_kernel<32><<<rgrid, kBkBlock, 0, s>>>(*p, N, K, X, K > 0 ? MOV : nullptr, mov_stride, h_min, phi, rows);
    const int ql = (forced == 8 && nb + 4 <= 128) ? 8 : ((forced == 32 || nb + 4 > 128 || (long)N * 32 <= (long)sm_count_of_current() * 512) ? 32 : 8);
    const unsigned qgrid = (unsigned)((N + kBkBlock / ql - 1) / (kBkBlock / ql));
#define GQ(L, R) backup_qp_kernel<L, R><<<qgrid, kBkBlock, 0, s>>>(*p, N, X, Uref, rows, h_min, U, status, intervene, active, words);
    if (ql == 8) {
      if (nb + 4 <= 64) GQ(8, 8) else GQ(8, 16)
    } else {
      if (nb + 4 <= 64) GQ(32, 2) else if (nb + 4 <= 128) GQ(32, 4) else GQ(32, 8)
    }
#undef GQ
    CK(cudaGetLastError());
    return SCB_OK;
  }
  // one fused launch: lane group of 8 for batches that fill the device, a warp per agent for small ones (rows spread
  // thinner, shorter QP scan) and for horizons whose rows do not fit 8 x 16 registers
  int lanes = (nb + 4 <= 128 && (long)N * 32 > (long)sm_count_of_current() * 512) ? 8 : 32;
  if (forced == 8 && nb + 4 <= 128) lanes = 8;
  if (forced == 32) lanes = 32;
  const int groups = kBkBlock / lanes;
  const size_t smem = (size_t)groups * bk_agent_doubles(nb) * sizeof(double);
  const unsigned grid = (unsigned)((N + groups - 1) / groups);
#define GO(L, R)                                                                                                          \
  {                                                                                                                       \
    auto kern = backupcbf_kernel<L, R>;                                                                                   \
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
      return SCB_ERR_TOO_LARGE;                                                                                           \
    kern<<<grid, kBkBlock, smem, s>>>(*p, N, K, X, Uref, K > 0 ? MOV : nullptr, mov_stride, U, status, intervene, h_min, phi, \
                                      rows, active, words);                                                               \
  }
  if (lanes == 8) {
    if (nb + 4 <= 64) GO(8, 8) else GO(8, 16)
  } else {
    if (nb + 4 <= 64) GO(32, 2) else if (nb + 4 <= 128) GO(32, 4) else GO(32, 8)
  }
#undef GO
  CK(cudaGetLastError());
  return SCB_OK;
}

extern "C" int scb_backupcbf_solve_host(scb_ctx* c, const scb_backup_params* p, int N, int K, const double* X,
                                        const double* Uref, const double* MOV, long mov_stride, double* U, int32_t* status,
                                        int32_t* intervene, double* h_min, double* phi, double* rows, uint64_t* active) {
  if (!c) return SCB_ERR_BAD_ARG;
  int rc = backup_check(p, N, K);
  if (rc != SCB_OK) return rc;
  if (N == 0) return SCB_OK;
  if (!X || !Uref || !U || !status || (K > 0 && !MOV) || mov_stride < 0 || (mov_stride > 0 && mov_stride < (long)K * kBkMov)) return SCB_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  const int nb = p->n_backup, words = scb_backup_active_words(nb);
  const size_t mov_el = (K == 0) ? 0 : (mov_stride == 0 ? (size_t)K * kBkMov : (size_t)N * (size_t)mov_stride);
  const size_t need = padded((size_t)N * 4 * 8) + padded((size_t)N * 2 * 8) * 2 + padded(mov_el * 8) + padded((size_t)N * 4) * 2 +
                      padded((size_t)N * 8) + padded((size_t)N * nb * 4 * 8) + padded((size_t)N * nb * 3 * 8) +
                      padded((size_t)N * words * 8);
  rc = ctx_reserve(c, need);
  if (rc != SCB_OK) return rc;
  Carver cv{c->dbuf, 0};
  double* dX = cv.take<double>((size_t)N * 4);
  double* dUr = cv.take<double>((size_t)N * 2);
  double* dU = cv.take<double>((size_t)N * 2);
  double* dM = cv.take<double>(mov_el);
  int32_t* dS = cv.take<int32_t>(N);
  int32_t* dI = cv.take<int32_t>(N);
  double* dH = cv.take<double>(N);
  double* dP = cv.take<double>((size_t)N * nb * 4);
  double* dR = cv.take<double>((size_t)N * nb * 3);
  uint64_t* dA = cv.take<uint64_t>((size_t)N * words);
  H2D(dX, X, (size_t)N * 4, double);
  H2D(dUr, Uref, (size_t)N * 2, double);
  if (mov_el) H2D(dM, MOV, mov_el, double);
  rc = scb_backupcbf_solve(p, N, K, dX, dUr, mov_el ? dM : nullptr, mov_stride, dU, dS, intervene ? dI : nullptr, dH,
                           phi ? dP : nullptr, dR, active ? dA : nullptr, c->stream);       // (row buffer given: two launches)
  if (rc != SCB_OK) return rc;
  c->launches += 2;
  D2H(U, dU, (size_t)N * 2, double);
  D2H(status, dS, N, int32_t);
  if (intervene) D2H(intervene, dI, N, int32_t);
  if (h_min) D2H(h_min, dH, N, double);
  if (phi) D2H(phi, dP, (size_t)N * nb * 4, double);
  if (rows) D2H(rows, dR, (size_t)N * nb * 3, double);
  if (active) D2H(active, dA, (size_t)N * words, uint64_t);
  CK(cudaStreamSynchronize(c->stream));
  return SCB_OK;
}

// ---------------------------------------------------------------------------------------- gatekeeper / MPS
extern "C" int scb_shield_step(const scb_shield_params* p, const scb_shield_state* st, int N, int K, const double* X,
                               const double* NOMX, const double* NOMU, const int32_t* nom_len, const double* MOV,
                               long mov_stride, const double* STAT, double* U, int32_t* using_backup, void* stream) {
  if (!p || !st) return SCB_ERR_BAD_ARG;
  int rc = backup_check(&p->scene, N, K);
  if (rc != SCB_OK) return rc;
  if (p->nom_cap < 0 || (p->mode != 0 && p->mode != 1) || !(p->event_offset >= 0.0)) return SCB_ERR_BAD_ARG;
  if (N == 0) return SCB_OK;
  if (!X || !U || !st->CU || !st->clen || !st->cidx || !st->nsteps || !st->next_event || !st->cbuf || (p->nom_cap > 0 && (!NOMX || !NOMU)) ||
      (K > 0 && !MOV) || mov_stride < 0 || (mov_stride > 0 && mov_stride < (long)K * kBkMov))
    return SCB_ERR_BAD_ARG;
  if (p->nom_cap == 0 && !NOMX) return SCB_ERR_BAD_ARG;            // (NOMX always holds at least the start state)
  cudaStream_t s = (cudaStream_t)stream;
  // gatekeeper: one candidate per lane (T / discount + 2 of them); MPS has a single candidate: a thread per agent
  // (8 lanes beat 32 at 22 candidates: 3.8 vs 4.7 ms per 65 536-agent step -- the longest nominal horizon is valid for most
  // agents, so most of a warp's 22 speculative rollouts are wasted; with 8 lanes a lane walks candidates c, c + 8, c + 16)
  int lanes = p->mode == 1 ? 1 : 8;
  if (p->mode == 0 && (long)N * 8 <= (long)sm_count_of_current() * 64) lanes = 32;      // few agents: latency, not throughput
  if (const char* e = getenv("SCB_SHIELD_LANES")) { const int v = atoi(e); if (v == 1 || v == 8 || v == 32) lanes = v; }
  const double* mov = K > 0 ? MOV : nullptr;
  // two-launch search (gatekeeper, work list given, batch large enough to matter; SCB_SHIELD_TWO_PHASE=0/1 overrides)
  bool two = p->mode == 0 && st->work && N >= 1024;
  if (const char* e = getenv("SCB_SHIELD_TWO_PHASE")) two = p->mode == 0 && st->work && e[0] == '1';
  if (two) {
    CK(cudaMemsetAsync(st->work, 0, sizeof(int32_t), s));
    shield_step_kernel<1, 1><<<(unsigned)((N + kBkBlock - 1) / kBkBlock), kBkBlock, 0, s>>>(*p, *st, N, K, X, NOMX, NOMU, nom_len, mov, mov_stride, STAT, U, using_backup, st->work);
    const int l2 = lanes == 32 ? 32 : 8;
    const unsigned g2 = (unsigned)((N + kBkBlock / l2 - 1) / (kBkBlock / l2));        // sized for "every agent pending"; surplus groups exit at once
    if (l2 == 32) shield_step_kernel<32, 2><<<g2, kBkBlock, 0, s>>>(*p, *st, N, K, X, NOMX, NOMU, nom_len, mov, mov_stride, STAT, U, using_backup, st->work);
    else          shield_step_kernel<8, 2><<<g2, kBkBlock, 0, s>>>(*p, *st, N, K, X, NOMX, NOMU, nom_len, mov, mov_stride, STAT, U, using_backup, st->work);
    CK(cudaGetLastError());
    return SCB_OK;
  }
  const int groups = kBkBlock / lanes;
  const unsigned grid = (unsigned)((N + groups - 1) / groups);
  if (lanes == 32)     shield_step_kernel<32, 0><<<grid, kBkBlock, 0, s>>>(*p, *st, N, K, X, NOMX, NOMU, nom_len, mov, mov_stride, STAT, U, using_backup, nullptr);
  else if (lanes == 8) shield_step_kernel<8, 0><<<grid, kBkBlock, 0, s>>>(*p, *st, N, K, X, NOMX, NOMU, nom_len, mov, mov_stride, STAT, U, using_backup, nullptr);
  else                 shield_step_kernel<1, 0><<<grid, kBkBlock, 0, s>>>(*p, *st, N, K, X, NOMX, NOMU, nom_len, mov, mov_stride, STAT, U, using_backup, nullptr);
  CK(cudaGetLastError());
  return SCB_OK;
}
