// scb_mpc_kernels.cuh -- __global__ wrapper + launch for the MPC-CBF path (scb_mpc.cuh).
//
// One lane group (LANES = 32: a full warp) per agent; each group owns a private workspace of
// MpcLayout::total doubles in dynamic shared memory (~35 KB at H = 8, M = 16), so a CTA carries as
// many agents as fit in the SM's 227 KB and the grid is persistent (one CTA per SM, agents
// grid-strided).  Everything the interior-point loop touches after the initial obstacle load
// stays on chip; HBM sees the inputs once and U/status once.
#pragma once
// (the kernel and its per-model launcher live in scb_mpc_impl.cuh; each model is its own translation unit,
// scb_mpc_inst.cu compiled with -DSCB_MPC_INST=<model id>, so the library builds in parallel)

#include <cuda_runtime.h>

#include "scb_core.cuh"

namespace scb {

constexpr int kMpcMaxObs = 64;
constexpr int kMpcMaxH = 16;
// defined in scb_mpc_impl.cuh, explicitly instantiated per model (scb_mpc_inst.cu)
template <int MODEL>
int mpc_launch_m(const scb_params& p, int N, int M, int H, const double* X, const double* Uref,
                 const double* goal, const double* u_prev, const int32_t* track, const double* OBS, long stride,
                 const int32_t* nobs, double* U, int32_t* status, double* pred_x, double* pred_u,
                 int32_t* iters, double* kkt, int* counter, void* workspace, size_t workspace_bytes, cudaStream_t s,
                 int sm_count, int* count_only);

// scheduling scratch of one launch: histogram [1024] + bin per agent [N] + schedule [N], int32
inline size_t mpc_workspace_bytes(long N) { return (size_t)(1024 + 2 * (N > 0 ? N : 0)) * sizeof(int32_t); }

// Per-launch work counters: a small ring of zero-initialised ints per device; each launch takes the next slot and
// re-zeroes it with a stream-ordered memset just before the kernel, so concurrent streams and CUDA-graph capture are
// safe (a slot is only reused after kRing further launches on that device).
constexpr int kRing = 4096;
inline int* mpc_counter_slot(cudaStream_t s) {
  static int* ring[64] = {nullptr};
  static unsigned next[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!ring[dev]) {
    if (cudaMalloc((void**)&ring[dev], kRing * sizeof(int)) != cudaSuccess) { ring[dev] = nullptr; return nullptr; }
    cudaMemset(ring[dev], 0, kRing * sizeof(int));
  }
  int* slot = ring[dev] + (__sync_fetch_and_add(&next[dev], 1u) % kRing);
  if (cudaMemsetAsync(slot, 0, sizeof(int), s) != cudaSuccess) return nullptr;
  return slot;
}

#ifndef SCB_MPC_NO_DISPATCH
constexpr int kMpcSeBaseId = 100;       // == kMpcSeBase (scb_mpc.cuh): general-row variants of SI / DU / DI for superellipsoid rows
// fast path + (when p.mpc_superellipsoid) the general-row launch for the agents that have a superellipsoid row
template <int MODEL>
inline int mpc_launch_se(const scb_params& p, int N, int M, int H, const double* X, const double* Uref, const double* goal,
                         const double* u_prev, const int32_t* track, const double* OBS, long stride, const int32_t* nobs,
                         double* U, int32_t* status, double* pred_x, double* pred_u, int32_t* iters, double* kkt,
                         int* counter, void* workspace, size_t workspace_bytes, cudaStream_t s, int sm_count, int* count_only) {
  int rc = mpc_launch_m<MODEL>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status, pred_x, pred_u, iters,
                               kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
  if (rc != SCB_OK || !p.mpc_superellipsoid) return rc;
  int n2 = 0;
  int* counter2 = count_only ? nullptr : mpc_counter_slot(s);
  if (!counter2 && !count_only) return SCB_ERR_ALLOC;
  rc = mpc_launch_m<kMpcSeBaseId + MODEL>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status, pred_x, pred_u,
                                          iters, kkt, counter2, nullptr, 0, s, sm_count, count_only ? &n2 : nullptr);
  if (count_only) *count_only += n2;
  return rc;
}
   // (a per-model translation unit must not see references to the other models' launchers)
inline int mpc_launch(const scb_params& p, int N, int M, int H, const double* X, const double* Uref, const double* goal,
                      const double* u_prev, const int32_t* track, const double* OBS, long stride, const int32_t* nobs,
                      double* U, int32_t* status, double* pred_x, double* pred_u, int32_t* iters, double* kkt,
                      void* workspace, size_t workspace_bytes, cudaStream_t s, int sm_count, int* count_only = nullptr) {
  if (M > kMpcMaxObs || H > kMpcMaxH || H * p.nu > 64) return SCB_ERR_TOO_LARGE;
  int* counter = count_only ? nullptr : mpc_counter_slot(s);
  if (!counter && !count_only) return SCB_ERR_ALLOC;
  switch (p.model) {
    case SCB_SINGLE_INTEGRATOR_2D:
      return mpc_launch_se<SCB_SINGLE_INTEGRATOR_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                    pred_x, pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    case SCB_DYNAMIC_UNICYCLE_2D:
      return mpc_launch_se<SCB_DYNAMIC_UNICYCLE_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                   pred_x, pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    case SCB_KINEMATIC_BICYCLE_2D:
      return mpc_launch_m<SCB_KINEMATIC_BICYCLE_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                    pred_x, pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    case SCB_DOUBLE_INTEGRATOR_2D:
      return mpc_launch_se<SCB_DOUBLE_INTEGRATOR_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                    pred_x, pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    case SCB_QUAD_2D:
      return mpc_launch_m<SCB_QUAD_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status, pred_x,
                                       pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    case SCB_UNICYCLE_2D:
      return mpc_launch_m<SCB_UNICYCLE_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status, pred_x,
                                           pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    case SCB_KINEMATIC_BICYCLE_2D_C3BF:
      return mpc_launch_m<SCB_KINEMATIC_BICYCLE_2D_C3BF>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                         pred_x, pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    case SCB_KINEMATIC_BICYCLE_2D_DPCBF:
      return mpc_launch_m<SCB_KINEMATIC_BICYCLE_2D_DPCBF>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                          pred_x, pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    case SCB_QUAD_3D:
      return mpc_launch_m<SCB_QUAD_3D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status, pred_x,
                                       pred_u, iters, kkt, counter, workspace, workspace_bytes, s, sm_count, count_only);
    default:
      return SCB_ERR_UNSUPPORTED;     // Manipulator2D: no agent_barrier_dt in the reference
  }
}

#endif

}  // namespace scb
