// scb_mpc_kernels.cuh -- __global__ wrapper + launch for the MPC-CBF path (scb_mpc.cuh).
//
// One lane group (LANES = 32: a full warp) per agent; each group owns a private workspace of
// MpcLayout::total doubles in dynamic shared memory (17 KB at H = 8, M = 16; 40 KB at H = 10, M = 64).  The grid is
// persistent: one-warp CTAs, as many per SM as registers and shared memory allow, for launches of a few waves; CTAs of
// up to 8 agent-warps, one per SM, for launches of four waves or more (mpc_launch_m in scb_mpc_impl.cuh).  Everything
// the interior-point loop touches after the initial obstacle load stays on chip; HBM sees the inputs once and U/status once.
#pragma once
// (the kernel and its per-model launcher live in scb_mpc_impl.cuh; each model is its own translation unit,
// scb_mpc_inst.cu compiled with -DSCB_MPC_INST=<model id>, so the library builds in parallel)

#include <cuda_runtime.h>

#include "scb_core.cuh"

namespace scb {

constexpr int kMpcMaxObs = 64;
// horizon limit: the stage-wise (Riccati) solver is O(H) in workspace and time, so the real limit is the shared-memory
// budget of one agent (mpc_launch_m returns SCB_ERR_TOO_LARGE when a single workspace exceeds it); VTOL2D uses H = 30
constexpr int kMpcMaxH = 32;
// all caller arrays of one MPC launch (device pointers; see include/scb.h scb_mpccbf_solve_ws)
struct MpcIO {
  const double* X; const double* Uref; const double* goal; const double* u_prev; const int32_t* track;
  const double* OBS; long stride; const int32_t* nobs;
  double* U; int32_t* status; double* pred_x; double* pred_u; int32_t* iters; double* kkt;
  uint64_t* active;      // [N, active_words] or NULL: bit k*M + j = CBF row (stage k, obstacle slot j), then the simple bounds
  double* omega;         // optimal-decay MPC only: [N, H, 2] or NULL
};
// defined in scb_mpc_impl.cuh, explicitly instantiated per model (scb_mpc_inst.cu)
template <int MODEL>
int mpc_launch_m(const scb_params& p, int N, int M, int H, const MpcIO& io, int* counter, void* workspace,
                 size_t workspace_bytes, cudaStream_t s, int sm_count, int* count_only);

// Caller-owned scratch of one call (scb_mpccbf_workspace_bytes): work counters [kMpcWsHead ints: one per launch of the
// call, zeroed by a stream-ordered memset], histogram [1024], bin per agent [N], schedule [N], all int32.  The library
// owns no device memory and no global state: without a workspace the kernel strides the agents statically.
constexpr int kMpcWsHead = 16;
inline size_t mpc_workspace_bytes(long N) { return (size_t)(kMpcWsHead + 1024 + 2 * (N > 0 ? N : 0)) * sizeof(int32_t); }

#ifndef SCB_MPC_NO_DISPATCH
constexpr int kMpcSeBaseId = 100;       // == kMpcSeBase (scb_mpc.cuh): general-row variants of SI / DU / DI for superellipsoid rows
// fast path + (when p.mpc_superellipsoid) the general-row launch for the agents that have a superellipsoid row
template <int MODEL>
inline int mpc_launch_se(const scb_params& p, int N, int M, int H, const MpcIO& io, int* counter, void* workspace,
                         size_t workspace_bytes, cudaStream_t s, int sm_count, int* count_only) {
  int rc = mpc_launch_m<MODEL>(p, N, M, H, io, counter, workspace, workspace_bytes, s, sm_count, count_only);
  if (rc != SCB_OK || !p.mpc_superellipsoid) return rc;
  int n2 = 0;
  int* counter2 = counter ? counter + 1 : nullptr;
  rc = mpc_launch_m<kMpcSeBaseId + MODEL>(p, N, M, H, io, counter2, nullptr, 0, s, sm_count, count_only ? &n2 : nullptr);
  if (count_only) *count_only += n2;
  return rc;
}
   // (a per-model translation unit must not see references to the other models' launchers)
inline int mpc_launch(const scb_params& p, int N, int M, int H, const MpcIO& io, void* workspace, size_t workspace_bytes,
                      cudaStream_t s, int sm_count, int* count_only = nullptr) {
  if (M > kMpcMaxObs || H > kMpcMaxH) return SCB_ERR_TOO_LARGE;
  int* counter = nullptr;                   // dynamic agent scheduling needs the caller's workspace
  if (!count_only && workspace && workspace_bytes >= mpc_workspace_bytes(N)) {
    counter = (int*)workspace;
    if (cudaMemsetAsync(counter, 0, kMpcWsHead * sizeof(int), s) != cudaSuccess) return SCB_ERR_CUDA;
  }
#define SCB_GO(FN, MODEL) case MODEL: return FN<MODEL>(p, N, M, H, io, counter, workspace, workspace_bytes, s, sm_count, count_only);
  if (p.od_mpc) {                       // optimal-decay MPC-CBF: omega1, omega2 as extra stage inputs (scb_mpc.cuh MpcModelOD)
    constexpr int kOd = 200;            // == kMpcOdBase
    switch (p.model) {
      case SCB_DYNAMIC_UNICYCLE_2D: return mpc_launch_m<kOd + SCB_DYNAMIC_UNICYCLE_2D>(p, N, M, H, io, counter, workspace, workspace_bytes, s, sm_count, count_only);
      case SCB_KINEMATIC_BICYCLE_2D: return mpc_launch_m<kOd + SCB_KINEMATIC_BICYCLE_2D>(p, N, M, H, io, counter, workspace, workspace_bytes, s, sm_count, count_only);
      case SCB_QUAD_2D: return mpc_launch_m<kOd + SCB_QUAD_2D>(p, N, M, H, io, counter, workspace, workspace_bytes, s, sm_count, count_only);
      case SCB_VTOL_2D: return mpc_launch_m<kOd + SCB_VTOL_2D>(p, N, M, H, io, counter, workspace, workspace_bytes, s, sm_count, count_only);
      default: return SCB_ERR_UNSUPPORTED;     // optimal_decay_mpc_cbf.py:19-20 (Quad3D: see include/scb.h)
    }
  }
  switch (p.model) {
    SCB_GO(mpc_launch_se, SCB_SINGLE_INTEGRATOR_2D)
    SCB_GO(mpc_launch_se, SCB_DYNAMIC_UNICYCLE_2D)
    SCB_GO(mpc_launch_m, SCB_KINEMATIC_BICYCLE_2D)
    SCB_GO(mpc_launch_se, SCB_DOUBLE_INTEGRATOR_2D)
    SCB_GO(mpc_launch_m, SCB_QUAD_2D)
    SCB_GO(mpc_launch_m, SCB_UNICYCLE_2D)
    SCB_GO(mpc_launch_m, SCB_KINEMATIC_BICYCLE_2D_C3BF)
    SCB_GO(mpc_launch_m, SCB_KINEMATIC_BICYCLE_2D_DPCBF)
    SCB_GO(mpc_launch_m, SCB_QUAD_3D)
    SCB_GO(mpc_launch_m, SCB_VTOL_2D)
    default:
      return SCB_ERR_UNSUPPORTED;     // Manipulator2D: no agent_barrier_dt in the reference
  }
#undef SCB_GO
}

#endif

}  // namespace scb
