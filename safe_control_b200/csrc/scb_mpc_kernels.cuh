// scb_mpc_kernels.cuh -- __global__ wrapper + launch for the MPC-CBF path (scb_mpc.cuh).
//
// One lane group (LANES = 32: a full warp) per agent; each group owns a private workspace of
// MpcLayout::total doubles in dynamic shared memory (~35 KB at H = 8, M = 16), so a CTA carries as
// many agents as fit in the SM's 227 KB and the grid is persistent (one CTA per SM, agents
// grid-strided).  Everything the interior-point loop touches after the initial obstacle load
// stays on chip; HBM sees the inputs once and U/status once.
#pragma once

#include <cuda_runtime.h>

#include "scb_mpc.cuh"

namespace scb {

constexpr int kMpcMaxObs = 64;
constexpr int kMpcMaxH = 16;
constexpr int kMpcLanes = 32;

#ifndef SCB_MPC_MAXTHREADS
#define SCB_MPC_MAXTHREADS 256       // measured (cfg3): 256 threads (255 regs, 8 agent-warps) 8.3 ms, 384 (168 regs + spills) 8.9 ms, 512 9.5 ms
#endif
constexpr int kMpcMaxGroups = SCB_MPC_MAXTHREADS / kMpcLanes;

template <int MODEL, int LANES>
__global__ void __launch_bounds__(SCB_MPC_MAXTHREADS, 1) mpc_kernel(const __grid_constant__ scb_params p, int N, int M, int H, int ws_doubles,
                           const double* __restrict__ X, const double* __restrict__ Uref,
                           const double* __restrict__ goal, const double* __restrict__ u_prev,
                           const int32_t* __restrict__ track, const double* __restrict__ OBS, long stride,
                           const int32_t* __restrict__ nobs, double* __restrict__ U, int32_t* __restrict__ status,
                           double* __restrict__ pred_x, double* __restrict__ pred_u, int32_t* __restrict__ iters,
                           double* __restrict__ kkt, int* __restrict__ next_agent) {
  extern __shared__ double smem[];
  using Mod = MpcModel<MODEL>;
  constexpr int NX = Mod::NX, NU = Mod::NU;
  const int gpb = blockDim.x / LANES;
  const int grp = threadIdx.x / LANES;
  double* ws = smem + (size_t)grp * ws_doubles;
  // Work distribution: the first wave is static, afterwards a group that finishes early pulls the next agent
  // from a global counter (iteration counts vary 10..60 per agent; static striding left ~30 % of the wave idle).
  const long first_dynamic = (long)gridDim.x * gpb;
  for (long a = (long)blockIdx.x * gpb + grp; a < N;) {
    if (track && track[a] == 0) {
      // state_machine != 'track': return u_ref untouched, no solve (mpc_cbf.py:379-381)
      if ((threadIdx.x & (LANES - 1)) == 0) {
        for (int i = 0; i < NU; ++i) U[a * NU + i] = Uref[a * NU + i];
        status[a] = SCB_OPTIMAL;
        if (iters) iters[a] = 0;
        if (kkt) kkt[a] = 0.0;
      }
    } else {
    mpc_agent<MODEL, LANES>(p, H, M, nobs ? nobs[a] : M, X + a * NX, goal + a * Mod::NGOAL, u_prev + a * NU, OBS + a * stride, ws,
                            U + a * NU, status + a, pred_x ? pred_x + a * (H + 1) * NX : nullptr,
                            pred_u ? pred_u + a * H * NU : nullptr, iters ? iters + a : nullptr,
                            kkt ? kkt + a : nullptr);
    }
    int nxt = 0;
    if ((threadIdx.x & (LANES - 1)) == 0) nxt = atomicAdd(next_agent, 1);
    nxt = __shfl_sync(Grp<LANES>::gmask(), nxt, 0, LANES);
    a = first_dynamic + nxt;
  }
}

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

template <int MODEL>
inline int mpc_launch_m(const scb_params& p, int N, int M, int H, const double* X, const double* Uref,
                        const double* goal, const double* u_prev, const int32_t* track, const double* OBS, long stride,
                        const int32_t* nobs, double* U, int32_t* status, double* pred_x, double* pred_u,
                        int32_t* iters, double* kkt, cudaStream_t s, int sm_count) {
  using Mod = MpcModel<MODEL>;
  const MpcLayout L = mpc_layout<Mod::NX, Mod::NU, Mod::VBOUND, Mod::LINEAR, Mod::AUX>(H, M);
  const size_t per = (size_t)L.total * sizeof(double);
  const size_t budget = 220 * 1024;
  int gpb = (int)(budget / per);
  if (gpb < 1) return SCB_ERR_TOO_LARGE;
  if (gpb > kMpcMaxGroups) gpb = kMpcMaxGroups;
  const long need = ((long)N + gpb - 1) / gpb;
  if (need < sm_count) {                       // small batch: spread over all SMs first
    gpb = (int)(((long)N + sm_count - 1) / sm_count);
    if (gpb < 1) gpb = 1;
  }
  const size_t smem = per * gpb;
  auto kern = mpc_kernel<MODEL, kMpcLanes>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SCB_ERR_TOO_LARGE;
  long blocks = ((long)N + gpb - 1) / gpb;
  if (blocks > sm_count) blocks = sm_count;
  int* counter = mpc_counter_slot(s);
  if (!counter) return SCB_ERR_ALLOC;
  kern<<<(int)blocks, gpb * kMpcLanes, smem, s>>>(p, N, M, H, L.total, X, Uref, goal, u_prev, track, OBS, stride, nobs, U,
                                                  status, pred_x, pred_u, iters, kkt, counter);
  return SCB_OK;
}

inline int mpc_launch(const scb_params& p, int N, int M, int H, const double* X, const double* Uref, const double* goal,
                      const double* u_prev, const int32_t* track, const double* OBS, long stride, const int32_t* nobs,
                      double* U, int32_t* status, double* pred_x, double* pred_u, int32_t* iters, double* kkt,
                      cudaStream_t s, int sm_count) {
  if (M > kMpcMaxObs || H > kMpcMaxH || H * p.nu > 64) return SCB_ERR_TOO_LARGE;
  switch (p.model) {
    case SCB_SINGLE_INTEGRATOR_2D:
      return mpc_launch_m<SCB_SINGLE_INTEGRATOR_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                    pred_x, pred_u, iters, kkt, s, sm_count);
    case SCB_DYNAMIC_UNICYCLE_2D:
      return mpc_launch_m<SCB_DYNAMIC_UNICYCLE_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                   pred_x, pred_u, iters, kkt, s, sm_count);
    case SCB_KINEMATIC_BICYCLE_2D:
      return mpc_launch_m<SCB_KINEMATIC_BICYCLE_2D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status,
                                                    pred_x, pred_u, iters, kkt, s, sm_count);
    case SCB_QUAD_3D:
      return mpc_launch_m<SCB_QUAD_3D>(p, N, M, H, X, Uref, goal, u_prev, track, OBS, stride, nobs, U, status, pred_x,
                                       pred_u, iters, kkt, s, sm_count);
    default:
      return SCB_ERR_UNSUPPORTED;     // C3BF (collision-cone) barriers in MPC: not yet
  }
}

}  // namespace scb
