// scb_track_kernels.cuh -- __global__ wrappers + launch sequence of the device-side control loop.
//
// One control_step() = track_pre_kernel -> [dyn_obs_kernel] -> the controller's solve kernel ->
// track_post_kernel on one stream.  A warp owns one agent in the pre/post kernels (the scene scan is
// strided over its lanes, the selection keys live in the warp's slice of shared memory); the solve
// kernels are the ones behind scb_*_solve and read the Uref/OBS/nobs buffers the pre kernel filled.
#pragma once

#include <cuda_runtime.h>

#include "scb_qp.cuh"
#include "scb_track.cuh"

namespace scb {

constexpr int kTrackBlock = 128;                 // 4 agent-warps per CTA
constexpr int kTrackMaxScene = 1024;             // selection keys: 4 warps x 1024 x 8 B = 32 KB of shared memory

template <int MODEL>
__global__ void __launch_bounds__(kTrackBlock)
track_pre_kernel(const __grid_constant__ scb_params p, const __grid_constant__ scb_track t) {
  extern __shared__ double keys_smem[];
  double* keys = keys_smem + (size_t)(threadIdx.x >> 5) * t.K;
  constexpr int WPB = kTrackBlock / 32;
  for (long a = (long)blockIdx.x * WPB + (threadIdx.x >> 5); a < t.N; a += (long)gridDim.x * WPB) {
    track_pre_agent<MODEL, 32>(p, t, a, keys, t.SCENE);
    __syncwarp();                                // keys are reused by the warp's next agent
  }
}

template <int MODEL>
__global__ void __launch_bounds__(kTrackBlock)
track_post_kernel(const __grid_constant__ scb_params p, const __grid_constant__ scb_track t) {
  constexpr int WPB = kTrackBlock / 32;
  for (long a = (long)blockIdx.x * WPB + (threadIdx.x >> 5); a < t.N; a += (long)gridDim.x * WPB)
    track_post_agent<MODEL, 32>(p, t, a, t.SCENE);
}

// step_dyn_obs (dynamic_env/main.py:54-58)
__global__ void dyn_obs_kernel(double* scene, int K, double dt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < K) {
    scene[j * 7 + 0] += scene[j * 7 + 3] * dt;
    scene[j * 7 + 1] += scene[j * 7 + 4] * dt;
  }
}

template <int MODEL>
__global__ void __launch_bounds__(kTrackBlock)
select_kernel(const __grid_constant__ scb_params p, int N, int K, int M, const double* __restrict__ X,
              const double* __restrict__ yaw, const double* SCENE, long sstride, double* OBS, int32_t* nobs,
              int32_t* idx) {
  using ML = ModelLoop<MODEL>;
  extern __shared__ double keys_smem[];
  double* keys = keys_smem + (size_t)(threadIdx.x >> 5) * K;
  constexpr int WPB = kTrackBlock / 32;
  for (long a = (long)blockIdx.x * WPB + (threadIdx.x >> 5); a < N; a += (long)gridDim.x * WPB) {
    double x[ML::NX];
#pragma unroll
    for (int i = 0; i < ML::NX; ++i) x[i] = X[a * ML::NX + i];
    const double psi = ML::yaw_of(x, yaw ? yaw[a] : 0.0);
    const int no = select_agent<32>(K, M, SCENE + a * sstride, x[0], x[1], yaw ? yaw[a] : psi, ML::half_angle(), keys,
                                    OBS + (size_t)a * M * 7, idx ? idx + (size_t)a * M : nullptr);
    if ((threadIdx.x & 31) == 0) nobs[a] = no;
    __syncwarp();
  }
}

// n sequential step_dyn_obs updates (same arithmetic as n dyn_obs_kernel launches)
__global__ void dyn_obs_steps_kernel(double* scene, int K, double dt, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < K) {
    double x = scene[j * 7 + 0], y = scene[j * 7 + 1];
    const double vx = scene[j * 7 + 3], vy = scene[j * 7 + 4];
    for (int k = 0; k < n; ++k) { x += vx * dt; y += vy * dt; }
    scene[j * 7 + 0] = x; scene[j * 7 + 1] = y;
  }
}

// ---- fused closed loop (QP controllers) -----------------------------------------------------------------------
// run_all_steps in ONE launch: a warp owns one agent for all n_steps control steps; per step it runs the same three
// bodies the 3-launch path runs (track_pre_agent -> cbfqp_agent / odcbf_agent -> track_post_agent), separated by
// __syncwarp() instead of kernel boundaries.  Agents never interact, so CTAs need no grid-wide synchronisation:
// every CTA keeps its own copy of the scene in shared memory and, when the obstacles move, steps that copy itself
// (all copies evolve identically; the global SCENE is advanced afterwards by dyn_obs_steps_kernel).  The solve reads
// the rows this kernel wrote with plain loads (NC = false).  Results are bit-identical to the 3-launch path.
constexpr int kFusedMaxScene = 512;              // (7 + 4) * K doubles of shared memory per CTA <= 45 KB

template <int MODEL, int CTRL, int NW, int RPL>
__global__ void __launch_bounds__(kTrackBlock)
track_fused_kernel(const __grid_constant__ scb_params p, const __grid_constant__ scb_track t, int n_steps) {
  extern __shared__ double fused_smem[];
  constexpr int WPB = kTrackBlock / 32;
  constexpr int NX = ModelLoop<MODEL>::NX, NU = ModelLoop<MODEL>::NU;
  const int K = t.K, M = t.M;
  double* scene = fused_smem;
  double* keys = fused_smem + (size_t)K * 7 + (size_t)(threadIdx.x >> 5) * K;
  const long a = (long)blockIdx.x * WPB + (threadIdx.x >> 5);
  const bool live = a < t.N;
  for (int i = threadIdx.x; i < K * 7; i += kTrackBlock) scene[i] = t.SCENE[i];
  __syncthreads();
  const int words = (M + 2 * NU + 63) / 64;            // scb_active_words(M, nu)
  for (int k = 0; k < n_steps; ++k) {
    if (live) track_pre_agent<MODEL, 32>(p, t, a, keys, scene);
    if (t.dynamic_obs) {                         // step_dyn_obs after the selection (dynamic_env/main.py:152)
      __syncthreads();
      for (int j = threadIdx.x; j < K; j += kTrackBlock) {
        scene[j * 7 + 0] += scene[j * 7 + 3] * p.dt;
        scene[j * 7 + 1] += scene[j * 7 + 4] * p.dt;
      }
      __syncthreads();
    } else {
      __syncwarp();
    }
    if (live && !t.done[a]) {
      if (CTRL == SCB_CTRL_CBF_QP)
        cbfqp_agent<MODEL, 32, RPL, false>(p, M, t.nobs[a], t.X + a * NX, t.Uref + a * NU, t.OBS + (size_t)a * M * 7,
                                           t.U + a * NU, t.status + a, t.active ? t.active + a * words : nullptr, words);
      else
        odcbf_agent<MODEL, NW, 32, RPL, false>(p, M, t.nobs[a], t.X + a * NX, t.Uref + a * NU, t.OBS + (size_t)a * M * 7,
                                               t.U + a * NU, nullptr, nullptr, t.status + a, t.active ? t.active + a : nullptr);
    }
    __syncwarp();
    if (live) track_post_agent<MODEL, 32>(p, t, a, scene);
    __syncwarp();
  }
}

template <int MODEL, int CTRL, int NW>
inline bool launch_fused(const scb_params& p, const scb_track& t, int n_steps, cudaStream_t s) {
  const int rows = (CTRL == SCB_CTRL_CBF_QP) ? t.M + 4 : (t.M > 1 ? t.M : 1);
  const int rpl = rows <= 32 ? 1 : (rows <= 64 ? 2 : 0);
  if (rpl == 0 || t.K > kFusedMaxScene) return false;
  const int grid = (t.N + (kTrackBlock / 32) - 1) / (kTrackBlock / 32);
  const size_t smem = (size_t)(7 + kTrackBlock / 32) * (size_t)(t.K > 0 ? t.K : 1) * sizeof(double);
  if (rpl == 1) track_fused_kernel<MODEL, CTRL, NW, 1><<<grid, kTrackBlock, smem, s>>>(p, t, n_steps);
  else track_fused_kernel<MODEL, CTRL, NW, 2><<<grid, kTrackBlock, smem, s>>>(p, t, n_steps);
  if (t.dynamic_obs && t.K > 0) dyn_obs_steps_kernel<<<(t.K + 127) / 128, 128, 0, s>>>(t.SCENE, t.K, p.dt, n_steps);
  return true;
}

inline int track_grid(int N, int sm_count) {
  const long blocks = ((long)N + (kTrackBlock / 32) - 1) / (kTrackBlock / 32);
  const long cap = (long)sm_count * 16;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

template <int MODEL>
inline void launch_pre(const scb_params& p, const scb_track& t, cudaStream_t s, int sm_count) {
  const size_t smem = (size_t)(kTrackBlock / 32) * (size_t)(t.K > 0 ? t.K : 1) * sizeof(double);
  track_pre_kernel<MODEL><<<track_grid(t.N, sm_count), kTrackBlock, smem, s>>>(p, t);
}

template <int MODEL>
inline void launch_post(const scb_params& p, const scb_track& t, cudaStream_t s, int sm_count) {
  track_post_kernel<MODEL><<<track_grid(t.N, sm_count), kTrackBlock, 0, s>>>(p, t);
}

}  // namespace scb
