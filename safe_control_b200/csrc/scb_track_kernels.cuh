// scb_track_kernels.cuh -- __global__ wrappers + launch sequence of the device-side control loop.
//
// One control_step() = track_pre_kernel -> [dyn_obs_kernel] -> the controller's solve kernel ->
// track_post_kernel on one stream.  A warp owns one agent in the pre/post kernels (the scene scan is
// strided over its lanes, the selection keys live in the warp's slice of shared memory); the solve
// kernels are the ones behind scb_*_solve and read the Uref/OBS/nobs buffers the pre kernel filled.
#pragma once

#include <cuda_runtime.h>

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
    track_pre_agent<MODEL, 32>(p, t, a, keys);
    __syncwarp();                                // keys are reused by the warp's next agent
  }
}

template <int MODEL>
__global__ void __launch_bounds__(kTrackBlock)
track_post_kernel(const __grid_constant__ scb_params p, const __grid_constant__ scb_track t) {
  constexpr int WPB = kTrackBlock / 32;
  for (long a = (long)blockIdx.x * WPB + (threadIdx.x >> 5); a < t.N; a += (long)gridDim.x * WPB)
    track_post_agent<MODEL, 32>(p, t, a);
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
