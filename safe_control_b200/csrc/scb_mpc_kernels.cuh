// scb_mpc_kernels.cuh -- launch side of the MPC-CBF path (placeholder until scb_mpc.cuh lands).
#pragma once
#include <cuda_runtime.h>
#include "scb_core.cuh"

namespace scb {
constexpr int kMpcMaxObs = 64;
constexpr int kMpcMaxH = 16;

inline int mpc_launch(const scb_params&, int, int, int, const double*, const double*, const double*, const double*,
                      const int32_t*, const double*, long, const int32_t*, double*, int32_t*, double*, double*,
                      int32_t*, double*, cudaStream_t, int) {
  return SCB_ERR_UNSUPPORTED;
}
}  // namespace scb
