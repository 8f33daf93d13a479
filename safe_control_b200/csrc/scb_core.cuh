// scb_core.cuh -- lane-group primitives shared by every kernel.
//
// A "group" is LANES consecutive lanes of a warp that cooperate on ONE agent's
// problem: constraint rows are strided across the lanes, the tiny dense algebra
// (n <= 4) is replicated in every lane, and the only cross-lane traffic is
// xor-shuffle reductions + a broadcast of the winning row.  LANES is a launch-
// time tuning knob: 32 (one warp per QP) minimises latency for small batches,
// 4-8 maximises throughput for large ones, and LANES == 1 is plain sequential
// code -- which is also what lets tests/_hostsim compile the *same* source with
// g++ and check the kernel math on a CPU-only box (test aid, never shipped).
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/scb.h"

#if defined(__CUDACC__)
#define SCB_HD __host__ __device__ __forceinline__
#define SCB_D __device__ __forceinline__
#define SCB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SCB_HD_NOINLINE __attribute__((noinline))
#define SCB_HD inline
#define SCB_D inline
#endif

namespace scb {

constexpr double kInf = 1e300;

template <int LANES>
struct Grp {
  static_assert(LANES == 1 || LANES == 2 || LANES == 4 || LANES == 8 || LANES == 16 || LANES == 32, "LANES");

  static SCB_HD int lane() {
#if defined(__CUDA_ARCH__)
    return (LANES == 1) ? 0 : (int)(threadIdx.x & (LANES - 1));
#else
    return 0;
#endif
  }

  // shuffle mask of THIS group only: groups of one warp may diverge (different
  // active-set iteration counts), so a full-warp mask would deadlock.
  static SCB_HD unsigned gmask() {
#if defined(__CUDA_ARCH__)
    return (LANES == 32) ? 0xffffffffu
                         : (((1u << (LANES & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(LANES - 1)));
#else
    return 0u;
#endif
  }

  static SCB_HD double sum(double v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(gmask(), v, o, 32);
#endif
    return v;
  }

  static SCB_HD double vmin(double v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES >> 1; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(gmask(), v, o, 32));
#endif
    return v;
  }

  // argmin with deterministic tie-break on the smaller index; every lane gets (v, idx)
  static SCB_HD void argmin(double& v, int& idx) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES >> 1; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(gmask(), v, o, 32);
      int oi = __shfl_xor_sync(gmask(), idx, o, 32);
      if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
#endif
  }

  // value held by group-lane `src` -> all lanes of the group
  static SCB_HD double bcast(double v, int src) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) v = __shfl_sync(gmask(), v, src, LANES);
#endif
    (void)src;
    return v;
  }

  static SCB_HD uint32_t bcast_u32(uint32_t v, int src) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) v = __shfl_sync(gmask(), v, src, LANES);
#endif
    (void)src;
    return v;
  }

  static SCB_HD uint32_t or_reduce(uint32_t v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES >> 1; o > 0; o >>= 1) v |= __shfl_xor_sync(gmask(), v, o, 32);
#endif
    return v;
  }
};

SCB_HD void sincos_pair(double th, double& s, double& c) {
#if defined(__CUDA_ARCH__)
  sincos(th, &s, &c);
#else
  s = sin(th);
  c = cos(th);
#endif
}

// out-of-line versions for the (instruction-cache bound) MPC kernel: one copy of the slow paths
#if defined(__CUDACC__)
#define SCB_PHASE __host__ __device__ __noinline__
#else
#define SCB_PHASE inline
#endif
#ifndef SCB_MPC_OUTLINE_MATH
#define SCB_MPC_OUTLINE_MATH 0
#endif
#if SCB_MPC_OUTLINE_MATH
static SCB_PHASE void sincos_call(double a, double* s, double* c) { sincos_pair(a, *s, *c); }
static SCB_PHASE double log_call(double a) { return log(a); }
#else
SCB_HD void sincos_call(double a, double* s, double* c) { sincos_pair(a, *s, *c); }
SCB_HD double log_call(double a) { return log(a); }
#endif

SCB_HD double rsqrt_pos(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

SCB_HD double ld(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// NC = true: read-only (non-coherent) load of a kernel INPUT; NC = false: plain load, for buffers the same kernel
// wrote earlier (the fused closed-loop kernel solves on rows / inputs it produced itself a few lines above).
template <bool NC>
SCB_HD double ldx(const double* p) {
#if defined(__CUDA_ARCH__)
  if (NC) return __ldg(p);
#endif
  return *p;
}

}  // namespace scb
