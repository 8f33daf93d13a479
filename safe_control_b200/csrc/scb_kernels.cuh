// scb_kernels.cuh -- __global__ wrappers + launch dispatch for the QP paths.
//
// Grid mapping: a lane group (LANES lanes of one warp) owns one agent; groups grid-stride
// over the batch.  The host picks LANES from the batch size: 32 (warp per QP, lowest
// latency) while 32*N threads still fit the machine a few times over, 8 for large
// batches where throughput matters (4 QPs per warp, fewer redundant replicas of the
// replicated NV x NV algebra).  RPL (rows per lane) = ceil((M + 2 nu) / LANES).
#pragma once

#include <cuda_runtime.h>

#include "scb_qp.cuh"
#include "scb_tma.cuh"

namespace scb {

constexpr int kBlock = 128;
// __launch_bounds__ min CTAs/SM (register cap) for the QP kernels, measured on B200 (tools/sweep_qp.py, M = 16):
// warp-per-QP (latency regime, N = 1024): 4 -> 5.28 us, 6 -> 5.46 us;  4-8 lanes/QP (N = 1M): 4 -> 487/553 us,
// 6 -> 477/503 us.  Override both with -DSCB_QP_MINBLOCKS=n.
#ifdef SCB_QP_MINBLOCKS
#define SCB_QP_MINB(LANES) SCB_QP_MINBLOCKS
#else
#define SCB_QP_MINB(LANES) ((LANES) == 32 ? 4 : 6)
#endif

template <int MODEL, int LANES, int RPL, bool EAGER = false>
__global__ void __launch_bounds__(kBlock, SCB_QP_MINB(LANES))
cbfqp_kernel(const __grid_constant__ scb_params p, int N, int M, const double* __restrict__ X,
             const double* __restrict__ Uref, const double* __restrict__ OBS, long stride,
             const int32_t* __restrict__ nobs, double* __restrict__ U, int32_t* __restrict__ status,
             uint64_t* __restrict__ active, int words, const int32_t* __restrict__ skip) {
  constexpr int NX = ModelCT<MODEL>::NX, NU = ModelCT<MODEL>::NU;
  constexpr int GPB = kBlock / LANES;
  for (long a = (long)blockIdx.x * GPB + threadIdx.x / LANES; a < N; a += (long)gridDim.x * GPB) {
    if (skip && skip[a]) continue;     // closed loop: agents whose run has ended keep the outputs of their last step
    cbfqp_agent<MODEL, LANES, RPL, true, EAGER>(p, M, nobs ? nobs[a] : M, X + a * NX, Uref + a * NU, OBS + a * stride,
                                   U + a * NU, status + a, active ? active + a * words : nullptr, words);
  }
}

// ---- bulk-async staged variant (lane-group geometries, per-agent obstacle lists) ------------------------------------
// A warp owns G = 32 / LANES consecutive agents per pass.  Lane 0 asks the TMA engine for the G obstacle blocks
// (one cp.async.bulk of 56 M bytes per agent into the warp's private shared-memory slots, completion on the warp's
// mbarrier); while the copy is in flight the lanes fetch X / Uref / nobs (small, coalesced); then every lane reads its
// rows from shared memory and assembles them into registers.  As soon as the rows are in registers the slots are free:
// the copy of the warp's NEXT pass is issued before the QP solve, so the DRAM round trip overlaps the active-set
// iterations instead of preceding them.  Slot stride = 56 M bytes + padding such that the 16 lanes of a half-warp
// (16 / LANES agents x LANES rows, one 8-byte field each) hit 16 distinct bank pairs: stride = 8 LANES bytes mod 128.
// Preconditions (checked by the launcher, otherwise the LDG kernel runs): OBS 16-byte aligned, stride even, M even.
inline int tma_slot_doubles(int M, int lanes) {
  int bytes = 56 * M;
  const int want = (8 * lanes) & 127;                  // 32 (4 lanes), 64 (8 lanes), 0 (16 lanes)
  while ((bytes & 127) != want) bytes += 16;
  return bytes / 8;
}

// (min CTAs per SM, measured at 1 Mi agents, us per launch at 6 / 5 / 4: 4 lanes 301.9 / 291.2 / 289.6, 8 lanes 389.5 / 426.4 / 421.1;
//  profiles/r2_sweep_minblocks_v1.txt)
template <int MODEL, int LANES, int RPL>
__global__ void __launch_bounds__(kBlock, (LANES == 4) ? 4 : SCB_QP_MINB(LANES))
cbfqp_tma_kernel(const __grid_constant__ scb_params p, int N, int M, const double* __restrict__ X,
                 const double* __restrict__ Uref, const double* __restrict__ OBS, long stride,
                 const int32_t* __restrict__ nobs, double* __restrict__ U, int32_t* __restrict__ status,
                 uint64_t* __restrict__ active, int words, const int32_t* __restrict__ skip, int slot_doubles) {
  constexpr int NX = ModelCT<MODEL>::NX, NU = ModelCT<MODEL>::NU;
  constexpr int G = 32 / LANES, WPB = kBlock / 32;
  extern __shared__ __align__(128) double tma_stage[];
  __shared__ uint64_t bars[WPB];
  const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31, grp = lane32 / LANES;
  double* slots = tma_stage + (size_t)warp * G * slot_doubles;
  uint64_t* bar = &bars[warp];
  if (lane32 == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  __syncwarp();
  const uint32_t bytes = (uint32_t)(56 * M);
  const long npass = ((long)N + G - 1) / G, nwarps = (long)gridDim.x * WPB;
  auto issue = [&](long ps) {
    if (lane32 == 0 && ps < npass) {
      const long a0 = ps * G;
      const int cnt = (int)((N - a0) < G ? (N - a0) : G);
      mbar_arrive_expect_tx(bar, bytes * cnt);
      for (int g = 0; g < cnt; ++g) bulk_copy_g2s(slots + (size_t)g * slot_doubles, OBS + (a0 + g) * stride, bytes, bar);
    }
  };
  long pass = (long)blockIdx.x * WPB + warp;
  issue(pass);
  uint32_t parity = 0;
  for (; pass < npass; pass += nwarps) {
    const long a = pass * G + grp;
    const bool live = a < N && !(skip && skip[a]);
    // small per-agent inputs while the bulk copy is in flight
    int no = 0;
    double xs[NX], ur[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) ur[i] = 0.0;
    if (live) {
      no = nobs ? __ldg(nobs + a) : M;
#pragma unroll
      for (int i = 0; i < NX; ++i) xs[i] = __ldg(X + a * NX + i);
#pragma unroll
      for (int i = 0; i < NU; ++i) ur[i] = __ldg(Uref + a * NU + i);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    double r0[RPL], r1[RPL], rb[RPL];
    cbfqp_rows_staged<MODEL, LANES, RPL>(p, M, live ? no : 0, xs, slots + (size_t)grp * slot_doubles, r0, r1, rb);
    __syncwarp();                  // every lane has its rows in registers: the slots are free
    fence_proxy_async();
    issue(pass + nwarps);          // next pass streams in during the solve below
    if (!live) continue;
    if (no < 0) {                  // obs_list is None -> u_ref unclipped (cbf_qp.py:113-118)
      if ((lane32 & (LANES - 1)) == 0) {
#pragma unroll
        for (int i = 0; i < NU; ++i) U[a * NU + i] = ur[i];
        status[a] = SCB_OPTIMAL;
        if (active) for (int w = 0; w < words; ++w) active[a * words + w] = 0ull;
      }
      continue;
    }
    cbfqp_finish<MODEL, LANES, RPL>(p, M + 2 * NU, ur, r0, r1, rb, U + a * NU, status + a,
                                    active ? active + a * words : nullptr, words);
  }
}

// Manipulator2D: warp per arm (rows = M link-circle rows + 6 box rows, RPL = ceil(rows / 32))
template <int RPL>
__global__ void __launch_bounds__(kBlock)
manipqp_kernel(const __grid_constant__ scb_params p, int N, int M, const double* __restrict__ X,
               const double* __restrict__ Uref, const double* __restrict__ OBS, long stride,
               const int32_t* __restrict__ nobs, double* __restrict__ U, int32_t* __restrict__ status,
               uint64_t* __restrict__ active, int words) {
  constexpr int GPB = kBlock / 32;
  for (long a = (long)blockIdx.x * GPB + threadIdx.x / 32; a < N; a += (long)gridDim.x * GPB)
    manipqp_agent<32, RPL>(p, M, nobs ? nobs[a] : M, X + a * 3, Uref + a * 3, OBS + a * stride, U + a * 3, status + a,
                           active ? active + a * words : nullptr, words);
}

__global__ void __launch_bounds__(256)
manip_rows_kernel(const __grid_constant__ scb_params p, int N, int M, const double* __restrict__ X,
                  const double* __restrict__ OBS, long stride, const int32_t* __restrict__ nobs,
                  double* __restrict__ A, double* __restrict__ b) {
  const long total = (long)N * M;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long a = t / M;
    const int r = (int)(t - a * M);
    int no = nobs ? nobs[a] : M;
    no = no < 0 ? 0 : (no > M ? M : no);
    double q[3];
    for (int i = 0; i < 3; ++i) q[i] = __ldg(X + a * 3 + i);
    ManipArm arm;
    manip_prep(q, arm);
    double av[3], bv;
    manip_row<true>(p, arm, OBS + a * stride, no, r, av, bv);
    for (int i = 0; i < 3; ++i) A[t * 3 + i] = av[i];
    b[t] = bv;
  }
}

template <int MODEL>
__global__ void __launch_bounds__(256)
cbfqp_rows_kernel(const __grid_constant__ scb_params p, int N, int M, const double* __restrict__ X,
                  const double* __restrict__ OBS, long stride, const int32_t* __restrict__ nobs,
                  double* __restrict__ A, double* __restrict__ b) {
  constexpr int NX = ModelCT<MODEL>::NX, NU = ModelCT<MODEL>::NU;
  const long total = (long)N * M;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long a = t / M;
    const int r = (int)(t - a * M);
    int no = nobs ? nobs[a] : M;
    no = no < 0 ? 0 : (no > M ? M : no);
    double xs[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = __ldg(X + a * NX + i);
    AgentCT g;
    ModelCT<MODEL>::prep(p, xs, g);
    double av[NU], bv;
    cbfqp_row<MODEL>(p, g, OBS + a * stride, M, no, r, av, bv);
#pragma unroll
    for (int i = 0; i < NU; ++i) A[t * NU + i] = av[i];
    b[t] = bv;
  }
}

template <int MODEL, int NW, int LANES, int RPL>
__global__ void __launch_bounds__(kBlock)
odcbf_kernel(const __grid_constant__ scb_params p, int N, int M, const double* __restrict__ X,
             const double* __restrict__ Uref, const double* __restrict__ OBS, long stride,
             const int32_t* __restrict__ nobs, double* __restrict__ U, double* __restrict__ omega,
             int32_t* __restrict__ sel, int32_t* __restrict__ status, uint64_t* __restrict__ active,
             const int32_t* __restrict__ skip) {
  constexpr int GPB = kBlock / LANES;
  for (long a = (long)blockIdx.x * GPB + threadIdx.x / LANES; a < N; a += (long)gridDim.x * GPB) {
    if (skip && skip[a]) continue;
    odcbf_agent<MODEL, NW, LANES, RPL>(p, M, nobs ? nobs[a] : M, X + a * ModelCT<MODEL>::NX, Uref + a * 2, OBS + a * stride,
                                       U + a * 2, omega ? omega + a * 2 : nullptr, sel ? sel + a : nullptr,
                                       status + a, active ? active + a : nullptr);
  }
}

// bulk-async staged variant of the optimal-decay kernel: the nearest-obstacle scan AND the row of the chosen obstacle
// read shared memory, so the second, dependent DRAM round trip of the LDG kernel (scan x, y -> pick -> load the row)
// disappears; at BASELINE config 4 (8192 agents x 32 obstacles = 15 MB) every warp's copy is in flight at once.
template <int MODEL, int NW, int LANES, int RPL>
__global__ void __launch_bounds__(kBlock, 4)       // <= 128 registers: config 4's 512 CTAs are resident in one wave (4 per SM)
odcbf_tma_kernel(const __grid_constant__ scb_params p, int N, int M, const double* __restrict__ X,
                 const double* __restrict__ Uref, const double* __restrict__ OBS, long stride,
                 const int32_t* __restrict__ nobs, double* __restrict__ U, double* __restrict__ omega,
                 int32_t* __restrict__ sel, int32_t* __restrict__ status, uint64_t* __restrict__ active,
                 const int32_t* __restrict__ skip, int slot_doubles) {
  constexpr int NX = ModelCT<MODEL>::NX;
  constexpr int G = 32 / LANES, WPB = kBlock / 32;
  extern __shared__ __align__(128) double tma_stage[];
  __shared__ uint64_t bars[WPB];
  const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31, grp = lane32 / LANES;
  double* slots = tma_stage + (size_t)warp * G * slot_doubles;
  uint64_t* bar = &bars[warp];
  if (lane32 == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  __syncwarp();
  const uint32_t bytes = (uint32_t)(56 * M);
  const long npass = ((long)N + G - 1) / G, nwarps = (long)gridDim.x * WPB;
  uint32_t parity = 0;
  for (long pass = (long)blockIdx.x * WPB + warp; pass < npass; pass += nwarps) {
    if (lane32 == 0) {
      const long a0 = pass * G;
      const int cnt = (int)((N - a0) < G ? (N - a0) : G);
      mbar_arrive_expect_tx(bar, bytes * cnt);
      for (int g = 0; g < cnt; ++g) bulk_copy_g2s(slots + (size_t)g * slot_doubles, OBS + (a0 + g) * stride, bytes, bar);
    }
    const long a = pass * G + grp;
    const bool live = a < N && !(skip && skip[a]);
    int no = 0;
    double xs[NX], ur[2] = {0.0, 0.0};
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = 0.0;
    if (live) {
      no = nobs ? __ldg(nobs + a) : M;
#pragma unroll
      for (int i = 0; i < NX; ++i) xs[i] = __ldg(X + a * NX + i);
      ur[0] = __ldg(Uref + a * 2); ur[1] = __ldg(Uref + a * 2 + 1);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    if (live)
      odcbf_agent<MODEL, NW, LANES, RPL, false>(p, M, no, xs, ur, slots + (size_t)grp * slot_doubles, U + a * 2,
                                                omega ? omega + a * 2 : nullptr, sel ? sel + a : nullptr, status + a,
                                                active ? active + a : nullptr);
    __syncwarp();                  // all lanes are done with the slots before the next pass overwrites them
    fence_proxy_async();
  }
}

struct LaunchGeom {
  int lanes, rpl, grid;
};

// total rows -> (LANES, RPL); returns false if beyond the compiled instantiations.
// Compiled (LANES, RPL): (32,1) (32,2) (32,4)  (8,3) (8,4) (8,8)  (4,5) (4,8); RPL slots are fully unrolled, so the
// smallest RPL >= ceil(rows / LANES) is used (rows = 20 for M = 16: (8,3) and (4,5) waste no slot).
// `force_lanes` (0 = auto) is a tuning override (env SCB_QP_LANES), honoured only when an instantiation exists.
inline int pick_rpl(int lanes, int rows) {
  const int need = (rows + lanes - 1) / lanes;
  if (lanes == 32) return need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : 0;
  if (lanes == 8) return need <= 3 ? 3 : need <= 4 ? 4 : need <= 8 ? 8 : 0;
  if (lanes == 4) return need <= 5 ? 5 : need <= 8 ? 8 : 0;
  return 0;
}

inline bool pick_geom(long N, int rows, int sm_count, LaunchGeom& g, int force_lanes = 0) {
  // measured on B200 (tools/sweep_qp.py, M = 16, us per launch, lanes 32 / 8 / 4):
  //   N = 1024: 4.42 / 5.14 / 7.28     N = 8192: 10.8 / 6.47 / 8.27     N = 1M: 1186 / 420 / 390
  // -> warp per QP up to ~16 warps per SM, 8 lanes per QP for mid-size batches, 4 lanes per QP beyond ~32k agents.
  const bool small = N <= (long)sm_count * 16;
  int lanes = (small || rows > 64) ? 32 : ((rows <= 32 && N > 32768) ? 4 : 8);
  if ((force_lanes == 32 || force_lanes == 8 || force_lanes == 4) && pick_rpl(force_lanes, rows) > 0) lanes = force_lanes;
  g.lanes = lanes;
  g.rpl = pick_rpl(lanes, rows);
  if (g.rpl == 0) return false;
  const long groups_per_block = kBlock / g.lanes;
  long blocks = (N + groups_per_block - 1) / groups_per_block;
  const long cap = (long)sm_count * 16;                     // persistent-ish: <= 16 CTAs of 128 thr per SM
  g.grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
  return true;
}

}  // namespace scb
