// scb_mpc_impl.cuh -- __global__ wrapper + per-model launcher of the MPC-CBF path (body: scb_mpc.cuh).
//
// One lane group per agent; each group owns a private workspace of MpcLayout::total doubles in dynamic shared
// memory, so a CTA carries as many agents as fit in the SM's 227 KB and the grid is persistent (one CTA per SM,
// agents pulled from a work counter).  Everything the interior-point loop touches after the initial obstacle
// load stays on chip; HBM sees the inputs once and U/status once.
#pragma once

#include <stdlib.h>

#include "scb_mpc.cuh"
#include "scb_mpc_kernels.cuh"

namespace scb {

#ifndef SCB_MPC_LANES
#define SCB_MPC_LANES 32             // lanes per agent (16: two agents per warp, for models whose Riccati stage fits: NV + 1 <= 16)
#endif
template <int MODEL>
constexpr int mpc_lanes() {
  using Mod = MpcModel<MODEL>;
  return (Mod::NX + 2 * Mod::NU + 1 <= SCB_MPC_LANES) ? SCB_MPC_LANES : 32;
}

#ifndef SCB_MPC_MAXTHREADS
#define SCB_MPC_MAXTHREADS 256       // packed CTAs: at most 8 agent-warps (round 1, cfg3: 256 threads 8.3 ms, 384 with spills 8.9 ms, 512 9.5 ms)
#endif

template <int MODEL, int LANES>
__global__ void __launch_bounds__(SCB_MPC_MAXTHREADS, 1) mpc_kernel(const __grid_constant__ scb_params p, int N, int M, int H, int ws_doubles, int gpb,
                           const __grid_constant__ MpcIO io, int active_words, int* __restrict__ next_agent,
                           const int32_t* __restrict__ order, const __grid_constant__ MpcLayout lay) {
  const double* __restrict__ X = io.X; const double* __restrict__ Uref = io.Uref; const double* __restrict__ goal = io.goal;
  const double* __restrict__ u_prev = io.u_prev; const int32_t* __restrict__ track = io.track;
  const double* __restrict__ OBS = io.OBS; const long stride = io.stride; const int32_t* __restrict__ nobs = io.nobs;
  double* __restrict__ U = io.U; int32_t* __restrict__ status = io.status; double* __restrict__ pred_x = io.pred_x;
  double* __restrict__ pred_u = io.pred_u; int32_t* __restrict__ iters = io.iters; double* __restrict__ kkt = io.kkt;
  uint64_t* __restrict__ active = io.active;
  extern __shared__ double smem[];
  using Mod = MpcModel<MODEL>;
  constexpr int NX = Mod::NX, NU = Mod::NU;
  constexpr bool kIsSe = MODEL >= kMpcSeBase && MODEL < kMpcOdBase;       // (200 + id: the optimal-decay variants)
  constexpr bool kHasSeVariant = MODEL == SCB_SINGLE_INTEGRATOR_2D || MODEL == SCB_DYNAMIC_UNICYCLE_2D || MODEL == SCB_DOUBLE_INTEGRATOR_2D;
  const int grp = threadIdx.x / LANES;
  if (grp >= gpb) return;             // padding lanes of the last warp (gpb * LANES is rounded up to whole warps)
  double* ws = smem + (size_t)grp * ws_doubles;
  // Work distribution: the first wave is static, afterwards a group that finishes early pulls the next agent
  // from a global counter (iteration counts vary 10..60 per agent; static striding left ~30 % of the wave idle).
  // With a schedule (`order`, hardest-looking agents first, see mpc_key_kernel) slot q runs agent order[q].
  const long first_dynamic = (long)gridDim.x * gpb;
  for (long q = (long)blockIdx.x * gpb + grp; q < N;) {
    const long a = order ? (long)order[q] : q;
    // agents with a superellipsoid row belong to the general-row variant of the model (second launch), all others
    // to the fast path: each kernel skips what is not its kind
    bool mine = !(track && track[a] < 0), se_unsupported = false;     // track < 0: skip (outputs untouched)
    if (mine) if constexpr (kIsSe || kHasSeVariant) {
      const int no = nobs ? min(max(nobs[a], 0), M) : M;
      unsigned se = 0;
      for (int j = threadIdx.x & (LANES - 1); j < no; j += LANES) se |= (__ldg(OBS + a * stride + j * 7 + 6) >= 0.5) ? 1u : 0u;
      se = Grp<LANES>::or_reduce(se);
      if (kIsSe) mine = (se != 0) && !(track && track[a] == 0);
      else if (se != 0 && !(track && track[a] == 0)) { mine = false; se_unsupported = (p.mpc_superellipsoid == 0); }
    }
    if (!mine) {
      if (se_unsupported && (threadIdx.x & (LANES - 1)) == 0) {      // flagged rows without mpc_superellipsoid: refuse loudly
        for (int i = 0; i < NU; ++i) U[a * NU + i] = fmin(fmax(u_prev[a * NU + i], p.u_lb[i]), p.u_ub[i]);
        status[a] = SCB_NUMERICAL;
        if (iters) iters[a] = 0;
        if (kkt) kkt[a] = kInf;
        if (active) for (int q2 = 0; q2 < active_words; ++q2) active[a * active_words + q2] = 0ull;
      }
    } else if (track && track[a] == 0) {
      // state_machine != 'track': return u_ref untouched, no solve (mpc_cbf.py:379-381)
      if ((threadIdx.x & (LANES - 1)) == 0) {
        for (int i = 0; i < NU; ++i) U[a * NU + i] = Uref[a * NU + i];
        status[a] = SCB_OPTIMAL;
        if (iters) iters[a] = 0;
        if (kkt) kkt[a] = 0.0;
        if (active) for (int q2 = 0; q2 < active_words; ++q2) active[a * active_words + q2] = 0ull;
      }
    } else {
    mpc_agent<MODEL, LANES>(p, lay, nobs ? nobs[a] : M, X + a * NX, goal + a * Mod::NGOAL, u_prev + a * NU, OBS + a * stride, ws,
                            U + a * NU, status + a, pred_x ? pred_x + a * (H + 1) * NX : nullptr,
                            pred_u ? pred_u + a * H * NU : nullptr, iters ? iters + a : nullptr,
                            kkt ? kkt + a : nullptr, active ? active + a * active_words : nullptr);
    }
    if (!next_agent) { q += first_dynamic; continue; }     // no workspace: static stride
    int nxt = 0;
    if ((threadIdx.x & (LANES - 1)) == 0) nxt = atomicAdd(next_agent, 1);
    nxt = __shfl_sync(Grp<LANES>::gmask(), nxt, 0, LANES);
    q = first_dynamic + nxt;
  }
}

// ---- schedule: hardest-looking agents first ---------------------------------------------------------------
// Iteration counts are long-tailed (cfg3: mean 17, 99 % <= 22, max ~58) and an agent is a serial job, so a straggler
// that STARTS late sets the kernel's duration: simulated on cfg3's own iteration counts, index order costs 106
// iteration-slots per group against 64 ideal and 73 for longest-first.  The stragglers are the agents whose CBF rows
// are violated or nearly so at the cold start, so the key is the kernel's own constraint function at stage 0,
//   key = min_j cbf_j(x_0, u_prev),
// quantised monotonically into kMpcBins bins; a counting sort over the bins gives the schedule (ascending key).
// Only the ORDER in which agents are started changes -- every agent's result is independent of it.
constexpr int kMpcBins = 1024;

template <int MODEL>
__global__ void mpc_key_kernel(const __grid_constant__ scb_params p, int N, const double* __restrict__ X,
                               const double* __restrict__ u_prev, const int32_t* __restrict__ track,
                               const double* __restrict__ OBS, long stride, const int32_t* __restrict__ nobs, int M,
                               int32_t* __restrict__ bin_of, int32_t* __restrict__ hist) {
  using Mod = MpcModel<MODEL>;
  constexpr int NX = Mod::NX, NU = Mod::NU, NY = Mod::NY;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  int bin = kMpcBins - 1;                                  // not solved (state machine != track): last
  if (!track || track[a] > 0) {
    double w0, w1, w2, Wsum;
    if (Mod::REL == 2) {
      const double g1 = p.alpha1 + p.alpha2, g2 = p.alpha1 * p.alpha2;
      w2 = 1.0; w1 = g1 - 2.0; w0 = 1.0 - g1 + g2; Wsum = g2;
    } else {
      w2 = 0.0; w1 = 1.0; w0 = p.alpha - 1.0; Wsum = p.alpha;
    }
    double y[NY], P1, Q1, P2, Q2;
#pragma unroll
    for (int i = 0; i < NX; ++i) y[i] = __ldg(X + (long)a * NX + i);
#pragma unroll
    for (int i = 0; i < NU; ++i) y[NX + i] = __ldg(u_prev + (long)a * NU + i);
    if constexpr (Mod::LINEAR) {
      // (the linear model's barrier points need its RK4 matrices; the plain barrier value at x_0 ranks well enough)
      P1 = y[0]; Q1 = y[1]; P2 = y[0]; Q2 = y[1];
      w0 = Wsum; w1 = 0.0; w2 = 0.0;
    } else {
      double F[NX];
      TrigCompute trig;
      Mod::stage(p, nullptr, y, F, P1, Q1, P2, Q2, trig);
    }
    const int no = nobs ? min(max(nobs[a], 0), M) : M;
    const double* ob = OBS + (long)a * stride;
    double key = 1e6;
    for (int j = 0; j < no; ++j) {
      const double ox = __ldg(ob + j * 7), oy = __ldg(ob + j * 7 + 1), d = __ldg(ob + j * 7 + 2) + p.radius;
      double v = -Wsum * Mod::beta() * d * d;
      double dx = y[0] - ox, dy = y[1] - oy;
      v = fma(w0, dx * dx + dy * dy, v);
      dx = P1 - ox; dy = Q1 - oy;
      v = fma(w1, dx * dx + dy * dy, v);
      dx = P2 - ox; dy = Q2 - oy;
      v = fma(w2, dx * dx + dy * dy, v);
      key = fmin(key, v);
    }
    if (!(key == key)) key = -1e6;                         // NaN inputs: start them first, they exit at once
    const double q = key / (fabs(key) + 1.0);              // monotone map to (-1, 1)
    bin = (int)((q + 1.0) * 0.5 * (kMpcBins - 2));
    bin = min(max(bin, 0), kMpcBins - 2);
  }
  bin_of[a] = bin;
  atomicAdd(hist + bin, 1);
}

// one CTA: exclusive scan of the histogram, then scatter (order within a bin is arbitrary)
static __global__ void __launch_bounds__(kMpcBins) mpc_order_kernel(int N, const int32_t* __restrict__ bin_of,
                                                            const int32_t* __restrict__ hist, int32_t* __restrict__ order) {
  __shared__ int cursor[kMpcBins];
  __shared__ int wsum[kMpcBins / 32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int h = hist[t];
  int v = h;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
  if (lane == 31) wsum[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int s = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += u; }
    wsum[lane] = s;
  }
  __syncthreads();
  cursor[t] = v - h + (wid > 0 ? wsum[wid - 1] : 0);
  __syncthreads();
  for (int a = t; a < N; a += kMpcBins) order[atomicAdd(&cursor[bin_of[a]], 1)] = a;
}

template <int MODEL>
int mpc_launch_m(const scb_params& p, int N, int M, int H, const MpcIO& io, int* counter, void* workspace,
                 size_t workspace_bytes, cudaStream_t s, int sm_count, int* count_only) {
  using Mod = MpcModel<MODEL>;
  const MpcLayout L = mpc_layout<Mod, false>(H, M);
  const size_t per = (size_t)L.total * sizeof(double);
  const size_t budget = 220 * 1024;
  constexpr int kLanes = mpc_lanes<MODEL>();
  constexpr int kMpcMaxGroups = SCB_MPC_MAXTHREADS / kLanes;
  int slots = (int)(budget / per);                       // agent-warps one SM can hold (shared memory)
  if (slots < 1) return SCB_ERR_TOO_LARGE;
  if (slots > kMpcMaxGroups) slots = kMpcMaxGroups;
  // CTA shape.  One agent-warp per CTA (batches of a few waves): the warps of a CTA are independent, so small CTAs lose nothing, and an SM
  // slot is released the moment ITS agent queue runs dry instead of when the slowest of `slots` warps is done -- which is
  // what lets the CTAs of another launch (the next model group of a mixed batch, on another stream) move in during the
  // tail.  SCB_MPC_GPB=n packs n agent-warps per CTA (round 1's shape: n = slots, one CTA per SM).
  // Launches of >= 4 waves keep round 1's packed CTAs (measured 3 % faster there: config 5 on one GPU 201.8 vs 208.7 ms).
  int gpb = ((long)N >= 4L * sm_count * slots) ? slots : 1;
  if (const char* e = getenv("SCB_MPC_GPB")) { const int v = atoi(e); if (v >= 1) gpb = v < slots ? v : slots; }
  auto kern = mpc_kernel<MODEL, kLanes>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per * (gpb > 1 ? gpb : 1)));
  if (e != cudaSuccess) {
    if (!count_only) return SCB_ERR_TOO_LARGE;
    cudaGetLastError();                                  // launch-count query on a box without a device: the estimate above stands
  } else if (gpb == 1 && !getenv("SCB_MPC_SLOTS8")) {
    // one-warp CTAs are not bound to 256 threads per SM: as many as registers and shared memory allow (the occupancy
    // calculator knows the kernel's register count; 144 registers -> 14 warps, 220 KB / 16.7 KB -> 13 at the cfg3 shape)
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kLanes, per) == cudaSuccess && occ > 0) {
      const int by_smem = (int)(budget / per);
      slots = occ < by_smem ? occ : by_smem;
      if (slots > 16) slots = 16;
    }
  }
  const int cta_per_sm = slots / gpb > 0 ? slots / gpb : 1;
  const size_t smem = per * gpb;
  long blocks = ((long)N + gpb - 1) / gpb;
  if (blocks > (long)sm_count * cta_per_sm) blocks = (long)sm_count * cta_per_sm;
  // schedule (needs the caller's workspace; without one, or when every agent starts in the first wave, index order)
  const int32_t* order = nullptr;
  const bool scheduled = workspace && workspace_bytes >= mpc_workspace_bytes(N) && (long)N > blocks * gpb;
  if (count_only) { *count_only = scheduled ? 3 : 1; return SCB_OK; }       // (nothing launched; works without a device)
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SCB_ERR_TOO_LARGE;
  if (scheduled) {
    int32_t* hist = (int32_t*)workspace + kMpcWsHead;
    int32_t* bin_of = hist + kMpcBins;
    int32_t* ord = bin_of + N;
    if (cudaMemsetAsync(hist, 0, kMpcBins * sizeof(int32_t), s) != cudaSuccess) return SCB_ERR_CUDA;
    mpc_key_kernel<MODEL><<<(N + 127) / 128, 128, 0, s>>>(p, N, io.X, io.u_prev, io.track, io.OBS, io.stride, io.nobs, M, bin_of, hist);
    mpc_order_kernel<<<1, kMpcBins, 0, s>>>(N, bin_of, hist, ord);
    order = ord;
  }
  kern<<<(int)blocks, ((gpb * kLanes + 31) / 32) * 32, smem, s>>>(p, N, M, H, L.total, gpb, io, mpc_active_words<Mod>(H, M), counter, order, L);
  return SCB_OK;
}


#define SCB_MPC_INSTANTIATE(MODEL)                                                                                   \
  template int mpc_launch_m<MODEL>(const scb_params&, int, int, int, const MpcIO&, int*, void*, size_t, cudaStream_t, int, int*);

}  // namespace scb
