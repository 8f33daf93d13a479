// Backup-CBF QP kernel: one lane group per agent (scb_backup.cuh), groups packed into 128-thread CTAs.
#pragma once
#include "scb_backup.cuh"
#include "scb_shield.cuh"

namespace scb {

constexpr int kBkBlock = 128;

SCB_HD int bk_agent_doubles(int n_backup) {            // scratch + rows, padded to a multiple of 4 doubles + 4 (bank spread)
  return ((kBkScratch + 3 * n_backup + 3) & ~3) + 4;
}

template <int LANES, int RPL>
__global__ void __launch_bounds__(kBkBlock)
backupcbf_kernel(const scb_backup_params p, int N, int K, const double* __restrict__ X, const double* __restrict__ Uref,
                 const double* __restrict__ MOV, long mov_stride, double* __restrict__ U, int32_t* __restrict__ status,
                 int32_t* __restrict__ intervene, double* __restrict__ h_min, double* __restrict__ phi,
                 double* __restrict__ rows_out, uint64_t* __restrict__ active, int words) {
  extern __shared__ double scb_bk_smem_[];
  constexpr int kGroups = kBkBlock / LANES;
  const int g = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const long agent = (long)blockIdx.x * kGroups + g;
  if (agent >= N) return;                              // (whole groups leave together)
  const int nb = p.n_backup;
  double* scr = scb_bk_smem_ + (size_t)g * bk_agent_doubles(nb);
  double* rows = scr + kBkScratch;
  BackupOut o;
  backup_agent<LANES, RPL>(p, X + agent * 4, Uref + agent * 2, MOV ? MOV + agent * mov_stride : nullptr, MOV ? K : 0, scr,
                           rows, phi ? phi + agent * (long)nb * 4 : nullptr, o);
  if (rows_out)
    for (int k = lane; k < 3 * nb; k += LANES) rows_out[agent * (long)nb * 3 + k] = rows[k];
  if (lane == 0) {
    U[agent * 2] = o.u0; U[agent * 2 + 1] = o.u1;
    status[agent] = o.status;
    if (intervene) intervene[agent] = o.intervene;
    if (h_min) h_min[agent] = o.h_min;
    if (active) {
      for (int w = 0; w < words; ++w) {
        uint64_t bits = 0ull;
        if (o.w0 >= 0 && o.lam0 > 0.0 && (o.w0 >> 6) == w) bits |= 1ull << (o.w0 & 63);
        if (o.w1 >= 0 && o.lam1 > 0.0 && (o.w1 >> 6) == w) bits |= 1ull << (o.w1 & 63);
        active[agent * words + w] = bits;
      }
    }
  }
}

// ---- the same work as two launches: rollout + rows (few registers, 352 B of shared memory per agent -> the SM holds 8x
// more agents than the fused kernel, and the rollout is a chain of dependent fp64 sqrt / div latencies), then the QP ----
template <int LANES>
__global__ void __launch_bounds__(kBkBlock)
backup_rollout_kernel(const scb_backup_params p, int N, int K, const double* __restrict__ X, const double* __restrict__ MOV,
                      long mov_stride, double* __restrict__ h_min, double* __restrict__ phi, double* __restrict__ rows) {
  // agents per CTA: kBkBlock / LANES, or 6 per warp for LANES = 5 (lanes 30, 31 of every warp idle)
  constexpr int kPerWarp = (LANES == 5) ? 6 : 32 / LANES;
  constexpr int kGroups = (kBkBlock / 32) * kPerWarp;
  __shared__ double scr_all[kGroups * kBkScratch];
  const int wl = threadIdx.x & 31;
  if (LANES == 5 && wl >= 30) return;
  const int g = (threadIdx.x >> 5) * kPerWarp + wl / LANES;
  long agent = (long)blockIdx.x * kGroups + g;
  bool valid = agent < N;
  if (!valid) {
    if (LANES != 5) return;                            // (whole power-of-two groups leave together)
    agent = N - 1;                                     // the 30 lanes of a warp synchronise as one group: keep them all, write nothing
  }
  const int nb = p.n_backup;
  const double h = backup_rollout<LANES>(p, X + agent * 4, MOV ? MOV + agent * mov_stride : nullptr, MOV ? K : 0,
                                         scr_all + g * kBkScratch, valid ? rows + agent * (long)nb * 3 : nullptr,
                                         (phi && valid) ? phi + agent * (long)nb * 4 : nullptr);
  if (valid && wl % LANES == 0) h_min[agent] = h;
}

template <int LANES, int RPL>
__global__ void __launch_bounds__(kBkBlock)
backup_qp_kernel(const scb_backup_params p, int N, const double* __restrict__ X, const double* __restrict__ Uref,
                 const double* __restrict__ rows, const double* __restrict__ h_min, double* __restrict__ U,
                 int32_t* __restrict__ status, int32_t* __restrict__ intervene, uint64_t* __restrict__ active, int words) {
  constexpr int kGroups = kBkBlock / LANES;
  const int g = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const long agent = (long)blockIdx.x * kGroups + g;
  if (agent >= N) return;
  BackupOut o;
  backup_qp<LANES, RPL>(p, X + agent * 4, Uref + agent * 2, rows + agent * (long)p.n_backup * 3, h_min[agent], o);
  if (lane == 0) {
    U[agent * 2] = o.u0; U[agent * 2 + 1] = o.u1;
    status[agent] = o.status;
    if (intervene) intervene[agent] = o.intervene;
    if (active) {
      for (int w = 0; w < words; ++w) {
        uint64_t bits = 0ull;
        if (o.w0 >= 0 && o.lam0 > 0.0 && (o.w0 >> 6) == w) bits |= 1ull << (o.w0 & 63);
        if (o.w1 >= 0 && o.lam1 > 0.0 && (o.w1 >> 6) == w) bits |= 1ull << (o.w1 & 63);
        active[agent * words + w] = bits;
      }
    }
  }
}

// ---- gatekeeper / MPS: one control step of N agents, one lane group per agent (scb_shield.cuh) ----
// PHASE 0: the whole step in one launch.  PHASE 1 / 2: the two-launch search (see shield_agent); `work` = [count, agent ids...].
template <int LANES, int PHASE>
__global__ void __launch_bounds__(kBkBlock)
shield_step_kernel(const scb_shield_params sp, const scb_shield_state st, int N, int K, const double* __restrict__ X,
                   const double* __restrict__ NOMX, const double* __restrict__ NOMU, const int32_t* __restrict__ nom_len,
                   const double* __restrict__ MOV, long mov_stride, const double* __restrict__ STAT, double* __restrict__ U,
                   int32_t* __restrict__ using_backup, int32_t* __restrict__ work) {
  constexpr int kGroups = kBkBlock / LANES;
  const int g = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  long a = (long)blockIdx.x * kGroups + g;
  if (PHASE == 2) {
    if (a >= work[0]) return;                          // (work[0] = number of pending agents, written by the phase-1 launch)
    a = work[1 + a];
  } else if (a >= N) {
    return;
  }
  const int T = sp.nom_cap, Nb = sp.scene.n_backup;
  ShieldIO io;
  io.x = X + a * 4;
  io.nomx = NOMX + a * (long)(T + 1) * 4;
  io.nomu = NOMU + a * (long)T * 2;
  int nl = nom_len ? nom_len[a] : T + 1;
  io.nom_len = nl < 0 ? 0 : (nl > T + 1 ? T + 1 : nl);
  io.mov = MOV ? MOV + a * mov_stride : nullptr; io.K = MOV ? K : 0;
  io.stat = STAT ? STAT + a * 5 : nullptr;
  const int cb = st.cbuf[a] & 1;
  const long lu = (long)(T + Nb) * 2, lx = (long)(T + Nb + 1) * 4;
  io.cu = st.CU + (a * 2 + cb) * lu; io.cu_spare = st.CU + (a * 2 + (cb ^ 1)) * lu;
  io.cx = st.CX ? st.CX + (a * 2 + cb) * lx : nullptr; io.cx_spare = st.CX ? st.CX + (a * 2 + (cb ^ 1)) * lx : nullptr;
  int clen = st.clen[a], cidx = st.cidx[a], nsteps = st.nsteps[a], ub = 0;
  double ne = st.next_event[a], u[2];
  bool flip = false;
  const bool done = shield_agent<LANES, PHASE>(sp, io, clen, cidx, nsteps, ne, u, ub, flip);
  if (lane == 0) {
    st.clen[a] = clen; st.cidx[a] = cidx; st.nsteps[a] = nsteps; st.next_event[a] = ne;
    if (!done) {                                       // (phase 1 only) queue the agent for the lane-group search
      work[1 + atomicAdd(work, 1)] = (int32_t)a;
      return;
    }
    if (flip) st.cbuf[a] = cb ^ 1;
    U[a * 2] = u[0]; U[a * 2 + 1] = u[1];
    if (using_backup) using_backup[a] = ub;
  }
}

}  // namespace scb
