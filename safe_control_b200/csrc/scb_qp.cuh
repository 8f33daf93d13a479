// scb_qp.cuh -- per-agent bodies of the two QP controllers (device + host-sim).
//
//   cbfqp_agent  : CBFQP.solve_control_problem            position_control/cbf_qp.py:108-199
//   odcbf_agent  : OptimalDecayCBFQP.solve_control_problem position_control/optimal_decay_cbf_qp.py:132-159
//
// Both are "fused": the constraint rows are assembled straight into the lane registers
// the solver reads (they never touch HBM), then solved exactly (scb_gi.cuh).
#pragma once

#include "scb_gi.cuh"
#include "scb_models.cuh"

namespace scb {

// Row r of the CBF-QP, r in [0, M + 2 NU):
//   r <  M      : CBF row of obstacle slot r (vacuous 0 >= 0 when r >= nobs, cbf_qp.py:110-111)
//   r = M + 2i  : u_i <= ub_i      r = M + 2i + 1 : u_i >= lb_i          (cbf_qp.py:54-73)
// `pre` != nullptr: the obstacle row was loaded by the caller before `nobs` was known (see cbfqp_agent).
template <int MODEL, bool NC = true>
SCB_HD void cbfqp_row(const scb_params& p, const AgentCT& g, const double* obs, int M, int nobs, int r,
                      double* a, double& b, const double* pre = nullptr) {
  constexpr int NU = ModelCT<MODEL>::NU;
#pragma unroll
  for (int i = 0; i < NU; ++i) a[i] = 0.0;
  b = 0.0;
  if (r < M) {
    if (r < nobs) {
      double o[7];
#pragma unroll
      for (int q = 0; q < 7; ++q) o[q] = pre ? pre[q] : ldx<NC>(obs + (size_t)r * 7 + q);
      RowOut ro;
      ModelCT<MODEL>::row(p, g, o, ro);
#pragma unroll
      for (int i = 0; i < NU; ++i) a[i] = ro.a[i];
      b = ro.b;
    }
  } else if (r < M + 2 * NU) {
    const int q = r - M, i = q >> 1;
    if (q & 1) {
#pragma unroll
      for (int t = 0; t < NU; ++t) if (t == i) { a[t] = 1.0; b = -p.u_lb[t]; }
    } else {
#pragma unroll
      for (int t = 0; t < NU; ++t) if (t == i) { a[t] = -1.0; b = p.u_ub[t]; }
    }
  }
}

template <int MODEL, int LANES, int RPL>
SCB_HD void cbfqp_finish(const scb_params& p, int mrows, const double* ur, const double (&r0)[RPL], const double (&r1)[RPL],
                         const double (&rb)[RPL], double* U, int32_t* status, uint64_t* active, int words);

template <int MODEL, int LANES, int RPL, bool NC = true, bool EAGER = false>
SCB_HD void cbfqp_agent(const scb_params& p, int M, int nobs, const double* x, const double* uref,
                        const double* obs, double* U, int32_t* status, uint64_t* active, int words) {
  using Mod = ModelCT<MODEL>;
  using G = Grp<LANES>;
  constexpr int NU = Mod::NU;
  const int lane = G::lane();

  // EAGER (warp per agent, inputs in device memory): a small batch is a chain of latencies (cfg2: 5.2 us for 1 MB), so
  // the obstacle-row loads are issued BEFORE anything looks at `nobs` (slot r < M always exists in OBS) and their DRAM
  // round trip overlaps the one that fetches nobs instead of following it: 5.18 -> 4.68 us per 1024-agent step.
  // Lane-group launches (large batches) and the zero-copy host path (rows would cross PCIe: e2e 18.0 -> 14.5 M/s when
  // tried) stay lazy -- rows beyond nobs are never read.  (Staging the block through shared memory with coalesced
  // loads was measured too: 5.66 us, the extra store / barrier / load round trip costs more than the sectors it saves.)
  constexpr bool kEager = EAGER && (LANES == 32) && NC && (RPL <= 2);
  double pre[kEager ? RPL : 1][7];
  if (kEager) {
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      const int r = j * LANES + lane;
#pragma unroll
      for (int q = 0; q < 7; ++q) pre[j][q] = (r < M) ? ldx<NC>(obs + (size_t)r * 7 + q) : 0.0;
    }
  }

  double ur[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i) ur[i] = ldx<NC>(uref + i);

  if (nobs < 0) {                    // obs_list is None -> u_ref unclipped (cbf_qp.py:113-118)
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NU; ++i) U[i] = ur[i];
      *status = SCB_OPTIMAL;
      if (active) for (int w = 0; w < words; ++w) active[w] = 0ull;
    }
    return;
  }
  if (nobs > M) nobs = M;            // "Stop if we exceed allocated constraints" (cbf_qp.py:128-129)

  double xs[Mod::NX];
#pragma unroll
  for (int i = 0; i < Mod::NX; ++i) xs[i] = ldx<NC>(x + i);
  AgentCT g;
  Mod::prep(p, xs, g);

  // rows, normalised to unit normals (slack = signed distance in u-space), one per lane slot
  static_assert(NU == 2, "the fused CBF-QP kernel covers the 2-input models");
  double r0[RPL], r1[RPL], rb[RPL];
  const int mrows = M + 2 * NU;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    double a[NU], b;
    cbfqp_row<MODEL, NC>(p, g, obs, M, nobs, j * LANES + lane, a, b, kEager ? pre[j] : nullptr);
    const double n2 = a[0] * a[0] + a[1] * a[1];
    const double inv = (n2 > 0.0) ? rsqrt_pos(n2) : 1.0;
    r0[j] = a[0] * inv; r1[j] = a[1] * inv; rb[j] = b * inv;
  }
  cbfqp_finish<MODEL, LANES, RPL>(p, mrows, ur, r0, r1, rb, U, status, active, words);
}

// second half of cbfqp_agent: exact 2-variable QP over the normalised rows held in registers + outputs
template <int MODEL, int LANES, int RPL>
SCB_HD void cbfqp_finish(const scb_params& p, int mrows, const double* ur, const double (&r0)[RPL], const double (&r1)[RPL],
                         const double (&rb)[RPL], double* U, int32_t* status, uint64_t* active, int words) {
  using G = Grp<LANES>;
  const int lane = G::lane();
  Qp2Out q;
  gi_solve2<LANES, RPL>(2.0, ur[0], ur[1], r0, r1, rb, mrows, 8 * mrows + 16, q);

  if (lane == 0) {
    double v0 = q.x0, v1 = q.x1;
    if (q.status != SCB_OPTIMAL) {     // never leave the box / never NaN on failure
      v0 = fmin(fmax(v0, p.u_lb[0]), p.u_ub[0]);
      v1 = fmin(fmax(v1, p.u_lb[1]), p.u_ub[1]);
    }
    U[0] = v0; U[1] = v1;
    *status = q.status;
    if (active) {
      for (int w = 0; w < words; ++w) {
        uint64_t bits = 0ull;
        if (q.w0 >= 0 && q.lam0 > 0.0 && (q.w0 >> 6) == w) bits |= 1ull << (q.w0 & 63);
        if (q.w1 >= 0 && q.lam1 > 0.0 && (q.w1 >> 6) == w) bits |= 1ull << (q.w1 & 63);
        active[w] = bits;
      }
    }
  }
}

// The same rows as cbfqp_agent's first half, from an obstacle block that is already ON CHIP (shared memory filled by
// a bulk-async copy, scb_kernels.cuh cbfqp_tma_kernel): plain loads, every lane of the warp takes the same path
// (dead agents / nobs < 0 produce vacuous rows without touching `obs`), so the caller may warp-synchronise after it.
template <int MODEL, int LANES, int RPL>
SCB_HD void cbfqp_rows_staged(const scb_params& p, int M, int nobs, const double* xs, const double* obs, double (&r0)[RPL],
                              double (&r1)[RPL], double (&rb)[RPL]) {
  using Mod = ModelCT<MODEL>;
  constexpr int NU = Mod::NU;
  static_assert(NU == 2, "the fused CBF-QP kernel covers the 2-input models");
  const int lane = Grp<LANES>::lane();
  AgentCT g;
  Mod::prep(p, xs, g);
  if (nobs > M) nobs = M;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    double a[NU], b;
    cbfqp_row<MODEL, false>(p, g, obs, M, nobs, j * LANES + lane, a, b, nullptr);
    const double n2 = a[0] * a[0] + a[1] * a[1];
    const double inv = (n2 > 0.0) ? rsqrt_pos(n2) : 1.0;
    r0[j] = a[0] * inv; r1[j] = a[1] * inv; rb[j] = b * inv;
  }
}

// ----------------------------------------------------------------------------------------
// Manipulator2D CBF-QP (cbf_qp.py:96-105 problem, 131-149 rows): 3 joint velocities, |u_i| <= w_max, and one
// row PER LINK CIRCLE of every obstacle (robots/manipulator2D.py:110-198): the arm is covered by circles every
// 10/60 m along each link -- 9 + 9 + 7 = 25 per obstacle (ceil(70/10 = 7.000000000000001) = 8 steps on link 2,
// ceil(5.000000000000001) = 6 on link 3) -- and the reference stacks them obstacle by obstacle until its `num_obs`
// rows are full.  Here M is that row budget AND the number of obstacle slots of OBS; row r belongs to obstacle
// r / 25, circle r % 25.  Active bits: bit r = CBF row r, bit M + 2i / M + 2i + 1 = joint i at +/- w_max.
struct ManipArm {
  double jx[4], jy[4];        // joint positions P_0 (base) .. P_3 (end effector)
};
SCB_HD void manip_prep(const double* q, ManipArm& arm) {
  const double L[3] = {80.0 / 60.0, 70.0 / 60.0, 50.0 / 60.0};                   // :15-16
  arm.jx[0] = 0.0; arm.jy[0] = 0.0;
  double ang = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    ang += q[i];
    double s, c; sincos_pair(ang, s, c);
    arm.jx[i + 1] = arm.jx[i] + L[i] * c;
    arm.jy[i + 1] = arm.jy[i] + L[i] * s;
  }
}
// row r < M of the arm's QP (zero row when its obstacle is beyond nobs)
template <bool NC = true>
SCB_HD void manip_row(const scb_params& p, const ManipArm& arm, const double* obs, int nobs, int r, double* a, double& b) {
  a[0] = 0.0; a[1] = 0.0; a[2] = 0.0; b = 0.0;
  const int j = r / 25, c = r - j * 25;
  if (j >= nobs) return;
  const int link = (c < 9) ? 0 : (c < 18 ? 1 : 2);
  const int k = c - (link == 0 ? 0 : (link == 1 ? 9 : 18));
  const double t = (double)k / (double)(link == 2 ? 6 : 8);                      // j / num_steps (:133-134)
  double sx = 0.0, sy = 0.0, ex = 0.0, ey = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) if (i == link) { sx = arm.jx[i]; sy = arm.jy[i]; ex = arm.jx[i + 1]; ey = arm.jy[i + 1]; }
  const double cx = sx + t * (ex - sx), cy = sy + t * (ey - sy);                 // p_start + t * [dx, dy]
  const double ox = ldx<NC>(obs + (size_t)j * 7), oy = ldx<NC>(obs + (size_t)j * 7 + 1), orad = ldx<NC>(obs + (size_t)j * 7 + 2);
  const double dmin = p.radius + orad;
  const double dx = cx - ox, dy = cy - oy;
  const double h = (dx * dx + dy * dy) - 1.3 * (dmin * dmin);                    // beta = 1.3 (:163)
  // dh/dq_k = 2 [dx, dy] . [-(cy - P_k.y), cx - P_k.x]  for joints k <= link (get_points_jacobian :140-160)
#pragma unroll
  for (int kk = 0; kk < 3; ++kk)
    if (kk <= link) a[kk] = 2.0 * dx * (-(cy - arm.jy[kk])) + 2.0 * dy * (cx - arm.jx[kk]);
  b = (p.cbf_mode == 1) ? h / p.dt : p.alpha * h;                                // f = 0, g = I (:138-146)
}

template <int LANES, int RPL, bool NC = true>
SCB_HD void manipqp_agent(const scb_params& p, int M, int nobs, const double* x, const double* uref,
                          const double* obs, double* U, int32_t* status, uint64_t* active, int words) {
  using G = Grp<LANES>;
  constexpr int NU = 3;
  const int lane = G::lane();
  double ur[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i) ur[i] = ldx<NC>(uref + i);
  if (nobs < 0) {                    // obs_list is None -> u_ref unclipped (cbf_qp.py:113-118)
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NU; ++i) U[i] = ur[i];
      *status = SCB_OPTIMAL;
      if (active) for (int w = 0; w < words; ++w) active[w] = 0ull;
    }
    return;
  }
  if (nobs > M) nobs = M;
  double q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) q[i] = ldx<NC>(x + i);
  ManipArm arm;
  manip_prep(q, arm);
  double ra[RPL][NU], rb[RPL];
  const int mrows = M + 2 * NU;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    const int r = j * LANES + lane;
    ra[j][0] = 0.0; ra[j][1] = 0.0; ra[j][2] = 0.0; rb[j] = 0.0;
    if (r < M) {
      manip_row<NC>(p, arm, obs, nobs, r, ra[j], rb[j]);
    } else if (r < mrows) {
      const int qq = r - M, i = qq >> 1;
#pragma unroll
      for (int t = 0; t < NU; ++t)
        if (t == i) { ra[j][t] = (qq & 1) ? 1.0 : -1.0; rb[j] = (qq & 1) ? -p.u_lb[t] : p.u_ub[t]; }
    }
  }
  const double hd[NU] = {2.0, 2.0, 2.0};
  QpOut<NU> out;
  gi_solve<NU, LANES, RPL>(hd, ur, ra, rb, mrows, 8 * mrows + 16, out);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      double v = out.x[i];
      if (out.status != SCB_OPTIMAL) v = fmin(fmax(v, p.u_lb[i]), p.u_ub[i]);
      U[i] = v;
    }
    *status = out.status;
    if (active) {
      for (int w = 0; w < words; ++w) {
        uint64_t bits = 0ull;
#pragma unroll
        for (int a = 0; a < NU; ++a)
          if (a < out.wk && out.lam[a] > 0.0 && (out.widx[a] >> 6) == w) bits |= 1ull << (out.widx[a] & 63);
        active[w] = bits;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// Optimal-decay CBF-QP.  Variables z = [u (2), omega1 (, omega2)], ONE CBF row
// (optimal_decay_cbf_qp.py:61) built from the nearest valid obstacle of the agent's list
// (tracking.py:585-586), 4 box rows.  NW = number of omega variables (1: C3BF, 2: DU/KB).
// Active bits: bit 0 = CBF row, bit 1+2i = u_i upper, bit 2+2i = u_i lower.
template <int MODEL, int NW, int LANES, int RPL, bool NC = true>
SCB_HD void odcbf_agent(const scb_params& p, int M, int nobs, const double* x, const double* uref,
                        const double* obs, double* U, double* omega, int32_t* sel, int32_t* status,
                        uint64_t* active) {
  using Mod = ModelCT<MODEL>;
  using G = Grp<LANES>;
  constexpr int NU = 2, NV = NU + NW;
  const int lane = G::lane();
  if (nobs > M || nobs < 0) nobs = (nobs < 0) ? 0 : M;

  double xs[Mod::NX];
#pragma unroll
  for (int i = 0; i < Mod::NX; ++i) xs[i] = ldx<NC>(x + i);
  AgentCT g;
  Mod::prep(p, xs, g);

  // nearest obstacle by centre distance (first row of tracking.py:get_nearest_unpassed_obs' sort)
  double bestd = kInf;
  int bi = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    const int r = j * LANES + lane;
    if (r < nobs) {
      const double dx = ldx<NC>(obs + (size_t)r * 7) - g.px, dy = ldx<NC>(obs + (size_t)r * 7 + 1) - g.py;
      const double d2 = dx * dx + dy * dy;
      if (d2 < bestd) { bestd = d2; bi = r; }
    }
  }
  G::argmin(bestd, bi);
  const bool has = (bi != 0x7fffffff);

  // the single CBF row, replicated in every lane
  double ra[5][NV], rb[5];
#pragma unroll
  for (int r = 0; r < 5; ++r) {
#pragma unroll
    for (int i = 0; i < NV; ++i) ra[r][i] = 0.0;
    rb[r] = 0.0;
  }
  if (has) {
    double o[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) o[q] = ldx<NC>(obs + (size_t)bi * 7 + q);
    if (MODEL == SCB_KINEMATIC_BICYCLE_2D_C3BF) {                       // optimal_decay_cbf_qp.py:139-143
      double h, dh[4];
      ModelCT<SCB_KINEMATIC_BICYCLE_2D_C3BF>::barrier(p, g, o, h, dh);
      ra[0][0] = dh[3];
      ra[0][1] = dh[0] * (-g.v * g.s) + dh[1] * (g.v * g.c) + dh[2] * (g.v / p.rear_ax_dist);
      rb[0] = dh[0] * g.fx + dh[1] * g.fy;
      ra[0][2] = p.alpha * h;
    } else if (MODEL == SCB_DYNAMIC_UNICYCLE_2D) {                       // :144-149
      double h, hd, dhd[4];
      ModelCT<SCB_DYNAMIC_UNICYCLE_2D>::barrier(p, g, o, h, hd, dhd);
      ra[0][0] = dhd[3]; ra[0][1] = dhd[2];
      rb[0] = dhd[0] * g.fx + dhd[1] * g.fy;
      if (NW == 2) {
        ra[0][2] = (p.alpha1 + p.alpha2) * hd;
        ra[0][NV - 1] = (p.alpha1 * p.alpha2) * h;
      }
    } else if (MODEL == SCB_QUAD_2D) {                                   // :144-149 (same branch as DynamicUnicycle2D)
      double h, hd, dhd[4];
      ModelCT<SCB_QUAD_2D>::barrier(p, g, o, h, hd, dhd);
      const double a = dhd[2] * (-g.s / p.mass) + dhd[3] * (g.c / p.mass);
      ra[0][0] = a; ra[0][1] = a;
      rb[0] = dhd[0] * g.fx + dhd[1] * g.fy + dhd[3] * (-p.gravity);
      if (NW == 2) {
        ra[0][2] = (p.alpha1 + p.alpha2) * hd;
        ra[0][NV - 1] = (p.alpha1 * p.alpha2) * h;
      }
    }
    // plain KinematicBicycle2D: the reference has no branch -> zero row
  }
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    ra[1 + 2 * i][i] = -1.0; rb[1 + 2 * i] = p.u_ub[i];
    ra[2 + 2 * i][i] = 1.0;  rb[2 + 2 * i] = -p.u_lb[i];
  }

  double hd[NV], z0[NV];
#pragma unroll
  for (int i = 0; i < NU; ++i) { hd[i] = 2.0; z0[i] = ldx<NC>(uref + i); }
  hd[NU] = 2.0 * p.p_sb1; z0[NU] = p.omega1_0;
  if (NW == 2) { hd[NV - 1] = 2.0 * p.p_sb2; z0[NV - 1] = p.omega2_0; }

  QpOut<NV> q;
  gi_solve<NV, 1, 5>(hd, z0, ra, rb, 5, 64, q);     // 5 rows: replicated per lane, no shuffles

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      double v = q.x[i];
      if (q.status != SCB_OPTIMAL) v = fmin(fmax(v, p.u_lb[i]), p.u_ub[i]);
      U[i] = v;
    }
    if (omega) {
      omega[0] = q.x[NU];
      omega[1] = (NW == 2) ? q.x[NV - 1] : nan("");
    }
    if (sel) *sel = has ? bi : -1;
    *status = q.status;
    if (active) {
      uint64_t bits = 0ull;
#pragma unroll
      for (int a = 0; a < NV; ++a)
        if (a < q.wk && q.lam[a] > 0.0) bits |= 1ull << q.widx[a];
      *active = bits;
    }
  }
}

}  // namespace scb
