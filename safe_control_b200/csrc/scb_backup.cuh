// Backup-CBF QP for the double integrator in the evade scene: rollout of the backup policy with forward-difference
// sensitivities, one CBF row per backup step + the terminal row, exact 2-variable QP in scaled inputs, fall-backs.
//
//   /root/reference/position_control/backup_cbf_qp.py:236-318  _integrate_backup_trajectory
//   /root/reference/position_control/backup_cbf_qp.py:341-458  _h_safety / _grad_h_safety
//   /root/reference/position_control/backup_cbf_qp.py:460-553  _h_terminal / _grad_h_terminal
//   /root/reference/position_control/backup_cbf_qp.py:563-794  solve_control_problem
//   /root/reference/position_control/backup_controller.py:456-575  EvadeBackupController.compute_control
//   /root/reference/robots/double_integrator2D.py:79-107  step
//
// One lane group per agent.  The reference does (5 closed-loop steps + 6 barrier evaluations) per backup step, 120 backup
// steps per call in the evade scenario; the 5 step variants (nominal + one per perturbed state) run on 5 lanes, the 4
// barrier variants (h, x + eps, y + eps, t + dt) on 4 lanes, and the rows are produced while the rollout advances (row i
// needs phi[i-1], phi[i], phi[i+1], S_i only), so nothing but the 3 numbers of each row is kept: [N, 3] doubles of shared
// memory per agent.  The rows then go to registers, lane-strided, for the same exact active-set QP the CBF-QP kernels use.
//
// Arithmetic: every quantity that is finite-differenced (eps = 1e-5 amplifies one ulp of h to 1e-11 of a gradient) is
// computed in the reference's own operation order with products rounded before they are added (nmul: never contracted
// into an FMA), so the rollout, the barrier values and the rows reproduce the numpy results to the last bit or two.
#pragma once
#include "scb_core.cuh"
#include "scb_gi.cuh"

namespace scb {

SCB_HD double nmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double t = a * b;          // a rounded product the compiler cannot fuse into the add that follows
  return t;
#endif
}

constexpr double kBkEps = 1e-5;       // backup_cbf_qp.py:285, 448, 543
constexpr int kBkMov = 8;             // doubles per moving obstacle: x, y, vx, vy, length, width, radius, kind

SCB_HD void bk_clamp(const scb_backup_params& p, double& ax, double& ay) {          // backup_controller.py:568-575
  const double a_mag = sqrt(nmul(ax, ax) + nmul(ay, ay));
  if (a_mag > p.a_max) {
    ax = nmul(ax, p.a_max) / a_mag;
    ay = nmul(ay, p.a_max) / a_mag;
  }
}

// EvadeBackupController.compute_control (backup_controller.py:456-566)
SCB_HD void bk_policy(const scb_backup_params& p, const double* s, double& ax, double& ay) {
  const double x = s[0], y = s[1], vx = s[2], vy = s[3];
  if (p.use_goal && p.goal_x_min <= x && x <= p.goal_x_max && p.goal_y_min <= y && y <= p.goal_y_max) {     // :476-484
    ax = nmul(-p.Kd, vx); ay = nmul(-p.Kd, vy);
    bk_clamp(p, ax, ay);
    return;
  }
  const double x_min = p.pocket_x_min, x_max = p.pocket_x_max, y_min = p.pocket_y_min, y_max = p.pocket_y_max;
  const double cx = p.center_x, cy = p.center_y;
  const double margin = p.radius + 0.1;                                                                       // :496
  const double ex = x - cx, ey = y - cy;
  const double dist_to_center = sqrt(nmul(ex, ex) + nmul(ey, ey));
  const bool x_in = (x_min + margin <= x) && (x <= x_max - margin);
  if (x_in && (y_min + margin <= y) && (y <= y_max - margin) && dist_to_center < 1.0) {                       // :502-508
    ax = nmul(-p.Kd, vx); ay = nmul(-p.Kd, vy);
  } else if (x_min - 2.0 <= x && x <= x_max + 2.0) {                                                          // :513-542
    const double error_x = cx - x;
    double error_y;
    if (x_in) {
      error_y = cy - y;
    } else {
      const double target_y = (y > y_min) ? fmax(y, 3.0) : 0.0;
      error_y = target_y - y;
    }
    ax = nmul(p.Kp, error_x) - nmul(p.Kd, vx);
    ay = nmul(p.Kp, error_y) - nmul(p.Kd, vy);
  } else {                                                                                                    // :546-563
    const double target_y = (y > y_min && x > x_max) ? fmax(y, 3.0) : 0.0;
    const double error_x = cx - x;
    const double error_y = target_y - y;
    const double sgn = (error_x > 0.0) ? 1.0 : ((error_x < 0.0) ? -1.0 : 0.0);
    ax = nmul(nmul(p.Kp, sgn), fmin(fabs(error_x), 3.0)) - nmul(p.Kd, vx);
    ay = nmul(p.Kp, error_y) - nmul(p.Kd, vy);
  }
  bk_clamp(p, ax, ay);
}

// one closed-loop step: DoubleIntegrator2D.step under the backup policy (double_integrator2D.py:79-107)
SCB_HD void bk_step(const scb_backup_params& p, const double* s, double* n) {
  double ax, ay;
  bk_policy(p, s, ax, ay);
  n[0] = s[0] + nmul(s[2], p.dt);
  n[1] = s[1] + nmul(s[3], p.dt);
  n[2] = s[2] + nmul(ax, p.dt);
  n[3] = s[3] + nmul(ay, p.dt);
  const double v_mag = sqrt(nmul(n[2], n[2]) + nmul(n[3], n[3]));
  if (v_mag > p.v_max) {
    const double scale = p.v_max / v_mag;
    n[2] = nmul(n[2], scale);
    n[3] = nmul(n[3], scale);
  }
}

// _h_safety, evade branch (backup_cbf_qp.py:356-389) + moving obstacles at time t (:418-442)
SCB_HD double bk_h_safety(const scb_backup_params& p, double px, double py, double t, const double* mov, int K) {
  const double r = p.radius;
  double h = py + p.half_width - r;
  h = fmin(h, px - r);
  h = fmin(h, p.hallway_length - px - r);
  if (p.pocket_x_min <= px && px <= p.pocket_x_max) {
    h = fmin(h, p.pocket_y_max - py - r);
    if (py > p.half_width) h = fmin(h, fmin(px - p.pocket_x_min - r, p.pocket_x_max - px - r));
  } else {
    h = fmin(h, p.half_width - py - r);
  }
  for (int k = 0; k < K; ++k) {
    const double* o = mov + (size_t)k * kBkMov;
    const int kind = (int)o[7];
    if (kind == 0) continue;
    const double ox = o[0] + nmul(o[2], t), oy = o[1] + nmul(o[3], t);
    if (kind == 1) {
      const double dx = fmax(fabs(px - ox) - o[4] / 2, 0.0), dy = fmax(fabs(py - oy) - o[5] / 2, 0.0);
      h = fmin(h, sqrt(nmul(dx, dx) + nmul(dy, dy)) - r - p.safety_margin);
    } else {
      const double dx = px - ox, dy = py - oy;
      h = fmin(h, sqrt(nmul(dx, dx) + nmul(dy, dy)) - r - o[6] - p.safety_margin);
    }
  }
  return h;
}

// _h_terminal, evade branch (backup_cbf_qp.py:472-536)
SCB_HD double bk_h_terminal(const scb_backup_params& p, const double* s, const double* mov, int K) {
  const double margin = p.radius + 0.2;
  double h = fmin(fmin(s[0] - p.pocket_x_min - margin, p.pocket_x_max - s[0] - margin),
                  fmin(s[1] - p.pocket_y_min - margin, p.pocket_y_max - s[1] - margin));
  h = fmin(h, p.v_max - sqrt(nmul(s[2], s[2]) + nmul(s[3], s[3])));
  return fmin(h, bk_h_safety(p, s[0], s[1], p.backup_horizon, mov, K));
}

// Scratch of one agent: SCR doubles.  [0, 20) step variants / A columns, [20, 36) S, [36, 44) barrier variants
constexpr int kBkScratch = 44;

// Lane groups of the rollout: the usual power-of-two groups (Grp<>), or LANES = 5 -- six agents per warp, lanes 30 and 31
// idle (they leave the kernel before the first barrier): exactly one lane per step variant, 30 of 32 lanes busy in the
// step phase and 24 in the barrier phase instead of 20 and 16 with groups of 8.  The six agents of a warp run the same
// number of backup steps, so the 30 lanes synchronise as one group.
constexpr unsigned kBkMask5 = 0x3fffffffu;
template <int LANES>
SCB_HD int bk_lane() {
#if defined(__CUDA_ARCH__)
  if (LANES == 5) return (int)((threadIdx.x & 31u) % 5u);
#endif
  return Grp<LANES == 5 ? 1 : LANES>::lane();
}
template <int LANES>
SCB_HD void bk_sync() {
#if defined(__CUDA_ARCH__)
  if (LANES == 5) __syncwarp(kBkMask5);
  else if (LANES > 1) __syncwarp(Grp<LANES == 5 ? 1 : LANES>::gmask());
#endif
}

// closed-loop step from x and its forward-difference Jacobian: scr[0..3] = x_next, scr[4 + 4 k + r] = A[r][k]
template <int LANES>
SCB_HD void bk_step_fd(const scb_backup_params& p, const double* x, double* scr) {
  const int lane = bk_lane<LANES>();
  for (int v = lane; v < 5; v += LANES) {
    const double xp[4] = {x[0] + (v == 1 ? kBkEps : 0.0), x[1] + (v == 2 ? kBkEps : 0.0), x[2] + (v == 3 ? kBkEps : 0.0),
                          x[3] + (v == 4 ? kBkEps : 0.0)};
    bk_step(p, xp, scr + 4 * v);
  }
  bk_sync<LANES>();
  for (int v = lane; v < 5; v += LANES)
    if (v > 0)
      for (int r = 0; r < 4; ++r) scr[4 * v + r] = (scr[4 * v + r] - scr[r]) / kBkEps;
  bk_sync<LANES>();
}

// S <- A S (S at scr + 20, row-major); lane c owns column c
template <int LANES>
SCB_HD void bk_advance_S(double* scr) {
  const int lane = bk_lane<LANES>();
  double* S = scr + 20;
  for (int c = lane; c < 4; c += LANES) {
    const double s0 = S[c], s1 = S[4 + c], s2 = S[8 + c], s3 = S[12 + c];
    for (int r = 0; r < 4; ++r)
      S[4 * r + c] = scr[4 + r] * s0 + scr[8 + r] * s1 + scr[12 + r] * s2 + scr[16 + r] * s3;
  }
  bk_sync<LANES>();
}

struct BackupOut {
  double u0, u1, h_min;
  int status, intervene, w0, w1;
  double lam0, lam1;
};

// rows[3 r + {0, 1, 2}] = lhs0, lhs1, rhs of row r (r = i - 1 for backup step i = 1 .. N-1; r = N - 1 the terminal row);
// phi (may be null) [N, 4] receives the backup trajectory.
// first half: the rollout and the rows; returns _last_h_min (:587-590).  `rows` may be shared or global memory.
template <int LANES>
SCB_HD double backup_rollout(const scb_backup_params& p, const double* x0, const double* mov, int K, double* scr,
                             double* rows, double* phi) {
  const int lane = bk_lane<LANES>();
  const int N = p.n_backup;
  const double dt = p.dt;
  double* S = scr + 20;
  double* hv = scr + 36;

  double x[4] = {x0[0], x0[1], x0[2], x0[3]}, prev[4] = {x0[0], x0[1], x0[2], x0[3]};
  for (int c = lane; c < 16; c += LANES) S[c] = (c % 5 == 0) ? 1.0 : 0.0;
  double h_min = bk_h_safety(p, x[0], x[1], 0.0, mov, K);                       // i = 0 of the status scan (:587-590)
  if (phi && lane == 0) for (int r = 0; r < 4; ++r) phi[r] = x[r];
  bk_sync<LANES>();
  if (N > 1) {                                                                  // phi[1], S_1 = A_0
    bk_step_fd<LANES>(p, x, scr);
    bk_advance_S<LANES>(scr);
    for (int r = 0; r < 4; ++r) x[r] = scr[r];
    bk_sync<LANES>();
  }
  for (int i = 1; i < N; ++i) {
    const double t_i = nmul((double)i, dt);
    if (phi && lane == 0) for (int r = 0; r < 4; ++r) phi[4 * i + r] = x[r];
    // barrier variants at phi[i]: h, x + eps, y + eps, t + dt  (:623-632, 446-458; the velocity components of the gradient
    // are (h - h) / eps = 0 exactly)
    for (int v = lane; v < 4; v += LANES)
      hv[v] = bk_h_safety(p, x[0] + (v == 1 ? kBkEps : 0.0), x[1] + (v == 2 ? kBkEps : 0.0), v == 3 ? t_i + dt : t_i, mov, K);
    double xn[4] = {x[0], x[1], x[2], x[3]};
    if (i < N - 1) {
      bk_step_fd<LANES>(p, x, scr);                                             // (syncs: hv is visible after it)
      for (int r = 0; r < 4; ++r) xn[r] = scr[r];
    } else {
      bk_sync<LANES>();
    }
    {
      const double h_val = hv[0];
      const double gx = (hv[1] - h_val) / kBkEps, gy = (hv[2] - h_val) / kBkEps;
      const double dh_dt = (hv[3] - h_val) / dt;
      double f0, f1;                                                            // f_pi (:636-639)
      if (i < N - 1) { f0 = (xn[0] - x[0]) / dt; f1 = (xn[1] - x[1]) / dt; }
      else           { f0 = (x[0] - prev[0]) / dt; f1 = (x[1] - prev[1]) / dt; }
      const double gS0 = gx * S[0] + gy * S[4], gS1 = gx * S[1] + gy * S[5];
      const double gS2 = gx * S[2] + gy * S[6], gS3 = gx * S[3] + gy * S[7];
      const double rhs = -(gS0 * x0[2] + gS1 * x0[3]) + (gx * f0 + gy * f1) - dh_dt - nmul(p.alpha, h_val);   // :646-647
      if (lane == 0 && rows) { rows[3 * (i - 1)] = gS2; rows[3 * (i - 1) + 1] = gS3; rows[3 * (i - 1) + 2] = rhs; }
      h_min = fmin(h_min, h_val);
    }
    bk_sync<LANES>();                                                           // S and hv are read before they change
    if (i < N - 1) {
      bk_advance_S<LANES>(scr);
      for (int r = 0; r < 4; ++r) { prev[r] = x[r]; x[r] = xn[r]; }
    }
  }
  // terminal row at phi[N-1] with S_{N-1} (:655-665): forward differences of h_terminal in all 4 states
  {
    for (int v = lane; v < 5; v += LANES) {
      const double xp[4] = {x[0] + (v == 1 ? kBkEps : 0.0), x[1] + (v == 2 ? kBkEps : 0.0), x[2] + (v == 3 ? kBkEps : 0.0),
                            x[3] + (v == 4 ? kBkEps : 0.0)};
      scr[v] = bk_h_terminal(p, xp, mov, K);
    }
    bk_sync<LANES>();
    const double h_T = scr[0];
    double g[4];
    for (int r = 0; r < 4; ++r) g[r] = (scr[1 + r] - h_T) / kBkEps;
    double gS[4];
    for (int c = 0; c < 4; ++c) gS[c] = g[0] * S[c] + g[1] * S[4 + c] + g[2] * S[8 + c] + g[3] * S[12 + c];
    const double rhs = -(gS[0] * x0[2] + gS[1] * x0[3] + nmul(p.alpha_terminal, h_T));
    if (lane == 0 && rows) { rows[3 * (N - 1)] = gS[2]; rows[3 * (N - 1) + 1] = gS[3]; rows[3 * (N - 1) + 2] = rhs; }
    h_min = fmin(h_min, h_T);
    bk_sync<LANES>();
  }
  return h_min;
}

// second half: the QP over the rows + the fall-backs
template <int LANES, int RPL>
SCB_HD void backup_qp(const scb_backup_params& p, const double* x0, const double* uref, const double* rows, double h_min,
                      BackupOut& out) {
  using G = Grp<LANES>;
  const int lane = G::lane();
  const int N = p.n_backup;
  // ---- QP in scaled inputs z = u / a_max, weighted by Q_u (:676-733); here v = Q_u z so that the cost is ||v - v_ref||^2 ----
  const double us = p.a_max;
  const double q[2] = {p.q0, p.q1};
  double ur[2] = {uref[0], uref[1]};
  double r0[RPL], r1[RPL], rb[RPL];
  const int mrows = N + 4;
  bool any = false;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    const int r = j * LANES + lane;
    double a0 = 0.0, a1 = 0.0, b = 1.0;                                        // vacuous
    if (r < N) {
      const double l0 = rows[3 * r], l1 = rows[3 * r + 1];
      if (sqrt(l0 * l0 + l1 * l1) > 1e-6) {                                     // the reference drops flat rows (:650, 663)
        a0 = l0 * us / q[0]; a1 = l1 * us / q[1]; b = -rows[3 * r + 2];
        any = true;
      }
    } else if (r < N + 4) {
      const int k = r - N;                                                      // z_k >= -1 (k = 0, 1), z_k <= 1 (k = 2, 3)
      const double sg = (k < 2) ? 1.0 : -1.0;
      a0 = (k & 1) ? 0.0 : sg; a1 = (k & 1) ? sg : 0.0; b = q[k & 1];
    }
    const double n2 = a0 * a0 + a1 * a1;
    const double inv = (n2 > 0.0) ? 1.0 / sqrt(n2) : 1.0;
    r0[j] = a0 * inv; r1[j] = a1 * inv; rb[j] = b * inv;
  }
  any = G::or_reduce(any ? 1u : 0u) != 0u;

  out.h_min = h_min; out.w0 = out.w1 = -1; out.lam0 = out.lam1 = 0.0;
  if (!any) {                                                                   // no rows: u_ref as it is (:785-789)
    out.u0 = ur[0]; out.u1 = ur[1]; out.status = SCB_OPTIMAL; out.intervene = 0;
    return;
  }
  ur[0] = fmin(fmax(ur[0], -us), us); ur[1] = fmin(fmax(ur[1], -us), us);       // :702
  const double inv_us = 1.0 / us;
  const double c0 = nmul(inv_us, ur[0]) * q[0], c1 = nmul(inv_us, ur[1]) * q[1];
  Qp2Out qp;
  gi_solve2<LANES, RPL>(2.0, c0, c1, r0, r1, rb, mrows, 8 * mrows + 16, qp);
  if (qp.status == SCB_OPTIMAL) {
    out.u0 = us * (qp.x0 / q[0]); out.u1 = us * (qp.x1 / q[1]);                 // :752
    const double d0 = qp.x0 - c0, d1 = qp.x1 - c1;
    out.intervene = sqrt(d0 * d0 + d1 * d1) > 0.1;                              // :757-766
    out.status = SCB_OPTIMAL;
    out.w0 = qp.w0; out.w1 = qp.w1; out.lam0 = qp.lam0; out.lam1 = qp.lam1;
  } else {                                                                      // :768-783
    out.status = SCB_INFEASIBLE;
    if (h_min > 0.01) {
      out.u0 = ur[0]; out.u1 = ur[1]; out.intervene = 0;
    } else {
      bk_policy(p, x0, out.u0, out.u1);
      out.intervene = 1;
    }
  }
}

}  // namespace scb

namespace scb {
template <int LANES, int RPL>
SCB_HD void backup_agent(const scb_backup_params& p, const double* x0, const double* uref, const double* mov, int K,
                         double* scr, double* rows, double* phi, BackupOut& out) {
  const double h_min = backup_rollout<LANES>(p, x0, mov, K, scr, rows, phi);
  backup_qp<LANES, RPL>(p, x0, uref, rows, h_min, out);
}
}  // namespace scb
