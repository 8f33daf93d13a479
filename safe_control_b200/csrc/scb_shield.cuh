// Trajectory-rollout shields for the double integrator in the evade scene: gatekeeper (backward search over the nominal
// horizon) and MPS (one nominal step), one control step of N agents per launch, shield state resident in HBM.
//
//   /root/reference/shielding/gatekeeper.py:271-307  _forward_simulate_backup
//   /root/reference/shielding/gatekeeper.py:309-367  _generate_candidate_trajectory (external nominal trajectory)
//   /root/reference/shielding/gatekeeper.py:380-471  _is_collision, _check_moving_obstacle_collision
//   /root/reference/shielding/gatekeeper.py:499-551  _is_candidate_valid, _update_committed_trajectory
//   /root/reference/shielding/gatekeeper.py:553-672  Gatekeeper.solve_control_problem
//   /root/reference/shielding/mps.py:59-160          MPS.solve_control_problem
//   /root/reference/envs/evade_env.py:408-485        check_collision, check_obstacle_collision
//
// One lane group per agent, ONE CANDIDATE PER LANE: candidate c switches from the nominal trajectory to the backup policy
// after max(T - c * discount, 0) nominal steps; the lane checks the nominal prefix (read from HBM), then rolls the backup
// policy out from the switching state (bk_step of scb_backup.cuh: the reference's arithmetic, operation for operation)
// checking every state against walls, the bullet where it is now and the moving obstacles at t = k dt, and stops at the
// first collision.  The reference walks the candidates from the longest nominal horizon down and takes the first valid
// one: here every lane keeps the first valid candidate of its own (strided, ascending) list and a group-wide minimum
// picks the winner -- the same candidate.  The winner's backup leg is rolled out once more to write the committed input
// (and, on request, state) trajectory: storing 32 speculative 120-step trajectories per agent would cost 61 KB each.
#pragma once
#include "scb_backup.cuh"

namespace scb {

// EvadeEnv.check_collision (evade_env.py:408-452)
SCB_HD bool sh_wall_hit(const scb_backup_params& p, double px, double py, double r) {
  if (py - r < -p.half_width) return true;
  if (py + r > p.half_width) {
    if (p.pocket_x_min <= px && px <= p.pocket_x_max) {
      if (py + r > p.pocket_y_max) return true;
      if (px - r < p.pocket_x_min && py > p.half_width) return true;
      if (px + r > p.pocket_x_max && py > p.half_width) return true;
    } else {
      return true;
    }
  }
  if (px - r < 0.0) return true;
  if (px + r > p.hallway_length) return true;
  return false;
}

// circle against an axis-aligned rectangle (evade_env.py:476-483, gatekeeper.py:452-462)
SCB_HD bool sh_rect_hit(double px, double py, double x_min, double x_max, double y_min, double y_max, double r) {
  const double cx = fmin(fmax(px, x_min), x_max), cy = fmin(fmax(py, y_min), y_max);
  const double dx = px - cx, dy = py - cy;
  return sqrt(nmul(dx, dx) + nmul(dy, dy)) < r;
}

// Gatekeeper._is_collision (gatekeeper.py:380-424) for candidate state k (obstacles at t = k dt)
SCB_HD bool sh_collides(const scb_backup_params& p, double px, double py, int k, const double* mov, int K, const double* stat) {
  if (sh_wall_hit(p, px, py, p.radius)) return true;
  if (stat && stat[4] != 0.0 && sh_rect_hit(px, py, stat[0], stat[1], stat[2], stat[3], p.radius)) return true;
  const double rr = p.radius + p.safety_margin;
  const double t = nmul((double)k, p.dt);
  for (int q = 0; q < K; ++q) {
    const double* o = mov + (size_t)q * kBkMov;
    const int kind = (int)o[7];
    if (kind == 0) continue;
    const double ox = o[0] + nmul(o[2], t), oy = o[1] + nmul(o[3], t);
    if (kind == 1) {
      if (sh_rect_hit(px, py, ox - o[4] / 2, ox + o[4] / 2, oy - o[5] / 2, oy + o[5] / 2, rr)) return true;
    } else {
      const double dx = px - ox, dy = py - oy;
      if (sqrt(nmul(dx, dx) + nmul(dy, dy)) < rr + o[6]) return true;
    }
  }
  return false;
}

// backup leg from s (n_backup steps); k0 = candidate index of the first backup state.  check: stop at the first collision
// and return false; cu / cx (may be null): write the inputs / states of the leg.
SCB_HD bool sh_backup_leg(const scb_backup_params& p, const double* s0, int k0, bool check, const double* mov, int K,
                          const double* stat, double* cu, double* cx) {
  double s[4] = {s0[0], s0[1], s0[2], s0[3]};
  for (int j = 0; j < p.n_backup; ++j) {
    double ax, ay, n[4];
    bk_policy(p, s, ax, ay);
    // (bk_step evaluates the policy itself; inlined here so that the input can be stored)
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
    if (cu) { cu[2 * j] = ax; cu[2 * j + 1] = ay; }
    if (cx) { cx[4 * j] = n[0]; cx[4 * j + 1] = n[1]; cx[4 * j + 2] = n[2]; cx[4 * j + 3] = n[3]; }
    if (check && sh_collides(p, n[0], n[1], k0 + j, mov, K, stat)) return false;
    s[0] = n[0]; s[1] = n[1]; s[2] = n[2]; s[3] = n[3];
  }
  return true;
}

struct ShieldIO {
  const double* x;        // [4] current state
  const double* nomx;     // [T + 1, 4] nominal states (nomx[0] = the state the nominal rollout started from)
  const double* nomu;     // [T, 2]
  int nom_len;            // states available (0: no nominal trajectory)
  const double* mov; int K;
  const double* stat;     // [5] or null
  double* cu;             // [T + n_backup, 2] committed inputs (the active one of the agent's two buffers)
  double* cx;             // [T + n_backup + 1, 4] committed states or null
  double* cu_spare;       // the other buffer: a new commitment is built there, then the two are swapped
  double* cx_spare;
};

// one control step of one agent; the scalar state is passed by reference and stored by the caller
// `flip` is set when the commitment moved to the spare buffer (the caller toggles the agent's buffer index).
// PHASE 0: the whole step.  The search can also be split in two launches (scb_shield_step with a work list):
//   PHASE 1 (a thread per agent): everything, but only candidate 0 is tried; when it fails and there are more candidates
//            the function returns false -- nothing but the first-call commitment has been changed, the agent is `pending`;
//   PHASE 2 (a lane group per pending agent): the remaining candidates 1 .. n_cand - 1 and the rest of the step.
// Candidate 0 wins for most agents, and 32 agents per warp walking their own 220 states run converged; the speculative
// lane-per-candidate search then only pays for the agents that need it.
template <int LANES, int PHASE = 0>
SCB_HD bool shield_agent(const scb_shield_params& sp, const ShieldIO& io, int& clen, int& cidx, int& nsteps, double& next_event,
                         double* u_out, int& using_backup, bool& flip) {
  using G = Grp<LANES>;
  const scb_backup_params& p = sp.scene;
  const int lane = G::lane();
  const int Nb = p.n_backup;
  const double dt = p.dt;
  const int nl = io.nom_len;
  double* cu = io.cu;
  flip = false;

  if (PHASE != 2 && clen < 0) {                     // first call: commit the pure backup trajectory (gatekeeper.py:571-583)
    if (lane == 0) {
      if (io.cx) { io.cx[0] = io.x[0]; io.cx[1] = io.x[1]; io.cx[2] = io.x[2]; io.cx[3] = io.x[3]; }
      sh_backup_leg(p, io.x, 1, false, nullptr, 0, nullptr, io.cu, io.cx ? io.cx + 4 : nullptr);
    }
    clen = Nb; nsteps = 0; cidx = 0; next_event = 0.0;
    bk_sync<LANES>();
  }

  const bool mps = sp.mode == 1;
  const bool event = (PHASE == 2) || (mps ? (nl > 1) : ((double)cidx >= next_event / dt));   // mps.py:92-95; gatekeeper.py:590
  if (event) {
    const int max_steps = mps ? 1 : (nl > 0 ? nl - 1 : 0);                      // gatekeeper.py:592-599; mps.py:88
    const int disc = sp.discount_steps > 0 ? sp.discount_steps : 1;
    const int n_cand = mps ? 1 : max_steps / disc + 2;                          // gatekeeper.py:605
    constexpr int kNone = 0x7fffffff;
    int best = kNone;
    const int c_begin = (PHASE == 2) ? 1 : 0, c_end = (PHASE == 1) ? 1 : n_cand;
    for (int c = c_begin + lane; c < c_end && best == kNone; c += LANES) {
      int steps = max_steps - c * disc;
      if (steps < 0) steps = 0;
      const int n_use = (nl > 0) ? ((steps + 1 < nl) ? steps + 1 : nl) : 1;     // gatekeeper.py:328-341
      const double* base = (nl > 0) ? io.nomx : io.x;
      bool ok = true;
      for (int k = 0; k < n_use && ok; ++k) ok = !sh_collides(p, base[4 * k], base[4 * k + 1], k, io.mov, io.K, io.stat);
      // candidate 0 (the longest nominal horizon, the winner for most agents) writes its backup leg straight into the spare
      // buffer while it is being checked: when it wins, nothing has to be rolled out twice
      const bool spec = (c == 0);
      if (ok) ok = sh_backup_leg(p, base + 4 * (n_use - 1), n_use, true, io.mov, io.K, io.stat,
                                 spec ? io.cu_spare + 2 * (n_use - 1) : nullptr,
                                 (spec && io.cx_spare) ? io.cx_spare + 4 * n_use : nullptr);
      if (ok) best = c;
    }
    best = (int)G::vmin((double)best);
    if (PHASE == 1 && best == kNone && n_cand > 1) return false;                // pending: phase 2 tries the other candidates
    if (best != kNone) {                            // _update_committed_trajectory (gatekeeper.py:529-551), built in the spare buffer
      int steps = max_steps - best * disc;
      if (steps < 0) steps = 0;
      const int n_use = (nl > 0) ? ((steps + 1 < nl) ? steps + 1 : nl) : 1;
      const int actual = n_use - 1;
      const double* base = (nl > 0) ? io.nomx : io.x;
      bk_sync<LANES>();                             // (candidate 0's speculative leg is complete and visible)
      for (int k = lane; k < 2 * actual; k += LANES) io.cu_spare[k] = io.nomu[k];
      if (io.cx_spare) for (int k = lane; k < 4 * n_use; k += LANES) io.cx_spare[k] = base[k];
      if (best != 0 && lane == 0)
        sh_backup_leg(p, base + 4 * (n_use - 1), n_use, false, nullptr, 0, nullptr, io.cu_spare + 2 * actual,
                      io.cx_spare ? io.cx_spare + 4 * n_use : nullptr);
      clen = actual + Nb; nsteps = actual; cidx = 0; next_event = sp.event_offset;
      cu = io.cu_spare; flip = true;
      bk_sync<LANES>();
    } else {
      next_event = nmul((double)cidx, dt) + sp.event_offset;                    // gatekeeper.py:654; mps.py:127
    }
  }

  double u0, u1;
  if (cidx < clen) { u0 = cu[2 * cidx]; u1 = cu[2 * cidx + 1]; }          // gatekeeper.py:656-667
  else bk_policy(p, io.x, u0, u1);
  u_out[0] = u0; u_out[1] = u1;
  cidx += 1;
  if (mps) {                                        // mps.py:142-153 (external nominal trajectory: u_ref = nominal_u_traj[0])
    bool match = false;
    if (nl > 1) {
      const double d0 = u0 - io.nomu[0], d1 = u1 - io.nomu[1];
      match = sqrt(d0 * d0 + d1 * d1) < 1e-2;
    }
    using_backup = match ? 0 : 1;
  } else {                                          // gatekeeper.py:741-744, evaluated after the increment
    using_backup = (cidx >= (int)(nmul((double)nsteps, dt) / dt)) ? 1 : 0;
  }
  return true;
}

}  // namespace scb
