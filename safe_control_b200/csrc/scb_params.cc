// scb_params.cc -- the reference's per-(model, controller) defaults as a POD.
// No CUDA here: compiled into libscb.so by nvcc and into the CPU host-sim test aid by g++.
//
// Sources of every constant:
//   robots/robot.py:49                         radius 0.25 (set before the model ctor)
//   robots/single_integrator2D.py:41-42        v_max 1.0
//   robots/dynamic_unicycle2D.py:38-40         a_max 0.5, w_max 0.5, v_max 1.0
//   robots/kinematic_bicycle2D.py:44-53        wheel_base 0.4, rear_ax_dist 0.2, v_max 3.5, a_max 5.0,
//                                              delta_max 32 deg, beta_max = atan(L_r/L tan(delta_max)), v_min 0.2
//   robots/quad3D.py:53-61                     mass 3, Ix=Iy=Iz 0.5, L 0.3, nu 0.1, u in [-10, 10], g 9.8
//   position_control/cbf_qp.py:12-35           alpha 1.0 (SI) / 1.5 (C3BF); alpha1 = alpha2 = 1.5 (DU, KB)
//   position_control/optimal_decay_cbf_qp.py:17-50   alpha(1,2) 0.5, omega0 1.0, p_sb 1e4
//   position_control/mpc_cbf.py:19-39,49-82    Q, R, alpha per model
#include <math.h>
#include <string.h>

#include "../../include/scb.h"

extern "C" {

int scb_version(void) { return SCB_VERSION; }

size_t scb_params_sizeof(void) { return sizeof(scb_params); }

/* offsetof of a field by name (-1: unknown): lets a binding verify its struct mirror field by field, not only by size */
#include <stddef.h>
long scb_params_offsetof(const char* field) {
  if (!field) return -1;
#define F(name) if (strcmp(field, #name) == 0) return (long)offsetof(scb_params, name);
  F(model) F(cbf_mode) F(nx) F(nu) F(dt) F(radius) F(alpha) F(alpha1) F(alpha2) F(u_lb) F(u_ub) F(v_min) F(v_max) F(rear_ax_dist) F(omega1_0) F(omega2_0) F(p_sb1) F(p_sb2) F(Q) F(R) F(mass) F(Ix) F(Iy) F(Iz) F(arm_L) F(nu_coef) F(gravity) F(mpc_max_iter) F(mpc_superellipsoid) F(mpc_tol) F(S_wing) F(rho) F(C_L0) F(C_Lalpha) F(blend_M) F(alpha_0) F(C_Ldelta_e) F(C_D0) F(C_Dalpha) F(C_Ddelta_e) F(C_m0) F(C_malpha) F(C_mdelta_e) F(chord) F(k_front) F(k_rear) F(k_pusher) F(ell_f) F(ell_r) F(pitch_max) F(descent_speed_max) F(od_mpc) F(od_sum_rterms)
#undef F
  return -1;
}
long scb_track_offsetof(const char* field) {
  if (!field) return -1;
#define F(name) if (strcmp(field, #name) == 0) return (long)offsetof(scb_track, name);
  F(controller) F(N) F(K) F(M) F(W) F(H) F(enable_rotation) F(dynamic_obs) F(att_velocity_tracking) F(mpc_strict) F(reached_threshold) F(rotation_threshold) F(k_omega) F(k_a) F(k_v) F(k_a_stop) F(w_max) F(att_kp) F(wheel_base) F(delta_max) F(X) F(yaw) F(sm) F(wp_idx) F(WP) F(nwp) F(goal) F(has_goal) F(u_att) F(u_prev) F(ret) F(done) F(nsteps) F(SCENE) F(Uref) F(OBS) F(nobs) F(U) F(status) F(active) F(track_flag) F(mpc_iters) F(mpc_ws) F(mpc_ws_bytes) F(mpc_fail)
#undef F
  return -1;
}


const char* scb_strerror(int err) {
  switch (err) {
    case SCB_OK: return "ok";
    case SCB_ERR_BAD_ARG: return "bad argument (null pointer, negative size or unknown model)";
    case SCB_ERR_UNSUPPORTED: return "model/controller pair not supported (the reference has no branch for it either)";
    case SCB_ERR_TOO_LARGE: return "obstacle slots or horizon beyond the compiled limits (see scb_limits)";
    case SCB_ERR_CUDA: return "CUDA launch or copy failed (see scb_last_cuda_error)";
    case SCB_ERR_NO_DEVICE: return "no CUDA device";
    case SCB_ERR_ALLOC: return "allocation failed";
    default: return "unknown error";
  }
}

int scb_model_dims(int model, int* nx, int* nu) {
  int x, u;
  switch (model) {
    case SCB_SINGLE_INTEGRATOR_2D: x = 2; u = 2; break;
    case SCB_DYNAMIC_UNICYCLE_2D:
    case SCB_KINEMATIC_BICYCLE_2D:
    case SCB_KINEMATIC_BICYCLE_2D_C3BF: x = 4; u = 2; break;
    case SCB_QUAD_3D: x = 12; u = 4; break;
    case SCB_DOUBLE_INTEGRATOR_2D:
    case SCB_KINEMATIC_BICYCLE_2D_DPCBF: x = 4; u = 2; break;
    case SCB_QUAD_2D: x = 6; u = 2; break;
    case SCB_UNICYCLE_2D: x = 3; u = 2; break;
    case SCB_MANIPULATOR_2D: x = 3; u = 3; break;
    case SCB_VTOL_2D: x = 6; u = 4; break;
    default: return SCB_ERR_BAD_ARG;
  }
  if (nx) *nx = x;
  if (nu) *nu = u;
  return SCB_OK;
}

int scb_active_words(int M, int nu) { return (M + 2 * nu + 63) / 64; }

int scb_mpc_active_words(const scb_params* p, int M, int H) {
  if (!p || M < 0 || H < 1) return SCB_ERR_BAD_ARG;
  int nx = 0, nu = 0;
  if (scb_model_dims(p->model, &nx, &nu) != SCB_OK) return SCB_ERR_BAD_ARG;
  // models whose MPC bounds the velocity state (mpc_cbf.py:193-199, 205-211): DynamicUnicycle2D, KinematicBicycle2D*
  const bool vbound = p->model == SCB_DYNAMIC_UNICYCLE_2D || p->model == SCB_KINEMATIC_BICYCLE_2D ||
                      p->model == SCB_KINEMATIC_BICYCLE_2D_C3BF || p->model == SCB_KINEMATIC_BICYCLE_2D_DPCBF;
  const int nsb = vbound ? 2 : (p->model == SCB_VTOL_2D ? 5 : 0);       // state-bound rows per node (mpc_cbf.py:193-232)
  return (H * M + 2 * H * nu + nsb * H + 63) / 64;
}


size_t scb_shield_params_sizeof(void) { return sizeof(scb_shield_params); }
size_t scb_backup_params_sizeof(void) { return sizeof(scb_backup_params); }
int scb_backup_active_words(int n_backup) { return n_backup < 1 ? SCB_ERR_BAD_ARG : (n_backup + 4 + 63) / 64; }

// examples/evade/test_evade.py:60-100 (EnvironmentConfig, RobotConfig, SimulationConfig), envs/evade_env.py:62-76
void scb_backup_params_default(scb_backup_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  const double hallway_length = 60.0, hallway_width = 4.0, pocket_x = 25.0, pocket_length = 10.0, pocket_width = 4.0,
               goal_length = 5.0;
  p->hallway_length = hallway_length;
  p->half_width = hallway_width / 2;
  p->pocket_x_min = pocket_x;
  p->pocket_x_max = pocket_x + pocket_length;
  p->pocket_y_min = p->half_width;
  p->pocket_y_max = p->half_width + pocket_width;
  p->center_x = (p->pocket_x_min + p->pocket_x_max) / 2;
  p->center_y = (p->pocket_y_min + p->pocket_y_max) / 2;
  p->goal_x_min = hallway_length - goal_length;
  p->goal_x_max = hallway_length;
  p->goal_y_min = -p->half_width;
  p->goal_y_max = p->half_width;
  p->use_goal = 1;
  p->radius = 0.5; p->a_max = 2.0; p->v_max = 1.5; p->safety_margin = 0.5;
  p->Kp = 2.0; p->Kd = 2.0;
  p->dt = 0.1; p->backup_horizon = 12.0;
  p->n_backup = (int)(p->backup_horizon / p->dt);
  p->alpha = 1.0; p->alpha_terminal = 2.0;
  p->q0 = 1.0; p->q1 = 1.0;
}


int scb_params_default(scb_params* p, int model, const char* controller) {
  if (!p || !controller) return SCB_ERR_BAD_ARG;
  memset(p, 0, sizeof(*p));
  int nx, nu;
  if (scb_model_dims(model, &nx, &nu) != SCB_OK) return SCB_ERR_BAD_ARG;
  const bool qp = strcmp(controller, "cbf_qp") == 0;
  const bool od = strcmp(controller, "optimal_decay_cbf_qp") == 0;
  const bool odm = strcmp(controller, "optimal_decay_mpc_cbf") == 0;
  const bool mpc = strcmp(controller, "mpc_cbf") == 0 || odm;
  if (!qp && !od && !mpc) return SCB_ERR_BAD_ARG;
  p->model = model; p->nx = nx; p->nu = nu;
  p->dt = 0.05;
  p->radius = 0.25;
  p->gravity = 9.8;
  p->mpc_max_iter = 150;
  p->mpc_tol = 1e-8;
  p->v_min = -1e300; p->v_max = 1e300;
  switch (model) {
    case SCB_SINGLE_INTEGRATOR_2D:
      p->u_lb[0] = p->u_lb[1] = -1.0; p->u_ub[0] = p->u_ub[1] = 1.0;
      if (qp) p->alpha = 1.0;
      if (mpc) { p->alpha = 0.05; p->Q[0] = p->Q[1] = 50; p->R[0] = p->R[1] = 5; }
      if (od) return SCB_ERR_UNSUPPORTED;             // optimal_decay_cbf_qp.py:51-52 raises
      break;
    case SCB_MANIPULATOR_2D:                           // manipulator2D.py:19-20, cbf_qp.py:34-35, 96-105
      if (!qp) return SCB_ERR_UNSUPPORTED;             // no agent_barrier_dt (no MPC), no optimal-decay branch
      for (int i = 0; i < 3; ++i) { p->u_lb[i] = -2.0; p->u_ub[i] = 2.0; }
      p->alpha = 1.0;
      break;
    case SCB_UNICYCLE_2D:                              // unicycle2D.py:40-41, cbf_qp.py:14-15,58-61, mpc_cbf.py:22-24,53-55,188-192
      if (od) return SCB_ERR_UNSUPPORTED;              // optimal_decay_cbf_qp.py has no Unicycle2D branch (raises)
      p->u_lb[0] = -1.0; p->u_ub[0] = 1.0; p->u_lb[1] = -0.5; p->u_ub[1] = 0.5;
      if (qp) p->alpha = 1.0;
      if (mpc) { p->alpha = 0.05; p->Q[0] = p->Q[1] = 50; p->Q[2] = 0.01; p->R[0] = p->R[1] = 0.5; }
      break;
    case SCB_DYNAMIC_UNICYCLE_2D:
      p->u_lb[0] = -0.5; p->u_ub[0] = 0.5; p->u_lb[1] = -0.5; p->u_ub[1] = 0.5;
      p->v_max = 1.0; p->v_min = -1.0;
      if (qp) p->alpha1 = p->alpha2 = 1.5;
      if (od) { p->alpha1 = p->alpha2 = 0.5; p->omega1_0 = p->omega2_0 = 1.0; p->p_sb1 = p->p_sb2 = 1e4; }
      if (mpc) { p->alpha1 = p->alpha2 = 0.15; p->Q[0] = p->Q[1] = 50; p->Q[2] = 0.01; p->Q[3] = 30; p->R[0] = p->R[1] = 0.5; }
      break;
    case SCB_KINEMATIC_BICYCLE_2D:
    case SCB_KINEMATIC_BICYCLE_2D_C3BF:
    case SCB_KINEMATIC_BICYCLE_2D_DPCBF: {
      const double Lr = 0.2, L = 0.4, dmax = 32.0 * M_PI / 180.0;
      const double bmax = atan(Lr / L * tan(dmax));
      p->rear_ax_dist = Lr;
      p->u_lb[0] = -5.0; p->u_ub[0] = 5.0; p->u_lb[1] = -bmax; p->u_ub[1] = bmax;
      p->v_min = 0.2; p->v_max = 3.5;
      const bool c3 = model != SCB_KINEMATIC_BICYCLE_2D;          // C3BF and DPCBF: relative degree 1
      if (model == SCB_KINEMATIC_BICYCLE_2D_DPCBF && od) return SCB_ERR_UNSUPPORTED;    // optimal_decay_cbf_qp.py:51-52 raises
      if (qp) { if (c3) p->alpha = 1.5; else p->alpha1 = p->alpha2 = 1.5; }
      if (od) {
        p->omega1_0 = 1.0; p->p_sb1 = 1e4;
        if (c3) p->alpha = 0.5; else { p->alpha1 = p->alpha2 = 0.5; p->omega2_0 = 1.0; p->p_sb2 = 1e4; }
      }
      if (mpc) {
        if (c3) p->alpha = 0.15; else p->alpha1 = p->alpha2 = 0.1;
        p->Q[0] = p->Q[1] = 50; p->Q[2] = 1; p->Q[3] = 1; p->R[0] = 0.5; p->R[1] = 5000.0;
      }
      break;
    }
    case SCB_DOUBLE_INTEGRATOR_2D:                     // double_integrator2D.py:40-44, cbf_qp.py:18-20,66-69
      if (od) return SCB_ERR_UNSUPPORTED;              // optimal decay raises NotCompatibleError
      p->u_lb[0] = p->u_lb[1] = -1.0; p->u_ub[0] = p->u_ub[1] = 1.0;
      p->v_max = 1.0; p->v_min = -1.0;
      p->alpha1 = p->alpha2 = 1.5;
      if (mpc) {                                       // mpc_cbf.py:28-30, 60-63
        p->alpha1 = p->alpha2 = 0.2;
        p->Q[0] = p->Q[1] = 50; p->Q[2] = p->Q[3] = 20; p->R[0] = p->R[1] = 0.5;
      }
      break;
    case SCB_QUAD_2D:                                  // quad2D.py:40-46, cbf_qp.py:30-32,74-79, optimal_decay_cbf_qp.py:38-45
      p->mass = 1.0; p->Iy = 0.01; p->gravity = 9.81;
      if (mpc) {                                       // mpc_cbf.py:34-36, 74-77
        p->alpha1 = p->alpha2 = 0.15;
        p->Q[0] = p->Q[1] = 25; p->Q[2] = 50; p->Q[3] = p->Q[4] = 10; p->Q[5] = 50; p->R[0] = p->R[1] = 0.5;
      }
      p->u_lb[0] = p->u_lb[1] = 1.0; p->u_ub[0] = p->u_ub[1] = 10.0;
      if (qp) p->alpha1 = p->alpha2 = 1.5;
      if (od) { p->alpha1 = p->alpha2 = 0.5; p->omega1_0 = p->omega2_0 = 1.0; p->p_sb1 = p->p_sb2 = 1e4; }
      break;
    case SCB_VTOL_2D: {                                // robots/vtol2D.py:57-110; mpc_cbf.py:40-43, 83-87, 222-232
      if (!mpc) return SCB_ERR_UNSUPPORTED;            // agent_barrier is not implemented (vtol2D.py:458-460)
      p->mass = 11.0; p->Iy = 1.135; p->gravity = 9.81;
      p->S_wing = 0.55; p->rho = 1.2682; p->C_L0 = 0.23; p->C_Lalpha = 5.61; p->blend_M = 50.0; p->alpha_0 = 15.0 * M_PI / 180.0;
      p->C_Ldelta_e = 0.13; p->C_D0 = 0.043; p->C_Dalpha = 0.03; p->C_Ddelta_e = 0.0;
      p->C_m0 = 0.0135; p->C_malpha = -2.74; p->C_mdelta_e = -0.99; p->chord = 0.18994;
      p->k_front = 70.0; p->k_rear = 70.0; p->k_pusher = 60.0; p->ell_f = 0.5; p->ell_r = 0.5;
      p->v_max = 15.0; p->v_min = -15.0; p->pitch_max = 15.0; p->descent_speed_max = 5.0;
      for (int i = 0; i < 3; ++i) { p->u_lb[i] = 0.0; p->u_ub[i] = 1.0; }     // throttle_min / max
      p->u_lb[3] = -0.5; p->u_ub[3] = 0.5;                                       // elevator_min / max
      const double Q[6] = {10, 10, 250, 10, 10, 50};
      for (int i = 0; i < 6; ++i) p->Q[i] = Q[i];
      p->R[0] = p->R[1] = p->R[2] = 0.5; p->R[3] = 50000.0;
      p->alpha1 = p->alpha2 = 0.05;
      break;
    }
    case SCB_QUAD_3D: {
      if (!mpc) return SCB_ERR_UNSUPPORTED;            // agent_barrier raises, quad3D.py:269-273
      p->mass = 3.0; p->Ix = p->Iy = p->Iz = 0.5; p->arm_L = 0.3; p->nu_coef = 0.1;
      for (int i = 0; i < 4; ++i) { p->u_lb[i] = -10.0; p->u_ub[i] = 10.0; p->R[i] = 1.0; }
      const double Q[12] = {30, 30, 5, 20, 20, 1, 10, 10, 10, 20, 20, 1};
      for (int i = 0; i < 12; ++i) p->Q[i] = Q[i];
      p->alpha = 0.15;
      break;
    }
  }
  if (odm) {
    // optimal_decay_mpc_cbf.py:19-20 admits DynamicUnicycle2D, KinematicBicycle2D, Quad2D, Quad3D, VTOL2D; built here are the
    // relative-degree-2 ones, where the omegas enter the CBF row (:296-300).  (Quad3D's row has no omega, :292-295.)
    if (model != SCB_DYNAMIC_UNICYCLE_2D && model != SCB_KINEMATIC_BICYCLE_2D && model != SCB_QUAD_2D && model != SCB_VTOL_2D)
      return SCB_ERR_UNSUPPORTED;
    p->od_mpc = 1; p->od_sum_rterms = 0;
    p->omega1_0 = p->omega2_0 = 1.0; p->p_sb1 = p->p_sb2 = 10.0;                   // :87-90
    switch (model) {                                                               // gains :60-85, weights :28-50
      case SCB_DYNAMIC_UNICYCLE_2D: p->alpha1 = p->alpha2 = 0.01; break;
      case SCB_KINEMATIC_BICYCLE_2D: p->alpha1 = p->alpha2 = 0.05; p->R[1] = 50.0; break;
      case SCB_QUAD_2D: p->alpha1 = p->alpha2 = 0.15; break;
      case SCB_VTOL_2D: p->alpha1 = p->alpha2 = 0.35; break;
    }
  }
  return SCB_OK;
}

}  // extern "C"
