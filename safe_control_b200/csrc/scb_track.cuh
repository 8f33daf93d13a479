// scb_track.cuh -- per-agent bodies of the closed loop around the solve (device + host-sim).
//
// Everything LocalTrackingController.control_step() (tracking.py:559-668) does either side of
// pos_controller.solve_control_problem(), one lane group per agent:
//
//   track_pre_agent   state machine + update_goal (tracking.py:497-535, 569-578),
//                     get_nearest_unpassed_obs (:345-403), u_ref from nominal_input / stop /
//                     rotate_to (:589-604)
//   track_post_agent  VelocityTrackingYaw (:621-624), is_collide_unknown (:445-495), status
//                     check (:627-634), robot.step (:637), second collision check (:640-646),
//                     return code (:666-668)
//
// Model laws are transcribed from robots/<model>.py (line numbers at each function); the
// scalar work runs replicated in every lane, the loops over scene obstacles are strided over
// the lanes of the group.
#pragma once

#include "scb_core.cuh"

namespace scb {

constexpr double kPi = 3.14159265358979323846;

// numpy branch of angle_normalize: Python's floored `%` (e.g. robots/dynamic_unicycle2D.py:14-16)
SCB_HD double wrap_floor(double x) {
  const double two_pi = 2.0 * kPi;
  double r = fmod(x + kPi, two_pi);
  if (r != 0.0 && r < 0.0) r += two_pi;
  return r - kPi;
}

SCB_HD double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }   // np.clip

// ------------------------------------------------------------------------------------------
// per-model laws used by the loop
template <int MODEL>
struct ModelLoop;

template <>
struct ModelLoop<SCB_SINGLE_INTEGRATOR_2D> {
  static constexpr int NX = 2, NU = 2, NPOS = 2;
  static constexpr bool HAS_ATT = true;                       // yaw kept outside the state (robots/robot.py:72-73)
  static SCB_HD double half_angle() { return kPi; }           // tracking.py:352-353
  static SCB_HD double yaw_of(const double*, double yaw) { return yaw; }
  // single_integrator2D.py:72-90
  static SCB_HD void nominal(const scb_params& p, const scb_track& t, const double* x, const double* g, double* u) {
    double e0 = g[0] - x[0], e1 = g[1] - x[1];
    const double s0 = (e0 > 0.0) - (e0 < 0.0), s1 = (e1 > 0.0) - (e1 < 0.0);
    e0 = s0 * fmax(fabs(e0) - 0.05, 0.0);
    e1 = s1 * fmax(fabs(e1) - 0.05, 0.0);
    double v0 = t.k_v * e0, v1 = t.k_v * e1;                  // facade passes k_v (robots/robot.py:403)
    const double mag = sqrt(v0 * v0 + v1 * v1), vmax = p.u_ub[0];
    if (mag > vmax) { v0 = v0 * vmax / mag; v1 = v1 * vmax / mag; }
    u[0] = v0; u[1] = v1;
  }
  static SCB_HD void stop(const scb_params&, const scb_track&, const double*, double* u) { u[0] = 0.0; u[1] = 0.0; }   // :99-102
  static SCB_HD bool has_stopped(const double*) { return true; }                                      // :104-106
  // rotate state: u_att = rotate_to(yaw, theta) (:108-112), u_ref = stop() (tracking.py:592-594)
  static SCB_HD void rotate_to(const scb_params& p, const scb_track& t, const double* x, double yaw, double th,
                               double* u, double& u_att) {
    u_att = clipd(2.0 * wrap_floor(th - yaw), -t.w_max, t.w_max);
    stop(p, t, x, u);
  }
  static SCB_HD void step(const scb_params& p, double* x, const double* u) {                            // :64-66
    x[0] = x[0] + u[0] * p.dt; x[1] = x[1] + u[1] * p.dt;
  }
  // VelocityTrackingYaw follows the commanded velocity (velocity_tracking_yaw.py:41-43)
  static SCB_HD void att_velocity(const double*, const double* u, double& vx, double& vy) { vx = u[0]; vy = u[1]; }
};

// DoubleIntegrator2D (robots/double_integrator2D.py): X = [x, y, vx, vy], yaw kept outside the state (robots/robot.py:80-82)
template <>
struct ModelLoop<SCB_DOUBLE_INTEGRATOR_2D> {
  static constexpr int NX = 4, NU = 2, NPOS = 2;
  static constexpr bool HAS_ATT = true;
  static SCB_HD double half_angle() { return kPi; }           // tracking.py:352-353
  static SCB_HD double yaw_of(const double*, double yaw) { return yaw; }
  // :116-143; the facade passes (d_min, k_v, k_a) (robots/robot.py:408-409); a_max = the cbf_qp input bound
  static SCB_HD void nominal(const scb_params& p, const scb_track& t, const double* x, const double* g, double* u) {
    double e0 = g[0] - x[0], e1 = g[1] - x[1];
    const double s0 = (e0 > 0.0) - (e0 < 0.0), s1 = (e1 > 0.0) - (e1 < 0.0);
    e0 = s0 * fmax(fabs(e0) - 0.05, 0.0);
    e1 = s1 * fmax(fabs(e1) - 0.05, 0.0);
    double v0 = t.k_v * e0, v1 = t.k_v * e1;
    const double vm = sqrt(v0 * v0 + v1 * v1);
    if (vm > p.v_max) { v0 = v0 * p.v_max / vm; v1 = v1 * p.v_max / vm; }
    double a0 = t.k_a * (v0 - x[2]), a1 = t.k_a * (v1 - x[3]);
    const double am = sqrt(a0 * a0 + a1 * a1), amax = p.u_ub[0];
    if (am > amax) { a0 = a0 * amax / am; a1 = a1 * amax / am; }
    u[0] = a0; u[1] = a1;
  }
  static SCB_HD void stop(const scb_params&, const scb_track& t, const double* x, double* u) {          // :147-153
    u[0] = t.k_a_stop * (0.0 - x[2]); u[1] = t.k_a_stop * (0.0 - x[3]);
  }
  static SCB_HD bool has_stopped(const double* x) { return sqrt(x[2] * x[2] + x[3] * x[3]) < 0.05; }     // :155-156
  static SCB_HD void rotate_to(const scb_params& p, const scb_track& t, const double* x, double yaw, double th,
                               double* u, double& u_att) {                                              // :158-162
    u_att = clipd(2.0 * wrap_floor(th - yaw), -t.w_max, t.w_max);
    stop(p, t, x, u);
  }
  static SCB_HD void step(const scb_params& p, double* x, const double* u) {                            // :80-108
    const double vx = x[2], vy = x[3];
    x[0] = x[0] + vx * p.dt; x[1] = x[1] + vy * p.dt;
    double nvx = vx + u[0] * p.dt, nvy = vy + u[1] * p.dt;
    const double vm = sqrt(nvx * nvx + nvy * nvy);
    if (vm > p.v_max) { const double sc = p.v_max / vm; nvx *= sc; nvy *= sc; }
    x[2] = nvx; x[3] = nvy;
  }
  // VelocityTrackingYaw follows the STATE velocity for this model (velocity_tracking_yaw.py:44-50, preview_time = 0)
  static SCB_HD void att_velocity(const double* x, const double*, double& vx, double& vy) { vx = x[2]; vy = x[3]; }
};

// Unicycle2D (robots/unicycle2D.py): X = [x, y, theta], U = [v, omega]
template <>
struct ModelLoop<SCB_UNICYCLE_2D> {
  static constexpr int NX = 3, NU = 2, NPOS = 2;
  static constexpr bool HAS_ATT = false;
  static SCB_HD double half_angle() { return 0.6 * kPi; }     // angle_unpassed = 1.2 pi (tracking.py:354-355)
  static SCB_HD double yaw_of(const double* x, double) { return x[2]; }
  // :70-86; the facade passes (d_min, k_omega, k_v) (robots/robot.py:404-405)
  static SCB_HD void nominal(const scb_params&, const scb_track& t, const double* x, const double* g, double* u) {
    const double dx = x[0] - g[0], dy = x[1] - g[1];
    const double dist = fmax(sqrt(dx * dx + dy * dy) - 0.05, 0.05);
    const double err = wrap_floor(atan2(g[1] - x[1], g[0] - x[0]) - x[2]);
    u[1] = t.k_omega * err;
    u[0] = (fabs(err) > 90.0 * (kPi / 180.0)) ? 0.0 : t.k_v * dist * cos(err);
  }
  static SCB_HD void stop(const scb_params&, const scb_track&, const double*, double* u) { u[0] = 0.0; u[1] = 0.0; }   // :88-89
  static SCB_HD bool has_stopped(const double*) { return true; }                                                      // :91-93
  static SCB_HD void rotate_to(const scb_params&, const scb_track&, const double* x, double, double th, double* u,
                               double&) {                                                                            // :95-98
    u[0] = 0.0; u[1] = 2.0 * wrap_floor(th - x[2]);
  }
  static SCB_HD void step(const scb_params& p, double* x, const double* u) {                                          // :65-68
    double s, c; sincos_pair(x[2], s, c);
    x[0] = x[0] + (c * u[0]) * p.dt;
    x[1] = x[1] + (s * u[0]) * p.dt;
    x[2] = wrap_floor(x[2] + u[1] * p.dt);
  }
};

// Quad2D (robots/quad2D.py): X = [x, z, theta, x_dot, z_dot, theta_dot], U = rotor forces [F_r, F_l]
template <>
struct ModelLoop<SCB_QUAD_2D> {
  static constexpr int NX = 6, NU = 2, NPOS = 2;
  static constexpr bool HAS_ATT = false;
  static SCB_HD double half_angle() { return kPi; }           // tracking.py:352-353
  static SCB_HD double yaw_of(const double* x, double) { return x[2]; }
  // cascaded PD law with the function's own default gains (:87-150; the facade passes no gains, robots/robot.py:410-413)
  static SCB_HD void nominal(const scb_params& p, const scb_track&, const double* x, const double* g, double* u) {
    const double grav = 9.81;
    const double adx = 3.0 * (g[0] - x[0]) + 0.5 * (-x[3]);
    const double adz = (0.1 * (g[1] - x[1]) + 0.5 * (-x[4])) + grav;
    const double T = p.mass * sqrt(adx * adx + adz * adz);
    const double th_d = -atan2(adx, adz);
    double e = th_d - x[2];
    e = atan2(sin(e), cos(e));
    const double tau = clipd(0.05 * e + 0.05 * (-x[5]), -1.0, 1.0);
    u[0] = clipd((T + tau / p.radius) / 2.0, p.u_lb[0], p.u_ub[0]);
    u[1] = clipd((T - tau / p.radius) / 2.0, p.u_lb[0], p.u_ub[0]);
  }
  static SCB_HD void stop(const scb_params& p, const scb_track& t, const double* x, double* u) {        // :152-161
    const double here[2] = {x[0], x[1]};
    nominal(p, t, x, here, u);
  }
  static SCB_HD bool has_stopped(const double* x) { return sqrt(x[3] * x[3] + x[4] * x[4]) < 0.05; }     // :163-165
  static SCB_HD void rotate_to(const scb_params&, const scb_track&, const double* x, double, double th, double* u,
                               double&) {                                                              // :167-171
    u[0] = 0.0; u[1] = 2.0 * wrap_floor(th - x[2]);
  }
  static SCB_HD void step(const scb_params& p, double* x, const double* u) {                            // :46-90
    double s, c; sincos_pair(x[2], s, c);
    const double us = u[0] + u[1];
    const double f3 = 0.0 + (-s / p.mass) * u[0] + (-s / p.mass) * u[1];
    const double f4 = -9.81 + (c / p.mass) * u[0] + (c / p.mass) * u[1];
    const double f5 = 0.0 + (p.radius / p.Iy) * u[0] + (-(p.radius / p.Iy)) * u[1];
    (void)us;
    const double x3 = x[3], x4 = x[4], x5 = x[5];
    x[0] = x[0] + x3 * p.dt; x[1] = x[1] + x4 * p.dt;
    x[2] = wrap_floor(x[2] + x5 * p.dt);
    x[3] = x3 + f3 * p.dt; x[4] = x4 + f4 * p.dt; x[5] = x5 + f5 * p.dt;
  }
};

template <>
struct ModelLoop<SCB_DYNAMIC_UNICYCLE_2D> {
  static constexpr int NX = 4, NU = 2, NPOS = 2;
  static constexpr bool HAS_ATT = false;
  static SCB_HD double half_angle() { return 0.6 * kPi; }     // angle_unpassed = 1.2 pi (tracking.py:354-355)
  static SCB_HD double yaw_of(const double* x, double) { return x[2]; }
  // dynamic_unicycle2D.py:80-104
  static SCB_HD void nominal(const scb_params& p, const scb_track& t, const double* x, const double* g, double* u) {
    const double dx = x[0] - g[0], dy = x[1] - g[1];
    const double dist = fmax(sqrt(dx * dx + dy * dy) - 0.05, 0.0);
    const double th_d = atan2(g[1] - x[1], g[0] - x[0]);
    const double err = wrap_floor(th_d - x[2]);
    const double omega = t.k_omega * err;
    double v;
    if (fabs(err) > 90.0 * (kPi / 180.0)) v = 0.0;            // np.deg2rad(90) = 90 * (pi / 180)
    else v = fmin(t.k_v * dist * cos(err), p.v_max);
    u[0] = t.k_a * (v - x[3]);
    u[1] = omega;
  }
  static SCB_HD void stop(const scb_params&, const scb_track& t, const double* x, double* u) {          // :106-111
    u[0] = t.k_a_stop * (0.0 - x[3]); u[1] = 0.0;            // stop() is called without gains (robots/robot.py:424): 1.0 or nominal_k_a
  }
  static SCB_HD bool has_stopped(const double* x) { return fabs(x[3]) < 0.05; }                         // :113-114
  static SCB_HD void rotate_to(const scb_params&, const scb_track&, const double* x, double, double th, double* u,
                               double&) {                                                              // :116-119
    u[0] = 0.0; u[1] = 2.0 * wrap_floor(th - x[2]);
  }
  static SCB_HD void step(const scb_params& p, double* x, const double* u) {                            // :75-78
    double s, c; sincos_pair(x[2], s, c);
    const double v = x[3];
    x[0] = x[0] + (v * c) * p.dt;
    x[1] = x[1] + (v * s) * p.dt;
    x[2] = wrap_floor(x[2] + u[1] * p.dt);
    x[3] = v + u[0] * p.dt;
  }
};

template <>
struct ModelLoop<SCB_KINEMATIC_BICYCLE_2D> {
  static constexpr int NX = 4, NU = 2, NPOS = 2;
  static constexpr bool HAS_ATT = false;
  static SCB_HD double half_angle() { return kPi; }           // tracking.py:356-357
  static SCB_HD double yaw_of(const double* x, double) { return x[2]; }
  // kinematic_bicycle2D.py:125-147; the facade passes (d_min, k_omega, k_a, k_v) positionally (robots/robot.py:406-407)
  static SCB_HD void nominal(const scb_params& p, const scb_track& t, const double* x, const double* g, double* u) {
    const double dx = x[0] - g[0], dy = x[1] - g[1];
    const double dist = fmax(sqrt(dx * dx + dy * dy) - 0.05, 0.05);
    const double th_d = atan2(g[1] - x[1], g[0] - x[0]);
    const double err = wrap_floor(th_d - x[2]);
    const double delta = clipd(t.k_omega * err, -t.delta_max, t.delta_max);
    const double beta = atan((p.rear_ax_dist / t.wheel_base) * tan(delta));       // :55-59
    const double hs = fmax(0.0, cos(err));
    const double v = clipd(t.k_v * dist * hs, p.v_min, p.v_max);
    u[0] = t.k_a * (v - x[3]);
    u[1] = beta;
  }
  static SCB_HD void stop(const scb_params&, const scb_track&, const double*, double* u) { u[0] = 0.0; u[1] = 0.0; }   // :149-150
  static SCB_HD bool has_stopped(const double* x) { return fabs(x[3]) < 0.05; }                         // :152-153
  static SCB_HD void rotate_to(const scb_params&, const scb_track&, const double* x, double, double th, double* u,
                               double&) {                                                              // :155-158
    u[0] = 0.0; u[1] = 2.0 * wrap_floor(th - x[2]);
  }
  static SCB_HD void step(const scb_params& p, double* x, const double* u) {                            // :112-123, f :75-91, g :93-110
    double s, c; sincos_pair(x[2], s, c);
    const double v = x[3];
    x[0] = x[0] + (v * c + (-v * s) * u[1]) * p.dt;
    x[1] = x[1] + (v * s + (v * c) * u[1]) * p.dt;
    x[2] = wrap_floor(x[2] + ((v / p.rear_ax_dist) * u[1]) * p.dt);
    x[3] = clipd(v + u[0] * p.dt, p.v_min, p.v_max);
  }
};

template <>
struct ModelLoop<SCB_KINEMATIC_BICYCLE_2D_C3BF> : ModelLoop<SCB_KINEMATIC_BICYCLE_2D> {};
template <>
struct ModelLoop<SCB_KINEMATIC_BICYCLE_2D_DPCBF> : ModelLoop<SCB_KINEMATIC_BICYCLE_2D> {};

template <>
struct ModelLoop<SCB_QUAD_3D> {
  static constexpr int NX = 12, NU = 4, NPOS = 3;
  static constexpr bool HAS_ATT = false;
  static SCB_HD double half_angle() { return kPi; }
  static SCB_HD double yaw_of(const double* x, double) { return x[5]; }       // robots/robot.py:450-451
  // u = pinv(B2) w with B2 = [[1,1,1,1],[0,L,0,-L],[L,0,-L,0],[nu,-nu,nu,-nu]] (quad3D.py:70-79): B2 is invertible,
  // so pinv(B2) = B2^-1, written out
  static SCB_HD void wrench_to_u(const scb_params& p, double F, double ty, double tx, double tz, double* u) {
    const double f4 = 0.25 * F, a = tx / (2.0 * p.arm_L), b = ty / (2.0 * p.arm_L), c = tz / (4.0 * p.nu_coef);
    u[0] = f4 + a + c; u[1] = f4 + b - c; u[2] = f4 - a + c; u[3] = f4 - b - c;
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = clipd(u[i], p.u_lb[i], p.u_ub[i]);
  }
  static SCB_HD void nominal(const scb_params& p, const scb_track&, const double* x, const double* g, double* u) {   // :160-206
    const double kp = 1.0, kd = 2.0, ka = 5.0;
    const double ax = kp * (g[0] - x[0]) + kd * (-x[6]);
    const double ay = kp * (g[1] - x[1]) + kd * (-x[7]);
    const double az = kp * (g[2] - x[2]) + kd * (-x[8]);
    const double th_des = ax / p.gravity, ph_des = -ay / p.gravity, F = p.mass * az;
    const double ty = p.Iy * (ka * (th_des - x[3]) + kd * (-x[9]));
    const double tx = p.Ix * (ka * (ph_des - x[4]) + kd * (-x[10]));
    const double tz = p.Iz * (ka * (0.0 - x[5]) + kd * (-x[11]));
    wrench_to_u(p, F, ty, tx, tz, u);
  }
  static SCB_HD void stop(const scb_params& p, const scb_track&, const double* x, double* u) {          // :208-236
    const double k = 1.0;
    const double ax = -k * x[6], ay = -k * x[7], az = -k * x[8];
    const double th_des = ax / p.gravity, ph_des = -ay / p.gravity, F = p.mass * az;
    const double ty = p.Iy * k * (th_des - x[3] - x[9] / k);
    const double tx = p.Ix * k * (ph_des - x[4] - x[10] / k);
    const double tz = p.Iz * k * (0.0 - x[5] - x[11] / k);
    wrench_to_u(p, F, ty, tx, tz, u);
  }
  static SCB_HD bool has_stopped(const double* x) {                                                    // :238-242
    const double lv = sqrt(x[6] * x[6] + x[7] * x[7] + x[8] * x[8]);
    const double av = sqrt(x[9] * x[9] + x[10] * x[10] + x[11] * x[11]);
    return lv < 0.05 && av < 0.05;
  }
  static SCB_HD void rotate_to(const scb_params& p, const scb_track&, const double* x, double, double th, double* u,
                               double&) {                                                              // :244-268
    const double k = 2.0;
    const double F = p.mass * p.gravity;
    const double ty = p.Iy * k * (0.0 - x[3] - x[9] / k);
    const double tx = p.Ix * k * (0.0 - x[4] - x[10] / k);
    const double tz = p.Iz * k * (th - x[5] - x[11] / k);
    wrench_to_u(p, F, ty, tx, tz, u);
  }
  // xdot = A x + B u (quad3D.py:81-97)
  static SCB_HD void deriv(const scb_params& p, const double* x, const double* u, double* d) {
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = x[6 + i];
    d[6] = p.gravity * x[3];
    d[7] = -p.gravity * x[4];
    const double L = p.arm_L, nu = p.nu_coef;
    d[8] = (1.0 / p.mass) * (u[0] + u[1] + u[2] + u[3]);
    d[9] = (1.0 / p.Iy) * (L * u[1] - L * u[3]);
    d[10] = (1.0 / p.Ix) * (L * u[0] - L * u[2]);
    d[11] = (1.0 / p.Iz) * (nu * u[0] - nu * u[1] + nu * u[2] - nu * u[3]);
  }
  static SCB_HD void step(const scb_params& p, double* x, const double* u) {                            // RK4 + wrap, :121-158
    double k1[12], k2[12], k3[12], k4[12], t[12];
    const double dt = p.dt;
    deriv(p, x, u, k1);
#pragma unroll
    for (int i = 0; i < 12; ++i) t[i] = x[i] + (dt / 2.0) * k1[i];
    deriv(p, t, u, k2);
#pragma unroll
    for (int i = 0; i < 12; ++i) t[i] = x[i] + (dt / 2.0) * k2[i];
    deriv(p, t, u, k3);
#pragma unroll
    for (int i = 0; i < 12; ++i) t[i] = x[i] + dt * k3[i];
    deriv(p, t, u, k4);
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = x[i] + (dt / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    x[3] = wrap_floor(x[3]); x[4] = wrap_floor(x[4]); x[5] = wrap_floor(x[5]);
  }
};

// ------------------------------------------------------------------------------------------
// get_nearest_unpassed_obs (tracking.py:345-403) for one agent.  `keys` = K doubles of scratch
// private to the group (shared memory on the device).  Returns nobs (-1: no obstacles at all).
template <int LANES>
SCB_HD int select_agent(int K, int M, const double* scene, double px, double py, double yaw, double half,
                        double* keys, double* OBS, int32_t* idx) {
  using G = Grp<LANES>;
  const int lane = G::lane();
  if (K <= 0) {
    for (int r = lane; r < M; r += LANES) {
      OBS[r * 7 + 0] = 1000.0; OBS[r * 7 + 1] = 1000.0;
#pragma unroll
      for (int q = 2; q < 7; ++q) OBS[r * 7 + q] = 0.0;
      if (idx) idx[r] = -1;
    }
    return -1;                                                // all_obs empty -> None (:371-372)
  }
  // pass 1: distance + "unpassed" test (:381-391); a non-kept obstacle is stored with the sign bit set (exact)
  uint32_t any = 0u;
  for (int j = lane; j < K; j += LANES) {
    const double ox = scene[j * 7], oy = scene[j * 7 + 1];
    const double vx = ox - px, vy = oy - py;
    const double d = sqrt(vx * vx + vy * vy);
    const double diff = fabs(wrap_floor(atan2(vy, vx) - yaw));
    const bool keep = diff <= half;
    any |= keep ? 1u : 0u;
    keys[j] = keep ? d : -d;
  }
  any = G::or_reduce(any);
#if defined(__CUDA_ARCH__)
  if (LANES > 1) __syncwarp(G::gmask());
#endif
  // pass 2: rank among the candidates (kept ones, or all when nothing was kept, :393-397); ties -> lower index
  // (np.argsort on distinct floats; equal distances have measure zero and the QP does not depend on row order)
  int cnt = 0;
  for (int j = lane; j < K; j += LANES) {
    cnt += (!any || !signbit(keys[j])) ? 1 : 0;
  }
  cnt = (int)G::sum((double)cnt);
  const int take = cnt < M ? cnt : M;
  for (int j = lane; j < K; j += LANES) {
    const double kraw = keys[j];
    if (any && signbit(kraw)) continue;
    const double kj = fabs(kraw);
    int rank = 0;
    for (int i = 0; i < K; ++i) {
      const double r_ = keys[i];
      if (any && signbit(r_)) continue;
      const double ki = fabs(r_);
      rank += (ki < kj || (ki == kj && i < j)) ? 1 : 0;
    }
    if (rank < M) {
#pragma unroll
      for (int q = 0; q < 7; ++q) OBS[rank * 7 + q] = scene[j * 7 + q];
      if (idx) idx[rank] = j;
    }
  }
  for (int r = take + lane; r < M; r += LANES) {              // pad (mpc_cbf.py:346 dummy row)
    OBS[r * 7 + 0] = 1000.0; OBS[r * 7 + 1] = 1000.0;
#pragma unroll
    for (int q = 2; q < 7; ++q) OBS[r * 7 + q] = 0.0;
    if (idx) idx[r] = -1;
  }
  return take;
}

// is_collide_unknown over the known obstacles (tracking.py:445-495)
template <int LANES>
SCB_HD bool collides(int K, const double* scene, double px, double py, double radius) {
  using G = Grp<LANES>;
  uint32_t hit = 0u;
  for (int j = G::lane(); j < K; j += LANES) {
    const double* o = scene + j * 7;
    const double flag = o[6];
    // _known_obs_geometry (:405-420): superellipsoid iff isclose(flag, 1) and e >= 2, else circle
    const bool se = (fabs(flag - 1.0) <= 1e-8 + 1e-5) && !(fabs(flag) <= 1e-8) && o[4] >= 2.0;
    if (!se) {
      const double dx = px - o[0], dy = py - o[1];
      if (sqrt(dx * dx + dy * dy) < (o[2] + radius)) hit = 1u;
    } else {
      double st, ct; sincos_pair(o[5], st, ct);
      const double xp = ct * (px - o[0]) + st * (py - o[1]);
      const double yp = -st * (px - o[0]) + ct * (py - o[1]);
      const double h = pow(xp / (o[2] + radius), o[4]) + pow(yp / (o[3] + radius), o[4]) - 1.0;
      if (h <= 0.0) hit = 1u;
    }
  }
  return G::or_reduce(hit) != 0u;
}

// update_goal (tracking.py:497-535).  Returns has_goal; goal[0..NPOS) written when true.
template <int MODEL>
SCB_HD bool update_goal(const scb_track& t, const double* x, double yaw, const double* wp, int nwp, int& sm,
                        int& wi, double& u_att, double* goal) {
  using ML = ModelLoop<MODEL>;
  if (sm == SCB_SM_ROTATE && wi < nwp) {
    const double* rg = wp + (size_t)wi * 3;
    const double goal_angle = atan2(rg[1] - x[1], rg[0] - x[0]);
    if (MODEL == SCB_QUAD_2D) sm = SCB_SM_TRACK;              // "Those skip 'rotate' state" (:512-513)
    if (!t.enable_rotation) sm = SCB_SM_TRACK;
    if (fabs(yaw - goal_angle) > t.rotation_threshold) {
#pragma unroll
      for (int i = 0; i < ML::NPOS; ++i) goal[i] = rg[i];
      return true;
    }
    sm = SCB_SM_TRACK;
    u_att = nan("");                                          // self.u_att = None (:519)
  }
  if (wi >= nwp) return false;
  {
    const double* g = wp + (size_t)wi * 3;
    const double dx = x[0] - g[0], dy = x[1] - g[1];
    if (sqrt(dx * dx + dy * dy) < t.reached_threshold) {      // goal_reached (:262-267)
      wi += 1;
      if (wi >= nwp) { sm = SCB_SM_IDLE; return false; }
    }
  }
  const double* g = wp + (size_t)wi * 3;
#pragma unroll
  for (int i = 0; i < ML::NPOS; ++i) goal[i] = g[i];
  return true;
}

// ------------------------------------------------------------------------------------------
// Everything before the solve.  All lanes of the group call this; lane 0 stores the scalars.
// `scene` = the obstacle array to select from (t.SCENE, or the fused kernel's shared-memory copy of it).
template <int MODEL, int LANES>
SCB_HD void track_pre_agent(const scb_params& p, const scb_track& t, long a, double* keys, const double* scene) {
  using ML = ModelLoop<MODEL>;
  using G = Grp<LANES>;
  constexpr int NX = ML::NX, NU = ML::NU;
  const int lane = G::lane();
  if (t.done[a]) {                                            // frozen: run_all_steps' loop has broken for this agent
    if (lane == 0 && t.track_flag) t.track_flag[a] = -1;      // (the MPC kernel skips it: outputs stay those of its last step)
    return;
  }
  double x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = t.X[a * NX + i];
  const double yaw = ML::yaw_of(x, t.yaw[a]);
  int sm = t.sm[a], wi = t.wp_idx[a];
  const int nwp = t.nwp[a];
  const double* wp = t.WP + (size_t)a * t.W * 3;
  double u_att = t.u_att[a];
  double goal[3] = {0.0, 0.0, 0.0};
  bool has_goal = t.has_goal[a] != 0;
  if (has_goal) {
#pragma unroll
    for (int i = 0; i < ML::NPOS; ++i) goal[i] = t.goal[a * ML::NPOS + i];
  }
#if defined(__CUDA_ARCH__)
  if (LANES > 1) __syncwarp(G::gmask());                      // every lane has read the state before lane 0 rewrites it
#endif

  // state machine (tracking.py:569-578)
  if (sm == SCB_SM_STOP) {
    if (ML::has_stopped(x)) {
      sm = t.enable_rotation ? SCB_SM_ROTATE : SCB_SM_TRACK;
      has_goal = update_goal<MODEL>(t, x, yaw, wp, nwp, sm, wi, u_att, goal);
    }
  } else {
    has_goal = update_goal<MODEL>(t, x, yaw, wp, nwp, sm, wi, u_att, goal);
  }

  // obstacle selection (tracking.py:583)
  const int no = select_agent<LANES>(t.K, t.M, scene, x[0], x[1], yaw, ML::half_angle(), keys,
                                     t.OBS + (size_t)a * t.M * 7, nullptr);

  // nominal input (tracking.py:589-604)
  double u[NU];
  if (sm == SCB_SM_ROTATE && has_goal) {
    const double goal_angle = atan2(goal[1] - x[1], goal[0] - x[0]);
    ML::rotate_to(p, t, x, yaw, goal_angle, u, u_att);
  } else if (!has_goal) {
    ML::stop(p, t, x, u);
  } else {
    ML::nominal(p, t, x, goal, u);
  }

  if (lane == 0) {
    t.sm[a] = sm; t.wp_idx[a] = wi; t.u_att[a] = u_att;
    t.has_goal[a] = has_goal ? 1 : 0;
#pragma unroll
    for (int i = 0; i < ML::NPOS; ++i) t.goal[a * ML::NPOS + i] = goal[i];
#pragma unroll
    for (int i = 0; i < NU; ++i) t.Uref[a * NU + i] = u[i];
    t.nobs[a] = no;
    if (t.track_flag) t.track_flag[a] = (sm == SCB_SM_TRACK) ? 1 : 0;
  }
}

// Everything after the solve.
template <int MODEL, int LANES>
SCB_HD void track_post_agent(const scb_params& p, const scb_track& t, long a, const double* scene) {
  using ML = ModelLoop<MODEL>;
  using G = Grp<LANES>;
  constexpr int NX = ML::NX, NU = ML::NU;
  if (t.done[a]) return;
  const int lane = G::lane();
  double x[NX], u[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = t.X[a * NX + i];
#pragma unroll
  for (int i = 0; i < NU; ++i) u[i] = t.U[a * NU + i];
  double yaw = t.yaw[a];
  double u_att = t.u_att[a];
  const int sm = t.sm[a];
  const bool has_goal = t.has_goal[a] != 0;
  // the reference's MPCCBF.status is hard-wired 'optimal' (mpc_cbf.py:10,400): by default only the QP controllers can
  // fail here.  Our interior-point solver is not IPOPT, so the per-agent MPC status stays visible: `mpc_fail` counts the
  // control steps whose solve did not end SCB_OPTIMAL, and with `mpc_strict` such a step returns -2 like a failed QP.
  const bool is_mpc = (t.controller == SCB_CTRL_MPC_CBF);
  const bool solved = (t.status[a] == SCB_OPTIMAL);
  const bool ok = is_mpc ? (t.mpc_strict ? solved : true) : solved;
#if defined(__CUDA_ARCH__)
  if (LANES > 1) __syncwarp(G::gmask());
#endif

  // attitude controller, integrators only (tracking.py:621-624; velocity_tracking_yaw.py:35-62)
  if (ML::HAS_ATT && sm == SCB_SM_TRACK && t.att_velocity_tracking && t.enable_rotation) {
    double vx = 0.0, vy = 0.0;
    if constexpr (ML::HAS_ATT) ML::att_velocity(x, u, vx, vy);
    const double speed = hypot(vx, vy);
    if (speed < 1e-2) u_att = 0.0;
    else u_att = clipd(t.att_kp * wrap_floor(atan2(vy, vx) - yaw), -t.w_max, t.w_max);
  }

  int ret;
  bool collide = collides<LANES>(t.K, scene, x[0], x[1], p.radius);
  if (!ok || collide) {
    ret = -2;                                                 // :627-634 (no step)
  } else {
    ML::step(p, x, u);                                        // :637
    if (ML::HAS_ATT) { if (!(u_att != u_att)) yaw = wrap_floor(yaw + u_att * p.dt); }      // step_rotate, robots/robot.py:446-448
    else yaw = ML::yaw_of(x, yaw);
    collide = collides<LANES>(t.K, scene, x[0], x[1], p.radius);
    if (collide) ret = -2;                                    // :640-646
    else ret = (!has_goal && sm != SCB_SM_STOP) ? -1 : 0;     // :666-668
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) t.X[a * NX + i] = x[i];
    t.yaw[a] = yaw;
    t.u_att[a] = u_att;
    t.ret[a] = ret;
    t.nsteps[a] += 1;
    if (ret != 0) t.done[a] = 1;
    if (is_mpc && !solved && t.mpc_fail) t.mpc_fail[a] += 1;
    if (is_mpc && sm == SCB_SM_TRACK && ok) {
#pragma unroll
      for (int i = 0; i < NU; ++i) t.u_prev[a * NU + i] = u[i];
    }
  }
}

}  // namespace scb
