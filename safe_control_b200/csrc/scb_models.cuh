// scb_models.cuh -- continuous-time CBF row assembly, one (agent, obstacle) pair per call.
//
// Each ModelCT<MODEL>::row() produces the row (a[0..nu), b) of the reference's
//     A1[row] u + b1[row] >= 0                         (position_control/cbf_qp.py:152-183)
// i.e. rel-degree 1:  a = dh/dx g,      b = dh/dx f + alpha h            (:164-165)
//      rel-degree 2:  a = d(hdot)/dx g, b = d(hdot)/dx f + (a1+a2) hdot + a1 a2 h   (:180-183)
//      cbf_mode 'hard': b = h/dt + dh/dx f   or   h/dt^2 + 2 hdot/dt + d(hdot)/dx f   (:161,177)
// from the model's f, g and agent_barrier:
//   SingleIntegrator2D        robots/single_integrator2D.py:44-62, 114-146
//   DynamicUnicycle2D         robots/dynamic_unicycle2D.py:42-73, 121-186
//   KinematicBicycle2D        robots/kinematic_bicycle2D.py:75-110, 160-173
//   KinematicBicycle2D_C3BF   dynamic_env/kinematic_bicycle2D_c3bf.py:15-75
//   KinematicBicycle2D_DPCBF  dynamic_env/kinematic_bicycle2D_dpcbf.py:16-84
//   DoubleIntegrator2D        robots/double_integrator2D.py:46-79, 167-222
//   Quad2D                    robots/quad2D.py:46-82, 166-177
//   Unicycle2D                robots/unicycle2D.py:43-63, 100-125
// The arithmetic follows the reference's operation order where that is cheap, so rows
// agree with the numpy path to a few ulp.  Quad3D has no continuous barrier
// (quad3D.py:269-273) and is rejected on the host.
#pragma once

#include "scb_core.cuh"

namespace scb {

// per-agent quantities shared by all of the agent's rows
struct AgentCT {
  double px, py;     // position
  double c, s, v;    // cos(theta), sin(theta), speed (models with heading)
  double fx, fy;     // f(x)[0:2] = v c, v s
  double th;         // heading itself (models whose rows need it)
};

struct RowOut {
  double a[4];
  double b;
};

template <int MODEL>
struct ModelCT;

// Superellipsoid obstacles (flag == 1) are the rare, pow()-heavy branch: keep them out of line so
// the unrolled per-lane row loops of the fused kernels carry one call, not RPL inlined copies.
// h, dh/dx, dh/dy and (optionally) the second derivatives hxx, hxy, hyy at position (px, py).
// Everything is passed and returned BY VALUE so the callers' obstacle row stays in registers.
struct SEOut { double h, gx, gy, hxx, hxy, hyy; };
SCB_HD_NOINLINE SEOut superellipsoid_terms(double px, double py, double ox, double oy, double a, double b, double e,
                                           double th, double radius) {
  SEOut r;
  double st, ct; sincos_pair(th, st, ct);
  const double xp = ct * (px - ox) + st * (py - oy);
  const double yp = -st * (px - ox) + ct * (py - oy);
  const double ar = a + radius, br = b + radius;
  const double ae = pow(ar, e), be = pow(br, e);
  r.h = pow(xp / ar, e) + pow(yp / br, e) - 1.0;
  const double ga = e * pow(xp, e - 1.0), gb = e * pow(yp, e - 1.0);
  r.gx = ga * (ct / ae) + gb * (-st / be);
  r.gy = ga * (st / ae) + gb * (ct / be);
  const double ka = (e * (e - 1.0) / ae) * pow(xp, e - 2.0), kb = (e * (e - 1.0) / be) * pow(yp, e - 2.0);
  r.hxx = ka * ct * ct + kb * st * st;
  r.hxy = (ka - kb) * ct * st;
  r.hyy = ka * st * st + kb * ct * ct;
  return r;
}

// ---------------------------------------------------------------------------------------
template <>
struct ModelCT<SCB_SINGLE_INTEGRATOR_2D> {
  static constexpr int NX = 2, NU = 2;
  static SCB_HD void prep(const scb_params&, const double* x, AgentCT& g) {
    g.px = x[0]; g.py = x[1]; g.c = 1.0; g.s = 0.0; g.v = 0.0; g.fx = 0.0; g.fy = 0.0;
  }
  static SCB_HD void row(const scb_params& p, const AgentCT& g, const double* o, RowOut& r) {
    double h = 0.0, d0 = 0.0, d1 = 0.0;
    const double flag = o[6];
    if (flag == 0.0) {                                        // single_integrator2D.py:120-127
      const double dx = g.px - o[0], dy = g.py - o[1];
      const double dmin = o[2] + p.radius;
      h = (dx * dx + dy * dy) - 1.01 * (dmin * dmin);
      d0 = 2.0 * dx; d1 = 2.0 * dy;
    } else if (flag == 1.0) {                                 // :128-143
      const SEOut se = superellipsoid_terms(g.px, g.py, o[0], o[1], o[2], o[3], o[4], o[5], p.radius);
      h = se.h; d0 = se.gx; d1 = se.gy;
    }
    r.a[0] = d0; r.a[1] = d1;                                 // g = I, f = 0
    r.b = (p.cbf_mode == 1) ? h / p.dt : p.alpha * h;
  }
};

// ---------------------------------------------------------------------------------------
// Unicycle2D (robots/unicycle2D.py): X = [x, y, theta], U = [v, omega], f = 0, g = [[c, 0], [s, 0], [0, 1]].
// h = |p - o|^2 - beta d^2 - sigma(s),  s = (p - o) . (c, s)   (:107-125): the sigma term is what gives omega a
// non-zero coefficient (relative degree 1).  Circle obstacles only; the flag column is never read.
// (The reference's own call path hands this model obstacle ROWS, on which its `obs[2][0]` raises IndexError --
//  tracking.py:611-616 -> cbf_qp.py:156 -> unicycle2D.py:109; implemented is the intended column semantics, the
//  fixtures in tests/golden/ref_*3.npz were generated by feeding the reference columns.)
template <>
struct ModelCT<SCB_UNICYCLE_2D> {
  static constexpr int NX = 3, NU = 2;
  static SCB_HD void prep(const scb_params&, const double* x, AgentCT& g) {
    g.px = x[0]; g.py = x[1]; g.th = x[2];
    sincos_pair(x[2], g.s, g.c);
    g.v = 0.0; g.fx = 0.0; g.fy = 0.0;
  }
  static SCB_HD void row(const scb_params& p, const AgentCT& g, const double* o, RowOut& r) {
    const double k1 = 0.5, k2 = 1.8;                          // :36-37
    const double dx = g.px - o[0], dy = g.py - o[1];
    const double dmin = o[2] + p.radius;
    const double sv = dx * g.c + dy * g.s;
    const double ex = exp(k1 - sv);
    const double sig = k2 * (ex - 1.0) / (ex + 1.0);          // sigma(s)      :100-102
    const double dsig = -k2 * ex / (1.0 + ex) * (1.0 - sig / k2);   // sigma'(s)  :104-105
    const double h = ((dx * dx + dy * dy) - 1.01 * (dmin * dmin)) - sig;
    const double d0 = 2.0 * dx - dsig * g.c, d1 = 2.0 * dy - dsig * g.s;
    const double d2 = -dsig * (-g.s * dx + g.c * dy);
    r.a[0] = d0 * g.c + d1 * g.s;                             // dh/dx g
    r.a[1] = d2;
    r.b = (p.cbf_mode == 1) ? h / p.dt : p.alpha * h;         // f = 0
  }
};

// ---------------------------------------------------------------------------------------
// shared by DU and KB: h, hdot, d(hdot)/dx for the distance barrier (circle), beta differs
SCB_HD void hocbf_circle(const AgentCT& g, const double* o, double radius, double beta,
                         double& h, double& hd, double* dhd) {
  const double dx = g.px - o[0], dy = g.py - o[1];
  const double dmin = o[2] + radius;
  h = (dx * dx + dy * dy) - beta * (dmin * dmin);
  hd = 2.0 * dx * g.fx + 2.0 * dy * g.fy;
  dhd[0] = 2.0 * g.fx;
  dhd[1] = 2.0 * g.fy;
  dhd[2] = 2.0 * dx * (-g.v * g.s) + 2.0 * dy * (g.v * g.c);
  dhd[3] = 2.0 * dx * g.c + 2.0 * dy * g.s;
}

SCB_HD double rel2_b(const scb_params& p, double h, double hd, double lf) {
  if (p.cbf_mode == 1) return h / (p.dt * p.dt) + 2.0 * hd / p.dt + lf;
  return lf + (p.alpha1 + p.alpha2) * hd + (p.alpha1 * p.alpha2) * h;
}

template <>
struct ModelCT<SCB_DYNAMIC_UNICYCLE_2D> {
  static constexpr int NX = 4, NU = 2;
  static SCB_HD void prep(const scb_params&, const double* x, AgentCT& g) {
    g.px = x[0]; g.py = x[1]; g.v = x[3];
    sincos_pair(x[2], g.s, g.c);
    g.fx = g.v * g.c; g.fy = g.v * g.s;
  }
  // h, hdot, d(hdot)/dx by obstacle flag; anything else -> vacuous zeros
  static SCB_HD void barrier(const scb_params& p, const AgentCT& g, const double* o, double& h, double& hd, double* dhd) {
    h = 0.0; hd = 0.0; dhd[0] = dhd[1] = dhd[2] = dhd[3] = 0.0;
    const double flag = o[6];
    if (flag == 0.0) {                                        // dynamic_unicycle2D.py:136-146
      hocbf_circle(g, o, p.radius, 1.01, h, hd, dhd);
    } else if (flag == 1.0) {                                 // :148-183
      const SEOut se = superellipsoid_terms(g.px, g.py, o[0], o[1], o[2], o[3], o[4], o[5], p.radius);
      h = se.h;
      hd = se.gx * g.fx + se.gy * g.fy;
      dhd[0] = se.hxx * g.fx + se.hxy * g.fy;
      dhd[1] = se.hxy * g.fx + se.hyy * g.fy;
      dhd[2] = se.gx * (-g.v * g.s) + se.gy * (g.v * g.c);
      dhd[3] = se.gx * g.c + se.gy * g.s;
    }
  }
  static SCB_HD void row(const scb_params& p, const AgentCT& g, const double* o, RowOut& r) {
    double h, hd, dhd[4];
    barrier(p, g, o, h, hd, dhd);
    // g = [[0,0],[0,0],[0,1],[1,0]]  (:64-73)
    r.a[0] = dhd[3];
    r.a[1] = dhd[2];
    const double lf = dhd[0] * g.fx + dhd[1] * g.fy;
    r.b = rel2_b(p, h, hd, lf);
  }
};

template <>
struct ModelCT<SCB_KINEMATIC_BICYCLE_2D> {
  static constexpr int NX = 4, NU = 2;
  static SCB_HD void prep(const scb_params& p, const double* x, AgentCT& g) {
    ModelCT<SCB_DYNAMIC_UNICYCLE_2D>::prep(p, x, g);
  }
  static SCB_HD void row(const scb_params& p, const AgentCT& g, const double* o, RowOut& r) {
    double h, hd, dhd[4];
    hocbf_circle(g, o, p.radius, 1.1, h, hd, dhd);            // kinematic_bicycle2D.py:160-173 (flag ignored)
    // g = [[0,-v s],[0, v c],[0, v/L_r],[1, 0]]  (:93-110)
    r.a[0] = dhd[3];
    r.a[1] = dhd[0] * (-g.v * g.s) + dhd[1] * (g.v * g.c) + dhd[2] * (g.v / p.rear_ax_dist);
    const double lf = dhd[0] * g.fx + dhd[1] * g.fy;
    r.b = rel2_b(p, h, hd, lf);
  }
};

template <>
struct ModelCT<SCB_KINEMATIC_BICYCLE_2D_C3BF> {
  static constexpr int NX = 4, NU = 2;
  static SCB_HD void prep(const scb_params& p, const double* x, AgentCT& g) {
    ModelCT<SCB_DYNAMIC_UNICYCLE_2D>::prep(p, x, g);
  }
  // h and the reference's hand-written dh/dx incl. its +eps terms (kinematic_bicycle2D_c3bf.py:43-73)
  static SCB_HD void barrier(const scb_params& p, const AgentCT& g, const double* o, double& h, double* dh) {
    const double ovx = o[3], ovy = o[4];
    const double ego = (o[2] + p.radius) * 1.0;
    const double prx = o[0] - g.px, pry = o[1] - g.py;
    const double vrx = ovx - g.v * g.c, vry = ovy - g.v * g.s;
    const double pm = sqrt(prx * prx + pry * pry);
    const double vm = sqrt(vrx * vrx + vry * vry);
    const double eps = 1e-6;
    const double sq = sqrt(fmax(pm * pm - ego * ego, eps));
    const double cosphi = sq / (pm + eps);
    h = (prx * vrx + pry * vry) + pm * vm * cosphi;
    dh[0] = -vrx - vm * prx / (sq + eps);
    dh[1] = -vry - vm * pry / (sq + eps);
    dh[2] = g.v * g.s * prx - g.v * g.c * pry + (sq + eps) / vm * (g.v * (ovx * g.s - ovy * g.c));
    dh[3] = -g.c * prx - g.s * pry + (sq + eps) / vm * (g.v - (ovx * g.c + ovy * g.s));
  }
  static SCB_HD void row(const scb_params& p, const AgentCT& g, const double* o, RowOut& r) {
    double h, dh[4];
    barrier(p, g, o, h, dh);
    r.a[0] = dh[3];
    r.a[1] = dh[0] * (-g.v * g.s) + dh[1] * (g.v * g.c) + dh[2] * (g.v / p.rear_ax_dist);
    const double lf = dh[0] * g.fx + dh[1] * g.fy;
    r.b = (p.cbf_mode == 1) ? h / p.dt + lf : lf + p.alpha * h;
  }
};

// ---------------------------------------------------------------------------------------
// Dynamic-parabolic CBF.  The reference's hand-written dh/dx is NOT the gradient of its h (it drops the
// sqrt(s^2-1)/ego_dim factor): transcribed literally, never differentiated.
template <>
struct ModelCT<SCB_KINEMATIC_BICYCLE_2D_DPCBF> {
  static constexpr int NX = 4, NU = 2;
  static SCB_HD void prep(const scb_params& p, const double* x, AgentCT& g) {
    ModelCT<SCB_DYNAMIC_UNICYCLE_2D>::prep(p, x, g);
    g.th = x[2];
  }
  static SCB_HD void barrier(const scb_params& p, const AgentCT& g, const double* o, double& h, double* dh) {
    const double k_lambda = 0.1, k_mu = 0.5, sm = 1.05;                        // ctor defaults (:11), s (:16)
    const double ovx = o[3], ovy = o[4];
    const double ego = (o[2] + p.radius) * sm;
    const double prx = o[0] - g.px, pry = o[1] - g.py;
    const double vrx = ovx - g.v * g.c, vry = ovy - g.v * g.s;
    const double pm = sqrt(prx * prx + pry * pry);
    const double vm = sqrt(vrx * vrx + vry * vry);
    const double rot = atan2(pry, prx);
    double sr, cr; sincos_pair(rot, sr, cr);
    const double vnx = cr * vrx + sr * vry;                                     // R @ v_rel (:57-63)
    const double vny = -sr * vrx + cr * vry;
    const double eps = 1e-6;
    const double dsafe = fmax(pm * pm - ego * ego, eps);
    const double sd = sqrt(dsafe);
    const double k2 = sqrt(sm * sm - 1.0) / ego;
    const double lam = k_lambda * sd / vm * k2;
    const double mu = k_mu * sd * k2;
    h = vnx + lam * (vny * vny) + mu;
    const double pm2 = pm * pm;
    double srt, crt; sincos_pair(rot - g.th, srt, crt);
    dh[0] = pry * vny / pm2 - k_lambda * prx * (vny * vny) / vm / sd - 2.0 * k_lambda * sd / vm * vny * pry / pm2 * vnx - k_mu * prx / sd;
    dh[1] = -prx * vny / pm2 - k_lambda * pry * (vny * vny) / vm / sd + 2.0 * k_lambda * sd / vm * vny * prx / pm2 * vnx - k_mu * pry / sd;
    dh[2] = -g.v * srt - k_lambda * sd * g.v * (ovx * g.s - ovy * g.c) * (vny * vny) / (vm * vm * vm) - 2.0 * k_lambda * sd * vny * g.v * crt / vm;
    dh[3] = -crt - k_lambda * sd / (vm * vm * vm) * (g.v - ovx * g.c - ovy * g.s) * (vny * vny) - 2.0 * k_lambda * sd * vny * srt / vm;
  }
  static SCB_HD void row(const scb_params& p, const AgentCT& g, const double* o, RowOut& r) {
    double h, dh[4];
    barrier(p, g, o, h, dh);
    r.a[0] = dh[3];
    r.a[1] = dh[0] * (-g.v * g.s) + dh[1] * (g.v * g.c) + dh[2] * (g.v / p.rear_ax_dist);
    const double lf = dh[0] * g.fx + dh[1] * g.fy;
    r.b = (p.cbf_mode == 1) ? h / p.dt + lf : lf + p.alpha * h;
  }
};

// ---------------------------------------------------------------------------------------
// DoubleIntegrator2D: X = [x, y, vx, vy], f = [vx, vy, 0, 0], g = [[0,0],[0,0],[1,0],[0,1]]
template <>
struct ModelCT<SCB_DOUBLE_INTEGRATOR_2D> {
  static constexpr int NX = 4, NU = 2;
  static SCB_HD void prep(const scb_params&, const double* x, AgentCT& g) {
    g.px = x[0]; g.py = x[1]; g.c = 1.0; g.s = 0.0; g.v = 0.0; g.th = 0.0;
    g.fx = x[2]; g.fy = x[3];
  }
  static SCB_HD void barrier(const scb_params& p, const AgentCT& g, const double* o, double& h, double& hd, double* dhd) {
    h = 0.0; hd = 0.0; dhd[0] = dhd[1] = dhd[2] = dhd[3] = 0.0;
    const double flag = o[6];
    if (flag == 0.0) {                                        // double_integrator2D.py:172-184
      const double dx = g.px - o[0], dy = g.py - o[1];
      const double dmin = o[2] + p.radius;
      h = (dx * dx + dy * dy) - 1.01 * (dmin * dmin);
      hd = 2.0 * dx * g.fx + 2.0 * dy * g.fy;
      dhd[0] = 2.0 * g.fx; dhd[1] = 2.0 * g.fy; dhd[2] = 2.0 * dx; dhd[3] = 2.0 * dy;
    } else if (flag == 1.0) {                                 // :185-220
      const SEOut se = superellipsoid_terms(g.px, g.py, o[0], o[1], o[2], o[3], o[4], o[5], p.radius);
      h = se.h;
      hd = se.gx * g.fx + se.gy * g.fy;
      dhd[0] = se.hxx * g.fx + se.hxy * g.fy;
      dhd[1] = se.hxy * g.fx + se.hyy * g.fy;
      dhd[2] = se.gx; dhd[3] = se.gy;
    }
  }
  static SCB_HD void row(const scb_params& p, const AgentCT& g, const double* o, RowOut& r) {
    double h, hd, dhd[4];
    barrier(p, g, o, h, hd, dhd);
    r.a[0] = dhd[2]; r.a[1] = dhd[3];
    const double lf = dhd[0] * g.fx + dhd[1] * g.fy;
    r.b = rel2_b(p, h, hd, lf);
  }
};

// ---------------------------------------------------------------------------------------
// Quad2D: X = [x, z, theta, vx, vz, theta_dot], U = [f_right, f_left]; f = [vx, vz, theta_dot, 0, -g, 0],
// g columns = [0, 0, 0, -sin(theta)/m, cos(theta)/m, +-r/I] (quad2D.py:46-82).  Circle barrier on (x, z) only
// (flag ignored, :166-177): dh_dot/dx = [2 vx, 2 vz, 0, 2 dx, 2 dz, 0].
template <>
struct ModelCT<SCB_QUAD_2D> {
  static constexpr int NX = 6, NU = 2;
  static SCB_HD void prep(const scb_params&, const double* x, AgentCT& g) {
    g.px = x[0]; g.py = x[1]; g.th = x[2]; g.v = 0.0;
    sincos_pair(x[2], g.s, g.c);
    g.fx = x[3]; g.fy = x[4];
  }
  static SCB_HD void barrier(const scb_params& p, const AgentCT& g, const double* o, double& h, double& hd, double* dhd) {
    const double dx = g.px - o[0], dy = g.py - o[1];
    const double dmin = o[2] + p.radius;
    h = (dx * dx + dy * dy) - 1.01 * (dmin * dmin);
    hd = 2.0 * dx * g.fx + 2.0 * dy * g.fy;
    dhd[0] = 2.0 * g.fx; dhd[1] = 2.0 * g.fy; dhd[2] = 2.0 * dx; dhd[3] = 2.0 * dy;    // w.r.t. (x, z, vx, vz)
  }
  static SCB_HD void row(const scb_params& p, const AgentCT& g, const double* o, RowOut& r) {
    double h, hd, dhd[4];
    barrier(p, g, o, h, hd, dhd);
    const double a = dhd[2] * (-g.s / p.mass) + dhd[3] * (g.c / p.mass);
    r.a[0] = a; r.a[1] = a;
    const double lf = dhd[0] * g.fx + dhd[1] * g.fy + dhd[3] * (-p.gravity);
    r.b = rel2_b(p, h, hd, lf);
  }
};

}  // namespace scb
