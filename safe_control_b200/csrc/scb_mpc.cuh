// scb_mpc.cuh -- per-agent body of the MPC-CBF path (device + host-sim).
//
// Replaces MPCCBF.solve_control_problem (position_control/mpc_cbf.py:366-402), i.e. the NLP that
// do-mpc assembles (mpc_cbf.py:108-160, 162-259, 295-325) and IPOPT solves:
//
//   min  sum_{k<H} [(x_k-g)'Q(x_k-g) + sum_i R_i (u_k,i - u_{k-1,i})^2] + (x_H-g)'Q(x_H-g)
//   s.t. x_{k+1} = x_k + (f(x_k) + g(x_k) u_k) dt                      (plain Euler, :135-141)
//        cbf_j(x_k, u_k) >= 0, k < H, j < num_obs   built from the model's own step()   (:308-325)
//        u_lb <= u_k <= u_ub,  |x_k[3]| <= v_max                        (:183-221)
//
// Method: the states are eliminated by the rollout (single shooting, z = (u_0..u_{H-1})), and the
// resulting inequality-constrained NLP is solved by a primal-dual interior-point loop with slacks
// (g(z) - s = 0, s > 0), monotone barrier schedule, fraction-to-the-boundary rule, an l1-merit
// backtracking line search and diagonal regularisation when the reduced Hessian is not positive
// definite -- the same family of method as IPOPT, so it converges to a KKT point of the same NLP
// from the same cold start (x_k = x_init, u_k = u_prev, mpc_cbf.py:368-369).
//
// Exact second derivatives, obtained structurally instead of densely:
//   * per stage, second-order jets (scb_jet.cuh) of the Euler map F and of the barrier points
//     p1 = step(x,u), p2 = step(step(x,u),u) in the 6 stage variables y = (x, u);
//   * every circle barrier is  c_j = sum_i w_i |p_i - o_j|^2 - W beta d_j^2, hence its gradient and
//     Hessian are AFFINE in the obstacle centre:  grad c_j = gE - o_jx gX - o_jy gY,
//     hess c_j = KE - o_jx KX - o_jy KY, with (gE,KE), (gX,KX), (gY,KY) the jets of
//     E = sum w_i |p_i|^2, 2 sum w_i p_ix, 2 sum w_i p_iy.  The M obstacle rows of a stage therefore
//     enter the Newton system only through 12 weighted sums over j;
//   * the stage Hessians G_k (6x6) are condensed with the state sensitivities S_k = dx_k/dz into the
//     n x n reduced Hessian (n = H nu <= 32), factored by Cholesky.
// Work split in a lane group: lanes stride over stages / (stage, obstacle) pairs / matrix columns;
// all scratch lives in a per-agent workspace (shared memory on the device).
#pragma once

#include <type_traits>

#include "scb_jet.cuh"

namespace scb {

// double overloads so the stage maps below serve both jets (derivatives) and plain values (line search)
// (double overloads of the jet vocabulary: scb_jet.cuh)

template <int MODEL>
struct MpcModel;

// SingleIntegrator2D: f = 0, g = I (robots/single_integrator2D.py:44-62), step :64-66, barrier_dt :148-195.
// Relative degree 1: cbf = d_h + alpha h_k (mpc_cbf.py:312-315).
template <>
struct MpcModel<SCB_SINGLE_INTEGRATOR_2D> {
  static constexpr int NX = 2, NU = 2, NY = 4, REL = 1, NGOAL = 2, AUX = 0, NTRIG = 0;
  static constexpr bool VBOUND = false, LINEAR = false, GENERAL = false;
  static SCB_HD double beta() { return 1.01; }
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double*, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR&) {
    jaxpy(F[0], y[0], p.dt, y[2]);
    jaxpy(F[1], y[1], p.dt, y[3]);
    P1 = F[0]; Q1 = F[1]; P2 = F[0]; Q2 = F[1];
  }
};

// Unicycle2D: f = 0, g robots/unicycle2D.py:43-63, step = Euler + wrap :65-68, barrier_dt :127-145 (plain circle h, no
// sigma term; rel. degree 1).  MPC weights / gain / bounds mpc_cbf.py:22-24, 53-55, 188-192.
template <>
struct MpcModel<SCB_UNICYCLE_2D> {
  static constexpr int NX = 3, NU = 2, NY = 5, REL = 1, NGOAL = 2, AUX = 0, NTRIG = 1;
  static constexpr bool VBOUND = false, LINEAR = false, GENERAL = false;
  static SCB_HD double beta() { return 1.01; }
  // y = (px, py, theta, v, omega)
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double*, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR& trig) {
    T s, c, vc, vs;
    trig(s, c, y[2]);
    jmul(vc, y[3], c); jmul(vs, y[3], s);
    jaxpy(F[0], y[0], p.dt, vc);
    jaxpy(F[1], y[1], p.dt, vs);
    jaxpy(F[2], y[2], p.dt, y[4]);
    P1 = F[0]; Q1 = F[1]; P2 = F[0]; Q2 = F[1];
  }
};

// DynamicUnicycle2D: f, g robots/dynamic_unicycle2D.py:42-73, step :75-78, barrier_dt :188-238
template <>
struct MpcModel<SCB_DYNAMIC_UNICYCLE_2D> {
  static constexpr int NX = 4, NU = 2, NY = 6, REL = 2, NGOAL = 2, AUX = 0, NTRIG = 2;
  static constexpr bool VBOUND = true, LINEAR = false, GENERAL = false;
  static SCB_HD double beta() { return 1.01; }
  // y = (px, py, theta, v, a, omega).  F = Euler map; (P1,Q1), (P2,Q2) = positions after 1 and 2 own steps.
  // trig(s, c, angle) supplies sin/cos (computed, or replayed from the per-stage cache; scb_jet.cuh)
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double*, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR& trig) {
    T s, c, vc, vs;
    trig(s, c, y[2]);
    jmul(vc, y[3], c); jmul(vs, y[3], s);
    jaxpy(F[0], y[0], p.dt, vc);
    jaxpy(F[1], y[1], p.dt, vs);
    jaxpy(F[2], y[2], p.dt, y[5]);
    jaxpy(F[3], y[3], p.dt, y[4]);
    P1 = F[0]; Q1 = F[1];
    T s1, c1, v1c, v1s;
    trig(s1, c1, F[2]);
    jmul(v1c, F[3], c1); jmul(v1s, F[3], s1);
    jaxpy(P2, P1, p.dt, v1c);
    jaxpy(Q2, Q1, p.dt, v1s);
  }
};

// KinematicBicycle2D: f, g robots/kinematic_bicycle2D.py:75-110, step (clips v) :112-123, barrier_dt :175-199
template <>
struct MpcModel<SCB_KINEMATIC_BICYCLE_2D> {
  static constexpr int NX = 4, NU = 2, NY = 6, REL = 2, NGOAL = 2, AUX = 0, NTRIG = 2;
  static constexpr bool VBOUND = true, LINEAR = false, GENERAL = false;
  static SCB_HD double beta() { return 1.1; }
  template <class T, class TR>
  static SCB_HD void euler(const scb_params& p, const T& px, const T& py, const T& th, const T& v, const T& a,
                           const T& b, T* F, TR& trig) {
    T s, c;
    trig(s, c, th);
    euler_sc(p, px, py, th, v, a, b, s, c, F);
  }
  template <class T>
  static SCB_HD void euler_sc(const scb_params& p, const T& px, const T& py, const T& th, const T& v, const T& a,
                              const T& b, const T& s, const T& c, T* F) {
    T vc, vs, vsb, vcb, t0, t1, vb;
    jmul(vc, v, c); jmul(vs, v, s);
    jmul(vsb, vs, b); jmul(vcb, vc, b);
    jaxpy(t0, vc, -1.0, vsb);                 // v c - v s beta
    jaxpy(t1, vs, 1.0, vcb);                  // v s + v c beta
    jaxpy(F[0], px, p.dt, t0);
    jaxpy(F[1], py, p.dt, t1);
    jmul(vb, v, b);
    jaxpy(F[2], th, p.dt / p.rear_ax_dist, vb);
    jaxpy(F[3], v, p.dt, a);
  }
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double*, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR& trig) {
    euler(p, y[0], y[1], y[2], y[3], y[4], y[5], F, trig);
    P1 = F[0]; Q1 = F[1];
    T v1, G[4];
    jclip(v1, F[3], p.v_min, p.v_max);        // the model's own step clips v (:116-121)
    euler(p, P1, Q1, F[2], v1, y[4], y[5], G, trig);
    P2 = G[0]; Q2 = G[1];
  }
};

// ---- models whose barrier is NOT a weighted sum of squared distances ("general" rows) ---------------------------
// KinematicBicycle2D_C3BF / _DPCBF (dynamic_env/kinematic_bicycle2D_{c3bf,dpcbf}.py): dynamics, own step, MPC weights
// and bounds of KinematicBicycle2D (mpc_cbf.py:31-33, 205-211), relative degree 1 (alpha = 0.15, :68-73), barrier
//   cbf_j = h(step(x_k, u_k); o_j) - h(x_k; o_j) + alpha h(x_k; o_j)        (mpc_cbf.py:312-315)
// with h a collision-cone / dynamic-parabolic function of the FULL state and of the obstacle's velocity.  There is no
// 12-sums structure: values, gradients and Hessian entries of every (stage, obstacle) row come from the same jets,
// evaluated per row (MpcSolver's GENERAL branches).  states(): S[0] = x_k, S[1] = own step, with their heading sin/cos.
template <int MODEL>
struct MpcModelKBGeneral {
  using KB = MpcModel<SCB_KINEMATIC_BICYCLE_2D>;
  static constexpr int NX = 4, NU = 2, NY = 6, REL = 1, NGOAL = 2, AUX = 0, NTRIG = 2, NPT = 2;
  static constexpr bool VBOUND = true, LINEAR = false, GENERAL = true;
  static SCB_HD double beta() { return 1.01; }
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double*, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR& trig) {
    KB::euler(p, y[0], y[1], y[2], y[3], y[4], y[5], F, trig);
    P1 = F[0]; Q1 = F[1]; P2 = F[0]; Q2 = F[1];
  }
  template <class T, class TR>
  static SCB_HD void states(const scb_params& p, const T* y, T* F, T (*S)[NX], T* SN, T* CS, TR& trig) {
    trig(SN[0], CS[0], y[2]);
    KB::euler_sc(p, y[0], y[1], y[2], y[3], y[4], y[5], SN[0], CS[0], F);
#pragma unroll
    for (int i = 0; i < NX; ++i) S[0][i] = y[i];
    S[1][0] = F[0]; S[1][1] = F[1]; S[1][2] = F[2];
    jclip(S[1][3], F[3], p.v_min, p.v_max);                  // the model's own step clips v (kinematic_bicycle2D.py:116-121)
    trig(SN[1], CS[1], F[2]);
  }
  // h(x; obs) of the DISCRETE barrier; ob = raw obstacle row [x, y, r, vx, vy, ., .]
  template <class T>
  static SCB_HD void hfun(const scb_params& p, const T* s, const T& sn, const T& cs, const double* ob, T& h) {
    T px, py, vx, vy, t, pm2, vm2, pm, vm, dot;
    jscale(px, s[0], -1.0); jaddc(px, px, ob[0]);            // p_rel = o - p
    jscale(py, s[1], -1.0); jaddc(py, py, ob[1]);
    jmul(t, s[3], cs); jscale(vx, t, -1.0); jaddc(vx, vx, ob[3]);   // v_rel = o_vel - v (cos, sin)
    jmul(t, s[3], sn); jscale(vy, t, -1.0); jaddc(vy, vy, ob[4]);
    jmul(pm2, px, px); jmul(t, py, py); jadd(pm2, pm2, t);
    jmul(vm2, vx, vx); jmul(t, vy, vy); jadd(vm2, vm2, t);
    jsqrt0(pm, pm2); jsqrt0(vm, vm2);
    if (MODEL == SCB_KINEMATIC_BICYCLE_2D_C3BF) {
      // h = <p_rel, v_rel> + |p_rel| |v_rel| sqrt(max(|p_rel|^2 - ego^2, 0)) / |p_rel|,  ego = (r + R) 1.01   (c3bf.py:82-108)
      const double ego = (ob[2] + p.radius) * 1.01;
      T d, sq, ipm;
      jaddc(d, pm2, -ego * ego); jsqrt0(sq, d);
      jmul(dot, px, vx); jmul(t, py, vy); jadd(dot, dot, t);
      jrecip(ipm, pm);
      jmul(t, pm, vm); jmul(t, t, sq); jmul(t, t, ipm);
      jadd(h, dot, t);
    } else {
      // dynamic-parabolic CBF (dpcbf.py:86-136): v_rel rotated into the line-of-sight frame,
      //   h = v_n,x + lambda v_n,y^2 + mu,  lambda = k_l sqrt(d_safe) / |v_rel|,  mu = k_m sqrt(d_safe),
      //   d_safe = max(|p_rel|^2 - ego^2, 1e-6),  ego = (r + R) s,  s = 1.05,  k_l, k_m = (0.1, 0.5) sqrt(s^2 - 1) / ego
      const double sm = 1.05, ego = (ob[2] + p.radius) * sm;
      const double kl = 0.1 * sqrt(sm * sm - 1.0) / ego, km = 0.5 * sqrt(sm * sm - 1.0) / ego;
      T ipm, vnx, vny, d, sd, ivm, lam;
      jrecip(ipm, pm);
      jmul(vnx, px, vx); jmul(t, py, vy); jadd(vnx, vnx, t); jmul(vnx, vnx, ipm);      // cos(rot) vx + sin(rot) vy
      jmul(vny, px, vy); jmul(t, py, vx); jaxpy(vny, vny, -1.0, t); jmul(vny, vny, ipm);   // -sin(rot) vx + cos(rot) vy
      jaddc(d, pm2, -ego * ego);
      if (jval(d) < 1e-6) jconst(d, 1e-6);
      jsqrt0(sd, d);
      jrecip(ivm, vm);
      jmul(lam, sd, ivm); jscale(lam, lam, kl);
      jmul(t, vny, vny); jmul(t, t, lam);
      jadd(h, vnx, t);
      jaxpy(h, h, km, sd);
    }
  }
};
template <> struct MpcModel<SCB_KINEMATIC_BICYCLE_2D_C3BF> : MpcModelKBGeneral<SCB_KINEMATIC_BICYCLE_2D_C3BF> {};
template <> struct MpcModel<SCB_KINEMATIC_BICYCLE_2D_DPCBF> : MpcModelKBGeneral<SCB_KINEMATIC_BICYCLE_2D_DPCBF> {};

// DoubleIntegrator2D: f, g robots/double_integrator2D.py:46-78, step (rescales the velocity to |v| <= v_max) :80-108,
// barrier_dt :223-272 (circle rows; rel. degree 2).  MPC weights / gains / bounds mpc_cbf.py:28-30, 60-63, 200-204.
template <>
struct MpcModel<SCB_DOUBLE_INTEGRATOR_2D> {
  static constexpr int NX = 4, NU = 2, NY = 6, REL = 2, NGOAL = 2, AUX = 0, NTRIG = 0;
  static constexpr bool VBOUND = false, LINEAR = false, GENERAL = false;
  static SCB_HD double beta() { return 1.01; }
  // y = (px, py, vx, vy, ax, ay).  The model's own step scales (vx, vy) by v_max / |v| when |v| > v_max (CasADi
  // if_else, :84-95); the positions after one own step are the Euler ones, after two they use the scaled velocity.
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double*, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR&) {
    jaxpy(F[0], y[0], p.dt, y[2]);
    jaxpy(F[1], y[1], p.dt, y[3]);
    jaxpy(F[2], y[2], p.dt, y[4]);
    jaxpy(F[3], y[3], p.dt, y[5]);
    P1 = F[0]; Q1 = F[1];
    T q, t, sc, v1x, v1y;
    jmul(q, F[2], F[2]); jmul(t, F[3], F[3]); jaxpy(q, q, 1.0, t);       // |v_1|^2
    const double qv = jval(q), vmax = p.v_max;
    if (qv > vmax * vmax) {
      const double r = 1.0 / sqrt(qv), r2 = r * r;                       // scale = v_max q^(-1/2)
      jchain(sc, q, vmax * r, -0.5 * vmax * r * r2, 0.75 * vmax * r * r2 * r2);
      jmul(v1x, F[2], sc); jmul(v1y, F[3], sc);
    } else {
      v1x = F[2]; v1y = F[3];
    }
    jaxpy(P2, P1, p.dt, v1x);
    jaxpy(Q2, Q1, p.dt, v1y);
  }
};

// Quad2D: f, g robots/quad2D.py:46-85 (x, z, theta, xdot, zdot, thetadot; two rotor forces), step = Euler + wrap :87-90,
// barrier_dt :178-203 (circle rows on (x, z); rel. degree 2).  MPC weights / gains / bounds mpc_cbf.py:34-36, 74-77, 212-216.
template <>
struct MpcModel<SCB_QUAD_2D> {
  static constexpr int NX = 6, NU = 2, NY = 8, REL = 2, NGOAL = 2, AUX = 0, NTRIG = 1;
  static constexpr bool VBOUND = false, LINEAR = false, GENERAL = false;
  static SCB_HD double beta() { return 1.01; }
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double*, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR& trig) {
    T s, c, usum, udif, sx, cz;
    trig(s, c, y[2]);
    jaxpy(usum, y[6], 1.0, y[7]);            // u_right + u_left
    jaxpy(udif, y[6], -1.0, y[7]);           // u_right - u_left
    jmul(sx, s, usum); jmul(cz, c, usum);
    jaxpy(F[0], y[0], p.dt, y[3]);
    jaxpy(F[1], y[1], p.dt, y[4]);
    jaxpy(F[2], y[2], p.dt, y[5]);
    jaxpy(F[3], y[3], -p.dt / p.mass, sx);
    jaxpy(F[4], y[4], p.dt / p.mass, cz);
    T gconst; jconst(gconst, -p.gravity * p.dt);
    jaxpy(F[4], F[4], 1.0, gconst);
    jaxpy(F[5], y[5], p.dt * p.radius / p.Iy, udif);
    P1 = F[0]; Q1 = F[1];
    jaxpy(P2, P1, p.dt, F[3]);
    jaxpy(Q2, Q1, p.dt, F[4]);
  }
};

// VTOL2D (quadplane in the x-z plane, robots/vtol2D.py): X = [x, z, theta, x_dot, z_dot, theta_dot], U = [front, rear,
// pusher rotor throttle, elevator]; control affine:  f = baseline aerodynamics (elevator 0) + gravity (:115-193),
// g = rotor thrusts rotated by theta and the aerodynamic forces / moment AT elevator deflection 1 (:198-300 -- the full
// value, transcribed literally, not the increment).  step = Euler + wrap (:302-310); barrier_dt (:462-488): circle h on
// (x, z), relative degree 2, positions after two own steps = p_1 + v_1 dt, so the aerodynamics are evaluated once per
// stage.  MPC weights / gains / bounds / horizon 30: mpc_cbf.py:40-43, 83-87, 222-232.
template <>
struct MpcModel<SCB_VTOL_2D> {
  static constexpr int NX = 6, NU = 4, NY = 10, REL = 2, NGOAL = 2, AUX = 0, NTRIG = 6;   // 3 sin/cos pairs + alpha, 2 exp
  static constexpr bool VBOUND = false, LINEAR = false, GENERAL = false;
  static SCB_HD double beta() { return 1.01; }
  // lift, drag, pitching moment at elevator deflection de  (_lift_blending :351-377, _lift_drag_moment :379-412)
  template <class T>
  static SCB_HD void lift_drag_moment(const scb_params& p, const T& V2, const T& al, const T& CLa, double de, T& L, T& D, T& Mo) {
    T CL, CD, CM, t, q;
    jaddc(CL, CLa, p.C_Ldelta_e * de);
    jmul(t, al, al); jscale(CD, t, p.C_Dalpha); jaddc(CD, CD, p.C_D0 + p.C_Ddelta_e * de);
    jscale(CM, al, p.C_malpha); jaddc(CM, CM, p.C_m0 + p.C_mdelta_e * de);
    jscale(q, V2, 0.5 * p.rho);                                  // qbar = 0.5 rho V^2
    jmul(L, q, CL); jscale(L, L, p.S_wing);
    jmul(D, q, CD); jscale(D, D, p.S_wing);
    jmul(Mo, q, CM); jscale(Mo, Mo, p.S_wing * p.chord);
  }
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double*, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR& trig) {
    const T &th = y[2], &xd = y[3], &zd = y[4];
    T s, c, ub, wb, nwb, t, V2, al;
    trig(s, c, th);
    jmul(ub, c, xd); jmul(t, s, zd); jadd(ub, ub, t);            // u_b =  c x_dot + s z_dot   (:333-343)
    jmul(wb, c, zd); jmul(t, s, xd); jaxpy(wb, wb, -1.0, t);     // w_b = -s x_dot + c z_dot
    jmul(V2, ub, ub); jmul(t, wb, wb); jadd(V2, V2, t);          // V^2
    jscale(nwb, wb, -1.0);
    const double alv = trig.memo([&] { return atan2(-jval(wb), jval(ub)); });
    jatan2(al, nwb, ub, alv);                                    // alpha = atan2(-w_b, u_b)
    // blended lift coefficient
    T CLlin, sa, ca, CLnl, e1, e2, num, den, sig, CLa, a1, a2;
    jscale(CLlin, al, p.C_Lalpha); jaddc(CLlin, CLlin, p.C_L0);
    trig(sa, ca, al);
    jmul(CLnl, sa, ca); jscale(CLnl, CLnl, 2.0);
    jaddc(a1, al, -p.alpha_0); jscale(a1, a1, -p.blend_M);       // -M (alpha - alpha_0)
    jaddc(a2, al, p.alpha_0); jscale(a2, a2, p.blend_M);         //  M (alpha + alpha_0)
    const double e1v = trig.memo([&] { return exp(jval(a1)); }), e2v = trig.memo([&] { return exp(jval(a2)); });
    jchain(e1, a1, e1v, e1v, e1v); jchain(e2, a2, e2v, e2v, e2v);
    jadd(num, e1, e2); jaddc(num, num, 1.0);                     // 1 + tmp1 + tmp2
    T d1, d2, iden;
    jaddc(d1, e1, 1.0); jaddc(d2, e2, 1.0); jmul(den, d1, d2);
    jrecip(iden, den); jmul(sig, num, iden);                     // sigma
    jaxpy(t, CLnl, -1.0, CLlin); jmul(t, sig, t); jadd(CLa, CLlin, t);     // (1 - sigma) CL_lin + sigma CL_nl
    // forces: baseline (delta_e = 0) and the elevator column (delta_e = 1), wind -> inertial by theta + alpha
    T L0, D0, M0, L1, D1, M1, hd, sh, ch;
    lift_drag_moment(p, V2, al, CLa, 0.0, L0, D0, M0);
    lift_drag_moment(p, V2, al, CLa, 1.0, L1, D1, M1);
    jadd(hd, th, al);
    trig(sh, ch, hd);
    T fx0, fz0, fx1, fz1;
    jmul(fx0, ch, D0); jscale(fx0, fx0, -1.0); jmul(t, sh, L0); jaxpy(fx0, fx0, -1.0, t);    // c (-D) - s L
    jmul(fz0, sh, D0); jscale(fz0, fz0, -1.0); jmul(t, ch, L0); jadd(fz0, fz0, t);           // s (-D) + c L
    jmul(fx1, ch, D1); jscale(fx1, fx1, -1.0); jmul(t, sh, L1); jaxpy(fx1, fx1, -1.0, t);
    jmul(fz1, sh, D1); jscale(fz1, fz1, -1.0); jmul(t, ch, L1); jadd(fz1, fz1, t);
    const double im = 1.0 / p.mass, iI = 1.0 / p.Iy;
    // accelerations: f + g u
    T xdd, zdd, tdd, rot, g3;
    jscale(xdd, fx0, im);
    jscale(zdd, fz0, im); jaddc(zdd, zdd, -p.gravity);           // (fz - m g) / m
    jscale(tdd, M0, iI);
    jaxpy(rot, y[6], p.k_rear / p.k_front, y[7]); jscale(rot, rot, p.k_front * im);   // (k_f u0 + k_r u1) / m
    jmul(t, s, rot); jaxpy(xdd, xdd, -1.0, t);                   // front / rear rotors: (-s, c) k / m
    jmul(t, c, rot); jadd(zdd, zdd, t);
    jscale(g3, y[8], p.k_pusher * im);                           // pusher: (c, s) k_p / m
    jmul(t, c, g3); jadd(xdd, xdd, t);
    jmul(t, s, g3); jadd(zdd, zdd, t);
    jmul(t, fx1, y[9]); jaxpy(xdd, xdd, im, t);                  // elevator column
    jmul(t, fz1, y[9]); jaxpy(zdd, zdd, im, t);
    jaxpy(tdd, tdd, p.ell_f * p.k_front * iI, y[6]);
    jaxpy(tdd, tdd, -p.ell_r * p.k_rear * iI, y[7]);
    jmul(t, M1, y[9]); jaxpy(tdd, tdd, iI, t);
    jaxpy(F[0], y[0], p.dt, xd);
    jaxpy(F[1], y[1], p.dt, zd);
    jaxpy(F[2], th, p.dt, y[5]);
    jaxpy(F[3], xd, p.dt, xdd);
    jaxpy(F[4], zd, p.dt, zdd);
    jaxpy(F[5], y[5], p.dt, tdd);
    P1 = F[0]; Q1 = F[1];
    jaxpy(P2, P1, p.dt, F[3]);                                   // second own step moves the position by v_1 dt
    jaxpy(Q2, Q1, p.dt, F[4]);
  }
};

// ---- optimal-decay MPC-CBF (position_control/optimal_decay_mpc_cbf.py) ---------------------------------------------
// Same dynamics / state cost / bounds as MPC-CBF, plus two extra STAGE INPUTS omega1, omega2 (:122-124) that scale the
// class-K gains of the relative-degree-2 discrete CBF row (:296-300):
//     dd_h + (alpha1 omega1 + alpha2 omega2) d_h + alpha1 alpha2 omega1 omega2 h_k >= 0,
// and an input cost that is NOT a rate penalty: the reference passes do-mpc two expression rterms (:178-185),
// sum_i R_i u_i^2 and p_sb1 (omega1 - 1)^2 + p_sb2 (omega2 - 1)^2, evaluated at u_k (see MpcSolver::Ra / ut).
// The row weights depend on the stage inputs, so every (stage, obstacle) row goes through the general-row jets with the
// omegas riding along in the (otherwise unused) heading slots of the barrier states.  Inputs of the NLP per stage:
// [u (NUB, bounded), omega1, omega2 (unbounded)].
constexpr int kMpcOdBase = 200;
template <int BASE>
struct MpcModelOD : MpcModel<BASE> {
  using B = MpcModel<BASE>;
  static_assert(B::REL == 2 && !B::LINEAR, "optimal-decay rows exist for the relative-degree-2 models");
  static constexpr int NX = B::NX, NUB = B::NU, NU = B::NU + 2, NY = NX + NU, NPT = 3;
  static constexpr bool GENERAL = true, OD = true;
  template <class T, class TR>
  static SCB_HD void stage(const scb_params& p, const double* aux, const T* y, T* F, T& P1, T& Q1, T& P2, T& Q2, TR& trig) {
    B::stage(p, aux, y, F, P1, Q1, P2, Q2, trig);          // (the base map reads y[0 .. NX + NUB) only)
  }
  template <class T, class TR>
  static SCB_HD void states(const scb_params& p, const T* y, T* F, T (*S)[NX], T* SN, T* CS, TR& trig) {
    T P1, Q1, P2, Q2;
    B::stage(p, nullptr, y, F, P1, Q1, P2, Q2, trig);
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
#pragma unroll
      for (int c = 0; c < NX; ++c) jconst(S[i][c], 0.0);
      jconst(SN[i], 0.0); jconst(CS[i], 0.0);
    }
    S[0][0] = y[0]; S[0][1] = y[1];
    S[1][0] = P1; S[1][1] = Q1;
    S[2][0] = P2; S[2][1] = Q2;
    SN[0] = y[NX + NUB]; CS[0] = y[NX + NUB + 1];          // omega1, omega2
  }
  template <class T>
  static SCB_HD void hfun(const scb_params& p, const T* s, const T&, const T&, const double* ob, T& h) {
    T dx, dy, t;
    jaddc(dx, s[0], -ob[0]); jaddc(dy, s[1], -ob[1]);
    const double d = ob[2] + p.radius;
    jmul(h, dx, dx); jmul(t, dy, dy); jadd(h, h, t); jaddc(h, h, -B::beta() * d * d);
  }
};
template <> struct MpcModel<kMpcOdBase + SCB_DYNAMIC_UNICYCLE_2D> : MpcModelOD<SCB_DYNAMIC_UNICYCLE_2D> {};
template <> struct MpcModel<kMpcOdBase + SCB_KINEMATIC_BICYCLE_2D> : MpcModelOD<SCB_KINEMATIC_BICYCLE_2D> {};
template <> struct MpcModel<kMpcOdBase + SCB_QUAD_2D> : MpcModelOD<SCB_QUAD_2D> {};
template <> struct MpcModel<kMpcOdBase + SCB_VTOL_2D> : MpcModelOD<SCB_VTOL_2D> {};

// number of BOUNDED inputs (the first NUB of the NU stage inputs) and the optimal-decay flag, with defaults
template <class Mod, class = void> struct MpcNub { static constexpr int value = Mod::NU; };
template <class Mod> struct MpcNub<Mod, std::void_t<decltype(Mod::NUB)>> { static constexpr int value = Mod::NUB; };
template <class Mod, class = void> struct MpcOd { static constexpr bool value = false; };
template <class Mod> struct MpcOd<Mod, std::void_t<decltype(Mod::OD)>> { static constexpr bool value = Mod::OD; };

// state-bound rows of the MPC per node k = 1..H (mpc_cbf.py:193-199, 205-211: |x[3]| <= v_max for the unicycle / bicycle
// models; :227-232 for VTOL2D): row r of a node is  sgn * x[var] + off >= 0
template <class Mod>
struct MpcStateBounds {
  static constexpr int NSB = Mod::VBOUND ? 2 : 0;
  static SCB_HD void get(const scb_params& p, int r, int& var, double& sgn, double& off) {
    var = 3; sgn = (r & 1) ? 1.0 : -1.0; off = p.v_max;
  }
};
template <>
struct MpcStateBounds<MpcModel<SCB_VTOL_2D>> {
  static constexpr int NSB = 5;
  static SCB_HD void get(const scb_params& p, int r, int& var, double& sgn, double& off) {
    const double pitch = p.pitch_max * 3.14159 / 180.0;          // (the reference's own constant, mpc_cbf.py:231-232)
    switch (r) {
      case 0: var = 3; sgn = -1.0; off = p.v_max; break;
      case 1: var = 3; sgn = 1.0; off = p.v_max; break;
      case 2: var = 4; sgn = 1.0; off = p.descent_speed_max; break;
      case 3: var = 2; sgn = -1.0; off = pitch; break;
      default: var = 2; sgn = 1.0; off = pitch; break;
    }
  }
};

// Superellipsoid obstacles in MPC (SingleIntegrator2D / DynamicUnicycle2D / DoubleIntegrator2D): their agent_barrier_dt
// picks h per obstacle with if_else(obs[6] < 0.5, circle, superellipsoid) (e.g. dynamic_unicycle2D.py:204-228), and the
// superellipsoid h = (|x'| / (a + R))^e + (|y'| / (b + R))^e - 1 (guards: a, b >= 1e-3, e >= 2) has no 12-sums
// structure.  Agents with at least one flagged row run through these "general row" variants of the same models
// (internal ids kMpcSeBase + base id; selected per agent, see mpc_kernel), everyone else through the fast path.
constexpr int kMpcSeBase = 100;
template <int BASE>
struct MpcModelSE : MpcModel<BASE> {
  using B = MpcModel<BASE>;
  static constexpr int NX = B::NX, NPT = B::REL + 1;
  static constexpr bool GENERAL = true;
  template <class T, class TR>
  static SCB_HD void states(const scb_params& p, const T* y, T* F, T (*S)[NX], T* SN, T* CS, TR& trig) {
    T P1, Q1, P2, Q2;
    B::stage(p, nullptr, y, F, P1, Q1, P2, Q2, trig);
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
#pragma unroll
      for (int c = 0; c < NX; ++c) jconst(S[i][c], 0.0);
      jconst(SN[i], 0.0); jconst(CS[i], 0.0);
    }
    S[0][0] = y[0]; S[0][1] = y[1];
    S[1][0] = P1; S[1][1] = Q1;
    if constexpr (B::REL == 2) { S[2][0] = P2; S[2][1] = Q2; }
  }
  template <class T>
  static SCB_HD void hfun(const scb_params& p, const T* s, const T&, const T&, const double* ob, T& h) {
    T dx, dy, t;
    jaddc(dx, s[0], -ob[0]); jaddc(dy, s[1], -ob[1]);
    if (ob[6] < 0.5) {                                        // circle
      const double d = ob[2] + p.radius;
      jmul(h, dx, dx); jmul(t, dy, dy); jadd(h, h, t); jaddc(h, h, -B::beta() * d * d);
    } else {                                                  // superellipsoid [x, y, a, b, e, theta, 1]
      const double a = fmax(fabs(ob[2]), 1e-3), b = fmax(fabs(ob[3]), 1e-3), e = fmax(fabs(ob[4]), 2.0);
      double st, ct; sincos_pair(ob[5], st, ct);
      T xp, yp;
      jscale(xp, dx, ct); jaxpy(xp, xp, st, dy);
      jscale(yp, dy, ct); jaxpy(yp, yp, -st, dx);
      T px, py;
      jpowabs(px, xp, e, 1.0 / (a + p.radius));
      jpowabs(py, yp, e, 1.0 / (b + p.radius));
      jadd(h, px, py); jaddc(h, h, -1.0);
    }
  }
};
template <> struct MpcModel<kMpcSeBase + SCB_SINGLE_INTEGRATOR_2D> : MpcModelSE<SCB_SINGLE_INTEGRATOR_2D> {};
template <> struct MpcModel<kMpcSeBase + SCB_DYNAMIC_UNICYCLE_2D> : MpcModelSE<SCB_DYNAMIC_UNICYCLE_2D> {};
template <> struct MpcModel<kMpcSeBase + SCB_DOUBLE_INTEGRATOR_2D> : MpcModelSE<SCB_DOUBLE_INTEGRATOR_2D> {};

// Quad3D: linear 12-state model xdot = A x + B u (robots/quad3D.py:81-97); the MPC uses plain Euler
// (mpc_cbf.py:135-141) while the barrier uses the model's own RK4 step (quad3D.py:121-158, 275-297), which for a
// linear system is the constant affine map x1 = Ad x + Bd u.  Relative degree 1: cbf = d_h + alpha h_k.  Because
// everything is linear the stage derivatives are written down directly (no jets); per-agent constants live in
// the AUX block: Ae = I + dt A (12x12), Be = dt B (12x4), rows 0 and 1 of [Ad | Bd] (2 x 16).
template <>
struct MpcModel<SCB_QUAD_3D> {
  static constexpr int NX = 12, NU = 4, NY = 16, REL = 1, NGOAL = 3, AUX = 144 + 48 + 32, NTRIG = 0;
  static constexpr bool VBOUND = false, LINEAR = true, GENERAL = false;
  static SCB_HD double beta() { return 1.01; }

  static SCB_HD void setup_aux(const scb_params& p, double* aux) {
    double A[144], B[48];
    for (int i = 0; i < 144; ++i) A[i] = 0.0;
    for (int i = 0; i < 48; ++i) B[i] = 0.0;
    for (int i = 0; i < 6; ++i) A[i * 12 + 6 + i] = 1.0;
    A[6 * 12 + 3] = p.gravity; A[7 * 12 + 4] = -p.gravity;
    // B = B1 B2 (quad3D.py:70-97)
    const double L = p.arm_L, nu = p.nu_coef;
    const double B2[4][4] = {{1, 1, 1, 1}, {0, L, 0, -L}, {L, 0, -L, 0}, {nu, -nu, nu, -nu}};
    const double b1[4] = {1.0 / p.mass, 1.0 / p.Iy, 1.0 / p.Ix, 1.0 / p.Iz};
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) B[(8 + r) * 4 + c] = b1[r] * B2[r][c];
    double* Ae = aux; double* Be = aux + 144; double* R = aux + 192;
    for (int i = 0; i < 12; ++i)
      for (int j = 0; j < 12; ++j) Ae[i * 12 + j] = (i == j ? 1.0 : 0.0) + p.dt * A[i * 12 + j];
    for (int i = 0; i < 48; ++i) Be[i] = p.dt * B[i];
    // rows 0, 1 of Ad = sum_{m<=4} dt^m/m! A^m and Bd = (sum_{m<=3} dt^{m+1}/(m+1)! A^m) B
    for (int row = 0; row < 2; ++row) {
      double q[12], acc_a[12], acc_b[12];
      for (int j = 0; j < 12; ++j) { q[j] = (j == row) ? 1.0 : 0.0; acc_a[j] = q[j]; acc_b[j] = p.dt * q[j]; }
      double ca = 1.0, cb = p.dt;
      for (int m = 1; m <= 4; ++m) {
        double qn[12];
        for (int j = 0; j < 12; ++j) {
          double v = 0.0;
          for (int t = 0; t < 12; ++t) v = fma(q[t], A[t * 12 + j], v);
          qn[j] = v;
        }
        for (int j = 0; j < 12; ++j) q[j] = qn[j];
        ca *= p.dt / m;
        for (int j = 0; j < 12; ++j) acc_a[j] = fma(ca, q[j], acc_a[j]);
        if (m <= 3) {
          cb *= p.dt / (m + 1);
          for (int j = 0; j < 12; ++j) acc_b[j] = fma(cb, q[j], acc_b[j]);
        }
      }
      for (int j = 0; j < 12; ++j) R[row * 16 + j] = acc_a[j];
      for (int c = 0; c < 4; ++c) {
        double v = 0.0;
        for (int t = 0; t < 12; ++t) v = fma(acc_b[t], B[t * 4 + c], v);
        R[row * 16 + 12 + c] = v;
      }
    }
  }

  // plain-value stage map (rollout / line search)
  template <class TR>
  static SCB_HD void stage(const scb_params&, const double* aux, const double* y, double* F, double& P1, double& Q1,
                           double& P2, double& Q2, TR&) {
    const double* Ae = aux; const double* Be = aux + 144; const double* R = aux + 192;
    for (int i = 0; i < 12; ++i) {
      double v = 0.0;
      for (int j = 0; j < 12; ++j) v = fma(Ae[i * 12 + j], y[j], v);
      for (int j = 0; j < 4; ++j) v = fma(Be[i * 4 + j], y[12 + j], v);
      F[i] = v;
    }
    double a = 0.0, b = 0.0;
    for (int j = 0; j < 16; ++j) { a = fma(R[j], y[j], a); b = fma(R[16 + j], y[j], b); }
    P1 = a; Q1 = b; P2 = a; Q2 = b;
  }
};

template <>
struct MpcStateBounds<MpcModel<kMpcOdBase + SCB_VTOL_2D>> : MpcStateBounds<MpcModel<SCB_VTOL_2D>> {};

// ---------------------------------------------------------------------------------------------
// workspace layout (in doubles), computed identically on host and device
struct MpcLayout {
  int H, M, n, NS;
  int X, Z, A, B, JE, JX, JY, PT, OB, C, L, DS, DL, CT, SS, SL, SDS, SDL, SUM, G, GAM, MU, RD, DZ, PM, PV, KG, KF,
      MM, PT2, DY, ZT, XT, RG, AUX, AS, BS, TR, TRS, GR, SP, OBS7;
  int total;
};

// SEQ = true: single-lane (host-sim) build, which needs one extra scratch block (PT2) because one lane plays all
// matrix columns of the Riccati stage in turn.  Linear models keep ONE copy of (A, B) (their AUX block) instead of H
// identical ones (AS = BS = 0), and the Riccati value function is double-buffered (the forward sweep only needs the
// gains), so the workspace is O(H) only in what really differs per stage.
template <class Mod, bool SEQ = false>
SCB_HD MpcLayout mpc_layout(int H, int M) {
  constexpr int NX = Mod::NX, NU = Mod::NU, AUXN = Mod::AUX, NTRIG = Mod::NTRIG;
  constexpr bool VBOUND = Mod::VBOUND, LINEAR = Mod::LINEAR, GENERAL = Mod::GENERAL;
  constexpr int NY = NX + NU, NH = NY * (NY + 1) / 2;
  MpcLayout L;
  L.H = H; L.M = M; L.n = H * NU; L.NS = 2 * H * MpcNub<Mod>::value + MpcStateBounds<Mod>::NSB * H;
  int o = 0;
  auto take = [&](int cnt) { int r = o; o += cnt; return r; };
  L.X = take((H + 1) * NX);  L.Z = take(H * NU);
  L.AUX = take(AUXN);
  if (LINEAR) { L.A = L.AUX; L.B = L.AUX + NX * NX; L.AS = 0; L.BS = 0; }
  else { L.A = take(H * NX * NX); L.B = take(H * NX * NU); L.AS = NX * NX; L.BS = NX * NU; }
  L.JE = take(GENERAL ? 0 : H * NY); L.JX = take(GENERAL ? 0 : H * NY); L.JY = take(GENERAL ? 0 : H * NY);   // gE, gX, gY
  L.PT = take(GENERAL ? 0 : H * 6);        L.OB = take(GENERAL ? 0 : M * 3);
  // general rows: gradient of every (stage, obstacle) row, the stage's barrier states (+ sin, cos), raw obstacle rows
  L.GR = take(GENERAL ? H * M * NY : 0); L.SP = take(GENERAL ? H * 2 * (NX + 2) : 0); L.OBS7 = take(GENERAL ? M * 7 : 0);
  L.C = take(H * M); L.L = take(H * M); L.DS = take(H * M); L.DL = take(H * M); L.CT = take(H * M);
  L.SS = take(L.NS); L.SL = take(L.NS); L.SDS = take(L.NS); L.SDL = take(L.NS);
  L.SUM = take(GENERAL ? 0 : H * 12);
  L.G = take((H + 1) * NH);  L.GAM = take((H + 1) * NY);
  L.MU = take((H + 1) * NX);
  L.RD = take(L.n); L.DZ = take(L.n);
  {
    const int NXT = NX + NU, NV = NXT + NU;
    L.PM = take(2 * NXT * NXT); L.PV = take(2 * NXT);     // P_{k+1} / P_k ping-pong (slot k & 1)
    L.KG = take(H * NU * NXT); L.KF = take(H * NU);
    L.MM = take(NU * NV);                               // published input columns of a Riccati stage
    L.PT2 = take(SEQ ? (NV + 1) * NV : 0);
  }
  L.DY = take((H + 1) * NY);
  L.TRS = 2 * NTRIG; L.TR = take(H * L.TRS);              // sin/cos of every stage at the current iterate
  L.ZT = take(L.n); L.XT = take((H + 1) * NX); L.RG = take(L.n);
  L.total = o;
  return L;
}

// words of the MPC active mask (include/scb.h scb_mpc_active_words): H*M CBF rows, then the NS simple bounds
template <class Mod>
SCB_HD int mpc_active_words(int H, int M) {
  return (H * M + 2 * H * MpcNub<Mod>::value + MpcStateBounds<Mod>::NSB * H + 63) / 64;
}

// simple (bound) constraint q:  g_q = sgn * y_k[var] + off >= 0
struct SimpleCon { int k, var; double sgn, off; };
template <class Mod>
SCB_HD SimpleCon decode_simple(const scb_params& p, int H, int q) {
  constexpr int NX = Mod::NX, NUB = MpcNub<Mod>::value, NSB = MpcStateBounds<Mod>::NSB;
  SimpleCon c;
  if (q < 2 * H * NUB) {
    c.k = q / (2 * NUB);
    const int r = q - c.k * 2 * NUB, i = r >> 1;
    c.var = NX + i;
    if (r & 1) { c.sgn = 1.0; c.off = -p.u_lb[i]; } else { c.sgn = -1.0; c.off = p.u_ub[i]; }
  } else {
    const int t = q - 2 * H * NUB, nsb = NSB > 0 ? NSB : 1;
    c.k = 1 + t / nsb;
    MpcStateBounds<Mod>::get(p, t - (c.k - 1) * nsb, c.var, c.sgn, c.off);
  }
  return c;
}

// Optional per-phase cycle counters (device debug build, -DSCB_MPC_PROFILE): warp 0 of block 0 accumulates
// clock64() deltas at phase boundaries into g_mpc_prof; read back with scb_debug_mpc_profile().
#if defined(SCB_MPC_PROFILE) && defined(__CUDACC__)
__device__ long long g_mpc_prof[24];
SCB_HD long long prof_clock() {
#if defined(__CUDA_ARCH__)
  return clock64();
#else
  return 0;
#endif
}
SCB_HD void prof_add(int i, long long& tlast) {
#if defined(__CUDA_ARCH__)
  if (blockIdx.x == 0 && threadIdx.x == 0) { const long long t = clock64(); g_mpc_prof[i] += t - tlast; tlast = t; }
#endif
}
#define SCB_PH(i) prof_add(i, tlast)
#define SCB_PH_INIT long long tlast = prof_clock()
#else
#define SCB_PH(i) do { } while (0)
#define SCB_PH_INIT do { } while (0)
#endif

// Loops with run-time trip counts (H, M) are NOT unrolled: measured on cfg3, unrolling the lane-strided loops x2 / x4
// (17.9 k / 27 k SASS instructions instead of 12.9 k) made the kernel 1.8x / 2.8x slower -- instruction fetch, not
// latency, is what this kernel waits for.
#ifndef SCB_MPC_UNROLL
#define SCB_MPC_UNROLL 1
#endif
#define SCB_PRAGMA_(x) _Pragma(#x)
#define SCB_PRAGMA(x) SCB_PRAGMA_(x)
#if defined(__CUDACC__)
#define SCB_LANE_UNROLL SCB_PRAGMA(unroll SCB_MPC_UNROLL)
#define SCB_LOOP SCB_PRAGMA(unroll 1)
#else
#define SCB_LANE_UNROLL
#define SCB_LOOP
#endif

// The MPC kernel is instruction-cache bound when everything is inlined (27 k SASS instructions = 430 KB with the
// loops unrolled x4 ran 2.8x slower than 12.9 k): the phases that are called from several places are real
// functions on the device, and the transcendental / reduction helpers exist once.
// (SCB_PHASE, sincos_call, log_call: scb_core.cuh)

// SCB_MPC_PHASE: the solver's phases; inline by default (out-of-line phases measured slower: cfg3 6.4 vs 5.2 ms,
// `this` and the layout then live in local memory), -DSCB_MPC_PHASES_OUTLINE to compare.
#if defined(SCB_MPC_PHASES_OUTLINE)
#define SCB_MPC_PHASE SCB_PHASE
#else
#define SCB_MPC_PHASE SCB_HD
#endif

#ifndef SCB_MPC_OUTLINE_RED
#define SCB_MPC_OUTLINE_RED 0
#endif

// barrier schedule (IPOPT's defaults: mu_init 0.1, kappa_mu 0.2, theta_mu 1.5, kappa_epsilon 10)
#ifndef SCB_MPC_MU0
#define SCB_MPC_MU0 0.1
#endif
#ifndef SCB_MPC_KAPPA_MU
#define SCB_MPC_KAPPA_MU 0.2
#endif
#ifndef SCB_MPC_KAPPA_EPS
#define SCB_MPC_KAPPA_EPS 10.0
#endif
#ifndef SCB_MPC_CRAWL_ITERS
#define SCB_MPC_CRAWL_ITERS 15
#endif
#ifndef SCB_MPC_CRAWL_ALPHA
#define SCB_MPC_CRAWL_ALPHA 1e-4
#endif
#ifndef SCB_MPC_NOPROGRESS
#define SCB_MPC_NOPROGRESS 50
#endif
#ifndef SCB_MPC_STALL_BT
#define SCB_MPC_STALL_BT 15
#endif
#ifndef SCB_MPC_STALL_ITERS
#define SCB_MPC_STALL_ITERS 6
#endif

template <int MODEL, int LANES>
struct MpcSolver {
  using Mod = MpcModel<MODEL>;
  using G = Grp<LANES>;
  static constexpr int NX = Mod::NX, NU = Mod::NU, NY = Mod::NY, NH = NY * (NY + 1) / 2;
  static constexpr int NPT = Mod::REL + 1;     // barrier states per stage (general rows)
  static constexpr int NUB = MpcNub<Mod>::value;          // bounded inputs (optimal decay: omega1, omega2 follow, unbounded)
  static constexpr bool OD = MpcOd<Mod>::value;

  const scb_params& p;
  const MpcLayout& L;
  // The workspace.  On the device it is the group's slice of dynamic shared memory, addressed as (shared array + offset)
  // at every use so that the compiler emits LDS / STS with the offset folded in; a stored generic pointer made it emit
  // generic LD / ST (897 of them in the DynamicUnicycle2D kernel), which resolve the address space at run time.
#if defined(__CUDA_ARCH__) && !defined(SCB_MPC_GENERIC_WS)
  unsigned wofs;
  SCB_HD double* wbase() const { extern __shared__ double scb_mpc_smem_[]; return scb_mpc_smem_ + wofs; }
#else
  double* wptr;
  SCB_HD double* wbase() const { return wptr; }
#endif
#define w (wbase())
  int H, M, n, lane;
  double w0, w1, w2, Wsum;      // c_j = w0 h(p0) + w1 h(p1) + w2 h(p2),  h = |p - o|^2 - beta d^2
  double goal[NX];
  double uprev[NU];
  double Qs[NX], Rs[NU];        // cost weights times the objective scaling factor (IPOPT-style gradient-based scaling)
  // optimal decay only: the input cost is sum_i Ra_i (u_i - ut_i)^2 at every stage instead of the input-RATE penalty --
  // the two expression rterms of optimal_decay_mpc_cbf.py:178-185.  do-mpc's MPC.set_rterm ASSIGNS the expression, so the
  // second call (the omega penalty) replaces the first (sum R_i u_i^2): Ra = [0 ..., p_sb1, p_sb2], ut = [0 ..., omega1_0,
  // omega2_0].  scb_params.od_sum_rterms = 1 keeps both terms (the evident intent of the reference).  UNPINNED: do-mpc is
  // not installable here.
  double Ra[NU], ut[NU];
  bool gauss_newton;            // assemble stage Hessians without the (possibly indefinite) curvature terms
  double floor_cur;             // slack floor mu / nu of the current iteration (general rows re-derive their weights)

  SCB_HD MpcSolver(const scb_params& p_, const MpcLayout& L_, double* w_) : p(p_), L(L_) {
#if defined(__CUDA_ARCH__) && !defined(SCB_MPC_GENERIC_WS)
    extern __shared__ double scb_mpc_smem_[];
    wofs = (unsigned)(w_ - scb_mpc_smem_);
#else
    wptr = w_;
#endif
    H = L.H; M = L.M; n = L.n; lane = G::lane();
    if (Mod::REL == 2) {
      const double g1 = p.alpha1 + p.alpha2, g2 = p.alpha1 * p.alpha2;
      w2 = 1.0; w1 = g1 - 2.0; w0 = 1.0 - g1 + g2; Wsum = g2;  // dd_h + (a1+a2) d_h + a1 a2 h_k  (mpc_cbf.py:320-321)
    } else {
      w2 = 0.0; w1 = 1.0; w0 = p.alpha - 1.0; Wsum = p.alpha;  // d_h + alpha h_k                 (mpc_cbf.py:314-315)
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) Qs[i] = p.Q[i];
#pragma unroll
    for (int i = 0; i < NU; ++i) { Rs[i] = (!OD && i < 4) ? p.R[i] : 0.0; Ra[i] = 0.0; ut[i] = 0.0; }
    if constexpr (OD) {
#pragma unroll
      for (int i = 0; i < NUB; ++i) Ra[i] = (p.od_sum_rterms && i < 4) ? p.R[i] : 0.0;
      Ra[NUB] = p.p_sb1; ut[NUB] = p.omega1_0;
      Ra[NUB + 1] = p.p_sb2; ut[NUB + 1] = p.omega2_0;
    }
    gauss_newton = false; floor_cur = 0.0;
  }

  // group reductions as real functions (20 call sites x a 5-step double-precision butterfly is ~1.7 k instructions inline)
#if SCB_MPC_OUTLINE_RED
#define SCB_MPC_RED SCB_PHASE
#else
#define SCB_MPC_RED SCB_HD
#endif
  static SCB_MPC_RED double gsum(double v) { return G::sum(v); }
  static SCB_MPC_RED double gmin(double v) { return G::vmin(v); }
  static SCB_MPC_RED double gmax(double v) { return -G::vmin(-v); }

  static SCB_HD void sync() {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) __syncwarp(G::gmask());
#endif
  }

  // ---- rollout + cost at the control sequence `z` (lane 0), states -> xs ----
  SCB_MPC_PHASE double rollout(const double* z, double* xs) const {
    double Jc = 0.0;
    double x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xs[i];
    SCB_LOOP
    for (int k = 0; k < H; ++k) {
#pragma unroll
      for (int i = 0; i < NX; ++i) { const double e = x[i] - goal[i]; Jc = fma(Qs[i] * e, e, Jc); }
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const double d = z[k * NU + i] - (k == 0 ? uprev[i] : z[(k - 1) * NU + i]);
        Jc = fma(Rs[i] * d, d, Jc);
        if constexpr (OD) { const double e = z[k * NU + i] - ut[i]; Jc = fma(Ra[i] * e, e, Jc); }
      }
      double y[NY], F[NX], a, b, c, d;
#pragma unroll
      for (int i = 0; i < NX; ++i) y[i] = x[i];
#pragma unroll
      for (int i = 0; i < NU; ++i) y[NX + i] = z[k * NU + i];
      TrigCompute trig;
      Mod::stage(p, w + L.AUX, y, F, a, b, c, d, trig);
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = F[i]; xs[(k + 1) * NX + i] = F[i]; }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) { const double e = x[i] - goal[i]; Jc = fma(Qs[i] * e, e, Jc); }
    return Jc;
  }

  // ---- general rows (Mod::GENERAL): one (stage, obstacle) row evaluated with any jet flavour -------------------
  // c = sum_i w_i h(S_i; o_j):  (w0, w1) = (alpha - 1, 1) for relative degree 1, (w0, w1, w2) for 2 (mpc_cbf.py:312-321)
  template <class T>
  SCB_HD void general_row(const T (*S)[NX], const T* SN, const T* CS, const double* ob, T& c) const {
    if constexpr (OD) {
      // dd_h + (alpha1 omega1 + alpha2 omega2) d_h + alpha1 alpha2 omega1 omega2 h_k   (optimal_decay_mpc_cbf.py:296-300)
      T h0, h1, h2, sg, tt, dh, ddh, tmp;
      Mod::hfun(p, S[0], SN[0], CS[0], ob, h0);
      Mod::hfun(p, S[1], SN[1], CS[1], ob, h1);
      Mod::hfun(p, S[2], SN[2], CS[2], ob, h2);
      const T& o1 = SN[0]; const T& o2 = CS[0];
      jscale(sg, o1, p.alpha1); jaxpy(sg, sg, p.alpha2, o2);
      jmul(tt, o1, o2); jscale(tt, tt, p.alpha1 * p.alpha2);
      jaxpy(dh, h1, -1.0, h0);
      jaxpy(ddh, h2, -2.0, h1); jadd(ddh, ddh, h0);
      jmul(tmp, sg, dh); jadd(c, ddh, tmp);
      jmul(tmp, tt, h0); jadd(c, c, tmp);
      return;
    }
    T h0, h1;
    Mod::hfun(p, S[0], SN[0], CS[0], ob, h0);
    Mod::hfun(p, S[1], SN[1], CS[1], ob, h1);
    jscale(c, h0, w0);
    jaxpy(c, c, w1, h1);
    if constexpr (Mod::REL == 2) {
      T h2;
      Mod::hfun(p, S[2], SN[2], CS[2], ob, h2);
      jaxpy(c, c, w2, h2);
    }
  }

  // barrier points of every stage at (xs, z) -> w[L.PT]; CBF values -> dst[H*M]
  SCB_MPC_PHASE void points_and_cbf(const double* z, const double* xs, double* dst) const {
    if constexpr (Mod::GENERAL) {
      // the stage's barrier states S_0, S_1 and their heading sin/cos (plain values) -> scratch at the END of the
      // gradient block of each stage is not available here (trial points): recompute per (stage, obstacle) pair
      SCB_LANE_UNROLL
      for (int t = lane; t < H * M; t += LANES) {
        const int k = t / M, j = t - k * M;
        double y[NY], F[NX], S[NPT][NX], SN[NPT], CS[NPT], c;
#pragma unroll
        for (int i = 0; i < NX; ++i) y[i] = xs[k * NX + i];
#pragma unroll
        for (int i = 0; i < NU; ++i) y[NX + i] = z[k * NU + i];
        TrigCompute trig;
        Mod::states(p, y, F, S, SN, CS, trig);
        general_row(S, SN, CS, w + L.OBS7 + j * 7, c);
        dst[t] = c;
      }
      sync();
      return;
    } else {
    SCB_LANE_UNROLL
    for (int k = lane; k < H; k += LANES) {
      double y[NY], F[NX], P1, Q1, P2, Q2;
#pragma unroll
      for (int i = 0; i < NX; ++i) y[i] = xs[k * NX + i];
#pragma unroll
      for (int i = 0; i < NU; ++i) y[NX + i] = z[k * NU + i];
      TrigCompute trig;
      Mod::stage(p, w + L.AUX, y, F, P1, Q1, P2, Q2, trig);
      double* pt = w + L.PT + k * 6;
      pt[0] = y[0]; pt[1] = y[1]; pt[2] = P1; pt[3] = Q1; pt[4] = P2; pt[5] = Q2;
    }
    sync();
    SCB_LANE_UNROLL
    for (int t = lane; t < H * M; t += LANES) {
      const int k = t / M, j = t - k * M;
      const double* pt = w + L.PT + k * 6;
      const double* ob = w + L.OB + j * 3;
      double v = -Wsum * ob[2];
      double dx = pt[0] - ob[0], dy = pt[1] - ob[1];
      v = fma(w0, dx * dx + dy * dy, v);
      dx = pt[2] - ob[0]; dy = pt[3] - ob[1];
      v = fma(w1, dx * dx + dy * dy, v);
      dx = pt[4] - ob[0]; dy = pt[5] - ob[1];
      v = fma(w2, dx * dx + dy * dy, v);
      dst[t] = v;
    }
    sync();
    }
  }

  SCB_HD double simple_value(const SimpleCon& c, const double* z, const double* xs) const {
    const double yv = (c.var < NX) ? xs[c.k * NX + c.var] : z[c.k * NU + (c.var - NX)];
    return fma(c.sgn, yv, c.off);
  }

  // ---- first derivatives of every stage at the current iterate (lanes over stages) ----
  // A_k, B_k and the gradients gE, gX, gY (see header).  Curvature is NOT stored: stage_hessians() re-evaluates
  // the stage with second-order jets once the multipliers and costates that weight it are known and keeps only
  // the contracted 21-entry stage Hessian (saves 4 NX + 3 Hessians per stage of workspace).
  template <class JT>
  SCB_HD void barrier_jets(const JT* y, const JT& P1, const JT& Q1, const JT& P2, const JT& Q2, JT& E, JT& PX, JT& PY) const {
    // E = sum_i w_i (P_i^2 + Q_i^2),  PX = 2 sum w_i P_i,  PY = 2 sum w_i Q_i
    JT t;
    jmul(E, y[0], y[0]); jmul(t, y[1], y[1]); jaxpy(E, E, 1.0, t); jscale(E, E, w0);
    jmul(t, P1, P1); jaxpy(E, E, w1, t); jmul(t, Q1, Q1); jaxpy(E, E, w1, t);
    jmul(t, P2, P2); jaxpy(E, E, w2, t); jmul(t, Q2, Q2); jaxpy(E, E, w2, t);
    jscale(PX, y[0], 2.0 * w0); jaxpy(PX, PX, 2.0 * w1, P1); jaxpy(PX, PX, 2.0 * w2, P2);
    jscale(PY, y[1], 2.0 * w0); jaxpy(PY, PY, 2.0 * w1, Q1); jaxpy(PY, PY, 2.0 * w2, Q2);
  }

  SCB_MPC_PHASE void stage_derivatives() {
    const double* xs = w + L.X;
    const double* z = w + L.Z;
    if constexpr (Mod::LINEAR) {
      // linear model: A, B constant; barrier points affine in y
      const double* aux = w + L.AUX;
      const double* R0 = aux + NX * NX + NX * NU;
      const double* R1 = R0 + NY;
      SCB_LANE_UNROLL
      for (int k = lane; k < H; k += LANES) {
        double y[NY];
#pragma unroll
        for (int i = 0; i < NX; ++i) y[i] = xs[k * NX + i];
#pragma unroll
        for (int i = 0; i < NU; ++i) y[NX + i] = z[k * NU + i];
        // (A_k, B_k) = (Ae, Be) of the AUX block for every stage: L.A / L.B alias it with stride 0
        double P1 = 0.0, Q1 = 0.0;
        for (int i = 0; i < NY; ++i) { P1 = fma(R0[i], y[i], P1); Q1 = fma(R1[i], y[i], Q1); }
        double* je = w + L.JE + k * NY;
        double* jx = w + L.JX + k * NY;
        double* jy = w + L.JY + k * NY;
        for (int i = 0; i < NY; ++i) {
          const double d0 = (i == 0) ? 1.0 : 0.0, d1 = (i == 1) ? 1.0 : 0.0;
          je[i] = 2.0 * w0 * (y[0] * d0 + y[1] * d1) + 2.0 * w1 * (P1 * R0[i] + Q1 * R1[i]);
          jx[i] = 2.0 * w0 * d0 + 2.0 * w1 * R0[i];
          jy[i] = 2.0 * w0 * d1 + 2.0 * w1 * R1[i];
        }
      }
      sync();
    } else if constexpr (Mod::GENERAL) {
      // pass 0 (lanes over stages): sin/cos of the headings of S_0 and S_1 -> trig cache
      SCB_LANE_UNROLL
      for (int k = lane; k < H; k += LANES) {
        double y[NY], F[NX], S[NPT][NX], SN[NPT], CS[NPT];
#pragma unroll
        for (int i = 0; i < NX; ++i) y[i] = xs[k * NX + i];
#pragma unroll
        for (int i = 0; i < NU; ++i) y[NX + i] = z[k * NU + i];
        TrigStore trig(w + L.TR + k * L.TRS);
        Mod::states(p, y, F, S, SN, CS, trig);
      }
      sync();
      // pass 1 (lanes over (stage, variable)): column i of A_k / B_k and d/dy_i of EVERY row of the stage
      SCB_LANE_UNROLL
      for (int t = lane; t < H * NY; t += LANES) {
        const int k = t / NY, i = t - k * NY;
        JetG y[NY], F[NX], S[NPT][NX], SN[NPT], CS[NPT];
#pragma unroll
        for (int m = 0; m < NX; ++m) jvar_entry(y[m], xs[k * NX + m], m, i);
#pragma unroll
        for (int m = 0; m < NU; ++m) jvar_entry(y[NX + m], z[k * NU + m], NX + m, i);
        TrigLoad trig(w + L.TR + k * L.TRS);
        Mod::states(p, y, F, S, SN, CS, trig);
        double* A = w + L.A + k * L.AS;
        double* B = w + L.B + k * L.BS;
        if (i < NX) {
#pragma unroll
          for (int c = 0; c < NX; ++c) A[c * NX + i] = F[c].g;
        } else {
#pragma unroll
          for (int c = 0; c < NX; ++c) B[c * NU + (i - NX)] = F[c].g;
        }
        SCB_LOOP
        for (int j = 0; j < M; ++j) {
          JetG c;
          general_row(S, SN, CS, w + L.OBS7 + j * 7, c);
          w[L.GR + (k * M + j) * NY + i] = c.g;
        }
      }
      sync();
    } else {
      // pass 0 (lanes over stages): the stage's sin/cos at the current iterate -> trig cache
      if constexpr (Mod::NTRIG > 0) {
        SCB_LANE_UNROLL
        for (int k = lane; k < H; k += LANES) {
          double y[NY], F[NX], P1, Q1, P2, Q2;
#pragma unroll
          for (int i = 0; i < NX; ++i) y[i] = xs[k * NX + i];
#pragma unroll
          for (int i = 0; i < NU; ++i) y[NX + i] = z[k * NU + i];
          TrigStore trig(w + L.TR + k * L.TRS);
          Mod::stage(p, w + L.AUX, y, F, P1, Q1, P2, Q2, trig);
        }
        sync();
      }
      // pass 1 (lanes over (stage, variable) pairs): column i of A_k / B_k and entry i of gE, gX, gY by entry jets
      SCB_LANE_UNROLL
      for (int t = lane; t < H * NY; t += LANES) {
        const int k = t / NY, i = t - k * NY;
        JetG y[NY], F[NX], P1, Q1, P2, Q2;
#pragma unroll
        for (int m = 0; m < NX; ++m) jvar_entry(y[m], xs[k * NX + m], m, i);
#pragma unroll
        for (int m = 0; m < NU; ++m) jvar_entry(y[NX + m], z[k * NU + m], NX + m, i);
        TrigLoad trig(w + L.TR + k * L.TRS);
        Mod::stage(p, w + L.AUX, y, F, P1, Q1, P2, Q2, trig);
        double* A = w + L.A + k * L.AS;
        double* B = w + L.B + k * L.BS;
        if (i < NX) {
#pragma unroll
          for (int c = 0; c < NX; ++c) A[c * NX + i] = F[c].g;
        } else {
#pragma unroll
          for (int c = 0; c < NX; ++c) B[c * NU + (i - NX)] = F[c].g;
        }
        JetG E, PX, PY;
        barrier_jets(y, P1, Q1, P2, Q2, E, PX, PY);
        w[L.JE + t] = E.g; w[L.JX + t] = PX.g; w[L.JY + t] = PY.g;
      }
      sync();
    }
  }

  // gradient of the input-rate term sum R (u_k - u_{k-1})^2 at z -> out[n]
  SCB_MPC_PHASE void rate_gradient(const double* z, double* out) const {
    SCB_LANE_UNROLL
    for (int t = lane; t < n; t += LANES) {
      const int k = t / NU, i = t - k * NU;
      const double um = (k == 0) ? uprev[i] : z[(k - 1) * NU + i];
      double g = 2.0 * Rs[i] * (z[t] - um);
      if (k + 1 < H) g -= 2.0 * Rs[i] * (z[(k + 1) * NU + i] - z[t]);
      if constexpr (OD) g += 2.0 * Ra[i] * (z[t] - ut[i]);
      out[t] = g;
    }
    sync();
  }

  // adjoint sweep: stage gradients gam[(H+1)*NY] -> z-gradient out[n]; costates -> w[L.MU] (lane 0)
  SCB_MPC_PHASE void adjoint(const double* gam, double* out) {
    if (lane == 0) {
      double mu[NX];
      double* MU = w + L.MU;
#pragma unroll
      for (int i = 0; i < NX; ++i) { mu[i] = gam[H * NY + i]; MU[H * NX + i] = mu[i]; }
      SCB_LOOP
      for (int k = H - 1; k >= 0; --k) {
        const double* A = w + L.A + k * L.AS;
        const double* B = w + L.B + k * L.BS;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = gam[k * NY + NX + i];
#pragma unroll
          for (int c = 0; c < NX; ++c) v = fma(B[c * NU + i], mu[c], v);
          out[k * NU + i] = v;
        }
        double m2[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double v = gam[k * NY + i];
#pragma unroll
          for (int c = 0; c < NX; ++c) v = fma(A[c * NX + i], mu[c], v);
          m2[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) { mu[i] = m2[i]; MU[k * NX + i] = m2[i]; }
      }
    }
    sync();
  }

  // weighted obstacle sums of stage k: which = 0 (sigma moments, 6) / 1 (lambda, 3) / 2 (rhs weights, 3)
  // segmented xor-shuffle sum over the `seg` (power of two) consecutive lanes this lane belongs to
  static SCB_HD double seg_sum(double v, int seg) {
#if defined(__CUDA_ARCH__)
    SCB_LOOP
    for (int o = seg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(G::gmask(), v, o, 32);
#endif
    (void)seg;
    return v;
  }

  // mode 0: lambda sums only (dual residual / costates); 1: + sigma moments and Newton-rhs weights
  SCB_MPC_PHASE void stage_sums(double mu_bar, bool with_rhs, double floor_s = 0.0) {
    if constexpr (Mod::GENERAL) { floor_cur = floor_s; return; }     // general rows: no sums structure, see stage_gradients / stage_hessians
    // A segment of `seg` lanes (smallest power of two >= M, capped at LANES) owns one stage at a time, so
    // LANES/seg stages are processed per pass and each of the <= 12 partial sums needs log2(seg) shuffle steps.
    int seg = 1;
    while (seg < M && seg < LANES) seg <<= 1;
    const int spp = LANES / seg;                       // stages per pass
    const int sub = lane / seg, jl = lane - sub * seg;
    SCB_LOOP
    for (int k0 = 0; k0 < H; k0 += spp) {
      const int k = k0 + sub;
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, l0 = 0, l1 = 0, l2 = 0, r0 = 0, r1 = 0, r2 = 0;
      if (k < H) {
        SCB_LOOP
        for (int j = jl; j < M; j += seg) {
          const double ox = w[L.OB + j * 3], oy = w[L.OB + j * 3 + 1];
          const double lam = w[L.L + k * M + j];
          l0 += lam; l1 = fma(lam, ox, l1); l2 = fma(lam, oy, l2);
          if (with_rhs) {
            const double inv = w[L.DS + k * M + j], sig = lam * inv;
            a0 += sig; a1 = fma(sig, ox, a1); a2 = fma(sig, oy, a2);
            a3 = fma(sig * ox, ox, a3); a4 = fma(sig * ox, oy, a4); a5 = fma(sig * oy, oy, a5);
            const double g = w[L.C + k * M + j];
            const double wt = mu_bar * inv - sig * (g - fmax(g, floor_s));
            r0 += wt; r1 = fma(wt, ox, r1); r2 = fma(wt, oy, r2);
          }
        }
      }
      l0 = seg_sum(l0, seg); l1 = seg_sum(l1, seg); l2 = seg_sum(l2, seg);
      if (with_rhs) {
        a0 = seg_sum(a0, seg); a1 = seg_sum(a1, seg); a2 = seg_sum(a2, seg);
        a3 = seg_sum(a3, seg); a4 = seg_sum(a4, seg); a5 = seg_sum(a5, seg);
        r0 = seg_sum(r0, seg); r1 = seg_sum(r1, seg); r2 = seg_sum(r2, seg);
      }
      if (jl == 0 && k < H) {
        double* sm = w + L.SUM + k * 12;
        sm[6] = l0; sm[7] = l1; sm[8] = l2;
        if (with_rhs) {
          sm[0] = a0; sm[1] = a1; sm[2] = a2; sm[3] = a3; sm[4] = a4; sm[5] = a5;
          sm[9] = r0; sm[10] = r1; sm[11] = r2;
        }
      }
    }
    sync();
  }

  // stage gradient for the dual residual / costates:  gam = grad l_k - sum lam grad g
  // (sign = +1)  or for the Newton rhs:  gam = -grad l_k + sum wt grad g   (sign = -1, weights 9..11)
  SCB_MPC_PHASE void stage_gradients(bool rhs, double mu_bar) {
    double* gam = w + L.GAM;
    const double* xs = w + L.X;
    SCB_LANE_UNROLL
    for (int t = lane; t < (H + 1) * NY; t += LANES) {
      const int k = t / NY, i = t - k * NY;
      double v = 0.0;
      if (i < NX) v = 2.0 * Qs[i] * (xs[k * NX + i] - goal[i]);
      if (rhs) v = -v;
      if (k < H) {
        double cg;                                                    // sum_j wt_j grad c_j
        if constexpr (Mod::GENERAL) {
          cg = 0.0;
          SCB_LOOP
          for (int j = 0; j < M; ++j) {
            const double lam = w[L.L + k * M + j];
            double wt = lam;
            if (rhs) {
              const double inv = w[L.DS + k * M + j], g = w[L.C + k * M + j];
              wt = mu_bar * inv - lam * inv * (g - fmax(g, floor_cur));
            }
            cg = fma(wt, w[L.GR + (k * M + j) * NY + i], cg);
          }
        } else {
          const double* sm = w + L.SUM + k * 12 + (rhs ? 9 : 6);
          const double ge = w[L.JE + k * NY + i], gx = w[L.JX + k * NY + i], gy = w[L.JY + k * NY + i];
          cg = sm[0] * ge - sm[1] * gx - sm[2] * gy;
        }
        v += rhs ? cg : -cg;
      }
      gam[t] = v;
    }
    sync();
    // simple bounds
    SCB_LANE_UNROLL
    for (int q = lane; q < L.NS; q += LANES) {
      const SimpleCon c = decode_simple<Mod>(p, H, q);
      const double s = w[L.SS + q], lam = w[L.SL + q], is = w[L.SDS + q];
      double wt;
      if (rhs) {
        const double g = simple_value(c, w + L.Z, w + L.X);
        wt = mu_bar * is - (lam * is) * (g - s);
      } else {
        wt = -lam;
      }
      // distinct q may hit the same entry (upper/lower of one variable): serialise per group
      atomic_add_ws(gam + c.k * NY + c.var, wt * c.sgn);
    }
    sync();
  }

  static SCB_HD void atomic_add_ws(double* addr, double v) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) atomicAdd(addr, v); else *addr += v;
#else
    *addr += v;
#endif
  }

  // stage Hessians G_k = hess l_k - sum lam hess c + sum sigma grad c grad c' + bound sigmas + costate curvature
  SCB_MPC_PHASE void stage_hessians() {
    double* Gm = w + L.G;
    // All curvature of stage k is the Hessian of ONE scalar function of y,
    //   Psi_k = sum_c mu_{k+1,c} F_c(y) - Lam0 E(y) + LamX PX(y) + LamY PY(y)      (skipped in Gauss-Newton mode).
    // Lanes over (stage, packed Hessian entry) pairs: the curvature entry comes from an ENTRY jet (4 doubles per
    // quantity instead of a 28-double full jet, scb_jet.cuh) replaying the stage's cached sin/cos; the cost Hessian,
    // the barrier terms sum sigma grad c grad c' and the linear model's constant curvature are added in the same pass.
    SCB_LANE_UNROLL
    for (int t = lane; t < (H + 1) * NH; t += LANES) {
      const int k = t / NH, e = t - k * NH;
      // unpack (i, j) of packed entry e
      int i = 0, rem = e;
      while (rem >= NY - i) { rem -= NY - i; ++i; }
      const int j = i + rem;
      double v = 0.0;
      if constexpr (Mod::GENERAL) {
        if (k < H) {
          JetH y[NY], F[NX], S[NPT][NX], SN[NPT], CS[NPT];
#pragma unroll
          for (int m = 0; m < NX; ++m) jvar_entry(y[m], w[L.X + k * NX + m], m, i, j);
#pragma unroll
          for (int m = 0; m < NU; ++m) jvar_entry(y[NX + m], w[L.Z + k * NU + m], NX + m, i, j);
          TrigLoad trig(w + L.TR + k * L.TRS);
          Mod::states(p, y, F, S, SN, CS, trig);
          if (!gauss_newton) {
            const double* mu = w + L.MU + (k + 1) * NX;
#pragma unroll
            for (int c = 0; c < NX; ++c) v = fma(mu[c], F[c].h, v);
          }
          SCB_LOOP
          for (int jo = 0; jo < M; ++jo) {
            JetH c;
            general_row(S, SN, CS, w + L.OBS7 + jo * 7, c);
            const double lam = w[L.L + k * M + jo], sig = lam * w[L.DS + k * M + jo];
            v = fma(sig * c.gi, c.gj, v);                              // + sigma grad c grad c'
            if (!gauss_newton) v = fma(-lam, c.h, v);                  // - lambda hess c
          }
        }
        if (i == j && i < NX) v += 2.0 * Qs[i];
        Gm[t] = v;
        continue;
      }
      if constexpr (!Mod::LINEAR) {
        if (k < H && !gauss_newton) {
          JetH y[NY], F[NX], P1, Q1, P2, Q2, E, PX, PY;
#pragma unroll
          for (int m = 0; m < NX; ++m) jvar_entry(y[m], w[L.X + k * NX + m], m, i, j);
#pragma unroll
          for (int m = 0; m < NU; ++m) jvar_entry(y[NX + m], w[L.Z + k * NU + m], NX + m, i, j);
          TrigLoad trig(w + L.TR + k * L.TRS);
          Mod::stage(p, w + L.AUX, y, F, P1, Q1, P2, Q2, trig);
          barrier_jets(y, P1, Q1, P2, Q2, E, PX, PY);
          const double* sm = w + L.SUM + k * 12;
          const double* mu = w + L.MU + (k + 1) * NX;
          v = -sm[6] * E.h + sm[7] * PX.h + sm[8] * PY.h;
#pragma unroll
          for (int c = 0; c < NX; ++c) v = fma(mu[c], F[c].h, v);
        }
      }
      if (i == j && i < NX) v += 2.0 * Qs[i];
      if (k < H) {
        const double* sm = w + L.SUM + k * 12;
        const double* je = w + L.JE + k * NY;
        const double* jx = w + L.JX + k * NY;
        const double* jy = w + L.JY + k * NY;
        if (Mod::LINEAR && !gauss_newton) {
          // hess E is constant: 2 w0 (e0 e0' + e1 e1') + 2 w1 (R0 R0' + R1 R1');  PX, PY are linear
          const double* R0 = w + L.AUX + NX * NX + NX * NU;
          const double* R1 = R0 + NY;
          const double d = (i == j && i < 2) ? 1.0 : 0.0;
          v -= sm[6] * (2.0 * w0 * d + 2.0 * w1 * (R0[i] * R0[j] + R1[i] * R1[j]));
        }
        // + sum_j sigma_j grad c_j grad c_j'
        const double ei = je[i], ej = je[j], xi = jx[i], xj = jx[j], yi = jy[i], yj = jy[j];
        v += sm[0] * ei * ej - sm[1] * (ei * xj + xi * ej) - sm[2] * (ei * yj + yi * ej) + sm[3] * xi * xj +
             sm[4] * (xi * yj + yi * xj) + sm[5] * yi * yj;
      }
      Gm[t] = v;
    }
    sync();
    SCB_LANE_UNROLL
    for (int q = lane; q < L.NS; q += LANES) {
      const SimpleCon c = decode_simple<Mod>(p, H, q);
      atomic_add_ws(Gm + c.k * NH + hidx<NY>(c.var, c.var), w[L.SL + q] * w[L.SDS + q]);
    }
    sync();
  }

  // ---- Newton step by a stage-wise (Riccati) factorisation ------------------------------------------------
  // The Newton system of the condensed problem is the optimality system of the equality-constrained QP
  //     min  sum_k 1/2 v_k' Hs_k v_k + h_k' v_k ,   v_k = (dxt_k, du_k),   dxt_{k+1} = Fm_k v_k,  dxt_0 = 0
  // over the augmented state xt_k = (x_k, u_{k-1}) (the input-RATE cost couples neighbouring inputs), with
  // Hs_k = stage Hessian G_k + rate terms, Fm_k = [[A_k 0 B_k],[0 0 I]].  The backward Riccati sweep is the block
  // LDL' of that system: it is positive definite iff every Muu_k is, costs O(H (nx+nu)^3) instead of the O(H^2)
  // sensitivities + dense n x n Cholesky of a condensed factorisation, and needs O(H) scratch.
  static constexpr int NXT = NX + NU, NV = NXT + NU;

  // y-index (x, u) of the stage variable v-index, or -1 for the u_{k-1} block
  static SCB_HD int v2y(int c) { return c < NX ? c : (c < NXT ? -1 : NX + (c - NXT)); }

  // backward sweep with diagonal shift `delta` on Muu; false if some Muu_k is not positive definite.
  // Lane c owns COLUMN c of the stage matrix M_k = Hs_k + F_k' P_{k+1} F_k (c < NV) and lane NV the vector
  // m_k = h_k + F_k' p_{k+1}: t_c = P f_c and M[:, c] = Hs[:, c] + F' t_c need no cross-lane data, so a stage is
  // two barriers (publish the input columns of M; publish P_k, K_k), everything else lives in registers.
  // Hs_k[b][c]: stage Hessian G_k on the (x, u_k) rows/cols + the input-rate terms coupling u_{k-1} and u_k (+ shift)
  SCB_HD double hs_entry(int k, int b, int c, double delta) const {
    const int yb = v2y(b), yc = v2y(c);
    const int ub = (b >= NXT) ? b - NXT : (b >= NX ? b - NX : -1), uc = (c >= NXT) ? c - NXT : (c >= NX ? c - NX : -1);
    double v = 0.0;
    if (yb >= 0 && yc >= 0) { const int lo = yb < yc ? yb : yc, hi = yb < yc ? yc : yb; v = w[L.G + k * NH + hidx<NY>(lo, hi)]; }
    if (ub >= 0 && ub == uc) {
      double rr = 0.0;
#pragma unroll
      for (int i = 0; i < NU; ++i) if (i == ub) rr = 2.0 * Rs[i];
      v += ((b >= NXT) == (c >= NXT)) ? rr : -rr;
      if constexpr (OD) {
        if (b == c && b >= NXT) {
#pragma unroll
          for (int i = 0; i < NU; ++i) if (i == ub) v += 2.0 * Ra[i];
        }
      }
    }
    if (b == c && b >= NXT) v += delta;
    return v;
  }
  // h_k[b] = -(Newton rhs of the stage) + rate-term gradient
  SCB_HD double hv_entry(int k, int b) const {
    const int yb = v2y(b);
    const int ub = (b >= NXT) ? b - NXT : (b >= NX ? b - NX : -1);
    double v = (yb >= 0) ? -w[L.GAM + k * NY + yb] : 0.0;
    if (ub >= 0) {
      const double* z = w + L.Z;
      double rr = 0.0, du = 0.0;
#pragma unroll
      for (int i = 0; i < NU; ++i)
        if (i == ub) { rr = 2.0 * Rs[i]; du = z[k * NU + i] - (k == 0 ? uprev[i] : z[(k - 1) * NU + i]); }
      v += (b >= NXT) ? rr * du : -rr * du;
      if constexpr (OD) {
        if (b >= NXT) {
#pragma unroll
          for (int i = 0; i < NU; ++i) if (i == ub) v += 2.0 * Ra[i] * (z[k * NU + i] - ut[i]);
        }
      }
    }
    return v;
  }

  SCB_MPC_PHASE bool riccati_backward(double delta) {
    double* PM = w + L.PM;
    double* PV = w + L.PV;
    const double* gam = w + L.GAM;                  // = -grad l_k + sum w grad g  (the Newton rhs per stage)
    const double* z = w + L.Z;
    // terminal: P_H = G_H (x block), p_H = -gam_H
    SCB_LANE_UNROLL
    for (int t = lane; t < NXT * NXT; t += LANES) {
      const int r = t / NXT, c = t - r * NXT;
      double v = 0.0;
      if (r < NX && c < NX) { const int lo = r < c ? r : c, hi = r < c ? c : r; v = w[L.G + H * NH + hidx<NY>(lo, hi)]; }
      PM[(H & 1) * NXT * NXT + t] = v;
    }
    SCB_LANE_UNROLL
    for (int t = lane; t < NXT; t += LANES) PV[(H & 1) * NXT + t] = (t < NX) ? -gam[H * NY + t] : 0.0;
    sync();
    bool ok = true;
    double* MU_ = w + L.MM;                         // published input columns: MU_[i * NV + r] = M[r][NXT + i]
    SCB_LOOP
    for (int k = H - 1; k >= 0; --k) {
      const double* Pn = PM + ((k + 1) & 1) * NXT * NXT;
      const double* pn = PV + ((k + 1) & 1) * NXT;
      const double* A = w + L.A + k * L.AS;         // A[a * NX + c]
      const double* B = w + L.B + k * L.BS;         // B[a * NU + i]
      // lane c: t = P f_c (or p), M[:, c] = Hs[:, c] + F' t  (or m = h + F' p), with F_k = [[A 0 B], [0 0 I]] used
      // block-wise (no dense copy of F_k or Hs_k: a lane builds the 8 entries of its own column of Hs_k on the fly)
      double Mc[NV];
#pragma unroll
      for (int b2 = 0; b2 < NV; ++b2) Mc[b2] = 0.0;
      SCB_LOOP
      for (int c = lane; c <= NV; c += LANES) {     // (one pass: NV + 1 <= LANES on the device; sequential on the host)
        const bool isvec = (c == NV);
        const bool isx = (c < NX), isu = (c >= NXT && !isvec);
        const int ci = isu ? c - NXT : 0, cx = isx ? c : 0;
        double fcol[NX];                             // rows < NX of column c of F_k
#pragma unroll
        for (int a2 = 0; a2 < NX; ++a2) fcol[a2] = isx ? A[a2 * NX + cx] : (isu ? B[a2 * NU + ci] : 0.0);
        double tc[NXT];
#pragma unroll
        for (int r = 0; r < NXT; ++r) {
          double v = isu ? Pn[r * NXT + NX + ci] : 0.0;
#pragma unroll
          for (int a2 = 0; a2 < NX; ++a2) v = fma(Pn[r * NXT + a2], fcol[a2], v);
          tc[r] = isvec ? pn[r] : v;
        }
#pragma unroll
        for (int b2 = 0; b2 < NV; ++b2) {
          double v = isvec ? hv_entry(k, b2) : hs_entry(k, b2, c, delta);
          if (b2 < NX) {
#pragma unroll
            for (int a2 = 0; a2 < NX; ++a2) v = fma(A[a2 * NX + b2], tc[a2], v);
          } else if (b2 >= NXT) {
            v += tc[NX + (b2 - NXT)];
#pragma unroll
            for (int a2 = 0; a2 < NX; ++a2) v = fma(B[a2 * NU + (b2 - NXT)], tc[a2], v);
          }
          Mc[b2] = v;
        }
        if (c >= NXT && !isvec) {
#pragma unroll
          for (int r = 0; r < NV; ++r) MU_[(c - NXT) * NV + r] = Mc[r];
        }
        if (LANES == 1) {                            // host-sim: one lane plays all columns -> stash them
#pragma unroll
          for (int r = 0; r < NV; ++r) w[L.PT2 + c * NV + r] = Mc[r];
        }
      }
      sync();
      // Muu = L L' (NU x NU) from the published input columns, every lane redundantly
      double Lu[NU][NU];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
#pragma unroll
        for (int j = 0; j < NU; ++j) Lu[i][j] = 0.0;
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double d = MU_[j * NV + NXT + j];
#pragma unroll
        for (int t = 0; t < NU; ++t) if (t < j) d -= Lu[j][t] * Lu[j][t];
        if (!(d > 1e-300) || !(d < 1e300)) ok = false;
        const double inv = ok ? rsqrt_pos(d) : 0.0;
        Lu[j][j] = inv;                               // diagonal holds 1 / L_jj
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          if (i > j) {
            double v = MU_[j * NV + NXT + i];
#pragma unroll
            for (int t = 0; t < NU; ++t) if (t < j) v -= Lu[i][t] * Lu[j][t];
            Lu[i][j] = v * inv;
          }
        }
      }
      if (!ok) break;
      // gains for this lane's column (K[:, c] or kff), then its column of P_k (or p_k)
      double* Kg = w + L.KG + k * NU * NXT;
      double* Kf = w + L.KF + k * NU;
      double* Pk = PM + (k & 1) * NXT * NXT;
      double* pk = PV + (k & 1) * NXT;
      SCB_LANE_UNROLL
      for (int c = lane; c <= NV; c += LANES) {
        if (c >= NXT && c < NV) continue;            // input columns carry no gain
        if (LANES == 1) {
#pragma unroll
          for (int r = 0; r < NV; ++r) Mc[r] = w[L.PT2 + c * NV + r];
        }
        double rhs[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) rhs[i] = -Mc[NXT + i];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = rhs[i];
#pragma unroll
          for (int t = 0; t < NU; ++t) if (t < i) v -= Lu[i][t] * rhs[t];
          rhs[i] = v * Lu[i][i];
        }
#pragma unroll
        for (int ii = NU - 1; ii >= 0; --ii) {
          double v = rhs[ii];
#pragma unroll
          for (int t = 0; t < NU; ++t) if (t > ii) v -= Lu[t][ii] * rhs[t];
          rhs[ii] = v * Lu[ii][ii];
        }
        const bool isvec = (c == NV);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          if (isvec) Kf[i] = rhs[i]; else Kg[i * NXT + c] = rhs[i];
        }
#pragma unroll
        for (int r = 0; r < NXT; ++r) {
          double v = Mc[r];
#pragma unroll
          for (int i = 0; i < NU; ++i) v = fma(MU_[i * NV + r], rhs[i], v);
          if (isvec) pk[r] = v; else Pk[r * NXT + c] = v;
        }
      }
      sync();
    }
    sync();
    return ok;
  }

  // forward sweep: dz (inputs) and the stage directions dy_k = (dx_k, du_k)   (lane 0; O(H (nx+nu)^2))
  SCB_MPC_PHASE void riccati_forward() {
    if (lane == 0) {
      double dxt[NXT];
#pragma unroll
      for (int i = 0; i < NXT; ++i) dxt[i] = 0.0;
      SCB_LOOP
      for (int k = 0; k < H; ++k) {
        const double* Kg = w + L.KG + k * NU * NXT;
        const double* Kf = w + L.KF + k * NU;
        double du[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double v = Kf[i];
#pragma unroll
          for (int c = 0; c < NXT; ++c) v = fma(Kg[i * NXT + c], dxt[c], v);
          du[i] = v;
          w[L.DZ + k * NU + i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) w[L.DY + k * NY + i] = dxt[i];
#pragma unroll
        for (int i = 0; i < NU; ++i) w[L.DY + k * NY + NX + i] = du[i];
        double nx_[NX];
        const double* A = w + L.A + k * L.AS;
        const double* B = w + L.B + k * L.BS;
#pragma unroll
        for (int a = 0; a < NX; ++a) {
          double v = 0.0;
#pragma unroll
          for (int c = 0; c < NX; ++c) v = fma(A[a * NX + c], dxt[c], v);
#pragma unroll
          for (int c = 0; c < NU; ++c) v = fma(B[a * NU + c], du[c], v);
          nx_[a] = v;
        }
#pragma unroll
        for (int a = 0; a < NX; ++a) dxt[a] = nx_[a];
#pragma unroll
        for (int i = 0; i < NU; ++i) dxt[NX + i] = du[i];
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) w[L.DY + H * NY + i] = dxt[i];
#pragma unroll
      for (int i = 0; i < NU; ++i) w[L.DY + H * NY + NX + i] = 0.0;
    }
    sync();
  }

  // ------------------------------------------------------------------------------------------
  // Problem data + cold start: goal padding, obstacle table, z = u_prev, rollout and the CBF values of every
  // (stage, obstacle) pair at the cold start (-> w[L.C]).  Returns the unscaled objective.  Split from solve() so
  // the CPU host-sim can compare the kernel's own statement (Euler map, stage cost, CBF rows) with the values the
  // reference's mpc_cbf.py hands to do-mpc (tests/golden/ref_mpc_statement.npz).
  SCB_HD double init(int nobs, const double* x0, const double* goal_in, int ngoal, const double* up, const double* obs,
                     bool push_start = true) {
#pragma unroll
    for (int i = 0; i < NX; ++i) goal[i] = (i < ngoal) ? ld(goal_in + i) : 0.0;      // goal padded with zeros (mpc_cbf.py:267)
#pragma unroll
    for (int i = 0; i < NU; ++i) uprev[i] = ld(up + i);
    const double beta = Mod::beta();
    if constexpr (Mod::LINEAR) {
      if (lane == 0) Mod::setup_aux(p, w + L.AUX);
      sync();
    }
    // obstacles: (ox, oy, beta d^2); missing slots = the reference's dummy [1000, 1000, 0, ...] (mpc_cbf.py:346-364)
    SCB_LANE_UNROLL
    for (int j = lane; j < M; j += LANES) {
      if constexpr (Mod::GENERAL) {
        // raw rows; missing slots = the reference's dummy [1000, 1000, 0, 0, 0, 0, 0]
#pragma unroll
        for (int q = 0; q < 7; ++q) w[L.OBS7 + j * 7 + q] = (j < nobs) ? ld(obs + j * 7 + q) : (q < 2 ? 1000.0 : 0.0);
      } else {
        double ox = 1000.0, oy = 1000.0, r = 0.0;
        if (j < nobs) { ox = ld(obs + j * 7); oy = ld(obs + j * 7 + 1); r = ld(obs + j * 7 + 2); }
        const double d = r + p.radius;
        w[L.OB + j * 3] = ox; w[L.OB + j * 3 + 1] = oy; w[L.OB + j * 3 + 2] = beta * d * d;
      }
    }
    // cold start: u_k = u_prev (mpc_cbf.py:368-369)
    SCB_LANE_UNROLL
    for (int t = lane; t < n; t += LANES) {
      // ... moved strictly inside the input box the way IPOPT does before its first iteration (bound_push =
      // bound_frac = 1e-2): Quad2D's u_prev = 0 start violates f_min <= u, and a start ON a bound has no interior.
      const int i = t % NU;
      double lb = 0.0, ub = 0.0, u0 = 0.0;
#pragma unroll
      for (int m = 0; m < NU; ++m) if (m == i) { u0 = uprev[m]; if (m < NUB && m < 4) { lb = p.u_lb[m]; ub = p.u_ub[m]; } }
      const double pl = fmin(1e-2 * fmax(1.0, fabs(lb)), 1e-2 * (ub - lb));
      const double pu = fmin(1e-2 * fmax(1.0, fabs(ub)), 1e-2 * (ub - lb));
      // (push_start false: the statement probes of the tests; the unbounded omegas start where they are)
      w[L.Z + t] = (push_start && i < NUB) ? fmax(lb + pl, fmin(u0, ub - pu)) : u0;
    }
    SCB_LANE_UNROLL
    for (int i = lane; i < NX; i += LANES) { w[L.X + i] = ld(x0 + i); w[L.XT + i] = w[L.X + i]; }
    sync();
    double Jcur = 0.0;
    if (lane == 0) Jcur = rollout(w + L.Z, w + L.X);
    Jcur = G::bcast(Jcur, 0);
    sync();
    points_and_cbf(w + L.Z, w + L.X, w + L.C);
    return Jcur;
  }

  SCB_HD void solve(int nobs, const double* x0, const double* goal_in, int ngoal, const double* up, const double* obs,
                    double* U, int32_t* status, double* pred_x, double* pred_u, int32_t* iters, double* kkt,
                    uint64_t* active = nullptr) {
    double Jcur = init(nobs, x0, goal_in, ngoal, up, obs);

    // objective scaling as IPOPT's default gradient-based scaling: max |dJ/dz| at the start <= 100
    {
      stage_derivatives();
      double* gam = w + L.GAM;
      SCB_LANE_UNROLL
      for (int t = lane; t < (H + 1) * NY; t += LANES) {
        const int k = t / NY, i = t - k * NY;
        gam[t] = (i < NX) ? 2.0 * Qs[i] * (w[L.X + k * NX + i] - goal[i]) : 0.0;
      }
      sync();
      adjoint(gam, w + L.RD);
      rate_gradient(w + L.Z, w + L.RG);
      double g_abs = 0.0;
      SCB_LANE_UNROLL
      for (int t = lane; t < n; t += LANES) g_abs = fmax(g_abs, fabs(w[L.RD + t] + w[L.RG + t]));
      g_abs = gmax(g_abs);
      const double sf = (g_abs > 100.0) ? 100.0 / g_abs : 1.0;
#pragma unroll
      for (int i = 0; i < NX; ++i) Qs[i] *= sf;
#pragma unroll
      for (int i = 0; i < NU; ++i) { Rs[i] *= sf; Ra[i] *= sf; }
      Jcur *= sf;
      sync();
    }
    double mu_bar = SCB_MPC_MU0, nu_pen = 10.0;
    const double tol = p.mpc_tol;
    // slacks are not independent iterates: s_i = max(g_i(z), mu/nu) is re-derived from the constraint values
    // every iteration, so satisfied rows carry no primal residual however non-linear they are, and only rows
    // below the floor act as (linearly penalised) violations.  The line search runs on the matching
    // penalty-barrier merit  psi(z) = J(z) + sum_i rho(g_i(z)),  rho(g) = -mu log g  (g >= mu/nu), linear below.
    auto reset_slacks = [&](double floor_) {
      // s -> S, 1/s -> DS (the only division per constraint per iteration; everything else multiplies)
      SCB_LANE_UNROLL
      for (int t = lane; t < H * M; t += LANES) w[L.DS + t] = 1.0 / fmax(w[L.C + t], floor_);
      SCB_LANE_UNROLL
      for (int q = lane; q < L.NS; q += LANES) {
        const SimpleCon c = decode_simple<Mod>(p, H, q);
        const double sv = fmax(simple_value(c, w + L.Z, w + L.X), floor_);
        w[L.SS + q] = sv; w[L.SDS + q] = 1.0 / sv;
      }
      sync();
    };
    reset_slacks(mu_bar / nu_pen);
    SCB_LANE_UNROLL
    for (int t = lane; t < H * M; t += LANES) w[L.L + t] = mu_bar * w[L.DS + t];
    SCB_LANE_UNROLL
    for (int q = lane; q < L.NS; q += LANES) w[L.SL + q] = mu_bar * w[L.SDS + q];
    sync();

    int it = 0, st = SCB_MAXITER, it_best = 0, tiny_steps = 0, crawl = 0;
    double err = kInf, err_best = kInf;
    const int max_iter = p.mpc_max_iter > 0 ? p.mpc_max_iter : 150;
    SCB_PH_INIT;
    SCB_LOOP
    for (; it < max_iter; ++it) {
      stage_derivatives();
      SCB_PH(0);
      stage_sums(mu_bar, false);
      SCB_PH(1);
      stage_gradients(false, mu_bar);
      adjoint(w + L.GAM, w + L.RD);                 // costates + d/dz of (J_stage - lam' g)
      rate_gradient(w + L.Z, w + L.RG);
      SCB_PH(2);
      // residuals (e_p = constraint violation, e_c = complementarity)
      double e_d = 0.0, e_p = 0.0, e_c = 0.0, e_cm = 0.0, lam_max = 0.0;
      SCB_LANE_UNROLL
      for (int t = lane; t < n; t += LANES) e_d = fmax(e_d, fabs(w[L.RD + t] + w[L.RG + t]));
      SCB_LANE_UNROLL
      for (int t = lane; t < H * M; t += LANES) {
        const double g = w[L.C + t], lam = w[L.L + t], sg = fmax(g, 0.0);
        e_p = fmax(e_p, -g); e_c = fmax(e_c, sg * lam); e_cm = fmax(e_cm, fabs(sg * lam - mu_bar));
        lam_max = fmax(lam_max, lam);
      }
      SCB_LANE_UNROLL
      for (int q = lane; q < L.NS; q += LANES) {
        const SimpleCon c = decode_simple<Mod>(p, H, q);
        const double g = simple_value(c, w + L.Z, w + L.X), lam = w[L.SL + q], sg = fmax(g, 0.0);
        e_p = fmax(e_p, -g); e_c = fmax(e_c, sg * lam); e_cm = fmax(e_cm, fabs(sg * lam - mu_bar));
        lam_max = fmax(lam_max, lam);
      }
      e_d = gmax(e_d); e_p = gmax(e_p); e_c = gmax(e_c); e_cm = gmax(e_cm);
      lam_max = gmax(lam_max);
      err = fmax(e_d, fmax(e_p, e_c));
      if (!(err == err) || !(lam_max < 1e200)) { st = SCB_NUMERICAL; break; }
      if (err <= tol) { st = SCB_OPTIMAL; break; }
      if (err < 0.5 * err_best) { err_best = err; it_best = it; }
      else if (it - it_best > SCB_MPC_NOPROGRESS) break;   // error not halved for that many iterations: give up
      // monotone barrier update (Fiacco-McCormick with IPOPT's kappa_mu = 0.2, theta_mu = 1.5, kappa_eps = 10)
      while (fmax(e_d, fmax(e_p, e_cm)) <= SCB_MPC_KAPPA_EPS * mu_bar && mu_bar > tol / 10.0) {
        mu_bar = fmax(tol / 10.0, fmin(SCB_MPC_KAPPA_MU * mu_bar, mu_bar * sqrt(mu_bar)));
        e_cm = 0.0;
        SCB_LANE_UNROLL
        for (int t = lane; t < H * M; t += LANES) e_cm = fmax(e_cm, fabs(fmax(w[L.C + t], 0.0) * w[L.L + t] - mu_bar));
        SCB_LANE_UNROLL
        for (int q = lane; q < L.NS; q += LANES) {
          const SimpleCon c = decode_simple<Mod>(p, H, q);
          e_cm = fmax(e_cm, fabs(fmax(simple_value(c, w + L.Z, w + L.X), 0.0) * w[L.SL + q] - mu_bar));
        }
        e_cm = gmax(e_cm);
      }
      nu_pen = fmax(nu_pen, 1.1 * lam_max);
      if (nu_pen > 1e12) { st = SCB_INFEASIBLE; break; }
      const double floor_s = mu_bar / nu_pen;
      reset_slacks(floor_s);
      // keep every multiplier within kappa_Sigma = 1e10 of mu/s (IPOPT's safeguard), using the fresh 1/s
      SCB_LANE_UNROLL
      for (int t = lane; t < H * M; t += LANES) {
        const double c0 = mu_bar * w[L.DS + t];
        w[L.L + t] = fmin(fmax(w[L.L + t], 1e-10 * c0), 1e10 * c0);
      }
      SCB_LANE_UNROLL
      for (int q = lane; q < L.NS; q += LANES) {
        const double c0 = mu_bar * w[L.SDS + q];
        w[L.SL + q] = fmin(fmax(w[L.SL + q], 1e-10 * c0), 1e10 * c0);
      }
      sync();
      SCB_PH(3);
      // Newton system: exact Lagrangian Hessian first; if the reduced matrix is not positive definite,
      // fall back to the Gauss-Newton stage Hessians (PSD by construction) before any diagonal shift
      stage_sums(mu_bar, true, floor_s);
      SCB_PH(9);
      gauss_newton = false;
      stage_hessians();
      SCB_PH(4);
      stage_gradients(true, mu_bar);
      SCB_PH(5);
      double delta = 0.0;
      bool pd = riccati_backward(0.0);
      SCB_PH(6);
      if (!pd) {
        gauss_newton = true;
        stage_hessians();
        gauss_newton = false;
        int tries = 0;
        pd = riccati_backward(0.0);
        while (!pd && tries < 12) { delta = (delta == 0.0) ? 1e-8 : delta * 100.0; pd = riccati_backward(delta); ++tries; }
        if (delta == 0.0) delta = -1.0;             // marks "Gauss-Newton step" in traces
      }
      if (!pd) { st = SCB_NUMERICAL; break; }
      SCB_PH(7);
      riccati_forward();
      SCB_PH(8);
      // linearised constraint change dg (stored in DS), multiplier direction, fraction to the boundary,
      // and the directional derivative of the merit
      const double tau = fmax(0.99, 1.0 - mu_bar);
      double ap = 1.0, ad = 1.0, dpsi = 0.0;
      SCB_LANE_UNROLL
      for (int t = lane; t < H * M; t += LANES) {
        const int k = t / M, j = t - k * M;
        const double* ob = w + L.OB + j * 3;
        const double* dy = w + L.DY + k * NY;
        double dg = 0.0;
#pragma unroll
        for (int i = 0; i < NY; ++i) {
          double gi;
          if constexpr (Mod::GENERAL) gi = w[L.GR + t * NY + i];
          else gi = w[L.JE + k * NY + i] - ob[0] * w[L.JX + k * NY + i] - ob[1] * w[L.JY + k * NY + i];
          dg = fma(gi, dy[i], dg);
        }
        const double g = w[L.C + t], s = fmax(g, floor_s), lam = w[L.L + t];
        const double ds = dg + (g - s), dl = -((s * lam - mu_bar) + lam * ds) * w[L.DS + t];
        w[L.DL + t] = dl;
        if (ds < 0.0) ap = fmin(ap, -tau * s / ds);
        if (dl < 0.0) ad = fmin(ad, -tau * lam / dl);
        dpsi += (g >= floor_s) ? -mu_bar * dg / g : -nu_pen * dg;
      }
      SCB_LANE_UNROLL
      for (int q = lane; q < L.NS; q += LANES) {
        const SimpleCon c = decode_simple<Mod>(p, H, q);
        const double dg = c.sgn * w[L.DY + c.k * NY + c.var];
        const double g = simple_value(c, w + L.Z, w + L.X), s = w[L.SS + q], lam = w[L.SL + q];
        const double ds = dg + (g - s), dl = -((s * lam - mu_bar) + lam * ds) * w[L.SDS + q];
        w[L.SDL + q] = dl;
        if (ds < 0.0) ap = fmin(ap, -tau * s / ds);
        if (dl < 0.0) ad = fmin(ad, -tau * lam / dl);
        dpsi += (g >= floor_s) ? -mu_bar * dg / g : -nu_pen * dg;
      }
      ap = gmin(ap); ad = gmin(ad);
      dpsi = gsum(dpsi);
      SCB_PH(10);
      // directional derivative of the cost: grad J . dz = sum_k grad l_k . dy_k + rate_grad . dz
      double dJ = 0.0;
      SCB_LANE_UNROLL
      for (int t = lane; t < (H + 1) * NX; t += LANES) {
        const int k = t / NX, i = t - k * NX;
        dJ = fma(2.0 * Qs[i] * (w[L.X + t] - goal[i]), w[L.DY + k * NY + i], dJ);
      }
      SCB_LANE_UNROLL
      for (int t = lane; t < n; t += LANES) dJ = fma(w[L.RG + t], w[L.DZ + t], dJ);
      dJ = gsum(dJ);
      dpsi += dJ;
      const double rho_lin0 = -mu_bar * log_call(floor_s) + nu_pen * floor_s;     // rho(g) = rho_lin0 - nu g  below the floor
      auto merit_terms = [&](const double* cb, const double* zz, const double* xx) {
        double acc = 0.0;
        SCB_LANE_UNROLL
        for (int t = lane; t < H * M; t += LANES) {
          const double g = cb[t];
          acc += (g >= floor_s) ? -mu_bar * log_call(g) : rho_lin0 - nu_pen * g;
        }
        SCB_LANE_UNROLL
        for (int q = lane; q < L.NS; q += LANES) {
          const SimpleCon c = decode_simple<Mod>(p, H, q);
          const double g = simple_value(c, zz, xx);
          acc += (g >= floor_s) ? -mu_bar * log_call(g) : rho_lin0 - nu_pen * g;
        }
        return gsum(acc);
      };
      const double psi0 = merit_terms(w + L.C, w + L.Z, w + L.X) + Jcur;
      SCB_PH(11);
      // backtracking
      double alpha = ap, Jt = Jcur;
      int bt = 0;
      SCB_LOOP
      for (; bt < 20; ++bt) {
        sync();                                     // every lane is done reading the previous trial point (merit terms)
        SCB_LANE_UNROLL
        for (int t = lane; t < n; t += LANES) w[L.ZT + t] = fma(alpha, w[L.DZ + t], w[L.Z + t]);
        sync();
        if (lane == 0) Jt = rollout(w + L.ZT, w + L.XT);
        Jt = G::bcast(Jt, 0);
        sync();
        points_and_cbf(w + L.ZT, w + L.XT, w + L.CT);
        const double psi = merit_terms(w + L.CT, w + L.ZT, w + L.XT) + Jt;
        if (psi <= psi0 + 1e-4 * alpha * fmin(dpsi, 0.0) + 1e-13 * fabs(psi0)) break;
        alpha *= 0.5;
      }
#if defined(SCB_MPC_TRACE) && !defined(__CUDA_ARCH__)
      printf("[mpc] it=%3d J=%.6f ed=%.2e ep=%.2e ec=%.2e mu=%.1e ap=%.2e ad=%.2e alpha=%.2e bt=%d delta=%.1e dpsi=%.2e nu=%.2e\n", it, Jcur, e_d, e_p, e_c,
             mu_bar, ap, ad, alpha, bt, delta, dpsi, nu_pen);
#endif
      SCB_PH(12);
      // stalled line search: the step was cut by >= 2^-12 (or to nothing) several iterations in a row.  This is what a
      // kink of the model's own step does (KinematicBicycle2D clips v inside the barrier, kinematic_bicycle2D.py:116-121):
      // the Newton direction is a descent direction of a smooth model the merit does not follow, every further
      // iteration pays ~20 trial rollouts and moves by 1e-6.  Give up instead of repeating that 50 times.  (Only at
      // feasible iterates: far from feasibility a few heavily damped steps in a row are normal and recover.)
      // Local infeasibility, detected early: an agent that STARTS inside a barrier set (a CBF row violated at x_0 whatever
      // the inputs) has every step cut to ~1e-5 by the fraction-to-the-boundary rule and crawls for 50-70 iterations with
      // the violation unchanged before the stagnation exit fires.  Such agents are 2 % of the BASELINE scenes but, being
      // the longest-running ones, they set the duration of a small batch (config 5 on 8 GPUs: 2731 agents per launch).
      // SCB_MPC_CRAWL_ITERS consecutive iterations with a violated row and a primal step below 1e-4 end the solve with
      // the status it would have reached anyway (SCB_INFEASIBLE).
      crawl = (e_p > 1e-6 && alpha < SCB_MPC_CRAWL_ALPHA) ? crawl + 1 : 0;
      if (crawl >= SCB_MPC_CRAWL_ITERS) { st = SCB_INFEASIBLE; break; }
      tiny_steps = (alpha < 1e-10 || (bt >= SCB_MPC_STALL_BT && e_p <= 1e-9)) ? tiny_steps + 1 : 0;
      if (tiny_steps >= SCB_MPC_STALL_ITERS) { st = (e_p > 1e-6) ? SCB_INFEASIBLE : SCB_MAXITER; break; }
      // accept: z, x, g; multipliers move with their own step and are kept within kappa_Sigma of mu/g
      SCB_LANE_UNROLL
      for (int t = lane; t < n; t += LANES) w[L.Z + t] = w[L.ZT + t];
      SCB_LANE_UNROLL
      for (int t = lane; t < (H + 1) * NX; t += LANES) w[L.X + t] = w[L.XT + t];
      sync();
      SCB_LANE_UNROLL
      for (int t = lane; t < H * M; t += LANES) {
        w[L.C + t] = w[L.CT + t];
        w[L.L + t] = fma(ad, w[L.DL + t], w[L.L + t]);
      }
      SCB_LANE_UNROLL
      for (int q = lane; q < L.NS; q += LANES) {
        w[L.SL + q] = fma(ad, w[L.SDL + q], w[L.SL + q]);
      }
      Jcur = Jt;
      sync();
      SCB_PH(13);
    }
    // active set at exit: row i is active <=> its multiplier dominates its value (lam_i > g_i; at a converged point
    // lam_i g_i ~ mu <= 1e-9, so either g_i < 3e-5 < lam_i or the reverse).  Bit k*M + j = CBF row of (stage k, obstacle
    // slot j); bit H*M + q = simple bound q in decode_simple's order.
    if (active) {
      const int total = H * M + L.NS, words = (total + 63) >> 6;
      SCB_LANE_UNROLL
      for (int wd = lane; wd < words; wd += LANES) {
        uint64_t bits = 0ull;
        SCB_LOOP
        for (int b = 0; b < 64; ++b) {
          const int t = wd * 64 + b;
          if (t >= total) break;
          double g, lam;
          if (t < H * M) { g = w[L.C + t]; lam = w[L.L + t]; }
          else {
            const SimpleCon c = decode_simple<Mod>(p, H, t - H * M);
            g = simple_value(c, w + L.Z, w + L.X); lam = w[L.SL + (t - H * M)];
          }
          if (lam > fmax(g, 0.0)) bits |= 1ull << b;
        }
        active[wd] = bits;
      }
    }
    // outputs
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double v = w[L.Z + i];
        if (!(v == v)) v = uprev[i];
        U[i] = (i < NUB && i < 4) ? fmin(fmax(v, p.u_lb[i]), p.u_ub[i]) : v;
      }
      if (st == SCB_MAXITER && err > 1e-4) {
        // distinguish "did not converge" from "locally infeasible" by the primal residual
        double e_p = 0.0;
        SCB_LOOP
        for (int t = 0; t < H * M; ++t) e_p = fmax(e_p, -w[L.C + t]);
        if (e_p > 1e-6) st = SCB_INFEASIBLE;
      }
      *status = st;
      if (iters) *iters = it;
      if (kkt) *kkt = err;
      if (pred_x) for (int t = 0; t < (H + 1) * NX; ++t) pred_x[t] = w[L.X + t];
      if (pred_u) for (int t = 0; t < n; ++t) pred_u[t] = w[L.Z + t];
    }
    sync();
  }
#undef w
};

// entry point shared by the kernel and the host-sim
template <int MODEL, int LANES>
SCB_HD void mpc_agent(const scb_params& p, int H, int M, int nobs, const double* x0, const double* goal,
                      const double* uprev, const double* obs, double* workspace, double* U, int32_t* status,
                      double* pred_x, double* pred_u, int32_t* iters, double* kkt, uint64_t* active = nullptr) {
  using Mod = MpcModel<MODEL>;
  const MpcLayout L = mpc_layout<Mod, LANES == 1>(H, M);
  MpcSolver<MODEL, LANES> s(p, L, workspace);
  if (nobs < 0) nobs = 0;
  if (nobs > M) nobs = M;
  s.solve(nobs, x0, goal, Mod::NGOAL, uprev, obs, U, status, pred_x, pred_u, iters, kkt, active);
}

// The same with the layout supplied by the caller.  The kernel passes its __grid_constant__ copy: the 47 workspace offsets
// are then read from the constant bank (LDC / uniform registers) next to the access that needs them, instead of from a
// struct on the thread's local-memory stack (53 M local loads per cfg3 launch when the layout was built per agent).
template <int MODEL, int LANES>
SCB_HD void mpc_agent(const scb_params& p, const MpcLayout& L, int nobs, const double* x0, const double* goal,
                      const double* uprev, const double* obs, double* workspace, double* U, int32_t* status,
                      double* pred_x, double* pred_u, int32_t* iters, double* kkt, uint64_t* active = nullptr) {
  using Mod = MpcModel<MODEL>;
  MpcSolver<MODEL, LANES> s(p, L, workspace);
  if (nobs < 0) nobs = 0;
  if (nobs > L.M) nobs = L.M;
  s.solve(nobs, x0, goal, Mod::NGOAL, uprev, obs, U, status, pred_x, pred_u, iters, kkt, active);
}

}  // namespace scb
