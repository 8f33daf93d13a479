// scb_jet.cuh -- forward-mode jets over the stage variables y = (x, u).  Used by the MPC path to obtain
// EXACT first and second derivatives of the Euler dynamics and of the discrete-time barriers from the same
// few lines that define them (no hand-derived Hessians to get wrong), one derivative ENTRY per lane.
#pragma once

#include "scb_core.cuh"

namespace scb {

template <int N>
SCB_HD int hidx(int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }   // packed upper triangle, row-major, i <= j

// ---- entry jets ---------------------------------------------------------------------------------------
// Every operation of forward-mode differentiation (sum, product, chain rule) is ENTRY-WISE in the derivative index:
// gradient entry i of a result needs only the values and gradient entry i of the operands, Hessian entry (i, j) only
// values, gradient entries i, j and Hessian entry (i, j).  So instead of one lane carrying a whole second-order jet
// (1 + 6 + 21 = 28 doubles at 6 stage variables, ~100 flops per product, 255 registers + spills in the first version
// of the kernel) a lane evaluates the same expression for ONE entry with a 2- or 4-double jet: the MPC kernel spreads
// the (stage, entry) pairs of its derivative passes over the lanes of the agent's group this way.
struct JetG { double v, g; };              // value, d/dy_i
struct JetH { double v, gi, gj, h; };      // value, d/dy_i, d/dy_j, d2/(dy_i dy_j)

SCB_HD void jvar_entry(JetG& r, double val, int idx, int i) { r.v = val; r.g = (idx == i) ? 1.0 : 0.0; }
SCB_HD void jvar_entry(JetH& r, double val, int idx, int i, int j) {
  r.v = val; r.gi = (idx == i) ? 1.0 : 0.0; r.gj = (idx == j) ? 1.0 : 0.0; r.h = 0.0;
}
SCB_HD void jconst(JetG& r, double c) { r.v = c; r.g = 0.0; }
SCB_HD void jconst(JetH& r, double c) { r.v = c; r.gi = 0.0; r.gj = 0.0; r.h = 0.0; }
SCB_HD void jaxpy(JetG& r, const JetG& a, double s, const JetG& b) { r.v = a.v + s * b.v; r.g = a.g + s * b.g; }
SCB_HD void jaxpy(JetH& r, const JetH& a, double s, const JetH& b) {
  r.v = a.v + s * b.v; r.gi = a.gi + s * b.gi; r.gj = a.gj + s * b.gj; r.h = a.h + s * b.h;
}
SCB_HD void jscale(JetG& r, const JetG& a, double s) { r.v = s * a.v; r.g = s * a.g; }
SCB_HD void jscale(JetH& r, const JetH& a, double s) { r.v = s * a.v; r.gi = s * a.gi; r.gj = s * a.gj; r.h = s * a.h; }
SCB_HD void jmul(JetG& r, const JetG& a, const JetG& b) {
  const double av = a.v, bv = b.v;
  r.g = av * b.g + bv * a.g; r.v = av * bv;
}
SCB_HD void jmul(JetH& r, const JetH& a, const JetH& b) {
  const double av = a.v, bv = b.v, agi = a.gi, agj = a.gj, bgi = b.gi, bgj = b.gj;
  r.h = av * b.h + bv * a.h + agi * bgj + agj * bgi;
  r.gi = av * bgi + bv * agi; r.gj = av * bgj + bv * agj; r.v = av * bv;
}
SCB_HD void jchain(JetG& r, const JetG& a, double f0, double f1, double) { r.g = f1 * a.g; r.v = f0; }
SCB_HD void jchain(JetH& r, const JetH& a, double f0, double f1, double f2) {
  const double agi = a.gi, agj = a.gj;
  r.h = f1 * a.h + f2 * agi * agj; r.gi = f1 * agi; r.gj = f1 * agj; r.v = f0;
}
SCB_HD void jclip(JetG& r, const JetG& a, double lo, double hi) {
  if (a.v > hi) jconst(r, hi); else if (a.v < lo) jconst(r, lo); else r = a;
}
SCB_HD void jclip(JetH& r, const JetH& a, double lo, double hi) {
  if (a.v > hi) jconst(r, hi); else if (a.v < lo) jconst(r, lo); else r = a;
}
// plain values speak the same vocabulary, so one stage map serves values (line search) and every jet flavour
SCB_HD void jsincos(double& s, double& c, double a) { sincos_pair(a, s, c); }
SCB_HD void jmul(double& r, double a, double b) { r = a * b; }
SCB_HD void jaxpy(double& r, double a, double s, double b) { r = a + s * b; }
SCB_HD void jscale(double& r, double a, double s) { r = s * a; }
SCB_HD void jclip(double& r, double a, double lo, double hi) { r = a > hi ? hi : (a < lo ? lo : a); }
SCB_HD double jval(const JetG& a) { return a.v; }
SCB_HD double jval(const JetH& a) { return a.v; }
SCB_HD double jval(double a) { return a; }
SCB_HD void jchain(double& r, const double&, double f0, double, double) { r = f0; }
SCB_HD void jconst(double& r, double c) { r = c; }
SCB_HD void jvar_entry(double& r, double val, int, int) { r = val; }
SCB_HD void jvar_entry(double& r, double val, int, int, int) { r = val; }

// bivariate chain rule r = f(a, b): value f0, first partials fa, fb, second partials faa, fab, fbb
SCB_HD void jchain2(double& r, const double&, const double&, double f0, double, double, double, double, double) { r = f0; }
SCB_HD void jchain2(JetG& r, const JetG& a, const JetG& b, double f0, double fa, double fb, double, double, double) {
  r.g = fa * a.g + fb * b.g; r.v = f0;
}
SCB_HD void jchain2(JetH& r, const JetH& a, const JetH& b, double f0, double fa, double fb, double faa, double fab, double fbb) {
  const double agi = a.gi, agj = a.gj, bgi = b.gi, bgj = b.gj;
  r.h = fa * a.h + fb * b.h + faa * agi * agj + fab * (agi * bgj + agj * bgi) + fbb * bgi * bgj;
  r.gi = fa * agi + fb * bgi; r.gj = fa * agj + fb * bgj; r.v = f0;
}
// atan2(y, x) with a precomputed value `val` (= atan2(y.v, x.v)): partials  x / r2,  -y / r2;  second partials
// -2xy / r2^2 (yy),  (y^2 - x^2) / r2^2 (yx),  2xy / r2^2 (xx)
template <class T>
SCB_HD void jatan2(T& r, const T& y, const T& x, double val) {
  const double yv = jval(y), xv = jval(x), r2 = xv * xv + yv * yv, i2 = 1.0 / r2, i4 = i2 * i2;
  jchain2(r, y, x, val, xv * i2, -yv * i2, -2.0 * xv * yv * i4, (yv * yv - xv * xv) * i4, 2.0 * xv * yv * i4);
}

// sqrt(max(a, 0)) and 1/a through the chain rule, for every jet flavour (double, JetG, JetH)
template <class T>
SCB_HD void jsqrt0(T& r, const T& a) {
  const double v = jval(a);
  if (v > 1e-300) { const double s = sqrt(v); jchain(r, a, s, 0.5 / s, -0.25 / (s * v)); }
  else jconst(r, 0.0);                                  // the flat side of the kink (CasADi: fmax(., 0))
}
template <class T>
SCB_HD void jrecip(T& r, const T& a) {
  const double inv = 1.0 / jval(a);
  jchain(r, a, inv, -inv * inv, 2.0 * inv * inv * inv);
}
// (|a| k)^e, e >= 2
template <class T>
SCB_HD void jpowabs(T& r, const T& a, double e, double k) {
  const double x = jval(a), ax = fabs(x) * k, sg = (x < 0.0) ? -1.0 : 1.0;
  const double p2 = pow(ax, e - 2.0);                    // ax^(e-2) (1 at e = 2)
  jchain(r, a, p2 * ax * ax, e * p2 * ax * k * sg, e * (e - 1.0) * p2 * k * k);
}
template <class T>
SCB_HD void jadd(T& r, const T& a, const T& b) { jaxpy(r, a, 1.0, b); }
template <class T>
SCB_HD void jaddc(T& r, const T& a, double c) { T k; jconst(k, c); jaxpy(r, a, 1.0, k); }

// sin/cos providers for the stage maps: compute, compute + remember, or replay the remembered values (the entry-jet
// passes evaluate one stage many times at the same point; the transcendental is paid once per stage and iterate)
struct TrigCompute {
  // a scalar the stage map needs at the iterate (exp, atan2 ...): compute / compute + remember / replay, like sin and cos
  template <class F>
  SCB_HD double memo(F f) { return f(); }
  template <class T>
  SCB_HD void operator()(T& s, T& c, const T& a) {
    double sv, cv;
    sincos_call(jval(a), &sv, &cv);
    T a0 = a;                                  // (s or c may alias a)
    jchain(s, a0, sv, cv, -sv);
    jchain(c, a0, cv, -sv, -cv);
  }
};
struct TrigStore {
  double* buf;
  int n;
  SCB_HD explicit TrigStore(double* b) : buf(b), n(0) {}
  template <class F>
  SCB_HD double memo(F f) { const double v = f(); buf[2 * n] = v; buf[2 * n + 1] = 0.0; ++n; return v; }
  template <class T>
  SCB_HD void operator()(T& s, T& c, const T& a) {
    double sv, cv;
    sincos_call(jval(a), &sv, &cv);
    buf[2 * n] = sv; buf[2 * n + 1] = cv; ++n;
    T a0 = a;
    jchain(s, a0, sv, cv, -sv);
    jchain(c, a0, cv, -sv, -cv);
  }
};
struct TrigLoad {
  const double* buf;
  int n;
  SCB_HD explicit TrigLoad(const double* b) : buf(b), n(0) {}
  template <class F>
  SCB_HD double memo(F) { const double v = buf[2 * n]; ++n; return v; }
  template <class T>
  SCB_HD void operator()(T& s, T& c, const T& a) {
    const double sv = buf[2 * n], cv = buf[2 * n + 1];
    ++n;
    T a0 = a;
    jchain(s, a0, sv, cv, -sv);
    jchain(c, a0, cv, -sv, -cv);
  }
};

}  // namespace scb
