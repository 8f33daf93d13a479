// scb_jet.cuh -- second-order forward-mode jets (value, gradient, packed Hessian) over N
// stage variables y = (x, u).  Used by the MPC path to obtain EXACT first and second
// derivatives of the Euler dynamics and of the discrete-time barrier points from the same
// few lines that define them (no hand-derived Hessians to get wrong); N = 6 for the
// 4-state / 2-input models, so a jet is 28 doubles and a product ~150 flops.
#pragma once

#include "scb_core.cuh"

namespace scb {

template <int N>
struct Jet {
  static constexpr int NH = N * (N + 1) / 2;
  double v;
  double g[N];
  double h[NH];   // upper triangle, row-major: (i,j), i <= j  ->  i*N - i*(i-1)/2 + (j - i)
};

template <int N>
SCB_HD int hidx(int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }

template <int N>
SCB_HD void jconst(Jet<N>& r, double c) {
  r.v = c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = 0.0;
#pragma unroll
  for (int i = 0; i < Jet<N>::NH; ++i) r.h[i] = 0.0;
}

template <int N>
SCB_HD void jvar(Jet<N>& r, double val, int idx) {
  jconst(r, val);
#pragma unroll
  for (int i = 0; i < N; ++i) if (i == idx) r.g[i] = 1.0;
}

// r = a + s * b
template <int N>
SCB_HD void jaxpy(Jet<N>& r, const Jet<N>& a, double s, const Jet<N>& b) {
  r.v = a.v + s * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] + s * b.g[i];
#pragma unroll
  for (int i = 0; i < Jet<N>::NH; ++i) r.h[i] = a.h[i] + s * b.h[i];
}

template <int N>
SCB_HD void jscale(Jet<N>& r, const Jet<N>& a, double s) {
  r.v = s * a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = s * a.g[i];
#pragma unroll
  for (int i = 0; i < Jet<N>::NH; ++i) r.h[i] = s * a.h[i];
}

template <int N>
SCB_HD void jmul(Jet<N>& r, const Jet<N>& a, const Jet<N>& b) {
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.v * b.g[i] + b.v * a.g[i];
  int t = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = i; j < N; ++j, ++t)
      r.h[t] = a.v * b.h[t] + b.v * a.h[t] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  }
}

// r = f(a) given f(a.v) = f0, f'(a.v) = f1, f''(a.v) = f2
template <int N>
SCB_HD void jchain(Jet<N>& r, const Jet<N>& a, double f0, double f1, double f2) {
  r.v = f0;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = f1 * a.g[i];
  int t = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = i; j < N; ++j, ++t) r.h[t] = f1 * a.h[t] + f2 * a.g[i] * a.g[j];
  }
}

template <int N>
SCB_HD void jsincos(Jet<N>& s, Jet<N>& c, const Jet<N>& a) {
  double sv, cv;
  sincos_pair(a.v, sv, cv);
  jchain(s, a, sv, cv, -sv);
  jchain(c, a, cv, -sv, -cv);
}

// clip(a, lo, hi) as CasADi's fmax(fmin(a, hi), lo): identity inside, constant outside
template <int N>
SCB_HD void jclip(Jet<N>& r, const Jet<N>& a, double lo, double hi) {
  if (a.v > hi) jconst(r, hi);
  else if (a.v < lo) jconst(r, lo);
  else r = a;
}

// ---- first-order jets (value + gradient): same vocabulary, used where no curvature is needed ----------
template <int N>
struct Jet1 {
  double v;
  double g[N];
};

template <int N>
SCB_HD void jvar(Jet1<N>& r, double val, int idx) {
  r.v = val;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = (i == idx) ? 1.0 : 0.0;
}
template <int N>
SCB_HD void jconst(Jet1<N>& r, double c) {
  r.v = c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = 0.0;
}
template <int N>
SCB_HD void jaxpy(Jet1<N>& r, const Jet1<N>& a, double s, const Jet1<N>& b) {
  r.v = a.v + s * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] + s * b.g[i];
}
template <int N>
SCB_HD void jscale(Jet1<N>& r, const Jet1<N>& a, double s) {
  r.v = s * a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = s * a.g[i];
}
template <int N>
SCB_HD void jmul(Jet1<N>& r, const Jet1<N>& a, const Jet1<N>& b) {
  const double av = a.v, bv = b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = av * b.g[i] + bv * a.g[i];
  r.v = av * bv;
}
template <int N>
SCB_HD void jsincos(Jet1<N>& s, Jet1<N>& c, const Jet1<N>& a) {
  double sv, cv;
  sincos_pair(a.v, sv, cv);
#pragma unroll
  for (int i = 0; i < N; ++i) { s.g[i] = cv * a.g[i]; c.g[i] = -sv * a.g[i]; }
  s.v = sv; c.v = cv;
}
template <int N>
SCB_HD void jclip(Jet1<N>& r, const Jet1<N>& a, double lo, double hi) {
  if (a.v > hi) jconst(r, hi);
  else if (a.v < lo) jconst(r, lo);
  else r = a;
}

}  // namespace scb
