// scb_gi.cuh -- exact dual active-set (Goldfarb-Idnani) solver for the tiny strictly
// convex QPs of the CBF-QP / optimal-decay paths:
//
//     min  1/2 sum_i hd_i (x_i - x0_i)^2     s.t.   a_r . x + b_r >= 0 ,  r = 0..mrows-1
//
// with NV <= 4 variables and a diagonal Hessian -- exactly the problems the reference
// builds with cvxpy and hands to GUROBI (position_control/cbf_qp.py:47-106,190;
// optimal_decay_cbf_qp.py:56-130,156).
//
// Why an active-set method and not an interior-point loop here: with 2-4 variables the
// optimum has at most NV active rows, the dual method reaches it in (typically) 1-3
// pivots of O(mrows) work, it terminates at the EXACT vertex (so the active-constraint
// indices are well defined and reproducible bit for bit), and it certifies infeasibility
// exactly instead of by a residual heuristic.  An IPM would need ~10-15 dependent Newton
// steps for a less precise answer.  (The MPC path, where the problem is a real NLP, does
// use a primal-dual interior-point loop: scb_mpc.cuh.)
//
// Work split inside a lane group: row r lives in registers of lane r % LANES (slot
// r / LANES); the most-violated-row search is a lane-local scan + one xor-shuffle
// argmin; the working set (<= NV rows) and all NV x NV algebra are replicated in every
// lane, so there is no shared memory and no barrier.
#pragma once

#include "scb_core.cuh"

namespace scb {

template <int NV>
struct QpOut {
  double x[NV];
  double lam[NV];   // multipliers of the working set
  int widx[NV];     // row indices of the working set
  int wk;           // working-set size
  int status;       // scb_status
  int iters;        // constraint additions + drops
};

// in-place Cholesky solve of the leading k x k block of S (k <= NV), rhs -> solution.
// returns false on a non-positive pivot.
template <int NV>
SCB_HD bool chol_solve_small(double (&S)[NV][NV], double (&r)[NV], int k) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (j < k) {
      double d = S[j][j];
#pragma unroll
      for (int t = 0; t < NV; ++t)
        if (t < j) d -= S[j][t] * S[j][t];
      if (!(d > 0.0)) return false;
      const double dj = sqrt(d);
      S[j][j] = dj;
      const double inv = 1.0 / dj;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (i > j && i < k) {
          double v = S[i][j];
#pragma unroll
          for (int t = 0; t < NV; ++t)
            if (t < j) v -= S[i][t] * S[j][t];
          S[i][j] = v * inv;
        }
      }
    }
  }
  // forward: L y = r
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < k) {
      double v = r[i];
#pragma unroll
      for (int t = 0; t < NV; ++t)
        if (t < i) v -= S[i][t] * r[t];
      r[i] = v / S[i][i];
    }
  }
  // backward: L' x = y
#pragma unroll
  for (int ii = NV - 1; ii >= 0; --ii) {
    if (ii < k) {
      double v = r[ii];
#pragma unroll
      for (int t = 0; t < NV; ++t)
        if (t > ii && t < k) v -= S[t][ii] * r[t];
      r[ii] = v / S[ii][ii];
    }
  }
  return true;
}

// ra[j][i], rb[j]: slot j of THIS lane holds row r = j * LANES + lane (rows >= mrows ignored).
template <int NV, int LANES, int RPL>
SCB_HD void gi_solve(const double (&hd)[NV], const double (&x0)[NV],
                     const double (&ra)[RPL][NV], const double (&rb)[RPL],
                     int mrows, int max_iter, QpOut<NV>& out) {
  using G = Grp<LANES>;
  const int lane = G::lane();
  constexpr int kNone = 0x7fffffff;

  double hinv[NV], x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { hinv[i] = 1.0 / hd[i]; x[i] = x0[i]; }

  double rn[RPL];   // 1 / ||a_r||  (0 for an all-zero normal)
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) n2 += ra[j][i] * ra[j][i];
    rn[j] = (n2 > 0.0) ? 1.0 / sqrt(n2) : 0.0;
  }

  double Wa[NV][NV], lam[NV];
  int Wi[NV];
  int k = 0;
#pragma unroll
  for (int a = 0; a < NV; ++a) { Wi[a] = -1; lam[a] = 0.0; }
  int status = SCB_OPTIMAL, it = 0;

  while (true) {
    // ---- most violated row (normalised slack), excluding the working set ----
    double bestv = 0.0;
    int bi = kNone;
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      const int r = j * LANES + lane;
      bool inW = false;
#pragma unroll
      for (int a = 0; a < NV; ++a) inW = inW || (a < k && Wi[a] == r);
      if (r < mrows && !inW) {
        double s = rb[j];
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fma(ra[j][i], x[i], s);
        const double sn = (rn[j] > 0.0) ? s * rn[j] : s;
        const double tol = 1e-12 * (1.0 + fabs(rb[j]) * rn[j]);
        if (sn < -tol && sn < bestv) { bestv = sn; bi = r; }
      }
    }
    G::argmin(bestv, bi);
#if defined(SCB_GI_TRACE) && !defined(__CUDA_ARCH__)
    printf("[gi] it=%d k=%d W=(%d,%d) x=(%.6g,%.6g) pick=%d sn=%.3e\n", it, k, Wi[0], NV > 1 ? Wi[1] : -1, x[0], x[1], bi, bestv);
#endif
    if (bi == kNone) break;                       // primal feasible -> optimal
    if (++it > max_iter) { status = SCB_MAXITER; break; }

    // ---- fetch row bi into every lane ----
    double ap[NV], bp = 0.0;
    {
      const int src = bi % LANES, slot = bi / LANES;
      double mine[NV], mb = 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) mine[i] = 0.0;
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        if (j == slot) {
#pragma unroll
          for (int i = 0; i < NV; ++i) mine[i] = ra[j][i];
          mb = rb[j];
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) ap[i] = G::bcast(mine[i], src);
      bp = G::bcast(mb, src);
    }
    double sp = bp;
#pragma unroll
    for (int i = 0; i < NV; ++i) sp = fma(ap[i], x[i], sp);
    double lam_p = 0.0;

    // ---- pivot until row bi is satisfied with all multipliers >= 0 ----
    bool done = false;
    while (!done) {
      double d[NV], z[NV], r[NV];
      double dap = 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) { d[i] = hinv[i] * ap[i]; dap = fma(d[i], ap[i], dap); }
#pragma unroll
      for (int a = 0; a < NV; ++a) r[a] = 0.0;
      if (k > 0) {
        double S[NV][NV];
#pragma unroll
        for (int a = 0; a < NV; ++a) {
#pragma unroll
          for (int b = 0; b < NV; ++b) {
            double v = 0.0;
            if (a < k && b <= a) {
#pragma unroll
              for (int i = 0; i < NV; ++i) v = fma(Wa[a][i] * hinv[i], Wa[b][i], v);
            }
            S[a][b] = v;
          }
          double v = 0.0;
          if (a < k) {
#pragma unroll
            for (int i = 0; i < NV; ++i) v = fma(Wa[a][i], d[i], v);
          }
          r[a] = v;
        }
        if (!chol_solve_small<NV>(S, r, k)) { status = SCB_NUMERICAL; done = true; break; }
      }
      double zap = 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < NV; ++a)
          if (a < k) v = fma(Wa[a][i], r[a], v);
        z[i] = d[i] - hinv[i] * v;
        zap = fma(z[i], ap[i], zap);
      }
      // z is the H^-1-projection of a_p onto the null space of the working rows: identically 0
      // once k == NV; otherwise zap/dap = sin^2(angle(a_p, span W)), and below ~1e-9 the
      // computed value is rounding noise of the k x k solve (cond ~ 1/sin^2) -> treat as dependent.
      const bool zzero = (k >= NV) || !(zap > 1e-9 * dap);
      double t1 = kInf;
      int l = -1;
#pragma unroll
      for (int a = 0; a < NV; ++a) {
        if (a < k && r[a] > 0.0) {
          const double t = lam[a] / r[a];
          if (t < t1) { t1 = t; l = a; }
        }
      }
      const double t2 = zzero ? kInf : (-sp / zap);
#if defined(SCB_GI_TRACE) && !defined(__CUDA_ARCH__)
      printf("[gi]   k=%d sp=%.6g zap=%.3e dap=%.3e zzero=%d t1=%.6g l=%d t2=%.6g r=(%.3e,%.3e) lam=(%.3e,%.3e)\n", k, sp, zap, dap, (int)zzero, t1, l, t2, r[0], NV > 1 ? r[1] : 0.0, lam[0], NV > 1 ? lam[1] : 0.0);
#endif
      if (l < 0 && zzero) { status = SCB_INFEASIBLE; done = true; break; }
      if (t2 <= t1) {
        // full step: row bi becomes active
#pragma unroll
        for (int i = 0; i < NV; ++i) x[i] = fma(t2, z[i], x[i]);
#pragma unroll
        for (int a = 0; a < NV; ++a)
          if (a < k) lam[a] = fmax(lam[a] - t2 * r[a], 0.0);
        lam_p += t2;
#pragma unroll
        for (int a = 0; a < NV; ++a) {
          if (a == k) {
#pragma unroll
            for (int i = 0; i < NV; ++i) Wa[a][i] = ap[i];
            Wi[a] = bi;
            lam[a] = lam_p;
          }
        }
        ++k;
        done = true;
      } else {
        // partial step: multiplier of working row l hits zero -> drop it, retry
        if (!zzero) {
#pragma unroll
          for (int i = 0; i < NV; ++i) x[i] = fma(t1, z[i], x[i]);
          sp = fma(t1, zap, sp);
        }
        lam_p += t1;
#pragma unroll
        for (int a = 0; a < NV; ++a)
          if (a < k) lam[a] = fmax(lam[a] - t1 * r[a], 0.0);
#pragma unroll
        for (int a = 0; a < NV - 1; ++a) {
          if (a >= l && a < k - 1) {
#pragma unroll
            for (int i = 0; i < NV; ++i) Wa[a][i] = Wa[a + 1][i];
            Wi[a] = Wi[a + 1];
            lam[a] = lam[a + 1];
          }
        }
        --k;
#pragma unroll
        for (int a = 0; a < NV; ++a)
          if (a == k) { Wi[a] = -1; lam[a] = 0.0; }
        if (++it > max_iter) { status = SCB_MAXITER; done = true; }
      }
    }
    if (status != SCB_OPTIMAL) break;
  }

#pragma unroll
  for (int i = 0; i < NV; ++i) out.x[i] = x[i];
#pragma unroll
  for (int a = 0; a < NV; ++a) { out.lam[a] = (a < k) ? lam[a] : 0.0; out.widx[a] = (a < k) ? Wi[a] : -1; }
  out.wk = k;
  out.status = status;
  out.iters = it;
}

// ----------------------------------------------------------------------------------------
// NV = 2 specialisation (the CBF-QP of every 2-input model): the working set has 0, 1 or 2
// rows, so each pivot is a closed form -- no generic k x k factorisation, no guarded
// unrolled loops.  Rows must be PRE-NORMALISED (||a_r|| = 1, or a_r = 0 for a constant
// row), which makes the slack a signed distance and removes a multiply + a register per
// row; the multipliers returned are those of the normalised rows (same sign pattern, which
// is all the active mask needs).  Hessian = hd * I (hd = 2 for ||u - u_ref||^2).
struct Qp2Out {
  double x0, x1;
  double lam0, lam1;
  int w0, w1;        // working-set row indices (-1 = empty)
  int status, iters;
};

template <int LANES, int RPL>
SCB_HD void gi_solve2(double hd, double x0, double x1, const double (&r0)[RPL], const double (&r1)[RPL],
                      const double (&rb)[RPL], int mrows, int max_iter, Qp2Out& out) {
  using G = Grp<LANES>;
  const int lane = G::lane();
  constexpr int kNone = 0x7fffffff;
  const double hinv = 1.0 / hd;
  // working set
  double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0, l0 = 0.0, l1 = 0.0;
  int w0 = -1, w1 = -1, k = 0;
  int status = SCB_OPTIMAL, it = 0;

  while (true) {
    // most violated row (signed distance), working rows excluded
    double bestv = 0.0;
    int bi = kNone;
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      const int r = j * LANES + lane;
      const double sdist = fma(r1[j], x1, fma(r0[j], x0, rb[j]));
      const double tol = -1e-12 * (1.0 + fabs(rb[j]));
      if (r < mrows && r != w0 && r != w1 && sdist < tol && sdist < bestv) { bestv = sdist; bi = r; }
    }
    G::argmin(bestv, bi);
    if (bi == kNone) break;
    if (++it > max_iter) { status = SCB_MAXITER; break; }

    // fetch the row into every lane
    double p0 = 0.0, p1 = 0.0, pb = 0.0;
    {
      const int slot = bi / LANES;
#pragma unroll
      for (int j = 0; j < RPL; ++j)
        if (j == slot) { p0 = r0[j]; p1 = r1[j]; pb = rb[j]; }
      const int src = bi % LANES;
      p0 = G::bcast(p0, src); p1 = G::bcast(p1, src); pb = G::bcast(pb, src);
    }
    double sp = fma(p1, x1, fma(p0, x0, pb));
    double lp = 0.0;
    const double pp = p0 * p0 + p1 * p1;            // 1 for a normalised row, 0 for a constant row

    // All rows are unit-normal (pp == 1) or constant (pp == 0), so the Gram entries of the working set are 1 and
    // every pivot needs ONE division: step lengths are compared by cross-multiplication before dividing.
    bool done = false;
    while (!done) {
      if (k == 0) {
        if (!(pp > 0.0)) { status = SCB_INFEASIBLE; break; }       // constant row b < 0
        // z = hinv p, t = -sp / (hinv pp)  ->  x += t z = -(sp / pp) p
        const double step = -sp / pp;
        x0 = fma(step, p0, x0); x1 = fma(step, p1, x1);
        a00 = p0; a01 = p1; w0 = bi; l0 = lp + step * hd; k = 1;      // (lp: dual steps already taken for this row while rows were dropped)
        done = true;
      } else if (k == 1) {
        // r = a0.p (|a0| = 1);  z = hinv (p - r a0);  z.p = hinv (pp - r^2)
        const double r = a00 * p0 + a01 * p1;
        const double g = pp - r * r;                                // sin^2 of the angle between the rows
        const bool zzero = !(g > 1e-9 * pp);
        const double zap = hinv * g;
        const bool has1 = r > 0.0;
        if (!has1 && zzero) { status = SCB_INFEASIBLE; break; }
        // full step iff t2 = -sp/zap <= t1 = l0/r   <=>   -sp r <= l0 zap   (r, zap > 0)
        const bool full = !zzero && (!has1 || (-sp * r <= l0 * zap));
        if (full) {
          const double t2 = -sp / zap;
          const double z0 = hinv * fma(-r, a00, p0), z1 = hinv * fma(-r, a01, p1);
          x0 = fma(t2, z0, x0); x1 = fma(t2, z1, x1);
          l0 = fmax(l0 - t2 * r, 0.0); lp += t2;
          a10 = p0; a11 = p1; w1 = bi; l1 = lp; k = 2;
          done = true;
        } else {
          const double t1 = l0 / r;
          if (!zzero) {
            const double z0 = hinv * fma(-r, a00, p0), z1 = hinv * fma(-r, a01, p1);
            x0 = fma(t1, z0, x0); x1 = fma(t1, z1, x1); sp = fma(t1, zap, sp);
          }
          lp += t1;
          w0 = -1; l0 = 0.0; k = 0;
          if (++it > max_iter) { status = SCB_MAXITER; break; }
        }
      } else {
        // two unit rows span the plane (z = 0): solve [[1,c],[c,1]] (ra, rb) = (q0, q1),  c = a0.a1
        const double c = a00 * a10 + a01 * a11;
        const double q0 = a00 * p0 + a01 * p1, q1 = a10 * p0 + a11 * p1;
        const double det = 1.0 - c * c;
        if (!(det > 0.0)) { status = SCB_NUMERICAL; break; }
        const double na = q0 - c * q1, nb = q1 - c * q0;            // ra = na / det, rb = nb / det (same signs)
        const bool pa = na > 0.0, pb = nb > 0.0;
        if (!pa && !pb) { status = SCB_INFEASIBLE; break; }
        // ta = l0 det / na, tb = l1 det / nb: pick the smaller by cross-multiplication, divide once
        const bool drop_a = pa && (!pb || (l0 * nb <= l1 * na));
        const double tq = drop_a ? l0 / na : l1 / nb;               // = t / det
        lp += tq * det;
        l0 = fmax(l0 - tq * na, 0.0); l1 = fmax(l1 - tq * nb, 0.0);
        if (drop_a) { a00 = a10; a01 = a11; w0 = w1; l0 = l1; }     // drop row 0: row 1 moves down
        w1 = -1; l1 = 0.0; k = 1;
        if (++it > max_iter) { status = SCB_MAXITER; break; }
      }
    }
    if (status != SCB_OPTIMAL) break;
  }
  out.x0 = x0; out.x1 = x1;
  out.lam0 = (k > 0) ? l0 : 0.0; out.lam1 = (k > 1) ? l1 : 0.0;
  out.w0 = (k > 0) ? w0 : -1; out.w1 = (k > 1) ? w1 : -1;
  out.status = status; out.iters = it;
}

}  // namespace scb
