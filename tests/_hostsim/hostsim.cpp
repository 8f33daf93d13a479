// hostsim.cpp -- CPU build of the kernels' per-agent bodies (LANES = 1).  TEST AID ONLY:
// lets the `-m "not gpu"` suite check the CUDA source's math against the oracle on a box
// with no GPU.  It is built by tests/_hostsim/build.py into tests/_hostsim/build/, is never
// imported by safe_control_b200, and is not a fallback: the product fails loudly without
// libscb.so + a CUDA device.
#include "../../safe_control_b200/csrc/scb_qp.cuh"
#ifdef SCB_HOSTSIM_MPC
#include "../../safe_control_b200/csrc/scb_mpc.cuh"
#endif

using namespace scb;

template <int MODEL, int RPL>
static void run_cbfqp(const scb_params& p, int N, int M, const double* X, const double* Uref, const double* OBS,
                      long stride, const int32_t* nobs, double* U, int32_t* status, uint64_t* active) {
  const int words = scb_active_words(M, ModelCT<MODEL>::NU);
  for (int i = 0; i < N; ++i)
    cbfqp_agent<MODEL, 1, RPL>(p, M, nobs ? nobs[i] : M, X + (size_t)i * ModelCT<MODEL>::NX,
                               Uref + (size_t)i * ModelCT<MODEL>::NU, OBS + (size_t)i * stride,
                               U + (size_t)i * ModelCT<MODEL>::NU, status + i,
                               active ? active + (size_t)i * words : nullptr, words);
}

template <int MODEL, int NW>
static void run_od(const scb_params& p, int N, int M, const double* X, const double* Uref, const double* OBS,
                   long stride, const int32_t* nobs, double* U, double* omega, int32_t* sel, int32_t* status,
                   uint64_t* active) {
  for (int i = 0; i < N; ++i)
    odcbf_agent<MODEL, NW, 1, 128>(p, M, nobs ? nobs[i] : M, X + (size_t)i * 4, Uref + (size_t)i * 2,
                                   OBS + (size_t)i * stride, U + (size_t)i * 2, omega ? omega + (size_t)i * 2 : nullptr,
                                   sel ? sel + i : nullptr, status + i, active ? active + i : nullptr);
}

extern "C" {

int hostsim_cbfqp_rows(const scb_params* p, int N, int M, const double* X, const double* OBS, long stride,
                       const int32_t* nobs, double* A, double* b) {
  for (int i = 0; i < N; ++i) {
    for (int r = 0; r < M; ++r) {
      double a[4] = {0, 0, 0, 0}, bb = 0;
      AgentCT g;
      const int no = nobs ? (nobs[i] < 0 ? 0 : (nobs[i] > M ? M : nobs[i])) : M;
      switch (p->model) {
#define ROWCASE(MODEL)                                                                         \
  case MODEL:                                                                                  \
    ModelCT<MODEL>::prep(*p, X + (size_t)i * ModelCT<MODEL>::NX, g);                           \
    cbfqp_row<MODEL>(*p, g, OBS + (size_t)i * stride, M, no, r, a, bb);                        \
    break;
        ROWCASE(SCB_SINGLE_INTEGRATOR_2D)
        ROWCASE(SCB_DYNAMIC_UNICYCLE_2D)
        ROWCASE(SCB_KINEMATIC_BICYCLE_2D)
        ROWCASE(SCB_KINEMATIC_BICYCLE_2D_C3BF)
        default: return SCB_ERR_UNSUPPORTED;
      }
      for (int t = 0; t < p->nu; ++t) A[((size_t)i * M + r) * p->nu + t] = a[t];
      b[(size_t)i * M + r] = bb;
    }
  }
  return 0;
}

int hostsim_cbfqp_solve(const scb_params* p, int N, int M, const double* X, const double* Uref, const double* OBS,
                        long stride, const int32_t* nobs, double* U, int32_t* status, uint64_t* active) {
  if (M + 4 > 128) return SCB_ERR_TOO_LARGE;
  switch (p->model) {
    case SCB_SINGLE_INTEGRATOR_2D: run_cbfqp<SCB_SINGLE_INTEGRATOR_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_DYNAMIC_UNICYCLE_2D: run_cbfqp<SCB_DYNAMIC_UNICYCLE_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_KINEMATIC_BICYCLE_2D: run_cbfqp<SCB_KINEMATIC_BICYCLE_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_KINEMATIC_BICYCLE_2D_C3BF: run_cbfqp<SCB_KINEMATIC_BICYCLE_2D_C3BF, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    default: return SCB_ERR_UNSUPPORTED;
  }
  return 0;
}

int hostsim_odcbf_solve(const scb_params* p, int N, int M, const double* X, const double* Uref, const double* OBS,
                        long stride, const int32_t* nobs, double* U, double* omega, int32_t* sel, int32_t* status,
                        uint64_t* active) {
  if (M > 128) return SCB_ERR_TOO_LARGE;
  switch (p->model) {
    case SCB_DYNAMIC_UNICYCLE_2D: run_od<SCB_DYNAMIC_UNICYCLE_2D, 2>(*p, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active); break;
    case SCB_KINEMATIC_BICYCLE_2D: run_od<SCB_KINEMATIC_BICYCLE_2D, 2>(*p, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active); break;
    case SCB_KINEMATIC_BICYCLE_2D_C3BF: run_od<SCB_KINEMATIC_BICYCLE_2D_C3BF, 1>(*p, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active); break;
    default: return SCB_ERR_UNSUPPORTED;
  }
  return 0;
}

#ifdef SCB_HOSTSIM_MPC
int hostsim_mpccbf_solve(const scb_params* p, int N, int M, int H, const double* X, const double* goal,
                         const double* u_prev, const double* OBS, long stride, const int32_t* nobs, double* U,
                         int32_t* status, double* pred_x, double* pred_u, int32_t* iters, double* kkt) {
  const int nx = p->nx, nu = p->nu;
  for (int i = 0; i < N; ++i) {
    const int no = nobs ? nobs[i] : M;
    double* px = pred_x ? pred_x + (size_t)i * (H + 1) * nx : nullptr;
    double* pu = pred_u ? pred_u + (size_t)i * H * nu : nullptr;
    switch (p->model) {
#define MPCCASE(MODEL)                                                                                          \
  case MODEL: {                                                                                                 \
    using Mod = MpcModel<MODEL>;                                                                                \
    const MpcLayout L = mpc_layout<Mod::NX, Mod::NU, Mod::VBOUND, Mod::LINEAR, Mod::AUX>(H, M);                                        \
    double* ws = new double[L.total];                                                                           \
    for (int t = 0; t < L.total; ++t) ws[t] = 0.0;                                                              \
    mpc_agent<MODEL, 1>(*p, H, M, no, X + (size_t)i * nx, goal + (size_t)i * Mod::NGOAL, u_prev + (size_t)i * nu,        \
                        OBS + (size_t)i * stride, ws, U + (size_t)i * nu, status + i, px, pu,                   \
                        iters ? iters + i : nullptr, kkt ? kkt + i : nullptr);                                  \
    delete[] ws;                                                                                                \
  } break;
      MPCCASE(SCB_SINGLE_INTEGRATOR_2D)
      MPCCASE(SCB_DYNAMIC_UNICYCLE_2D)
      MPCCASE(SCB_KINEMATIC_BICYCLE_2D)
      MPCCASE(SCB_QUAD_3D)
      default: return SCB_ERR_UNSUPPORTED;
    }
  }
  return 0;
}
#endif

}  // extern "C"
