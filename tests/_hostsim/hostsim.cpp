// hostsim.cpp -- CPU build of the kernels' per-agent bodies (LANES = 1).  TEST AID ONLY:
// lets the `-m "not gpu"` suite check the CUDA source's math against the oracle on a box
// with no GPU.  It is built by tests/_hostsim/build.py into tests/_hostsim/build/, is never
// imported by safe_control_b200, and is not a fallback: the product fails loudly without
// libscb.so + a CUDA device.
#include "../../safe_control_b200/csrc/scb_qp.cuh"
#include "../../safe_control_b200/csrc/scb_track.cuh"
#include "../../safe_control_b200/csrc/scb_backup.cuh"
#include "../../safe_control_b200/csrc/scb_shield.cuh"
#include <vector>
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
    odcbf_agent<MODEL, NW, 1, 128>(p, M, nobs ? nobs[i] : M, X + (size_t)i * ModelCT<MODEL>::NX, Uref + (size_t)i * 2,
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
        ROWCASE(SCB_DOUBLE_INTEGRATOR_2D)
        ROWCASE(SCB_QUAD_2D)
        ROWCASE(SCB_KINEMATIC_BICYCLE_2D_DPCBF)
        ROWCASE(SCB_UNICYCLE_2D)
        case SCB_MANIPULATOR_2D: {
          ManipArm arm; manip_prep(X + (size_t)i * 3, arm);
          manip_row<false>(*p, arm, OBS + (size_t)i * stride, no, r, a, bb);
        } break;
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
  if (M + 6 > 128) return SCB_ERR_TOO_LARGE;
  switch (p->model) {
    case SCB_SINGLE_INTEGRATOR_2D: run_cbfqp<SCB_SINGLE_INTEGRATOR_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_DYNAMIC_UNICYCLE_2D: run_cbfqp<SCB_DYNAMIC_UNICYCLE_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_KINEMATIC_BICYCLE_2D: run_cbfqp<SCB_KINEMATIC_BICYCLE_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_KINEMATIC_BICYCLE_2D_C3BF: run_cbfqp<SCB_KINEMATIC_BICYCLE_2D_C3BF, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_DOUBLE_INTEGRATOR_2D: run_cbfqp<SCB_DOUBLE_INTEGRATOR_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_QUAD_2D: run_cbfqp<SCB_QUAD_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_KINEMATIC_BICYCLE_2D_DPCBF: run_cbfqp<SCB_KINEMATIC_BICYCLE_2D_DPCBF, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_UNICYCLE_2D: run_cbfqp<SCB_UNICYCLE_2D, 128>(*p, N, M, X, Uref, OBS, stride, nobs, U, status, active); break;
    case SCB_MANIPULATOR_2D: {
      const int words = scb_active_words(M, 3);
      for (int i = 0; i < N; ++i)
        manipqp_agent<1, 128, false>(*p, M, nobs ? nobs[i] : M, X + (size_t)i * 3, Uref + (size_t)i * 3, OBS + (size_t)i * stride,
                                     U + (size_t)i * 3, status + i, active ? active + (size_t)i * words : nullptr, words);
    } break;
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
    case SCB_QUAD_2D: run_od<SCB_QUAD_2D, 2>(*p, N, M, X, Uref, OBS, stride, nobs, U, omega, sel, status, active); break;
    default: return SCB_ERR_UNSUPPORTED;
  }
  return 0;
}

#ifdef SCB_HOSTSIM_MPC
int hostsim_mpccbf_solve(const scb_params* p, int N, int M, int H, const double* X, const double* goal,
                         const double* u_prev, const double* OBS, long stride, const int32_t* nobs, double* U,
                         int32_t* status, double* pred_x, double* pred_u, int32_t* iters, double* kkt, uint64_t* active) {
  const int nx = p->nx, nu = p->nu + (p->od_mpc ? 2 : 0);
  const int aw = active ? scb_mpc_active_words(p, M, H) : 0;
  for (int i = 0; i < N; ++i) {
    const int no = nobs ? nobs[i] : M;
    double* px = pred_x ? pred_x + (size_t)i * (H + 1) * nx : nullptr;
    double* pu = pred_u ? pred_u + (size_t)i * H * nu : nullptr;
    // agents with a superellipsoid row (flag >= 0.5) take the general-row variant of their model, like mpc_kernel
    int model = p->model;
    if (model == SCB_SINGLE_INTEGRATOR_2D || model == SCB_DYNAMIC_UNICYCLE_2D || model == SCB_DOUBLE_INTEGRATOR_2D)
      for (int j = 0; j < no && j < M; ++j)
        if (OBS[(size_t)i * stride + j * 7 + 6] >= 0.5) { model = kMpcSeBase + p->model; break; }
    if (p->od_mpc) model = kMpcOdBase + p->model;             // optimal decay: nu = model inputs + 2 (omega1, omega2)
    switch (model) {
#define MPCCASE(MODEL)                                                                                          \
  case MODEL: {                                                                                                 \
    using Mod = MpcModel<MODEL>;                                                                                \
    const MpcLayout L = mpc_layout<Mod, true>(H, M);                                      \
    double* ws = new double[L.total];                                                                           \
    for (int t = 0; t < L.total; ++t) ws[t] = 0.0;                                                              \
    mpc_agent<MODEL, 1>(*p, H, M, no, X + (size_t)i * nx, goal + (size_t)i * Mod::NGOAL, u_prev + (size_t)i * nu,        \
                        OBS + (size_t)i * stride, ws, U + (size_t)i * nu, status + i, px, pu,                   \
                        iters ? iters + i : nullptr, kkt ? kkt + i : nullptr,                                   \
                        active ? active + (size_t)i * aw : nullptr);                                            \
    delete[] ws;                                                                                                \
  } break;
      MPCCASE(SCB_SINGLE_INTEGRATOR_2D)
      MPCCASE(SCB_DYNAMIC_UNICYCLE_2D)
      MPCCASE(SCB_KINEMATIC_BICYCLE_2D)
      MPCCASE(SCB_QUAD_3D)
      MPCCASE(SCB_DOUBLE_INTEGRATOR_2D)
      MPCCASE(SCB_QUAD_2D)
      MPCCASE(SCB_UNICYCLE_2D)
      MPCCASE(SCB_KINEMATIC_BICYCLE_2D_C3BF)
      MPCCASE(SCB_KINEMATIC_BICYCLE_2D_DPCBF)
      MPCCASE(SCB_VTOL_2D)
      MPCCASE(kMpcOdBase + SCB_DYNAMIC_UNICYCLE_2D)
      MPCCASE(kMpcOdBase + SCB_KINEMATIC_BICYCLE_2D)
      MPCCASE(kMpcOdBase + SCB_QUAD_2D)
      MPCCASE(kMpcOdBase + SCB_VTOL_2D)
      MPCCASE(kMpcSeBase + SCB_SINGLE_INTEGRATOR_2D)
      MPCCASE(kMpcSeBase + SCB_DYNAMIC_UNICYCLE_2D)
      MPCCASE(kMpcSeBase + SCB_DOUBLE_INTEGRATOR_2D)
      default: return SCB_ERR_UNSUPPORTED;
    }
  }
  return 0;
}

// The kernel's own problem statement at a probe point: horizon 1, cold start u_0 = u (passed as u_prev), so
// init() leaves x_1 = Euler(x, u) in the rollout, the stage cost of x_0 (+ terminal cost of x_1) in J and the CBF
// value of every obstacle slot in w[L.C].  Compared with what the reference's mpc_cbf.py hands to do-mpc.
int hostsim_mpc_statement(const scb_params* p, int M, int nobs, const double* x, const double* u, const double* goal,
                          const double* obs, double* x_next, double* stage_cost, double* cbf) {
  int model = p->model;
  if (model == SCB_SINGLE_INTEGRATOR_2D || model == SCB_DYNAMIC_UNICYCLE_2D || model == SCB_DOUBLE_INTEGRATOR_2D)
    for (int j = 0; j < nobs && j < M; ++j)
      if (obs[j * 7 + 6] >= 0.5) { model = kMpcSeBase + p->model; break; }
  if (p->od_mpc) model = kMpcOdBase + p->model;               // u then holds [u, omega1, omega2]
  switch (model) {
#define STCASE(MODEL)                                                                                           \
  case MODEL: {                                                                                                 \
    using Mod = MpcModel<MODEL>;                                                                                \
    const MpcLayout L = mpc_layout<Mod, true>(1, M);               \
    std::vector<double> ws(L.total, 0.0);                                                                       \
    MpcSolver<MODEL, 1> s(*p, L, ws.data());                                                                    \
    const double J = s.init(nobs, x, goal, Mod::NGOAL, u, obs, false);                                               \
    double term = 0.0;                                                                                          \
    for (int i = 0; i < Mod::NX; ++i) {                                                                         \
      x_next[i] = ws[L.X + Mod::NX + i];                                                                        \
      const double g = i < Mod::NGOAL ? goal[i] : 0.0;                                                          \
      term += p->Q[i] * (x_next[i] - g) * (x_next[i] - g);                                                      \
    }                                                                                                           \
    *stage_cost = J - term;                                                                                     \
    for (int j = 0; j < M; ++j) cbf[j] = ws[L.C + j];                                                           \
  } break;
    STCASE(SCB_SINGLE_INTEGRATOR_2D)
    STCASE(SCB_DYNAMIC_UNICYCLE_2D)
    STCASE(SCB_KINEMATIC_BICYCLE_2D)
    STCASE(SCB_QUAD_3D)
    STCASE(SCB_DOUBLE_INTEGRATOR_2D)
    STCASE(SCB_QUAD_2D)
    STCASE(SCB_UNICYCLE_2D)
    STCASE(SCB_KINEMATIC_BICYCLE_2D_C3BF)
    STCASE(SCB_KINEMATIC_BICYCLE_2D_DPCBF)
    STCASE(SCB_VTOL_2D)
    STCASE(kMpcOdBase + SCB_DYNAMIC_UNICYCLE_2D)
    STCASE(kMpcOdBase + SCB_KINEMATIC_BICYCLE_2D)
    STCASE(kMpcOdBase + SCB_QUAD_2D)
    STCASE(kMpcOdBase + SCB_VTOL_2D)
    STCASE(kMpcSeBase + SCB_SINGLE_INTEGRATOR_2D)
    STCASE(kMpcSeBase + SCB_DYNAMIC_UNICYCLE_2D)
    STCASE(kMpcSeBase + SCB_DOUBLE_INTEGRATOR_2D)
    default: return SCB_ERR_UNSUPPORTED;
  }
  return 0;
}
#endif


// ---- closed loop (scb_track.cuh), LANES = 1 ------------------------------------------------------
int hostsim_select_obstacles(const scb_params* p, int N, int K, int M, const double* X, const double* yaw,
                             const double* SCENE, long sstride, double* OBS, int32_t* nobs, int32_t* idx) {
  std::vector<double> keys(K > 0 ? K : 1);
  for (int a = 0; a < N; ++a) {
    switch (p->model) {
#define SELCASE(MODEL)                                                                                          \
  case MODEL: {                                                                                                 \
    using ML = ModelLoop<MODEL>;                                                                                \
    const double* x = X + (size_t)a * ML::NX;                                                                   \
    const double psi = yaw ? yaw[a] : ML::yaw_of(x, 0.0);                                                       \
    nobs[a] = select_agent<1>(K, M, SCENE + (size_t)a * sstride, x[0], x[1], psi, ML::half_angle(), keys.data(), \
                              OBS + (size_t)a * M * 7, idx ? idx + (size_t)a * M : nullptr);                     \
  } break;
      SELCASE(SCB_SINGLE_INTEGRATOR_2D)
      SELCASE(SCB_DYNAMIC_UNICYCLE_2D)
      SELCASE(SCB_KINEMATIC_BICYCLE_2D)
      SELCASE(SCB_KINEMATIC_BICYCLE_2D_C3BF)
      SELCASE(SCB_QUAD_3D)
      SELCASE(SCB_KINEMATIC_BICYCLE_2D_DPCBF)
      SELCASE(SCB_DOUBLE_INTEGRATOR_2D)
      SELCASE(SCB_UNICYCLE_2D)
      SELCASE(SCB_QUAD_2D)
      default: return SCB_ERR_UNSUPPORTED;
    }
  }
  return 0;
}

size_t hostsim_track_sizeof(void) { return sizeof(scb_track); }

// same sequence as control_step_impl in scb_api.cu: pre -> [dyn obs] -> solve -> post
int hostsim_control_step(const scb_params* p, const scb_track* t) {
  std::vector<double> keys(t->K > 0 ? t->K : 1);
  for (long a = 0; a < t->N; ++a) {
    switch (p->model) {
#define PRECASE(MODEL) case MODEL: track_pre_agent<MODEL, 1>(*p, *t, a, keys.data(), t->SCENE); break;
      PRECASE(SCB_SINGLE_INTEGRATOR_2D)
      PRECASE(SCB_DYNAMIC_UNICYCLE_2D)
      PRECASE(SCB_KINEMATIC_BICYCLE_2D)
      PRECASE(SCB_KINEMATIC_BICYCLE_2D_C3BF)
      PRECASE(SCB_QUAD_3D)
      PRECASE(SCB_KINEMATIC_BICYCLE_2D_DPCBF)
      PRECASE(SCB_DOUBLE_INTEGRATOR_2D)
      PRECASE(SCB_UNICYCLE_2D)
      PRECASE(SCB_QUAD_2D)
      default: return SCB_ERR_UNSUPPORTED;
    }
  }
  if (t->dynamic_obs)
    for (int j = 0; j < t->K; ++j) {
      t->SCENE[j * 7 + 0] += t->SCENE[j * 7 + 3] * p->dt;
      t->SCENE[j * 7 + 1] += t->SCENE[j * 7 + 4] * p->dt;
    }
  const long stride = 7L * t->M;
  int rc;
  if (t->controller == SCB_CTRL_CBF_QP) {
    const int words = scb_active_words(t->M, p->nu);
    rc = 0;
    for (int a = 0; a < t->N && rc == 0; ++a) {
      if (t->done[a]) continue;                                    // frozen agents keep their last outputs
      rc = hostsim_cbfqp_solve(p, 1, t->M, t->X + (size_t)a * p->nx, t->Uref + (size_t)a * p->nu, t->OBS + (size_t)a * stride,
                               stride, t->nobs + a, t->U + (size_t)a * p->nu, t->status + a,
                               t->active ? t->active + (size_t)a * words : nullptr);
    }
  } else if (t->controller == SCB_CTRL_OPTIMAL_DECAY) {
    rc = 0;
    for (int a = 0; a < t->N && rc == 0; ++a) {
      if (t->done[a]) continue;
      rc = hostsim_odcbf_solve(p, 1, t->M, t->X + (size_t)a * p->nx, t->Uref + (size_t)a * 2, t->OBS + (size_t)a * stride, stride,
                               t->nobs + a, t->U + (size_t)a * 2, nullptr, nullptr, t->status + a,
                               t->active ? t->active + a : nullptr);
    }
  } else {
#ifdef SCB_HOSTSIM_MPC
    const int nx = p->nx, nu = p->nu, ng = (p->model == SCB_QUAD_3D) ? 3 : 2;
    rc = 0;
    for (int a = 0; a < t->N && rc == 0; ++a) {
      if (t->done[a]) continue;
      if (t->track_flag[a] <= 0) {                                 // mpc_cbf.py:379-381
        for (int i = 0; i < nu; ++i) t->U[(size_t)a * nu + i] = t->Uref[(size_t)a * nu + i];
        t->status[a] = SCB_OPTIMAL;
        continue;
      }
      rc = hostsim_mpccbf_solve(p, 1, t->M, t->H, t->X + (size_t)a * nx, t->goal + (size_t)a * ng,
                                t->u_prev + (size_t)a * nu, t->OBS + (size_t)a * stride, stride, t->nobs + a,
                                t->U + (size_t)a * nu, t->status + a, nullptr, nullptr,
                                t->mpc_iters ? t->mpc_iters + a : nullptr, nullptr, nullptr);
    }
#else
    rc = SCB_ERR_UNSUPPORTED;
#endif
  }
  if (rc != 0) return rc;
  for (long a = 0; a < t->N; ++a) {
    switch (p->model) {
#define POSTCASE(MODEL) case MODEL: track_post_agent<MODEL, 1>(*p, *t, a, t->SCENE); break;
      POSTCASE(SCB_SINGLE_INTEGRATOR_2D)
      POSTCASE(SCB_DYNAMIC_UNICYCLE_2D)
      POSTCASE(SCB_KINEMATIC_BICYCLE_2D)
      POSTCASE(SCB_KINEMATIC_BICYCLE_2D_C3BF)
      POSTCASE(SCB_QUAD_3D)
      POSTCASE(SCB_KINEMATIC_BICYCLE_2D_DPCBF)
      POSTCASE(SCB_DOUBLE_INTEGRATOR_2D)
      POSTCASE(SCB_UNICYCLE_2D)
      POSTCASE(SCB_QUAD_2D)
      default: return SCB_ERR_UNSUPPORTED;
    }
  }
  return 0;
}

// Backup-CBF QP bodies (scb_backup.cuh), one agent after the other; same outputs as scb_backupcbf_solve
int hostsim_backupcbf_solve(const scb_backup_params* p, int N, int K, const double* X, const double* Uref, const double* MOV,
                            long mov_stride, double* U, int32_t* status, int32_t* intervene, double* h_min, double* phi,
                            double* rows, uint64_t* active) {
  const int nb = p->n_backup, words = scb_backup_active_words(nb);
  if (nb < 1 || nb + 4 > 256) return SCB_ERR_TOO_LARGE;
  std::vector<double> scr(kBkScratch), rw((size_t)3 * nb);
  for (long a = 0; a < N; ++a) {
    BackupOut o;
    backup_agent<1, 256>(*p, X + a * 4, Uref + a * 2, K > 0 ? MOV + a * mov_stride : nullptr, K, scr.data(), rw.data(),
                         phi ? phi + a * nb * 4 : nullptr, o);
    U[a * 2] = o.u0; U[a * 2 + 1] = o.u1; status[a] = o.status;
    if (intervene) intervene[a] = o.intervene;
    if (h_min) h_min[a] = o.h_min;
    if (rows) for (int k = 0; k < 3 * nb; ++k) rows[a * nb * 3 + k] = rw[k];
    if (active)
      for (int w = 0; w < words; ++w) {
        uint64_t bits = 0ull;
        if (o.w0 >= 0 && o.lam0 > 0.0 && (o.w0 >> 6) == w) bits |= 1ull << (o.w0 & 63);
        if (o.w1 >= 0 && o.lam1 > 0.0 && (o.w1 >> 6) == w) bits |= 1ull << (o.w1 & 63);
        active[a * words + w] = bits;
      }
  }
  return 0;
}

// gatekeeper / MPS bodies (scb_shield.cuh), one agent after the other on HOST arrays laid out like scb_shield_step's
int hostsim_shield_step(const scb_shield_params* sp, const scb_shield_state* st, int N, int K, const double* X, const double* NOMX,
                        const double* NOMU, const int32_t* nom_len, const double* MOV, long mov_stride, const double* STAT,
                        double* U, int32_t* using_backup) {
  const int T = sp->nom_cap, Nb = sp->scene.n_backup;
  for (long a = 0; a < N; ++a) {
    ShieldIO io;
    io.x = X + a * 4; io.nomx = NOMX + a * (long)(T + 1) * 4; io.nomu = NOMU + a * (long)T * 2;
    int nl = nom_len ? nom_len[a] : T + 1;
    io.nom_len = nl < 0 ? 0 : (nl > T + 1 ? T + 1 : nl);
    io.mov = (K > 0 && MOV) ? MOV + a * mov_stride : nullptr; io.K = (K > 0 && MOV) ? K : 0;
    io.stat = STAT ? STAT + a * 5 : nullptr;
    const int cb = st->cbuf[a] & 1;
    const long lu = (long)(T + Nb) * 2, lx = (long)(T + Nb + 1) * 4;
    io.cu = st->CU + (a * 2 + cb) * lu; io.cu_spare = st->CU + (a * 2 + (cb ^ 1)) * lu;
    io.cx = st->CX ? st->CX + (a * 2 + cb) * lx : nullptr; io.cx_spare = st->CX ? st->CX + (a * 2 + (cb ^ 1)) * lx : nullptr;
    int ub = 0;
    bool flip = false;
    shield_agent<1, 0>(*sp, io, st->clen[a], st->cidx[a], st->nsteps[a], st->next_event[a], U + a * 2, ub, flip);
    if (flip) st->cbuf[a] = cb ^ 1;
    if (using_backup) using_backup[a] = ub;
  }
  return 0;
}

}  // extern "C"
