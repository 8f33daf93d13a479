/* scb.h -- C ABI of libscb.so: the B200-native batched safety-filter solve.
 *
 * This is the drop-in boundary for the reference's per-step position-controller
 * solve (tkkim-robot/safe_control):
 *
 *   scb_cbfqp_solve*      replaces  CBFQP.solve_control_problem            position_control/cbf_qp.py:108-199
 *   scb_cbfqp_rows        replaces  the row loop of the same method        position_control/cbf_qp.py:122-185
 *   scb_odcbf_solve*      replaces  OptimalDecayCBFQP.solve_control_problem position_control/optimal_decay_cbf_qp.py:132-159
 *   scb_mpccbf_solve*     replaces  MPCCBF.solve_control_problem           position_control/mpc_cbf.py:366-402
 *   scb_params            replaces  the robot_spec / cbf_param dictionaries robots/*.py ctor setdefault cascades,
 *                                                                          cbf_qp.py:12-43, optimal_decay_cbf_qp.py:17-50,
 *                                                                          mpc_cbf.py:15-95
 *
 * for a BATCH of N independent agents (the reference solves one agent per call,
 * tracking.py:611-616).  Plain C types only; all arithmetic is IEEE float64 like
 * the reference's numpy / GUROBI / IPOPT path.
 *
 * Two families of entry points:
 *   *_solve       device pointers owned by the caller (e.g. torch tensors), asynchronous
 *                 on `stream`; nothing is allocated.
 *   *_solve_host  HOST pointers (e.g. numpy arrays): stages through the context's device
 *                 buffers (H2D, launch, D2H) and returns when the outputs are in host
 *                 memory.  This is what a reference-side ctypes binding calls.
 *
 * Array layouts (row-major, float64 unless noted):
 *   X      [N, nx]     agent states                       (robot.X, robots/robot.py:38)
 *   Uref   [N, nu]     nominal inputs                     (control_ref['u_ref'], tracking.py:607-609)
 *   OBS    [N, M, 7]   per-agent obstacle lists, or [M, 7] shared by all agents when
 *                      obs_stride_agent == 0.  Row = [x, y, r, vx, vy, -, flag=0] (circle) or
 *                      [x, y, a, b, e, theta, flag=1] (superellipsoid)  (README.md:133-138).
 *                      obs_stride_agent is in doubles (normally 7*M).
 *   nobs   [N] int32   number of valid rows per agent (NULL = all M); rows >= nobs[i] are
 *                      vacuous `0 >= 0` rows exactly like the reference's zeroed A1/b1
 *                      (cbf_qp.py:110-111).  nobs[i] < 0 means "obs_list is None": the QP
 *                      controller returns u_ref unclipped (cbf_qp.py:113-118).
 *   U      [N, nu]     filtered inputs (out)
 *   status [N] int32   (out) SCB_OPTIMAL / SCB_INFEASIBLE / SCB_MAXITER / SCB_NUMERICAL.
 *                      The reference's `.status == 'optimal'` <=> status == 0 (tracking.py:628).
 *   active [N, W] u64  (out, may be NULL) bitmask of the optimal working set;
 *                      W = scb_active_words(M, nu).  Bit j < M: CBF row of obstacle slot j;
 *                      bit M+2i: u_i at its upper bound; bit M+2i+1: u_i at its lower bound.
 *
 * All functions return 0 on success or a negative scb_error code; they never throw and
 * never touch errno.  Calls with distinct contexts/streams/workspaces are re-entrant; the
 * library owns no device memory and no global mutable state (the MPC kernel's dynamic agent
 * scheduling keeps its work counters in the caller's workspace, see scb_mpccbf_solve_ws).
 */
#ifndef SCB_H_
#define SCB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_VERSION 210

/* model ids (robot_spec['model'], robots/robot.py:65-175) */
enum scb_model {
  SCB_SINGLE_INTEGRATOR_2D = 0,     /* robots/single_integrator2D.py */
  SCB_DYNAMIC_UNICYCLE_2D = 1,      /* robots/dynamic_unicycle2D.py */
  SCB_KINEMATIC_BICYCLE_2D = 2,     /* robots/kinematic_bicycle2D.py */
  SCB_KINEMATIC_BICYCLE_2D_C3BF = 3,/* dynamic_env/kinematic_bicycle2D_c3bf.py (all three controllers) */
  SCB_QUAD_3D = 4,                  /* robots/quad3D.py (MPC only: agent_barrier raises, quad3D.py:269-273) */
  SCB_DOUBLE_INTEGRATOR_2D = 5,     /* robots/double_integrator2D.py        (cbf_qp, mpc_cbf) */
  SCB_QUAD_2D = 6,                  /* robots/quad2D.py                      (all three controllers) */
  SCB_KINEMATIC_BICYCLE_2D_DPCBF = 7,/* dynamic_env/kinematic_bicycle2D_dpcbf.py (cbf_qp, mpc_cbf) */
  SCB_UNICYCLE_2D = 8,              /* robots/unicycle2D.py (cbf_qp, mpc_cbf) */
  SCB_MANIPULATOR_2D = 9,           /* robots/manipulator2D.py (cbf_qp: 3 inputs, 25 link-circle rows per obstacle;
                                       M = CBFQP's num_obs = the ROW budget, cbf_qp.py:131-149) */
  SCB_VTOL_2D = 10,                 /* robots/vtol2D.py (mpc_cbf only: agent_barrier is not implemented, vtol2D.py:458-460;
                                       6 states, 4 inputs, horizon 30, mpc_cbf.py:40-43, 83-87, 222-232) */
  SCB_NUM_MODELS = 11
};

enum scb_status { SCB_OPTIMAL = 0, SCB_INFEASIBLE = 1, SCB_MAXITER = 2, SCB_NUMERICAL = 3 };

enum scb_error {
  SCB_OK = 0,
  SCB_ERR_BAD_ARG = -1,        /* NULL pointer, negative size, unknown model */
  SCB_ERR_UNSUPPORTED = -2,    /* model/controller pair the reference does not support either */
  SCB_ERR_TOO_LARGE = -3,      /* M or horizon beyond the compiled limits (see scb_limits) */
  SCB_ERR_CUDA = -4,           /* launch / memcpy failure; scb_last_cuda_error() has the code */
  SCB_ERR_NO_DEVICE = -5,
  SCB_ERR_ALLOC = -6
};

/* Everything the reference keeps in robot_spec / cbf_param for one (model, controller)
 * group, resolved once on the host.  POD, passed by pointer, copied at launch. */
typedef struct scb_params {
  int32_t model;          /* enum scb_model */
  int32_t cbf_mode;       /* 0 = 'cbf', 1 = 'hard'      robot_spec['cbf_mode'], cbf_qp.py:120 */
  int32_t nx, nu;         /* filled by scb_params_default */
  double dt;              /* tracking.py:39 */
  double radius;          /* robot_spec['radius'], robots/robot.py:49-50 */
  double alpha;           /* rel-degree-1 gain        cbf_qp.py:12-35 / mpc_cbf.py:49-82 / optimal_decay:17-50 */
  double alpha1, alpha2;  /* rel-degree-2 gains */
  double u_lb[4], u_ub[4];/* input box               cbf_qp.py:54-73, mpc_cbf.py:183-221 */
  double v_min, v_max;    /* KB step clip (kinematic_bicycle2D.py:116-121); MPC state bound |x[3]| <= v_max */
  double rear_ax_dist;    /* KB L_r (kinematic_bicycle2D.py:48) */
  double omega1_0, omega2_0, p_sb1, p_sb2;   /* optimal_decay_cbf_qp.py:17-50 */
  double Q[12];           /* MPC state weights (diagonal)  mpc_cbf.py:19-39 */
  double R[4];            /* MPC input-rate weights        mpc_cbf.py:19-39, 180 */
  double mass, Ix, Iy, Iz, arm_L, nu_coef, gravity;     /* quad3D.py:53-69; Quad2D: mass, Iy = inertia, gravity 9.81 (quad2D.py:40-46) */
  int32_t mpc_max_iter;   /* interior-point iteration cap (ours) */
  int32_t mpc_superellipsoid; /* mpc_cbf, SingleIntegrator2D / DynamicUnicycle2D / DoubleIntegrator2D: 1 = OBS may hold
                                 superellipsoid rows (flag 1; *_2D.py agent_barrier_dt if_else): the agents that have one
                                 are solved by a second launch with general rows.  0: such agents get SCB_NUMERICAL */
  double mpc_tol;         /* KKT tolerance (ours; IPOPT default 1e-8) */
  /* VTOL2D (robots/vtol2D.py:57-110; mass, Iy = inertia, gravity 9.81, v_max above): wing / aero coefficients, lift
   * blending (M, alpha_0), rotor gains and lever arms, safety limits of the MPC state bounds (mpc_cbf.py:227-232:
   * |x_dot| <= v_max, z_dot >= -descent_speed_max, |theta| <= pitch_max * 3.14159 / 180, pitch_max in degrees) */
  double S_wing, rho, C_L0, C_Lalpha, blend_M, alpha_0, C_Ldelta_e, C_D0, C_Dalpha, C_Ddelta_e, C_m0, C_malpha, C_mdelta_e,
         chord, k_front, k_rear, k_pusher, ell_f, ell_r, pitch_max, descent_speed_max;
  int32_t od_mpc;         /* 1: optimal-decay MPC-CBF (position_control/optimal_decay_mpc_cbf.py): every stage has two more
                             inputs omega1, omega2; the u_prev / Uref / U / pred_u arrays of scb_mpccbf_solve* then hold
                             nu + 2 columns [u, omega1, omega2].  Set by scb_params_default(.., "optimal_decay_mpc_cbf"). */
  int32_t od_sum_rterms;  /* optimal-decay MPC: 0 = do-mpc's assignment semantics of the two set_rterm calls (:178-185: the omega
                             penalty replaces sum R_i u_i^2), 1 = both terms (the reference's evident intent).  UNPINNED. */
} scb_params;

typedef struct scb_ctx scb_ctx;   /* opaque: device staging buffers + stream for the *_host calls */

/* ---- library / parameter helpers ------------------------------------------------------- */
int         scb_version(void);
size_t      scb_params_sizeof(void);                   /* sizeof(scb_params): lets a binding verify its struct mirror */
long        scb_params_offsetof(const char* field);    /* offsetof(scb_params, field) by name, -1 if unknown */
long        scb_track_offsetof(const char* field);     /* offsetof(scb_track, field) by name, -1 if unknown */
const char* scb_strerror(int err);
int         scb_last_cuda_error(void);                 /* cudaError_t of the last SCB_ERR_CUDA on this thread */
int         scb_device_count(void);
/* Fill `p` with the reference's defaults for (model, controller); controller is one of
 * "cbf_qp", "optimal_decay_cbf_qp", "mpc_cbf", "optimal_decay_mpc_cbf".  Returns SCB_ERR_UNSUPPORTED for pairs the
 * reference has no branch for. */
int         scb_params_default(scb_params* p, int model, const char* controller);
int         scb_model_dims(int model, int* nx, int* nu);
int         scb_active_words(int M, int nu);           /* ceil((M + 2 nu) / 64) */
/* compiled limits: max obstacle slots for the QP kernels, max slots and horizon for MPC */
int         scb_limits(int* max_obs_qp, int* max_obs_mpc, int* max_horizon);

/* measurement helper (bench.py): achieved FP64 FMA throughput of the current device in TFLOP/s (2 flops per DFMA), a
 * chain-parallel kernel timed with CUDA events on `stream`.  The measured denominator of the MPC kernels' roofline. */
int         scb_measure_fp64_peak(double* tflops, void* stream);

/* measurement helper (bench.py: roofline.latency_floor): microseconds per launch, inside a CUDA graph, of a kernel with
 * config 2's geometry (256 CTAs x 128 threads) that does nothing / one / two dependent cold-DRAM round trips per warp. */
int         scb_measure_latency_floor(double* empty_us, double* one_trip_us, double* two_trip_us, void* stream);

/* ---- context for the host-pointer calls ------------------------------------------------- */
int  scb_ctx_create(scb_ctx** out, int device);
void scb_ctx_destroy(scb_ctx* ctx);
/* number of kernel launches issued through this context so far (for benchmarks) */
long scb_ctx_launches(const scb_ctx* ctx);

/* ---- CBF-QP  (position_control/cbf_qp.py) ------------------------------------------------ */
/* constraint rows only: A [N, M, nu], b [N, M]   (A u + b >= 0) */
int scb_cbfqp_rows(const scb_params* p, int N, int M,
                   const double* X, const double* OBS, long obs_stride_agent, const int32_t* nobs,
                   double* A, double* b, void* stream);
int scb_cbfqp_solve(const scb_params* p, int N, int M,
                    const double* X, const double* Uref,
                    const double* OBS, long obs_stride_agent, const int32_t* nobs,
                    double* U, int32_t* status, uint64_t* active, void* stream);
int scb_cbfqp_solve_host(scb_ctx* ctx, const scb_params* p, int N, int M,
                         const double* X, const double* Uref,
                         const double* OBS, long obs_stride_agent, const int32_t* nobs,
                         double* U, int32_t* status, uint64_t* active);

/* ---- optimal-decay CBF-QP  (position_control/optimal_decay_cbf_qp.py) -------------------- */
/* ONE CBF row per agent, built from the nearest valid obstacle (by centre distance) of the
 * agent's list -- i.e. nearest_multi_obs[0] of tracking.py:585-586.  omega [N, 2] (out, may
 * be NULL): omega1 (, omega2; NaN for rel-degree-1 models).  sel [N] int32 (out, may be NULL):
 * the obstacle slot the row was built from (-1 if none). */
int scb_odcbf_solve(const scb_params* p, int N, int M,
                    const double* X, const double* Uref,
                    const double* OBS, long obs_stride_agent, const int32_t* nobs,
                    double* U, double* omega, int32_t* sel, int32_t* status, uint64_t* active, void* stream);
int scb_odcbf_solve_host(scb_ctx* ctx, const scb_params* p, int N, int M,
                         const double* X, const double* Uref,
                         const double* OBS, long obs_stride_agent, const int32_t* nobs,
                         double* U, double* omega, int32_t* sel, int32_t* status, uint64_t* active);

/* ---- MPC-CBF  (position_control/mpc_cbf.py) ---------------------------------------------- */
/* goal [N, 2] (3 for Quad3D: x, y, z), u_prev [N, nu] (last applied input, 0 at the first
 * call), track [N] int32 or NULL (0 => state_machine != 'track': return Uref untouched,
 * mpc_cbf.py:379-381; < 0 => skip the agent, its outputs are left untouched).
 * pred_x [N, H+1, nx], pred_u [N, H, nu], iters [N] may be NULL.
 * kkt [N] (out, may be NULL): final KKT error.
 * active [N, scb_mpc_active_words(p, M, H)] u64 (out, may be NULL): active set of the NLP at the returned point, one bit
 * per inequality row: bit k*M + j = CBF row of (stage k, obstacle slot j)  (mpc_cbf.py:301-325);  bit H*M + q = simple
 * bound q:  q < 2 H nu: stage k = q / (2 nu), input i = (q % (2 nu)) / 2, even = u_i at its upper bound, odd = lower
 * (mpc_cbf.py:183-232);  then, for the models with state bounds, nsb H bits: t = q - 2 H nu, node k = 1 + t / nsb, r = t % nsb:
 * DynamicUnicycle2D / KinematicBicycle2D* (nsb = 2): r = 0: v <= v_max, 1: v >= -v_max;  VTOL2D (nsb = 5): r = 0: x_dot <= v_max,
 * 1: x_dot >= -v_max, 2: z_dot >= -descent_speed_max, 3: theta <= pitch limit, 4: theta >= -pitch limit.
 * A row is active <=> its multiplier exceeds its value at exit. */
int scb_mpc_active_words(const scb_params* p, int M, int H);
int scb_mpccbf_solve(const scb_params* p, int N, int M, int H,
                     const double* X, const double* Uref, const double* goal, const double* u_prev,
                     const int32_t* track,
                     const double* OBS, long obs_stride_agent, const int32_t* nobs,
                     double* U, int32_t* status, double* pred_x, double* pred_u,
                     int32_t* iters, double* kkt, uint64_t* active, void* stream);
/* Same call with a caller-owned device scratch of scb_mpccbf_workspace_bytes(N) bytes (contents undefined before
 * and after; one workspace per concurrent call).  With it the kernel starts the agents in order of their cold-start
 * CBF margin (most violated first) instead of index order and hands agents to warps dynamically (the work counters
 * live in the workspace): per-agent results are identical, the batch finishes sooner because the long-tailed
 * iteration counts no longer leave a late-started straggler running alone.  workspace == NULL or too small: index
 * order with a static stride, exactly scb_mpccbf_solve. */
size_t scb_mpccbf_workspace_bytes(int N);
/* kernel launches one scb_mpccbf_solve[_ws] call issues (1, or 3 with a schedule: key, counting sort, solve) */
int scb_mpccbf_launch_count(const scb_params* p, int N, int M, int H, int with_workspace);
int scb_mpccbf_solve_ws(const scb_params* p, int N, int M, int H,
                        const double* X, const double* Uref, const double* goal, const double* u_prev,
                        const int32_t* track,
                        const double* OBS, long obs_stride_agent, const int32_t* nobs,
                        double* U, int32_t* status, double* pred_x, double* pred_u,
                        int32_t* iters, double* kkt, uint64_t* active,
                        void* workspace, size_t workspace_bytes, void* stream);
int scb_mpccbf_solve_host(scb_ctx* ctx, const scb_params* p, int N, int M, int H,
                          const double* X, const double* Uref, const double* goal, const double* u_prev,
                          const int32_t* track,
                          const double* OBS, long obs_stride_agent, const int32_t* nobs,
                          double* U, int32_t* status, double* pred_x, double* pred_u,
                          int32_t* iters, double* kkt, uint64_t* active);

/* ---- Backup-CBF QP  (position_control/backup_cbf_qp.py) ---------------------------------- */
/* BackupCBF.solve_control_problem (backup_cbf_qp.py:563-794) for N double-integrator agents in the evade scene
 * (envs/evade_env.py; examples/evade/test_evade.py --algo backupcbf): per agent, rollout of the backup policy
 * (EvadeBackupController, position_control/backup_controller.py:420-575) over n_backup = int(backup_horizon / dt) steps of
 * DoubleIntegrator2D.step with forward-difference sensitivities (:236-318), one CBF row per backup step + the terminal
 * row (:613-673), QP in inputs scaled by a_max with |z| <= 1 (:676-733) and the reference's two fall-backs when the QP
 * is infeasible (:768-783).  The other scenes of that file (drift car: 8-state Fiala-tyre model) are not covered. */
typedef struct scb_backup_params {
  /* EvadeEnv geometry (evade_env.py:62-76) */
  double hallway_length, half_width;
  double pocket_x_min, pocket_x_max, pocket_y_min, pocket_y_max;
  double center_x, center_y;                       /* pocket centre = target of the backup policy */
  double goal_x_min, goal_x_max, goal_y_min, goal_y_max;   /* goal_bounds of EvadeBackupController (test_evade.py:308-313) */
  /* robot_spec (test_evade.py:74-88) + BackupCBF parameters (backup_cbf_qp.py:91-110) */
  double radius, a_max, v_max, safety_margin;
  double Kp, Kd;                                   /* backup_controller.py:449-450 */
  double dt, backup_horizon;
  double alpha, alpha_terminal;                    /* class-K gains, 1.0 and 2.0 */
  double q0, q1;                                   /* Q_u */
  int32_t use_goal;                                /* goal_bounds is not None */
  int32_t n_backup;                                /* N = int(backup_horizon / dt), computed by the caller; <= 252 */
} scb_backup_params;

/* the evade example's defaults (test_evade.py:60-100) */
void scb_backup_params_default(scb_backup_params* p);
size_t scb_backup_params_sizeof(void);
/* uint64 words of one agent's active mask: bit r < n_backup - 1 = safety row of backup step r + 1, bit n_backup - 1 = the
 * terminal row, bits n_backup + {0, 1} = z_{0,1} >= -1, n_backup + {2, 3} = z_{0,1} <= 1 */
int  scb_backup_active_words(int n_backup);

/* X [N, 4], Uref [N, 2] (the nominal input, backup_cbf_qp.py:173-180).  MOV [N, K, 8] (mov_stride_agent = 8 K) or one
 * shared [K, 8] list (stride 0): moving obstacles [x, y, vx, vy, length, width, radius, kind] (kind 0 absent / inactive,
 * 1 rectangle, 2 circle) at x + vx t, y + vy t -- the reference's callable t -> dict (test_evade.py:373-385).
 * Out: U [N, 2]; status [N] (SCB_OPTIMAL, or SCB_INFEASIBLE = QP infeasible, U is the reference's fall-back: clipped u_ref
 * when h_min > 0.01, else the backup policy's input); intervene [N] = is_using_backup() (:757-766); h_min [N] = _last_h_min;
 * optional phi [N, n_backup, 4] = latest_backup_trajectory, rows [N, n_backup, 3] = (lhs_0, lhs_1, rhs) of every row before
 * the reference's ||lhs|| > 1e-6 filter, active [N, scb_backup_active_words]. */
/* With `rows` and `h_min` given the solve is two launches (rollout + rows, then the QP: 8x more agents resident per SM);
 * without, one fused launch that keeps the rows in shared memory. */
int scb_backupcbf_solve(const scb_backup_params* p, int N, int K,
                        const double* X, const double* Uref, const double* MOV, long mov_stride_agent,
                        double* U, int32_t* status, int32_t* intervene, double* h_min,
                        double* phi, double* rows, uint64_t* active, void* stream);
int scb_backupcbf_solve_host(scb_ctx* ctx, const scb_backup_params* p, int N, int K,
                             const double* X, const double* Uref, const double* MOV, long mov_stride_agent,
                             double* U, int32_t* status, int32_t* intervene, double* h_min,
                             double* phi, double* rows, uint64_t* active);

/* ---- gatekeeper / MPS shields  (shielding/gatekeeper.py, shielding/mps.py) ----------------- */
/* Gatekeeper.solve_control_problem (gatekeeper.py:553-672) and MPS.solve_control_problem (mps.py:59-160) for N
 * double-integrator agents in the evade scene, with an external nominal trajectory per agent (set_nominal_trajectory, the
 * way examples/evade/test_evade.py:424-426 drives both): candidate = nominal prefix + backup-policy rollout, validated
 * state by state against the walls, the bullet's current hitbox and the moving obstacles at t = k dt; gatekeeper searches
 * the nominal horizon backwards in steps of `discount_steps`, MPS tries one nominal step every call.  The shield's state
 * (committed trajectory, event timing) stays in caller-owned device arrays between calls. */
typedef struct scb_shield_params {
  scb_backup_params scene;        /* geometry, backup policy, robot, dt, n_backup = int(backup_horizon / dt), safety_margin
                                     (the `safety_margin` constructor argument, gatekeeper.py:45) */
  double  event_offset;           /* gatekeeper.py:67 */
  int32_t mode;                   /* 0 gatekeeper, 1 MPS */
  int32_t discount_steps;         /* max(1, int(horizon_discount / dt)), horizon_discount = 5 dt by default (:68, 601) */
  int32_t nom_cap;                /* T: capacity of the nominal buffers in STEPS */
  int32_t reserved;
} scb_shield_params;

typedef struct scb_shield_state { /* device pointers, caller-owned, persistent across calls */
  double*  CU;                    /* [N, 2, T + n_backup, 2]      committed_u_traj, double-buffered: agent i's is CU[i, cbuf[i]] */
  double*  CX;                    /* [N, 2, T + n_backup + 1, 4]  committed_x_traj (same buffering), or NULL */
  int32_t* clen;                  /* [N] len(committed_u_traj); -1 = no committed trajectory yet (:571) */
  int32_t* cidx;                  /* [N] current_time_idx */
  int32_t* nsteps;                /* [N] actual_nominal_steps (committed_horizon = nsteps dt) */
  double*  next_event;            /* [N] next_event_time */
  int32_t* cbuf;                  /* [N] which of the two buffers holds the committed trajectory (start: 0) */
  int32_t* work;                  /* [N + 1] scratch (contents irrelevant between calls), or NULL.  With it a gatekeeper step of
                                     a large batch is two launches: candidate 0 of every agent with a thread per agent, then
                                     the remaining candidates, a lane group per agent that still needs one. */
} scb_shield_state;

size_t scb_shield_params_sizeof(void);
/* One control step.  X [N, 4]; NOMX [N, T + 1, 4], NOMU [N, T, 2] = nominal_x_traj / nominal_u_traj, nom_len [N] (may be
 * NULL = T + 1) = number of nominal STATES available (0 = empty trajectory); MOV / mov_stride_agent as in
 * scb_backupcbf_solve; STAT [N, 5] (may be NULL) = x_min, x_max, y_min, y_max, active of the obstacle hitbox checked with
 * the bare robot radius at every candidate state (evade_env.py:454-485: the bullet where it is now).
 * Out: U [N, 2]; using_backup [N] = is_using_backup() after the call (gatekeeper.py:741-744, mps.py:55-57). */
int scb_shield_step(const scb_shield_params* p, const scb_shield_state* s, int N, int K,
                    const double* X, const double* NOMX, const double* NOMU, const int32_t* nom_len,
                    const double* MOV, long mov_stride_agent, const double* STAT,
                    double* U, int32_t* using_backup, void* stream);

/* ---- closed loop: the rest of LocalTrackingController.control_step()  (tracking.py:559-668) ---- */
/* Everything either side of the solve, for N agents on the device, so that run_all_steps
 * (tracking.py:711-747) never leaves the GPU:
 *   waypoint state machine + update_goal            tracking.py:497-535, 569-578
 *   get_nearest_unpassed_obs                        tracking.py:345-403
 *   nominal_input / stop / rotate_to / has_stopped  robots/<model>.py via robots/robot.py:401-433
 *   VelocityTrackingYaw (SingleIntegrator2D yaw)    attitude_control/velocity_tracking_yaw.py:35-62
 *   is_collide_unknown (known obstacles)            tracking.py:445-495
 *   robot.step                                      robots/robot.py:441-453 -> robots/<model>.py step
 *   step_dyn_obs                                    dynamic_env/main.py:54-58, 152
 *   return code                                     tracking.py:627-668
 */
enum scb_state_machine { SCB_SM_IDLE = 0, SCB_SM_TRACK = 1, SCB_SM_STOP = 2, SCB_SM_ROTATE = 3 };  /* tracking.py:49 */
enum scb_controller { SCB_CTRL_CBF_QP = 0, SCB_CTRL_OPTIMAL_DECAY = 1, SCB_CTRL_MPC_CBF = 2 };

/* Obstacle selection only.  SCENE [K, 7] is the caller's `self.obs` (shared by all agents, or
 * [N, K, 7] with scene_stride_agent = 7 K); yaw [N] is robot.yaw (NULL: the model's own heading
 * state, 0 for SingleIntegrator2D).  Out: OBS [N, M, 7] = the (at most M) nearest unpassed
 * obstacles in distance order, rows beyond nobs[i] padded with the reference's dummy
 * [1000, 1000, 0, 0, 0, 0, 0] (mpc_cbf.py:346); nobs [N] (-1 when K == 0: "obs is None");
 * idx [N, M] int32 (may be NULL) = scene index of every selected row (-1 = pad). */
int scb_select_obstacles(const scb_params* p, int N, int K, int M,
                         const double* X, const double* yaw,
                         const double* SCENE, long scene_stride_agent,
                         double* OBS, int32_t* nobs, int32_t* idx, void* stream);

/* All device pointers, caller-owned (e.g. torch tensors); float64 unless noted.  The solve
 * buffers (Uref, OBS, nobs, U, status) are the same arrays the *_solve entry points take. */
typedef struct scb_track {
  int32_t controller;            /* enum scb_controller */
  int32_t N, K, M, W, H;         /* agents, scene obstacles, obstacle slots (num_constraints, tracking.py:134-138),
                                    max waypoints per agent, MPC horizon */
  int32_t enable_rotation;       /* tracking.py:41 */
  int32_t dynamic_obs;           /* 1: SCENE[:, 0:2] += SCENE[:, 3:5] dt after the selection (dynamic_env/main.py:152) */
  int32_t att_velocity_tracking; /* SingleIntegrator2D: 1 = VelocityTrackingYaw drives yaw in 'track' (tracking.py:156-181) */
  int32_t mpc_strict;            /* mpc_cbf only.  0 (default) = the reference's semantics: MPCCBF.status is hard-wired
                                    'optimal' (mpc_cbf.py:10, 400), so the returned input is always stepped.  1 = an MPC solve
                                    that did not end SCB_OPTIMAL makes control_step return -2 without stepping, exactly like
                                    a failed QP (tracking.py:627-634).  Either way `mpc_fail` counts such solves.
                                    (Superellipsoid rows in MPC are enabled by scb_params.mpc_superellipsoid.) */
  double reached_threshold;      /* 0.3  tracking.py:52 */
  double rotation_threshold;     /* 0.1  tracking.py:50 */
  double k_omega, k_a, k_v;      /* nominal_input gains (robots/robot.py:401; optimal decay: 3.0, 0.5, 0.5 tracking.py:601-602) */
  double k_a_stop;               /* DynamicUnicycle2D stop() gain: robot_spec['nominal_k_a'] or 1.0 (dynamic_unicycle2D.py:106-111) */
  double w_max;                  /* SingleIntegrator2D rotate_to / VelocityTrackingYaw clip */
  double att_kp;                 /* VelocityTrackingYaw kp (1.5) */
  double wheel_base, delta_max;  /* KinematicBicycle2D nominal_input (kinematic_bicycle2D.py:55-59, 125-147) */
  /* per-agent tracker state (in/out) */
  double*  X;                    /* [N, nx]   robot.X */
  double*  yaw;                  /* [N]       robot.yaw */
  int32_t* sm;                   /* [N]       state_machine */
  int32_t* wp_idx;               /* [N]       current_goal_index */
  const double*  WP;             /* [N, W, 3] waypoints (after filter_waypoints) */
  const int32_t* nwp;            /* [N] */
  double*  goal;                 /* [N, 2]    ([N, 3] for Quad3D) self.goal, valid when has_goal; same layout as scb_mpccbf_solve's goal */
  int32_t* has_goal;             /* [N]       0 <=> self.goal is None */
  double*  u_att;                /* [N]       self.u_att; NaN <=> None */
  double*  u_prev;               /* [N, nu]   MPC: last applied MPC input (do-mpc u0) */
  int32_t* ret;                  /* [N]       control_step() return value of the last executed step: 0, -1, -2 */
  int32_t* done;                 /* [N]       latched: run_all_steps' loop has broken (ret in {-1,-2}); such agents are frozen */
  int32_t* nsteps;               /* [N]       control steps executed so far */
  /* world */
  double*  SCENE;                /* [K, 7]    self.obs, shared; stepped in place when dynamic_obs */
  /* solve buffers */
  double*  Uref;                 /* [N, nu] */
  double*  OBS;                  /* [N, M, 7] */
  int32_t* nobs;                 /* [N] */
  double*  U;                    /* [N, nu]   get_control_input() */
  int32_t* status;               /* [N] */
  uint64_t* active;              /* [N, scb_active_words(M, nu)] or NULL */
  int32_t* track_flag;           /* [N]       MPC only: scratch (state_machine == 'track') */
  int32_t* mpc_iters;            /* [N]       MPC only, may be NULL */
  void*    mpc_ws;               /* MPC only, may be NULL: scheduling scratch, scb_mpccbf_workspace_bytes(N) bytes */
  uint64_t mpc_ws_bytes;
  int32_t* mpc_fail;             /* [N]       MPC only, may be NULL (in/out): number of control steps of this agent whose MPC solve
                                              did not end SCB_OPTIMAL (infeasible / iteration limit / numerical) */
} scb_track;

size_t scb_track_sizeof(void);
/* One control_step() for every agent that is not done: 3 launches (+1 when dynamic_obs) on `stream`. */
int scb_control_step(const scb_params* p, const scb_track* t, void* stream);
/* n_steps control steps back to back (run_all_steps' loop body; per-agent break = the done latch).  For the QP
 * controllers with M <= 60 and K <= 512 this is ONE kernel launch (a warp keeps its agent for all n_steps steps,
 * each CTA steps its own shared-memory copy of a moving scene); otherwise n_steps x scb_control_step.  Both give
 * bit-identical results.  Environment SCB_TRACK_FUSED=0 forces the per-step path. */
int scb_run_all_steps(const scb_params* p, const scb_track* t, int n_steps, void* stream);
/* number of kernel launches scb_run_all_steps(p, t, n_steps) issues (for benchmarks' gpu_launches) */
long scb_run_all_steps_launches(const scb_params* p, const scb_track* t, int n_steps);

#ifdef __cplusplus
}
#endif
#endif /* SCB_H_ */
