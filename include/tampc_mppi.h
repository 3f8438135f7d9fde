/*
 * tampc_mppi.h — C ABI of the B200-native MPPI rollout hot path (libtampc_mppi.so).
 *
 * Drop-in boundary for the `pytorch_mppi` `MPPI.command()` call that tampc's controllers make
 * (reference: tampc/controller/online_controller.py:497,93 and tampc/controller/controller.py:428).
 * The reference hands MPPI three Python callbacks; here the same information crosses the boundary as
 * plain-old-data descriptors (see tampc_b200/spec.py for the host-side mirror):
 *
 *   tampc_model_desc  <- OnlineMPC._apply_dynamics (online_controller.py:60-79) = bad-state guard,
 *                        HybridDynamicsModel nominal network (dynamics/hybrid_model.py:204-221,
 *                        dynamics/model.py:147-170) behind the preprocessor chain of util.py:408-411
 *                        (RobustMinMax scalers + InvariantTransformer/RexExtract, transform/invariant.py:
 *                        628-640,792-795,926-931), constrain_state (scripts/push_main.py:190-193)
 *   tampc_cost_desc   <- MPC._running_cost / _terminal_cost (controller.py:245-259) = ComposeCost of
 *                        CostQROnlineTorch + TrapSetCost (cost.py:11-36,154-189,213-235), or the recovery
 *                        cost GoalSetCost/min_cost (cost.py:114-147; online_controller.py:370-402)
 *   tampc_mppi_config <- pytorch_mppi.MPPI.__init__ keywords + ExperimentalMPPI extras
 *                        (controller.py:316-321,418-421; scripts/push_main.py:101-112)
 *
 * Conventions: every entry point returns 0 (TAMPC_OK) or a negative tampc_status; the message is available
 * from tampc_mppi_last_error().  All host pointers are float64 / int32 plain arrays owned by the caller and
 * are not retained past the call.  A handle is bound to one CUDA device and is not thread-safe.  There is no
 * CPU fallback: without a usable sm_100 device tampc_mppi_create fails.
 */
#ifndef TAMPC_MPPI_H_
#define TAMPC_MPPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAMPC_ABI_VERSION 1

#if defined(__GNUC__)
#define TAMPC_API __attribute__((visibility("default")))
#else
#define TAMPC_API
#endif

#define TAMPC_MAX_NX 8
#define TAMPC_MAX_NU 4
#define TAMPC_MAX_TRAPS 64
#define TAMPC_MAX_GOAL_SETS 2
#define TAMPC_MAX_SET_GOALS 8
#define TAMPC_MAX_LAYERS 4
#define TAMPC_MAX_HORIZON 128
#define TAMPC_GP_MAX_POINTS 64 /* OnlineGPMixing.max_data_points defaults to 50 (online_model.py:327) */
#define TAMPC_GP_MAX_DIM 8
#define TAMPC_MAX_ROLLOUT_SAMPLES 32
#define TAMPC_GP_FIT_MAX_POINTS 50 /* hyper-parameter fit: OnlineGPMixing.max_data_points (online_model.py:327) */

typedef enum tampc_status {
  TAMPC_OK = 0,
  TAMPC_ERR_INVALID = -1,      /* bad argument / inconsistent descriptor */
  TAMPC_ERR_UNSUPPORTED = -2,  /* shape or option with no compiled kernel (never a silent fallback) */
  TAMPC_ERR_CUDA = -3,         /* CUDA runtime error (message carries cudaGetErrorString) */
  TAMPC_ERR_STATE = -4,        /* call order: model / cost / nominal not set yet */
  TAMPC_ERR_NO_DEVICE = -5
} tampc_status;

/* arithmetic of the rollout kernel */
typedef enum tampc_precision {
  TAMPC_PREC_F64 = 0,   /* networks, state and cost in float64 — the reference's arithmetic (controller.py:197);
                           network layers run on the FP64 tensor pipe (DMMA) where the shape has an mma kernel */
  TAMPC_PREC_MIXED = 1, /* networks in float32 (packed FFMA2), state + cost accumulation in float64 */
  TAMPC_PREC_F32 = 2,   /* everything in float32 */
  TAMPC_PREC_TF32X3 = 3, /* networks on the tensor cores as 3xTF32 (error-compensated split, ~float32 accuracy),
                            state and cost terms in float32, cost accumulated in float64 */
  TAMPC_PREC_F16X3 = 4   /* the same split on the fp16 tensor path (mma.sync.m16n8k16: 11 significant bits like TF32,
                            half the tensor instructions); low parts carried scaled by 2^11, high parts saturate at
                            65504.  Shapes with resident-weight tensor kernels only */
} tampc_precision;

/* where the per-sample action perturbations come from */
typedef enum tampc_noise_mode {
  TAMPC_NOISE_PHILOX = 0,   /* on-device Philox4x32-10 keyed by (seed, call, global sample, t): N(mu, Sigma) */
  TAMPC_NOISE_INJECTED = 1, /* caller supplies the (K,T,nu) draw (the reference's noise_dist.sample((K,T))) */
  TAMPC_NOISE_ACTIONS = 2   /* caller supplies final actions (K,T,nu): _compute_rollout_costs(perturbed_actions) */
} tampc_noise_mode;

typedef enum tampc_dyn_kind {
  TAMPC_DYN_MLP = 0,     /* x' = x + inv_out( net( aff_in([x,u]) ) ) */
  TAMPC_DYN_REX = 1,     /* RexExtract latent dynamics */
  TAMPC_DYN_PENDULUM = 2 /* scripts/pendulum.py:223-241 */
} tampc_dyn_kind;

typedef enum tampc_cost_kind {
  TAMPC_COST_GOAL_TRAP = 0, /* MPC._running_cost: QR goal cost + trap-set cost (+ terminal) */
  TAMPC_COST_RECOVERY = 1,  /* OnlineMPPI._recovery_running_cost: weighted min-distance goal sets */
  TAMPC_COST_APF = 2        /* APFVO / APFSP potentials on the NEXT state only (online_controller.py:554-664; cost.py:236-289):
                               [use_goal] d'Qd  +  gain * sum_i clamp((1/rho_i - 1/rho0)^2, 0, 1e5) [rho_i <= rho0]  +  [helicoid] */
} tampc_cost_kind;

/* Sequential(Linear, LeakyReLU, ..., Linear).  dims[0..n_layers]; w_off/b_off index the float64 `params`
 * array passed to tampc_mppi_set_model; W is row-major (out, in) as in torch.nn.Linear.weight. */
typedef struct tampc_mlp_desc {
  int32_t n_layers;
  int32_t dims[TAMPC_MAX_LAYERS + 1];
  int32_t w_off[TAMPC_MAX_LAYERS];
  int32_t b_off[TAMPC_MAX_LAYERS];
} tampc_mlp_desc;

/* per-dim affine y = x*scale + offset; off < 0 means identity */
typedef struct tampc_affine_desc {
  int32_t scale_off;
  int32_t offset_off;
} tampc_affine_desc;

typedef struct tampc_model_desc {
  int32_t kind; /* tampc_dyn_kind */
  int32_t nx, nu;
  int32_t nv;              /* REX: latent output width (decoder emits nx*nv) */
  double leaky_slope;      /* LeakyReLU negative slope shared by all networks, in [0,1] */
  tampc_mlp_desc net;      /* MLP: dynamics net; REX: latent dynamics z'->v' */
  tampc_mlp_desc encoder;  /* REX: (x,u)->z */
  tampc_mlp_desc decoder;  /* REX: e -> C (nx*nv) */
  int32_t ext_w_off, ext_b_off, ne; /* REX: x_extractor Linear(nx, ne) */
  tampc_affine_desc aff_in;  /* MLP: on [x,u]; REX: on z (forward) */
  tampc_affine_desc aff_out; /* MLP: on dx (inverted); REX: on v (inverted) */
  tampc_affine_desc aff_y;   /* REX: pre-invariant y scaler on dx (inverted) */
  uint32_t wrap_mask;        /* constrain_state: state dims passed through angle_normalize */
  int32_t has_bad_state_guard;
  double bad_state_norm;     /* online_controller.py:63 (1e5) */
  double pendulum[6];        /* g, m, l, dt, |u| clamp, |thdot| clamp */
} tampc_model_desc;

/* Online GP residual on top of the nominal network: OnlineGPMixing with independent outputs
 * (tampc/dynamics/online_model.py:286-315,324-457), frozen at its current data and hyper-parameters.  Per output d:
 * kernel s_d exp(-|a-b|^2 / (2 l_d^2)), Gaussian noise n_d, mean function = the nominal model.  The library prepares
 * alpha_d = (K_d + n_d I)^-1 resid_d and the inverse Cholesky factor; the rollout kernel evaluates the posterior mean
 * and variance per sample-step and draws Normal(mean, stddev) (online_model.py:435-449).  The hyper-parameter fit
 * (online_model.py:412-430) stays with the caller. */
typedef enum tampc_gp_space {
  TAMPC_GP_AT_NETWORK = 0, /* inputs = the rows the network sees, outputs add to what it emits (GP built on the nominal
                              datasource, hybrid_model.py:110-112) */
  TAMPC_GP_ORIGINAL = 1    /* inputs = [x,u]*in_scale + in_offset, outputs are scaled dx: dx_i += delta_i / out_scale_i
                              (temporary local model behind util.no_tsf_preprocessor(), hybrid_model.py:189-199) */
} tampc_gp_space;

typedef struct tampc_gp_desc {
  int32_t n_points; /* 0 switches the residual off */
  int32_t n_in, n_out;
  int32_t space;    /* tampc_gp_space */
  int32_t sample;   /* OnlineGPMixing.sample_dynamics: 1 = draw, 0 = posterior mean */
  int32_t x_off, resid_off; /* (n_points, n_in) inputs and (n_points, n_out) targets minus the mean function there,
                               row-major, in the float64 `params` array */
  double lengthscale[TAMPC_GP_MAX_DIM], outputscale[TAMPC_GP_MAX_DIM], noise[TAMPC_GP_MAX_DIM];
  double in_scale[TAMPC_GP_MAX_DIM], in_offset[TAMPC_GP_MAX_DIM], out_scale[TAMPC_GP_MAX_DIM]; /* TAMPC_GP_ORIGINAL */
  /* a sampled model makes ExperimentalMPPI's M rollout replicas differ (controller.py:317,355-380) */
  int32_t rollout_samples; /* M in [1, TAMPC_MAX_ROLLOUT_SAMPLES] */
  double rollout_var_cost, rollout_var_discount;
} tampc_gp_desc;

typedef struct tampc_cost_desc {
  int32_t kind; /* tampc_cost_kind */
  int32_t nx, nu;
  uint32_t ang_mask;                /* compare_to_goal dims wrapped once into [-pi,pi] */
  double dist_scale[TAMPC_MAX_NX];  /* state_dist(diff) = || diff * dist_scale || */
  uint32_t usim_mask;               /* action dims of the clamped cosine similarity */
  /* goal + trap-set program */
  double Q[TAMPC_MAX_NX * TAMPC_MAX_NX]; /* row-major nx*nx in the leading block */
  double R[TAMPC_MAX_NU * TAMPC_MAX_NU];
  double goal[TAMPC_MAX_NX];
  int32_t use_trap_cost;
  int32_t n_traps;
  double trap_x[TAMPC_MAX_TRAPS][TAMPC_MAX_NX];
  double trap_u[TAMPC_MAX_TRAPS][TAMPC_MAX_NU];
  double trap_weight;
  int32_t has_terminal;
  double terminal_multiplier;
  /* recovery program */
  int32_t n_sets;
  int32_t set_size[TAMPC_MAX_GOAL_SETS];
  double set_goals[TAMPC_MAX_GOAL_SETS][TAMPC_MAX_SET_GOALS][TAMPC_MAX_NX];
  double set_weight[TAMPC_MAX_GOAL_SETS];
  double recovery_scale;
  /* APF program (TAMPC_COST_APF): goal / Q as above (no action term: the reference passes U=None); repulsive states in
   * trap_x[0..n_traps) (APFVO's trap_set holds states only, online_controller.py:580-581) */
  int32_t apf_use_goal;   /* 1: add the goal cost d'Qd (APFVO always; APFSP when no obstacle is active) */
  int32_t apf_helicoid;   /* 1: APFHelicoid2D(rxy, oxy) (cost.py:268-289) */
  int32_t apf_clockwise;
  double apf_max_dist;    /* ArtificialRepulsionCost.max_dist_influence (rho0) */
  double apf_gain;        /* ArtificialRepulsionCost.gain */
  double apf_rxy[2], apf_oxy[2];
} tampc_cost_desc;

typedef struct tampc_mppi_config {
  int32_t abi_version; /* TAMPC_ABI_VERSION */
  int32_t device;      /* CUDA ordinal */
  int32_t nx, nu;
  int32_t num_samples; /* K: samples of the WHOLE controller */
  int32_t sample_offset, num_local; /* this handle rolls out samples [offset, offset+num_local) (multi-GPU shard) */
  int32_t max_horizon; /* capacity for change_horizon */
  int32_t precision;   /* tampc_precision */
  double lambda_;
  double noise_mu[TAMPC_MAX_NU];
  double noise_sigma[TAMPC_MAX_NU * TAMPC_MAX_NU]; /* covariance, row-major nu*nu */
  int32_t has_bounds;
  double u_min[TAMPC_MAX_NU], u_max[TAMPC_MAX_NU];
  double u_init[TAMPC_MAX_NU];
  int32_t sample_null_action;
  uint64_t seed;
  void* stream; /* cudaStream_t to run on; NULL = a stream owned by the handle */
} tampc_mppi_config;

/* One rank's contribution to the softmax update: min cost, its global sample index, the normaliser and the
 * weighted post-clamp noise sum, both relative to THIS rank's beta (combine rescales by exp(-(beta_g-beta)/lambda)). */
typedef struct tampc_partial {
  double beta;
  double eta;
  int64_t argmin;
  int64_t count;
  double wsum[TAMPC_MAX_HORIZON * TAMPC_MAX_NU];
} tampc_partial;

typedef enum tampc_query {
  TAMPC_Q_COST_TOTAL = 0, /* double[num_local]            : ExperimentalMPPI.cost_total */
  TAMPC_Q_PERTURBED = 1,  /* double[num_local*T*nu]       : .perturbed_action (post-clamp) */
  TAMPC_Q_NOMINAL = 2,    /* double[T*nu]                 : .U */
  TAMPC_Q_OMEGA = 3,      /* double[num_local]            : .omega */
  TAMPC_Q_STATS = 4,      /* double[4]: beta, eta, argmin (global index), launches so far */
  TAMPC_Q_STATES = 5,     /* double[num_local*T*nx]       : rollout states of the last command_ex(keep_states=1) */
  TAMPC_Q_NOISE = 6       /* double[num_local*T*nu]       : .noise = perturbed_action - U, post-clamp (controller.py:394) */
} tampc_query;

typedef struct tampc_mppi tampc_mppi;

TAMPC_API const char* tampc_mppi_last_error(const tampc_mppi* h); /* h may be NULL: error of the last failed create */
TAMPC_API int tampc_mppi_abi_version(void);

TAMPC_API int tampc_mppi_create(const tampc_mppi_config* cfg, tampc_mppi** out);
TAMPC_API void tampc_mppi_destroy(tampc_mppi* h);

/* dynamics / cost / nominal sequence; each may be called again at any time between commands
 * (the reference mutates cost, horizon and trap set between calls: online_controller.py:400-402,416-418,353,462) */
TAMPC_API int tampc_mppi_set_model(tampc_mppi* h, const tampc_model_desc* desc, const double* params, size_t n_params);
TAMPC_API int tampc_mppi_set_cost(tampc_mppi* h, const tampc_cost_desc* desc);
/* online GP residual (replaces OnlineDynamicsModel.predict, online_model.py:67-99, inside the rollout); call after
 * every OnlineGPMixing update (online_model.py:386-410).  desc->n_points == 0 returns to the nominal network. */
TAMPC_API int tampc_mppi_set_gp(tampc_mppi* h, const tampc_gp_desc* desc, const double* params, size_t n_params);
/* Hyper-parameter fit of that GP on the device: `iters` Adam steps (lr as given, torch defaults otherwise) on the negative
 * exact marginal log-likelihood — replaces OnlineGPMixing._fit_params (online_model.py:412-430; 15 iterations per control
 * step, 150 at a reset).  Handle-free: runs on `device`.  X (n_points, n_in) inputs, resid (n_points, n_out) targets minus
 * the mean function; raw_params / adam_m / adam_v hold [raw lengthscale[n_out] | raw outputscale[n_out] | raw task
 * noise[n_out] | raw global noise] (softplus-constrained, noises additionally +1e-4 each: gpytorch's parameterisation) and
 * are updated in place together with *adam_step; loss_out[0..1] = loss at the last and the last-but-one iteration. */
TAMPC_API int tampc_gp_fit(int32_t device, int32_t n_points, int32_t n_in, int32_t n_out, const double* X, const double* resid, int32_t iters,
                           double lr, double* raw_params, double* adam_m, double* adam_v, int64_t* adam_step, double* loss_out);
TAMPC_API const char* tampc_gp_fit_last_error(void);
/* standard normals behind the GP draws of the NEXT command / rollout, (num_local, M, T, n_out) float64, instead of the
 * on-device Philox stream (the equivalent of patching Normal.sample in the reference; parity tests).  One-shot. */
TAMPC_API int tampc_mppi_set_gp_draws(tampc_mppi* h, const double* eps, size_t n);
TAMPC_API int tampc_mppi_set_nominal(tampc_mppi* h, const double* U, int32_t horizon); /* (horizon, nu): MPPI.U, sets T */
TAMPC_API int tampc_mppi_get_nominal(tampc_mppi* h, double* U);

/* MPPI.command(state): shift U, sample + roll out K trajectories, softmax-update U, return U[0] in action[nu].
 * noise: NULL (Philox) or host float64 (K_local,T,nu) when noise_mode is INJECTED. */
TAMPC_API int tampc_mppi_command(tampc_mppi* h, const double* state, int32_t noise_mode, const double* noise, double* action);

/* Split form for multi-GPU sharding: begin() rolls out this handle's shard and reduces it to ONE tampc_partial
 * written to `partial_dev` (caller-owned DEVICE memory of sizeof(tampc_partial), e.g. a slot of the all-gather
 * buffer); the host plumbing all-gathers the partials (NCCL via torch.distributed); finish() combines `world`
 * partials in rank order (bitwise identical on every rank) and updates U.  Work is enqueued on the handle's
 * stream; begin() does not synchronise. */
TAMPC_API int tampc_mppi_command_begin(tampc_mppi* h, const double* state, int32_t noise_mode, const double* noise, int32_t keep_states,
                             void* partial_dev);
TAMPC_API int tampc_mppi_command_finish(tampc_mppi* h, const void* partials_dev, int32_t world, double* action);

/* Peer-memory exchange (one node, NVLink): replaces the all-gather between begin() and finish() by P2P stores done
 * inside the merge kernel.  comm_init allocates this rank's exchange buffer and returns its CUDA IPC handle
 * (TAMPC_IPC_HANDLE_BYTES); the host plumbing all-gathers the handles of the world (any transport) and passes them
 * to comm_connect in rank order.  After that tampc_mppi_command() of a sharded handle is complete by itself: every
 * rank calls it with the same state and obtains the same action. */
#define TAMPC_IPC_HANDLE_BYTES 64
#define TAMPC_MAX_RANKS 8
TAMPC_API int tampc_mppi_comm_init(tampc_mppi* h, int32_t rank, int32_t world, void* handle_out);
TAMPC_API int tampc_mppi_comm_connect(tampc_mppi* h, const void* handles);
/* Unmap the peers (after a failed connect on any rank, or a peer-exchange timeout: the command that timed out leaves
 * the nominal sequence as it was and further sharded commands are refused until comm_init + comm_connect ran again). */
TAMPC_API int tampc_mppi_comm_disconnect(tampc_mppi* h);

/* ExperimentalMPPI._compute_rollout_costs(perturbed_actions): costs of caller-given actions (K_local,T,nu), no U update */
TAMPC_API int tampc_mppi_rollout_costs(tampc_mppi* h, const double* state, const double* actions, int32_t horizon, double* cost_out,
                             double* states_out /* may be NULL; (K_local,T,nx) */);
/* MPPI.get_rollouts(state)[0]: the T states reached by the current nominal U (controller.py:434-435) */
TAMPC_API int tampc_mppi_get_rollouts(tampc_mppi* h, const double* state, double* states_out /* (T,nx) */);

/* ControllerWithModelPrediction.predict_next_state (controller.py:312-313, online_controller.py:47-58): one step of the
 * nominal dynamics (guard, network, constrain_state) from `state` under `action`; next_state[nx]. */
TAMPC_API int tampc_mppi_predict(tampc_mppi* h, const double* state, const double* action, double* next_state);
/* OnlineMPC.predict_next_state (online_controller.py:47-58): the reference swaps the ORIGINAL nominal network in for this
 * call (hybrid dynamics with class NOMINAL) — no online GP residual, no bad-state guard, no constrain_state.  Its result
 * feeds diff_predicted, which drives the trap / local-model state machine, so it must stay the nominal prediction even
 * while a GP residual is installed for the rollouts. */
TAMPC_API int tampc_mppi_predict_nominal(tampc_mppi* h, const double* state, const double* action, double* next_state);

/* One-step action selection of the APF baselines (online_controller.py:583-590, 657-664): push `n` candidate actions
 * (n <= num_local, host (n, nu)) through _apply_dynamics from `state`, evaluate the installed cost program on the next
 * states, take the argmin ON THE DEVICE (lowest index wins ties, like torch.argmin).  index_out: winning row; cost_out:
 * its cost; next_state_out[nx]: the state it leads to.  One rollout launch + one argmin launch, one small D2H. */
TAMPC_API int tampc_mppi_select_action(tampc_mppi* h, const double* state, const double* actions, int32_t n, int64_t* index_out, double* cost_out,
                                       double* next_state_out);
/* goal_cost(x) and trap_cost(x) of OnlineMPPI.normalize_trapset_cost_to_state (online_controller.py:517-530) for n states
 * (host (n, nx)): the QR goal cost without an action term and the raw trap-set cost with c_u = 1 (U=None, cost.py:177),
 * unweighted.  Uses the installed GOAL_TRAP cost program. */
TAMPC_API int tampc_mppi_state_costs(tampc_mppi* h, const double* states, int32_t n, double* goal_cost_out, double* trap_cost_out);

TAMPC_API int tampc_mppi_query(tampc_mppi* h, int32_t what, double* out, size_t n_doubles);
/* which rollout kernel family and arithmetic a command of this handle runs (static string; "none" before set_model):
 * the precision asked for in the config is a request, e.g. TF32X3 on the analytic pendulum runs the FMA kernel */
TAMPC_API const char* tampc_mppi_kernel_name(tampc_mppi* h);
/* development aid: 64 SM-clock stamps left by the roles of the tcgen05 kernel's first CTA at horizon step 2 when the
 * handle's model was installed with TAMPC_TC_PROF=1 in the environment (tools/tc_timeline.py); zeros otherwise */
TAMPC_API int tampc_mppi_debug_stamps(tampc_mppi* h, int64_t* out);
/* device pointer to the per-sample costs of the last rollout (double[num_local]); valid until the next call */
TAMPC_API int tampc_mppi_device_costs(tampc_mppi* h, void** dev_ptr);
TAMPC_API int tampc_mppi_synchronize(tampc_mppi* h);

/* Device-resident variant used for kernel-only timing: state already on the device, no host copies, no sync.
 * Equivalent to command() minus the H2D of `state` and the D2H of the action. */
TAMPC_API int tampc_mppi_command_resident(tampc_mppi* h, int32_t noise_mode);
TAMPC_API int tampc_mppi_set_state_resident(tampc_mppi* h, const double* state);
/* Only the fused rollout kernel of a resident command (no reduction, U untouched): what bench.py times for the
 * roofline of the dominant kernel. */
TAMPC_API int tampc_mppi_rollout_resident(tampc_mppi* h, int32_t noise_mode);

#ifdef __cplusplus
}
#endif
#endif /* TAMPC_MPPI_H_ */
