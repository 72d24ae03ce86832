/*
 * rsrl_b200.h — C ABI of the B200-native vectorised RL step engine.
 *
 * This is the drop-in boundary for ONE hot path of tspooner/rsrl (all file:line
 * citations are relative to the reference checkout):
 *
 *     Domain::transition            rsrl_domains/src/lib.rs:436-446
 *       -> basis.project            lfa 0.15 (external crate; call sites rsrl/src/fa/linear.rs:310,322,337,389)
 *       -> LFA evaluate             rsrl/src/fa/linear.rs:303-324,353-363
 *       -> TD error + SGD update    rsrl/src/control/td/{q_learning.rs:51-71,sarsa.rs:53-75,expected_sarsa.rs:45-66}
 *       -> Greedy/eps-greedy sample rsrl/src/policies/{greedy.rs:74-84,epsilon_greedy.rs:69-83}
 *
 * The reference has no FFI on this path: the boundary it exposes is the Rust
 * trait surface (Domain, Handler<&Transition>, Policy, Enumerable,
 * Parameterised).  Each entry point below names the trait method(s) it
 * replaces; `shim/src/lib.rs` and INTEGRATION.md show the `extern "C"` block a
 * maintainer would add on the Rust side.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types, no exceptions cross the ABI;
 *   - every function returns an int status (0 = RSRL_OK, < 0 = error) unless it
 *     returns a pointer/version; the message is in rsrl_last_error() (thread local);
 *   - all `double*`/`int32_t*` buffers are HOST memory owned by the caller and
 *     only have to stay alive for the duration of the call;
 *   - states are row-major  N x D  f64 (reference: Vec<f64> per env),
 *     weights are row-major  F x A  f64 == Parameterised::weights_view()
 *     (rsrl/src/fa/linear.rs:293-301, rsrl/src/params/mod.rs:116-134);
 *     PER_ENV weight mode: N x F x A;
 *   - a handle is not thread-safe (mirrors Shared<T> = Rc<RefCell<T>> being
 *     !Send + !Sync, rsrl/src/core.rs:13-15);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with RSRL_ENODEVICE.
 */
#ifndef RSRL_B200_H
#define RSRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSRL_ABI_VERSION 1
#define RSRL_MAX_DIM 4      /* state dimension of the supported domains (2/4/4) */
#define RSRL_MAX_ACTIONS 3  /* MountainCar 3, CartPole 2, Acrobot 3 */

typedef enum rsrl_status {
    RSRL_OK = 0,
    RSRL_EINVAL = -1,       /* bad argument / inconsistent config */
    RSRL_ECUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    RSRL_ENOMEM = -3,
    RSRL_EUNSUPPORTED = -4, /* combination not built */
    RSRL_ENONFINITE = -5,   /* a Q vector had no valid maximum (all NaN): the reference panics
                               in utils.rs:70-76 `expect("No valid maxima ...")` */
    RSRL_ECOMM = -6,        /* NCCL / peer-memory error */
    RSRL_ENODEVICE = -7     /* no CUDA device: there is no CPU fallback */
} rsrl_status_t;

/* rsrl_domains/src/{mountain_car/discrete.rs, cart_pole.rs, acrobot.rs} */
typedef enum rsrl_domain {
    RSRL_MOUNTAIN_CAR = 0, RSRL_CART_POLE = 1, RSRL_ACROBOT = 2,
    /* component-level only (rsrl_domain_ex_*): rsrl_domains/src/mountain_car/continuous.rs, rsrl_domains/src/hiv.rs */
    RSRL_CONTINUOUS_MOUNTAIN_CAR = 3, RSRL_HIV = 4
} rsrl_domain_t;
/* lfa::basis::{Fourier, Polynomial, TileCoding} (+ .with_bias()) */
typedef enum rsrl_basis { RSRL_FOURIER = 0, RSRL_POLYNOMIAL = 1, RSRL_TILE_CODING = 2 } rsrl_basis_t;
/* rsrl/src/control/td/{q_learning,sarsa,expected_sarsa,sarsa_lambda,q_lambda,pal}.rs, rsrl/src/prediction/td/{td,td_lambda}.rs */
typedef enum rsrl_algo {
    RSRL_QLEARNING = 0, RSRL_SARSA = 1, RSRL_EXPECTED_SARSA = 2,
    RSRL_SARSA_LAMBDA = 3, RSRL_Q_LAMBDA = 4, RSRL_TD_LAMBDA = 5, RSRL_TD0 = 6,
    RSRL_PAL = 7, /* persistent advantage learning, control/td/pal.rs:35-59 (update error = alpha * residual) */
    /* Two weight tables (rsrl_engine_{get,set}_aux_weights reaches the second one):
     * GreedyGQ (control/td/greedy_gq.rs:73-141; examples/greedy_gq.rs): fa_q = the weights (SGD `lr`), fa_td = the aux weights
     *   (SGD `alpha`); three updates per transition: fa_q(s, a) += td_error, fa_q(s', argmax Q(s')) += -gamma * td_est, fa_td(s, a) += td_error - td_est.
     * A2C (examples/a2c.rs:24-66 = control/ac.rs:100-114 + policies/softmax.rs:113-129,146-160 with a SARSA critic): the critic's Q = the
     *   weights (SGD `lr`), the Gibbs policy's own LFA = the aux weights; actions are sampled from softmax(theta^T phi(s) / tau) (tau in
     *   `epsilon`, policy must be RSRL_SOFTMAX); advantage = Q(s)[a] - sum_i Q(s)_i pi_i after the critic update; theta += (alpha * advantage) * grad_log pi(a | s). */
    RSRL_GREEDY_GQ = 8, RSRL_A2C = 9
} rsrl_algo_t;
/* rsrl/src/policies/{greedy,epsilon_greedy,random,softmax}.rs.  RSRL_SOFTMAX (= Gibbs, softmax.rs:38): probabilities
 * softmax_stable(Q(s), tau) (softmax.rs:15-36), sample by inverse CDF (policies/mod.rs:46-61), mode = argmax_first of the
 * probabilities (softmax.rs:141); its temperature tau travels in the `epsilon` field / argument. */
typedef enum rsrl_policy { RSRL_GREEDY = 0, RSRL_EPSILON_GREEDY = 1, RSRL_RANDOM = 2, RSRL_SOFTMAX = 3 } rsrl_policy_t;
/* rsrl/src/traces.rs:196-240  Accumulate / Saturate ("replacing") / Dutch */
typedef enum rsrl_trace_rule { RSRL_TRACE_ACCUMULATE = 0, RSRL_TRACE_REPLACE = 1, RSRL_TRACE_DUTCH = 2 } rsrl_trace_rule_t;
/* SHARED: one agent learns from N envs (W replicated per GPU, dW summed);
 * PER_ENV: N independent reference agents, one W each. */
typedef enum rsrl_weight_mode { RSRL_SHARED = 0, RSRL_PER_ENV = 1 } rsrl_weight_mode_t;
/* SHARED mode only: W += lr * sum_i(...) (SUM; N = 1 is exactly the reference) or lr/N_global * sum (MEAN) */
typedef enum rsrl_update_scale { RSRL_SCALE_SUM = 0, RSRL_SCALE_MEAN = 1 } rsrl_update_scale_t;
/* arithmetic type of features / Q / weights / traces.  Physics is always f64. */
typedef enum rsrl_dtype { RSRL_F32 = 0, RSRL_F64 = 1 } rsrl_dtype_t;
/* DEFAULT: Domain::default() (every episode starts from the same state, e.g.
 * mountain_car/discrete.rs:68-70); UNIFORM: each component ~ U[init_lo, init_hi) */
typedef enum rsrl_init_mode { RSRL_INIT_DEFAULT = 0, RSRL_INIT_UNIFORM = 1 } rsrl_init_mode_t;

typedef struct rsrl_config {
    uint32_t struct_size;       /* = sizeof(rsrl_config_t); ABI guard */
    int32_t  domain;            /* rsrl_domain_t */
    int32_t  basis;             /* rsrl_basis_t */
    int32_t  basis_order;       /* Fourier / Polynomial order (examples/q_learning.rs:24 uses 5) */
    int32_t  n_tilings;         /* TileCoding: tilings T */
    int32_t  tiles_per_dim;     /* TileCoding: tiles per dimension per tiling */
    int32_t  memory_size;       /* TileCoding: hashed table rows M (power of two) */
    int32_t  algo;              /* rsrl_algo_t */
    int32_t  policy;            /* rsrl_policy_t (behaviour policy; also SARSA's in-handle policy) */
    int32_t  trace_rule;        /* rsrl_trace_rule_t */
    int32_t  weight_mode;       /* rsrl_weight_mode_t */
    int32_t  update_scale;      /* rsrl_update_scale_t */
    int32_t  dtype;             /* rsrl_dtype_t */
    int32_t  init_mode;         /* rsrl_init_mode_t */
    int32_t  device;            /* CUDA device ordinal */
    int32_t  record_td_error;   /* != 0: keep last TD error per env (Response{error}, q_learning.rs:17-20) */
    int64_t  n_envs;            /* envs owned by this engine (this GPU's shard) */
    int64_t  env_offset;        /* global id of local env 0 */
    int64_t  n_envs_global;     /* envs over all shards (RNG keys + MEAN scale); 0 => n_envs */
    int64_t  max_episode_steps; /* 0 = uncapped like examples/q_learning.rs:40 */
    uint64_t seed;
    double   lr;                /* SGD(lr) of the LFA (examples/q_learning.rs:25) */
    double   alpha;             /* agent step size: ExpectedSARSA (expected_sarsa.rs:64), PAL (pal.rs:49-57), lambda agents */
    double   gamma;
    double   lambda;
    double   epsilon;           /* EpsilonGreedy: epsilon; Softmax: temperature tau (|tau| >= 1e-7, softmax.rs:61-64) */
    double   init_lo[RSRL_MAX_DIM];
    double   init_hi[RSRL_MAX_DIM];
} rsrl_config_t;

typedef struct rsrl_stats {
    int64_t total_steps;        /* env-steps executed by this engine since reset */
    int64_t total_episodes;     /* episodes finished (terminal or capped) */
    int64_t terminal_episodes;  /* ... of which ended in a terminal observation */
    int64_t batch_steps;        /* batched steps t since reset */
    int64_t kernel_launches;    /* CUDA kernels launched by this engine since create */
    int32_t nonfinite;          /* sticky: some env saw a Q vector with no valid maximum */
    int32_t reserved;
} rsrl_stats_t;

typedef struct rsrl_engine rsrl_engine_t;

/* ---- library ---- */
int         rsrl_version(void);                 /* RSRL_ABI_VERSION */
const char* rsrl_last_error(void);              /* thread-local message of the last failing call */
int         rsrl_device_count(void);            /* 0 when no CUDA device is visible */
/* fills *cfg with examples/q_learning.rs:18-32: MountainCar, Fourier(5)+bias, SGD(0.001),
 * gamma 0.9, Greedy, seed 0, 1 env, SHARED/SUM, f32, default start, uncapped */
int         rsrl_config_default(rsrl_config_t* cfg);
/* D, A, F for a config (F = weights_dim().0, A = weights_dim().1; params/mod.rs:116-134) */
int         rsrl_config_dims(const rsrl_config_t* cfg, int32_t* dim, int32_t* n_actions, int64_t* n_features);

/* ---- engine: the fused transition -> handle -> sample loop (examples/q_learning.rs:34-55) ---- */
int rsrl_engine_create(const rsrl_config_t* cfg, rsrl_engine_t** out);
int rsrl_engine_destroy(rsrl_engine_t* e);
/* init_states: N x D f64 or NULL (=> cfg.init_mode). Zeroes W, traces, counters. */
int rsrl_engine_reset(rsrl_engine_t* e, const double* init_states);
/* k fused batched steps, asynchronous on the engine's stream */
int rsrl_engine_step(rsrl_engine_t* e, int64_t k_steps);
int rsrl_engine_sync(rsrl_engine_t* e);         /* RSRL_ENONFINITE if the sticky flag is set */
void* rsrl_engine_stream(rsrl_engine_t* e);     /* cudaStream_t of the engine (for event timing) */

int rsrl_engine_get_states(rsrl_engine_t* e, double* out /* N x D */);
int rsrl_engine_set_states(rsrl_engine_t* e, const double* in /* N x D */);
int rsrl_engine_get_actions(rsrl_engine_t* e, int32_t* out /* N; -1 before the first step */);
int rsrl_engine_get_episode_steps(rsrl_engine_t* e, int32_t* out /* N */);
/* Parameterised::weights() — F x A (SHARED) or N x F x A (PER_ENV), f64 row-major */
int rsrl_engine_get_weights(rsrl_engine_t* e, double* out);
int rsrl_engine_set_weights(rsrl_engine_t* e, const double* in);
/* second weight table of the two-table agents (GreedyGQ: fa_td, A2C: the policy's LFA), same shape as the weights */
int rsrl_engine_get_aux_weights(rsrl_engine_t* e, double* out);
int rsrl_engine_set_aux_weights(rsrl_engine_t* e, const double* in);
/* Trace::buffer (traces.rs:6-12) — N x F x A (Q traces) or N x F (TD(lambda)) */
int rsrl_engine_get_traces(rsrl_engine_t* e, double* out);
int rsrl_engine_set_traces(rsrl_engine_t* e, const double* in);
int rsrl_engine_get_td_errors(rsrl_engine_t* e, double* out /* N; needs record_td_error */);
int rsrl_engine_get_stats(rsrl_engine_t* e, rsrl_stats_t* out);
/* per-env episode bookkeeping: episodes finished, length of the last finished episode,
 * rolling hash h = h * 1000003 + len over all finished episode lengths (bit-exact step-count check) */
int rsrl_engine_get_env_stats(rsrl_engine_t* e, int32_t* n_episodes, int32_t* last_len, uint64_t* len_hash);
int rsrl_engine_set_epsilon(rsrl_engine_t* e, double epsilon);  /* examples/sarsa_lambda.rs:68 decays it per episode (Softmax: tau) */

/* ---- engine, trait-level (un-fused) entry points on the engine's weights ---- */
/* Function<(S,)>::evaluate for VectorLFA (fa/linear.rs:303-311): q_out N x A */
int rsrl_engine_evaluate(rsrl_engine_t* e, int64_t n, const double* states, double* q_out);
/* Policy::sample (policies/mod.rs:65-78) with the engine's policy; `draw` selects the RNG counter */
int rsrl_engine_sample(rsrl_engine_t* e, int64_t n, const double* states, uint64_t draw, int32_t* actions_out);
/* Policy::mode = Enumerable::find_max (greedy.rs:83, core.rs:96-105) */
int rsrl_engine_mode(rsrl_engine_t* e, int64_t n, const double* states, int32_t* actions_out);
/* Handler<&Transition>::handle for the configured agent, batch of n transitions applied as
 * one batched step (all TD errors use the weights before the call); td_out may be NULL */
int rsrl_engine_handle(rsrl_engine_t* e, int64_t n, const double* from_states, const int32_t* actions,
                       const double* rewards, const double* to_states, const uint8_t* terminal,
                       uint64_t draw, double* td_out);

/* ---- Domain::rollout / Trajectory (rsrl_domains/src/lib.rs:334-409,448-479) ----
 * n independent rollouts with the engine's CURRENT weights held fixed (no learning): env i starts from init_states[i] (NULL: the
 * config's start distribution, draw index `draw`) and follows pi = Policy::mode (greedy != 0: greedy.rs:83 find_max / softmax.rs:141;
 * A2C: the policy table) or Policy::sample with RNG draw (draw + step) until a terminal observation or until step_limit - 1 steps
 * are recorded (lib.rs:472-476: `iter.take(sl.saturating_sub(1))`; step_limit <= 0: T_max steps).  Layout = Trajectory{start, steps}:
 *   start_out    n x D          Trajectory::start (state of the first observation)
 *   next_out     n x T_max x D  steps[j].0 (state of the observation after step j)
 *   actions_out  n x T_max      steps[j].1        rewards_out n x T_max  steps[j].2
 *   terminal_out n x T_max      1 where steps[j].0 is Observation::Terminal
 *   len_out      n              Trajectory::n_transitions(); entries j >= len are untouched
 * T_max = max(step_limit - 1, 1) is the caller's row length (the first step is always taken, lib.rs:460-462). */
int rsrl_engine_rollout(rsrl_engine_t* e, int64_t n, const double* init_states, int64_t step_limit, int32_t greedy, uint64_t draw,
                        double* start_out, double* next_out, int32_t* actions_out, double* rewards_out, uint8_t* terminal_out,
                        int32_t* len_out);

/* ---- introspection for the parity tests ---- */
/* Launch shape of the fused loop: out = {persistent kernel in use, its mode, CTAs, cluster size, clusters, threads per CTA,
 * LL lanes per row, reducer lanes per row group, slots per reducer lane, PER_ENV weights in shared memory, world, rank,
 * peers attached, dynamic shared memory bytes, TileCoding engine, large-basis engine (1 + tensor-core bits),
 * counting (fixed-point) exchange in use, 0...}.
 * The order of the fp32 dW sums is a function of these numbers; oracle/oracle32.cpp replays it on the host. */
int rsrl_engine_get_launch_shape(rsrl_engine_t* e, int32_t out[24]);
/* The elementary functions of the device arithmetic (csrc/device.cuh "rsrl math") evaluated on the GPU:
 * fn 0 cos (f64), 1 sin (f64), 2 sin(pi x) (fp32), 3 cos(pi x) (fp32), 4 exp (fp32). */
int rsrl_math_probe(int32_t fn, int64_t n, const double* x, double* out);

/* ---- multi-GPU (one process per GPU; SHARED mode exchanges dW every step) ---- */
int rsrl_comm_unique_id(uint8_t out[128]);                      /* ncclGetUniqueId on rank 0 */
int rsrl_engine_comm_init(rsrl_engine_t* e, const uint8_t id[128], int rank, int world);
/* In-kernel exchange over NVLink peer memory (preferred; world <= 8 GPUs of one box): every rank exports the
 * cudaIpc handle of its dW mailbox, the host gathers the handles (any transport) and attaches them; from then
 * on rsrl_engine_step keeps the persistent kernel and sums dW across GPUs inside it (fp32: order-independent fixed-point
 * sums; f64: in rank order).  Attach before the engine's first step (RSRL_EINVAL afterwards: the exchange tables count
 * arrivals from step 0).  All ranks must call reset / step with the same arguments and in lockstep. */
int rsrl_engine_peer_export(rsrl_engine_t* e, uint8_t handle_out[64]);
int rsrl_engine_peer_attach(rsrl_engine_t* e, const uint8_t* handles /* world x 64 */, int rank, int world);

/* ---- stateless component entry points (host buffers; computed on the GPU) ---- */
/* Domain::{state_space, action_space, default} */
int rsrl_domain_info(int32_t domain, int32_t* dim, int32_t* n_actions, double* lo, double* hi, double* start);
/* Domain::step (lib.rs:434): states updated in place; reward/terminal per env */
int rsrl_domain_step(int32_t domain, int64_t n, double* states_inout, const int32_t* actions,
                     double* rewards_out, uint8_t* terminal_out);
/* Domain::emit(): Observation::Terminal? */
int rsrl_domain_is_terminal(int32_t domain, int64_t n, const double* states, uint8_t* terminal_out);
/* ---- the other ODE / continuous-action domains, as batched Domain::step / Domain::emit (the engine does not drive them) ----
 * ContinuousMountainCar (D = 2, n_actions reported as 0: the action is a double, clipped onto [-1, 1] like Interval::map_onto,
 * continuous.rs:41-48) and HIVTreatment (D = 6 raw state T1, T1S, T2, T2S, V, E; 4 actions = ALL_ACTIONS hiv.rs:35; 1000 RK4
 * sub-steps per step :58-69; observation = clip(-5, log10(state), 8) :131-135; reward from the observation :141-148).
 * `actions` (int32) is read by HIV, `actions_continuous` (f64) by ContinuousMountainCar; obs_out may be NULL. */
int rsrl_domain_ex_info(int32_t domain, int32_t* dim, int32_t* n_actions, double* lo /* D */, double* hi, double* start);
int rsrl_domain_ex_step(int32_t domain, int64_t n, double* states_inout, const int32_t* actions, const double* actions_continuous,
                        double* obs_out, double* rewards_out, uint8_t* terminal_out);
int rsrl_domain_ex_emit(int32_t domain, int64_t n, const double* states, double* obs_out, uint8_t* terminal_out);
const char* rsrl_domain_ex_last_error(void);
/* Basis::project: features_out N x F f64 (dense; TileCoding writes 1.0 at active rows) */
int rsrl_basis_project(const rsrl_config_t* cfg, int64_t n, const double* states, double* features_out);
/* LFA::evaluate: q_out N x A = phi(s)^T W, W is F x A */
int rsrl_lfa_evaluate(const rsrl_config_t* cfg, int64_t n, const double* states, const double* weights, double* q_out);
/* LFA::update_index + SGD for a batch: W[:,a_i] += lr * err_i * phi(s_i), summed over i */
int rsrl_lfa_update_index(const rsrl_config_t* cfg, int64_t n, const double* states, const int32_t* actions,
                          const double* errors, double* weights_inout);
/* Policy::sample on explicit Q vectors (the MockQ tests of greedy.rs:96-168): q is N x A */
int rsrl_policy_sample(int32_t policy, double epsilon, uint64_t seed, uint64_t draw, int64_t env_offset,
                       int64_t n, int32_t n_actions, const double* q, int32_t* actions_out);
/* Function<(S,)>::evaluate of the policy: probabilities N x A (greedy.rs:30-44, epsilon_greedy.rs:38-45) */
int rsrl_policy_probs(int32_t policy, double epsilon, int64_t n, int32_t n_actions, const double* q, double* probs_out);
/* Enumerable::find_max: exact compare, last maximal index wins (core.rs:96-105) */
int rsrl_policy_mode(int64_t n, int32_t n_actions, const double* q, int32_t* actions_out);
/* Trace::update (traces.rs:127-129): z <- rule(z, grad), elementwise over n values */
int rsrl_trace_update(int32_t rule, double gamma, double lambda, double alpha, int64_t n,
                      double* z_inout, const double* grad);
/* Philox4x32-10 block on the device: out[i] = philox(key = seed, counter = (env_i, draw_lo, draw_hi, stream)) */
int rsrl_philox(uint64_t seed, uint64_t draw, uint32_t stream, int64_t env_offset, int64_t n, uint32_t* out /* n x 4 */);

#ifdef __cplusplus
}
#endif
#endif /* RSRL_B200_H */
