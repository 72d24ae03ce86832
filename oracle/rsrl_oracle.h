/*
 * rsrl_oracle.h — CPU restatement of the rsrl hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 * The product (rsrl_b200/) never includes, links or calls anything in oracle/.
 *
 * PARITY PINNING (see DESIGN.md "Oracle"):
 *   pinned by the reference's own tests   : CartPole RK4 dynamics (cart_pole.rs:144-183),
 *       MountainCar terminal predicate (mountain_car/discrete.rs:109-137), initial
 *       observations, Greedy / EpsilonGreedy argmax + probabilities (greedy.rs:96-168,
 *       epsilon_greedy.rs:116-145), trace decay (traces.rs:112-126,135-148).
 *   PARITY UNPINNED (no reference test, crate source not under /root/reference, no Rust
 *       toolchain here): lfa 0.15 Fourier / Polynomial / TileCoding / LFA / SGD, the TD
 *       control handlers (never tested in the reference), Acrobot dynamics, rand 0.7 streams.
 *       These follow the reference source text (handlers, Acrobot) or the published
 *       algorithm of the crate (Konidaris Fourier basis, plain SGD); TileCoding and
 *       Polynomial input conventions are project-defined.
 */
#ifndef RSRL_ORACLE_H
#define RSRL_ORACLE_H

#include <stdint.h>
#include "../include/rsrl_b200.h" /* rsrl_config_t + enums only (interface, not implementation) */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- domains (rsrl_domains) ---- */
int  orc_domain_dim(int domain);
int  orc_domain_n_actions(int domain);
void orc_domain_limits(int domain, double* lo, double* hi);   /* Domain::state_space() */
void orc_domain_default(int domain, double* s);               /* Default::default() */
int  orc_domain_is_terminal(int domain, const double* s);     /* emit() is Observation::Terminal */
void orc_domain_step(int domain, double* s, int action, double* reward, int* terminal); /* Domain::step */

/* ContinuousMountainCar (continuous.rs) / HIVTreatment (hiv.rs): raw state in, observation out */
int  orc_domain_ex_dim(int domain);
void orc_domain_ex_default(int domain, double* s);
void orc_domain_ex_emit(int domain, const double* s, double* obs, int* terminal);
void orc_domain_ex_step(int domain, double* s, double action, double* obs, double* reward, int* terminal);

/* ---- bases (lfa crate, restated from its published algorithm) ---- */
int64_t orc_basis_n_features(const rsrl_config_t* cfg);
void orc_fourier_coefficients(int order, int dim, double* coef /* (F-1) x dim */);
void orc_basis_project(const rsrl_config_t* cfg, const double* s, double* phi /* F, dense */);
int  orc_tile_indices(const rsrl_config_t* cfg, const double* s, int32_t* idx /* T */); /* returns #unique */

/* ---- LFA (fa/linear.rs wrappers over lfa::LFA) ---- */
void orc_lfa_evaluate(const rsrl_config_t* cfg, const double* W, int A, const double* s, double* q);
void orc_lfa_update_index(const rsrl_config_t* cfg, double* W, int A, const double* s, int a, double lr_err);

/* ---- argmax family (utils.rs, core.rs) ---- */
int  orc_argmaxima(const double* v, int n, int* ixs, double* max_out);  /* utils.rs:6-21; returns count */
int  orc_find_max(const double* v, int n, double* max_out);             /* core.rs:96-105 */
int  orc_argmax_first(const double* v, int n, double* max_out);         /* utils.rs:23-34 */

/* ---- RNG: Philox4x32-10 (project-defined stream; the reference's rand 0.7 stream is not reproduced) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_draw(uint64_t seed, uint64_t env, uint64_t draw, uint32_t stream, uint32_t out[4]);

/* ---- policies (policies/{greedy,epsilon_greedy,random}.rs) ---- */
void orc_policy_probs(int policy, double epsilon, const double* q, int n, double* p);
int  orc_policy_sample(int policy, double epsilon, const double* q, int n, const uint32_t rnd[4], int* nonfinite);

/* ---- traces (traces.rs) ---- */
void orc_trace_update(int rule, double gamma, double lambda, double alpha, int64_t n, double* z, const double* grad);

/* ---- batched engine: the N-env restatement of examples/q_learning.rs:34-55 ---- */
typedef struct orc_engine orc_engine_t;
orc_engine_t* orc_engine_create(const rsrl_config_t* cfg);
void orc_engine_destroy(orc_engine_t* e);
void orc_engine_reset(orc_engine_t* e, const double* init_states);
void orc_engine_step(orc_engine_t* e, int64_t k);
void orc_engine_get_states(orc_engine_t* e, double* out);
void orc_engine_set_states(orc_engine_t* e, const double* in);
void orc_engine_get_actions(orc_engine_t* e, int32_t* out);
void orc_engine_get_episode_steps(orc_engine_t* e, int32_t* out);
void orc_engine_get_weights(orc_engine_t* e, double* out);
void orc_engine_set_weights(orc_engine_t* e, const double* in);
void orc_engine_get_aux_weights(orc_engine_t* e, double* out); /* GreedyGQ fa_td / A2C policy LFA */
void orc_engine_set_aux_weights(orc_engine_t* e, const double* in);
int64_t orc_engine_rollout(orc_engine_t* e, int64_t i, const double* init_state, int64_t step_limit, int greedy, uint64_t draw,
                           double* start_out, double* next_out, int32_t* actions_out, double* rewards_out, uint8_t* terminal_out);
void orc_engine_get_traces(orc_engine_t* e, double* out);
void orc_engine_set_traces(orc_engine_t* e, const double* in);
void orc_engine_get_td_errors(orc_engine_t* e, double* out);
void orc_engine_get_stats(orc_engine_t* e, rsrl_stats_t* out);
void orc_engine_get_env_stats(orc_engine_t* e, int32_t* n_episodes, int32_t* last_len, uint64_t* len_hash);
void orc_engine_set_epsilon(orc_engine_t* e, double eps);
/* min over envs and steps of the gap between the best and second-best Q used in an action choice
 * since reset — tells a test how close the run came to a tie (tolerance-limited parity). */
double orc_engine_min_gap(orc_engine_t* e);
/* exchange hook for the world_size>1 tests: the SHARED-mode step is split in two halves */
void orc_engine_step_local(orc_engine_t* e, double* dW_out);       /* phase 1: returns this shard's dW */
void orc_engine_step_apply(orc_engine_t* e, const double* dW_sum); /* phase 2: W += dW_sum, bookkeeping */
/* trait-level handle() on explicit transitions (Handler<&Transition>::handle) */
void orc_engine_handle(orc_engine_t* e, int64_t n, const double* from, const int32_t* actions, const double* rewards,
                       const double* to, const uint8_t* terminal, uint64_t draw, double* td_out);

/* ---- CPU baseline: `threads` independent reference-shaped single-env agents, `steps` env-steps each;
 * returns wall seconds of the stepping alone (engines are created and threads started before the clock); *out_steps = env-steps executed */
double orc_baseline_run(const rsrl_config_t* cfg, int threads, int64_t envs_per_thread, int64_t steps, int64_t* out_steps);
/* the same agents as a fused scalar loop: one projection per env-step, no heap allocation (the best simple CPU implementation;
 * Q-learning / SARSA / ExpectedSARSA).  Reported next to the reference-shaped number, BASELINE.md section 3. */
double orc_baseline_run_fused(const rsrl_config_t* cfg, int threads, int64_t envs_per_thread, int64_t steps, int64_t* out_steps);
/* one env on one core (examples/q_learning.rs itself); ep_lens receives the first n_ep_lens episode lengths (-1: not reached) */
double orc_baseline_run_single(const rsrl_config_t* cfg, int64_t steps, int32_t* ep_lens, int n_ep_lens, int64_t* out_steps);

#ifdef __cplusplus
}
#endif
#endif
