// core.cuh — one env, one batched step: phases B-D (behaviour action, Domain::transition, TD error under W_t) and
// phase F (episode bookkeeping).  Host-compilable (hostdev.h): oracle/oracle32.cpp builds this same code with g++.
#pragma once
#include "device.cuh"

namespace rsrl {

struct Counters {
    unsigned long long episodes;
    unsigned long long terminal_episodes;
    int nonfinite;
    int pad;
};

__host__ __device__ __forceinline__ void counter_add(unsigned long long* p) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, 1ull);
#else
    *p += 1ull;
#endif
}

struct StepArgs {
    // per-env state (HBM): states are N x D f64 row-major == the ABI layout
    double* states;
    int32_t* actions;
    int32_t* ep_steps;
    int32_t* n_ep;
    int32_t* last_len;
    unsigned long long* len_hash;
    void* td;  // R[N] or nullptr
    // parameters
    void* W;         // SHARED: R[F*AW] (index k*AW + a == Parameterised::weights_view row-major)
                     // PER_ENV: R[F*AW][N] (env fastest: coalesced)
    void* z;         // traces R[F*AW][N] or nullptr
    void* partials;  // SHARED: R[grid][F*AW]
    Counters* counters;
    long long* phase_prof;  // optional [grid][8] cycle counters (RSRL_B200_PHASE_PROFILE=1), else nullptr
    // external transitions (Handler::handle entry point); nullptr for the fused loop
    const double* ext_from;
    const int32_t* ext_actions;
    const double* ext_rewards;
    const double* ext_to;
    const uint8_t* ext_term;
    int64_t n;
    int64_t env_offset;
    uint64_t t;  // batched step index == RNG draw counter
    int64_t max_ep;
    int algo, trace_rule, init_mode, pad0;
    PolicyParams pol;
    double gamma, lr_scaled /* lr / scale */, alpha, inv_scale /* 1 / scale */, lambda, epsilon;
    double init_lo[RSRL_MAX_DIM], init_hi[RSRL_MAX_DIM];
};

__host__ __device__ constexpr bool algo_has_trace(int algo) {
    return algo == RSRL_SARSA_LAMBDA || algo == RSRL_Q_LAMBDA || algo == RSRL_TD_LAMBDA;
}

// Phases B-D for one env: behaviour action, Domain::transition, TD error under W_t.
template <typename R>
struct CoreOut {
    R coef;       // scaled error multiplying phi(s) (or the trace) in the update
    R residual;   // TD error (Response{error})
    int act;
    bool reset_before, terminated, nonfinite;
};

// evalS evaluates Q at the from-state (it may also record phi(s) rows for the update), evalN at s'.
// have_tab_s: tab_s already holds the tables of s (carried over from the previous step's s').
// tab_n returns the tables of s' (valid unless the transition was terminal).
// step_pre (optional): Dom::step_pre(s) computed by the caller ahead of time (persistent.cuh: in the shadow of the grid exchange).
// prep(state, tab) builds the basis tables of a state (Fourier/Polynomial grid tables or tile rows).
template <typename R, int DOM, int AW, bool EXT, class Tab, class Prep, class EvalS, class EvalN>
__host__ __device__ __forceinline__ void env_core(const StepArgs& a, uint64_t t, uint64_t g, double* s, Prep prep, EvalS evalS, EvalN evalN,
                                         Tab& tab_s, Tab& tab_n, bool have_tab_s, CoreOut<R>& o, int ext_act, double ext_reward,
                                         bool ext_term, const double* ext_to, const double* step_pre = nullptr) {
    using Dom = Domain<DOM>;
    constexpr int D = Dom::D;
    constexpr bool TDPRED = AW == 1;  // TD(0)/TD(lambda) state-value prediction: W is F x 1
    o.nonfinite = false;
    o.reset_before = false;

    if (!have_tab_s) prep(s, tab_s);

    // ---- B: behaviour action and Q(s_t, a_t) under W_t ----
    R q[AW];
    evalS(tab_s, q);
    if (EXT) {
        o.act = ext_act;
    } else if (TDPRED) {
        PolicyParams rp = a.pol;
        rp.policy = RSRL_RANDOM;
        o.act = policy_sample<R, Dom::A>(rp, q, g, t, STREAM_BEHAVIOUR, o.nonfinite);
    } else {
        o.act = policy_sample<R, AW>(a.pol, q, g, t, STREAM_BEHAVIOUR, o.nonfinite);
    }
    R qsa = q[0];
    if (!TDPRED) {
#pragma unroll
        for (int c = 0; c < AW; ++c) if (c == o.act) qsa = q[c];
        if (a.algo == RSRL_Q_LAMBDA) o.reset_before = o.act != argmax_first<R, AW>(q);  // q_lambda.rs:68
    }

    // ---- C: Domain::transition ----
    double reward;
    if (EXT) {
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = ext_to[d];
        reward = ext_reward;
        o.terminated = ext_term;
    } else {
        if (Dom::kHasPre && step_pre) Dom::step_post(s, o.act, *step_pre, reward, o.terminated);  // Dom::step_pre(s) taken ahead of time
        else Dom::step(s, o.act, reward, o.terminated);
    }

    // ---- D: TD error with W_t ----
    if (o.terminated) {
        o.residual = (R)reward - qsa;
    } else {
        prep(s, tab_n);
        R nq[AW];
        evalN(tab_n, nq);
        R target;
        if (TDPRED) {
            target = nq[0];
        } else if (a.algo == RSRL_QLEARNING || a.algo == RSRL_Q_LAMBDA) {
            find_max<R, AW>(nq, target);                                                          // q_learning.rs:59
        } else if (a.algo == RSRL_SARSA || a.algo == RSRL_SARSA_LAMBDA) {
            const int na = policy_sample<R, AW>(a.pol, nq, g, t, STREAM_TARGET, o.nonfinite);     // sarsa.rs:61
            target = nq[0];
#pragma unroll
            for (int c = 0; c < AW; ++c) if (c == na) target = nq[c];
        } else if (a.algo == RSRL_PAL) {                                                          // pal.rs:44-52 (literal: nqs[a_star])
            const int a_star = argmax_first<R, AW>(q), na_star = argmax_first<R, AW>(nq);
            R nq_astar = nq[0], q_astar = q[0], nq_nastar = nq[0], nq_act = nq[0];
#pragma unroll
            for (int c = 0; c < AW; ++c) {
                if (c == a_star) { nq_astar = nq[c]; q_astar = q[c]; }
                if (c == na_star) nq_nastar = nq[c];
                if (c == o.act) nq_act = nq[c];
            }
            const R td_error = (R)reward + (R)a.gamma * nq_astar - qsa;
            const R al_error = td_error - (R)a.alpha * (q_astar - qsa);
            const R alt = td_error - (R)a.alpha * (nq_nastar - nq_act);
            o.residual = (al_error > alt || alt != alt) ? al_error : alt;                                // f64::max (pal.rs:52)
            target = (R)0;
        } else {                                                                                  // expected_sarsa.rs:52-56
            R p[AW];
            policy_probs<R, AW>(a.pol.policy, (R)a.epsilon, nq, p);
            target = (R)0;
#pragma unroll
            for (int c = 0; c < AW; ++c) target = target + nq[c] * p[c];
        }
        if (a.algo != RSRL_PAL) o.residual = (R)reward + (R)a.gamma * target - qsa;
    }
    if (a.algo == RSRL_SARSA_LAMBDA || a.algo == RSRL_Q_LAMBDA) o.coef = (R)(a.alpha * a.inv_scale) * o.residual;  // bypasses SGD lr
    else if (a.algo == RSRL_TD_LAMBDA) o.coef = (R)a.inv_scale * o.residual;                                       // td_lambda.rs:56-59
    else if (a.algo == RSRL_EXPECTED_SARSA || a.algo == RSRL_PAL) o.coef = (R)a.lr_scaled * ((R)a.alpha * o.residual);  // expected_sarsa.rs:64, pal.rs:57
    else o.coef = (R)a.lr_scaled * o.residual;
}

// Phase F: episode bookkeeping / auto-reset (examples/q_learning.rs:37,49-51). Returns the new ep counter.
template <class Dom>
__host__ __device__ __forceinline__ int env_bookkeeping(const StepArgs& a, uint64_t t, int64_t i, uint64_t g, double* s, int ep,
                                               bool terminated, bool* was_reset = nullptr) {
    ep += 1;
    const bool ended = terminated || (a.max_ep > 0 && ep >= a.max_ep);
    if (was_reset) *was_reset = ended;
    if (ended) {
        a.n_ep[i] += 1;
        a.last_len[i] = ep;
        a.len_hash[i] = a.len_hash[i] * 1000003ull + (unsigned long long)ep;
        counter_add(&a.counters->episodes);
        if (terminated) counter_add(&a.counters->terminal_episodes);
        ep = 0;
        fresh_state<Dom>(s, a.init_mode, a.init_lo, a.init_hi, a.pol.seed, g, t + 1);
    }
    return ep;
}

}  // namespace rsrl
