// twotable.cuh — the agents with TWO weight tables, one batched step per launch (sm_100a):
//
//   GreedyGQ  rsrl/src/control/td/greedy_gq.rs:73-141 (examples/greedy_gq.rs)     table 0 = fa_q, table 1 = fa_td
//       qsa = Q(s)[a]; td_est = V(s)[a]; (na, qnsna) = find_max Q(s'); td_error = r + gamma qnsna - qsa (terminal: r - qsa)
//       fa_q(s, a) += td_error;  fa_q(s', na) += -gamma td_est (non-terminal);  fa_td(s, a) += td_error - td_est
//   A2C       rsrl/examples/a2c.rs:24-66 = control/td/sarsa.rs:53-75 (critic) + control/ac.rs:100-114 + policies/softmax.rs:113-129,146-160
//       table 0 = the critic's Q, table 1 = the Gibbs policy's own LFA theta.  a ~ softmax(theta^T phi(s) / tau);
//       critic: SARSA with na ~ policy(s'); advantage = Q'(s)[a] - sum_i Q'(s)_i pi(s)_i on the UPDATED Q (a2c.rs:55-56 runs eval.handle
//       first); theta[:, c] += (alpha advantage) * (1[c == a] - pi_c) * phi(s)   (grad_log pi, the optimiser is bypassed)
//
// Batched semantics as everywhere: all envs read the tables of step t; SHARED weights sum the
// contributions (CTA partials -> reduce_partials_kernel), PER_ENV weights apply them in place in the reference's order.
// In SHARED mode the A2C critic closure sees W_t plus the env's own critic update — for one env that is the reference's
// in-place update.  CPU statement: oracle/rsrl_oracle.c greedy_gq_handle / a2c_handle.
#pragma once
#include "kernels.cuh"

namespace rsrl {

template <typename R, int DOM, int BASIS, int P, int MODE, bool EXT>
__global__ void __launch_bounds__(256) two_table_step_kernel(const StepArgs a) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    constexpr int D = Dom::D, F = GB::F, AW = Dom::A, FA = F * AW;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int BLOCK = blockDim.x, RS = BLOCK + 1;
    R* Wsm = reinterpret_cast<R*>(smem_raw);       // [2][FA] (SHARED)
    R* red_s = Wsm + ((2 * FA + 3) & ~3);          // [F][RS] phi(s)
    R* red_n = red_s + (size_t)F * RS;             // [F][RS] phi(s') (GreedyGQ)
    R* dc = red_n + (size_t)F * RS;                // [3][AW][BLOCK]: table 0 from s, table 0 from s', table 1 from s

    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * BLOCK + tid;
    const int64_t N = a.n;
    const bool active = i < N;
    const uint64_t g = (uint64_t)(a.env_offset + i);
    const bool gq = a.algo == RSRL_GREEDY_GQ;

    if (MODE == RSRL_SHARED) {
        for (int j = tid; j < 2 * FA; j += BLOCK) Wsm[j] = static_cast<const R*>(a.W)[j];
        __syncthreads();
    }
    R* Wg = static_cast<R*>(a.W);
    // table t (0 / 1), flat index j = k * AW + c
    auto Wat = [&](int t, int j) -> R { return MODE == RSRL_SHARED ? Wsm[t * FA + j] : Wg[((int64_t)t * FA + j) * N + i]; };
    auto evalT = [&](int t, const typename GB::Tab& tab, R* q) {
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = (R)0;
        GB::for_each(tab, [&](int k, R phi) {
#pragma unroll
            for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, Wat(t, k * AW + c), q[c]);
        });
    };

    typename GB::Tab tab_s, tab_n;
    R c1 = (R)0, c2 = (R)0, err3 = (R)0, sf[AW];
    int act = 0, na = 0;
    bool terminated = false;
#pragma unroll
    for (int c = 0; c < AW; ++c) sf[c] = (R)0;

    if (active) {
        double s[D];
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = EXT ? a.ext_from[i * D + d] : a.states[i * D + d];
        grid_prepare<R, Dom, P, BASIS>(s, tab_s);
        R qs[AW], xs[AW];
        evalT(0, tab_s, qs);
        evalT(1, tab_s, xs);
        bool nonfinite = false;
        if (EXT) act = a.ext_actions[i];
        else act = policy_sample<R, AW>(a.pol, gq ? qs : xs, g, a.t, STREAM_BEHAVIOUR, nonfinite);  // GQ: behaviour policy on Q; A2C: Gibbs on theta
        R qsa = qs[0], xsa = xs[0];
#pragma unroll
        for (int c = 0; c < AW; ++c) if (c == act) { qsa = qs[c]; xsa = xs[c]; }

        double reward;
        if (EXT) {
#pragma unroll
            for (int d = 0; d < D; ++d) s[d] = a.ext_to[i * D + d];
            reward = a.ext_rewards[i];
            terminated = a.ext_term[i] != 0;
        } else {
            Dom::step(s, act, reward, terminated);
        }
        R residual;
        if (terminated) {
            residual = (R)reward - qsa;                                                  // greedy_gq.rs:81, sarsa.rs:58
        } else {
            grid_prepare<R, Dom, P, BASIS>(s, tab_n);
            R nq[AW];
            evalT(0, tab_n, nq);
            R target;
            if (gq) {
                na = find_max<R, AW>(nq, target);                                        // greedy_gq.rs:107
            } else {
                R nh[AW];
                evalT(1, tab_n, nh);
                na = policy_sample<R, AW>(a.pol, nh, g, a.t, STREAM_TARGET, nonfinite);  // sarsa.rs:61 with the Gibbs policy
                target = nq[0];
#pragma unroll
                for (int c = 0; c < AW; ++c) if (c == na) target = nq[c];
            }
            residual = (R)reward + (R)a.gamma * target - qsa;                            // greedy_gq.rs:109, sarsa.rs:62-64
        }
        c1 = (R)a.lr_scaled * residual;
        if (gq) {
            const R td_est = xsa;                                                        // greedy_gq.rs:78
            c2 = terminated ? (R)0 : (R)a.lr_scaled * (-(R)a.gamma * td_est);            // :117-124
            err3 = (R)(a.alpha * a.inv_scale) * (residual - td_est);                     // :127-133 (fa_td's SGD lr travels in `alpha`)
        } else {
            // critic closure a2c.rs:40-45 on the updated Q: only column `act` of Q(s) changed
            R qa = (R)0;
            GB::for_each(tab_s, [&](int k, R phi) { qa = O::mac(phi, O::mul_add_unfused(c1, phi, Wat(0, k * AW + act)), qa); });
            R ps[AW];
            softmax_probs<R, AW>((R)a.pol.tau, xs, ps);
            R ev = (R)0;
#pragma unroll
            for (int c = 0; c < AW; ++c) ev = ev + (c == act ? qa : qs[c]) * ps[c];
            err3 = (R)(a.alpha * a.inv_scale) * (qa - ev);                               // ac.rs:109-113: alpha * critic.target(t)
#pragma unroll
            for (int c = 0; c < AW; ++c) sf[c] = c == act ? ps[c] - (R)1 : ps[c];        // softmax.rs:117-118
        }
        if (a.td) static_cast<R*>(a.td)[i] = residual;
        if (nonfinite) atomicExch(&a.counters->nonfinite, 1);
        if (!EXT) {
            a.ep_steps[i] = env_bookkeeping<Dom>(a, a.t, i, g, s, a.ep_steps[i], terminated);
            a.actions[i] = act;
#pragma unroll
            for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
        }
    }

    if (MODE == RSRL_PER_ENV) {
        if (!active) return;
        // the reference's order: fa_q(s, a), then fa_q(s', na), then fa_td(s, a) / the policy
        GB::for_each(tab_s, [&](int k, R phi) {
            const int64_t idx = (int64_t)(k * AW + act) * N + i;
            Wg[idx] = O::mul_add_unfused(c1, phi, Wg[idx]);
        });
        if (gq) {
            if (!terminated)
                GB::for_each(tab_n, [&](int k, R phi) {
                    const int64_t idx = (int64_t)(k * AW + na) * N + i;
                    Wg[idx] = O::mul_add_unfused(c2, phi, Wg[idx]);
                });
            GB::for_each(tab_s, [&](int k, R phi) {
                const int64_t idx = ((int64_t)FA + k * AW + act) * N + i;
                Wg[idx] = O::mul_add_unfused(err3, phi, Wg[idx]);
            });
        } else {
            GB::for_each(tab_s, [&](int k, R phi) {
#pragma unroll
                for (int c = 0; c < AW; ++c) {  // theta += error * ((-sf[c]) * phi)  (softmax.rs:122-124, fa/linear.rs:193-195)
                    const int64_t idx = ((int64_t)FA + k * AW + c) * N + i;
                    Wg[idx] = O::mul_add_unfused(err3, (-sf[c]) * phi, Wg[idx]);
                }
            });
        }
        return;
    }

    // SHARED: CTA partials in a fixed order
    if (active) {
        GB::for_each(tab_s, [&](int k, R phi) { red_s[k * RS + tid] = phi; });
        if (gq && !terminated) GB::for_each(tab_n, [&](int k, R phi) { red_n[k * RS + tid] = phi; });
        else if (gq) { for (int k = 0; k < F; ++k) red_n[k * RS + tid] = (R)0; }
    } else {
        for (int k = 0; k < F; ++k) { red_s[k * RS + tid] = (R)0; if (gq) red_n[k * RS + tid] = (R)0; }
    }
#pragma unroll
    for (int c = 0; c < AW; ++c) {
        dc[(0 * AW + c) * BLOCK + tid] = (active && c == act) ? c1 : (R)0;
        dc[(1 * AW + c) * BLOCK + tid] = (active && gq && c == na) ? c2 : (R)0;
        dc[(2 * AW + c) * BLOCK + tid] = !active ? (R)0 : gq ? (c == act ? err3 : (R)0) : err3 * (-sf[c]);
    }
    __syncthreads();
    for (int j = tid; j < FA; j += BLOCK) {
        const int k = j / AW, c = j % AW;
        R acc0 = (R)0, acc1 = (R)0;
#pragma unroll 4
        for (int t2 = 0; t2 < BLOCK; ++t2) {
            const R ph = red_s[k * RS + t2];
            acc0 = O::fma(ph, dc[(0 * AW + c) * BLOCK + t2], acc0);
            acc1 = O::fma(ph, dc[(2 * AW + c) * BLOCK + t2], acc1);
        }
        if (gq) {
#pragma unroll 4
            for (int t2 = 0; t2 < BLOCK; ++t2) acc0 = O::fma(red_n[k * RS + t2], dc[(1 * AW + c) * BLOCK + t2], acc0);
        }
        static_cast<R*>(a.partials)[(int64_t)blockIdx.x * 2 * FA + j] = acc0;
        static_cast<R*>(a.partials)[(int64_t)blockIdx.x * 2 * FA + FA + j] = acc1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Domain::rollout (rsrl_domains/src/lib.rs:448-479): one thread per env, the weights held fixed, steps recorded in the
// Trajectory{start, steps} layout of the ABI (include/rsrl_b200.h: rsrl_engine_rollout).
// ---------------------------------------------------------------------------------------------------------------
struct RolloutArgs {
    int64_t n, env_offset, t_max /* row length */, take /* steps recorded at most; < 0: t_max */;
    const double* init;      // n x D or nullptr
    const void* W;           // the table the policy reads (A2C: the policy's own)
    int64_t w_env_stride;    // 0: shared; 1: per-env [FA][N]
    int greedy, init_mode;
    uint64_t draw;
    PolicyParams pol;
    double init_lo[RSRL_MAX_DIM], init_hi[RSRL_MAX_DIM];
    double* start_out; double* next_out; int32_t* actions_out; double* rewards_out; uint8_t* terminal_out; int32_t* len_out;
    Counters* counters;
};

template <typename R, int DOM, int BASIS, int P>
__global__ void __launch_bounds__(128) rollout_kernel(const RolloutArgs ra) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    constexpr int D = Dom::D, AW = Dom::A;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ra.n) return;
    const uint64_t g = (uint64_t)(ra.env_offset + i);
    const R* W = static_cast<const R*>(ra.W);
    double s[D];
    if (ra.init) {
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = ra.init[i * D + d];
    } else {
        fresh_state<Dom>(s, ra.init_mode, ra.init_lo, ra.init_hi, ra.pol.seed, g, ra.draw);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) ra.start_out[i * D + d] = s[d];   // let start = self.emit()
    const int64_t take = ra.take < 0 ? ra.t_max : ra.take;
    int64_t n = 0;
    bool terminated = false, nonfinite = false;
    for (int64_t j = 0;; ++j) {
        if (j > 0 && (terminated || j >= take)) break;   // Terminal => None; .take(sl - 1)
        typename GB::Tab tab;
        grid_prepare<R, Dom, P, BASIS>(s, tab);
        R q[AW];
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = (R)0;
        GB::for_each(tab, [&](int k, R phi) {
#pragma unroll
            for (int c = 0; c < AW; ++c) {
                const R w = ra.w_env_stride ? W[(int64_t)(k * AW + c) * ra.n + i] : W[k * AW + c];
                q[c] = O::mac(phi, w, q[c]);
            }
        });
        const int act = ra.greedy ? policy_mode<R, AW>(ra.pol.policy, (R)ra.pol.tau, q)
                                  : policy_sample<R, AW>(ra.pol, q, g, ra.draw + (uint64_t)j, STREAM_BEHAVIOUR, nonfinite);
        double reward;
        Dom::step(s, act, reward, terminated);
        if (j >= take) break;                              // step_limit == 1: the first step is executed but not recorded
        const int64_t o = i * ra.t_max + j;
#pragma unroll
        for (int d = 0; d < D; ++d) ra.next_out[o * D + d] = s[d];
        ra.actions_out[o] = act;
        ra.rewards_out[o] = reward;
        ra.terminal_out[o] = terminated ? 1 : 0;
        n = j + 1;
    }
    ra.len_out[i] = (int32_t)n;
    if (nonfinite) atomicExch(&ra.counters->nonfinite, 1);
}

}  // namespace rsrl
