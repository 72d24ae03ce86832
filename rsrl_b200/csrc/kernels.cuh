// kernels.cuh — fused per-step kernels of the rsrl hot path (sm_100a).
//
// One batched step t (semantics defined in DESIGN.md, restated on the CPU by
// oracle/rsrl_oracle.c:env_step) for every env i, all with the weights W_t:
//   B  a_t   = policy.sample(Q(s_t; W_t))              examples/q_learning.rs:38,47  greedy.rs:77-81
//      qsa   = Q(s_t; W_t)[a_t]                         control/td/q_learning.rs:53
//   C  s', r = Domain::step(s_t, a_t)                   rsrl_domains/src/lib.rs:436-446
//   D  delta = r + gamma * target(Q(s'; W_t)) - qsa     q_learning.rs:55-62, sarsa.rs:57-68, expected_sarsa.rs:48-59
//   E  dW[:, a_t] += (lr * err) * phi(s_t)              fa/linear.rs:379-391 -> lfa SGD
//   F  episode bookkeeping / auto-reset                 examples/q_learning.rs:37,49-51
// SHARED weights: E is reduced over the CTA in a fixed order into partials[block][F*A]; a
// second tiny kernel sums the partials in block order and applies W_{t+1} = W_t + dW
// (deterministic, bit-reproducible run to run).  PER_ENV weights: E is applied in place.
#pragma once
#include "core.cuh"

namespace rsrl {

template <typename R, int DOM, int BASIS, int P, int AW, int MODE, bool EXT>
__global__ void __launch_bounds__(256) fused_step_kernel(const StepArgs a) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    constexpr int D = Dom::D, F = GB::F, FA = F * AW;
    constexpr bool TDPRED = AW == 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int BLOCK = blockDim.x;  // <= 256; chosen by the host so that the reduce buffers fit in shared memory
    R* Wsm = reinterpret_cast<R*>(smem_raw);  // FA (SHARED)
    R* red = Wsm + ((FA + 3) & ~3);           // SHARED: phiT[F][BLOCK] or zT[FA][BLOCK]
    const int RS = BLOCK + 1;  // padded row stride: reducer lanes (different rows, same column) hit different banks
    R* dc = red + (size_t)(algo_has_trace(a.algo) ? FA : F) * RS;  // SHARED: coef[AW][BLOCK] (trace: [BLOCK])

    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * BLOCK + tid;
    const bool active = i < a.n;
    const int64_t N = a.n;
    const uint64_t g = (uint64_t)(a.env_offset + i);
    const bool traces = algo_has_trace(a.algo);

    if (MODE == RSRL_SHARED) {
        for (int j = tid; j < FA; j += BLOCK) Wsm[j] = static_cast<const R*>(a.W)[j];
        __syncthreads();
    }
    const R* Wg = static_cast<const R*>(a.W);
    auto Wat = [&](int j) -> R { return MODE == RSRL_SHARED ? Wsm[j] : Wg[(int64_t)j * N + i]; };
    auto evalQ = [&](const typename GB::Tab& tab, R* q) {
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = (R)0;
        GB::for_each(tab, [&](int k, R phi) {
#pragma unroll
            for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, Wat(k * AW + c), q[c]);
        });
    };

    typename GB::Tab tab_s, tab_n;
    CoreOut<R> o;
    o.coef = (R)0; o.act = 0; o.reset_before = false; o.terminated = false;

    if (active) {
        double s[D];
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = EXT ? a.ext_from[i * D + d] : a.states[i * D + d];
        auto prep = [](const double* st, typename GB::Tab& tb) { grid_prepare<R, Dom, P, BASIS>(st, tb); };
        env_core<R, DOM, AW, EXT>(a, a.t, g, s, prep, evalQ, evalQ, tab_s, tab_n, false, o, EXT ? a.ext_actions[i] : 0,
                                            EXT ? a.ext_rewards[i] : 0.0, EXT ? a.ext_term[i] != 0 : false,
                                            EXT ? a.ext_to + i * D : nullptr);
        if (a.td) static_cast<R*>(a.td)[i] = o.residual;
        if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
        if (!EXT) {
            a.ep_steps[i] = env_bookkeeping<Dom>(a, a.t, i, g, s, a.ep_steps[i], o.terminated);
            a.actions[i] = o.act;
#pragma unroll
            for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
        }
    }
    const R coef = o.coef;
    const int act = o.act;
    const bool reset_before = o.reset_before, terminated = o.terminated;

    // ---- E: update ----
    if (!traces) {
        if (MODE == RSRL_PER_ENV) {
            if (active) {
                R* Wm = static_cast<R*>(a.W);
                GB::for_each(tab_s, [&](int k, R phi) {
                    const int64_t idx = (int64_t)(k * AW + (TDPRED ? 0 : act)) * N + i;
                    Wm[idx] = O::mul_add_unfused(coef, phi, Wm[idx]);
                });
            }
        } else {
            // CTA reduce in a fixed order: phiT[k][tid], dc[a][tid] -> thread j sums over tid
            if (active) {
                GB::for_each(tab_s, [&](int k, R phi) { red[k * RS + tid] = phi; });
            } else {
#pragma unroll 4
                for (int k = 0; k < F; ++k) red[k * RS + tid] = (R)0;
            }
#pragma unroll
            for (int c = 0; c < AW; ++c) dc[c * BLOCK + tid] = (active && (TDPRED || c == act)) ? coef : (R)0;
            __syncthreads();
            for (int j = tid; j < FA; j += BLOCK) {
                const int k = j / AW, c = j % AW;
                R acc = (R)0;
#pragma unroll 8
                for (int t2 = 0; t2 < BLOCK; ++t2) acc = O::fma(red[k * RS + t2], dc[c * BLOCK + t2], acc);
                static_cast<R*>(a.partials)[(int64_t)blockIdx.x * FA + j] = acc;
            }
        }
    } else {
        // eligibility traces: z <- rule(rate * z + grad), W += coef * z, z.reset() on terminal
        // (traces.rs:127-129,196-240; sarsa_lambda.rs:68-90; q_lambda.rs:68-94; td_lambda.rs:50-70)
        R* Z = static_cast<R*>(a.z);
        const R rate = a.trace_rule == RSRL_TRACE_DUTCH ? (R)(a.gamma * a.lambda * (1.0 - a.alpha)) : (R)(a.gamma * a.lambda);
        if (active) {
            GB::for_each(tab_s, [&](int k, R phi) {
#pragma unroll
                for (int c = 0; c < AW; ++c) {
                    const int j = k * AW + c;
                    const int64_t idx = (int64_t)j * N + i;
                    R zv = reset_before ? (R)0 : Z[idx];
                    const R grad = (TDPRED || c == act) ? phi : (R)0;
                    zv = trace_rule<R>(a.trace_rule, rate, zv, grad);
                    if (MODE == RSRL_PER_ENV) {
                        R* Wm = static_cast<R*>(a.W);
                        Wm[idx] = O::mul_add_unfused(coef, zv, Wm[idx]);
                    } else {
                        red[j * RS + tid] = zv;
                    }
                    Z[idx] = terminated ? (R)0 : zv;
                }
            });
        } else if (MODE == RSRL_SHARED) {
            for (int j = 0; j < FA; ++j) red[j * RS + tid] = (R)0;
        }
        if (MODE == RSRL_SHARED) {
            dc[tid] = active ? coef : (R)0;
            __syncthreads();
            for (int j = tid; j < FA; j += BLOCK) {
                R acc = (R)0;
#pragma unroll 8
                for (int t2 = 0; t2 < BLOCK; ++t2) acc = O::fma(red[j * RS + t2], dc[t2], acc);
                static_cast<R*>(a.partials)[(int64_t)blockIdx.x * FA + j] = acc;
            }
        }
    }
}

// Sums the per-CTA partials in block order (deterministic) and either applies them
// (W += dW, single GPU) or writes dW for the cross-GPU exchange.
template <typename R>
__global__ void reduce_partials_kernel(const R* __restrict__ partials, int n_blocks, int fa, R* __restrict__ W,
                                       R* __restrict__ dW_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= fa) return;
    R a0 = (R)0, a1 = (R)0, a2 = (R)0, a3 = (R)0;
    int b = 0;
    for (; b + 4 <= n_blocks; b += 4) {  // 4 independent chains, fixed association
        a0 += partials[(int64_t)(b + 0) * fa + j];
        a1 += partials[(int64_t)(b + 1) * fa + j];
        a2 += partials[(int64_t)(b + 2) * fa + j];
        a3 += partials[(int64_t)(b + 3) * fa + j];
    }
    for (; b < n_blocks; ++b) a0 += partials[(int64_t)b * fa + j];
    const R g = (a0 + a1) + (a2 + a3);
    if (dW_out) dW_out[j] = g;
    else W[j] += g;
}

template <typename R>
__global__ void add_kernel(R* __restrict__ W, const R* __restrict__ dW, int fa) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < fa) W[j] += dW[j];
}

// ---------------------------------------------------------------------------
// component kernels (stateless entry points + engine evaluate/sample/mode)
// ---------------------------------------------------------------------------
template <int DOM>
__global__ void domain_step_kernel(int64_t n, double* states, const int32_t* actions, double* rewards, uint8_t* terminal) {
    using Dom = Domain<DOM>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[Dom::D];
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) s[d] = states[i * Dom::D + d];
    if (actions) {
        double r; bool term;
        Dom::step(s, actions[i], r, term);
#pragma unroll
        for (int d = 0; d < Dom::D; ++d) states[i * Dom::D + d] = s[d];
        rewards[i] = r;
        terminal[i] = term;
    } else {
        terminal[i] = Dom::is_terminal(s);  // Domain::emit() is Observation::Terminal
    }
}

// mode 0: features N x F (double out); 1: Q = phi^T W (N x AW); 2: policy sample; 3: find_max mode
template <typename R, int DOM, int BASIS, int P, int AW>
__global__ void basis_eval_kernel(int mode, int64_t n, const double* __restrict__ states, const R* __restrict__ W,
                                  int64_t w_env_stride /* 0: shared W; 1: per-env W[FA][N] */, double* __restrict__ out,
                                  int32_t* __restrict__ act_out, PolicyParams pol, uint64_t draw, int64_t env_offset,
                                  Counters* counters) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[Dom::D];
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) s[d] = states[i * Dom::D + d];
    typename GB::Tab tab;
    grid_prepare<R, Dom, P, BASIS>(s, tab);
    if (mode == 0) {
        GB::for_each(tab, [&](int k, R phi) { out[i * GB::F + k] = (double)phi; });
        return;
    }
    R q[AW];
#pragma unroll
    for (int c = 0; c < AW; ++c) q[c] = (R)0;
    GB::for_each(tab, [&](int k, R phi) {
#pragma unroll
        for (int c = 0; c < AW; ++c) {
            const R w = w_env_stride ? W[(int64_t)(k * AW + c) * n + i] : W[k * AW + c];
            q[c] = O::mac(phi, w, q[c]);
        }
    });
    if (mode == 1) {
#pragma unroll
        for (int c = 0; c < AW; ++c) out[i * AW + c] = (double)q[c];
    } else if (mode == 2) {
        bool nf = false;
        act_out[i] = policy_sample<R, AW>(pol, q, (uint64_t)(env_offset + i), draw, STREAM_BEHAVIOUR, nf);
        if (nf) atomicExch(&counters->nonfinite, 1);
    } else {
        R mx;
        (void)mx;
        act_out[i] = policy_mode<R, AW>(pol.policy, (R)pol.tau, q);
    }
}

// policies on explicit Q vectors (MockQ-style tests): mode 0 sample, 1 probs, 2 find_max
template <typename R, int A>
__global__ void policy_kernel(int mode, int64_t n, const double* __restrict__ q_in, PolicyParams pol, double eps,
                              uint64_t draw, int64_t env_offset, int32_t* __restrict__ act_out,
                              double* __restrict__ probs_out, Counters* counters) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R q[A];
#pragma unroll
    for (int c = 0; c < A; ++c) q[c] = (R)q_in[i * A + c];
    if (mode == 0) {
        bool nf = false;
        act_out[i] = policy_sample<R, A>(pol, q, (uint64_t)(env_offset + i), draw, STREAM_BEHAVIOUR, nf);
        if (nf) atomicExch(&counters->nonfinite, 1);
    } else if (mode == 1) {
        R p[A];
        policy_probs<R, A>(pol.policy, (R)eps, q, p);
#pragma unroll
        for (int c = 0; c < A; ++c) probs_out[i * A + c] = (double)p[c];
    } else {
        R mx;
        act_out[i] = find_max<R, A>(q, mx);
    }
}

template <typename R>
__global__ void trace_update_kernel(int rule, double rate, int64_t n, double* __restrict__ z, const double* __restrict__ grad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = (double)trace_rule<R>(rule, (R)rate, (R)z[i], (R)grad[i]);
}

static __global__ void philox_kernel(uint64_t seed, uint64_t draw, uint32_t stream, int64_t env_offset, int64_t n, uint32_t* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = draw4(seed, (uint64_t)(env_offset + i), draw, stream);
    out[i * 4 + 0] = r.x; out[i * 4 + 1] = r.y; out[i * 4 + 2] = r.z; out[i * 4 + 3] = r.w;
}

// init / reset: fresh states for step 0
template <int DOM>
__global__ void init_states_kernel(int64_t n, double* states, int init_mode, const double* lo4, const double* hi4,
                                   uint64_t seed, int64_t env_offset) {
    using Dom = Domain<DOM>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[Dom::D], lo[RSRL_MAX_DIM], hi[RSRL_MAX_DIM];
#pragma unroll
    for (int d = 0; d < RSRL_MAX_DIM; ++d) { lo[d] = lo4[d]; hi[d] = hi4[d]; }
    fresh_state<Dom>(s, init_mode, lo, hi, seed, (uint64_t)(env_offset + i), 0);
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) states[i * Dom::D + d] = s[d];
}

// dtype / layout conversion between the ABI (f64, env-major) and the device (R, [j][N] for per-env tensors)
template <typename R>
__global__ void export_kernel(const R* __restrict__ src, double* __restrict__ dst, int64_t n_env, int64_t fa, int transposed) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over dst (env-major)
    if (idx >= n_env * fa) return;
    const int64_t e = idx / fa, j = idx % fa;
    dst[idx] = (double)(transposed ? src[j * n_env + e] : src[idx]);
}
template <typename R>
__global__ void import_kernel(const double* __restrict__ src, R* __restrict__ dst, int64_t n_env, int64_t fa, int transposed) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_env * fa) return;
    const int64_t e = idx / fa, j = idx % fa;
    if (transposed) dst[j * n_env + e] = (R)src[idx];
    else dst[idx] = (R)src[idx];
}

}  // namespace rsrl
