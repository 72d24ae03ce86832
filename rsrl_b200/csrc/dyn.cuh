// dyn.cuh — tensor-grid bases of ANY order (run-time P), the coverage path behind the specialised kernels.
//
// `lfa::basis::{Fourier, Polynomial}` take any order; the fused kernels (kernels.cuh, persistent.cuh, fourier4.cuh) are
// instantiated for the orders the reference's examples and BASELINE use.  Every other (basis, order) pair runs here: the
// same per-env arithmetic (device.cuh: sincospi, angle addition, the Kronecker feature products in the same order — the
// orders that also exist as templates give bit-identical features), but with run-time loops, tables in local memory,
// W read from global memory and one warp per CTA whose dW partial is a fixed-order shuffle butterfly.  The rest of the
// step (reduce_partials_kernel, W += dW, exchange) is the per-step path of kernels.cuh.  Slow by design (F * A warp
// reductions per step); it exists so that no valid configuration of the reference fails.
#pragma once
#include "kernels.cuh"

namespace rsrl {

constexpr int kDynMaxOrder = 7;  // validate() admits basis_order 1..7

template <typename R, int D>
struct DynTab {
    R c[D][kDynMaxOrder];  // c[d][j] = cos(pi (j+1) x^_d)   (Polynomial: x_d^(j+1))
    R s[D][kDynMaxOrder];  // s[d][j] = sin(pi (j+1) x^_d)
};

struct DynBasis {
    int basis;  // RSRL_FOURIER / RSRL_POLYNOMIAL
    int P;      // order
    __host__ __device__ int n1() const { return P + 1; }
    __host__ __device__ int features(int D) const { int f = 1; for (int d = 0; d < D; ++d) f *= P + 1; return f; }
};

// device.cuh: grid_prepare_base + grid_expand with a run-time order
template <typename R, class Dom>
__device__ __forceinline__ void dyn_prepare(const double* st, const DynBasis& b, DynTab<R, Dom::D>& t) {
    using O = RealOps<R>;
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) {
        if (b.basis == RSRL_FOURIER) {
            const double num = dsub(st[d], Dom::lo(d));
            const R xh = sizeof(R) == 4 ? (R)dmul(num, 1.0 / (Dom::hi(d) - Dom::lo(d))) : (R)ddiv(num, dsub(Dom::hi(d), Dom::lo(d)));
            R s1, c1;
            O::sincospi(xh, &s1, &c1);
            t.c[d][0] = c1;
            t.s[d][0] = s1;
            for (int j = 1; j < b.P; ++j) {
                t.c[d][j] = O::fma(t.c[d][j - 1], c1, -(t.s[d][j - 1] * s1));
                t.s[d][j] = O::fma(t.s[d][j - 1], c1, t.c[d][j - 1] * s1);
            }
        } else {
            const R x = (R)st[d];
            t.c[d][0] = x;
            t.s[d][0] = (R)0;
            for (int j = 1; j < b.P; ++j) { t.c[d][j] = t.c[d][j - 1] * x; t.s[d][j] = (R)0; }
        }
    }
}

// e_d(c) = cos + i sin of the c-th multiple (c == 0: 1 + 0i); Polynomial: x^c + 0i
template <typename R, int D>
__device__ __forceinline__ void dyn_entry(const DynTab<R, D>& t, int d, int c, bool fourier, R& re, R& im) {
    re = c == 0 ? (R)1 : t.c[d][c - 1];
    im = (c == 0 || !fourier) ? (R)0 : t.s[d][c - 1];
}

// f(k, phi_k) for k = 0..F-1 in feature order; the arithmetic of GridBasis::for_each (device.cuh) term by term
template <typename R, int D, class Fn>
__device__ __forceinline__ void dyn_for_each(const DynTab<R, D>& t, const DynBasis& b, Fn f) {
    using O = RealOps<R>;
    const int P = b.P, N1 = P + 1;
    const bool fourier = b.basis == RSRL_FOURIER;
    if (D == 2) {
        for (int i0 = 0; i0 < N1; ++i0)
            for (int i1 = 0; i1 < N1; ++i1) {
                const int c0 = P - i0, c1 = P - i1;
                R phi;
                if (c0 == 0 && c1 == 0) phi = (R)1;
                else if (c0 == 0) phi = t.c[1][c1 - 1];
                else if (c1 == 0) phi = t.c[0][c0 - 1];
                else if (fourier) phi = O::fma(t.c[0][c0 - 1], t.c[1][c1 - 1], -(t.s[0][c0 - 1] * t.s[1][c1 - 1]));
                else phi = t.c[0][c0 - 1] * t.c[1][c1 - 1];
                f(i0 * N1 + i1, phi);
            }
    } else {
        for (int i0 = 0; i0 < N1; ++i0)
            for (int i1 = 0; i1 < N1; ++i1) {
                const int c0 = P - i0, c1 = P - i1;
                R a, bq, re01, im01;
                dyn_entry<R, D>(t, 0, c0, fourier, a, bq);
                if (c1 == 0) { re01 = a; im01 = bq; }
                else if (fourier) {
                    re01 = O::fma(a, t.c[1][c1 - 1], -(bq * t.s[1][c1 - 1]));
                    im01 = O::fma(a, t.s[1][c1 - 1], bq * t.c[1][c1 - 1]);
                } else { re01 = a * t.c[1][c1 - 1]; im01 = (R)0; }
                for (int i2 = 0; i2 < N1; ++i2) {
                    const int c2 = P - i2;
                    R re012, im012;
                    if (c2 == 0) { re012 = re01; im012 = im01; }
                    else if (fourier) {
                        re012 = O::fma(re01, t.c[2][c2 - 1], -(im01 * t.s[2][c2 - 1]));
                        im012 = O::fma(re01, t.s[2][c2 - 1], im01 * t.c[2][c2 - 1]);
                    } else { re012 = re01 * t.c[2][c2 - 1]; im012 = (R)0; }
                    for (int i3 = 0; i3 < N1; ++i3) {
                        const int c3 = P - i3;
                        R phi;
                        if (c3 == 0) phi = re012;
                        else if (fourier) phi = O::fma(re012, t.c[3][c3 - 1], -(im012 * t.s[3][c3 - 1]));
                        else phi = re012 * t.c[3][c3 - 1];
                        f(((i0 * N1 + i1) * N1 + i2) * N1 + i3, phi);
                    }
                }
            }
    }
}

// sum over the 32 lanes in a fixed order (xor butterfly 16, 8, 4, 2, 1): every lane gets the same value
template <typename R>
__device__ __forceinline__ R dyn_warp_sum(R v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// One batched step (kernels.cuh: fused_step_kernel) for a run-time (basis, order).  Launch: 32 threads per CTA, no shared
// memory; SHARED weights write partials[blockIdx.x][F * AW] for reduce_partials_kernel.
template <typename R, int DOM, int AW, int MODE, bool EXT>
__global__ void __launch_bounds__(32) dyn_step_kernel(const StepArgs a, const DynBasis b) {
    using Dom = Domain<DOM>;
    using O = RealOps<R>;
    using Tab = DynTab<R, Dom::D>;
    constexpr int D = Dom::D;
    constexpr bool TDPRED = AW == 1;
    const int F = b.features(D), FA = F * AW;
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * 32 + lane;
    const bool active = i < a.n;
    const int64_t N = a.n;
    const uint64_t g = (uint64_t)(a.env_offset + i);
    const bool traces = algo_has_trace(a.algo);
    const R* Wg = static_cast<const R*>(a.W);
    auto Wat = [&](int j) -> R { return MODE == RSRL_SHARED ? __ldg(Wg + j) : Wg[(int64_t)j * N + i]; };
    auto evalQ = [&](const Tab& tab, R* q) {
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = (R)0;
        dyn_for_each<R, D>(tab, b, [&](int k, R phi) {
#pragma unroll
            for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, Wat(k * AW + c), q[c]);
        });
    };
    Tab tab_s, tab_n;
#pragma unroll
    for (int d = 0; d < D; ++d)
        for (int j = 0; j < kDynMaxOrder; ++j) { tab_s.c[d][j] = (R)0; tab_s.s[d][j] = (R)0; }
    CoreOut<R> o;
    o.coef = (R)0; o.act = 0; o.reset_before = false; o.terminated = false;
    if (active) {
        double s[D];
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = EXT ? a.ext_from[i * D + d] : a.states[i * D + d];
        auto prep = [&](const double* st, Tab& tb) { dyn_prepare<R, Dom>(st, b, tb); };
        env_core<R, DOM, AW, EXT>(a, a.t, g, s, prep, evalQ, evalQ, tab_s, tab_n, false, o, EXT ? a.ext_actions[i] : 0,
                                  EXT ? a.ext_rewards[i] : 0.0, EXT ? a.ext_term[i] != 0 : false, EXT ? a.ext_to + i * D : nullptr);
        if (a.td) static_cast<R*>(a.td)[i] = o.residual;
        if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
        if (!EXT) {
            a.ep_steps[i] = env_bookkeeping<Dom>(a, a.t, i, g, s, a.ep_steps[i], o.terminated);
            a.actions[i] = o.act;
#pragma unroll
            for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
        }
    }
    const R coef = active ? o.coef : (R)0;
    const int act = o.act;
    const bool reset_before = o.reset_before, terminated = o.terminated;
    R* part = MODE == RSRL_SHARED ? static_cast<R*>(a.partials) + (int64_t)blockIdx.x * FA : nullptr;
    R* Wm = static_cast<R*>(a.W);

    // ---- E: update (every lane walks the features: the warp sums need all 32) ----
    if (!traces) {
        dyn_for_each<R, D>(tab_s, b, [&](int k, R phi) {
            if (MODE == RSRL_PER_ENV) {
                if (active) {
                    const int64_t idx = (int64_t)(k * AW + (TDPRED ? 0 : act)) * N + i;
                    Wm[idx] = O::mul_add_unfused(coef, phi, Wm[idx]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < AW; ++c) {
                    const R v = dyn_warp_sum<R>((active && (TDPRED || c == act)) ? coef * phi : (R)0);
                    if (lane == 0) part[k * AW + c] = v;
                }
            }
        });
    } else {
        // eligibility traces: z <- rule(rate * z + grad), W += coef * z, z.reset() on terminal (kernels.cuh)
        R* Z = static_cast<R*>(a.z);
        const R rate = a.trace_rule == RSRL_TRACE_DUTCH ? (R)(a.gamma * a.lambda * (1.0 - a.alpha)) : (R)(a.gamma * a.lambda);
        dyn_for_each<R, D>(tab_s, b, [&](int k, R phi) {
#pragma unroll
            for (int c = 0; c < AW; ++c) {
                const int j = k * AW + c;
                R zv = (R)0;
                if (active) {
                    const int64_t idx = (int64_t)j * N + i;
                    zv = reset_before ? (R)0 : Z[idx];
                    zv = trace_rule<R>(a.trace_rule, rate, zv, (TDPRED || c == act) ? phi : (R)0);
                    if (MODE == RSRL_PER_ENV) Wm[idx] = O::mul_add_unfused(coef, zv, Wm[idx]);
                    Z[idx] = terminated ? (R)0 : zv;
                }
                if (MODE == RSRL_SHARED) {
                    const R v = dyn_warp_sum<R>(coef * zv);
                    if (lane == 0) part[j] = v;
                }
            }
        });
    }
}

// component entry points (kernels.cuh: basis_eval_kernel): mode 0 features, 1 Q, 2 sample, 3 mode / find_max
template <typename R, int DOM, int AW>
__global__ void dyn_eval_kernel(int mode, int64_t n, const double* __restrict__ states, const R* __restrict__ W, int64_t w_env_stride,
                                double* __restrict__ out, int32_t* __restrict__ act_out, PolicyParams pol, uint64_t draw,
                                int64_t env_offset, Counters* counters, const DynBasis b) {
    using Dom = Domain<DOM>;
    using O = RealOps<R>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int F = b.features(Dom::D);
    double s[Dom::D];
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) s[d] = states[i * Dom::D + d];
    DynTab<R, Dom::D> tab;
    dyn_prepare<R, Dom>(s, b, tab);
    if (mode == 0) {
        dyn_for_each<R, Dom::D>(tab, b, [&](int k, R phi) { out[i * F + k] = (double)phi; });
        return;
    }
    R q[AW];
#pragma unroll
    for (int c = 0; c < AW; ++c) q[c] = (R)0;
    dyn_for_each<R, Dom::D>(tab, b, [&](int k, R phi) {
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, w_env_stride ? W[(int64_t)(k * AW + c) * n + i] : W[k * AW + c], q[c]);  // w_env_stride: 0 shared W, 1 per-env W[FA][n]
    });
    if (mode == 1) {
#pragma unroll
        for (int c = 0; c < AW; ++c) out[i * AW + c] = (double)q[c];
    } else if (mode == 2) {
        bool nf = false;
        act_out[i] = policy_sample<R, AW>(pol, q, (uint64_t)(env_offset + i), draw, STREAM_BEHAVIOUR, nf);
        if (nf) atomicExch(&counters->nonfinite, 1);
    } else {
        act_out[i] = policy_mode<R, AW>(pol.policy, (R)pol.tau, q);
    }
}

}  // namespace rsrl
