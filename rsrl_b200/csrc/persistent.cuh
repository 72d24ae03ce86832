// persistent.cuh — K batched steps in ONE launch (sm_100a): a persistent grid with one CTA per SM.
//
// Each CTA owns a contiguous slice of envs; with one env per thread the env state lives in
// registers for the whole launch (HBM traffic = one read + one write per launch).  SHARED weights
// need W_{t+1} = W_t + sum over ALL envs of the step-t updates before anybody can take step t+1, so
// every step ends in a grid-wide, fixed-order (bit-reproducible) reduction of F*A values:
//
//   CTA      : env threads drop phi(s_t) rows and their scaled TD error into shared memory; F-wide
//              reducer lanes sum them slot by slot (conflict-free rows, padded to an odd stride)
//   stage 1  : every CTA publishes its partial as LL words {payload, epoch} (8-byte stores, no fence)
//   stage 2  : one leader CTA per group of ~sqrt(G) CTAs spins on its members' words, sums them in
//              CTA order and publishes the group partial (double-buffered by step parity)
//   stage 3  : every CTA spins on the group partials, sums them in group order, updates its W copy
//
// No atomics, no fences, no cooperative-groups grid.sync(): two LL hops per step.  The kernel is
// launched with cudaLaunchCooperativeKernel so that all CTAs are co-resident (the spins need it).
#pragma once
#include "kernels.cuh"

namespace rsrl {

constexpr int kMaxFan = 16;  // max CTAs per group and max groups (G <= 256)

struct SyncArgs {
    uint2* stage1;  // [G][FA * WPV]
    uint2* stage2;  // [2][n_groups][FA * WPV]
    int group_size;
    int n_groups;
};

__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll(uint2* p, uint32_t payload, uint32_t epoch) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(epoch) : "memory");
}

template <typename R> struct LL;
template <> struct LL<float> {
    static constexpr int WPV = 1;
    __device__ __forceinline__ static void publish(uint2* slot, float v, uint32_t epoch) { st_ll(slot, __float_as_uint(v), epoch); }
    __device__ __forceinline__ static bool ready(const uint2* w, uint32_t epoch) { return w[0].y == epoch; }
    __device__ __forceinline__ static void load(const uint2* slot, uint2* w) { w[0] = ld_ll(slot); }
    __device__ __forceinline__ static float value(const uint2* w) { return __uint_as_float(w[0].x); }
};
template <> struct LL<double> {
    static constexpr int WPV = 2;
    __device__ __forceinline__ static void publish(uint2* slot, double v, uint32_t epoch) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(v);
        st_ll(slot, (uint32_t)u, epoch);
        st_ll(slot + 1, (uint32_t)(u >> 32), epoch);
    }
    __device__ __forceinline__ static bool ready(const uint2* w, uint32_t epoch) { return w[0].y == epoch && w[1].y == epoch; }
    __device__ __forceinline__ static void load(const uint2* slot, uint2* w) { w[0] = ld_ll(slot); w[1] = ld_ll(slot + 1); }
    __device__ __forceinline__ static double value(const uint2* w) {
        return __longlong_as_double((long long)(((unsigned long long)w[1].x << 32) | w[0].x));
    }
};

// sum of `cnt` (<= kMaxFan) LL values at slot0 + m * stride, m ascending; all loads are issued
// before the first flag check so the L2 round trips overlap.
template <typename R>
__device__ __forceinline__ R ll_gather_sum(const uint2* slot0, size_t stride, int cnt, uint32_t epoch) {
    using L = LL<R>;
    uint2 w[kMaxFan][L::WPV];
#pragma unroll
    for (int m = 0; m < kMaxFan; ++m)
        if (m < cnt) L::load(slot0 + m * stride, w[m]);
    R sum = (R)0;
#pragma unroll
    for (int m = 0; m < kMaxFan; ++m) {
        if (m < cnt) {
            while (!L::ready(w[m], epoch)) L::load(slot0 + m * stride, w[m]);
            sum += L::value(w[m]);
        }
    }
    return sum;
}

template <typename R, int DOM, int BASIS, int P, int AW, int MODE>
__global__ void __launch_bounds__(512, 1) persistent_kernel(const StepArgs a, const int k_steps, const SyncArgs sy) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    using L = LL<R>;
    constexpr int D = Dom::D, F = GB::F, FA = F * AW;
    constexpr int FP = F | 1;  // odd row stride: conflict-free row writes (lane = slot) and reads (lane = k)
    constexpr bool TDPRED = AW == 1;
    constexpr int FApad = (FA + 3) & ~3;

    const int tid = threadIdx.x, BLOCK = blockDim.x, G = gridDim.x, b = blockIdx.x;
    const int64_t N = a.n;
    const int64_t per_cta = (N + G - 1) / G;
    const int64_t base = (int64_t)b * per_cta;
    const int64_t end = base + per_cta < N ? base + per_cta : N;
    const int n_chunks = (int)((per_cta + BLOCK - 1) / BLOCK);
    const bool resident = n_chunks == 1;  // one env per thread: state stays in registers across steps

    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* Wsm = reinterpret_cast<R*>(smem_raw);  // [FApad]
    R* dc = Wsm + FApad;                      // [BLOCK][4]  scaled TD error per action column (0 elsewhere)
    R* red = dc + (size_t)BLOCK * 4;          // [BLOCK][FP] phi(s_t) rows
    const int nseg = BLOCK / F > 0 ? BLOCK / F : 1;
    const int seg_len = (BLOCK + nseg - 1) / nseg;
    R* segpart = red + (size_t)BLOCK * FP;    // [nseg][FA]

    if (MODE == RSRL_SHARED) {
        for (int j = tid; j < FA; j += BLOCK) Wsm[j] = static_cast<const R*>(a.W)[j];
        for (int j = tid; j < BLOCK * FP; j += BLOCK) red[j] = (R)0;  // rows of idle slots stay finite (x 0 = 0)
        __syncthreads();
    }
    const R* Wg = static_cast<const R*>(a.W);

    // resident env state
    double s[D];
    int ep = 0, act = -1;
    int64_t i = base + tid;
    bool active = i < end;
    if (resident && active) {
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = a.states[i * D + d];
        ep = a.ep_steps[i];
    }

    // reducer role: (seg, k) sums phi[slot][k] * dc[slot][:] over its slots
    const bool reducer = MODE == RSRL_SHARED && tid < nseg * F;
    const int rk = tid % F, rseg = tid / F;

    for (int step = 0; step < k_steps; ++step) {
        const uint64_t t = a.t + (uint64_t)step;
        R racc[AW];
#pragma unroll
        for (int c = 0; c < AW; ++c) racc[c] = (R)0;

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            if (!resident) {
                i = base + (int64_t)chunk * BLOCK + tid;
                active = i < end;
                if (active) {
#pragma unroll
                    for (int d = 0; d < D; ++d) s[d] = a.states[i * D + d];
                    ep = a.ep_steps[i];
                }
            }
            typename GB::Tab tab_s;
            CoreOut<R> o;
            o.coef = (R)0; o.act = 0; o.terminated = false;
            if (active) {
                const uint64_t g = (uint64_t)(a.env_offset + i);
                auto evalQ = [&](const typename GB::Tab& tab, R* q) {
#pragma unroll
                    for (int c = 0; c < AW; ++c) q[c] = (R)0;
                    GB::for_each(tab, [&](int k, R phi) {
#pragma unroll
                        for (int c = 0; c < AW; ++c) {
                            const R w = MODE == RSRL_SHARED ? Wsm[k * AW + c] : Wg[(int64_t)(k * AW + c) * N + i];
                            q[c] = O::fma(phi, w, q[c]);
                        }
                    });
                };
                env_core<R, DOM, BASIS, P, AW, false>(a, t, g, s, evalQ, tab_s, o, 0, 0.0, false, nullptr);
                if (a.td) static_cast<R*>(a.td)[i] = o.residual;
                if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
                ep = env_bookkeeping<Dom>(a, t, i, g, s, ep, o.terminated);
                act = o.act;
                if (!resident) {
                    a.ep_steps[i] = ep;
                    a.actions[i] = act;
#pragma unroll
                    for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
                }
            }
            if (MODE == RSRL_PER_ENV) {
                if (active) {
                    R* Wm = static_cast<R*>(a.W);
                    GB::for_each(tab_s, [&](int k, R phi) {
                        const int64_t idx = (int64_t)(k * AW + (TDPRED ? 0 : o.act)) * N + i;
                        Wm[idx] = O::mul_add_unfused(o.coef, phi, Wm[idx]);
                    });
                }
            } else {
                // env thread -> row `tid` of red / dc
                if (active) {
                    GB::for_each(tab_s, [&](int k, R phi) { red[tid * FP + k] = phi; });
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    dc[tid * 4 + c] = (active && c < AW && (TDPRED || c == o.act)) ? o.coef : (R)0;
                __syncthreads();
                if (reducer) {
                    const int s0 = rseg * seg_len;
                    const int s1 = s0 + seg_len < BLOCK ? s0 + seg_len : BLOCK;
                    for (int slot = s0; slot < s1; ++slot) {
                        const R phi = red[slot * FP + rk];
#pragma unroll
                        for (int c = 0; c < AW; ++c) racc[c] = O::fma(phi, dc[slot * 4 + c], racc[c]);
                    }
                }
                __syncthreads();
            }
        }

        if (MODE == RSRL_SHARED) {
            if (reducer) {
#pragma unroll
                for (int c = 0; c < AW; ++c) segpart[rseg * FA + rk * AW + c] = racc[c];
            }
            __syncthreads();
            for (int j = tid; j < FA; j += BLOCK) {
                R mine = (R)0;
                for (int sg = 0; sg < nseg; ++sg) mine += segpart[sg * FA + j];
                R dW = mine;
                if (G > 1) {
                    const uint32_t epoch = (uint32_t)(t + 1);
                    const int par = (int)(t & 1);
                    const int grp = b / sy.group_size;
                    L::publish(sy.stage1 + ((size_t)b * FA + j) * L::WPV, mine, epoch);
                    if (b % sy.group_size == 0) {
                        const int first = grp * sy.group_size;
                        const int cnt = G - first < sy.group_size ? G - first : sy.group_size;
                        const R gsum = ll_gather_sum<R>(sy.stage1 + ((size_t)first * FA + j) * L::WPV, (size_t)FA * L::WPV, cnt, epoch);
                        L::publish(sy.stage2 + (((size_t)par * sy.n_groups + grp) * FA + j) * L::WPV, gsum, epoch);
                    }
                    dW = ll_gather_sum<R>(sy.stage2 + ((size_t)par * sy.n_groups * FA + j) * L::WPV, (size_t)FA * L::WPV, sy.n_groups, epoch);
                }
                Wsm[j] += dW;
            }
            __syncthreads();
        }
    }

    if (resident && active) {
        a.ep_steps[i] = ep;
        a.actions[i] = act;
#pragma unroll
        for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
    }
    if (MODE == RSRL_SHARED && b == 0) {
        for (int j = tid; j < FA; j += BLOCK) static_cast<R*>(a.W)[j] = Wsm[j];
    }
}

}  // namespace rsrl
