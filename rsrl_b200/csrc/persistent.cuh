// persistent.cuh — K batched steps in ONE launch (sm_100a): a persistent grid with one CTA per SM.
//
// Each CTA owns a contiguous slice of envs; with one env per thread the env state and the Fourier
// tables of the current state live in registers for the whole launch (HBM traffic = one read + one
// write of the state per launch).  SHARED weights need W_{t+1} = W_t + sum over ALL envs of the
// step-t updates before anybody can take step t+1, so every step ends in a grid-wide, fixed-order
// (bit-reproducible) reduction of F*A values:
//
//   CTA    : env threads write phi(s_t) (feature-major rows, lane = env slot: conflict-free) while
//            they evaluate Q(s_t), and their scaled TD error per action column; (feature, segment)
//            reducer threads then sum 4 slots per LDS.128 in slot order.
//   hop 1  : every CTA publishes its partial as 16-byte LL lines {3 payload words, epoch}; the leader
//            of each group of ~sqrt(G) CTAs spins on its members' lines and sums them in CTA order.
//   hop 2  : leaders exchange group partials all-to-all (parity double-buffered), sum in group order.
//   hop 3  : each leader publishes the total; its members spin on one line set and update their W copy.
//
// No atomics, no fences, no grid.sync(): three LL hops per step, and only ~16K 16-byte polls in
// flight chip-wide (the first version polled 192K 8-byte words from every CTA and saturated L2:
// profiles/r01_persistent_v1.md).  Launched with cudaLaunchCooperativeKernel so that all CTAs
// are co-resident (the spins need it).
#pragma once
#include "kernels.cuh"

namespace rsrl {

constexpr int kMaxFan = 16;  // max CTAs per group and max groups (G <= 256)

struct SyncArgs {
    uint4* stage1;  // [G][NL]             member partials
    uint4* stage2;  // [2][n_groups][NL]   group partials (parity)
    uint4* stage3;  // [2][n_groups][NL]   totals (parity)
    int group_size;
    int n_groups;
};

__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t epoch) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(epoch) : "memory");
}

// one 16-byte LL line carries VPL values + the epoch in the last word
template <typename R> struct LL;
template <> struct LL<float> {
    static constexpr int VPL = 3;
    __device__ __forceinline__ static void publish(uint4* line, const float* v, uint32_t epoch) {
        st_ll(line, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), epoch);
    }
    __device__ __forceinline__ static void add(const uint4& w, float* acc) {
        acc[0] += __uint_as_float(w.x); acc[1] += __uint_as_float(w.y); acc[2] += __uint_as_float(w.z);
    }
};
template <> struct LL<double> {
    static constexpr int VPL = 1;
    __device__ __forceinline__ static void publish(uint4* line, const double* v, uint32_t epoch) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(v[0]);
        st_ll(line, (uint32_t)u, (uint32_t)(u >> 32), 0u, epoch);
    }
    __device__ __forceinline__ static void add(const uint4& w, double* acc) {
        acc[0] += __longlong_as_double((long long)(((unsigned long long)w.y << 32) | w.x));
    }
};

// acc = sum over m = 0..cnt-1 (ascending) of the LL line at line0 + m * stride.  Loads are issued
// four at a time before their flags are checked so the L2 round trips overlap.
template <typename R>
__device__ __forceinline__ void ll_gather_sum(const uint4* line0, size_t stride, int cnt, uint32_t epoch, R* acc) {
    using L = LL<R>;
#pragma unroll
    for (int v = 0; v < L::VPL; ++v) acc[v] = (R)0;
    for (int m0 = 0; m0 < cnt; m0 += 4) {
        uint4 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (m0 + u < cnt) w[u] = ld_ll(line0 + (size_t)(m0 + u) * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (m0 + u < cnt) {
                while (w[u].w != epoch) w[u] = ld_ll(line0 + (size_t)(m0 + u) * stride);
                L::add(w[u], acc);
            }
        }
    }
}

template <typename R> struct Vec16;  // 16-byte shared-memory vector of R
template <> struct Vec16<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct Vec16<double> { typedef double2 type; static constexpr int N = 2; };
__device__ __forceinline__ float vget(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ double vget(const double2& v, int i) { return i == 0 ? v.x : v.y; }

template <typename R, int DOM, int BASIS, int P, int AW, int MODE>
__global__ void __launch_bounds__(512, 1) persistent_kernel(const StepArgs a, const int k_steps, const SyncArgs sy, const int cap) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    using L = LL<R>;
    using V = Vec16<R>;
    typedef typename V::type vec_t;
    constexpr int D = Dom::D, F = GB::F, FA = F * AW;
    constexpr bool TDPRED = AW == 1;
    constexpr int FApad = (FA + 3) & ~3;
    constexpr int NL = (FA + L::VPL - 1) / L::VPL;  // LL lines per partial

    const int tid = threadIdx.x, BLOCK = blockDim.x, G = gridDim.x, b = blockIdx.x;
    const int64_t N = a.n;
    const int64_t per_cta = (N + G - 1) / G;
    const int64_t base = (int64_t)b * per_cta;
    const int64_t end = base + per_cta < N ? base + per_cta : N;
    const int n_chunks = (int)((per_cta + BLOCK - 1) / BLOCK);
    const bool resident = n_chunks == 1;  // one env per thread: state stays in registers across steps

    // shared memory (SHARED mode): cap = padded slot count, a multiple of V::N with cap / V::N odd, so
    // that the 8 lanes of a quarter warp reading red[k][slot..] at consecutive k hit 8 distinct 16-byte
    // bank groups.  Rows of slots >= BLOCK are zero and stay zero.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* Wsm = reinterpret_cast<R*>(smem_raw);  // [FApad]
    R* red = Wsm + FApad;                     // [F][cap]   phi(s_t), feature-major
    R* dcs = red + (size_t)F * cap;           // [AW][cap]  scaled TD error in the action's row, 0 elsewhere
    const int nseg = BLOCK / F > 0 ? BLOCK / F : 1;
    const int seg_len = (((cap + nseg - 1) / nseg) + V::N - 1) / V::N * V::N;
    R* segpart = dcs + (size_t)AW * cap;      // [nseg][FA]

    if (MODE == RSRL_SHARED) {
        for (int j = tid; j < FA; j += BLOCK) Wsm[j] = static_cast<const R*>(a.W)[j];
        for (int j = tid; j < (F + AW) * cap; j += BLOCK) red[j] = (R)0;
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
    typename GB::Tab tab_s, tab_n;
    bool have_tab = false;  // tab_s holds the tables of s (carried from the previous step's s')

    // reducer role: (seg, k) sums phi[k][slot] * dcs[:][slot] over its slots
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
                have_tab = false;
                if (active) {
#pragma unroll
                    for (int d = 0; d < D; ++d) s[d] = a.states[i * D + d];
                    ep = a.ep_steps[i];
                }
            }
            CoreOut<R> o;
            o.coef = (R)0; o.act = 0; o.terminated = false;
            if (active) {
                const uint64_t g = (uint64_t)(a.env_offset + i);
                auto evalS = [&](const typename GB::Tab& tab, R* q) {  // Q(s_t) and, SHARED, the phi(s_t) row
#pragma unroll
                    for (int c = 0; c < AW; ++c) q[c] = (R)0;
                    GB::for_each(tab, [&](int k, R phi) {
                        if (MODE == RSRL_SHARED) red[k * cap + tid] = phi;
#pragma unroll
                        for (int c = 0; c < AW; ++c) {
                            const R w = MODE == RSRL_SHARED ? Wsm[k * AW + c] : Wg[(int64_t)(k * AW + c) * N + i];
                            q[c] = O::fma(phi, w, q[c]);
                        }
                    });
                };
                auto evalN = [&](const typename GB::Tab& tab, R* q) {
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
                env_core<R, DOM, BASIS, P, AW, false>(a, t, g, s, evalS, evalN, tab_s, tab_n, have_tab, o, 0, 0.0, false, nullptr);
                if (a.td) static_cast<R*>(a.td)[i] = o.residual;
                if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
                if (MODE == RSRL_PER_ENV) {
                    R* Wm = static_cast<R*>(a.W);
                    GB::for_each(tab_s, [&](int k, R phi) {
                        const int64_t idx = (int64_t)(k * AW + (TDPRED ? 0 : o.act)) * N + i;
                        Wm[idx] = O::mul_add_unfused(o.coef, phi, Wm[idx]);
                    });
                }
                bool was_reset;
                ep = env_bookkeeping<Dom>(a, t, i, g, s, ep, o.terminated, &was_reset);
                act = o.act;
                have_tab = resident && !was_reset;  // s_{t+1} = s': reuse its tables
                if (have_tab) tab_s = tab_n;
                if (!resident) {
                    a.ep_steps[i] = ep;
                    a.actions[i] = act;
#pragma unroll
                    for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
                }
            }
            if (MODE == RSRL_SHARED) {
#pragma unroll
                for (int c = 0; c < AW; ++c) dcs[c * cap + tid] = (active && (TDPRED || c == o.act)) ? o.coef : (R)0;
                // (a slot idle in this chunk keeps a stale but finite phi row; its dcs entries are 0)
                __syncthreads();
                if (reducer) {
                    const int s0 = rseg * seg_len;
                    const int s1 = s0 + seg_len < cap ? s0 + seg_len : cap;
                    const R* prow = red + (size_t)rk * cap;
                    for (int slot = s0; slot < s1; slot += V::N) {
                        const vec_t pv = *reinterpret_cast<const vec_t*>(prow + slot);
                        vec_t dv[AW];
#pragma unroll
                        for (int c = 0; c < AW; ++c) dv[c] = *reinterpret_cast<const vec_t*>(dcs + (size_t)c * cap + slot);
#pragma unroll
                        for (int u = 0; u < V::N; ++u) {
#pragma unroll
                            for (int c = 0; c < AW; ++c) racc[c] = O::fma(vget(pv, u), vget(dv[c], u), racc[c]);
                        }
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
            for (int j = tid; j < NL; j += BLOCK) {
                R mine[L::VPL], dW[L::VPL];
#pragma unroll
                for (int v = 0; v < L::VPL; ++v) {
                    const int idx = j * L::VPL + v;
                    R m = (R)0;
                    if (idx < FA)
                        for (int sg = 0; sg < nseg; ++sg) m += segpart[sg * FA + idx];
                    mine[v] = m;
                    dW[v] = m;
                }
                if (G > 1) {
                    const uint32_t epoch = (uint32_t)(t + 1);
                    const int par = (int)(t & 1);
                    const int grp = b / sy.group_size;
                    const size_t pslot = ((size_t)par * sy.n_groups + grp) * NL + j;
                    L::publish(sy.stage1 + (size_t)b * NL + j, mine, epoch);
                    if (b % sy.group_size == 0) {
                        const int first = grp * sy.group_size;
                        const int cnt = G - first < sy.group_size ? G - first : sy.group_size;
                        R gsum[L::VPL];
                        ll_gather_sum<R>(sy.stage1 + (size_t)first * NL + j, (size_t)NL, cnt, epoch, gsum);
                        L::publish(sy.stage2 + pslot, gsum, epoch);
                        ll_gather_sum<R>(sy.stage2 + (size_t)par * sy.n_groups * NL + j, (size_t)NL, sy.n_groups, epoch, dW);
                        L::publish(sy.stage3 + pslot, dW, epoch);
                    } else {
                        ll_gather_sum<R>(sy.stage3 + pslot, (size_t)NL, 1, epoch, dW);
                    }
                }
#pragma unroll
                for (int v = 0; v < L::VPL; ++v) {
                    const int idx = j * L::VPL + v;
                    if (idx < FA) Wsm[idx] += dW[v];
                }
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
