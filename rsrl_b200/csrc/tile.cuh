// tile.cuh — TileCoding basis (sparse features) on the fused path: BASELINE config 3
// (CartPole / SARSA / tile coding).  SHARED weights only.
//
// W is a hashed table of M rows x A columns kept in shared memory by every CTA (32 KB fp32 for
// M = 4096, A = 2).  Q(s)[a] = sum of the T active rows; the update adds the scaled TD error to the
// T active rows of column a_t.  Contributions of different envs collide on rows, so dW is summed
// with 64-bit fixed-point RED atomics straight into an L2-resident table: integer addition is
// associative, hence the result is bit-reproducible regardless of arrival order (a float atomic would
// not be).  Four dW tables rotate by step (accumulate t, read t, idle, being cleared for t+2), one
// counter barrier per step separates "everybody added" from "everybody reads".
#pragma once
#include "kernels.cuh"

namespace rsrl {

struct TileArgs {
    TileParams tp;
    int dense;                    // 1: tile_dense_kernel (per-CTA shared-memory accumulation + dense reduce-scatter), 0: RED atomics
    unsigned long long* G;        // RED path: [4][M * AW] fixed-point dW accumulators; dense path: [grid + 1][M * AW] partials + totals
    unsigned long long* barrier;  // arrival counter, zeroed by the host before every launch
    unsigned long long barrier_base;  // batched steps (fused or handle) completed before this launch: rotation index of the dW tables
    double fx_scale;              // 2^40 (f32) / 2^44 (f64)
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid barrier (all CTAs co-resident: cooperative launch)
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
        while (ld_acquire_u64(counter) < target) {}
        __threadfence();
    }
    __syncthreads();
}

template <typename R, int DOM, int AW, bool EXT>
__global__ void __launch_bounds__(512, 1) tile_persistent_kernel(const StepArgs a, const int k_steps, const TileArgs ta) {
    using Dom = Domain<DOM>;
    constexpr int D = Dom::D;
    constexpr bool TDPRED = AW == 1;
    const int tid = threadIdx.x, BLOCK = blockDim.x, G = gridDim.x, b = blockIdx.x;
    const int M = ta.tp.memory_mask + 1, MA = M * AW;
    const int64_t N = a.n;
    const int64_t per_cta = (N + G - 1) / G;
    const int64_t base = (int64_t)b * per_cta;
    const int64_t end = base + per_cta < N ? base + per_cta : N;
    const int n_chunks = (int)((per_cta + BLOCK - 1) / BLOCK);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* Wsm = reinterpret_cast<R*>(smem_raw);  // [M][AW]
    for (int j = tid; j < MA; j += BLOCK) Wsm[j] = static_cast<const R*>(a.W)[j];
    __syncthreads();

    auto evalQ = [&](const TileTab& tab, R* q) {
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = (R)0;
#pragma unroll
        for (int j = 0; j < kMaxTilings; ++j) {
            if (j < tab.n && tab.idx[j] >= 0) {
#pragma unroll
                for (int c = 0; c < AW; ++c) q[c] += Wsm[tab.idx[j] * AW + c];  // activation 1.0
            }
        }
    };
    auto prep = [&](const double* st, TileTab& tb) { tile_prepare<Dom>(st, ta.tp, tb); };
    const double inv_fx = 1.0 / ta.fx_scale;

    for (int step = 0; step < k_steps; ++step) {
        const uint64_t t = a.t + (uint64_t)step;
        const unsigned long long rot = ta.barrier_base + (unsigned long long)step;
        unsigned long long* Gt = ta.G + (size_t)(rot & 3) * MA;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int64_t i = base + (int64_t)chunk * BLOCK + tid;
            if (i < end) {
                const uint64_t g = (uint64_t)(a.env_offset + i);
                double s[D];
#pragma unroll
                for (int d = 0; d < D; ++d) s[d] = EXT ? a.ext_from[i * D + d] : a.states[i * D + d];
                TileTab tab_s, tab_n;
                CoreOut<R> o;
                env_core<R, DOM, AW, EXT>(a, t, g, s, prep, evalQ, evalQ, tab_s, tab_n, false, o, EXT ? a.ext_actions[i] : 0,
                                          EXT ? a.ext_rewards[i] : 0.0, EXT ? a.ext_term[i] != 0 : false,
                                          EXT ? a.ext_to + i * D : nullptr);
                if (a.td) static_cast<R*>(a.td)[i] = o.residual;
                if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
                // dW[row, a_t] += coef for every active row (activation 1.0), fixed point
                const long long fx = __double2ll_rn((double)o.coef * ta.fx_scale);
                const int col = TDPRED ? 0 : o.act;
#pragma unroll
                for (int j = 0; j < kMaxTilings; ++j)
                    if (j < tab_s.n && tab_s.idx[j] >= 0) atomicAdd(Gt + (size_t)tab_s.idx[j] * AW + col, (unsigned long long)fx);
                if (!EXT) {
                    a.ep_steps[i] = env_bookkeeping<Dom>(a, t, i, g, s, a.ep_steps[i], o.terminated);
                    a.actions[i] = o.act;
#pragma unroll
                    for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
                }
            }
        }
        if (G > 1) grid_barrier(ta.barrier, ((unsigned long long)step + 1ull) * (unsigned long long)G);  // launch-local target: the host zeroes the counter before every launch (launches may differ in grid size)
        else { __threadfence(); __syncthreads(); }
        // every CTA applies the same dW to its W copy; the table for step t+2 is cleared cooperatively
        unsigned long long* Gz = ta.G + (size_t)((rot + 2) & 3) * MA;
        for (int j = tid; j < MA; j += BLOCK) {
            const long long v = (long long)__ldcg(Gt + j);
            Wsm[j] += (R)((double)v * inv_fx);
        }
        for (int j = b * BLOCK + tid; j < MA; j += G * BLOCK) Gz[j] = 0ull;
        __syncthreads();
    }
    if (b == 0)
        for (int j = tid; j < MA; j += BLOCK) static_cast<R*>(a.W)[j] = Wsm[j];
}

// ---------------------------------------------------------------------------------------------------------------
// Dense variant (default when W and a 64-bit accumulator table both fit in shared memory): the RED version spends its
// time in ~2.1 M L2 atomics per step that collide on a few hundred hot rows.  Here every CTA accumulates its envs'
// contributions into a shared-memory fixed-point table (integer adds: order independent => bit-reproducible), then the
// tables are summed across CTAs with a dense reduce-scatter through L2:
//   P[b][*] = CTA b's table  | grid barrier |  CTA b sums its slice of all P[*]  -> T  | grid barrier |  every CTA reads T.
// (round 1; ~28 MB of L2 traffic and two counter barriers per step instead of the atomics).  Round 2: the per-CTA tables are sparse, so
// each CTA pushes only its non-zero entries into a rotating global table with integer reductions: one barrier, ~1/50 of the traffic.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier_local(unsigned long long* counter, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
        while (ld_acquire_u64(counter) < target) {}
        __threadfence();
    }
    __syncthreads();
}

template <typename R, int DOM, int AW, bool EXT, int MAXT = 512, int TMAX = kMaxTilings>
__global__ void __launch_bounds__(MAXT, 1) tile_dense_kernel(const StepArgs a, const int k_steps, const TileArgs ta) {
    using Dom = Domain<DOM>;
    constexpr int D = Dom::D;
    constexpr bool TDPRED = AW == 1;
    const int tid = threadIdx.x, BLOCK = blockDim.x, G = gridDim.x, b = blockIdx.x, lane = tid & 31;
    const int M = ta.tp.memory_mask + 1, MA = M * AW;
    const int64_t N = a.n;
    const int64_t per_cta = (N + G - 1) / G;
    const int64_t base = (int64_t)b * per_cta;
    const int64_t end = base + per_cta < N ? base + per_cta : N;
    const int n_chunks = (int)((per_cta + BLOCK - 1) / BLOCK);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* Gs = reinterpret_cast<unsigned long long*>(smem_raw);   // [M][AW] fixed-point dW of this CTA, this step
    R* Wsm = reinterpret_cast<R*>(Gs + MA);                                      // [M][AW]
    for (int j = tid; j < MA; j += BLOCK) { Wsm[j] = static_cast<const R*>(a.W)[j]; Gs[j] = 0ull; }
    __syncthreads();

    auto evalQ = [&](const TileTab& tab, R* q) {
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = (R)0;
#pragma unroll
        for (int j = 0; j < TMAX; ++j) {
            if (j < tab.n && tab.idx[j] >= 0) {
#pragma unroll
                for (int c = 0; c < AW; ++c) q[c] += Wsm[tab.idx[j] * AW + c];  // activation 1.0
            }
        }
    };
    auto prep = [&](const double* st, TileTab& tb) { tile_prepare<Dom, TMAX>(st, ta.tp, tb); };
    const double inv_fx = 1.0 / ta.fx_scale;

    for (int step = 0; step < k_steps; ++step) {
        const uint64_t t = a.t + (uint64_t)step;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int64_t i = base + (int64_t)chunk * BLOCK + tid;
            const bool active = i < end;
            long long fx = 0;
            int col = 0;
            TileTab tab_s;
            tab_s.n = 0;
            if (active) {
                const uint64_t g = (uint64_t)(a.env_offset + i);
                double s[D];
#pragma unroll
                for (int d = 0; d < D; ++d) s[d] = EXT ? a.ext_from[i * D + d] : a.states[i * D + d];
                TileTab tab_n;
                CoreOut<R> o;
                env_core<R, DOM, AW, EXT>(a, t, g, s, prep, evalQ, evalQ, tab_s, tab_n, false, o, EXT ? a.ext_actions[i] : 0,
                                          EXT ? a.ext_rewards[i] : 0.0, EXT ? a.ext_term[i] != 0 : false,
                                          EXT ? a.ext_to + i * D : nullptr);
                if (a.td) static_cast<R*>(a.td)[i] = o.residual;
                if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
                fx = __double2ll_rn((double)o.coef * ta.fx_scale);
                col = TDPRED ? 0 : o.act;
                if (!EXT) {
                    a.ep_steps[i] = env_bookkeeping<Dom>(a, t, i, g, s, a.ep_steps[i], o.terminated);
                    a.actions[i] = o.act;
#pragma unroll
                    for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
                }
            }
            // dW[row, a_t] += coef for every active row (activation 1.0).  When the whole warp hits the same entry (envs
            // start in the same tiles) the warp adds once.  (Grouping the lanes by entry with match.any + three 32-bit warp
            // reductions per tiling instead: 132 us per step against 40 — measured, profiles/r02_cfg3_tile.md.)
#pragma unroll
            for (int j = 0; j < TMAX; ++j) {
                if (j >= ta.tp.n_tilings) break;  // uniform
                const int key = (active && j < tab_s.n && tab_s.idx[j] >= 0) ? tab_s.idx[j] * AW + col : -1;
                int same;
                __match_all_sync(0xffffffffu, key, &same);
                if (same) {
                    if (key >= 0) {
                        long long v = fx;
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                        if (lane == 0) atomicAdd(Gs + key, (unsigned long long)v);
                    }
                } else if (key >= 0) {
                    atomicAdd(Gs + key, (unsigned long long)fx);
                }
            }
        }
        __syncthreads();
        if (G > 1) {
            // sparse push: only the entries this CTA touched (a few hundred of M * A: the envs crowd into few tiles) are added to the
            // step's global table with 64-bit integer reductions (order independent); ONE grid barrier; every CTA reads the table.
            // Three tables rotate by step: accumulate + read t, (t + 1), and the one being cleared for t + 2 — it was read during step
            // t - 1, and every CTA has left that phase before it arrives at this step's barrier.
            const unsigned long long rot = ta.barrier_base + (unsigned long long)step;
            unsigned long long* Tt = ta.G + (size_t)(rot % 3ull) * MA;
            unsigned long long* Tz = ta.G + (size_t)((rot + 2ull) % 3ull) * MA;
            for (int j = tid; j < MA; j += BLOCK) {
                const unsigned long long v = Gs[j];
                if (v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(Tt + j), "l"(v) : "memory"); Gs[j] = 0ull; }
            }
            grid_barrier_local(ta.barrier, (unsigned long long)(step + 1) * (unsigned long long)G);
            for (int j = tid; j < MA; j += BLOCK) Wsm[j] += (R)((double)(long long)__ldcg(Tt + j) * inv_fx);
            for (int j = b * BLOCK + tid; j < MA; j += G * BLOCK) Tz[j] = 0ull;
        } else {
            for (int j = tid; j < MA; j += BLOCK) { Wsm[j] += (R)((double)(long long)Gs[j] * inv_fx); Gs[j] = 0ull; }
        }
        __syncthreads();
    }
    if (b == 0)
        for (int j = tid; j < MA; j += BLOCK) static_cast<R*>(a.W)[j] = Wsm[j];
}

// component entry points for the tile basis: mode 0 dense features (N x M), 1 Q, 2 sample, 3 find_max
template <typename R, int DOM, int AW>
__global__ void tile_eval_kernel(int mode, int64_t n, const double* __restrict__ states, const R* __restrict__ W, TileParams tp,
                                 double* __restrict__ out, int32_t* __restrict__ act_out, PolicyParams pol, uint64_t draw,
                                 int64_t env_offset, Counters* counters) {
    using Dom = Domain<DOM>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[Dom::D];
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) s[d] = states[i * Dom::D + d];
    TileTab tab;
    tile_prepare<Dom>(s, tp, tab);
    const int M = tp.memory_mask + 1;
    if (mode == 0) {
        for (int j = 0; j < tab.n; ++j) if (tab.idx[j] >= 0) out[i * M + tab.idx[j]] = 1.0;  // `out` is zero-filled by the host
        return;
    }
    R q[AW];
#pragma unroll
    for (int c = 0; c < AW; ++c) q[c] = (R)0;
    for (int j = 0; j < tab.n; ++j) {
        if (tab.idx[j] < 0) continue;
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] += W[tab.idx[j] * AW + c];
    }
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

}  // namespace rsrl
