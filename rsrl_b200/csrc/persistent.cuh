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
//            of each group of ~sqrt(G) CTAs collects its members' lines — one line per thread, so all
//            L2 round trips overlap — and sums them in CTA order through shared memory.
//   hop 2  : leaders publish the group partials (parity double-buffered); every CTA collects all of
//            them (again one line per thread), sums in group order and updates its W copy.
//            (Private per-destination mailboxes were tried: the 5 K strong stores per leader cost
//            3 K cycles, more than the shared lines' polling contention.)
//
// No atomics, no fences, no grid.sync(): two LL hops per step with ~64 K 16-byte polls in flight
// chip-wide.  (v2 polled 192 K 8-byte words with 12-16 dependent polls per thread and spent 4.7 us
// per step in the exchange: profiles/r01_persistent_v1.md.)  Launched with
// cudaLaunchCooperativeKernel so that all CTAs are co-resident (the spins need it).
#pragma once
#include "kernels.cuh"

namespace rsrl {

constexpr int kMaxFan = 16;  // max CTAs per group and max groups (G <= 256)
constexpr int kModeSharedTrace = 2;  // internal MODE: SHARED weights + per-env traces kept in shared memory

struct SyncArgs {
    uint4* stage1;  // [G][NL]             member partials
    uint4* stage2;  // [2][n_groups][NL]   group partials (parity)
    int group_size;
    int n_groups;
};

// Cross-GPU exchange (one process per GPU): every rank owns an inbox of 8-byte LL words
// {payload, epoch} that its peers write through NVLink (cudaIpc-mapped pointers).
constexpr int kMaxRanks = 8;
struct PeerArgs {
    uint2* inbox[kMaxRanks];  // inbox[r]: rank r's mailbox [2][world][FA * WPV] as seen from this GPU
    uint4* stage3;            // [2][NL] local broadcast of the all-GPU total (parity)
    int rank, world;
};

__device__ __forceinline__ uint2 ld_ll8_sys(const uint2* p) {
    uint2 v;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll8_sys(uint2* p, uint32_t payload, uint32_t epoch) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(epoch) : "memory");
}

__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t epoch) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(epoch) : "memory");
}

// one 16-byte LL line carries VPL values + the epoch in the last word
template <typename R> struct LL;
template <> struct LL<float> {
    static constexpr int VPL = 3;
    __device__ __forceinline__ static void publish(uint4* line, const float* v, uint32_t epoch) {
        st_ll(line, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), epoch);
    }
    __device__ __forceinline__ static void unpack(const uint4& w, float* out) {
        out[0] = __uint_as_float(w.x); out[1] = __uint_as_float(w.y); out[2] = __uint_as_float(w.z);
    }
};
template <> struct LL<double> {
    static constexpr int VPL = 1;
    __device__ __forceinline__ static void publish(uint4* line, const double* v, uint32_t epoch) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(v[0]);
        st_ll(line, (uint32_t)u, (uint32_t)(u >> 32), 0u, epoch);
    }
    __device__ __forceinline__ static void unpack(const uint4& w, double* out) {
        out[0] = __longlong_as_double((long long)(((unsigned long long)w.y << 32) | w.x));
    }
};

// sum of rows[m * stride], m = 0..cnt-1 ascending; loads issued four at a time (latency overlap), fixed association
template <typename R>
__device__ __forceinline__ R sum_rows(const R* rows, int stride, int cnt) {
    R acc = (R)0;
    int m = 0;
    for (; m + 4 <= cnt; m += 4) {
        const R v0 = rows[(size_t)(m + 0) * stride], v1 = rows[(size_t)(m + 1) * stride];
        const R v2 = rows[(size_t)(m + 2) * stride], v3 = rows[(size_t)(m + 3) * stride];
        acc = (((acc + v0) + v1) + v2) + v3;
    }
    for (; m < cnt; ++m) acc += rows[(size_t)m * stride];
    return acc;
}

template <typename R> struct Vec16;  // 16-byte shared-memory vector of R
template <> struct Vec16<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct Vec16<double> { typedef double2 type; static constexpr int N = 2; };
__device__ __forceinline__ float vget(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ double vget(const double2& v, int i) { return i == 0 ? v.x : v.y; }

template <typename R, int DOM, int BASIS, int P, int AW, int MODE>
__global__ void __launch_bounds__(512, 1) persistent_kernel(const StepArgs a, const int k_steps, const SyncArgs sy, const int cap, const PeerArgs pe) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    using L = LL<R>;
    using V = Vec16<R>;
    typedef typename V::type vec_t;
    constexpr int D = Dom::D, F = GB::F, FA = F * AW;
    constexpr bool TDPRED = AW == 1;
    constexpr bool SHAREDW = MODE != RSRL_PER_ENV;          // one replicated W, dW reduced over the grid
    constexpr bool TRACE = MODE == kModeSharedTrace;        // eligibility traces resident in shared memory
    constexpr int ROWS = TRACE ? FA : F;                    // rows of the reduce buffer: z (F*A) or phi(s_t) (F)
    constexpr int NDC = TRACE ? 1 : AW;                     // rows of scaled TD errors
    constexpr int WS = 4;            // padded row stride of the shared W copy: one LDS.128 per feature row
    constexpr int FApad = F * WS;
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
    R* red = Wsm + FApad;                     // [ROWS][cap] phi(s_t) feature-major, or the traces z[F*A][slot]
    R* dcs = red + (size_t)ROWS * cap;        // [NDC][cap]  scaled TD error in the action's row, 0 elsewhere
    const int nseg = BLOCK / ROWS > 0 ? BLOCK / ROWS : 1;
    const int seg_len = (((cap + nseg - 1) / nseg) + V::N - 1) / V::N * V::N;
    R* segpart = dcs + (size_t)NDC * cap;     // [nseg][FA]
    constexpr int NLV = NL * L::VPL;          // values per partial, padded to whole LL lines
    R* stgA = segpart + (size_t)nseg * FA;    // [kMaxFan + 1][NLV] hop-1 staging (leader) + the group partial
    R* stgB = stgA + (size_t)(kMaxFan + 1) * NLV;  // [kMaxFan][NLV] own partial, then hop-2 staging

    if (SHAREDW) {
        for (int j = tid; j < FA; j += BLOCK) Wsm[(j / AW) * WS + j % AW] = static_cast<const R*>(a.W)[j];
        for (int j = tid; j < (ROWS + NDC) * cap; j += BLOCK) red[j] = (R)0;
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
    if (TRACE && active) {  // this env's trace column: HBM -> shared memory once per launch
        const R* Z = static_cast<const R*>(a.z);
        for (int j = 0; j < FA; ++j) red[(size_t)j * cap + tid] = Z[(int64_t)j * N + i];
    }
    typename GB::Tab tab_s, tab_n;
    bool have_tab = false;  // tab_s holds the tables of s (carried from the previous step's s')

    // reducer role: (seg, k) sums phi[k][slot] * dcs[:][slot] over its slots
    const bool reducer = SHAREDW && tid < nseg * ROWS;
    const int rk = tid % ROWS, rseg = tid / ROWS;

    const bool prof = a.phase_prof != nullptr && tid == 0;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c0 = 0, c1 = 0, c2 = 0;
    for (int step = 0; step < k_steps; ++step) {
        const uint64_t t = a.t + (uint64_t)step;
        if (prof) c0 = clock64();
        R racc[2][NDC];
#pragma unroll
        for (int c = 0; c < NDC; ++c) racc[0][c] = racc[1][c] = (R)0;

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
                auto wrow = [&](int k, R* w) {  // W[k][0..AW): SHARED = 16-byte broadcast loads, PER_ENV = coalesced global
                    if (SHAREDW) {
                        const vec_t* p = reinterpret_cast<const vec_t*>(Wsm + k * WS);
                        vec_t v0 = p[0];
                        if (V::N == 4) {
#pragma unroll
                            for (int c = 0; c < AW; ++c) w[c] = vget(v0, c);
                        } else {
                            vec_t v1 = AW > 2 ? p[1] : v0;
#pragma unroll
                            for (int c = 0; c < AW; ++c) w[c] = c < 2 ? vget(v0, c) : vget(v1, c - 2);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < AW; ++c) w[c] = Wg[(int64_t)(k * AW + c) * N + i];
                    }
                };
                auto evalS = [&](const typename GB::Tab& tab, R* q) {  // Q(s_t) and, SHARED, the phi(s_t) row
#pragma unroll
                    for (int c = 0; c < AW; ++c) q[c] = (R)0;
                    GB::for_each(tab, [&](int k, R phi) {
                        if (SHAREDW && !TRACE) red[k * cap + tid] = phi;
                        R w[AW];
                        wrow(k, w);
#pragma unroll
                        for (int c = 0; c < AW; ++c) q[c] = O::fma(phi, w[c], q[c]);
                    });
                };
                auto evalN = [&](const typename GB::Tab& tab, R* q) {
#pragma unroll
                    for (int c = 0; c < AW; ++c) q[c] = (R)0;
                    GB::for_each(tab, [&](int k, R phi) {
                        R w[AW];
                        wrow(k, w);
#pragma unroll
                        for (int c = 0; c < AW; ++c) q[c] = O::fma(phi, w[c], q[c]);
                    });
                };
                auto prep = [](const double* st, typename GB::Tab& tb) { grid_prepare<R, Dom, P, BASIS>(st, tb); };
                env_core<R, DOM, AW, false>(a, t, g, s, prep, evalS, evalN, tab_s, tab_n, have_tab, o, 0, 0.0, false, nullptr);
                if (a.td) static_cast<R*>(a.td)[i] = o.residual;
                if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
                if (MODE == RSRL_PER_ENV) {
                    R* Wm = static_cast<R*>(a.W);
                    GB::for_each(tab_s, [&](int k, R phi) {
                        const int64_t idx = (int64_t)(k * AW + (TDPRED ? 0 : o.act)) * N + i;
                        Wm[idx] = O::mul_add_unfused(o.coef, phi, Wm[idx]);
                    });
                }
                if (TRACE) {
                    // z <- rule(rate * z + grad) on this env's column (traces.rs:127-129,196-240); the CTA reduce below
                    // then forms sum_i coef_i * z_i; z.reset() on terminal happens after the reduce.
                    const R rate = a.trace_rule == RSRL_TRACE_DUTCH ? (R)(a.gamma * a.lambda * (1.0 - a.alpha)) : (R)(a.gamma * a.lambda);
                    GB::for_each(tab_s, [&](int k, R phi) {
#pragma unroll
                        for (int c = 0; c < AW; ++c) {
                            R* zp = red + (size_t)(k * AW + c) * cap + tid;
                            const R zv = o.reset_before ? (R)0 : *zp;
                            *zp = trace_rule<R>(a.trace_rule, rate, zv, (TDPRED || c == o.act) ? phi : (R)0);
                        }
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
            if (prof) { c1 = clock64(); pc[0] += c1 - c0; c0 = c1; }
            if (SHAREDW) {
                if (TRACE) {
                    dcs[tid] = active ? o.coef : (R)0;
                } else {
#pragma unroll
                    for (int c = 0; c < AW; ++c) dcs[c * cap + tid] = (active && (TDPRED || c == o.act)) ? o.coef : (R)0;
                }
                // (a slot idle in this chunk keeps a stale but finite row; its dcs entries are 0)
                __syncthreads();
                if (prof) { c1 = clock64(); pc[1] += c1 - c0; c0 = c1; }
                if (reducer) {
                    const int s0 = rseg * seg_len;
                    const int s1 = s0 + seg_len < cap ? s0 + seg_len : cap;
                    const R* prow = red + (size_t)rk * cap;
                    // two interleaved accumulator sets (even / odd 16-byte groups) hide the LDS + FMA latency
                    for (int slot = s0; slot < s1; slot += 2 * V::N) {
                        const bool two = slot + V::N < s1;
                        const vec_t pv0 = *reinterpret_cast<const vec_t*>(prow + slot);
                        const vec_t pv1 = two ? *reinterpret_cast<const vec_t*>(prow + slot + V::N) : pv0;
                        vec_t dv0[NDC], dv1[NDC];
#pragma unroll
                        for (int c = 0; c < NDC; ++c) {
                            dv0[c] = *reinterpret_cast<const vec_t*>(dcs + (size_t)c * cap + slot);
                            dv1[c] = two ? *reinterpret_cast<const vec_t*>(dcs + (size_t)c * cap + slot + V::N) : dv0[c];
                        }
#pragma unroll
                        for (int u = 0; u < V::N; ++u) {
#pragma unroll
                            for (int c = 0; c < NDC; ++c) {
                                racc[0][c] = O::fma(vget(pv0, u), vget(dv0[c], u), racc[0][c]);
                                if (two) racc[1][c] = O::fma(vget(pv1, u), vget(dv1[c], u), racc[1][c]);
                            }
                        }
                    }
                }
                __syncthreads();
                if (TRACE && active && o.terminated) {  // trace.reset() (sarsa_lambda.rs:78, q_lambda.rs:81, td_lambda.rs:61)
                    for (int j = 0; j < FA; ++j) red[(size_t)j * cap + tid] = (R)0;
                }
            }
        }

        if (SHAREDW) {
            if (reducer) {
#pragma unroll
                for (int c = 0; c < NDC; ++c) segpart[rseg * FA + (TRACE ? rk : rk * AW + c)] = racc[0][c] + racc[1][c];
            }
            __syncthreads();
            if (prof) { c1 = clock64(); pc[2] += c1 - c0; c0 = c1; }
            const uint32_t epoch = (uint32_t)(t + 1);
            const int par = (int)(t & 1);
            const int grp = b / sy.group_size;
            // this CTA's partial: segment-order sum, loads issued four at a time so their latencies overlap
            for (int idx = tid; idx < NLV; idx += BLOCK) {
                R m = (R)0;
                if (idx < FA) {
                    int sg = 0;
                    for (; sg + 4 <= nseg; sg += 4) {
                        const R v0 = segpart[(sg + 0) * FA + idx], v1 = segpart[(sg + 1) * FA + idx];
                        const R v2 = segpart[(sg + 2) * FA + idx], v3 = segpart[(sg + 3) * FA + idx];
                        m = (((m + v0) + v1) + v2) + v3;
                    }
                    for (; sg < nseg; ++sg) m += segpart[sg * FA + idx];
                }
                stgB[idx] = m;
            }
            __syncthreads();
            if (G > 1) {
                // hop 1: publish the partial as LL lines
                for (int j = tid; j < NL; j += BLOCK) L::publish(sy.stage1 + (size_t)b * NL + j, stgB + j * L::VPL, epoch);
                if (prof) { c2 = clock64(); pc[5] += c2 - c0; }
                if (b % sy.group_size == 0) {  // group leader (CTA-uniform branch)
                    const int first = grp * sy.group_size;
                    const int cnt = G - first < sy.group_size ? G - first : sy.group_size;
                    for (int p = tid; p < cnt * NL; p += BLOCK) {  // every thread polls one line: all round trips overlap
                        const int m = p / NL, j = p % NL;
                        const uint4* line = sy.stage1 + (size_t)(first + m) * NL + j;
                        uint4 w = ld_ll(line);
                        while (w.w != epoch) w = ld_ll(line);
                        L::unpack(w, stgA + (size_t)m * NLV + j * L::VPL);
                    }
                    __syncthreads();
                    if (prof) { c1 = clock64(); pc[6] += c1 - c2; c2 = c1; }
                    for (int idx = tid; idx < NLV; idx += BLOCK) stgA[(size_t)kMaxFan * NLV + idx] = sum_rows<R>(stgA + idx, NLV, cnt);  // CTA order
                    __syncthreads();
                    for (int j = tid; j < NL; j += BLOCK)
                        L::publish(sy.stage2 + ((size_t)par * sy.n_groups + grp) * NL + j, stgA + (size_t)kMaxFan * NLV + j * L::VPL, epoch);
                    if (prof) { c1 = clock64(); pc[7] += c1 - c2; c2 = c1; }
                }
                // hop 2: every CTA collects all group partials (parity double-buffered) and sums in group order
                __syncthreads();  // everybody is done reading stgB (the hop-1 publish)
                for (int p = tid; p < sy.n_groups * NL; p += BLOCK) {
                    const int g2 = p / NL, j = p % NL;
                    const uint4* line = sy.stage2 + ((size_t)par * sy.n_groups + g2) * NL + j;
                    uint4 w = ld_ll(line);
                    while (w.w != epoch) w = ld_ll(line);
                    L::unpack(w, stgB + (size_t)g2 * NLV + j * L::VPL);
                }
                __syncthreads();
                for (int idx = tid; idx < NLV; idx += BLOCK) stgA[idx] = idx < FA ? sum_rows<R>(stgB + idx, NLV, sy.n_groups) : (R)0;
            } else {
                for (int idx = tid; idx < NLV; idx += BLOCK) stgA[idx] = stgB[idx];
            }
            __syncthreads();  // stgA[0..NLV) = this GPU's dW
            if (pe.world > 1) {
                // hop 3 (NVLink): CTA 0 of every GPU writes its dW into every rank's inbox, collects the
                // world's partials from its own inbox, sums them in rank order (=> bit-identical replicas)
                // and broadcasts the total to the local CTAs.  The transfer is part of this kernel.
                constexpr int WPV = sizeof(R) / 4;
                constexpr int FAW = FA * WPV;
                uint32_t* words = reinterpret_cast<uint32_t*>(stgB);  // [world][FAW]
                if (b == 0) {
                    const uint32_t* mine_w = reinterpret_cast<const uint32_t*>(stgA);
                    for (int p = tid; p < pe.world * FAW; p += BLOCK) {
                        const int r = p / FAW, w = p % FAW;
                        st_ll8_sys(pe.inbox[r] + ((size_t)par * pe.world + pe.rank) * FAW + w, mine_w[w], epoch);
                    }
                    for (int p = tid; p < pe.world * FAW; p += BLOCK) {
                        const int r = p / FAW, w = p % FAW;
                        const uint2* slot = pe.inbox[pe.rank] + ((size_t)par * pe.world + r) * FAW + w;
                        uint2 v = ld_ll8_sys(slot);
                        while (v.y != epoch) v = ld_ll8_sys(slot);
                        words[r * FAW + w] = v.x;
                    }
                    __syncthreads();
                    for (int idx = tid; idx < FA; idx += BLOCK) {
                        R tot = (R)0;
                        for (int r = 0; r < pe.world; ++r) tot += reinterpret_cast<const R*>(words + (size_t)r * FAW)[idx];  // rank order
                        stgA[idx] = tot;
                    }
                    __syncthreads();
                    if (G > 1)
                        for (int j = tid; j < NL; j += BLOCK) L::publish(pe.stage3 + (size_t)par * NL + j, stgA + j * L::VPL, epoch);
                } else {
                    for (int j = tid; j < NL; j += BLOCK) {
                        const uint4* line = pe.stage3 + (size_t)par * NL + j;
                        uint4 w = ld_ll(line);
                        while (w.w != epoch) w = ld_ll(line);
                        L::unpack(w, stgA + j * L::VPL);
                    }
                }
                __syncthreads();
            }
            for (int idx = tid; idx < FA; idx += BLOCK) Wsm[(idx / AW) * WS + idx % AW] += stgA[idx];
            if (prof) { c1 = clock64(); pc[3] += c1 - c0; c0 = c1; }
            __syncthreads();
            if (prof) { c1 = clock64(); pc[4] += c1 - c0; c0 = c1; }
        }
    }

    if (prof) {
#pragma unroll
        for (int q = 0; q < 8; ++q) a.phase_prof[b * 8 + q] += pc[q];
    }
    if (resident && active) {
        a.ep_steps[i] = ep;
        a.actions[i] = act;
#pragma unroll
        for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
    }
    if (TRACE && active) {
        R* Z = static_cast<R*>(a.z);
        for (int j = 0; j < FA; ++j) Z[(int64_t)j * N + i] = red[(size_t)j * cap + tid];
    }
    if (SHAREDW && b == 0) {
        for (int j = tid; j < FA; j += BLOCK) static_cast<R*>(a.W)[j] = Wsm[(j / AW) * WS + j % AW];
    }
}

}  // namespace rsrl
