// f4tc.cuh — tcgen05 (5th-gen tensor core) path for the order-7 Fourier basis on the 4-D domains
// (BASELINE config 4: Acrobot / ExpectedSARSA / Fourier(7)+bias, F = 8^4 = 4096, W = 4096 x A), dtype f32.
//
// The feature vector is the real part of a tensor product, phi_(i,j) = Re(u_i * v_j) with u, v complex vectors
// built from the per-dimension tables e_d[c] = exp(i pi c x^_d)  (device.cuh: grid_prepare).  Splitting the four
// digits of the feature index k = ((i0*8 + i1)*8 + i2)*8 + i3 (c_d = 7 - i_d) into a row part and a column part
// turns both hot contractions into dense GEMMs:
//
//   Q = Phi W       i = (i0,i1), j = (i2,i3):   Pr[env,(a,i)] = sum_j vr_j[env] W[(i,j),a]   (and Pi with vi)
//                                               Q_a[env] = sum_i ur_i Pr[env,(a,i)] - ui_i Pi[env,(a,i)]
//                   -> two [128 envs x 64] x [64 x 64A] GEMMs per tile (K = 64), epilogue contracts with u;
//   dW = Phi^T D    m = (i0,i1,b), n = (a,i2',i3), i2 = 4b + i2':
//                                               dW[m,n] = sum_env ur_m (vr_n' d_a) + ui_m (-vi_n' d_a)
//                   -> one [128 x K] x [K x 32A] GEMM with K = 2 x envs, accumulated in TMEM over all envs of a CTA.
//
// Both run as kind::tf32 tcgen05.mma with the operands split x = hi + lo (hi, lo both TF32-representable) and
// three passes hi*hi + lo*hi + hi*lo ("3xTF32"): per-product error ~2^-22, fp32 accumulation — fp32-grade results
// (tools/microbench/umma_tf32.cu measures 1.2e-5 max error on K = 64 sums of O(1) terms vs 7.6e-3 for plain TF32).
// Operand tiles are GENERATED on chip (never read from HBM) by the CTA's threads straight into the canonical
// no-swizzle K-major UMMA layout; the accumulators live in TMEM and are read back with tcgen05.ld.
// Reference semantics: identical to fourier4.cuh / kernels.cuh:env_core (file:line citations there).
#pragma once
#include "fourier4.cuh"

namespace rsrl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major: core matrix = 8 rows x 16 bytes stored as 128 contiguous
// bytes; LBO = byte stride between the two core matrices of one K = 8 step, SBO = byte stride between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a tensor-pipe completion that never arrives raises the engine's fault flag instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* fault) {
    for (int i = 0; i < (1 << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    atomicExch(fault, 1);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ uint32_t tmem_alloc(uint32_t* slot) {  // executed by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    return 0;
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(COLS) : "memory");
}
// 32 lanes x 32 columns: thread `lane` of the warp receives row (lane base + lane), 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// x = hi + lo with hi, lo TF32-representable (round to nearest, ties away): the tensor core then sees exact operands
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - hi);
}

}  // namespace tc

// e_d[c] as (re, im) with the c == 0 => 1 shortcut; tables hold c = 1..7
__device__ __forceinline__ void f4tc_e(const float (&tcs)[4][7], const float (&tsn)[4][7], int d, int c, float& re, float& im) {
    re = c == 0 ? 1.0f : tcs[d][c - 1];
    im = c == 0 ? 0.0f : tsn[d][c - 1];
}

// ---------------------------------------------------------------------------------------------------------------
// env kernel: Q(s_t), behaviour action, Domain::step, Q(s'), TD error -> coef / from_states / actions for the dW pass
// CTA = 128 threads = 128 envs per tile (thread = env = GEMM row = TMEM lane), persistent over tiles.
// ---------------------------------------------------------------------------------------------------------------
template <int AW>
struct F4tcEnvSmem {
    static constexpr int NB = AW * 64;                 // GEMM N: (a, i = (i0,i1))
    static constexpr int B_FLOATS = NB * 64;           // one of {hi, lo}
    static constexpr int UNIT_FLOATS = 128 * 32;       // A unit: 128 rows x 32 k, one of {hi, lo}
    static constexpr size_t bytes = (size_t)(2 * B_FLOATS + 4 * UNIT_FLOATS) * sizeof(float);
};

template <int DOM, bool EXT>
__global__ void __launch_bounds__(128, 1) f4tc_env_kernel(const StepArgs a, const F4Args fa, int n_tiles) {
    using Dom = Domain<DOM>;
    constexpr int D = 4, P = 7, AW = Dom::A;
    using SM = F4tcEnvSmem<AW>;
    constexpr int NB = SM::NB;
    constexpr uint32_t A_LBO = 16 * 128, B_LBO = (NB / 8) * 128, SBO = 128;
    constexpr uint32_t IDESC = tc::make_idesc(128, NB);
    constexpr int TMEM_COLS = 512;  // two accumulators of NB (<= 192) columns
    static_assert(Dom::D == 4, "4-D domains only");

    extern __shared__ __align__(128) unsigned char f4tc_smem[];
    float* Bhi = reinterpret_cast<float*>(f4tc_smem);
    float* Blo = Bhi + SM::B_FLOATS;
    float* Aun = Blo + SM::B_FLOATS;  // unit buffer ub: hi = Aun + ub * 2 * UNIT_FLOATS, lo = hi + UNIT_FLOATS
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    int* fault = &a.counters->pad;

    // ---- one-time setup: W -> B operand (row n = a*64 + i, column j), TMEM, barriers ----
    {
        const float* Wg = static_cast<const float*>(a.W);
        for (int idx = tid; idx < 4096 * AW; idx += 128) {
            const int k = idx / AW, c = idx - k * AW;
            const int i = k >> 6, j = k & 63, n = c * 64 + i;
            float hi, lo;
            tc::split(Wg[idx], hi, lo);
            const int o = (j >> 2) * (B_LBO / 4) + (n >> 3) * 32 + (n & 7) * 4 + (j & 3);
            Bhi[o] = hi;
            Blo[o] = lo;
        }
    }
    if (warp == 0) tc::tmem_alloc<TMEM_COLS>(&tmem_slot);
    if (tid == 0) {
        tc::mbar_init(tc::smem_u32(&bars[0]), 1);
        tc::mbar_init(tc::smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    uint32_t ph0 = 0, ph1 = 0;  // phase parity of bars[0], bars[1] (uniform over the CTA)

    struct Tab { float c[4][P], s[4][P]; };

    // Q(state of tab) for the thread's env; all 128 threads must call (block-level barriers inside)
    auto qeval = [&](const Tab& tb, float* q) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int part = u >> 1, kh = u & 1, ub = u & 1;
            float* Ahi = Aun + ub * 2 * SM::UNIT_FLOATS;
            float* Alo = Ahi + SM::UNIT_FLOATS;
            if (u >= 2) {  // the MMAs of unit u-2 have finished reading this buffer
                if (ub == 0) { tc::mbar_wait(tc::smem_u32(&bars[0]), ph0, fault); ph0 ^= 1; }
                else { tc::mbar_wait(tc::smem_u32(&bars[1]), ph1, fault); ph1 ^= 1; }
            }
            // generate the unit: row = tid, k = (i2 - 4*kh)*8 + i3, value = part of e2[c2] * e3[c3]
#pragma unroll
            for (int i2l = 0; i2l < 4; ++i2l) {
                const int c2 = P - (kh * 4 + i2l);
                float e2r, e2i;
                f4tc_e(tb.c, tb.s, 2, c2, e2r, e2i);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int c3 = P - (g * 4 + x);
                        float e3r, e3i;
                        f4tc_e(tb.c, tb.s, 3, c3, e3r, e3i);
                        const float v = part == 0 ? fmaf(e2r, e3r, -(e2i * e3i)) : fmaf(e2r, e3i, e2i * e3r);
                        tc::split(v, hi[x], lo[x]);
                    }
                    const int kc = i2l * 2 + g;  // 16-byte K chunk
                    const int o = kc * (A_LBO / 4) + (tid >> 3) * 32 + (tid & 7) * 4;
                    *reinterpret_cast<float4*>(Ahi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(Alo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t acc = tmem + (uint32_t)(part * NB);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t abase = tc::smem_u32(pass == 1 ? Alo : Ahi);
                    const uint32_t bbase = tc::smem_u32(pass == 2 ? Blo : Bhi) + (uint32_t)kh * 8u * B_LBO;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc::umma_tf32(acc, tc::make_desc(abase + ks * 2 * A_LBO, A_LBO, SBO), tc::make_desc(bbase + ks * 2 * B_LBO, B_LBO, SBO),
                                      IDESC, (kh | pass | ks) != 0 ? 1u : 0u);
                }
                tc::umma_commit(tc::smem_u32(&bars[ub]));
            }
        }
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = 0.0f;
        // epilogue: q_a = sum_i ur_i Pr[a*64 + i] - ui_i Pi[a*64 + i], i = i0*8 + i1
        auto contract = [&](int part) {
            const uint32_t acc = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(part * NB);
            __syncwarp();  // tcgen05.ld is .sync.aligned: thread 0 rejoins its warp after issuing the MMAs
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float uu[32];
#pragma unroll
                for (int i0l = 0; i0l < 4; ++i0l) {
                    float e0r, e0i;
                    f4tc_e(tb.c, tb.s, 0, P - (h * 4 + i0l), e0r, e0i);
#pragma unroll
                    for (int i1 = 0; i1 < 8; ++i1) {
                        float e1r, e1i;
                        f4tc_e(tb.c, tb.s, 1, P - i1, e1r, e1i);
                        uu[i0l * 8 + i1] = part == 0 ? fmaf(e0r, e1r, -(e0i * e1i)) : -fmaf(e0r, e1i, e0i * e1r);
                    }
                }
#pragma unroll
                for (int c = 0; c < AW; ++c) {
                    float v[32];
                    tc::tmem_ld32(acc + (uint32_t)(c * 64 + h * 32), v);
#pragma unroll
                    for (int x = 0; x < 32; ++x) q[c] = fmaf(uu[x], v[x], q[c]);
                }
            }
        };
        // units 0,1 (the real-part accumulator) were waited for above (u = 3): contract them while units 2,3 run
        tc::fence_after_sync();
        contract(0);
        tc::mbar_wait(tc::smem_u32(&bars[0]), ph0, fault); ph0 ^= 1;  // unit 2 done
        tc::mbar_wait(tc::smem_u32(&bars[1]), ph1, fault); ph1 ^= 1;  // unit 3 done
        tc::fence_after_sync();
        contract(1);
        tc::fence_before_sync();  // the next call's MMAs overwrite the accumulators after its first __syncthreads
    };

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = (int64_t)tile * 128 + tid;
        const bool active = i < a.n;
        const uint64_t g = (uint64_t)(a.env_offset + (active ? i : 0));
        double s[D];
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = active ? (EXT ? a.ext_from[i * D + d] : a.states[i * D + d]) : Dom::start(d);
        if (active) {
#pragma unroll
            for (int d = 0; d < D; ++d) fa.from_states[i * D + d] = s[d];
        }
        Tab tab;
        f4_tables<float, Dom, P, RSRL_FOURIER>(s, tab.c, tab.s);

        // ---- B: behaviour action and Q(s_t, a_t) under W_t (kernels.cuh:env_core) ----
        float q[AW];
        qeval(tab, q);
        bool nonfinite = false;
        int act;
        if (EXT) act = active ? a.ext_actions[i] : 0;
        else act = policy_sample<float, AW>(a.pol, q, g, a.t, STREAM_BEHAVIOUR, nonfinite);
        float qsa = q[0];
#pragma unroll
        for (int c = 0; c < AW; ++c) if (c == act) qsa = q[c];

        // ---- C: Domain::transition ----
        double reward;
        bool terminated;
        if (EXT) {
            if (active) {
#pragma unroll
                for (int d = 0; d < D; ++d) s[d] = a.ext_to[i * D + d];
            }
            reward = active ? a.ext_rewards[i] : 0.0;
            terminated = active ? a.ext_term[i] != 0 : false;
        } else {
            Dom::step(s, act, reward, terminated);
        }

        // ---- D: TD error with W_t (Q(s') is evaluated for every row; terminal rows ignore it) ----
        f4_tables<float, Dom, P, RSRL_FOURIER>(s, tab.c, tab.s);
        float nq[AW];
        qeval(tab, nq);
        float residual;
        if (terminated) {
            residual = (float)reward - qsa;
        } else {
            float target;
            if (a.algo == RSRL_QLEARNING || a.algo == RSRL_Q_LAMBDA) {
                find_max<float, AW>(nq, target);
            } else if (a.algo == RSRL_SARSA || a.algo == RSRL_SARSA_LAMBDA) {
                const int na = policy_sample<float, AW>(a.pol, nq, g, a.t, STREAM_TARGET, nonfinite);
                target = nq[0];
#pragma unroll
                for (int c = 0; c < AW; ++c) if (c == na) target = nq[c];
            } else {
                float p[AW];
                policy_probs<float, AW>(a.pol.policy, (float)a.epsilon, nq, p);
                target = 0.0f;
#pragma unroll
                for (int c = 0; c < AW; ++c) target = target + nq[c] * p[c];
            }
            residual = (float)reward + (float)a.gamma * target - qsa;
        }
        const float coef = a.algo == RSRL_EXPECTED_SARSA ? (float)a.lr_scaled * ((float)a.alpha * residual) : (float)a.lr_scaled * residual;

        if (active) {
            if (a.td) static_cast<float*>(a.td)[i] = residual;
            if (nonfinite) atomicExch(&a.counters->nonfinite, 1);
            static_cast<float*>(fa.coef)[i] = coef;
            a.actions[i] = act;
            if (!EXT) {
                a.ep_steps[i] = env_bookkeeping<Dom>(a, a.t, i, g, s, a.ep_steps[i], terminated);
#pragma unroll
                for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
            }
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_free<TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// dW kernel: partials[cta][k*AW + a] = sum over the CTA's envs of phi_k(s_env) * d_a(env), d_a = coef if a == action.
// CTA = 256 threads; sub-tile = 32 envs; thread = (env lane, q = warp): A rows m = q*16 + i1*2 + b (i0 = q),
// B rows n = a*32 + i2'*8 + q (i3 = q).  K index = (part, env): two units per sub-tile (real, imaginary), double-buffered
// so that generating one unit overlaps the MMAs of the other.  The accumulator stays in TMEM for the whole kernel.
// ---------------------------------------------------------------------------------------------------------------
template <int AW>
struct F4tcDwSmem {
    static constexpr int NB = AW * 32;
    static constexpr uint32_t A_LBO = 16 * 128 + 16;        // +16 B: the 8 K chunks of a warp's scalar stores hit 8 distinct bank groups
    static constexpr uint32_t B_LBO = (NB / 8) * 128 + 16;
    static constexpr int A_FLOATS = 8 * A_LBO / 4;          // one of {hi, lo}: 8 K chunks (32 envs)
    static constexpr int B_FLOATS = 8 * B_LBO / 4;
    static constexpr int UNIT_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    static constexpr int TAB_FLOATS = 4 * 7 * 2 * 32;       // [d][c-1][{cos,sin}][env lane]
    static constexpr size_t bytes = (size_t)(2 * UNIT_FLOATS + TAB_FLOATS + 2 * 32) * sizeof(float);
};

template <int DOM>
__global__ void __launch_bounds__(256, 1) f4tc_dw_kernel(int64_t n, const double* __restrict__ from_states, const float* __restrict__ coef,
                                                         const int32_t* __restrict__ actions, float* __restrict__ partials, Counters* counters) {
    using Dom = Domain<DOM>;
    constexpr int P = 7, AW = Dom::A;
    using SM = F4tcDwSmem<AW>;
    constexpr int NB = SM::NB;
    constexpr uint32_t SBO = 128;
    constexpr uint32_t IDESC = tc::make_idesc(128, NB);
    constexpr int TMEM_COLS = 128;

    extern __shared__ __align__(128) unsigned char f4tc_smem[];
    float* units = reinterpret_cast<float*>(f4tc_smem);
    float* tabs = units + 2 * SM::UNIT_FLOATS;
    float* dco = tabs + SM::TAB_FLOATS;                    // [32] coef
    int* dact = reinterpret_cast<int*>(dco + 32);          // [32] action (-1: padding env)
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, lane = tid & 31, q = tid >> 5;
    int* fault = &counters->pad;

    if (q == 0) tc::tmem_alloc<TMEM_COLS>(&tmem_slot);
    if (tid == 0) {
        tc::mbar_init(tc::smem_u32(&bars[0]), 1);
        tc::mbar_init(tc::smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    const int64_t n_sub = (n + 31) / 32;
    const int64_t s_begin = n_sub * blockIdx.x / gridDim.x, s_end = n_sub * (blockIdx.x + 1) / gridDim.x;
    uint32_t ph[2] = {0, 0};
    bool used[2] = {false, false};
    bool first_mma = true;

    for (int64_t st = s_begin; st < s_end; ++st) {
        const int64_t env = st * 32 + lane;
        // ---- per-env tables: warp d < 4 builds dimension d of env `lane` ----
        if (q < 4) {
            float c1 = 1.0f, s1 = 0.0f;
            if (env < n) {
                const double x = from_states[env * 4 + q];
                const double lo = Dom::lo(q), hi = Dom::hi(q);
                const float xh = (float)dmul(dsub(x, lo), 1.0 / (hi - lo));  // == grid_prepare (f32)
                sincospif(xh, &s1, &c1);
            }
            float cj = c1, sj = s1;
            float* t = tabs + (q * 7) * 64 + lane;
            t[0] = cj; t[32] = sj;
#pragma unroll
            for (int j = 1; j < P; ++j) {
                const float cn = fmaf(cj, c1, -(sj * s1)), sn = fmaf(sj, c1, cj * s1);
                cj = cn; sj = sn;
                t[j * 64] = cj; t[j * 64 + 32] = sj;
            }
        } else if (q == 4) {
            dco[lane] = env < n ? coef[env] : 0.0f;
            dact[lane] = env < n ? actions[env] : -1;
        }
        __syncthreads();
        auto E = [&](int d, int c, float& re, float& im) {  // e_d[c] of env `lane`
            re = c == 0 ? 1.0f : tabs[(d * 7 + c - 1) * 64 + lane];
            im = c == 0 ? 0.0f : tabs[(d * 7 + c - 1) * 64 + 32 + lane];
        };
        // u_m = e0[7-q] * e1[7-i1] * (b == 0 ? e2[4] : 1), m = q*16 + i1*2 + b
        float ur[16], ui[16];
        {
            float e0r, e0i, e24r, e24i;
            E(0, P - q, e0r, e0i);
            E(2, 4, e24r, e24i);
#pragma unroll
            for (int i1 = 0; i1 < 8; ++i1) {
                float e1r, e1i;
                E(1, P - i1, e1r, e1i);
                float tr, ti;
                if (i1 == 7) { tr = e0r; ti = e0i; }
                else { tr = fmaf(e0r, e1r, -(e0i * e1i)); ti = fmaf(e0r, e1i, e0i * e1r); }
                ur[i1 * 2 + 1] = tr; ui[i1 * 2 + 1] = ti;                                            // b = 1: c2 high part 0
                ur[i1 * 2] = fmaf(tr, e24r, -(ti * e24i)); ui[i1 * 2] = fmaf(tr, e24i, ti * e24r);   // b = 0: times e2[4]
            }
        }
        // v_n' = e2[3-i2'] * e3[7-q], n' = i2'*8 + q
        float vr[4], vi[4];
        {
            float e3r, e3i;
            E(3, P - q, e3r, e3i);
#pragma unroll
            for (int i2l = 0; i2l < 4; ++i2l) {
                float e2r, e2i;
                E(2, 3 - i2l, e2r, e2i);
                if (i2l == 3) { vr[i2l] = e3r; vi[i2l] = e3i; }
                else { vr[i2l] = fmaf(e2r, e3r, -(e2i * e3i)); vi[i2l] = fmaf(e2r, e3i, e2i * e3r); }
            }
        }
        const float dc = dco[lane];
        const int act = dact[lane];

#pragma unroll
        for (int part = 0; part < 2; ++part) {
            float* Ahi = units + part * SM::UNIT_FLOATS;
            float* Alo = Ahi + SM::A_FLOATS;
            float* Bhi = Alo + SM::A_FLOATS;
            float* Blo = Bhi + SM::B_FLOATS;
            if (used[part]) { tc::mbar_wait(tc::smem_u32(&bars[part]), ph[part], fault); ph[part] ^= 1; }
            used[part] = true;
            const int ko = (lane >> 2) * (int)(SM::A_LBO / 4) + (lane & 3);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int m = q * 16 + r;
                float hi, lo;
                tc::split(part == 0 ? ur[r] : ui[r], hi, lo);
                const int o = ko + (m >> 3) * 32 + (m & 7) * 4;
                Ahi[o] = hi;
                Alo[o] = lo;
            }
            const int kb = (lane >> 2) * (int)(SM::B_LBO / 4) + (lane & 3);
#pragma unroll
            for (int c = 0; c < AW; ++c)
#pragma unroll
                for (int i2l = 0; i2l < 4; ++i2l) {
                    const int nrow = c * 32 + i2l * 8 + q;
                    float hi = 0.0f, lo = 0.0f;
                    if (c == act) tc::split(part == 0 ? vr[i2l] * dc : -(vi[i2l] * dc), hi, lo);
                    const int o = kb + (nrow >> 3) * 32 + (nrow & 7) * 4;
                    Bhi[o] = hi;
                    Blo[o] = lo;
                }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t abase = tc::smem_u32(pass == 1 ? Alo : Ahi), bbase = tc::smem_u32(pass == 2 ? Blo : Bhi);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc::umma_tf32(tmem, tc::make_desc(abase + ks * 2 * SM::A_LBO, SM::A_LBO, SBO),
                                      tc::make_desc(bbase + ks * 2 * SM::B_LBO, SM::B_LBO, SBO), IDESC, first_mma ? 0u : 1u);
                        first_mma = false;
                    }
                }
                tc::umma_commit(tc::smem_u32(&bars[part]));
            }
        }
        // the next sub-tile's table writes happen after the two __syncthreads above: every thread has read its tables
    }

    // ---- drain + epilogue: TMEM [m = i0*16 + i1*2 + b][n = a*32 + i2'*8 + i3] -> partials[cta][k*AW + a] ----
    float* out = partials + (size_t)blockIdx.x * 4096 * AW;
    if (s_begin >= s_end) {
        for (int j = tid; j < 4096 * AW; j += 256) out[j] = 0.0f;
    } else {
#pragma unroll
        for (int part = 0; part < 2; ++part)
            if (used[part]) { tc::mbar_wait(tc::smem_u32(&bars[part]), ph[part], fault); ph[part] ^= 1; }
        tc::fence_after_sync();
        if (q < 4) {
            __syncwarp();
            const int m = q * 32 + lane, i0 = m >> 4, i1 = (m >> 1) & 7, b = m & 1;
#pragma unroll
            for (int c = 0; c < AW; ++c) {
                float v[32];
                tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
                for (int x = 0; x < 32; ++x) {
                    const int i2 = 4 * b + (x >> 3), i3 = x & 7;
                    const int k = ((i0 * 8 + i1) * 8 + i2) * 8 + i3;
                    out[(size_t)k * AW + c] = v[x];
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (q == 0) tc::tmem_free<TMEM_COLS>(tmem);
}

}  // namespace rsrl
