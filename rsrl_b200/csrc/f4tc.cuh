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
// Both run as kind::tf32 tcgen05.mma with the operands split x = hi + lo (hi = x rounded to TF32, lo = x - hi) and
// three products hi*hi + lo*hi + hi*lo ("3xTF32"): per-product error ~2^-21, fp32 accumulation — fp32-grade results
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
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {  // non-blocking poll
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
// Bounded wait: a tensor-pipe completion that never arrives raises the engine's fault flag instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* fault) {
    for (int i = 0; i < (1 << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    atomicExch(fault, 1);
}
__device__ __forceinline__ void group_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
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
// 32 lanes x 32 columns: thread `lane` of the warp receives row (lane base + lane), 32 consecutive fp32 columns.
// tmem_ld32_issue only issues the load; the registers are valid after tmem_ld_wait() + tmem_ld_fence(r).
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// pins the uses of r[] behind the wait (the compiler may not hoist arithmetic on them above this statement)
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&r)[32]) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
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

// x = hi + lo: hi = x rounded to TF32 (10 mantissa bits, half-up in magnitude: two integer ops), lo = x - hi exactly
// (|lo| <= 2^-11 |x|; the tensor core drops lo's bits below 2^-10 |lo|, i.e. 2^-21 |x| at worst).
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}

}  // namespace tc

// e_d[c] as (re, im) with the c == 0 => 1 shortcut; tables hold c = 1..7
__device__ __forceinline__ void f4tc_e(const float (&tcs)[4][7], const float (&tsn)[4][7], int d, int c, float& re, float& im) {
    re = c == 0 ? 1.0f : tcs[d][c - 1];
    im = c == 0 ? 0.0f : tsn[d][c - 1];
}
// ---------------------------------------------------------------------------------------------------------------
// One batched step = three kernels + the dW pass:
//   f4tc_q_kernel<PHASE 0>  Q(s_t; W_t) for every env (tensor cores), tables of s_t for the dW pass
//   f4tc_phys_kernel        behaviour action, Q(s_t)[a_t], Domain::step (f64 RK4) at full occupancy
//   f4tc_q_kernel<PHASE 1>  Q(s'; W_t) (tensor cores), TD error -> coef, episode bookkeeping
// (The f64 physics used to sit between the two evaluations inside one kernel: 30 % of that kernel's time at 8 warps/SM
// with the tensor pipe idle; on its own it runs at 64 warps/SM.)
// f4tc_q_kernel: CTA = 288 threads = two independent groups of 128 + one MMA-issuer warp; a group owns one tile of
// 128 envs at a time (thread = env = GEMM row = TMEM lane), its own 32 KB operand buffer, 192 TMEM columns and a
// full / done mbarrier pair; both groups share the W operand.  A group never blocks on MMA issue: it publishes a unit
// (arrive on `full`), the issuer thread queues the unit's 12 MMAs and commits them to `done`.  While one group generates
// operands / contracts its accumulator on the CUDA cores the other group's MMAs run.
// ---------------------------------------------------------------------------------------------------------------
template <int AW>
struct F4tcEnvSmem {
    static constexpr int NB = AW * 64;                 // GEMM N: (a, i = (i0,i1))
    static constexpr int B_FLOATS = NB * 64;           // one of {hi, lo}
    static constexpr int UNIT_FLOATS = 128 * 16;       // A unit: 128 rows x 16 k (a quarter of K), one of {hi, lo}
    static constexpr size_t bytes = (size_t)(2 * B_FLOATS + 8 * UNIT_FLOATS) * sizeof(float);  // 2 groups x 2 buffers x {hi, lo}
};

template <int DOM, int PHASE, bool EXT>
__global__ void __launch_bounds__(288, 1) f4tc_q_kernel(const StepArgs a, const F4Args fa, int n_tiles) {
    using Dom = Domain<DOM>;
    constexpr int D = 4, P = 7, AW = Dom::A;
    using SM = F4tcEnvSmem<AW>;
    constexpr int NB = SM::NB;
    constexpr uint32_t A_LBO = 16 * 128, B_LBO = (NB / 8) * 128, SBO = 128;
    constexpr uint32_t IDESC = tc::make_idesc(128, NB);
    constexpr int TMEM_COLS = 512;  // group g accumulates in columns [256 g, 256 g + NB)
    static_assert(Dom::D == 4, "4-D domains only");

    extern __shared__ __align__(128) unsigned char f4tc_smem[];
    float* Bhi = reinterpret_cast<float*>(f4tc_smem);
    float* Blo = Bhi + SM::B_FLOATS;
    __shared__ __align__(8) unsigned long long bars[2][2];   // done[g][b]: the MMAs of the unit in buffer b of group g completed
    __shared__ __align__(8) unsigned long long fulls[2][2];  // full[g][b]: all 128 threads of group g stored their rows of buffer b
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, grp = (tid >> 7) & 1, gt = tid & 127, gwarp = gt >> 5;
    const bool issuer_warp = tid >= 256;
    float* Abase = Blo + SM::B_FLOATS;                        // [group][buffer][{hi, lo}][UNIT_FLOATS]
    float* Agrp = Abase + grp * 4 * SM::UNIT_FLOATS;
    int* fault = &a.counters->pad;

    // ---- one-time setup: W -> B operand (row n = a*64 + i, column j), TMEM, barriers ----
    {
        const float* Wg = static_cast<const float*>(a.W);
        for (int idx = tid; idx < 4096 * AW; idx += 288) {
            const int k = idx / AW, c = idx - k * AW;
            const int i = k >> 6, j = k & 63, n = c * 64 + i;
            float hi, lo;
            tc::split(Wg[idx], hi, lo);
            const int o = (j >> 2) * (B_LBO / 4) + (n >> 3) * 32 + (n & 7) * 4 + (j & 3);
            Bhi[o] = hi;
            Blo[o] = lo;
        }
    }
    if (tid < 32) tc::tmem_alloc<TMEM_COLS>(&tmem_slot);
    if (tid == 0) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            tc::mbar_init(tc::smem_u32(&bars[x >> 1][x & 1]), 1);
            tc::mbar_init(tc::smem_u32(&fulls[x >> 1][x & 1]), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    if (issuer_warp) {
        // ---- MMA issuer: serves whichever group has published a unit (units of a group alternate between its two buffers);
        //      3xTF32 = hi*hi + lo*hi + hi*lo over the unit's 2 K steps ----
        if (tid == 256) {
            int units[2], cnt[2] = {0, 0};
            uint32_t phf[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
                const int first = blockIdx.x * 2 + g2;
                units[g2] = first < n_tiles ? ((n_tiles - first - 1) / ((int)gridDim.x * 2) + 1) * 8 : 0;  // 2 parts x 4 K quarters per tile
            }
            // descriptors are loop invariant: built once (a thread that rebuilds them per MMA is issue bound, tools/microbench)
            uint64_t adesc[2][2][2][2], bdesc[2][4][2];  // A: [group][buffer][hi, lo][K step]; B: [hi, lo][K quarter][K step]
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            adesc[g2][b][hl][ks] = tc::make_desc(tc::smem_u32(Abase + (g2 * 4 + b * 2 + hl) * SM::UNIT_FLOATS) + ks * 2 * A_LBO, A_LBO, SBO);
#pragma unroll
            for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                for (int kq = 0; kq < 4; ++kq)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        bdesc[hl][kq][ks] = tc::make_desc(tc::smem_u32(hl ? Blo : Bhi) + (uint32_t)(kq * 4 + ks * 2) * B_LBO, B_LBO, SBO);
            int spins = 0;
            long long t_issue = 0, t_idle = 0, t_last = clock64();
            while (cnt[0] < units[0] || cnt[1] < units[1]) {
                bool progressed = false;
#pragma unroll
                for (int g2 = 0; g2 < 2; ++g2) {
                    const int b = cnt[g2] & 1;
                    if (cnt[g2] < units[g2] && tc::mbar_test_wait(tc::smem_u32(&fulls[g2][b]), b ? phf[g2][1] : phf[g2][0])) {
                        if (b) phf[g2][1] ^= 1; else phf[g2][0] ^= 1;
                        { const long long now = clock64(); t_idle += now - t_last; t_last = now; }
                        tc::fence_after_sync();
                        const int kq = cnt[g2] & 3;
                        const uint32_t acc2 = tmem_slot + (uint32_t)(g2 * 256);
                        // 3xTF32: hi*hi + lo*hi + hi*lo; kq, b are runtime: select among the precomputed descriptors
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const int ahl = pass == 1 ? 1 : 0, bhl = pass == 2 ? 1 : 0;
                                const uint64_t ad = b ? adesc[g2][1][ahl][ks] : adesc[g2][0][ahl][ks];
                                const uint64_t b01 = (kq & 1) ? bdesc[bhl][1][ks] : bdesc[bhl][0][ks];
                                const uint64_t b23 = (kq & 1) ? bdesc[bhl][3][ks] : bdesc[bhl][2][ks];
                                tc::umma_tf32(acc2, ad, (kq & 2) ? b23 : b01, IDESC, (kq | pass | ks) != 0 ? 1u : 0u);
                            }
                        }
                        tc::umma_commit(tc::smem_u32(&bars[g2][b]));
                        { const long long now = clock64(); t_issue += now - t_last; t_last = now; }
                        cnt[g2] += 1;
                        progressed = true;
                        spins = 0;
                    }
                }
                if (!progressed) {
                    __nanosleep(50);  // the poll loop shares a scheduler with two group warps: do not steal their issue slots
                    if (++spins > (1 << 24)) { atomicExch(fault, 1); break; }
                }
            }
            if (a.phase_prof) {  // issuer view: cycles blocked in MMA issue (queue full = tensor pipe busy) / waiting for a unit
                a.phase_prof[(size_t)blockIdx.x * 8 + 6] += t_issue;
                a.phase_prof[(size_t)blockIdx.x * 8 + 7] += t_idle;
            }
        }
    } else {
    const uint32_t acc = tmem_slot + (uint32_t)(grp * 256);
    const uint32_t acc_lane = acc + ((uint32_t)(gwarp * 32) << 16);
    const uint32_t bar0 = tc::smem_u32(&bars[grp][0]), bar1 = tc::smem_u32(&bars[grp][1]);
    const uint32_t full0 = tc::smem_u32(&fulls[grp][0]), full1 = tc::smem_u32(&fulls[grp][1]);
    uint32_t ph0 = 0, ph1 = 0;   // phase parity of the group's done barriers (uniform over the group)
    int pend0 = 0, pend1 = 0;    // units published into buffer b whose completion has not been waited for yet (0 or 1)  // phase parity of the group's barrier (uniform over the group)
    const int a_row = (gt >> 3) * 32 + (gt & 7) * 4;  // float offset of the thread's row inside a 16-byte K chunk column

    float tcs[4][P], tsn[4][P];  // tables of the state being evaluated
    // optional phase profile (RSRL_B200_PHASE_PROFILE=1): cycles of thread 0 per phase, summed over its tiles
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool profiling = a.phase_prof != nullptr && tid == 0;
    long long tprev = profiling ? clock64() : 0;
    auto mark = [&](int phase) {
        if (profiling) { const long long now = clock64(); prof[phase] += now - tprev; tprev = now; }
    };

    // Q of the state whose tables are in tcs/tsn; every thread of the group must call.  Rolled loops over the four operand
    // units (part, kh) keep a single copy of every phase in the instruction stream: unit (part, 1) is generated while the
    // MMAs of unit (part, 0) run; the accumulator is contracted with u after each part (real, then imaginary).
    auto qeval = [&](float* q) {
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = 0.0f;
        auto unit_compute = [&](int part, int kq, float (&hi)[16], float (&lo)[16]) {
            // v = e2[c2] * e3[c3]: real part e2r*e3r - e2i*e3i, imaginary part e2r*e3i + e2i*e3r = e2r*P + e2i*Q
            float Pv[8], Qv[8];
#pragma unroll
            for (int i3 = 0; i3 < 8; ++i3) {
                float e3r, e3i;
                f4tc_e(tcs, tsn, 3, P - i3, e3r, e3i);
                Pv[i3] = part ? e3i : e3r;
                Qv[i3] = part ? e3r : -e3i;
            }
#pragma unroll
            for (int i2l = 0; i2l < 2; ++i2l) {  // i2 = 2 kq + i2l, c2 = 7 - i2: warp-uniform select among the four quarters
                float e2r, e2i, t0r, t0i, t1r, t1i, t2r, t2i, t3r, t3i;
                f4tc_e(tcs, tsn, 2, P - i2l, t0r, t0i);
                f4tc_e(tcs, tsn, 2, P - 2 - i2l, t1r, t1i);
                f4tc_e(tcs, tsn, 2, P - 4 - i2l, t2r, t2i);
                f4tc_e(tcs, tsn, 2, P - 6 - i2l, t3r, t3i);
                e2r = kq == 0 ? t0r : kq == 1 ? t1r : kq == 2 ? t2r : t3r;
                e2i = kq == 0 ? t0i : kq == 1 ? t1i : kq == 2 ? t2i : t3i;
#pragma unroll
                for (int i3 = 0; i3 < 8; ++i3) tc::split(fmaf(e2r, Pv[i3], e2i * Qv[i3]), hi[i2l * 8 + i3], lo[i2l * 8 + i3]);
            }
        };
        auto wait_buf = [&](int b) {  // the MMAs of the last unit published into buffer b completed (no-op if already waited)
            if (b == 0) { if (pend0) { tc::mbar_wait(bar0, ph0, fault); ph0 ^= 1; pend0 = 0; } }
            else { if (pend1) { tc::mbar_wait(bar1, ph1, fault); ph1 ^= 1; pend1 = 0; } }
        };
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
#pragma unroll 1
            for (int kq = 0; kq < 4; ++kq) {
                const int b = kq & 1;
                float hi[16], lo[16];
                unit_compute(part, kq, hi, lo);  // overlaps the MMAs of the previous units
                mark(1);
                wait_buf(b);                     // buffer b free again
                mark(2);
                float* Ahi = Agrp + b * 2 * SM::UNIT_FLOATS;
                float* Alo = Ahi + SM::UNIT_FLOATS;
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) {
                    const int o = kc * (A_LBO / 4) + a_row;
                    *reinterpret_cast<float4*>(Ahi + o) = make_float4(hi[kc * 4], hi[kc * 4 + 1], hi[kc * 4 + 2], hi[kc * 4 + 3]);
                    *reinterpret_cast<float4*>(Alo + o) = make_float4(lo[kc * 4], lo[kc * 4 + 1], lo[kc * 4 + 2], lo[kc * 4 + 3]);
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                tc::mbar_arrive(b ? full1 : full0);  // publish the unit: the issuer warp queues its MMAs and commits them to done[b]
                if (b) pend1 = 1; else pend0 = 1;
                mark(4);
            }
            wait_buf(0);
            wait_buf(1);  // all four units of this part done: accumulator complete
            mark(2);
            {
            // q_a += sum_i uu_i P[a*64 + i], uu = Re(e0 e1) (real part) or -Im(e0 e1) (imaginary part) = e0r*Pu + e0i*Qu
            const int cpart = part;
            __syncwarp();  // tcgen05.ld is .sync.aligned: the issuing thread rejoins its warp
            tc::fence_after_sync();
            float Pu[8], Qu[8];
#pragma unroll
            for (int i1 = 0; i1 < 8; ++i1) {
                float e1r, e1i;
                f4tc_e(tcs, tsn, 1, P - i1, e1r, e1i);
                Pu[i1] = cpart ? -e1i : e1r;
                Qu[i1] = cpart ? -e1r : -e1i;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[2][32];  // two TMEM load buffers: the load of action c+1 is in flight during the FMAs of action c
                tc::tmem_ld32_issue(acc_lane + (uint32_t)(h * 32), r[0]);
                float uu[32];
#pragma unroll
                for (int i0l = 0; i0l < 4; ++i0l) {
                    float e0r, e0i;
                    f4tc_e(tcs, tsn, 0, P - (h * 4 + i0l), e0r, e0i);
#pragma unroll
                    for (int i1 = 0; i1 < 8; ++i1) uu[i0l * 8 + i1] = fmaf(e0r, Pu[i1], e0i * Qu[i1]);
                }
#pragma unroll
                for (int c = 0; c < AW; ++c) {
                    tc::tmem_ld_wait();
                    tc::tmem_ld_fence(r[c & 1]);
                    if (c + 1 < AW) tc::tmem_ld32_issue(acc_lane + (uint32_t)((c + 1) * 64 + h * 32), r[(c + 1) & 1]);
                    float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;  // four chains per action
#pragma unroll
                    for (int x = 0; x < 32; x += 4) {
                        q0 = fmaf(uu[x], __uint_as_float(r[c & 1][x]), q0);
                        q1 = fmaf(uu[x + 1], __uint_as_float(r[c & 1][x + 1]), q1);
                        q2 = fmaf(uu[x + 2], __uint_as_float(r[c & 1][x + 2]), q2);
                        q3 = fmaf(uu[x + 3], __uint_as_float(r[c & 1][x + 3]), q3);
                    }
                    q[c] += (q0 + q1) + (q2 + q3);
                }
            }
            tc::fence_before_sync();  // orders these TMEM reads before the next MMAs into the accumulator
            }
            mark(3);
        }
    };

    for (int tile = blockIdx.x * 2 + grp; tile < n_tiles; tile += gridDim.x * 2) {
        const int64_t i = (int64_t)tile * 128 + gt;
        const bool active = i < a.n;
        const uint64_t g = (uint64_t)(a.env_offset + (active ? i : 0));
        double s[D];
        const double* src = PHASE == 0 ? (EXT ? a.ext_from : a.states) : fa.next_states;
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = active ? src[i * D + d] : Dom::start(d);
        f4_tables<float, Dom, P, RSRL_FOURIER>(s, tcs, tsn);
        if (PHASE == 0 && active) {  // tables of s_t for the dW pass: [d][c-1][{cos, sin}][env], env fastest (coalesced)
            float* tb = fa.tabs + i;
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    tb[(size_t)((d * P + j) * 2) * a.n] = tcs[d][j];
                    tb[(size_t)((d * P + j) * 2 + 1) * a.n] = tsn[d][j];
                }
        }
        float q[AW];
        mark(0);
        qeval(q);
        if (PHASE == 0) {
            if (active) {
#pragma unroll
                for (int c = 0; c < AW; ++c) fa.q[i * 4 + c] = q[c];
            }
            mark(5);
            continue;
        }
        if (active) {
            // ---- D: TD error with W_t (kernels.cuh:env_core); Q(s') was evaluated for every row, terminal rows ignore it ----
            const float4 aux = reinterpret_cast<const float4*>(fa.aux)[i];   // {Q(s_t)[a_t], reward, terminal, -} from the physics kernel
            const float qsa = aux.x, reward = aux.y, q_astar = aux.w;
            const int tz = (int)aux.z, a_star = tz >> 1;
            const bool terminated = (tz & 1) != 0;
            bool nonfinite = false;
            float residual;
            if (terminated) {
                residual = reward - qsa;
            } else {
                float target;
                if (a.algo == RSRL_QLEARNING || a.algo == RSRL_Q_LAMBDA) {
                    find_max<float, AW>(q, target);
                } else if (a.algo == RSRL_SARSA || a.algo == RSRL_SARSA_LAMBDA) {
                    const int na = policy_sample<float, AW>(a.pol, q, g, a.t, STREAM_TARGET, nonfinite);
                    target = q[0];
#pragma unroll
                    for (int c = 0; c < AW; ++c) if (c == na) target = q[c];
                } else if (a.algo == RSRL_PAL) {  // pal.rs:44-52 (q = Q(s'))
                    const int na_star = argmax_first<float, AW>(q), act = a.actions[i];
                    float nq_astar = q[0], nq_nastar = q[0], nq_act = q[0];
#pragma unroll
                    for (int c = 0; c < AW; ++c) {
                        if (c == a_star) nq_astar = q[c];
                        if (c == na_star) nq_nastar = q[c];
                        if (c == act) nq_act = q[c];
                    }
                    const float td_error = reward + (float)a.gamma * nq_astar - qsa;
                    target = fmaxf(td_error - (float)a.alpha * (q_astar - qsa), td_error - (float)a.alpha * (nq_nastar - nq_act));
                } else {
                    float p[AW];
                    policy_probs<float, AW>(a.pol.policy, (float)a.epsilon, q, p);
                    target = 0.0f;
#pragma unroll
                    for (int c = 0; c < AW; ++c) target = target + q[c] * p[c];
                }
                residual = a.algo == RSRL_PAL ? target : reward + (float)a.gamma * target - qsa;
            }
            const float coef = (a.algo == RSRL_EXPECTED_SARSA || a.algo == RSRL_PAL) ? (float)a.lr_scaled * ((float)a.alpha * residual) : (float)a.lr_scaled * residual;
            if (a.td) static_cast<float*>(a.td)[i] = residual;
            if (nonfinite) atomicExch(&a.counters->nonfinite, 1);
            static_cast<float*>(fa.coef)[i] = coef;
            if (!EXT) {
                a.ep_steps[i] = env_bookkeeping<Dom>(a, a.t, i, g, s, a.ep_steps[i], terminated);
#pragma unroll
                for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
            }
        }
        mark(5);
    }

    if (profiling) {
#pragma unroll
        for (int x = 0; x < 6; ++x) a.phase_prof[(size_t)blockIdx.x * 8 + x] += prof[x];
    }
    }  // group threads
    tc::fence_before_sync();
    __syncthreads();
    if (tid < 32) tc::tmem_free<TMEM_COLS>(tmem_slot);
}

// behaviour action + Domain::transition for every env (phases B and C of kernels.cuh:env_core), one thread per env
template <int DOM, bool EXT>
__global__ void __launch_bounds__(128) f4tc_phys_kernel(const StepArgs a, const F4Args fa) {
    using Dom = Domain<DOM>;
    constexpr int D = 4, AW = Dom::A;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= a.n) return;
    const uint64_t g = (uint64_t)(a.env_offset + i);
    float q[AW];
#pragma unroll
    for (int c = 0; c < AW; ++c) q[c] = fa.q[i * 4 + c];
    bool nonfinite = false, terminated;
    int act;
    double reward, s[D];
    if (EXT) {
        act = a.ext_actions[i];
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = a.ext_to[i * D + d];
        reward = a.ext_rewards[i];
        terminated = a.ext_term[i] != 0;
    } else {
        act = policy_sample<float, AW>(a.pol, q, g, a.t, STREAM_BEHAVIOUR, nonfinite);
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = a.states[i * D + d];
        Dom::step(s, act, reward, terminated);
    }
    float qsa = q[0];
#pragma unroll
    for (int c = 0; c < AW; ++c) if (c == act) qsa = q[c];
#pragma unroll
    for (int d = 0; d < D; ++d) fa.next_states[i * D + d] = s[d];
    // PAL (pal.rs:44-50) also needs a* = argmax_first Q(s_t) and Q(s_t)[a*]: packed as terminal + 2 a*, Q(s_t)[a*]
    const int a_star = argmax_first<float, AW>(q);
    float q_astar = q[0];
#pragma unroll
    for (int c = 0; c < AW; ++c) if (c == a_star) q_astar = q[c];
    reinterpret_cast<float4*>(fa.aux)[i] = make_float4(qsa, (float)reward, (terminated ? 1.0f : 0.0f) + 2.0f * (float)a_star, q_astar);
    a.actions[i] = act;  // (EXT: the caller's action, consumed by the dW pass)
    if (nonfinite) atomicExch(&a.counters->nonfinite, 1);
}

// ---------------------------------------------------------------------------------------------------------------
// dW kernel: partials[cta][k*AW + a] = sum over the CTA's envs of phi_k(s_env) * d_a(env), d_a = coef if a == action.
// CTA = 16 generator warps + 1 MMA-issuer warp; sub-tile = 32 envs; generator thread = (env lane, q = warp 0..15):
// A rows m = i0*16 + i1*2 + b with i0 = q/2, i1 in 4(q%2)..+3; B rows n = a*32 + i2'*8 + i3 with i3 = q/2, i2' in
// 2(q%2)..+1.  K index = (part, env): two units per sub-tile (real, imaginary), double-buffered: the generators publish
// a unit (full barrier), the issuer queues its MMAs and commits them (done barrier) while the next unit is generated.
// 3xTF32 in two MMAs per K step: A_hi x [B_hi; B_lo] (N = 2 NB: hi*hi and hi*lo land in separate column ranges) and
// A_lo x B_hi (N = NB) — an SS-mode M = 128 MMA reads 4 KB of A from shared memory whatever N is (56 cycles at N = 96
// against a 48-cycle tensor floor, tools/microbench), so folding the third pass into a wider second operand saves a
// quarter of the tensor time and a third of the A traffic.  The accumulator stays in TMEM
// for the whole kernel; the two column ranges are added in the epilogue.
// ---------------------------------------------------------------------------------------------------------------
template <int AW>
struct F4tcDwSmem {
    static constexpr int NB = AW * 32;
    static constexpr uint32_t A_LBO = 16 * 128 + 16;        // +16 B: the 8 K chunks of a warp's scalar stores hit 8 distinct bank groups
    static constexpr uint32_t B_LBO = (2 * NB / 8) * 128 + 16;  // B tile rows: [0, NB) = hi part, [NB, 2 NB) = lo part
    static constexpr int A_FLOATS = 8 * A_LBO / 4;          // one of {hi, lo}: 8 K chunks (32 envs)
    static constexpr int B_FLOATS = 8 * B_LBO / 4;          // hi and lo rows together
    static constexpr int UNIT_FLOATS = 2 * A_FLOATS + B_FLOATS;
    static constexpr size_t bytes = (size_t)(2 * UNIT_FLOATS) * sizeof(float);
};

template <int DOM>
__global__ void __launch_bounds__(544, 1) f4tc_dw_kernel(int64_t n, const float* __restrict__ tabs_g, const float* __restrict__ coef,
                                                         const int32_t* __restrict__ actions, float* __restrict__ partials, Counters* counters,
                                                         long long* phase_prof) {
    using Dom = Domain<DOM>;
    constexpr int P = 7, AW = Dom::A;
    using SM = F4tcDwSmem<AW>;
    constexpr int NB = SM::NB;
    constexpr uint32_t SBO = 128;
    constexpr uint32_t IDESC_WIDE = tc::make_idesc(128, 2 * NB), IDESC_HALF = tc::make_idesc(128, NB);
    constexpr int TMEM_COLS = 256;  // 2 NB <= 192 accumulator columns

    extern __shared__ __align__(128) unsigned char f4tc_smem[];
    float* units = reinterpret_cast<float*>(f4tc_smem);
    __shared__ __align__(8) unsigned long long bars[2];   // done[part]: the MMAs of the unit in buffer `part` completed
    __shared__ __align__(8) unsigned long long fulls[2];  // full[part]: all 512 generator threads stored their elements
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, lane = tid & 31, q = tid >> 5, qi = q >> 1, hq = q & 1;
    int* fault = &counters->pad;

    if (q == 0) tc::tmem_alloc<TMEM_COLS>(&tmem_slot);
    if (tid == 0) {
        tc::mbar_init(tc::smem_u32(&bars[0]), 1);
        tc::mbar_init(tc::smem_u32(&bars[1]), 1);
        tc::mbar_init(tc::smem_u32(&fulls[0]), 512);
        tc::mbar_init(tc::smem_u32(&fulls[1]), 512);
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

    if (q == 16) {
        // ---- MMA issuer ----
        if (lane == 0) {
            uint32_t phf[2] = {0, 0};
            bool first_mma = true;
            for (int64_t st = s_begin; st < s_end; ++st) {
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    const float* Ahi = units + part * SM::UNIT_FLOATS;
                    const float* Alo = Ahi + SM::A_FLOATS;
                    const float* Bt = Alo + SM::A_FLOATS;
                    tc::mbar_wait(tc::smem_u32(&fulls[part]), phf[part], fault);
                    phf[part] ^= 1;
                    tc::fence_after_sync();
                    const uint32_t bbase = tc::smem_u32(Bt);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t bd = tc::make_desc(bbase + ks * 2 * SM::B_LBO, SM::B_LBO, SBO);
                        tc::umma_tf32(tmem, tc::make_desc(tc::smem_u32(Ahi) + ks * 2 * SM::A_LBO, SM::A_LBO, SBO), bd, IDESC_WIDE, first_mma ? 0u : 1u);
                        tc::umma_tf32(tmem, tc::make_desc(tc::smem_u32(Alo) + ks * 2 * SM::A_LBO, SM::A_LBO, SBO), bd, IDESC_HALF, 1u);
                        first_mma = false;
                    }
                    tc::umma_commit(tc::smem_u32(&bars[part]));
                }
            }
        }
    } else {
    // float offsets of this thread's element (k = lane) in its rows
    const int ka = (lane >> 2) * (int)(SM::A_LBO / 4) + (lane & 3);
    const int kb = (lane >> 2) * (int)(SM::B_LBO / 4) + (lane & 3);
    const int m0 = qi * 16 + hq * 8;  // rows m0 .. m0+7 = (i1 = 4 hq + r/2, b = r%2): one 8-row group
    const int oa = ka + (m0 >> 3) * 32;

    // Inputs of one sub-tile for this thread: the table entries e_d[c] of env `lane` it needs (written by the env kernel,
    // [d][c-1][{cos,sin}][env], env fastest => coalesced), the scaled TD error and the action.  Loaded one sub-tile ahead.
    struct In { float e0r, e0i, e24r, e24i, e3r, e3i, e1r[4], e1i[4], e2r[2], e2i[2], dc; int act; };
    auto load_in = [&](int64_t st, In& in) {
        const int64_t env = st * 32 + lane;
        const bool ok = st < s_end && env < n;
        auto E = [&](int d, int c, float& re, float& im) {
            re = c == 0 ? 1.0f : (ok ? tabs_g[(size_t)((d * 7 + c - 1) * 2) * n + env] : 0.0f);
            im = c == 0 ? 0.0f : (ok ? tabs_g[(size_t)((d * 7 + c - 1) * 2 + 1) * n + env] : 0.0f);
        };
        E(0, P - qi, in.e0r, in.e0i);
        E(2, 4, in.e24r, in.e24i);
        E(3, P - qi, in.e3r, in.e3i);
#pragma unroll
        for (int i1l = 0; i1l < 4; ++i1l) E(1, P - (hq * 4 + i1l), in.e1r[i1l], in.e1i[i1l]);
#pragma unroll
        for (int x = 0; x < 2; ++x) E(2, 3 - (hq * 2 + x), in.e2r[x], in.e2i[x]);
        in.dc = ok ? coef[env] : 0.0f;
        in.act = ok ? actions[env] : -1;
    };
    In cur, nxt;
    load_in(s_begin, cur);
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool profiling = phase_prof != nullptr && tid == 0;
    long long tprev = profiling ? clock64() : 0;
    auto mark = [&](int phase) {
        if (profiling) { const long long now = clock64(); prof[phase] += now - tprev; tprev = now; }
    };

    for (int64_t st = s_begin; st < s_end; ++st) {
        load_in(st + 1, nxt);
        // u_m = e0[7-i0] * e1[7-i1] * (b == 0 ? e2[4] : 1), rows r = (i1 - 4 hq)*2 + b
        float ur[8], ui[8];
#pragma unroll
        for (int i1l = 0; i1l < 4; ++i1l) {
            const float tr = fmaf(cur.e0r, cur.e1r[i1l], -(cur.e0i * cur.e1i[i1l])), ti = fmaf(cur.e0r, cur.e1i[i1l], cur.e0i * cur.e1r[i1l]);
            ur[i1l * 2 + 1] = tr; ui[i1l * 2 + 1] = ti;                                                            // b = 1: c2 high part 0
            ur[i1l * 2] = fmaf(tr, cur.e24r, -(ti * cur.e24i)); ui[i1l * 2] = fmaf(tr, cur.e24i, ti * cur.e24r);   // b = 0: times e2[4]
        }
        // v_n' = e2[3-i2'] * e3[7-i3], i3 = qi, i2' = 2 hq + x
        float vr[2], vi[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            vr[x] = fmaf(cur.e2r[x], cur.e3r, -(cur.e2i[x] * cur.e3i));
            vi[x] = fmaf(cur.e2r[x], cur.e3i, cur.e2i[x] * cur.e3r);
        }
        const float dc = cur.dc;
        const int act = cur.act;
        mark(0);

#pragma unroll
        for (int part = 0; part < 2; ++part) {
            float* Ahi = units + part * SM::UNIT_FLOATS;
            float* Alo = Ahi + SM::A_FLOATS;
            float* Bhi = Alo + SM::A_FLOATS;   // rows [0, NB)
            float* Blo = Bhi + (NB / 8) * 32;  // rows [NB, 2 NB) of the same tile
            if (used[part]) { tc::mbar_wait(tc::smem_u32(&bars[part]), ph[part], fault); ph[part] ^= 1; }
            used[part] = true;
            mark(1);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float hi, lo;
                tc::split(part == 0 ? ur[r] : ui[r], hi, lo);
                Ahi[oa + r * 4] = hi;
                Alo[oa + r * 4] = lo;
            }
#pragma unroll
            for (int c = 0; c < AW; ++c)
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    const int nrow = c * 32 + (hq * 2 + x) * 8 + qi;
                    float hi = 0.0f, lo = 0.0f;
                    if (c == act) tc::split(part == 0 ? vr[x] * dc : -(vi[x] * dc), hi, lo);
                    const int o = kb + (nrow >> 3) * 32 + (nrow & 7) * 4;
                    Bhi[o] = hi;
                    Blo[o] = lo;
                }
            mark(2);
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive(tc::smem_u32(&fulls[part]));  // publish: the issuer warp queues this unit's MMAs
            mark(4);
        }
        cur = nxt;
    }
    if (profiling) {
#pragma unroll
        for (int x = 0; x < 8; ++x) phase_prof[(size_t)blockIdx.x * 8 + x] += prof[x];
    }
    }  // generator warps

    // ---- drain + epilogue: TMEM [m = i0*16 + i1*2 + b][n = a*32 + i2'*8 + i3] -> partials[cta][k*AW + a] ----
    float* out = partials + (size_t)blockIdx.x * 4096 * AW;
    if (s_begin >= s_end) {
        for (int j = tid; j < 4096 * AW; j += 544) out[j] = 0.0f;
    } else if (q < 16) {
#pragma unroll
        for (int part = 0; part < 2; ++part)
            if (used[part]) { tc::mbar_wait(tc::smem_u32(&bars[part]), ph[part], fault); ph[part] ^= 1; }
        tc::fence_after_sync();
        if (q < 4) {
            __syncwarp();
            const int m = q * 32 + lane, i0 = m >> 4, i1 = (m >> 1) & 7, b = m & 1;
#pragma unroll
            for (int c = 0; c < AW; ++c) {
                float v[32], w[32];  // hi*hi + lo*hi columns, hi*lo columns
                tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
                tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(NB + c * 32), w);
#pragma unroll
                for (int x = 0; x < 32; ++x) {
                    const int i2 = 4 * b + (x >> 3), i3 = x & 7;
                    const int k = ((i0 * 8 + i1) * 8 + i2) * 8 + i3;
                    out[(size_t)k * AW + c] = v[x] + w[x];
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (q == 0) tc::tmem_free<TMEM_COLS>(tmem);
}

// Fixed-order sum of the per-CTA partials: block = 32 warps x 32 consecutive weights; warp w adds partials w, w+32, ...
// in ascending order (<= 5 loads per thread for 148 partials: latency, not bandwidth, bounds this kernel), the 32 warp
// sums are added in warp order.  W += sum (single GPU) or dW_out = sum (exchange).
__global__ void __launch_bounds__(1024) f4tc_reduce_kernel(const float* __restrict__ partials, int n_partials, int fa, float* __restrict__ W,
                                                           float* __restrict__ dW_out) {
    __shared__ float part[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, j = blockIdx.x * 32 + lane;
    float acc = 0.0f;
    if (j < fa)
        for (int p = w; p < n_partials; p += 32) acc += partials[(size_t)p * fa + j];
    part[w][lane] = acc;
    __syncthreads();
    if (w == 0 && j < fa) {
        float g = part[0][lane];
#pragma unroll
        for (int x = 1; x < 32; ++x) g += part[x][lane];
        if (dW_out) dW_out[j] = g;
        else W[j] += g;
    }
}

}  // namespace rsrl
