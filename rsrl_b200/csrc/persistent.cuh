// persistent.cuh — K batched steps in ONE launch (sm_100a): a persistent grid of thread-block clusters, one CTA per SM.
//
// Each CTA owns a contiguous slice of envs; with one env per thread the env state and the Fourier
// tables of the current state live in registers for the whole launch (HBM traffic = one read + one
// write of the state per launch).  SHARED weights need W_{t+1} = W_t + sum over ALL envs of the
// step-t updates before anybody can take step t+1, so every step ends in a grid-wide, fixed-order
// (bit-reproducible) reduction of NV = F*A values:
//
//   CTA     : env threads write phi(s_t) (feature-major rows, lane = env slot: conflict-free) while
//             they evaluate Q(s_t), and their scaled TD error per action column.  The CTA reduce is WARP-LOCAL: each warp
//             sums the 32 slots of its own envs right after stepping them (lane = row, FFMA2 over even / odd slots), the
//             warp partials are added in warp order by thread j after one block barrier (`part`, NV values).
//   hop A   : (cluster, DSMEM) every member CTA stores its partial into the cluster leader's shared memory with
//             st.async (16 bytes per thread, shared::cluster) that complete on the leader's mbarrier (complete_tx);
//             the leader sums the partials in rank order.  No polling: the waiters sleep on the mbarrier.
//   hop N   : (multi-GPU, NVLink) leader c of every GPU stores its cluster partial as 8-byte LL words
//             {payload, epoch} into slot (rank, c) of EVERY GPU's inbox through cudaIpc-mapped peer pointers and
//             sums the world's slot-c partials from its own inbox — all leaders of all GPUs in parallel.
//   hop B   : (L2) the leaders exchange their (world-)cluster partials as 16-byte LL lines {3 payload words,
//             epoch}; lane l of a row polls clusters l, l + lpr, ...; a butterfly gives the total.
//   hop C   : (cluster, DSMEM) the leader stores the total into every member's shared memory (st.async + the
//             member's mbarrier); every CTA applies W += total to its own copy of W.
//
// fp32 engines (the bench dtype) use a shorter exchange instead of hops A-C — COUNTING ACCUMULATORS in L2:
//   every CTA converts its partial to 2^-40 fixed point and adds (value << 8) + 1 to NV 64-bit words with relaxed
//   reductions (red.add.u64): the low byte counts arrivals, the upper 56 bits are a running (never reset, wrapping) sum.
//   Every CTA polls the NV words until the count says that all G partials of this step are in, and takes the difference
//   to the running sum it saw at the previous completion.  Data and flag share one naturally atomic 64-bit word (no
//   fence, no 16-byte assumption), integer addition is order independent (any arrival order gives the same bits), two
//   tables alternate by step parity so that a fast CTA's next-step contribution cannot overtake a slow reader, and the
//   whole exchange is ONE L2 hop.  Multi-GPU: the CTAs form groups with their own tables; each group leader forwards its
//   group's total to every GPU's world table through NVLink peer pointers (red.add.u64 at system scope) and every CTA
//   polls the world table instead.  While a CTA waits, its envs evaluate the action-independent part of the NEXT transition.
//   Measured against the cluster exchange in profiles/r02_persistent.md.
//
// Every CTA of every GPU performs the same additions in the same order, so all W copies stay bit-identical;
// the order is restated on the host by oracle/oracle32.cpp (bit-exact parity of the fp32 path).
// One DSMEM hop costs ~250 cycles, one L2 LL hop ~700 (tools/microbench/pingpong.cu); round 1's version
// (two L2 hops polled by all 148 CTAs, NVLink hop serialised behind CTA 0) is in profiles/r01_final.md.
#pragma once
#include <type_traits>
#include "kernels.cuh"

namespace rsrl {

constexpr int kModeSharedTrace = 2;  // internal MODE: SHARED weights + per-env traces kept in shared memory
constexpr int kMaxRanks = 8;
constexpr int kMaxClusters = 64;     // leaders that exchange through L2
constexpr int kMaxClusterSize = 16;
constexpr int kMaxGroups = 8;        // group tables of the multi-GPU counting exchange
constexpr int kAccStride = 16;       // 64-bit words between two accumulators: one 128-byte line each (atomics on one line serialise)
constexpr double kFxScale = 1099511627776.0;   // 2^40: fixed-point unit of the counting exchange
constexpr float kFxLimit = 16384.0f;           // |CTA partial| must stay below 2^14 (2^54 in fixed point; the field has 56 bits)

// two fused multiply-adds per issue slot (FFMA2, sm_100): each half is an ordinary IEEE fma, so the host replay stays scalar
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

struct SyncArgs {
    uint4* stage;   // [2][n_clusters][ROWS * LPW]   (world-)cluster partial rows (parity double-buffered)
    int cluster_size;
    int n_clusters;
    int lpr;        // LL lanes per row (power of two <= 8, ROWS * lpr <= blockDim)
    int cap;        // slot stride of the reduce rows when it is not a compile-time constant (persist_cap_static == 0): persist_cap(blockDim)
    int pe_smem;    // PER_ENV: keep every env's own W (F*A values, column `tid`) in shared memory for the whole launch
    int debug_skip; // development timing aid (RSRL_B200_DEBUG_SKIP): bit 0 skips the grid exchange, bit 1 the CTA reduce (wrong results)
    uint32_t epoch_base;  // exchange epoch before the launch's first step (monotonic over the engine's life, never reset)
    // fp32 exchange through counting accumulators in L2 (see the header comment); fx == 0: the cluster + LL-line exchange
    int fx;
    int ngroups;              // several GPUs: CTA groups per GPU (<= kMaxGroups), each with its own table and a leader that forwards its total
    int poll_delay_ns;        // pause between the reductions and the first poll / between two polls: polls that come before the last
    int poll_backoff_ns;      // partial has landed only queue in front of the reductions in L2
    int world_backoff_ns;     // several GPUs: pause between two polls of the world table
    unsigned long long* acc;  // [2][kMaxGroups][NV][kAccStride] running fixed-point sums of the CTA partials + arrival count in the low byte (parity double-buffered)
    long long* prev;          // [2 + 2 * kMaxGroups][NV] running sums at the previous completion: world table parity 0, 1, then every group table
};

// Cross-GPU exchange (one process per GPU): every rank owns an inbox of 8-byte LL words
// {payload, epoch} that its peers write through NVLink (cudaIpc-mapped pointers).
struct PeerArgs {
    uint2* inbox[kMaxRanks];  // inbox[r]: rank r's mailbox as seen from this GPU.  LL exchange: [2][world][n_clusters][NV * WPV] 8-byte LL words;
                              // counting exchange (fp32): the rank's world table [2][NV][kAccStride] of 64-bit accumulators
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
// 16-byte LL line: data and flag travel in one aligned 16-byte access.  The PTX memory model only promises
// single-copy atomicity per scalar element of a vector access; a B200 L2 serves an aligned 16-byte access as one
// 32-byte-sector transaction (the property NCCL's LL128 protocol is built on).  tests/test_gpu_parity.py checks it:
// every step of a full-size run against oracle32 bit for bit, and a 10^6-step run repeated with identical results.
__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t epoch) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(epoch) : "memory");
}

// ---- counting accumulators: value and arrival count in ONE 64-bit word, added with relaxed reductions ----
__device__ __forceinline__ void red_add_u64_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_u64_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_u64_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// fp32 <-> the 2^-40 fixed-point grid (one rounding each way; identical on the host: oracle/oracle32.cpp)
__host__ __device__ __forceinline__ long long fx_from_float(float x) {
#ifdef __CUDA_ARCH__
    return __double2ll_rn((double)x * kFxScale);
#else
    return (long long)llrint((double)x * kFxScale);
#endif
}
__host__ __device__ __forceinline__ float fx_to_float(long long q) { return (float)((double)q * (1.0 / kFxScale)); }
// the word holds (sum << 8) + count (mod 2^64): take `count` contributions off and read the 56-bit running sum
__host__ __device__ __forceinline__ long long fx_running_sum(unsigned long long word, unsigned long long count) {
    return (long long)(word - count) >> 8;
}
__host__ __device__ __forceinline__ long long fx_sext56(long long v) { return (long long)((unsigned long long)v << 8) >> 8; }

// ---- thread-block cluster / DSMEM primitives ----
__device__ __forceinline__ uint32_t cl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cl_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_mapa(uint32_t saddr, uint32_t rank) {  // my shared::cta address -> the same offset in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cl_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_expect_tx(uint32_t bar, uint32_t bytes) {  // one arrival + `bytes` pending transaction bytes
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cl_mbar_wait(uint32_t bar, uint32_t parity) {  // try_wait suspends the thread in hardware: not a busy poll
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
// bytes (multiple of 16) from my shared memory to the shared memory of another CTA of the cluster; completes on that CTA's mbarrier
__device__ __forceinline__ void cl_bulk_copy(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t mbar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void cl_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16 bytes from registers to the shared memory of another CTA of the cluster; counts 16 bytes on that CTA's mbarrier.
// Latency of a DSMEM store (~250 cycles); a bulk copy of the same data goes through the copy engine and costs several times that.
__device__ __forceinline__ void cl_st_async16(uint32_t dst_cluster_addr, uint4 v, uint32_t mbar_cluster_addr) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(dst_cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar_cluster_addr) : "memory");
}

// A "row" carries NV <= 3 values.  fp32: one 16-byte LL line {v0, v1, v2, epoch}; fp64: one line per value.
template <typename R, int NV> struct LLRow;
template <int NV> struct LLRow<float, NV> {
    static constexpr int LPW = 1;
    __device__ __forceinline__ static void publish(uint4* line, const float* v, uint32_t epoch) {
        st_ll(line, __float_as_uint(v[0]), NV > 1 ? __float_as_uint(v[1]) : 0u, NV > 2 ? __float_as_uint(v[2]) : 0u, epoch);
    }
    __device__ __forceinline__ static void load(const uint4* line, uint4* w) { w[0] = ld_ll(line); }
    __device__ __forceinline__ static bool ready(const uint4* w, uint32_t epoch) { return w[0].w == epoch; }
    __device__ __forceinline__ static void add(const uint4* w, float* acc) {
        acc[0] += __uint_as_float(w[0].x);
        if (NV > 1) acc[1] += __uint_as_float(w[0].y);
        if (NV > 2) acc[2] += __uint_as_float(w[0].z);
    }
};
template <int NV> struct LLRow<double, NV> {
    static constexpr int LPW = NV;
    __device__ __forceinline__ static void publish(uint4* line, const double* v, uint32_t epoch) {
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            const unsigned long long u = (unsigned long long)__double_as_longlong(v[c]);
            st_ll(line + c, (uint32_t)u, (uint32_t)(u >> 32), 0u, epoch);
        }
    }
    __device__ __forceinline__ static void load(const uint4* line, uint4* w) {
#pragma unroll
        for (int c = 0; c < NV; ++c) w[c] = ld_ll(line + c);
    }
    __device__ __forceinline__ static bool ready(const uint4* w, uint32_t epoch) {
        bool ok = true;
#pragma unroll
        for (int c = 0; c < NV; ++c) ok &= w[c].w == epoch;
        return ok;
    }
    __device__ __forceinline__ static void add(const uint4* w, double* acc) {
#pragma unroll
        for (int c = 0; c < NV; ++c) acc[c] += __longlong_as_double((long long)(((unsigned long long)w[c].y << 32) | w[c].x));
    }
};

// acc = sum, in the order m = first, first + step, ... (< cnt), of the rows published at base + m * stride.
// Up to MAXO lines are requested before the first one is checked, so that their L2 round trips overlap.
template <typename R, int NV>
__device__ __forceinline__ void ll_gather(const uint4* base, size_t stride, int first, int step, int cnt, uint32_t epoch, bool valid, R* acc) {
    using LR = LLRow<R, NV>;
    constexpr int MAXO = LR::LPW == 1 ? 5 : 2;
#pragma unroll
    for (int c = 0; c < NV; ++c) acc[c] = (R)0;
    if (!valid) return;
    for (int m0 = first; m0 < cnt; m0 += MAXO * step) {
        uint4 w[MAXO][LR::LPW];
#pragma unroll
        for (int o = 0; o < MAXO; ++o)
            if (m0 + o * step < cnt) LR::load(base + (size_t)(m0 + o * step) * stride, w[o]);
#pragma unroll
        for (int o = 0; o < MAXO; ++o) {
            if (m0 + o * step < cnt) {
                while (!LR::ready(w[o], epoch)) LR::load(base + (size_t)(m0 + o * step) * stride, w[o]);
                LR::add(w[o], acc);
            }
        }
    }
}

// fixed-order butterfly over the lpr adjacent lanes that own one row (executed by whole warps)
template <typename R, int NV>
__device__ __forceinline__ void row_butterfly(R* v, int lpr) {
    for (int off = 1; off < lpr; off <<= 1) {
#pragma unroll
        for (int c = 0; c < NV; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], off);
    }
}

template <typename R> struct Vec16;  // 16-byte shared-memory vector of R
template <> struct Vec16<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct Vec16<double> { typedef double2 type; static constexpr int N = 2; };
// acc[0] += p.even * d.even, acc[1] += p.odd * d.odd over the slots of one 16-byte group, in slot order
__device__ __forceinline__ void pair_fma(const float4& p, const float4& d, float* acc) {
    float2 a = make_float2(acc[0], acc[1]);
    a = ffma2(make_float2(p.x, p.y), make_float2(d.x, d.y), a);
    a = ffma2(make_float2(p.z, p.w), make_float2(d.z, d.w), a);
    acc[0] = a.x; acc[1] = a.y;
}
__device__ __forceinline__ void pair_fma(const double2& p, const double2& d, double* acc) {
    acc[0] = dfma(p.x, d.x, acc[0]);
    acc[1] = dfma(p.y, d.y, acc[1]);
}
__device__ __forceinline__ float vget(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ double vget(const double2& v, int i) { return i == 0 ? v.x : v.y; }

// CTA reduce: WARP-LOCAL.  Each warp reduces the 32 slots of its own envs right after it has stepped them — no CTA barrier
// between the env phase and the reduce, no idle warps, and a slow warp's env phase overlaps the other warps' reduces.
// Pass p: lane l owns row 32 p + l and sums the warp's 32 slots (16-byte loads; even / odd slots in the two halves of an
// FFMA2).  The row stride `cap` is 4 (mod 32) words, so the 32 rows of a pass fall into distinct bank groups.  A last pass
// with n < 32 rows gives every row S = persist_tail_split(n) adjacent lanes (32 / S slots each, butterfly over the S lanes).
// The warp partials go to wpart[warp][NV]; after one CTA barrier thread j adds wpart[0..NW)[j] in warp order.
constexpr int kPersistMaxBlock = 512;  // (the register file is per scheduler: 13-16 warps all mean 4 warps on one scheduler = 128 registers)
constexpr int kPersistMaxWarps = kPersistMaxBlock / 32;
__host__ __device__ constexpr int persist_tail_split(int nrow, int vn) {  // lanes per row in a pass of nrow <= 32 rows; vn = slots per 16 bytes
    int s = 1;
    while (2 * s * nrow <= 32 && 32 / (2 * s) >= vn) s *= 2;
    return s;
}
// Row stride of the reduce buffer in slots: >= block, a multiple of the 16-byte vector, 4 (mod 32) words.
__host__ __device__ constexpr int persist_cap(int block, int rsz) { return block + 16 / rsz; }
// A compile-time stride whenever the rows fit (every STS / LDS of the env phase gets an immediate offset instead of an address
// computation: -70 instructions per env-step on Fourier(5)); otherwise persist_cap(blockDim) at run time.
__host__ __device__ constexpr int persist_cap_static(int rows, int rsz, bool trace) {
    return (!trace && (long long)rows * persist_cap(kPersistMaxBlock, rsz) * rsz <= 96 * 1024) ? persist_cap(kPersistMaxBlock, rsz) : 0;
}

// NV values of R padded to a multiple of 16 bytes (bulk copies); host and device lay the shared memory out from this
__host__ __device__ constexpr int persist_nvp(int nv, int rsz) { return (nv * rsz + 15) / 16 * 16 / rsz; }

template <typename R, int DOM, int BASIS, int P, int AW, int MODE>
__global__ void __launch_bounds__(kPersistMaxBlock, 1) persistent_kernel(const StepArgs a, const int k_steps, const SyncArgs sy, const PeerArgs pe) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    using V = Vec16<R>;
    typedef typename V::type vec_t;
    constexpr int D = Dom::D, F = GB::F, FA = F * AW;
    constexpr bool TDPRED = AW == 1;
    constexpr bool SHAREDW = MODE != RSRL_PER_ENV;          // one replicated W, dW reduced over the grid
    constexpr bool TRACE = MODE == kModeSharedTrace;        // eligibility traces resident in shared memory
    constexpr int ROWS = TRACE ? FA : F;                    // rows of the reduce buffer: z (F*A) or phi(s_t) (F)
    constexpr int NDC = TRACE ? 1 : AW;                     // values per row = rows of scaled TD errors
    constexpr int NV = ROWS * NDC;                          // values of one dW partial (= F*A), index row * NDC + c == the flat F x A index
    constexpr int NVP = persist_nvp(NV, (int)sizeof(R));
    constexpr uint32_t NVB = NVP * sizeof(R);               // bytes of one partial (multiple of 16)
    constexpr int NCH = (int)(NVB / 16);                    // 16-byte chunks of one partial (one st.async each)
    constexpr bool FX = sizeof(R) == 4;                     // fp32: counting exchange; f64: cluster + LL lines (each instantiation carries one)
    constexpr int NVP8 = (NV + 1) / 2 * 2;                  // prevs rows (long long), padded to 16 bytes
    constexpr int WS = 4;            // padded row stride of the shared W copy: one LDS.128 per feature row
    constexpr int FApad = F * WS;
    using LR = LLRow<R, NDC>;
    constexpr int LPW = LR::LPW;     // LL lines per row
    constexpr int WPV = sizeof(R) / 4;

    const int tid = threadIdx.x, BLOCK = blockDim.x, G = gridDim.x, b = blockIdx.x;
    const int64_t N = a.n;
    const int64_t per_cta = (N + G - 1) / G;
    const int64_t base = (int64_t)b * per_cta < N ? (int64_t)b * per_cta : N;
    const int64_t end = base + per_cta < N ? base + per_cta : N;
    const int n_chunks = (int)((per_cta + BLOCK - 1) / BLOCK);
    const bool resident = n_chunks == 1;  // one env per thread: state stays in registers across steps
    constexpr int CAPT = SHAREDW ? persist_cap_static(ROWS, (int)sizeof(R), TRACE) : 0;
    const int lpr = sy.lpr, cap = CAPT ? CAPT : sy.cap;
    constexpr bool ALIAS_TAB = SHAREDW && !TRACE;  // phi(s_t) is in shared memory after evalS: the tables of s' may overwrite those of s
    const int CS = SHAREDW ? sy.cluster_size : 1;
    const int crank = CS > 1 ? (int)cl_ctarank() : 0;   // rank in the cluster; rank 0 leads
    const int cid = b / CS;                             // cluster index
    const bool leader = crank == 0;

    // shared memory (SHARED modes).  Row stride cap > BLOCK slots (persist_cap); slots >= BLOCK are never read.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(smem_raw);  // [0] leader: members' partials, [1] member: the total
    R* part = reinterpret_cast<R*>(smem_raw + 16);  // [NVP]     this CTA's dW partial (bulk-copy source)
    R* totbuf = part + NVP;                         // [NVP]     grid total (leader: bulk-copy source, member: destination)
    R* inbuf = totbuf + NVP;                        // [CS][NVP] leader: the members' partials; slot 0: the grid total (bulk-copy source)
    R* Wsm = inbuf + (size_t)CS * NVP;        // [FApad]
    R* red = Wsm + FApad;                     // [ROWS][cap]  phi(s_t) feature-major, or the traces z[F*A][slot]
    R* dcs = red + (size_t)ROWS * cap;        // [NDC][cap]   scaled TD error in the action's row, 0 elsewhere
    R* wpart = dcs + (size_t)NDC * cap;       // [BLOCK / 32][NVP] warp partials of the CTA reduce
    R* gath = wpart + (size_t)(BLOCK >> 5) * NVP;  // [n_clusters][NVP] leader: the clusters' partials gathered from L2 (only when n_clusters > 1)
    // counting exchange: running sums seen at the previous completion [4][NVP8] (local parity 0, 1; world parity 0, 1); it takes gath's place
    long long* prevs = reinterpret_cast<long long*>(gath);
    const uint32_t mb_in = cl_smem_u32(&mbars[0]), mb_tot = cl_smem_u32(&mbars[1]);

    if (SHAREDW) {
        for (int j = tid; j < FA; j += BLOCK) Wsm[(j / AW) * WS + j % AW] = static_cast<const R*>(a.W)[j];
        for (int j = tid; j < (ROWS + NDC) * cap + (BLOCK >> 5) * NVP; j += BLOCK) red[j] = (R)0;
        for (int j = tid; j < NVP; j += BLOCK) { part[j] = (R)0; totbuf[j] = (R)0; }
        if (FX) {
            // running sums at the previous completion: rows 0, 1 = the table this CTA polls (one GPU: the only table; several GPUs:
            // the table of the group it leads), rows 2, 3 = the world table.  sy.prev: [2][NV] world, then [kMaxGroups][2][NV] groups.
            const int myg = (pe.world > 1 && b < kMaxGroups) ? b : 0;
            for (int j = tid; j < 2 * NV; j += BLOCK) {
                prevs[(j / NV) * NVP8 + j % NV] = sy.prev[(size_t)(2 + 2 * myg) * NV + j];
                prevs[(2 + j / NV) * NVP8 + j % NV] = sy.prev[j];
            }
        }
        if (CS > 1) {
            if (tid == 0) {
                cl_mbar_init(mb_in, 1);
                cl_mbar_init(mb_tot, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            cl_sync();  // every CTA's barriers are initialised before anybody sends
        } else {
            __syncthreads();
        }
    }
    const R* Wg = static_cast<const R*>(a.W);
    R* Wpe = reinterpret_cast<R*>(smem_raw);  // PER_ENV + pe_smem: [FA][BLOCK], column tid = this env's weights
    const bool pe_smem = !SHAREDW && sy.pe_smem != 0;

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
    if (pe_smem && active) {  // this env's weights: HBM -> shared memory once per launch (read again only at the end)
        for (int j = 0; j < FA; ++j) Wpe[(size_t)j * BLOCK + tid] = Wg[(int64_t)j * N + i];
    }
    if (TRACE && active) {  // this env's trace column: HBM -> shared memory once per launch
        const R* Z = static_cast<const R*>(a.z);
        for (int j = 0; j < FA; ++j) red[(size_t)j * cap + tid] = Z[(int64_t)j * N + i];
    }
    typename GB::Tab tab_s, tab_n;
    bool have_tab = false;  // tab_s holds the tables of s (carried from the previous step's s')
    // The action-independent part of the next transition (MountainCar: the f64 cosine) is evaluated for s_{t+1} while the CTA waits
    // for the grid exchange of step t: work-neutral (it is needed at step t + 1 whatever the action), two registers, no shared memory.
    // (fp32 engines only: in the f64 cluster exchange the leaders have no idle window before their sends — measured 13.5 -> 14.7 us per step)
    constexpr bool PRE = Dom::kHasPre && SHAREDW && sizeof(R) == 4;
    double pre = 0.0;
    bool have_pre = false;
    auto precompute = [&]() {
        if constexpr (PRE) {
            if (resident && active) { pre = Dom::step_pre(s); have_pre = true; }
        }
    };
    precompute();

    // LL row ownership: lanes [row * lpr, (row + 1) * lpr) own row `row` in the exchanges between leaders
    const int row = tid / lpr, rl = tid % lpr;
    const bool row_valid = SHAREDW && row < ROWS;
    const bool warp_rows = SHAREDW && (tid & ~31) < ROWS * lpr;  // warp-uniform: this warp owns at least one row

    // phase profile (RSRL_B200_PHASE_PROFILE=1): thread 0 adds its cycle counts straight to global memory — no counters in registers
    const bool prof = a.phase_prof != nullptr && tid == 0;
    long long c0 = 0;
    auto tick = [&](int q) {
        const long long c1 = clock64();
        a.phase_prof[b * 8 + q] += c1 - c0;
        c0 = c1;
    };
    for (int step = 0; step < k_steps; ++step) {
        const uint64_t t = a.t + (uint64_t)step;
        if (prof) c0 = clock64();
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
                    } else if (pe_smem) {
#pragma unroll
                        for (int c = 0; c < AW; ++c) w[c] = Wpe[(size_t)(k * AW + c) * BLOCK + tid];
                    } else {
#pragma unroll
                        for (int c = 0; c < AW; ++c) w[c] = Wg[(int64_t)(k * AW + c) * N + i];
                    }
                };
                // fp32, SHARED: q[0..1] += phi * W[k][0..1] is ONE FFMA2 (phi broadcast from a scalar register, the W pair straight
                // from the LDS.128) — per component the same fma as the scalar form, so oracle32 replays it unchanged
                constexpr bool PACKQ = sizeof(R) == 4 && SHAREDW && AW >= 2 && AW <= 4;
                auto eval = [&](const typename GB::Tab& tab, R* q, auto rec) {
                    if constexpr (PACKQ) {
                        // (without this the compiler keeps the 108 W values of the first evaluation alive for the second one: 600 bytes
                        // of local-memory spills per thread instead of 36 LDS.128)
                        asm volatile("" ::: "memory");
                        float2 q01 = make_float2(0.0f, 0.0f), q23 = make_float2(0.0f, 0.0f);
                        float q2 = 0.0f;
                        GB::template for_each<true>(tab, [&](int k, R phi) {
                            if (decltype(rec)::value) red[k * cap + tid] = phi;
                            const float4 w = *reinterpret_cast<const float4*>(Wsm + k * WS);
                            q01 = ffma2(make_float2(phi, phi), make_float2(w.x, w.y), q01);
                            if (AW == 3) q2 = ffma(phi, w.z, q2);
                            if (AW == 4) q23 = ffma2(make_float2(phi, phi), make_float2(w.z, w.w), q23);
                        });
                        q[0] = q01.x; q[1] = q01.y;
                        if (AW == 3) q[2] = q2;
                        if (AW == 4) { q[2] = q23.x; q[3] = q23.y; }
                    } else {
#pragma unroll
                        for (int c = 0; c < AW; ++c) q[c] = (R)0;
                        GB::for_each(tab, [&](int k, R phi) {
                            if (decltype(rec)::value) red[k * cap + tid] = phi;
                            R w[AW];
                            wrow(k, w);
#pragma unroll
                            for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, w[c], q[c]);
                        });
                    }
                };
                // Q(s_t) and, SHARED without traces, the phi(s_t) row; Q(s')
                auto evalS = [&](const typename GB::Tab& tab, R* q) { eval(tab, q, std::integral_constant<bool, SHAREDW && !TRACE>()); };
                auto evalN = [&](const typename GB::Tab& tab, R* q) { eval(tab, q, std::integral_constant<bool, false>()); };
                auto prep = [](const double* st, typename GB::Tab& tb) { grid_prepare<R, Dom, P, BASIS>(st, tb); };
                env_core<R, DOM, AW, false>(a, t, g, s, prep, evalS, evalN, tab_s, ALIAS_TAB ? tab_s : tab_n, have_tab, o, 0, 0.0, false, nullptr,
                                            (PRE && have_pre) ? &pre : nullptr);
                have_pre = false;  // (consumed: a step without a fresh precompute() evaluates it inline)
                if (a.td) static_cast<R*>(a.td)[i] = o.residual;
                if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
                if (MODE == RSRL_PER_ENV) {
                    R* Wm = static_cast<R*>(a.W);
                    GB::for_each(tab_s, [&](int k, R phi) {
                        const int col = k * AW + (TDPRED ? 0 : o.act);
                        if (pe_smem) {
                            R* wp = Wpe + (size_t)col * BLOCK + tid;
                            *wp = O::mul_add_unfused(o.coef, phi, *wp);
                        } else {
                            const int64_t idx = (int64_t)col * N + i;
                            Wm[idx] = O::mul_add_unfused(o.coef, phi, Wm[idx]);
                        }
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
                if (have_tab && !ALIAS_TAB) tab_s = tab_n;
                if (!resident) {
                    a.ep_steps[i] = ep;
                    a.actions[i] = act;
#pragma unroll
                    for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
                }
            }
            if (prof) tick(0);
            if (SHAREDW) {
                if (TRACE) {
                    dcs[tid] = active ? o.coef : (R)0;
                } else {
#pragma unroll
                    for (int c = 0; c < AW; ++c) dcs[c * cap + tid] = (active && (TDPRED || c == o.act)) ? o.coef : (R)0;
                }
                // (a slot idle in this chunk keeps a stale but finite row; its dcs entries are 0)
                __syncwarp();
                // ---- warp-local reduce of this warp's 32 slots (see the header) ----
                if (!(sy.debug_skip & 2)) {
                    const int lane = tid & 31, slot0 = tid & ~31;
                    R* wp = wpart + (size_t)(tid >> 5) * NVP;
#pragma unroll
                    for (int r0 = 0; r0 < ROWS; r0 += 32) {
                        constexpr int VN = V::N;
                        const int nrow = ROWS - r0 < 32 ? ROWS - r0 : 32;          // compile-time after unrolling
                        const int S = persist_tail_split(nrow, VN), NSL = 32 / S;  // lanes per row, slots per lane
                        const int rloc = lane / S, sub = lane % S;
                        R acc[NDC][2];
#pragma unroll
                        for (int c = 0; c < NDC; ++c) acc[c][0] = acc[c][1] = (R)0;
                        if (rloc < nrow) {
                            const R* prow = red + (size_t)(r0 + rloc) * cap + slot0 + sub * NSL;
                            const R* drow = dcs + slot0 + sub * NSL;
#pragma unroll
                            for (int gq = 0; gq < NSL / VN; ++gq) {
                                const vec_t pv = *reinterpret_cast<const vec_t*>(prow + gq * VN);
#pragma unroll
                                for (int c = 0; c < NDC; ++c)
                                    pair_fma(pv, *reinterpret_cast<const vec_t*>(drow + (size_t)c * cap + gq * VN), acc[c]);
                            }
                        }
                        R v[NDC];
#pragma unroll
                        for (int c = 0; c < NDC; ++c) v[c] = acc[c][0] + acc[c][1];
#pragma unroll
                        for (int off = 1; off < S; off <<= 1) {
#pragma unroll
                            for (int c = 0; c < NDC; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], off);
                        }
                        if (rloc < nrow && sub == 0) {
#pragma unroll
                            for (int c = 0; c < NDC; ++c) {
                                R* pp = wp + (r0 + rloc) * NDC + c;
                                *pp = chunk == 0 ? v[c] : *pp + v[c];  // chunks in order
                            }
                        }
                    }
                }
                if (n_chunks > 1 || TRACE) __syncwarp();  // the warp's rows are rewritten by the next chunk / cleared below
                if (TRACE && active && o.terminated) {  // trace.reset() (sarsa_lambda.rs:78, q_lambda.rs:81, td_lambda.rs:61)
                    for (int j = 0; j < FA; ++j) red[(size_t)j * cap + tid] = (R)0;
                }
            }
        }

        if (SHAREDW) {
            __syncthreads();
            if (prof) tick(1);
            // CTA partial: the warp partials added in warp order (thread j owns value j)
            const int NW = BLOCK >> 5;
            for (int j = tid; j < NV; j += BLOCK) {
                R v[kPersistMaxWarps];  // all loads first: one shared-memory latency, not NW
#pragma unroll
                for (int w = 0; w < kPersistMaxWarps; ++w) v[w] = w < NW ? wpart[(size_t)w * NVP + j] : (R)0;
                R acc = v[0];
#pragma unroll
                for (int w = 1; w < kPersistMaxWarps; ++w) if (w < NW) acc += v[w];
                part[j] = acc;
            }
            if (!FX) { __syncthreads(); precompute(); }  // fp32: thread j alone reads part[j] below (reduction, poll, W update): no barrier
            if (prof) tick(2);
            const uint32_t epoch = sy.epoch_base + (uint32_t)step + 1u;
            const int par = (int)(epoch & 1u);
            const uint32_t ph = (uint32_t)(step & 1);  // one mbarrier completion per step
            const R* total = part;                     // one CTA (or exchange skipped): its partial is the total
            if constexpr (FX) {
              if ((G > 1 || pe.world > 1) && !(sy.debug_skip & 1)) {
                // ---- counting accumulators (fp32) ----
                // One GPU: every CTA adds to ONE table and every CTA polls it (one L2 hop).
                // Several GPUs: the CTAs form NG groups (CTA b belongs to group b % NG, CTA q < NG leads group q); a group's table
                // fills quickly (G / NG reductions per word), its leader forwards the group total to the world table of EVERY GPU
                // through the NVLink peer pointers, and every CTA polls its own GPU's world table (world * NG arrivals per step).
                // The NVLink flight of the early groups overlaps the local reductions of the late ones; nothing waits for a
                // whole-GPU total.  Integer sums: the result does not depend on the grouping.
                const int p = par;
                const unsigned long long k = ((unsigned long long)epoch + (unsigned long long)p) >> 1;  // steps of this parity so far, this one included
                const bool multi = pe.world > 1;
                constexpr int AST = kAccStride;  // one 128-byte line per accumulator: packed words or one bulk reduction
                                                 // (cp.reduce.async.bulk.add.u64) per CTA are slower (profiles/r02_persistent.md)
                const int NG = multi ? (sy.ngroups < G ? sy.ngroups : G) : 1;
                const int grp = b % NG;
                unsigned long long* gtab = sy.acc + ((size_t)p * kMaxGroups + grp) * NV * AST;  // my group's table
                for (int j = tid; j < NV; j += BLOCK) {
                    float x = (float)part[j];
                    if (!(fabsf(x) < kFxLimit)) { atomicExch(&a.counters->nonfinite, 1); x = 0.0f; }  // NaN / Inf / beyond the fixed-point range
                    red_add_u64_gpu(gtab + (size_t)j * AST, ((unsigned long long)fx_from_float(x) << 8) + 1ull);
                }
                if (!multi || b >= NG) precompute();  // (group leaders first forward their group's total)
                if (sy.poll_delay_ns > 0 && !multi) __nanosleep((unsigned)sy.poll_delay_ns);
                if (!multi || b < NG) {
                    const unsigned long long cnt = (unsigned long long)((G - grp + NG - 1) / NG) * k;  // CTAs grp, grp + NG, ... over k steps
                    for (int j = tid; j < NV; j += BLOCK) {
                        const unsigned long long* word = gtab + (size_t)j * AST;
                        unsigned long long w = ld_u64_gpu(word);
                        while ((w & 0xFFull) != (cnt & 0xFFull)) {
                            if (sy.poll_backoff_ns > 0) __nanosleep((unsigned)sy.poll_backoff_ns);
                            w = ld_u64_gpu(word);
                        }
                        const long long sum = fx_running_sum(w, cnt);
                        const long long dq = fx_sext56(sum - prevs[p * NVP8 + j]);
                        prevs[p * NVP8 + j] = sum;
                        if (!multi) {
                            Wsm[(j / AW) * WS + j % AW] += (R)fx_to_float(dq);
                        } else {  // group leader: the group total goes to every GPU's world table
                            for (int r = 0; r < pe.world; ++r) {
                                const int rr = (pe.rank + r) % pe.world;  // own table first, then the peers in a rank-rotated order
                                red_add_u64_sys(reinterpret_cast<unsigned long long*>(pe.inbox[rr]) + ((size_t)p * NV + j) * AST,
                                                ((unsigned long long)dq << 8) + 1ull);
                            }
                        }
                    }
                }
                if (multi) {
                    if (b < NG) precompute();
                    if (sy.poll_delay_ns > 0) __nanosleep((unsigned)sy.poll_delay_ns);
                    const unsigned long long cnt = (unsigned long long)pe.world * (unsigned long long)NG * k;
                    const unsigned long long* wtab = reinterpret_cast<const unsigned long long*>(pe.inbox[pe.rank]) + (size_t)p * NV * AST;
                    for (int j = tid; j < NV; j += BLOCK) {
                        const unsigned long long* word = wtab + (size_t)j * AST;
                        unsigned long long w = ld_u64_sys(word);
                        while ((w & 0xFFull) != (cnt & 0xFFull)) {
                            if (sy.world_backoff_ns > 0) __nanosleep((unsigned)sy.world_backoff_ns);  // 148 CTAs spinning on the lines the NVLink reductions land on
                            w = ld_u64_sys(word);
                        }
                        const long long sum = fx_running_sum(w, cnt);
                        const long long dq = fx_sext56(sum - prevs[(2 + p) * NVP8 + j]);
                        prevs[(2 + p) * NVP8 + j] = sum;
                        Wsm[(j / AW) * WS + j % AW] += (R)fx_to_float(dq);
                    }
                }
              } else {
                for (int j = tid; j < NV; j += BLOCK) Wsm[(j / AW) * WS + j % AW] += part[j];
                precompute();
              }
            } else {
            if ((G > 1 || pe.world > 1) && !(sy.debug_skip & 1)) {
                if (!leader) {
                    // hop A (send): my partial -> the leader's inbuf[crank]; then sleep until the total arrives (hop C)
                    if (tid == 0) cl_mbar_expect_tx(mb_tot, NVB);
                    for (int j = tid; j < NCH; j += BLOCK) {  // 16 bytes per thread, straight from `part` (complete after the barrier above)
                        const uint4 v = reinterpret_cast<const uint4*>(part)[j];
                        cl_st_async16(cl_mapa(cl_smem_u32(inbuf + (size_t)crank * NVP) + 16u * (uint32_t)j, 0), v, cl_mapa(mb_in, 0));
                    }
                    cl_mbar_wait(mb_tot, ph);
                } else {
                    // hop A (receive): cluster partial = the members' partials summed in rank order
                    if (CS > 1) {
                        if (tid == 0) cl_mbar_expect_tx(mb_in, (uint32_t)(CS - 1) * NVB);
                        cl_mbar_wait(mb_in, ph);
                    }
                    if (prof) tick(5);  // leader: members' partials arrived (hop A + skew between the cluster's CTAs)
                    for (int j = tid; j < NV; j += BLOCK) {
                        R acc = part[j];
                        for (int r = 1; r < CS; ++r) acc += inbuf[(size_t)r * NVP + j];
                        totbuf[j] = acc;
                    }
                    const int NCL = sy.n_clusters;
                    if (NCL > 1 || pe.world > 1) {
                        __syncthreads();  // (leader CTAs only; the branch is CTA-uniform) cluster partial complete in totbuf
                        R dW[NDC];
                        if (warp_rows) {
#pragma unroll
                            for (int c = 0; c < NDC; ++c) dW[c] = row_valid ? totbuf[row * NDC + c] : (R)0;
                            if (pe.world > 1) {
                                // hop N (NVLink): slot (my rank, cid) of every GPU's inbox <- my cluster partial; then sum the world's
                                // slot-cid partials: lane l <-> ranks l, l + lpr, ...; butterfly => the same order on every GPU
                                constexpr int RW = NDC * WPV;  // 32-bit payload words per row
                                if (row_valid && rl == 0) {
                                    uint32_t w[RW];
#pragma unroll
                                    for (int c = 0; c < NDC; ++c) {
                                        if (WPV == 1) w[c] = __float_as_uint((float)dW[c]);
                                        else { const unsigned long long u = (unsigned long long)__double_as_longlong((double)dW[c]); w[c * WPV] = (uint32_t)u; w[c * WPV + WPV - 1] = (uint32_t)(u >> 32); }
                                    }
                                    for (int r = 0; r < pe.world; ++r) {
                                        uint2* dst = pe.inbox[r] + ((((size_t)par * pe.world + pe.rank) * NCL + cid) * ROWS + row) * RW;
#pragma unroll
                                        for (int q = 0; q < RW; ++q) st_ll8_sys(dst + q, w[q], epoch);
                                    }
                                }
                                R tot[NDC];
#pragma unroll
                                for (int c = 0; c < NDC; ++c) tot[c] = (R)0;
                                if (row_valid) {
                                    for (int r = rl; r < pe.world; r += lpr) {
                                        const uint2* src = pe.inbox[pe.rank] + ((((size_t)par * pe.world + r) * NCL + cid) * ROWS + row) * RW;
                                        uint32_t w[RW];
#pragma unroll
                                        for (int q = 0; q < RW; ++q) {
                                            uint2 v = ld_ll8_sys(src + q);
                                            while (v.y != epoch) v = ld_ll8_sys(src + q);
                                            w[q] = v.x;
                                        }
#pragma unroll
                                        for (int c = 0; c < NDC; ++c) {
                                            if (WPV == 1) tot[c] += (R)__uint_as_float(w[c]);
                                            else tot[c] += (R)__longlong_as_double((long long)(((unsigned long long)w[c * WPV + WPV - 1] << 32) | w[c * WPV]));
                                        }
                                    }
                                }
                                __syncwarp();
                                row_butterfly<R, NDC>(tot, lpr);
#pragma unroll
                                for (int c = 0; c < NDC; ++c) dW[c] = tot[c];
                            }
                            if (NCL > 1 && row_valid && rl == 0) {
                                // hop B (L2), publish: my (world-)cluster partial as LL lines {payload, epoch}
                                LR::publish(sy.stage + ((size_t)par * NCL + cid) * ((size_t)ROWS * LPW) + (size_t)row * LPW, dW, epoch);
                            }
                        }
                        if (NCL > 1) {
                            // hop B, gather: ALL threads of the leader poll (<= 3 lines each, all requested before the first is checked: one
                            // L2 round trip) and drop the payloads into shared memory ...
                            const size_t rstride = (size_t)ROWS * LPW;  // lines between two producers
                            const uint4* st = sy.stage + (size_t)par * NCL * rstride;
                            constexpr int MAXL = 3;
                            for (int q0 = tid; q0 < ROWS * NCL; q0 += MAXL * BLOCK) {
                                uint4 w[MAXL][LPW];
#pragma unroll
                                for (int o = 0; o < MAXL; ++o) {
                                    const int q = q0 + o * BLOCK;
                                    if (q < ROWS * NCL) LR::load(st + (size_t)(q / ROWS) * rstride + (size_t)(q % ROWS) * LPW, w[o]);
                                }
#pragma unroll
                                for (int o = 0; o < MAXL; ++o) {
                                    const int q = q0 + o * BLOCK;
                                    if (q < ROWS * NCL) {
                                        const uint4* line = st + (size_t)(q / ROWS) * rstride + (size_t)(q % ROWS) * LPW;
                                        while (!LR::ready(w[o], epoch)) LR::load(line, w[o]);
                                        R v[NDC];
#pragma unroll
                                        for (int c = 0; c < NDC; ++c) v[c] = (R)0;
                                        LR::add(w[o], v);
#pragma unroll
                                        for (int c = 0; c < NDC; ++c) gath[(size_t)(q / ROWS) * NVP + (q % ROWS) * NDC + c] = v[c];
                                    }
                                }
                            }
                            __syncthreads();
                        }
                        if (warp_rows) {
                            if (NCL > 1) {
                                // ... where lane l of a row sums clusters l, l + lpr, ... in order; a butterfly over the lanes gives the total
#pragma unroll
                                for (int c = 0; c < NDC; ++c) dW[c] = (R)0;
                                if (row_valid) {
                                    for (int m = rl; m < NCL; m += lpr) {
#pragma unroll
                                        for (int c = 0; c < NDC; ++c) dW[c] += gath[(size_t)m * NVP + row * NDC + c];
                                    }
                                }
                                __syncwarp();
                                row_butterfly<R, NDC>(dW, lpr);
                            }
                            if (row_valid && rl == 0) {  // slot 0 of inbuf (no member writes it) takes the total
#pragma unroll
                                for (int c = 0; c < NDC; ++c) inbuf[row * NDC + c] = dW[c];
                            }
                        }
                        total = inbuf;
                    } else {
                        total = totbuf;  // one cluster, one GPU: the cluster partial is the total
                    }
                    if (prof) tick(6);  // leader: cluster sum + hops N / B (thread 0 is a row lane)
                    __syncthreads();  // total complete
                    // hop C: the total -> every member's totbuf
                    for (int q = tid; q < (CS - 1) * NCH; q += BLOCK) {
                        const uint32_t r = 1u + (uint32_t)(q / NCH), j = (uint32_t)(q % NCH);
                        const uint4 v = reinterpret_cast<const uint4*>(total)[j];
                        cl_st_async16(cl_mapa(cl_smem_u32(totbuf) + 16u * j, r), v, cl_mapa(mb_tot, r));
                    }
                }
                if (!leader) total = totbuf;
            }
            for (int j = tid; j < NV; j += BLOCK) Wsm[(j / AW) * WS + j % AW] += total[j];
            }
            if (prof) tick(3);
            __syncthreads();
            if (prof) tick(4);
        }
    }

    if (resident && active) {
        a.ep_steps[i] = ep;
        a.actions[i] = act;
#pragma unroll
        for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
    }
    if (pe_smem && active) {
        R* Wm = static_cast<R*>(a.W);
        for (int j = 0; j < FA; ++j) Wm[(int64_t)j * N + i] = Wpe[(size_t)j * BLOCK + tid];
    }
    if (TRACE && active) {
        R* Z = static_cast<R*>(a.z);
        for (int j = 0; j < FA; ++j) Z[(int64_t)j * N + i] = red[(size_t)j * cap + tid];
    }
    if (SHAREDW && b == 0) {
        for (int j = tid; j < FA; j += BLOCK) static_cast<R*>(a.W)[j] = Wsm[(j / AW) * WS + j % AW];
        if (FX) {  // every CTA holds the same running sums of the world table (one GPU: of the only table, saved below)
            for (int j = tid; j < 2 * NV; j += BLOCK) sy.prev[j] = prevs[(2 + j / NV) * NVP8 + j % NV];
        }
    }
    if (SHAREDW && FX && b < (pe.world > 1 ? (sy.ngroups < G ? sy.ngroups : G) : 1)) {  // group leaders: their group's table
        for (int j = tid; j < 2 * NV; j += BLOCK) sy.prev[(size_t)(2 + 2 * b) * NV + j] = prevs[(j / NV) * NVP8 + j % NV];
    }
    if (SHAREDW && CS > 1) cl_sync();  // nobody leaves while a bulk copy may still read or write its shared memory
}

}  // namespace rsrl
