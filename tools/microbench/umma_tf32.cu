// umma_tf32.cu — standalone check of the tcgen05 building blocks used by rsrl_b200/csrc/f4tc.cuh:
// thread-written operand tiles in the canonical no-swizzle UMMA layouts (K-major and MN-major),
// shared-memory descriptors, kind::tf32 MMAs into TMEM, the 3xTF32 split (hi*hi + lo*hi + hi*lo),
// tcgen05.ld epilogue; reports max error against an f64 host product and cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tf32 umma_tf32.cu && ./umma_tf32
// Every wait is bounded: a wrong descriptor produces a wrong number or a timeout flag, never a hang.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from TMEM (TS mode): 128 lanes x K columns of 32-bit tf32 values
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
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
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity, int max_iter) {
    for (int i = 0; i < max_iter; ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// canonical no-swizzle layouts, R rows (M or N) x K, 4-byte elements; byte offset of element (r, k)
__host__ __device__ inline uint32_t off_kmajor(int R, int r, int k, bool swap, int pad = 0) {
    const uint32_t mn_stride = swap ? (uint32_t)(64 / 4) * 128u : 128u;   // normal: row groups adjacent, K chunks far apart
    const uint32_t k_stride = swap ? 128u : (uint32_t)(R / 8) * 128u + (uint32_t)pad;
    return (uint32_t)(k / 4) * k_stride + (uint32_t)(r / 8) * mn_stride + (uint32_t)(r % 8) * 16u + (uint32_t)(k % 4) * 4u;
}
__host__ __device__ inline uint32_t off_mnmajor(int R, int r, int k) {
    return (uint32_t)(r / 4) * 128u + (uint32_t)(k / 8) * (uint32_t)(R / 4) * 128u + (uint32_t)(k % 8) * 16u + (uint32_t)(r % 4) * 4u;
}

struct Params {
    const float* A;  // [128][64]
    const float* B;  // [N][64]
    float* D;        // [128][N]
    int N, passes, a_mn, b_mn, swap, swap_desc, reps, rna, pad, ts;
    long long* cycles;
    int* flags;
};

__global__ void __launch_bounds__(128) umma_test_kernel(Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int M = 128, K = 64;
    const int N = p.N;
    float* Ahi = reinterpret_cast<float*>(smem);
    float* Alo = Ahi + M * K + 16 * 16;
    float* Bhi = Alo + M * K + 16 * 16;
    float* Blo = Bhi + N * K + 16 * 16;
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int idx = tid; idx < M * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        const float x = p.A[idx];
        const float hi = p.rna ? tf32_rna(x) : __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const uint32_t o = p.a_mn ? off_mnmajor(M, r, k) : off_kmajor(M, r, k, p.swap, p.pad);
        Ahi[o / 4] = hi;
        Alo[o / 4] = p.rna ? tf32_rna(x - hi) : x - hi;
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        const float x = p.B[idx];
        const float hi = p.rna ? tf32_rna(x) : __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const uint32_t o = p.b_mn ? off_mnmajor(N, r, k) : off_kmajor(N, r, k, p.swap, p.pad);
        Bhi[o / 4] = hi;
        Blo[o / 4] = p.rna ? tf32_rna(x - hi) : x - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core (async proxy)

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (p.ts) {  // thread = row: A[row][0..63] hi -> columns 256.., lo -> columns 320..
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < K; k0 += 8) {
            float hi[8], lo[8];
            for (int x = 0; x < 8; ++x) {
                const float v = p.A[tid * K + k0 + x];
                hi[x] = tf32_rna(v);
                lo[x] = tf32_rna(v - hi[x]);
            }
            tmem_st8(lane_base + 256 + k0, hi);
            tmem_st8(lane_base + 320 + k0, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    // K-major: LBO = K-direction stride between the two 16-byte chunks of one K = 8 step, SBO = stride between 8-row groups
    // MN-major: SBO = stride between 4-row groups, LBO = stride between 8-k groups
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step;
    if (p.a_mn) { a_sbo = 128; a_lbo = (M / 4) * 128; a_step = a_lbo; }
    else { a_sbo = p.swap ? (K / 4) * 128 : 128; a_lbo = p.swap ? 128 : (M / 8) * 128 + p.pad; a_step = 2 * a_lbo; }
    if (p.b_mn) { b_sbo = 128; b_lbo = (uint32_t)(N / 4) * 128; b_step = b_lbo; }
    else { b_sbo = p.swap ? (K / 4) * 128 : 128; b_lbo = p.swap ? 128 : (uint32_t)(N / 8) * 128 + p.pad; b_step = 2 * b_lbo; }
    if (p.swap_desc) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }

    long long t0 = 0, t1 = 0;
    bool ok = true;
    if (tid == 0) {
        // descriptors precomputed: the timed loop is MMA issue only (a single thread that also builds descriptors is
        // issue-bound at ~110-140 cycles per MMA)
        uint64_t adv[3][8], bdv[3][8];
        uint32_t atm[3][8];
        for (int pass = 0; pass < 3; ++pass)
            for (int ks = 0; ks < 8; ++ks) {
                const float* As = pass == 1 ? Alo : Ahi;
                const float* Bs = pass == 2 ? Blo : Bhi;
                adv[pass][ks] = make_desc(smem_u32(As) + ks * a_step, a_lbo, a_sbo);
                bdv[pass][ks] = make_desc(smem_u32(Bs) + ks * b_step, b_lbo, b_sbo);
                atm[pass][ks] = tmem + (pass == 1 ? 320 : 256) + ks * 8;
            }
        t0 = clock64();
        if (p.ts) {
            for (int rep = 0; rep < p.reps; ++rep) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    if (pass < p.passes) {
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) umma_tf32_ts(tmem, atm[pass][ks], bdv[pass][ks], idesc, (rep | pass | ks) != 0 ? 1u : 0u);
                    }
                }
            }
        } else {
            for (int rep = 0; rep < p.reps; ++rep) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    if (pass < p.passes) {
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) umma_tf32(tmem, adv[pass][ks], bdv[pass][ks], idesc, (rep | pass | ks) != 0 ? 1u : 0u);
                    }
                }
            }
        }
        umma_commit(smem_u32(&bar));
        ok = mbar_wait_bounded(smem_u32(&bar), 0, 20000000);
        t1 = clock64();
        p.cycles[0] = t1 - t0;
        if (!ok) atomicExch(p.flags, 1);
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) p.D[(size_t)tid * N + c * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    const int M = 128, K = 64;
    std::vector<float> hA(M * K), hBfull(256 * K);
    srand(1);
    for (auto& x : hA) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : hBfull) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    long long* dC;
    int* dF;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hBfull.size() * 4); cudaMalloc(&dD, M * 256 * 4);
    cudaMalloc(&dC, 8); cudaMalloc(&dF, 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hBfull.data(), hBfull.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);

    struct Case { const char* name; int N, passes, a_mn, b_mn, swap, swap_desc, reps, rna, pad, ts; };
    const Case cases[] = {
        {"SS K-major N=192 3xTF32 (correctness)", 192, 3, 0, 0, 0, 0, 1, 1, 0, 0},
        {"TS (A in TMEM) N=192 3xTF32 (correctness)", 192, 3, 0, 0, 0, 0, 1, 1, 0, 1},
        {"TS N=96 3xTF32 padded B (correctness)", 96, 3, 0, 0, 0, 0, 1, 1, 16, 1},
        {"SS N=192 timing reps=64", 192, 3, 0, 0, 0, 0, 64, 1, 0, 0},
        {"TS N=192 timing reps=64", 192, 3, 0, 0, 0, 0, 64, 1, 0, 1},
        {"SS N=96 timing reps=64", 96, 3, 0, 0, 0, 0, 64, 1, 0, 0},
        {"TS N=96 timing reps=64", 96, 3, 0, 0, 0, 0, 64, 1, 0, 1},
        {"TS N=64 timing reps=64", 64, 3, 0, 0, 0, 0, 64, 1, 0, 1},
    };
    for (const Case& c : cases) {
        Params p{dA, dB, dD, c.N, c.passes, c.a_mn, c.b_mn, c.swap, c.swap_desc, c.reps, c.rna, c.pad, c.ts, dC, dF};
        cudaMemset(dF, 0, 4);
        cudaMemset(dD, 0, M * 256 * 4);
        const size_t smem = (size_t)(2 * M * K + 2 * c.N * K) * 4 + 1024 + 4 * 16 * 64;
        umma_test_kernel<<<1, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-52s CUDA error: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        std::vector<float> hD((size_t)M * c.N);
        long long cyc; int flag;
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&flag, dF, 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < c.N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * (double)hBfull[n * K + k];
                maxerr = fmax(maxerr, fabs((double)hD[(size_t)m * c.N + n] / c.reps - ref));
            }
        const int n_mma = c.reps * c.passes * (K / 8);
        printf("%-52s max|err| %.3e  timeout %d  cycles %lld (%d MMAs, %.1f cyc/MMA)\n", c.name, maxerr, flag, cyc, n_mma, (double)cyc / n_mma);
    }
    return 0;
}
