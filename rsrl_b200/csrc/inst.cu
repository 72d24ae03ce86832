// inst.cu — explicit kernel instantiations for one (dtype, domain) pair.
// Compiled six times: -DRSRL_REAL=float|double -DRSRL_DOM=0|1|2 -DRSRL_SUFFIX=f32_d0 ...
#include <cstring>
#include "launch.h"

namespace rsrl {

typedef RSRL_REAL R;
constexpr int DOM = RSRL_DOM;

// dynamic shared memory opt-in is a per-device attribute of the function: remember what each device was given
static bool smem_configured(size_t* table, size_t smem, int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev & 63;
    return smem <= table[dev & 63];
}

template <int BASIS, int P, int AW, int MODE, bool EXT>
static cudaError_t launch_one(const StepArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    auto kern = fused_step_kernel<R, DOM, BASIS, P, AW, MODE, EXT>;
    static size_t configured[64] = {0};
    int dev;
    if (!smem_configured(configured, smem, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    kern<<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}

template <int BASIS, int P, int AW>
static cudaError_t launch_mode(int mode, bool ext, const StepArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    if (mode == RSRL_SHARED) return ext ? launch_one<BASIS, P, AW, RSRL_SHARED, true>(a, grid, block, smem, st)
                                        : launch_one<BASIS, P, AW, RSRL_SHARED, false>(a, grid, block, smem, st);
    return ext ? launch_one<BASIS, P, AW, RSRL_PER_ENV, true>(a, grid, block, smem, st)
               : launch_one<BASIS, P, AW, RSRL_PER_ENV, false>(a, grid, block, smem, st);
}

template <int BASIS, int P, int AW, int MODE>
static cudaError_t persist_one(const StepArgs& a, int k_steps, const SyncArgs& sy, const PeerArgs& pe, int grid, int block, size_t smem,
                               cudaStream_t st, int* max_clusters) {
    auto kern = persistent_kernel<R, DOM, BASIS, P, AW, MODE>;
    static size_t configured[64] = {0};
    static int coop_ok[64] = {0};  // 0 unknown, 1 cooperative + cluster launch accepted, -1 rejected by this driver
    int dev;
    if (!smem_configured(configured, smem, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    if (MODE != RSRL_PER_ENV && (grid > 1 || max_clusters)) {
        // SHARED weights: the CTAs wait for each other every step, so the whole grid has to be co-resident.
        const int cs = sy.cluster_size;
        if (cs > 1) {
            if (cs > 8) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
                if (e != cudaSuccess) return e;
            }
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute at[2];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            at[1].id = cudaLaunchAttributeCooperative;
            at[1].val.cooperative = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (max_clusters) return cudaOccupancyMaxActiveClusters(max_clusters, (const void*)kern, &cfg);
            if (coop_ok[dev] >= 0) {
                cfg.numAttrs = 2;
                cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, k_steps, sy, pe);
                if (e == cudaSuccess) { coop_ok[dev] = 1; return e; }
                if (coop_ok[dev] == 1) return e;
                cudaGetLastError();   // this driver does not combine the two attributes: co-residency was checked at create
                coop_ok[dev] = -1;    // (cudaOccupancyMaxActiveClusters), launch as plain clusters
                cfg.numAttrs = 1;
            }
            return cudaLaunchKernelEx(&cfg, kern, a, k_steps, sy, pe);
        }
        int per_sm = 0, sms = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem);
        if (e != cudaSuccess) return e;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (max_clusters) { *max_clusters = per_sm * sms; return cudaSuccess; }
        if ((long long)per_sm * sms < grid) return cudaErrorCooperativeLaunchTooLarge;
        void* args[] = {(void*)&a, (void*)&k_steps, (void*)&sy, (void*)&pe};
        return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(block), args, smem, st);
    }
    kern<<<grid, block, smem, st>>>(a, k_steps, sy, pe);
    return cudaGetLastError();
}

template <int BASIS, int P, int MODE, bool EXT>
static cudaError_t two_one(const StepArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    auto kern = two_table_step_kernel<R, DOM, BASIS, P, MODE, EXT>;
    static size_t configured[64] = {0};
    int dev;
    if (!smem_configured(configured, smem, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    kern<<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}

template <int BASIS, int P>
static cudaError_t rollout_one(const RolloutArgs& ra, cudaStream_t st) {
    const int block = 128, grid = (int)((ra.n + block - 1) / block);
    rollout_kernel<R, DOM, BASIS, P><<<grid, block, 0, st>>>(ra);
    return cudaGetLastError();
}

template <int BASIS, int P, int AW>
static cudaError_t eval_one(const EvalArgs& e, cudaStream_t st) {
    const int block = 128;
    const int grid = (int)((e.n + block - 1) / block);
    basis_eval_kernel<R, DOM, BASIS, P, AW><<<grid, block, 0, st>>>(e.mode, e.n, e.states, static_cast<const R*>(e.W),
                                                                     e.w_env_stride, e.out, e.act_out, e.pol, e.draw,
                                                                     e.env_offset, e.counters);
    return cudaGetLastError();
}

// (basis, order) pairs built for this domain.  MountainCar (D = 2) unrolls any order up to 7;
// the D = 4 domains are fully unrolled only up to order 3 here (Fourier(7) on Acrobot = 4096
// features takes the tiled path).
#if defined(RSRL_EMPTY)
#define RSRL_COMBOS(X)  // development builds: this domain is left out (RSRL_BUILD_DOMAINS)
#elif RSRL_DOM == 0
#define RSRL_COMBOS(X) X(RSRL_FOURIER, 1) X(RSRL_FOURIER, 2) X(RSRL_FOURIER, 3) X(RSRL_FOURIER, 5) X(RSRL_FOURIER, 7) \
                       X(RSRL_POLYNOMIAL, 2) X(RSRL_POLYNOMIAL, 3)
#else
#define RSRL_COMBOS(X) X(RSRL_FOURIER, 2) X(RSRL_FOURIER, 3) X(RSRL_POLYNOMIAL, 2)
#endif

#define RSRL_CAT_(a, b) a##b
#define RSRL_CAT(a, b) RSRL_CAT_(a, b)

bool RSRL_CAT(has_static_, RSRL_SUFFIX)(const BasisKey& k) {
#define X(B, P) if (k.basis == B && k.order == P) return true;
    RSRL_COMBOS(X)
#undef X
    return false;
}

#if !defined(RSRL_EMPTY)
// run-time-order kernels (dyn.cuh): one warp per CTA
template <int AW>
static cudaError_t dyn_launch(const BasisKey& k, int mode, bool ext, const StepArgs& a, int grid, cudaStream_t st) {
    const DynBasis b = {k.basis, k.order};
    if (mode == RSRL_SHARED) {
        if (ext) dyn_step_kernel<R, DOM, AW, RSRL_SHARED, true><<<grid, 32, 0, st>>>(a, b);
        else dyn_step_kernel<R, DOM, AW, RSRL_SHARED, false><<<grid, 32, 0, st>>>(a, b);
    } else {
        if (ext) dyn_step_kernel<R, DOM, AW, RSRL_PER_ENV, true><<<grid, 32, 0, st>>>(a, b);
        else dyn_step_kernel<R, DOM, AW, RSRL_PER_ENV, false><<<grid, 32, 0, st>>>(a, b);
    }
    return cudaGetLastError();
}
template <int AW>
static cudaError_t dyn_eval(const BasisKey& k, const EvalArgs& e, cudaStream_t st) {
    const DynBasis b = {k.basis, k.order};
    const int block = 128, grid = (int)((e.n + block - 1) / block);
    dyn_eval_kernel<R, DOM, AW><<<grid, block, 0, st>>>(e.mode, e.n, e.states, static_cast<const R*>(e.W), e.w_env_stride, e.out, e.act_out,
                                                        e.pol, e.draw, e.env_offset, e.counters, b);
    return cudaGetLastError();
}
#endif

cudaError_t RSRL_CAT(launch_fused_, RSRL_SUFFIX)(const BasisKey& k, int mode, bool ext, const StepArgs& a, int grid,
                                                 int block, size_t smem, cudaStream_t st) {
    constexpr int A = Domain<DOM>::A;
#define X(B, P)                                                                                  \
    if (k.basis == B && k.order == P) {                                                          \
        if (k.aw == A) return launch_mode<B, P, A>(mode, ext, a, grid, block, smem, st);         \
        if (k.aw == 1) return launch_mode<B, P, 1>(mode, ext, a, grid, block, smem, st);         \
    }
    RSRL_COMBOS(X)
#undef X
#if !defined(RSRL_EMPTY)
    if (k.basis != RSRL_TILE_CODING && k.order >= 1 && k.order <= kDynMaxOrder && block == 32 && smem == 0) {
        if (k.aw == A) return dyn_launch<A>(k, mode, ext, a, grid, st);
        if (k.aw == 1) return dyn_launch<1>(k, mode, ext, a, grid, st);
    }
#endif
    return cudaErrorInvalidDeviceFunction;
}

cudaError_t RSRL_CAT(launch_persist_, RSRL_SUFFIX)(const BasisKey& k, int mode, const StepArgs& a, int k_steps, const SyncArgs& sy,
                                                   const PeerArgs& pe, int grid, int block, size_t smem, cudaStream_t st, int* max_clusters) {
    constexpr int A = Domain<DOM>::A;
#define RSRL_PERSIST(B, P, AWV)                                                                                   \
    (mode == RSRL_SHARED ? persist_one<B, P, AWV, RSRL_SHARED>(a, k_steps, sy, pe, grid, block, smem, st, max_clusters)         \
     : mode == RSRL_PER_ENV ? persist_one<B, P, AWV, RSRL_PER_ENV>(a, k_steps, sy, pe, grid, block, smem, st, max_clusters)     \
                            : persist_one<B, P, AWV, kModeSharedTrace>(a, k_steps, sy, pe, grid, block, smem, st, max_clusters))
#define X(B, P)                                   \
    if (k.basis == B && k.order == P) {           \
        if (k.aw == A) return RSRL_PERSIST(B, P, A); \
        if (k.aw == 1) return RSRL_PERSIST(B, P, 1); \
    }
    RSRL_COMBOS(X)
#undef X
    return cudaErrorInvalidDeviceFunction;
}

cudaError_t RSRL_CAT(launch_two_, RSRL_SUFFIX)(const BasisKey& k, int mode, bool ext, const StepArgs& a, int grid, int block, size_t smem,
                                               cudaStream_t st) {
#if RSRL_DOM == 0 && !defined(RSRL_EMPTY)
    // GreedyGQ / A2C are instantiated for MountainCar (the domain of examples/greedy_gq.rs and examples/a2c.rs): the D = 4 bases
    // would add ten minutes of compile time for 256-feature fully unrolled kernels nobody asked for
#define X(B, P)                                                                                                                  \
    if (k.basis == B && k.order == P) {                                                                                          \
        if (mode == RSRL_SHARED) return ext ? two_one<B, P, RSRL_SHARED, true>(a, grid, block, smem, st)                          \
                                            : two_one<B, P, RSRL_SHARED, false>(a, grid, block, smem, st);                        \
        return ext ? two_one<B, P, RSRL_PER_ENV, true>(a, grid, block, smem, st) : two_one<B, P, RSRL_PER_ENV, false>(a, grid, block, smem, st); \
    }
    RSRL_COMBOS(X)
#undef X
#endif
    return cudaErrorInvalidDeviceFunction;
}

cudaError_t RSRL_CAT(launch_rollout_, RSRL_SUFFIX)(const BasisKey& k, const RolloutArgs& ra, cudaStream_t st) {
#define X(B, P) \
    if (k.basis == B && k.order == P) return rollout_one<B, P>(ra, st);
    RSRL_COMBOS(X)
#undef X
    return cudaErrorInvalidDeviceFunction;
}

cudaError_t RSRL_CAT(launch_eval_, RSRL_SUFFIX)(const BasisKey& k, const EvalArgs& e, cudaStream_t st) {
    constexpr int A = Domain<DOM>::A;
#define X(B, P)                                                   \
    if (k.basis == B && k.order == P) {                           \
        if (k.aw == A) return eval_one<B, P, A>(e, st);           \
        if (k.aw == 1) return eval_one<B, P, 1>(e, st);           \
    }
    RSRL_COMBOS(X)
#undef X
#if !defined(RSRL_EMPTY)
    if (k.basis != RSRL_TILE_CODING && k.order >= 1 && k.order <= kDynMaxOrder) {
        if (k.aw == A) return dyn_eval<A>(k, e, st);
        if (k.aw == 1) return dyn_eval<1>(k, e, st);
    }
#endif
    return cudaErrorInvalidDeviceFunction;
}

}  // namespace rsrl
