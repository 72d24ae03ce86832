// tile_inst.cu — instantiations of the TileCoding kernels for one dtype (-DRSRL_REAL=float|double -DRSRL_SUFFIX=f32|f64)
#include "launch.h"
#include "tile.cuh"

namespace rsrl {

typedef RSRL_REAL R;
#define RSRL_CAT_(a, b) a##b
#define RSRL_CAT(a, b) RSRL_CAT_(a, b)

template <int DOM, int AW, bool EXT>
static cudaError_t tile_one(const StepArgs& a, int k_steps, const TileArgs& ta, int grid, int block, size_t smem, cudaStream_t st) {
    // dense kernel: 512 threads (128 registers each) or up to 1024 (64 registers: spills, but twice the warps to hide the f64 latencies)
    // TMAX = 8 instantiation: half the unrolled tiling loops for the common <= 8 tilings
    auto kern = !ta.dense ? tile_persistent_kernel<R, DOM, AW, EXT>
              : ta.tp.n_tilings <= 8
                    ? (block > 768 ? tile_dense_kernel<R, DOM, AW, EXT, 1024, 8> : block > 512 ? tile_dense_kernel<R, DOM, AW, EXT, 768, 8> : tile_dense_kernel<R, DOM, AW, EXT, 512, 8>)
                    : (block > 768 ? tile_dense_kernel<R, DOM, AW, EXT, 1024> : block > 512 ? tile_dense_kernel<R, DOM, AW, EXT, 768> : tile_dense_kernel<R, DOM, AW, EXT, 512>);
    static size_t configured_v[64][7] = {{0}};  // per device (the opt-in is a per-device attribute) and kernel variant
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& configured = configured_v[dev & 63][!ta.dense ? 0 : (block > 768 ? 3 : block > 512 ? 2 : 1) + (ta.tp.n_tilings <= 8 ? 3 : 0)];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    if (grid > 1) {
        int per_sm = 0, dev = 0, sms = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem);
        if (e != cudaSuccess) return e;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if ((long long)per_sm * sms < grid) return cudaErrorCooperativeLaunchTooLarge;
        void* args[] = {(void*)&a, (void*)&k_steps, (void*)&ta};
        return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(block), args, smem, st);
    }
    kern<<<grid, block, smem, st>>>(a, k_steps, ta);
    return cudaGetLastError();
}

template <int DOM>
static cudaError_t tile_dom(int aw, bool ext, const StepArgs& a, int k, const TileArgs& ta, int grid, int block, size_t smem, cudaStream_t st) {
    constexpr int A = Domain<DOM>::A;
    if (aw == A) return ext ? tile_one<DOM, A, true>(a, k, ta, grid, block, smem, st) : tile_one<DOM, A, false>(a, k, ta, grid, block, smem, st);
    return ext ? tile_one<DOM, 1, true>(a, k, ta, grid, block, smem, st) : tile_one<DOM, 1, false>(a, k, ta, grid, block, smem, st);
}

cudaError_t RSRL_CAT(launch_tile_persist_, RSRL_SUFFIX)(int domain, int aw, bool ext, const StepArgs& a, int k, const TileArgs& ta,
                                                        int grid, int block, size_t smem, cudaStream_t st) {
    if (domain == RSRL_MOUNTAIN_CAR) return tile_dom<RSRL_MOUNTAIN_CAR>(aw, ext, a, k, ta, grid, block, smem, st);
    if (domain == RSRL_CART_POLE) return tile_dom<RSRL_CART_POLE>(aw, ext, a, k, ta, grid, block, smem, st);
    return tile_dom<RSRL_ACROBOT>(aw, ext, a, k, ta, grid, block, smem, st);
}

template <int DOM, int AW>
static cudaError_t tile_eval_one(const EvalArgs& e, const TileParams& tp, cudaStream_t st) {
    const int block = 128, grid = (int)((e.n + block - 1) / block);
    tile_eval_kernel<R, DOM, AW><<<grid, block, 0, st>>>(e.mode, e.n, e.states, static_cast<const R*>(e.W), tp, e.out, e.act_out, e.pol,
                                                         e.draw, e.env_offset, e.counters);
    return cudaGetLastError();
}

cudaError_t RSRL_CAT(launch_tile_eval_, RSRL_SUFFIX)(int domain, int aw, const EvalArgs& e, const TileParams& tp, cudaStream_t st) {
    if (domain == RSRL_MOUNTAIN_CAR) return aw == 1 ? tile_eval_one<RSRL_MOUNTAIN_CAR, 1>(e, tp, st) : tile_eval_one<RSRL_MOUNTAIN_CAR, 3>(e, tp, st);
    if (domain == RSRL_CART_POLE) return aw == 1 ? tile_eval_one<RSRL_CART_POLE, 1>(e, tp, st) : tile_eval_one<RSRL_CART_POLE, 2>(e, tp, st);
    return aw == 1 ? tile_eval_one<RSRL_ACROBOT, 1>(e, tp, st) : tile_eval_one<RSRL_ACROBOT, 3>(e, tp, st);
}

}  // namespace rsrl
