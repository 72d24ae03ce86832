// f4tc_inst.cu — instantiations + host launchers of the tcgen05 path (f4tc.cuh): order-7 Fourier, 4-D domains, f32
#include <cstdlib>
#include "f4tc_launch.h"
#include "f4tc.cuh"

namespace rsrl {

template <int DOM, int PHASE, bool EXT>
static cudaError_t q_one(const StepArgs& a, const F4Args& fa, int n_tiles, int grid, cudaStream_t st) {
    auto kern = f4tc_q_kernel<DOM, PHASE, EXT>;
    constexpr size_t smem = F4tcEnvSmem<Domain<DOM>::A>::bytes;
    static bool configured_v[64] = {false};  // per device: the opt-in is a per-device attribute of the function
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_v[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, 288, smem, st>>>(a, fa, n_tiles);
    return cudaGetLastError();
}

template <int DOM, bool EXT>
static cudaError_t env_one(const StepArgs& a, const F4Args& fa, int n_tiles, int grid, cudaStream_t st) {
    cudaError_t e = q_one<DOM, 0, EXT>(a, fa, n_tiles, grid, st);
    if (e != cudaSuccess) return e;
    f4tc_phys_kernel<DOM, EXT><<<(unsigned)((a.n + 127) / 128), 128, 0, st>>>(a, fa);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return q_one<DOM, 1, EXT>(a, fa, n_tiles, grid, st);
}

// Q(s_t) -> action + physics -> Q(s') + TD error: three launches
cudaError_t launch_f4tc_env(int domain, bool ext, const StepArgs& a, const F4Args& fa, int n_tiles, int grid, cudaStream_t st) {
    if (domain == RSRL_CART_POLE) return ext ? env_one<RSRL_CART_POLE, true>(a, fa, n_tiles, grid, st) : env_one<RSRL_CART_POLE, false>(a, fa, n_tiles, grid, st);
    if (domain == RSRL_ACROBOT) return ext ? env_one<RSRL_ACROBOT, true>(a, fa, n_tiles, grid, st) : env_one<RSRL_ACROBOT, false>(a, fa, n_tiles, grid, st);
    return cudaErrorInvalidDeviceFunction;
}

template <int DOM>
static cudaError_t dw_one(int64_t n, const float* tabs, const void* coef, const int32_t* actions, int grid, void* partials,
                          Counters* counters, long long* prof, cudaStream_t st) {
    auto kern = f4tc_dw_kernel<DOM>;
    constexpr size_t smem = F4tcDwSmem<Domain<DOM>::A>::bytes;
    static bool configured_v[64] = {false};  // per device: the opt-in is a per-device attribute of the function
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_v[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, 544, smem, st>>>(n, tabs, static_cast<const float*>(coef), actions, static_cast<float*>(partials), counters, prof);
    return cudaGetLastError();
}

cudaError_t launch_f4tc_dw(int domain, int64_t n, const float* tabs, const void* coef, const int32_t* actions, int grid,
                           void* partials, Counters* counters, long long* prof, cudaStream_t st) {
    if (domain == RSRL_CART_POLE) return dw_one<RSRL_CART_POLE>(n, tabs, coef, actions, grid, partials, counters, prof, st);
    if (domain == RSRL_ACROBOT) return dw_one<RSRL_ACROBOT>(n, tabs, coef, actions, grid, partials, counters, prof, st);
    return cudaErrorInvalidDeviceFunction;
}

cudaError_t launch_f4tc_reduce(const void* partials, int n_partials, int fa, void* W, void* dW_out, cudaStream_t st) {
    f4tc_reduce_kernel<<<(fa + 31) / 32, 1024, 0, st>>>(static_cast<const float*>(partials), n_partials, fa, static_cast<float*>(W),
                                                      static_cast<float*>(dW_out));
    return cudaGetLastError();
}

}  // namespace rsrl
