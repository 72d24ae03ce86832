// f4_inst.cu — instantiations of the large-basis (4-D, order 5 / 7) kernels for one dtype
// (-DRSRL_REAL=float|double -DRSRL_SUFFIX=f32|f64)
#include "launch.h"
#include "fourier4.cuh"

namespace rsrl {

typedef RSRL_REAL R;
#define RSRL_CAT_(a, b) a##b
#define RSRL_CAT(a, b) RSRL_CAT_(a, b)

template <int DOM, int P, bool EXT>
static cudaError_t env_one(const StepArgs& a, const F4Args& fa, int grid, int block, size_t smem, cudaStream_t st) {
    auto kern = f4_env_kernel<R, DOM, RSRL_FOURIER, P, Domain<DOM>::A, EXT>;
    static size_t configured[64] = {0};  // the opt-in is a per-device attribute of the function
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    kern<<<grid, block, smem, st>>>(a, fa);
    return cudaGetLastError();
}

#define F4_DISPATCH(CALL)                                                            \
    if (domain == RSRL_CART_POLE && order == 5) { CALL(RSRL_CART_POLE, 5) }          \
    if (domain == RSRL_CART_POLE && order == 7) { CALL(RSRL_CART_POLE, 7) }          \
    if (domain == RSRL_ACROBOT && order == 5) { CALL(RSRL_ACROBOT, 5) }              \
    if (domain == RSRL_ACROBOT && order == 7) { CALL(RSRL_ACROBOT, 7) }              \
    return cudaErrorInvalidDeviceFunction;

cudaError_t RSRL_CAT(launch_f4_env_, RSRL_SUFFIX)(int domain, int order, bool ext, const StepArgs& a, const F4Args& fa, int grid,
                                                  int block, size_t smem, cudaStream_t st) {
#define CALL(D, P) return ext ? env_one<D, P, true>(a, fa, grid, block, smem, st) : env_one<D, P, false>(a, fa, grid, block, smem, st);
    F4_DISPATCH(CALL)
#undef CALL
}

cudaError_t RSRL_CAT(launch_f4_dw_, RSRL_SUFFIX)(int domain, int order, int64_t n, const double* from_states, const void* coef,
                                                 const int32_t* actions, int n_seg, void* partials, cudaStream_t st) {
#define CALL(D, P)                                                                                                          \
    {                                                                                                                       \
        const int threads = (((P + 1) * (P + 1) * 4) + 31) / 32 * 32;                                                       \
        f4_dw_kernel<R, D, RSRL_FOURIER, P, Domain<D>::A><<<dim3(P + 1, n_seg), threads, 0, st>>>(                          \
            n, from_states, static_cast<const R*>(coef), actions, n_seg, static_cast<R*>(partials));                        \
        return cudaGetLastError();                                                                                          \
    }
    F4_DISPATCH(CALL)
#undef CALL
}

cudaError_t RSRL_CAT(launch_f4_eval_, RSRL_SUFFIX)(int domain, int order, const EvalArgs& e, cudaStream_t st) {
#define CALL(D, P)                                                                                                          \
    {                                                                                                                       \
        const int block = 128, grid = (int)((e.n + block - 1) / block);                                                     \
        const size_t smem = (size_t)2 * P * 2 * block * sizeof(R);                                                          \
        f4_eval_kernel<R, D, RSRL_FOURIER, P, Domain<D>::A><<<grid, block, smem, st>>>(                                     \
            e.mode, e.n, e.states, static_cast<const R*>(e.W), e.out, e.act_out, e.pol, e.draw, e.env_offset, e.counters);  \
        return cudaGetLastError();                                                                                          \
    }
    F4_DISPATCH(CALL)
#undef CALL
}

}  // namespace rsrl
