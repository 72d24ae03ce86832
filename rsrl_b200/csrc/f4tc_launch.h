// f4tc_launch.h — host launchers of the tcgen05 path of the order-7 Fourier basis (f4tc.cuh), f32 only
#pragma once
#include "launch.h"

namespace rsrl {
cudaError_t launch_f4tc_env(int domain, bool ext, const StepArgs&, const F4Args&, int n_tiles, int grid, cudaStream_t);
cudaError_t launch_f4tc_dw(int domain, int64_t n, const float* tabs, const void* coef, const int32_t* actions, int grid,
                           void* partials, Counters* counters, long long* prof, cudaStream_t);
cudaError_t launch_f4tc_reduce(const void* partials, int n_partials, int fa, void* W, void* dW_out, cudaStream_t);
}  // namespace rsrl
