// launch.h — type-erased host launchers for the templated kernels (one instantiation TU per dtype/domain).
#pragma once
#include <cuda_runtime.h>
#include "persistent.cuh"
#include "twotable.cuh"
#include "dyn.cuh"

namespace rsrl {

struct BasisKey {
    int dtype;   // rsrl_dtype_t
    int domain;  // rsrl_domain_t
    int basis;   // rsrl_basis_t
    int order;   // P
    int aw;      // weight columns: n_actions, or 1 for TD prediction
};

struct EvalArgs {
    int mode;  // 0 features, 1 Q, 2 sample, 3 find_max
    int64_t n;
    const double* states;
    const void* W;
    int64_t w_env_stride;
    double* out;
    int32_t* act_out;
    PolicyParams pol;
    uint64_t draw;
    int64_t env_offset;
    Counters* counters;
};

// returns cudaErrorInvalidDeviceFunction when the combination was not instantiated
typedef cudaError_t (*fused_launch_fn)(const BasisKey&, int weight_mode, bool ext, const StepArgs&, int grid, int block,
                                       size_t smem, cudaStream_t);
typedef cudaError_t (*eval_launch_fn)(const BasisKey&, const EvalArgs&, cudaStream_t);
// persistent K-step kernel; SHARED mode is launched cooperatively (all CTAs co-resident)
// max_clusters != nullptr: no launch, *max_clusters = co-resident clusters (CTAs when cluster_size == 1) of this shape
typedef cudaError_t (*persist_launch_fn)(const BasisKey&, int weight_mode, const StepArgs&, int k_steps, const SyncArgs&,
                                         const PeerArgs&, int grid, int block, size_t smem, cudaStream_t, int* max_clusters);

// two-table agents (GreedyGQ, A2C) and Domain::rollout
typedef cudaError_t (*two_launch_fn)(const BasisKey&, int weight_mode, bool ext, const StepArgs&, int grid, int block, size_t smem, cudaStream_t);
typedef cudaError_t (*rollout_launch_fn)(const BasisKey&, const RolloutArgs&, cudaStream_t);

// has_static_*: the (basis, order) pair is instantiated as templates in this unit; otherwise launch_fused_* / launch_eval_* run the
// run-time-order kernels of dyn.cuh (launch_fused_* then wants block == 32 and no shared memory)
#define RSRL_DECL_INST(SUFFIX)                                                                                      \
    bool has_static_##SUFFIX(const BasisKey&);                                                                      \
    cudaError_t launch_two_##SUFFIX(const BasisKey&, int, bool, const StepArgs&, int, int, size_t, cudaStream_t);   \
    cudaError_t launch_rollout_##SUFFIX(const BasisKey&, const RolloutArgs&, cudaStream_t);                         \
    cudaError_t launch_fused_##SUFFIX(const BasisKey&, int, bool, const StepArgs&, int, int, size_t, cudaStream_t); \
    cudaError_t launch_eval_##SUFFIX(const BasisKey&, const EvalArgs&, cudaStream_t);                                 \
    cudaError_t launch_persist_##SUFFIX(const BasisKey&, int, const StepArgs&, int, const SyncArgs&, const PeerArgs&, int, int, size_t, cudaStream_t, int*);
RSRL_DECL_INST(f32_d0) RSRL_DECL_INST(f32_d1) RSRL_DECL_INST(f32_d2)
RSRL_DECL_INST(f64_d0) RSRL_DECL_INST(f64_d1) RSRL_DECL_INST(f64_d2)

struct F4Args;
cudaError_t launch_f4_env_f32(int domain, int order, bool ext, const StepArgs&, const F4Args&, int grid, int block, size_t smem, cudaStream_t);
cudaError_t launch_f4_env_f64(int domain, int order, bool ext, const StepArgs&, const F4Args&, int grid, int block, size_t smem, cudaStream_t);
cudaError_t launch_f4_dw_f32(int domain, int order, int64_t n, const double* from_states, const void* coef, const int32_t* actions, int n_seg, void* partials, cudaStream_t);
cudaError_t launch_f4_dw_f64(int domain, int order, int64_t n, const double* from_states, const void* coef, const int32_t* actions, int n_seg, void* partials, cudaStream_t);
cudaError_t launch_f4_eval_f32(int domain, int order, const EvalArgs&, cudaStream_t);
cudaError_t launch_f4_eval_f64(int domain, int order, const EvalArgs&, cudaStream_t);

struct TileArgs;
struct TileParams;
cudaError_t launch_tile_persist_f32(int domain, int aw, bool ext, const StepArgs&, int k, const TileArgs&, int grid, int block, size_t smem, cudaStream_t);
cudaError_t launch_tile_persist_f64(int domain, int aw, bool ext, const StepArgs&, int k, const TileArgs&, int grid, int block, size_t smem, cudaStream_t);
cudaError_t launch_tile_eval_f32(int domain, int aw, const EvalArgs&, const TileParams&, cudaStream_t);
cudaError_t launch_tile_eval_f64(int domain, int aw, const EvalArgs&, const TileParams&, cudaStream_t);

}  // namespace rsrl
