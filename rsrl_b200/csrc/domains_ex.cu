// domains_ex.cu — the remaining ODE / continuous-action domains of rsrl_domains as batched component kernels:
//   ContinuousMountainCar  rsrl_domains/src/mountain_car/continuous.rs:8-85   (action in [-1, 1], FORCE_CAR = 0.0015)
//   HIVTreatment           rsrl_domains/src/hiv.rs:6-153                       (6 states, 4 actions, 1000 RK4 sub-steps of DT / 1000)
// Stateless entry points in the layout of rsrl_domain_step: the caller owns the (raw) states; one thread per env, f64,
// unfused operations in the reference's association order (oracle: rsrl_oracle.c orc_domain_ex_*).
#include <cstring>
#include <string>

#include "device.cuh"

namespace rsrl {
namespace {

// continuous.rs:41-48: a = action_space.map_onto(a) (bounded Interval: clip to [-1, 1]); v, x updates as the discrete car
__device__ __forceinline__ void cmc_step(double* s, double action, double& reward, bool& terminal) {
    const double a = dclip(-1.0, action, 1.0);
    const double dv = dadd(dmul(0.0015, a), dmul(-0.0025, cos64(dmul(3.0, s[0]))));
    s[1] = dclip(-0.07, dadd(s[1], dv), 0.07);
    s[0] = dclip(-1.2, dadd(s[0], s[1]), 0.6);
    terminal = s[0] >= 0.6;
    reward = terminal ? 0.0 : -1.0;
}

// hiv.rs:72-103 (state order T1, T1S, T2, T2S, V, E)
__device__ __forceinline__ void hiv_grad(double e0, double e1, const double* b, double* out) {
    constexpr double LAMBDA1 = 1e4, LAMBDA2 = 31.98, D1 = 0.01, D2 = 0.01, F = 0.34, K1 = 8e-7, K2 = 1e-4, DELTA = 0.7, M1 = 1e-5, M2 = 1e-5,
                     NT = 100.0, C = 13.0, RHO1 = 1.0, RHO2 = 1.0, LAMBDA_E = 1.0, BE = 0.3, KB = 100.0, DE = 0.25, KD = 500.0, DELTA_E = 0.1;
    const double t1 = b[0], t1s = b[1], t2 = b[2], t2s = b[3], v = b[4], e = b[5];
    const double tmp1 = dmul(dmul(dmul(dsub(1.0, e0), K1), v), t1);
    const double tmp2 = dmul(dmul(dmul(dsub(1.0, dmul(F, e0)), K2), v), t2);
    const double sum_ts = dadd(t1s, t2s);
    out[0] = dsub(dsub(LAMBDA1, dmul(D1, t1)), tmp1);
    out[1] = dsub(dsub(tmp1, dmul(DELTA, t1s)), dmul(dmul(M1, e), t1s));
    out[2] = dsub(dsub(LAMBDA2, dmul(D2, t2)), tmp2);
    out[3] = dsub(dsub(tmp2, dmul(DELTA, t2s)), dmul(dmul(M2, e), t2s));
    out[4] = dsub(dsub(dmul(dmul(dmul(dsub(1.0, e1), NT), DELTA), sum_ts), dmul(C, v)),
                  dmul(dadd(dmul(dmul(dmul(dsub(1.0, e0), RHO1), K1), t1), dmul(dmul(dmul(dsub(1.0, dmul(F, e0)), RHO2), K2), t2)), v));
    out[5] = dsub(dsub(dadd(LAMBDA_E, dmul(ddiv(dmul(BE, sum_ts), dadd(sum_ts, KB)), e)), dmul(ddiv(dmul(DE, sum_ts), dadd(sum_ts, KD)), e)),
                  dmul(DELTA_E, e));
}

// hiv.rs:131-135 emit(): clip(-5, log10(v), 8) per component
__device__ __forceinline__ void hiv_emit(const double* s, double* obs) {
#pragma unroll
    for (int d = 0; d < 6; ++d) obs[d] = dclip(-5.0, log10(s[d]), 8.0);
}

__device__ __forceinline__ void hiv_step(double* s, int action, double* obs, double& reward) {
    const double e0 = (action & 1) ? 0.7 : 0.0, e1 = (action & 2) ? 0.3 : 0.0;   // ALL_ACTIONS :35
    constexpr double DT_STEP = 5.0 / 1000.0;
    double y[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) y[d] = s[d];
#pragma unroll 1
    for (int it = 0; it < 1000; ++it) {   // ode.rs:1-43, same association as the 4-D version in device.cuh
        double k1[6], k2[6], k3[6], k4[6], tmp[6];
        hiv_grad(e0, e1, y, k1);
#pragma unroll
        for (int d = 0; d < 6; ++d) { k1[d] = dmul(k1[d], DT_STEP); tmp[d] = dadd(y[d], dmul(k1[d], 0.5)); }
        hiv_grad(e0, e1, tmp, k2);
#pragma unroll
        for (int d = 0; d < 6; ++d) { k2[d] = dmul(k2[d], DT_STEP); tmp[d] = dadd(y[d], dmul(k2[d], 0.5)); }
        hiv_grad(e0, e1, tmp, k3);
#pragma unroll
        for (int d = 0; d < 6; ++d) { k3[d] = dmul(k3[d], DT_STEP); tmp[d] = dadd(y[d], k3[d]); }
        hiv_grad(e0, e1, tmp, k4);
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            k4[d] = dmul(k4[d], DT_STEP);
            y[d] = dadd(y[d], ddiv(dadd(dadd(dadd(k1[d], dmul(2.0, k2[d])), dmul(2.0, k3[d])), k4[d]), 6.0));
        }
    }
#pragma unroll
    for (int d = 0; d < 6; ++d) s[d] = y[d];
    hiv_emit(s, obs);
    // hiv.rs:141-148: r = (1e3 E - 0.1 V - 2e4 eps0^2 - 2e3 eps1^2) / 1e5 on the OBSERVATION (log10 values), powi(2) = x * x
    const double r = dsub(dsub(dsub(dmul(1e3, obs[5]), dmul(0.1, obs[4])), dmul(2e4, dmul(e0, e0))), dmul(2e3, dmul(e1, e1)));
    reward = ddiv(r, 1e5);
}

}  // namespace

__global__ void domain_ex_kernel(int domain, int mode /* 0 step, 1 emit */, int64_t n, double* states, const int32_t* act_i, const double* act_f,
                                 double* obs, double* rewards, uint8_t* terminal) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (domain == RSRL_CONTINUOUS_MOUNTAIN_CAR) {
        double s[2] = {states[i * 2], states[i * 2 + 1]};
        if (mode == 0) {
            double r; bool t;
            cmc_step(s, act_f[i], r, t);
            states[i * 2] = s[0]; states[i * 2 + 1] = s[1];
            rewards[i] = r; terminal[i] = t;
        } else {
            terminal[i] = s[0] >= 0.6;
        }
        if (obs) { obs[i * 2] = s[0]; obs[i * 2 + 1] = s[1]; }
    } else {
        double s[6], o[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) s[d] = states[i * 6 + d];
        if (mode == 0) {
            double r;
            hiv_step(s, act_i[i], o, r);
#pragma unroll
            for (int d = 0; d < 6; ++d) states[i * 6 + d] = s[d];
            rewards[i] = r;
        } else {
            hiv_emit(s, o);
        }
        terminal[i] = 0;   // hiv.rs:131-135: always Observation::Full
        if (obs) {
#pragma unroll
            for (int d = 0; d < 6; ++d) obs[i * 6 + d] = o[d];
        }
    }
}

}  // namespace rsrl

static thread_local std::string g_err_ex;

extern "C" {

const char* rsrl_domain_ex_last_error(void) { return g_err_ex.c_str(); }

int rsrl_domain_ex_info(int32_t domain, int32_t* dim, int32_t* n_actions, double* lo, double* hi, double* start) {
    if (domain == RSRL_CONTINUOUS_MOUNTAIN_CAR) {
        if (dim) *dim = 2;
        if (n_actions) *n_actions = 0;   // continuous: Interval::bounded(-1, 1)
        const double l[2] = {-1.2, -0.07}, h[2] = {0.6, 0.07}, s[2] = {-0.5, 0.0};
        for (int d = 0; d < 2; ++d) { if (lo) lo[d] = l[d]; if (hi) hi[d] = h[d]; if (start) start[d] = s[d]; }
        return RSRL_OK;
    }
    if (domain == RSRL_HIV) {
        if (dim) *dim = 6;
        if (n_actions) *n_actions = 4;
        const double s[6] = {163573.0, 11945.0, 5.0, 46.0, 63919.0, 24.0};   // hiv.rs:105-109 (raw state; observations are log10)
        for (int d = 0; d < 6; ++d) { if (lo) lo[d] = -5.0; if (hi) hi[d] = 8.0; if (start) start[d] = s[d]; }
        return RSRL_OK;
    }
    g_err_ex = "unknown extended domain";
    return RSRL_EINVAL;
}

static int domain_ex_call(int32_t domain, int mode, int64_t n, double* states, const int32_t* act_i, const double* act_f, double* obs,
                          double* rewards, uint8_t* terminal) {
    if ((domain != RSRL_CONTINUOUS_MOUNTAIN_CAR && domain != RSRL_HIV) || n <= 0 || !states || !terminal) { g_err_ex = "bad argument"; return RSRL_EINVAL; }
    if (mode == 0 && (!rewards || (domain == RSRL_HIV ? !act_i : !act_f))) { g_err_ex = "bad argument"; return RSRL_EINVAL; }
    if (mode == 0 && domain == RSRL_HIV) for (int64_t i = 0; i < n; ++i) if (act_i[i] < 0 || act_i[i] > 3) { g_err_ex = "action out of range"; return RSRL_EINVAL; }
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0) { cudaGetLastError(); g_err_ex = "no CUDA device visible: rsrl_b200 has no CPU fallback"; return RSRL_ENODEVICE; }
    const int D = domain == RSRL_HIV ? 6 : 2;
    double *ds = nullptr, *dobs = nullptr, *dr = nullptr, *daf = nullptr;
    int32_t* dai = nullptr;
    uint8_t* dt = nullptr;
    cudaError_t ce = cudaSuccess;
    auto chk = [&](cudaError_t e) { if (ce == cudaSuccess) ce = e; };
    chk(cudaMalloc(&ds, n * D * sizeof(double))); chk(cudaMalloc(&dobs, n * D * sizeof(double))); chk(cudaMalloc(&dr, n * sizeof(double)));
    chk(cudaMalloc(&dt, n)); chk(cudaMalloc(&dai, n * sizeof(int32_t))); chk(cudaMalloc(&daf, n * sizeof(double)));
    if (ce == cudaSuccess) {
        chk(cudaMemcpy(ds, states, n * D * sizeof(double), cudaMemcpyHostToDevice));
        if (mode == 0 && act_i) chk(cudaMemcpy(dai, act_i, n * sizeof(int32_t), cudaMemcpyHostToDevice));
        if (mode == 0 && act_f) chk(cudaMemcpy(daf, act_f, n * sizeof(double), cudaMemcpyHostToDevice));
        const int threads = 128, blocks = (int)((n + threads - 1) / threads);
        rsrl::domain_ex_kernel<<<blocks, threads>>>(domain, mode, n, ds, dai, daf, dobs, dr, dt);
        chk(cudaGetLastError());
        if (mode == 0) { chk(cudaMemcpy(states, ds, n * D * sizeof(double), cudaMemcpyDeviceToHost)); chk(cudaMemcpy(rewards, dr, n * sizeof(double), cudaMemcpyDeviceToHost)); }
        if (obs) chk(cudaMemcpy(obs, dobs, n * D * sizeof(double), cudaMemcpyDeviceToHost));
        chk(cudaMemcpy(terminal, dt, n, cudaMemcpyDeviceToHost));
    }
    cudaFree(ds); cudaFree(dobs); cudaFree(dr); cudaFree(dt); cudaFree(dai); cudaFree(daf);
    if (ce != cudaSuccess) { g_err_ex = cudaGetErrorString(ce); return ce == cudaErrorMemoryAllocation ? RSRL_ENOMEM : RSRL_ECUDA; }
    return RSRL_OK;
}

int rsrl_domain_ex_step(int32_t domain, int64_t n, double* states_inout, const int32_t* actions, const double* actions_continuous,
                        double* obs_out, double* rewards_out, uint8_t* terminal_out) {
    return domain_ex_call(domain, 0, n, states_inout, actions, actions_continuous, obs_out, rewards_out, terminal_out);
}

int rsrl_domain_ex_emit(int32_t domain, int64_t n, const double* states, double* obs_out, uint8_t* terminal_out) {
    return domain_ex_call(domain, 1, n, const_cast<double*>(states), nullptr, nullptr, obs_out, nullptr, terminal_out);
}

}  // extern "C"
