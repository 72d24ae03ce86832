// fourier4.cuh — large Fourier / Polynomial bases on the 4-D domains (BASELINE config 4: Acrobot,
// ExpectedSARSA, Fourier(7)+bias => F = 8^4 = 4096 features, W = 4096 x 3).
//
// F no longer fits registers, so the step is split into two kernels per batched step (+ the shared
// fixed-order partial reduce of kernels.cuh):
//   f4_env_kernel : one thread per env. Features are generated on the fly from per-dimension sin/cos
//                   tables in nested order (outer two dims: tables in shared memory, real loops; inner
//                   two dims: tables in registers, 64 features unrolled) — ~1.3 complex multiplies per
//                   feature instead of a cos — and contracted with W (shared memory, one 16-byte broadcast
//                   load per feature row) for Q(s_t) and Q(s').  Writes the scaled TD error per env.
//   f4_dw_kernel  : dW = Phi^T D as a tiled contraction: CTA = (c0 slice of 512 features) x (env segment);
//                   thread = (c1, c2) pair x env lane, 8 x A register accumulators; env tables are staged
//                   per 64-env chunk in shared memory so the loads of a warp are mostly broadcasts.
//                   Fixed env order per lane + fixed lane/segment order => bit-reproducible.
// This is the CUDA-core version; the tcgen05 formulation (Q = Phi W, dW = Phi^T D with Phi generated
// into UMMA-layout tiles) is the planned replacement.
#pragma once
#include "persistent.cuh"

namespace rsrl {

// per-dimension tables: e[d][j] = (cos, sin)(pi (j+1) x^_d), j = 0..P-1  (Polynomial: x^(j+1), 0)
template <typename R, class Dom, int P, int BASIS>
__device__ __forceinline__ void f4_tables(const double* st, R (&tc)[4][P], R (&ts)[4][P]) {
    GridTables<R, 4, P, BASIS> t;
    grid_prepare<R, Dom, P, BASIS>(st, t);
#pragma unroll
    for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int j = 0; j < P; ++j) { tc[d][j] = t.c[d][j]; ts[d][j] = BASIS == RSRL_FOURIER ? t.s[d][j] : (R)0; }
}

template <typename R>
__device__ __forceinline__ void cmul(R ar, R ai, R br, R bi, R& cr, R& ci) {
    cr = RealOps<R>::fma(ar, br, -(ai * bi));
    ci = RealOps<R>::fma(ar, bi, ai * br);
}

struct F4Args {
    double* from_states;  // [N][4]  s_t saved for the dW pass
    void* coef;           // R[N]    scaled TD error
    // f4tc.cuh only (else nullptr):
    float* tabs;          // float[56][N] per-dimension sin/cos tables of s_t (env fastest)
    float* q;             // float[N][4]  Q(s_t; W_t)
    float* aux;           // float4[N]    {Q(s_t)[a_t], reward, terminal, -}
    double* next_states;  // [N][4]       s' before the episode bookkeeping
};

// Outer-dimension tables live in shared memory as columns of the calling thread:
// sm[(d * P + j) * 2 * BLOCK + {0: cos, 1: sin} * BLOCK + tid], d in {0, 1}.
template <typename R, int P, int AW, class Fn>
__device__ __forceinline__ void f4_for_each(const R* sm_outer, int BLOCK, int tid, const R (&tc)[4][P], const R (&ts)[4][P], Fn fn) {
    constexpr int N1 = P + 1;
    for (int i0 = 0; i0 < N1; ++i0) {
        const int c0 = P - i0;
        const R a_r = c0 == 0 ? (R)1 : sm_outer[((0 * P + c0 - 1) * 2 + 0) * BLOCK + tid];
        const R a_i = c0 == 0 ? (R)0 : sm_outer[((0 * P + c0 - 1) * 2 + 1) * BLOCK + tid];
        for (int i1 = 0; i1 < N1; ++i1) {
            const int c1 = P - i1;
            R z01r = a_r, z01i = a_i;
            if (c1 != 0) cmul<R>(a_r, a_i, sm_outer[((1 * P + c1 - 1) * 2 + 0) * BLOCK + tid], sm_outer[((1 * P + c1 - 1) * 2 + 1) * BLOCK + tid], z01r, z01i);
            const int kbase = (i0 * N1 + i1) * N1 * N1;
#pragma unroll
            for (int i2 = 0; i2 < N1; ++i2) {
                const int c2 = P - i2;
                R zr = z01r, zi = z01i;
                if (c2 != 0) cmul<R>(z01r, z01i, tc[2][c2 - 1], ts[2][c2 - 1], zr, zi);
#pragma unroll
                for (int i3 = 0; i3 < N1; ++i3) {
                    const int c3 = P - i3;
                    const R phi = c3 == 0 ? zr : RealOps<R>::fma(zr, tc[3][c3 - 1], -(zi * ts[3][c3 - 1]));
                    fn(kbase + i2 * N1 + i3, phi);
                }
            }
        }
    }
}

template <typename R, int DOM, int BASIS, int P, int AW, bool EXT>
__global__ void __launch_bounds__(128) f4_env_kernel(const StepArgs a, const F4Args fa) {
    using Dom = Domain<DOM>;
    using O = RealOps<R>;
    constexpr int D = 4, N1 = P + 1, F = N1 * N1 * N1 * N1, WS = 4;
    static_assert(Dom::D == 4, "4-D domains only");
    typedef typename Vec16<R>::type vec_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* Wsm = reinterpret_cast<R*>(smem_raw);        // [F][WS]
    R* outer = Wsm + (size_t)F * WS;                // [2][P][2][BLOCK] outer-dimension tables of the current state
    const int tid = threadIdx.x, BLOCK = blockDim.x;
    const int64_t i = (int64_t)blockIdx.x * BLOCK + tid;
    const bool active = i < a.n;
    for (int j = tid; j < F * AW; j += BLOCK) Wsm[(j / AW) * WS + j % AW] = static_cast<const R*>(a.W)[j];
    __syncthreads();
    if (!active) return;
    const uint64_t g = (uint64_t)(a.env_offset + i);

    struct Tab { R tc[4][P], ts[4][P]; };
    auto prep = [&](const double* st, Tab& tb) { f4_tables<R, Dom, P, BASIS>(st, tb.tc, tb.ts); };
    auto evalQ = [&](const Tab& tb, R* q) {
        // publish this thread's outer tables (dims 0, 1) so that the c0 / c1 loops can index them dynamically
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
            for (int j = 0; j < P; ++j) {
                outer[((d * P + j) * 2 + 0) * BLOCK + tid] = tb.tc[d][j];
                outer[((d * P + j) * 2 + 1) * BLOCK + tid] = tb.ts[d][j];
            }
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = (R)0;
        f4_for_each<R, P, AW>(outer, BLOCK, tid, tb.tc, tb.ts, [&](int k, R phi) {
            const vec_t* p = reinterpret_cast<const vec_t*>(Wsm + (size_t)k * WS);
            const vec_t v0 = p[0];
            if (Vec16<R>::N == 4) {
#pragma unroll
                for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, vget(v0, c), q[c]);
            } else {
                const vec_t v1 = AW > 2 ? p[1] : v0;
#pragma unroll
                for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, c < 2 ? vget(v0, c) : vget(v1, c - 2), q[c]);
            }
        });
    };

    double s[D];
#pragma unroll
    for (int d = 0; d < D; ++d) s[d] = EXT ? a.ext_from[i * D + d] : a.states[i * D + d];
#pragma unroll
    for (int d = 0; d < D; ++d) fa.from_states[i * D + d] = s[d];
    Tab tab_s, tab_n;
    CoreOut<R> o;
    env_core<R, DOM, AW, EXT>(a, a.t, g, s, prep, evalQ, evalQ, tab_s, tab_n, false, o, EXT ? a.ext_actions[i] : 0,
                              EXT ? a.ext_rewards[i] : 0.0, EXT ? a.ext_term[i] != 0 : false, EXT ? a.ext_to + i * D : nullptr);
    if (a.td) static_cast<R*>(a.td)[i] = o.residual;
    if (o.nonfinite) atomicExch(&a.counters->nonfinite, 1);
    static_cast<R*>(fa.coef)[i] = o.coef;
    a.actions[i] = o.act;  // (EXT: the caller's action, consumed by the dW pass)
    if (!EXT) {
        a.ep_steps[i] = env_bookkeeping<Dom>(a, a.t, i, g, s, a.ep_steps[i], o.terminated);
#pragma unroll
        for (int d = 0; d < D; ++d) a.states[i * D + d] = s[d];
    }
}

// dW partial for one (c0 slice, env segment): partials[seg][k * AW + a], k in the slice.
// blockDim = 256: thread = ((c1, c2) pair p = tid / 4 in 0..63, env lane l = tid % 4).
template <typename R, int DOM, int BASIS, int P, int AW>
__global__ void __launch_bounds__(256) f4_dw_kernel(int64_t n, const double* __restrict__ from_states, const R* __restrict__ coef,
                                                    const int32_t* __restrict__ actions, int n_seg, R* __restrict__ partials) {
    using Dom = Domain<DOM>;
    using O = RealOps<R>;
    constexpr int N1 = P + 1, F = N1 * N1 * N1 * N1, CH = 64, NE = 2 * 4 * P;  // table entries per env
    constexpr bool TDPRED = AW == 1;
    static_assert(N1 * N1 * 4 <= 256, "thread mapping: (P+1)^2 (c1,c2) pairs x 4 env lanes must fit one CTA");
    __shared__ R tabs[CH][NE + 1];   // [env in chunk][(d * P + j) * 2 + {cos, sin}] (+1: conflict-free fill)
    __shared__ R dsm[CH][4];         // scaled TD error in the env's action column
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x;       // c0 slice
    const int seg = blockIdx.y;
    const int c0 = P - i0;
    const int lane = tid & 3;
    const bool valid = (tid >> 2) < N1 * N1;  // blockDim is (P+1)^2 * 4 rounded up to whole warps
    const int p = valid ? tid >> 2 : 0;
    const int i1 = p / N1, i2 = p % N1, c1 = P - i1, c2 = P - i2;
    const int64_t per_seg = (n + n_seg - 1) / n_seg;
    const int64_t e0 = (int64_t)seg * per_seg, e1 = e0 + per_seg < n ? e0 + per_seg : n;

    R acc[N1][AW];
#pragma unroll
    for (int j = 0; j < N1; ++j)
#pragma unroll
        for (int c = 0; c < AW; ++c) acc[j][c] = (R)0;

    for (int64_t cbase = e0; cbase < e1; cbase += CH) {
        __syncthreads();
        if (tid < CH) {
            const int64_t i = cbase + tid;
            R tc[4][P], ts[4][P];
            R dv[4] = {(R)0, (R)0, (R)0, (R)0};
            if (i < e1) {
                double st[4];
#pragma unroll
                for (int d = 0; d < 4; ++d) st[d] = from_states[i * 4 + d];
                f4_tables<R, Dom, P, BASIS>(st, tc, ts);
                const R cf = coef[i];
                const int act = TDPRED ? 0 : actions[i];
#pragma unroll
                for (int c = 0; c < AW; ++c) dv[c] = c == act ? cf : (R)0;
            } else {
#pragma unroll
                for (int d = 0; d < 4; ++d)
#pragma unroll
                    for (int j = 0; j < P; ++j) { tc[d][j] = (R)0; ts[d][j] = (R)0; }
            }
#pragma unroll
            for (int d = 0; d < 4; ++d)
#pragma unroll
                for (int j = 0; j < P; ++j) { tabs[tid][(d * P + j) * 2] = tc[d][j]; tabs[tid][(d * P + j) * 2 + 1] = ts[d][j]; }
#pragma unroll
            for (int c = 0; c < 4; ++c) dsm[tid][c] = dv[c];
        }
        __syncthreads();
        for (int e = lane; valid && e < CH; e += 4) {  // env order per lane: ascending
            const R* tb = tabs[e];
            R zr = c0 == 0 ? (R)1 : tb[(0 * P + c0 - 1) * 2], zi = c0 == 0 ? (R)0 : tb[(0 * P + c0 - 1) * 2 + 1];
            if (c1 != 0) { R r2, i2v; cmul<R>(zr, zi, tb[(1 * P + c1 - 1) * 2], tb[(1 * P + c1 - 1) * 2 + 1], r2, i2v); zr = r2; zi = i2v; }
            if (c2 != 0) { R r2, i2v; cmul<R>(zr, zi, tb[(2 * P + c2 - 1) * 2], tb[(2 * P + c2 - 1) * 2 + 1], r2, i2v); zr = r2; zi = i2v; }
            R d[AW];
#pragma unroll
            for (int c = 0; c < AW; ++c) d[c] = dsm[e][c];
#pragma unroll
            for (int i3 = 0; i3 < N1; ++i3) {
                const int c3 = P - i3;
                const R phi = c3 == 0 ? zr : O::fma(zr, tb[(3 * P + c3 - 1) * 2], -(zi * tb[(3 * P + c3 - 1) * 2 + 1]));
#pragma unroll
                for (int c = 0; c < AW; ++c) acc[i3][c] = O::fma(phi, d[c], acc[i3][c]);
            }
        }
    }
    // lanes 0..3 of a pair are adjacent lanes of one warp: fixed-order butterfly
#pragma unroll
    for (int j = 0; j < N1; ++j)
#pragma unroll
        for (int c = 0; c < AW; ++c) {
            R v = acc[j][c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            acc[j][c] = v;
        }
    if (lane == 0 && valid) {
        R* out = partials + (size_t)seg * F * AW;
#pragma unroll
        for (int i3 = 0; i3 < N1; ++i3) {
            const int k = ((i0 * N1 + i1) * N1 + i2) * N1 + i3;
#pragma unroll
            for (int c = 0; c < AW; ++c) out[(size_t)k * AW + c] = acc[i3][c];
        }
    }
}

// component entry points for the large bases: mode 0 features (N x F), 1 Q, 2 sample, 3 find_max
template <typename R, int DOM, int BASIS, int P, int AW>
__global__ void __launch_bounds__(128) f4_eval_kernel(int mode, int64_t n, const double* __restrict__ states, const R* __restrict__ W,
                                                      double* __restrict__ out, int32_t* __restrict__ act_out, PolicyParams pol,
                                                      uint64_t draw, int64_t env_offset, Counters* counters) {
    using Dom = Domain<DOM>;
    using O = RealOps<R>;
    constexpr int N1 = P + 1, F = N1 * N1 * N1 * N1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* outer = reinterpret_cast<R*>(smem_raw);
    const int tid = threadIdx.x, BLOCK = blockDim.x;
    const int64_t i = (int64_t)blockIdx.x * BLOCK + tid;
    if (i >= n) return;
    double s[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) s[d] = states[i * 4 + d];
    R tc[4][P], ts[4][P];
    f4_tables<R, Dom, P, BASIS>(s, tc, ts);
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
        for (int j = 0; j < P; ++j) {
            outer[((d * P + j) * 2 + 0) * BLOCK + tid] = tc[d][j];
            outer[((d * P + j) * 2 + 1) * BLOCK + tid] = ts[d][j];
        }
    R q[AW];
#pragma unroll
    for (int c = 0; c < AW; ++c) q[c] = (R)0;
    f4_for_each<R, P, AW>(outer, BLOCK, tid, tc, ts, [&](int k, R phi) {
        if (mode == 0) { out[i * F + k] = (double)phi; return; }
#pragma unroll
        for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, W[(size_t)k * AW + c], q[c]);
    });
    if (mode == 1) {
#pragma unroll
        for (int c = 0; c < AW; ++c) out[i * AW + c] = (double)q[c];
    } else if (mode == 2) {
        bool nf = false;
        act_out[i] = policy_sample<R, AW>(pol, q, (uint64_t)(env_offset + i), draw, STREAM_BEHAVIOUR, nf);
        if (nf) atomicExch(&counters->nonfinite, 1);
    } else if (mode == 3) {
        R mx;
        (void)mx;
        act_out[i] = policy_mode<R, AW>(pol.policy, (R)pol.tau, q);
    }
}

}  // namespace rsrl
