// device.cuh — per-env device arithmetic of the rsrl hot path (sm_100a).
//
// Everything here is written from the reference's behaviour (file:line cited per
// function, paths relative to the reference checkout), not from its code shape:
// the reference allocates a Vec per observation / feature vector and re-projects
// the state four times per step; here one env lives in the registers of one
// thread, the Fourier features are generated from per-dimension sin/cos tables
// (Kronecker structure) and never touch memory.
#pragma once
#include "hostdev.h"
#include "../../include/rsrl_b200.h"

namespace rsrl {

// ---------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------
// f64 physics: explicit round-to-nearest ops (hostdev.h: dmul / dadd / dsub / ddiv) so that nvcc does not contract a*b+c
// into DFMA — the reference (Rust, no FMA contraction) rounds after every operation.
// clip!(lb, x, ub) = lb.max(ub.min(x))  (rsrl_domains/src/macros.rs:20-24); fmin/fmax drop NaN like Rust
// as two compare-selects: the same values as fmax(lb, fmin(ub, x)) for lb < ub, neither zero (NaN -> ub), half the instructions
__host__ __device__ __forceinline__ double dclip(double lb, double x, double ub) {
    const double t = x < ub ? x : ub;
    return t > lb ? t : lb;
}
// wrap!(lb, x, ub)  (macros.rs:3-18)
__host__ __device__ __forceinline__ double dwrap(double lb, double x, double ub) {
    double nx = x;
    const double diff = dsub(ub, lb);
    while (nx > ub) nx = dsub(nx, diff);
    while (nx < lb) nx = dadd(nx, diff);
    return nx;
}

#define RSRL_PI 3.14159265358979323846264338327950288

// ---------------------------------------------------------------------------
// rsrl math: the elementary functions of the hot path, written out so that the GPU and the host build
// (oracle/oracle32.cpp) round identically.  Accuracy (tests/test_arith32.py, against mpmath): sin64 / cos64
// < 1 ulp for |x| < 2^20 (3-term Cody-Waite reduction with FMA + the classic degree-13/14 minimax kernels
// on [-pi/4, pi/4]); sincospi32 < 1 ulp (reduction x - rint(2x)/2 is exact); exp32 < 1 ulp.
// Outside those ranges (never reached by the bounded domains) the platform libm is used.
// ---------------------------------------------------------------------------
// The f64 constants of the reduction and of the two kernels.  On the device they come from constant memory: as literals the compiler
// materialises each one with two UMOV per use (52 extra issue slots per MountainCar step, 190 per CartPole RK4 step); as a constant-bank
// operand they are free.  Same values on both sides.
#if defined(__CUDACC__)
static __constant__ double kMath64[16] = {
    0.63661977236758138, -1.5707963267948966, -6.123233995736766e-17, -1.4973849048591698e-33,
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
    8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
    -1.38888888888741095749e-03, 4.16666666666666019037e-02};
#endif
#ifdef __CUDA_ARCH__
#define RSRL_M64(i, literal) kMath64[i]
#else
#define RSRL_M64(i, literal) (literal)
#endif

// x = n*pi/2 + (r + lo), |r| <= pi/4 + eps, |lo| < ulp(r); returns n mod 4
__host__ __device__ __forceinline__ int rem_pio2_64(double x, double& r, double& lo) {
    const double n = rint(dmul(x, RSRL_M64(0, 0.63661977236758138)));          // 2/pi
    const double r1 = dfma(n, RSRL_M64(1, -1.5707963267948966), x);            // exact: x and n*P1 cancel to <= 53 bits
    r = dfma(n, RSRL_M64(2, -6.123233995736766e-17), r1);
    lo = dfma(n, RSRL_M64(2, -6.123233995736766e-17), dsub(r1, r));            // the rounding error of r ...
    lo = dfma(n, RSRL_M64(3, -1.4973849048591698e-33), lo);                    // ... and the third part of pi/2
    return (int)n & 3;
}
__host__ __device__ __forceinline__ double sin_kernel64(double r, double lo) {  // sin(r + lo)
    const double z = dmul(r, r);
    double p = RSRL_M64(4, 1.58969099521155010221e-10);
    p = dfma(p, z, RSRL_M64(5, -2.50507602534068634195e-08));
    p = dfma(p, z, RSRL_M64(6, 2.75573137070700676789e-06));
    p = dfma(p, z, RSRL_M64(7, -1.98412698298579493134e-04));
    p = dfma(p, z, RSRL_M64(8, 8.33333333332248946124e-03));
    p = dfma(p, z, RSRL_M64(9, -1.66666666666666324348e-01));
    const double corr = dfma(dmul(r, z), p, dmul(lo, dfma(z, -0.5, 1.0)));  // r^3 p(z) + lo cos(r)
    return dadd(r, corr);
}
__host__ __device__ __forceinline__ double cos_kernel64(double r, double lo) {  // cos(r + lo)
    const double z = dmul(r, r);
    double p = RSRL_M64(10, -1.13596475577881948265e-11);
    p = dfma(p, z, RSRL_M64(11, 2.08757232129817482790e-09));
    p = dfma(p, z, RSRL_M64(12, -2.75573143513906633035e-07));
    p = dfma(p, z, RSRL_M64(13, 2.48015872894767294178e-05));
    p = dfma(p, z, RSRL_M64(14, -1.38888888888741095749e-03));
    p = dfma(p, z, RSRL_M64(15, 4.16666666666666019037e-02));
    const double h = dfma(z, -0.5, 1.0);                           // 1 - z/2 and its rounding error e
    const double e = dfma(z, -0.5, dsub(1.0, h));
    return dadd(h, dadd(e, dfma(dmul(z, z), p, -dmul(r, lo))));    // + z^2 p(z) - lo sin(r)
}
__host__ __device__ __forceinline__ void sincos64(double x, double* s, double* c) {
    if (!(fabs(x) < 1048576.0)) { *s = ::sin(x); *c = ::cos(x); return; }  // out of the pinned range (incl. NaN / Inf)
    double r, lo;
    const int q = rem_pio2_64(x, r, lo);
    const double sk = sin_kernel64(r, lo), ck = cos_kernel64(r, lo);
    const double sv = (q & 1) ? ck : sk, cv = (q & 1) ? sk : ck;
    *s = (q & 2) ? -sv : sv;
    *c = ((q + 1) & 2) ? -cv : cv;
}
__host__ __device__ __forceinline__ double cos64(double x) {
    if (!(fabs(x) < 1048576.0)) return ::cos(x);
    double r, lo;
    const int q = rem_pio2_64(x, r, lo);
    const double v = (q & 1) ? sin_kernel64(r, lo) : cos_kernel64(r, lo);
    return ((q + 1) & 2) ? -v : v;
}

// sin(pi x), cos(pi x) in fp32.  t = rint(2x); r = x - t/2 is exact; |r| <= 1/4.
__host__ __device__ __forceinline__ void sincospi32(float x, float* s, float* c) {
    if (!(fabsf(x) < 4194304.0f)) x = fmul(x, 0.0f);  // huge: an even integer for every such float; NaN / Inf -> NaN
    const float t = rintf(fadd(x, x));
    const float r = ffma(t, -0.5f, x);
    const float z = fmul(r, r);
    float ps = -0.5890144109725952f;
    ps = ffma(ps, z, 2.5497612953186035f);
    ps = ffma(ps, z, -5.167707920074463f);
    // pi r = r * pi_hi + r * pi_lo, then the cubic and higher terms
    const float sk = ffma(r, 3.1415927410125732f, ffma(r, -8.742277657347586e-08f, fmul(fmul(r, z), ps)));
    float pc = 0.23136425018310547f;
    pc = ffma(pc, z, -1.3350505828857422f);
    pc = ffma(pc, z, 4.0587077140808105f);
    pc = ffma(pc, z, -4.934802055358887f);
    const float ck = ffma(pc, z, 1.0f);
    const int q = (int)t & 3;
    const float sv = (q & 1) ? ck : sk, cv = (q & 1) ? sk : ck;
    *s = (q & 2) ? -sv : sv;
    *c = ((q + 1) & 2) ? -cv : cv;
}

// e^x in fp32 (Softmax policy)
__host__ __device__ __forceinline__ float exp32(float x) {
    if (!(x <= 88.72283935546875f)) return x > 0.0f ? bits_to_float(0x7f800000u) : x;  // +Inf; NaN stays NaN
    if (x < -103.97208404541016f) return 0.0f;
    const float n = rintf(fmul(x, 1.4426950216293335f));
    float r = ffma(n, -0.693145751953125f, x);        // ln2 high part (trailing zero bits: n * hi is exact)
    r = ffma(n, -1.4286068203094172e-06f, r);
    float p = 0.0013814615085721016f;
    p = ffma(p, r, 0.008368710055947304f);
    p = ffma(p, r, 0.04166838899254799f);
    p = ffma(p, r, 0.1666652113199234f);
    p = ffma(p, r, 0.4999999403953552f);
    const float e = fadd(ffma(fmul(r, r), p, r), 1.0f);
    const int ni = (int)n, n1 = ni >> 1, n2 = ni - n1;  // two exact power-of-two scalings: denormal results round once
    return fmul(fmul(e, bits_to_float((uint32_t)(n1 + 127) << 23)), bits_to_float((uint32_t)(n2 + 127) << 23));
}

template <typename R> struct RealOps;
template <> struct RealOps<float> {
    __host__ __device__ __forceinline__ static void sincospi(float x, float* s, float* c) { sincospi32(x, s, c); }
    __host__ __device__ __forceinline__ static float fma(float a, float b, float c) { return ffma(a, b, c); }
    // Q accumulation q + phi * w: fused in fp32
    __host__ __device__ __forceinline__ static float mac(float a, float b, float c) { return ffma(a, b, c); }
    // w + c*x with the product rounded first (the reference's SGD `w = w + (lr*err) * phi` is not fused)
    __host__ __device__ __forceinline__ static float mul_add_unfused(float c, float x, float w) { return fadd(w, fmul(c, x)); }
    __host__ __device__ __forceinline__ static float abs(float a) { return fabsf(a); }
    __host__ __device__ __forceinline__ static float lowest() { return -3.402823466e+38f; }  // magnitude only matters vs 1e-7
    __host__ __device__ __forceinline__ static float clamp1(float x) { return fmaxf(-1.0f, fminf(1.0f, x)); }
    __host__ __device__ __forceinline__ static float exp(float x) { return exp32(x); }
    __host__ __device__ __forceinline__ static float max_nan(float a, float b) { return fmaxf(a, b); }  // f64::max: drops NaN
    __host__ __device__ __forceinline__ static float min_max(float x) { return fminf(x, 3.402823466e+38f); }
};
template <> struct RealOps<double> {
    __host__ __device__ __forceinline__ static void sincospi(double x, double* s, double* c) {
#if RSRL_HOST_BUILD
        *s = ::sin(RSRL_PI * x); *c = ::cos(RSRL_PI * x);  // the host build only instantiates fp32 (oracle32)
#else
        ::sincospi(x, s, c);
#endif
    }
    __host__ __device__ __forceinline__ static double fma(double a, double b, double c) { return dfma(a, b, c); }
    // dtype f64 is the reference's arithmetic: q + phi * w rounds the product first (ndarray dot / the oracle do not fuse)
    __host__ __device__ __forceinline__ static double mac(double a, double b, double c) { return dadd(c, dmul(a, b)); }
    __host__ __device__ __forceinline__ static double mul_add_unfused(double c, double x, double w) { return dadd(w, dmul(c, x)); }
    __host__ __device__ __forceinline__ static double abs(double a) { return fabs(a); }
    __host__ __device__ __forceinline__ static double lowest() { return -1.7976931348623157e+308; }  // f64::MIN
    __host__ __device__ __forceinline__ static double clamp1(double x) { return fmax(-1.0, fmin(1.0, x)); }
    __host__ __device__ __forceinline__ static double exp(double x) { return ::exp(x); }
    __host__ __device__ __forceinline__ static double max_nan(double a, double b) { return fmax(a, b); }
    __host__ __device__ __forceinline__ static double min_max(double x) { return fmin(x, 1.7976931348623157e+308); }
};

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. SC'11).  Project-defined stream layout
// (SURVEY 7.2): key = seed, counter = (global env id, draw lo, draw hi, stream).
// ---------------------------------------------------------------------------
enum : uint32_t { STREAM_INIT = 0, STREAM_BEHAVIOUR = 1, STREAM_TARGET = 2 };

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = umulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = umulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__host__ __device__ __forceinline__ uint4 draw4(uint64_t seed, uint64_t env, uint64_t draw, uint32_t stream) {
    return philox4x32_10(make_uint4((uint32_t)env, (uint32_t)draw, (uint32_t)(draw >> 32), stream),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__host__ __device__ __forceinline__ uint32_t word(const uint4& r, int d) { return d == 0 ? r.x : d == 1 ? r.y : d == 2 ? r.z : r.w; }

// ---------------------------------------------------------------------------
// Domains: one env step in registers.  `Domain::step` of rsrl_domains/src/lib.rs:434.
// ---------------------------------------------------------------------------
template <int DOM> struct Domain;

// rsrl_domains/src/mountain_car/discrete.rs:8-22,56-102
template <> struct Domain<RSRL_MOUNTAIN_CAR> {
    static constexpr int D = 2, A = 3;
    __host__ __device__ static constexpr double lo(int d) { return d == 0 ? -1.2 : -0.07; }
    __host__ __device__ static constexpr double hi(int d) { return d == 0 ? 0.6 : 0.07; }
    __host__ __device__ static constexpr double start(int d) { return d == 0 ? -0.5 : 0.0; }
    __host__ __device__ __forceinline__ static bool is_terminal(const double* s) { return s[0] >= 0.6; }
    __host__ __device__ __forceinline__ static double reward_of(bool terminal) { return terminal ? 0.0 : -1.0; }  // :88-92
    static constexpr bool kCheapStep = true;  // a transition costs ~100 instructions: all A candidates can be stepped ahead of time (persistent.cuh)
    // The action-independent part of the transition (the f64 cosine: half of its instructions) can be taken ahead of time: the
    // persistent kernel evaluates it for s_{t+1} while it waits for the grid exchange of step t.  step == step_post(step_pre).
    static constexpr bool kHasPre = true;
    __host__ __device__ __forceinline__ static double step_pre(const double* s) { return dmul(-0.0025, cos64(dmul(3.0, s[0]))); }
    template <bool ROLLED = false>
    __host__ __device__ __forceinline__ static void step(double* s, int action, double& reward, bool& terminal) {
        step_post(s, action, step_pre(s), reward, terminal);
    }
    __host__ __device__ __forceinline__ static void step_post(double* s, int action, double pre, double& reward, bool& terminal) {
        const double a = (double)(action - 1);                                        // ALL_ACTIONS = [-1, 0, 1]
        const double dv = dadd(dmul(0.001, a), pre);                                  // :58
        s[1] = dclip(-0.07, dadd(s[1], dv), 0.07);                                    // :63
        s[0] = dclip(-1.2, dadd(s[0], s[1]), 0.6);                                    // :64 (uses the new v)
        terminal = s[0] >= 0.6;                                                       // :77
        reward = terminal ? 0.0 : -1.0;                                               // :88-92
    }
};

// rsrl_domains/src/ode.rs:1-43 — k_i = f(...) * dx; y += (k1 + 2 k2 + 2 k3 + k4) / 6 in that association.
// ROLLED = true keeps one copy of the gradient code (a 4-trip loop, same operations in the same order, bit-identical
// results): used by kernels whose instruction footprint matters (f4tc.cuh).
template <bool ROLLED = false, class Grad>
__host__ __device__ __forceinline__ void runge_kutta4(Grad f, double* y, double dx) {
    if (ROLLED) {
        double k[4], tmp[4], sum[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { tmp[i] = y[i]; sum[i] = 0.0; }
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
            f(tmp, k);
            const bool mid = st == 1 || st == 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                k[i] = dmul(k[i], dx);
                sum[i] = st == 0 ? k[i] : dadd(sum[i], mid ? dmul(2.0, k[i]) : k[i]);
                tmp[i] = dadd(y[i], st < 2 ? dmul(k[i], 0.5) : k[i]);  // k/2.0 == k*0.5 bit for bit
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = dadd(y[i], ddiv(sum[i], 6.0));
        return;
    }
    double k1[4], k2[4], k3[4], k4[4], tmp[4];
    f(y, k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) { k1[i] = dmul(k1[i], dx); tmp[i] = dadd(y[i], dmul(k1[i], 0.5)); }  // k/2.0 == k*0.5 bit for bit
    f(tmp, k2);
#pragma unroll
    for (int i = 0; i < 4; ++i) { k2[i] = dmul(k2[i], dx); tmp[i] = dadd(y[i], dmul(k2[i], 0.5)); }
    f(tmp, k3);
#pragma unroll
    for (int i = 0; i < 4; ++i) { k3[i] = dmul(k3[i], dx); tmp[i] = dadd(y[i], k3[i]); }
    f(tmp, k4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        k4[i] = dmul(k4[i], dx);
        const double sum = dadd(dadd(dadd(k1[i], dmul(2.0, k2[i])), dmul(2.0, k3[i])), k4[i]);
        y[i] = dadd(y[i], ddiv(sum, 6.0));
    }
}

// rsrl_domains/src/cart_pole.rs:7-26,34-121
template <> struct Domain<RSRL_CART_POLE> {
    static constexpr int D = 4, A = 2;
    static constexpr double TWELVE_DEGREES = RSRL_PI / 15.0;  // consts.rs:10
    __host__ __device__ static constexpr double hi(int d) { return d == 0 ? 2.4 : d == 1 ? 6.0 : d == 2 ? TWELVE_DEGREES : 2.0; }
    __host__ __device__ static constexpr double lo(int d) { return -hi(d); }
    __host__ __device__ static constexpr double start(int) { return 0.0; }
    __host__ __device__ __forceinline__ static bool is_terminal(const double* s) {  // :83-97
        return s[0] <= -2.4 || s[0] >= 2.4 || s[2] <= -TWELVE_DEGREES || s[2] >= TWELVE_DEGREES;
    }
    __host__ __device__ __forceinline__ static double reward_of(bool terminal) { return terminal ? -1.0 : 0.0; }  // :23-24,103-107
    static constexpr bool kCheapStep = false, kHasPre = false;
    __host__ __device__ __forceinline__ static double step_pre(const double*) { return 0.0; }
    __host__ __device__ __forceinline__ static void step_post(double* s, int action, double, double& reward, bool& terminal) { step(s, action, reward, terminal); }
    template <bool ROLLED = false>
    __host__ __device__ __forceinline__ static void step(double* s, int action, double& reward, bool& terminal) {
        const double force = action == 0 ? -10.0 : 10.0;  // ALL_ACTIONS :26
        constexpr double POLE_MOMENT = 0.5 * 0.1, TOTAL_MASS = 1.0 + 0.1, FOUR_THIRDS = 4.0 / 3.0, G = 9.8;
        auto grad = [force](const double* b, double* out) {  // :52-72
            const double dx = b[1], theta = b[2], dtheta = b[3];
            double sin_t, cos_t;
            sincos64(theta, &sin_t, &cos_t);
            const double z = ddiv(dadd(force, dmul(dmul(dmul(POLE_MOMENT, dtheta), dtheta), sin_t)), TOTAL_MASS);
            const double numer = dsub(dmul(G, sin_t), dmul(cos_t, z));
            const double denom = dsub(dmul(FOUR_THIRDS, 0.5), dmul(dmul(POLE_MOMENT, cos_t), cos_t));
            out[0] = dx;
            out[2] = dtheta;
            out[3] = ddiv(numer, denom);
            out[1] = dsub(z, dmul(dmul(0.5, out[3]), cos_t));
        };
        double ns[4] = {s[0], s[1], s[2], s[3]};
        runge_kutta4<ROLLED>(grad, ns, 0.02);
        s[0] = dclip(-2.4, ns[0], 2.4);  // :44-49
        s[1] = dclip(-6.0, ns[1], 6.0);
        s[2] = dclip(-TWELVE_DEGREES, ns[2], TWELVE_DEGREES);
        s[3] = dclip(-2.0, ns[3], 2.0);
        terminal = is_terminal(s);
        reward = terminal ? -1.0 : 0.0;  // :23-24,103-107
    }
};

// rsrl_domains/src/acrobot.rs:8-36,51-152 — quirks kept verbatim
template <> struct Domain<RSRL_ACROBOT> {
    static constexpr int D = 4, A = 3;
    __host__ __device__ static constexpr double hi(int d) { return d < 2 ? RSRL_PI : d == 2 ? 4.0 * RSRL_PI : 9.0 * RSRL_PI; }
    __host__ __device__ static constexpr double lo(int d) { return -hi(d); }
    __host__ __device__ static constexpr double start(int) { return 0.0; }
    __host__ __device__ __forceinline__ static bool is_terminal(const double* s) {  // :56-58
        return dadd(cos64(s[0]), cos64(dadd(s[0], s[1]))) < -1.0;
    }
    __host__ __device__ __forceinline__ static double reward_of(bool terminal) { return terminal ? 0.0 : -1.0; }  // :32-33,134-138
    static constexpr bool kCheapStep = false, kHasPre = false;
    __host__ __device__ __forceinline__ static double step_pre(const double*) { return 0.0; }
    __host__ __device__ __forceinline__ static void step_post(double* s, int action, double, double& reward, bool& terminal) { step(s, action, reward, terminal); }
    template <bool ROLLED = false>
    __host__ __device__ __forceinline__ static void step(double* s, int action, double& reward, bool& terminal) {
        const double torque = (double)(action - 1);  // ALL_ACTIONS = [-1, 0, 1] :36
        constexpr double G = 9.8, PI_OVER_2 = RSRL_PI / 2.0;
        auto grad = [torque](const double* b, double* out) {  // :81-108 (M1=M2=L1=1, LC1=LC2=0.5, I1=I2=1)
            const double theta1 = b[0], theta2 = b[1], dtheta1 = b[2], dtheta2 = b[3];
            double sin_t2, cos_t2;
            sincos64(theta2, &sin_t2, &cos_t2);
            // d1 = M1*LC1*LC1 + M2*(L1*L1 + LC2*LC2 + 2*L1*LC2*cos_t2) + I1 + I2
            const double d1 = dadd(dadd(dadd(0.25, dmul(1.0, dadd(dadd(1.0, 0.25), dmul(1.0, cos_t2)))), 1.0), 1.0);
            // d2 = M2*(LC2*LC2 + L1*LC2*cos_t2) + I2
            const double d2 = dadd(dmul(1.0, dadd(0.25, dmul(0.5, cos_t2))), 1.0);
            // phi2 = M2*LC2*G*cos(theta1 + theta2 - PI/2)
            const double phi2 = dmul(dmul(0.5, G), cos64(dsub(dadd(theta1, theta2), PI_OVER_2)));
            // phi1 = -1*L1*LC2*dth2*dth2*sin_t2 - 2*M2*L1*LC2*dth2*dth1*sin_t2 + (M1*LC1 + M2*L1)*G*cos(th1 - PI/2) + phi2
            const double t1 = dmul(dmul(dmul(-0.5, dtheta2), dtheta2), sin_t2);
            const double t2 = dmul(dmul(dmul(1.0, dtheta2), dtheta1), sin_t2);  // 2.0*1*1*0.5 = 1.0 exactly
            const double t3 = dmul(dmul(1.5, G), cos64(dsub(theta1, PI_OVER_2)));
            const double phi1 = dadd(dadd(dsub(t1, t2), t3), phi2);
            out[0] = dtheta1;
            out[1] = dtheta2;
            // (torque + d2/d1*phi1 - M2*L1*LC2*dth1*dth1*sin_t2 - phi2) / (M2*LC2*LC2 + I2 - d2*d2/d1)
            const double num = dsub(dsub(dadd(torque, dmul(ddiv(d2, d1), phi1)), dmul(dmul(dmul(0.5, dtheta1), dtheta1), sin_t2)), phi2);
            const double den = dsub(dadd(0.25, 1.0), ddiv(dmul(d2, d2), d1));
            out[2] = ddiv(num, den);
            out[3] = ddiv(-dadd(dmul(d2, out[2]), phi1), d1);
        };
        double ns[4] = {s[0], s[1], s[2], s[3]};
        runge_kutta4<ROLLED>(grad, ns, 0.2);
        s[0] = dwrap(-RSRL_PI, ns[0], RSRL_PI);  // :64-78
        s[1] = dwrap(-RSRL_PI, ns[1], RSRL_PI);
        s[2] = dclip(-4.0 * RSRL_PI, ns[2], 4.0 * RSRL_PI);
        s[3] = dclip(-9.0 * RSRL_PI, ns[3], 9.0 * RSRL_PI);
        terminal = is_terminal(s);
        reward = terminal ? 0.0 : -1.0;  // :32-33,134-138
    }
};

// start state of the episode beginning at batched step t (Domain::default() or U[lo,hi) from Philox)
template <class Dom>
__host__ __device__ __forceinline__ void fresh_state(double* s, int init_mode, const double* init_lo, const double* init_hi,
                                            uint64_t seed, uint64_t g, uint64_t t) {
    if (init_mode == RSRL_INIT_DEFAULT) {
#pragma unroll
        for (int d = 0; d < Dom::D; ++d) s[d] = Dom::start(d);
    } else {
        const uint4 r = draw4(seed, g, t, STREAM_INIT);
#pragma unroll
        for (int d = 0; d < Dom::D; ++d)
            s[d] = dadd(init_lo[d], dmul(dsub(init_hi[d], init_lo[d]), dmul((double)word(r, d), 1.0 / 4294967296.0)));
    }
}

// ---------------------------------------------------------------------------
// Tensor-grid bases.  lfa::basis::Fourier::project computes, for every coefficient vector c in
// {0..P}^D \ {0} (descending lexicographic order) phi = cos(pi * sum_d c_d * x^_d), then
// .with_bias() appends 1.0.  Row k of that order has digits c = P - digit_d(k), and the skipped
// all-zero vector would land exactly on the bias slot k = F-1 (cos 0 = 1), so the whole feature
// vector is the real part of the tensor product  prod_d exp(i*pi*c_d*x^_d)  over the full grid.
// We build per-dimension tables cos/sin(pi*c*x^_d), c = 1..P, by angle addition and combine.
// Polynomial: phi = prod_d x_d^{c_d} on the raw state, same grid order (project-defined).
// ---------------------------------------------------------------------------
template <typename R, int D, int P, int BASIS>
struct GridTables {
    R c[D][P];  // c[d][j] = cos(pi*(j+1)*x^_d)   (Polynomial: x_d^(j+1))
    R s[D][P];  // s[d][j] = sin(pi*(j+1)*x^_d)   (Polynomial: unused)
};

// grid_prepare = grid_prepare_base (per dimension: cos / sin of pi x^_d, or the raw x_d) + grid_expand (angle addition / powers).
// The split lets the persistent kernel compute the base of a state ahead of time and expand it when the tables are needed.
template <typename R, class Dom, int P, int BASIS>
__host__ __device__ __forceinline__ void grid_prepare_base(const double* st, R (*base)[2]) {
    using O = RealOps<R>;
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) {
        if (BASIS == RSRL_FOURIER) {
            // scaled = (v - lo) / (hi - lo) in f64 exactly like the reference, then to the compute type.
            // f32: the quotient is rounded to fp32 anyway, so multiply by the f64 reciprocal (error 1e-16 << 6e-8)
            // instead of a ~20-instruction f64 division.
            const double num = dsub(st[d], Dom::lo(d));
            const R xh = sizeof(R) == 4 ? (R)dmul(num, 1.0 / (Dom::hi(d) - Dom::lo(d)))
                                        : (R)ddiv(num, dsub(Dom::hi(d), Dom::lo(d)));
            R s1, c1;
            O::sincospi(xh, &s1, &c1);
            base[d][0] = c1;
            base[d][1] = s1;
        } else {
            base[d][0] = (R)st[d];
            base[d][1] = (R)0;
        }
    }
}
template <typename R, class Dom, int P, int BASIS>
__host__ __device__ __forceinline__ void grid_expand(const R (*base)[2], GridTables<R, Dom::D, P, BASIS>& t) {
    using O = RealOps<R>;
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) {
        if (BASIS == RSRL_FOURIER) {
            const R c1 = base[d][0], s1 = base[d][1];
            t.c[d][0] = c1;
            t.s[d][0] = s1;
#pragma unroll
            for (int j = 1; j < P; ++j) {  // angle addition: (j+1)*theta = j*theta + theta
                t.c[d][j] = O::fma(t.c[d][j - 1], c1, -(t.s[d][j - 1] * s1));
                t.s[d][j] = O::fma(t.s[d][j - 1], c1, t.c[d][j - 1] * s1);
            }
        } else {
            const R x = base[d][0];
            t.c[d][0] = x;
#pragma unroll
            for (int j = 1; j < P; ++j) t.c[d][j] = t.c[d][j - 1] * x;
        }
    }
}
template <typename R, class Dom, int P, int BASIS>
__host__ __device__ __forceinline__ void grid_prepare(const double* st, GridTables<R, Dom::D, P, BASIS>& t) {
    R base[Dom::D][2];
    grid_prepare_base<R, Dom, P, BASIS>(st, base);
    grid_expand<R, Dom, P, BASIS>(base, t);
}

// complex helper on (re, im) pairs with compile-time "c == 0 => 1 + 0i" shortcut
template <typename R, int D, int P, int BASIS>
struct GridBasis {
    static constexpr int N1 = P + 1;
    static constexpr int F = (D == 2) ? N1 * N1 : N1 * N1 * N1 * N1;
    using Tab = GridTables<R, D, P, BASIS>;

    // calls f(k, phi_k) for k = 0..F-1 in feature order; fully unrolled => k is a compile-time constant
    // PAIR (device, fp32, D = 2, Fourier): two features per issue slot.  Measured: -2 % step time in the SHARED-weights kernel (14 warps per
    // SM, latency bound), +50 % in the PER_ENV kernel (FMUL2 / FFMA2 issue at a fraction of the scalar rate) — so only the former asks for it.
    template <bool PAIR = false, class Fn>
    __host__ __device__ __forceinline__ static void for_each(const Tab& t, Fn f) {
        using O = RealOps<R>;
#ifdef __CUDA_ARCH__
        if constexpr (PAIR && D == 2 && BASIS == RSRL_FOURIER && sizeof(R) == 4) {
            // fp32 on the device: two features per issue slot (FMUL2 + FFMA2 with the dimension-0 entry broadcast from a scalar
            // register).  Per feature the operations are those of the scalar form below — (-s0) * s1 == -(s0 * s1) exactly —
            // so the host build (oracle32) keeps the scalar form.
#pragma unroll
            for (int i0 = 0; i0 < N1; ++i0) {
                const int c0 = P - i0;
                if (c0 == 0) {
#pragma unroll
                    for (int i1 = 0; i1 < N1; ++i1) f(i0 * N1 + i1, i1 == P ? (R)1 : t.c[1][P - i1 - 1]);
                } else {
                    const float ca = t.c[0][c0 - 1], nsa = -t.s[0][c0 - 1];
#pragma unroll
                    for (int i1 = 0; i1 < P; i1 += 2) {
                        const int c1 = P - i1;
                        if (c1 >= 2) {
                            const float2 m = __fmul2_rn(make_float2(nsa, nsa), make_float2(t.s[1][c1 - 1], t.s[1][c1 - 2]));
                            const float2 ph = __ffma2_rn(make_float2(ca, ca), make_float2(t.c[1][c1 - 1], t.c[1][c1 - 2]), m);
                            f(i0 * N1 + i1, ph.x);
                            f(i0 * N1 + i1 + 1, ph.y);
                        } else {
                            f(i0 * N1 + i1, ffma(ca, t.c[1][0], fmul(nsa, t.s[1][0])));
                        }
                    }
                    f(i0 * N1 + P, ca);
                }
            }
            return;
        }
#endif
        if (D == 2) {
#pragma unroll
            for (int i0 = 0; i0 < N1; ++i0) {
#pragma unroll
                for (int i1 = 0; i1 < N1; ++i1) {
                    const int c0 = P - i0, c1 = P - i1;
                    R phi;
                    if (BASIS == RSRL_FOURIER) {
                        if (c0 == 0 && c1 == 0) phi = (R)1;
                        else if (c0 == 0) phi = t.c[1][c1 - 1];
                        else if (c1 == 0) phi = t.c[0][c0 - 1];
                        else phi = O::fma(t.c[0][c0 - 1], t.c[1][c1 - 1], -(t.s[0][c0 - 1] * t.s[1][c1 - 1]));
                    } else {
                        if (c0 == 0 && c1 == 0) phi = (R)1;
                        else if (c0 == 0) phi = t.c[1][c1 - 1];
                        else if (c1 == 0) phi = t.c[0][c0 - 1];
                        else phi = t.c[0][c0 - 1] * t.c[1][c1 - 1];
                    }
                    f(i0 * N1 + i1, phi);
                }
            }
        } else {
#pragma unroll
            for (int i0 = 0; i0 < N1; ++i0) {
#pragma unroll
                for (int i1 = 0; i1 < N1; ++i1) {
                    const int c0 = P - i0, c1 = P - i1;
                    // z01 = e(c0) * e(c1)
                    R re01, im01;
                    cmul(t, 0, c0, 1, c1, re01, im01);
#pragma unroll
                    for (int i2 = 0; i2 < N1; ++i2) {
                        const int c2 = P - i2;
                        R re012, im012;
                        cmul_acc(t, re01, im01, 2, c2, re012, im012);
#pragma unroll
                        for (int i3 = 0; i3 < N1; ++i3) {
                            const int c3 = P - i3;
                            R phi;
                            if (BASIS == RSRL_FOURIER) {
                                if (c3 == 0) phi = re012;
                                else phi = O::fma(re012, t.c[3][c3 - 1], -(im012 * t.s[3][c3 - 1]));
                            } else {
                                phi = c3 == 0 ? re012 : re012 * t.c[3][c3 - 1];
                            }
                            f(((i0 * N1 + i1) * N1 + i2) * N1 + i3, phi);
                        }
                    }
                }
            }
        }
    }

    __host__ __device__ __forceinline__ static void cmul(const Tab& t, int d0, int c0, int d1, int c1, R& re, R& im) {
        using O = RealOps<R>;
        const R a = c0 == 0 ? (R)1 : t.c[d0][c0 - 1], b = (c0 == 0 || BASIS != RSRL_FOURIER) ? (R)0 : t.s[d0][c0 - 1];
        if (c1 == 0) { re = a; im = b; return; }
        if (BASIS == RSRL_FOURIER) {
            re = O::fma(a, t.c[d1][c1 - 1], -(b * t.s[d1][c1 - 1]));
            im = O::fma(a, t.s[d1][c1 - 1], b * t.c[d1][c1 - 1]);
        } else { re = a * t.c[d1][c1 - 1]; im = (R)0; }
    }
    __host__ __device__ __forceinline__ static void cmul_acc(const Tab& t, R a, R b, int d1, int c1, R& re, R& im) {
        using O = RealOps<R>;
        if (c1 == 0) { re = a; im = b; return; }
        if (BASIS == RSRL_FOURIER) {
            re = O::fma(a, t.c[d1][c1 - 1], -(b * t.s[d1][c1 - 1]));
            im = O::fma(a, t.s[d1][c1 - 1], b * t.c[d1][c1 - 1]);
        } else { re = a * t.c[d1][c1 - 1]; im = (R)0; }
    }
};

// ---------------------------------------------------------------------------
// TileCoding (project-defined spec, identical to oracle/rsrl_oracle.c:orc_tile_indices; the lfa crate's
// hasher is a user-supplied BuildHasher, so there is no reference behaviour to match):
//   x^ = (x - lo) / (hi - lo);  q_d = floor(x^_d * P * T);  tiling t: coord_d = (q_d + t*(1+2d)) / T;
//   row = fmix32(FNV-style combine(t, coord)) & (M - 1);  active set = unique rows, activation 1.0.
// Integer work: bit-exact with the oracle.
// ---------------------------------------------------------------------------
constexpr int kMaxTilings = 16;
struct TileTab {
    int32_t idx[kMaxTilings];  // idx[t], t < n: the row of tiling t, or -1 when an earlier tiling already activated that row (the set of active
                               // rows is what counts: a HashMap of activations collapses duplicates); walk it in order and skip the -1s
    int n;
};
struct TileParams {
    int n_tilings, tiles_per_dim, memory_mask;
    uint32_t div_magic;  // ceil(2^32 / n_tilings): n / n_tilings == umulhi(n, div_magic) for 0 <= n < 2^24 (host: tile_div_magic)
};
__host__ __device__ __forceinline__ uint32_t tile_div_magic(int n_tilings) {
    return (uint32_t)((0x100000000ull + (uint64_t)n_tilings - 1ull) / (uint64_t)n_tilings);
}
// (q + offset) / n_tilings, 32 times per state: the run-time divisor costs ~20 instructions per division on the GPU (half of
// the TileCoding kernel's instructions); multiply-high by the precomputed reciprocal is exact for the small non-negative
// numerators of in-range states (n * (magic * T - 2^32) < 2^32 holds for n < 2^28, T <= 16); anything else divides.
__host__ __device__ __forceinline__ int32_t tile_div(int32_t n, const TileParams& tp) {
#ifdef __CUDA_ARCH__
    if (tp.div_magic != 0u && n >= 0 && n < (1 << 24)) return (int32_t)__umulhi((uint32_t)n, tp.div_magic);  // (one tiling: 2^32 does not fit, magic = 0)
#endif
    return n / tp.n_tilings;
}

__host__ __device__ __forceinline__ uint32_t tile_hash(uint32_t tiling, const int32_t* coord, int D) {
    uint32_t h = (tiling + 1u) * 0x9E3779B1u;
    for (int d = 0; d < D; ++d) h = (h ^ (uint32_t)coord[d]) * 0x85EBCA6Bu;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// TMAX: compile-time bound of the tiling loops (>= tp.n_tilings); 8 halves the unrolled hash / dedup code of 8-tiling configs
template <class Dom, int TMAX = kMaxTilings>
__host__ __device__ __forceinline__ void tile_prepare(const double* st, const TileParams& tp, TileTab& tab) {
    int32_t q[Dom::D];
#pragma unroll
    for (int d = 0; d < Dom::D; ++d) {
        const double xh = ddiv(dsub(st[d], Dom::lo(d)), dsub(Dom::hi(d), Dom::lo(d)));
        q[d] = (int32_t)floor(dmul(dmul(xh, (double)tp.tiles_per_dim), (double)tp.n_tilings));
    }
    tab.n = tp.n_tilings;
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
        if (t < tp.n_tilings) {
            int32_t coord[Dom::D];
#pragma unroll
            for (int d = 0; d < Dom::D; ++d) coord[d] = tile_div(q[d] + t * (1 + 2 * d), tp);
            const int32_t row = (int32_t)(tile_hash((uint32_t)t, coord, Dom::D) & (uint32_t)tp.memory_mask);
            bool dup = false;
#pragma unroll
            for (int j = 0; j < t; ++j) dup |= tab.idx[j] == row;  // (earlier duplicates are -1: they never match)
            tab.idx[t] = dup ? -1 : row;
        }
    }
}

// ---------------------------------------------------------------------------
// argmax family + policies
// ---------------------------------------------------------------------------
// utils.rs:6-21 argmaxima: returns the tie set as a bit mask; a value within 1e-7 of `max` joins
// without raising `max`.
template <typename R, int A>
__host__ __device__ __forceinline__ uint32_t argmaxima(const R* q, int& count) {
    using O = RealOps<R>;
    R mx = O::lowest();
    uint32_t mask = 0;
    count = 0;
#pragma unroll
    for (int i = 0; i < A; ++i) {
        if (O::abs(q[i] - mx) < (R)1e-7) { mask |= 1u << i; ++count; }
        else if (q[i] > mx) { mx = q[i]; mask = 1u << i; count = 1; }
    }
    return mask;
}

// core.rs:96-105 find_max: exact compare, the LAST maximal index wins, NaN replaces the accumulator
template <typename R, int A>
__host__ __device__ __forceinline__ int find_max(const R* q, R& mx) {
    int idx = 0;
    mx = q[0];
#pragma unroll
    for (int i = 1; i < A; ++i) if (!(mx > q[i])) { idx = i; mx = q[i]; }
    return idx;
}

// utils.rs:23-34 argmax_first: new best only if y - x > 1e-7, first wins
template <typename R, int A>
__host__ __device__ __forceinline__ int argmax_first(const R* q) {
    using O = RealOps<R>;
    int idx = 0;
    R x = O::lowest();
#pragma unroll
    for (int j = 0; j < A; ++j) if (q[j] - x > (R)1e-7) { idx = j; x = q[j]; }
    return idx;
}

__host__ __device__ __forceinline__ int nth_set_bit(uint32_t mask, int n) {
    for (int i = 0; i < n; ++i) mask &= mask - 1;
    return ffs32(mask) - 1;
}

struct PolicyParams {
    int policy;           // rsrl_policy_t
    uint32_t eps_thresh;  // floor(eps * 2^32)
    int eps_always;       // eps >= 1 (rand's gen_bool(1.0) is true without drawing)
    uint64_t seed;
    double tau;           // Softmax temperature (the config's `epsilon` field)
};

// softmax.rs:15-36 softmax_stable: c = fold(NAN, f64::max); v_i = exp((q_i - c) / tau); p_i = min(v_i / sum v, MAX)
template <typename R, int A>
__host__ __device__ __forceinline__ void softmax_probs(R tau, const R* q, R* p) {
    using O = RealOps<R>;
    R c = q[0];
#pragma unroll
    for (int i = 1; i < A; ++i) c = O::max_nan(c, q[i]);
    R z = (R)0;
#pragma unroll
    for (int i = 0; i < A; ++i) { p[i] = O::exp((q[i] - c) / tau); z += p[i]; }
#pragma unroll
    for (int i = 0; i < A; ++i) p[i] = O::min_max(p[i] / z);
}

// Policy::sample — greedy.rs:77-81 (argmax_choose_rng: RNG only on ties), epsilon_greedy.rs:74-80,
// random.rs:43-45.  rnd.x -> gen_bool(eps), rnd.y -> Uniform(0, A), rnd.z -> choose among maxima.
template <typename R, int A>
__host__ __device__ __forceinline__ int policy_sample(const PolicyParams& p, const R* q, uint64_t g, uint64_t draw,
                                             uint32_t stream, bool& nonfinite) {
    if (p.policy == RSRL_SOFTMAX) {
        // policies/mod.rs:46-61 sample_probs_with_rng: r = rng.gen::<f64>() (53 bits: rnd.x high, rnd.y low);
        // first index whose running sum exceeds r, else the last
        R pr[A];
        softmax_probs<R, A>((R)p.tau, q, pr);
        const uint4 rn = draw4(p.seed, g, draw, stream);
        const double r = (double)((((unsigned long long)rn.x << 32) | (unsigned long long)rn.y) >> 11) * (1.0 / 9007199254740992.0);
        R cum = (R)0;
#pragma unroll
        for (int i = 0; i < A; ++i) {
            cum = cum + pr[i];
            if ((double)cum > r) return i;
        }
        return A - 1;
    }
    if (p.policy != RSRL_GREEDY) {
        const uint4 r = draw4(p.seed, g, draw, stream);
        const bool explore = p.policy == RSRL_RANDOM || p.eps_always || r.x < p.eps_thresh;
        if (explore) return (int)umulhi32(r.y, (uint32_t)A);
        int cnt;
        const uint32_t mask = argmaxima<R, A>(q, cnt);
        if (cnt == 0) { nonfinite = true; return 0; }
        if (cnt == 1) return ffs32(mask) - 1;
        return nth_set_bit(mask, (int)umulhi32(r.z, (uint32_t)cnt));
    }
    int cnt;
    const uint32_t mask = argmaxima<R, A>(q, cnt);
    if (cnt == 0) { nonfinite = true; return 0; }
    if (cnt == 1) return ffs32(mask) - 1;
    const uint4 r = draw4(p.seed, g, draw, stream);  // rare: only on ties
    return nth_set_bit(mask, (int)umulhi32(r.z, (uint32_t)cnt));
}

// Function<(S,)>::evaluate of the policy (greedy.rs:30-44, epsilon_greedy.rs:38-45)
template <typename R, int A>
__host__ __device__ __forceinline__ void policy_probs(int policy, R eps, const R* q, R* p) {
    if (policy == RSRL_SOFTMAX) { softmax_probs<R, A>(eps, q, p); return; }  // eps carries tau
    if (policy == RSRL_RANDOM) {
#pragma unroll
        for (int i = 0; i < A; ++i) p[i] = (R)1 / (R)A;
        return;
    }
    int cnt;
    const uint32_t mask = argmaxima<R, A>(q, cnt);
    const R pg = (R)1 / (R)cnt;
#pragma unroll
    for (int i = 0; i < A; ++i) p[i] = ((mask >> i) & 1u) ? pg : (R)0;
    if (policy == RSRL_EPSILON_GREEDY) {
        const R pr = eps / (R)A;
#pragma unroll
        for (int i = 0; i < A; ++i) p[i] = pr + p[i] * ((R)1 - eps);
    }
}

// Policy::mode — greedy.rs:83 find_max; softmax.rs:141 argmax_first over the probabilities
template <typename R, int A>
__host__ __device__ __forceinline__ int policy_mode(int policy, R tau, const R* q) {
    if (policy == RSRL_SOFTMAX) {
        R p[A];
        softmax_probs<R, A>(tau, q, p);
        return argmax_first<R, A>(p);
    }
    R mx;
    return find_max<R, A>(q, mx);
}

// traces.rs:196-240
template <typename R>
__host__ __device__ __forceinline__ R trace_rule(int rule, R rate, R z, R grad) {
    const R v = RealOps<R>::mac(rate, z, grad);  // traces.rs:200,218,238 `rate * x + y`: unfused in f64 (the reference's ops), fused in fp32
    return rule == RSRL_TRACE_REPLACE ? RealOps<R>::clamp1(v) : v;
}

}  // namespace rsrl
