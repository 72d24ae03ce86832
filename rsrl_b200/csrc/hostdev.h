// hostdev.h — lets the per-env arithmetic headers (device.cuh, core.cuh) compile for the HOST as well.
//
// The device arithmetic of the bench dtype (f32 features / Q / weights, f64 physics) is written once, in
// device.cuh + core.cuh, with every rounding spelled out (explicit fma, no contraction: the .cu units
// that instantiate it are built with --fmad=false, the host build with -ffp-contract=off) and with its own
// elementary functions (cos/sin in f64, sincospi/exp in f32: device.cuh "rsrl math") instead of the CUDA
// and glibc math libraries, which round differently.  oracle/oracle32.cpp compiles these same headers with
// g++ — so a free-running f32 trajectory on the GPU can be checked BIT FOR BIT on the CPU (SURVEY 7.1
// step 0 "oracle32").  Nothing in the product links the host build.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define RSRL_HOST_BUILD 0
#else
#define RSRL_HOST_BUILD 1
#include <string.h>
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
#endif

namespace rsrl {

// ---- single-rounding primitives: identical bits on both sides ----
__host__ __device__ __forceinline__ double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
__host__ __device__ __forceinline__ double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
__host__ __device__ __forceinline__ double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
__host__ __device__ __forceinline__ double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
__host__ __device__ __forceinline__ double dfma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
__host__ __device__ __forceinline__ float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
__host__ __device__ __forceinline__ float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
__host__ __device__ __forceinline__ float ffma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
__host__ __device__ __forceinline__ uint32_t umulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
__host__ __device__ __forceinline__ int ffs32(uint32_t m) {  // 1-based index of the lowest set bit, 0 if none
#ifdef __CUDA_ARCH__
    return __ffs((int)m);
#else
    return __builtin_ffs((int)m);
#endif
}
__host__ __device__ __forceinline__ float bits_to_float(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

}  // namespace rsrl
