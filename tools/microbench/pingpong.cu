// L2 flag ping-pong between two CTAs on different SMs: one-way latency of the LL publish -> poll path.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu && ./pingpong
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 ld_gpu(const uint4* p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_gpu(uint4* p, unsigned e) { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(e) : "memory"); }
__device__ __forceinline__ unsigned ld_vol(const unsigned* p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__global__ void pingpong(uint4* a, uint4* b, int iters, long long* out, int peer_block) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) { st_gpu(a, i); while (ld_gpu(b).w != (unsigned)i) {} }
        out[0] = clock64() - t0;
    } else if (blockIdx.x == peer_block) {
        for (int i = 1; i <= iters; ++i) { while (ld_gpu(a).w != (unsigned)i) {} st_gpu(b, i); }
    }
}
// N pollers on the same line (hot spot) while one writer publishes: time until the LAST poller sees it
__global__ void fanout(uint4* a, unsigned* done, int iters, long long* out) {
    if (threadIdx.x != 0) return;
    const int G = gridDim.x;
    for (int i = 1; i <= iters; ++i) {
        if (blockIdx.x == 0) {
            long long t0 = clock64();
            st_gpu(a, i);
            while (ld_vol(done) < (unsigned)(i * (G - 1))) {}
            out[1] += clock64() - t0;
        } else {
            while (ld_gpu(a).w != (unsigned)i) {}
            atomicAdd(done, 1u);
        }
    }
}
int main() {
    uint4 *a, *b; long long* out; unsigned* done;
    cudaMalloc(&a, 4096); cudaMalloc(&b, 4096); cudaMalloc(&out, 64); cudaMalloc(&done, 4);
    cudaMemset(a, 0, 4096); cudaMemset(b, 0, 4096); cudaMemset(out, 0, 64); cudaMemset(done, 0, 4);
    int iters = 2000;
    for (int peer : {1, 2, 73, 147}) {
        cudaMemset(a, 0, 4096); cudaMemset(b, 0, 4096);
        pingpong<<<148, 32>>>(a, b + 64, iters, out, peer);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("pingpong block0 <-> block%d: round trip %.0f cycles, one way %.0f\n", peer, (double)h / iters, (double)h / iters / 2);
    }
    for (int g : {2, 13, 148}) {
        cudaMemset(a, 0, 4096); cudaMemset(done, 0, 4); cudaMemset(out, 0, 64);
        void* args[] = {&a, &done, &iters, &out};
        cudaLaunchCooperativeKernel((void*)fanout, dim3(g), dim3(32), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("fanout 1 writer -> %d pollers (+atomic ack): %.0f cycles per round (%s)\n", g - 1, (double)h[1] / iters, cudaGetErrorString(e));
    }
    return 0;
}
