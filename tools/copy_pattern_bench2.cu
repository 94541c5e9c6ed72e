// Micro-benchmark 2 (not product code): what does a PERSISTENT copy kernel need to match the huge-grid copy?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint4 ld(const uint4* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st(uint4* p, uint4 v) { asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
// persistent, per-warp contiguous tiles of ROWS rows, U rows in flight, THREADS per CTA, grid = 148 * CTAS
template <int ROWS, int U, int THREADS> __global__ void __launch_bounds__(THREADS) pat_a(const uint4* s, uint4* d, size_t ntiles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = THREADS / 32;
    for (size_t t = blockIdx.x + (size_t)warp * gridDim.x; t < ntiles; t += (size_t)gridDim.x * W) {
        const size_t v0 = t * ROWS * 32;
        for (int r = 0; r + U <= ROWS; r += U) {
            uint4 x[U];
#pragma unroll
            for (int u = 0; u < U; u++) x[u] = ld(s + v0 + (size_t)(r + u) * 32 + lane);
#pragma unroll
            for (int u = 0; u < U; u++) st(d + v0 + (size_t)(r + u) * 32 + lane, x[u]);
        }
    }
}
// software-pipelined: loads for the next group are issued before the stores of the current group
template <int ROWS, int U, int THREADS> __global__ void __launch_bounds__(THREADS) pat_p(const uint4* s, uint4* d, size_t ntiles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = THREADS / 32;
    for (size_t t = blockIdx.x + (size_t)warp * gridDim.x; t < ntiles; t += (size_t)gridDim.x * W) {
        const size_t v0 = t * ROWS * 32;
        uint4 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; u++) x[u] = ld(s + v0 + (size_t)u * 32 + lane);
        for (int r = 0; r < ROWS; r += 2 * U) {
#pragma unroll
            for (int u = 0; u < U; u++) y[u] = ld(s + v0 + (size_t)(r + U + u) * 32 + lane);
#pragma unroll
            for (int u = 0; u < U; u++) st(d + v0 + (size_t)(r + u) * 32 + lane, x[u]);
            if (r + 2 * U < ROWS) {
#pragma unroll
                for (int u = 0; u < U; u++) x[u] = ld(s + v0 + (size_t)(r + 2 * U + u) * 32 + lane);
            }
#pragma unroll
            for (int u = 0; u < U; u++) st(d + v0 + (size_t)(r + U + u) * 32 + lane, y[u]);
        }
    }
}
int main() {
    const size_t bytes = 8ull << 30, nvec = bytes / 16;
    uint4 *s, *d; cudaMalloc(&s, bytes); cudaMalloc(&d, bytes); cudaMemset(s, 1, bytes); cudaMemset(d, 0, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto report = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaEventRecord(a); for (int i = 0; i < 10; i++) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
        cudaError_t e = cudaGetLastError();
        printf("%-44s %7.3f ms  %7.1f GB/s  %5.1f%%  %s\n", name, ms, 2.0 * bytes / ms / 1e6, 2.0 * bytes / ms / 1e6 / 6538.9 * 100, e ? cudaGetErrorString(e) : "");
    };
    const size_t nt64 = nvec / (64 * 32);
    report("1024 thr x1 CTA/SM  U=4  (today)", [&] { pat_a<64, 4, 1024><<<148, 1024>>>(s, d, nt64); });
    report("1024 thr x1 CTA/SM  U=8", [&] { pat_a<64, 8, 1024><<<148, 1024>>>(s, d, nt64); });
    report("1024 thr x1 CTA/SM  U=16", [&] { pat_a<64, 16, 1024><<<148, 1024>>>(s, d, nt64); });
    report("1024 thr x2 CTA/SM  U=4", [&] { pat_a<64, 4, 1024><<<296, 1024>>>(s, d, nt64); });
    report("512 thr x1 CTA/SM  U=8", [&] { pat_a<64, 8, 512><<<148, 512>>>(s, d, nt64); });
    report("512 thr x1 CTA/SM  U=16", [&] { pat_a<64, 16, 512><<<148, 512>>>(s, d, nt64); });
    report("256 thr x8 CTA/SM  U=4", [&] { pat_a<64, 4, 256><<<148 * 8, 256>>>(s, d, nt64); });
    report("1024 thr x1  pipelined U=4+4", [&] { pat_p<64, 4, 1024><<<148, 1024>>>(s, d, nt64); });
    report("1024 thr x1  pipelined U=8+8", [&] { pat_p<64, 8, 1024><<<148, 1024>>>(s, d, nt64); });
    report("512 thr x1  pipelined U=8+8", [&] { pat_p<64, 8, 512><<<148, 512>>>(s, d, nt64); });
    report("1024 thr x2  pipelined U=4+4", [&] { pat_p<64, 4, 1024><<<296, 1024>>>(s, d, nt64); });
    return 0;
}
