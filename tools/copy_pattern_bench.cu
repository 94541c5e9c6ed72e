// copy_pattern_bench.cu -- micro-benchmark (not product code): which way of walking memory gets a plain 8 GiB copy closest to
// the HBM copy peak?  The finding behind pack_kernel's tile scheduler.  One file, four patterns:
//   grid    classic grid-stride vector copy over a huge grid (what library copy kernels do)
//   chunk   the same body launched once per chunk: the HARDWARE block scheduler hands out chunks in order      (106 % of peak)
//   static  persistent (148 x 1024 threads), every warp streams 32 KiB tiles dealt round robin                 (91-95 %)
//   dynamic persistent, every warp takes the NEXT tile from a global counter, tile size swept                   (101 % at 32 KiB)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/copy_pattern_bench.cu -o tools/copy_pattern_bench && tools/copy_pattern_bench [pattern ...]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
__device__ __forceinline__ uint4 ld(const uint4* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st(uint4* p, uint4 v) { asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__global__ void fill(uint4* p, size_t n) { for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint32_t x = (uint32_t)(i * 2654435761u); p[i] = make_uint4(x, x * 3 + 1, x ^ 0x9e3779b9u, x * 7); } }

template <int ROWS, int U> __device__ __forceinline__ void copy_tile(const uint4* s, uint4* d, size_t v0, int lane) {
    for (int r = 0; r < ROWS; r += U) {
        uint4 x[U];
#pragma unroll
        for (int u = 0; u < U; u++) x[u] = ld(s + v0 + (size_t)(r + u) * 32 + lane);
#pragma unroll
        for (int u = 0; u < U; u++) st(d + v0 + (size_t)(r + u) * 32 + lane, x[u]);
    }
}
// dynamic: each warp grabs tiles of ROWS rows (512 B each) from a global counter
template <int ROWS, int U> __global__ void __launch_bounds__(1024, 1) warp_dyn(const uint4* s, uint4* d, unsigned ntiles, unsigned* counter) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(counter, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= ntiles) break;
        copy_tile<ROWS, U>(s, d, (size_t)t * ROWS * 32, lane);
    }
}
// static: tiles dealt round robin over the persistent warps
template <int ROWS, int U> __global__ void __launch_bounds__(1024, 1) warp_static(const uint4* s, uint4* d, unsigned ntiles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned t = blockIdx.x + warp * gridDim.x; t < ntiles; t += gridDim.x * 32) copy_tile<ROWS, U>(s, d, (size_t)t * ROWS * 32, lane);
}
// chunk: one CTA per 1 MiB chunk (32 warps x one 32 KiB tile), handed out by the hardware block scheduler
template <int U> __global__ void __launch_bounds__(1024, 1) cta_chunk(const uint4* s, uint4* d) {
    copy_tile<64, U>(s, d, ((size_t)blockIdx.x * 32 + (threadIdx.x >> 5)) * 64 * 32, threadIdx.x & 31);
}
// grid: grid-stride copy, 4 vectors in flight per thread
__global__ void __launch_bounds__(256) grid_stride(const uint4* s, uint4* d, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) x[u] = ld(s + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; u++) st(d + i + u * stride, x[u]);
    }
}
int main(int argc, char** argv) {
    const size_t bytes = 8ull << 30, nvec = bytes / 16;
    uint4 *s, *d; cudaMalloc(&s, bytes); cudaMalloc(&d, bytes); fill<<<148 * 8, 256>>>(s, nvec); cudaMemset(d, 0, bytes);
    unsigned* ctr; cudaMalloc(&ctr, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto want = [&](const char* p) { if (argc < 2) return true; for (int i = 1; i < argc; i++) if (!strcmp(argv[i], p)) return true; return false; };
    auto report = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaEventRecord(a); for (int i = 0; i < 10; i++) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
        cudaError_t e = cudaGetLastError();
        printf("%-44s %7.3f ms  %7.1f GB/s  %5.1f%% of 6538.9  %s\n", name, ms, 2.0 * bytes / ms / 1e6, 2.0 * bytes / ms / 1e6 / 6538.9 * 100, e ? cudaGetErrorString(e) : "");
    };
    const unsigned nt64 = (unsigned)(nvec / (64 * 32));
    if (want("grid")) report("grid-stride, 148*32 CTAs of 256", [&] { grid_stride<<<148 * 32, 256>>>(s, d, nvec); });
    if (want("chunk")) report("one CTA per 1 MiB chunk (hardware scheduler)", [&] { cta_chunk<4><<<nt64 / 32, 1024>>>(s, d); });
    if (want("static")) report("static  64 rows (32 KiB tiles), U=4", [&] { warp_static<64, 4><<<148, 1024>>>(s, d, nt64); });
    if (want("dynamic")) {
        report("dynamic 64 rows (32 KiB tiles), U=4", [&] { cudaMemsetAsync(ctr, 0, 4); warp_dyn<64, 4><<<148, 1024>>>(s, d, nt64, ctr); });
        report("dynamic 16 rows (8 KiB tiles), U=4", [&] { cudaMemsetAsync(ctr, 0, 4); warp_dyn<16, 4><<<148, 1024>>>(s, d, (unsigned)(nvec / (16 * 32)), ctr); });
        report("dynamic 4 rows (2 KiB tiles), U=4", [&] { cudaMemsetAsync(ctr, 0, 4); warp_dyn<4, 4><<<148, 1024>>>(s, d, (unsigned)(nvec / (4 * 32)), ctr); });
        report("dynamic 256 rows (128 KiB tiles), U=4", [&] { cudaMemsetAsync(ctr, 0, 4); warp_dyn<256, 4><<<148, 1024>>>(s, d, (unsigned)(nvec / (256 * 32)), ctr); });
        report("dynamic 64 rows, U=8", [&] { cudaMemsetAsync(ctr, 0, 4); warp_dyn<64, 8><<<148, 1024>>>(s, d, nt64, ctr); });
    }
    return 0;
}
