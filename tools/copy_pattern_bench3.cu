// Micro-benchmark 3 (not product code): isolate why the huge-grid copy beats persistent ones.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint4 ld(const uint4* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st(uint4* p, uint4 v) { asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__global__ void fill(uint4* p, size_t n) { for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint32_t x = (uint32_t)(i * 2654435761u); p[i] = make_uint4(x, x * 3 + 1, x ^ 0x9e3779b9u, x * 7); } }
// body: block-contiguous chunk of T*U vectors starting at chunk*T*U; thread t takes t, t+T, ...
template <int T, int U> __device__ __forceinline__ void body_block(const uint4* s, uint4* d, size_t chunk) {
    const size_t base = chunk * T * U + threadIdx.x;
    uint4 x[U];
#pragma unroll
    for (int u = 0; u < U; u++) x[u] = ld(s + base + (size_t)u * T);
#pragma unroll
    for (int u = 0; u < U; u++) st(d + base + (size_t)u * T, x[u]);
}
// body: warp-contiguous: each warp takes U consecutive 512 B rows
template <int T, int U> __device__ __forceinline__ void body_warp(const uint4* s, uint4* d, size_t chunk) {
    const size_t base = chunk * T * U + (size_t)(threadIdx.x >> 5) * 32 * U + (threadIdx.x & 31);
    uint4 x[U];
#pragma unroll
    for (int u = 0; u < U; u++) x[u] = ld(s + base + (size_t)u * 32);
#pragma unroll
    for (int u = 0; u < U; u++) st(d + base + (size_t)u * 32, x[u]);
}
template <int T, int U, bool WARP> __global__ void __launch_bounds__(T) once(const uint4* s, uint4* d) { if (WARP) body_warp<T, U>(s, d, blockIdx.x); else body_block<T, U>(s, d, blockIdx.x); }
template <int T, int U, bool WARP> __global__ void __launch_bounds__(T) persist(const uint4* s, uint4* d, size_t nchunks) {
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) { if (WARP) body_warp<T, U>(s, d, c); else body_block<T, U>(s, d, c); }
}
int main() {
    const size_t bytes = 8ull << 30, nvec = bytes / 16;
    uint4 *s, *d; cudaMalloc(&s, bytes); cudaMalloc(&d, bytes); fill<<<148 * 8, 256>>>(s, nvec); cudaMemset(d, 0, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto report = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaEventRecord(a); for (int i = 0; i < 10; i++) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
        cudaError_t e = cudaGetLastError();
        printf("%-52s %7.3f ms  %7.1f GB/s  %5.1f%%  %s\n", name, ms, 2.0 * bytes / ms / 1e6, 2.0 * bytes / ms / 1e6 / 6538.9 * 100, e ? cudaGetErrorString(e) : "");
    };
    report("cudaMemcpy D2D (random data)", [&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); });
    report("once  256x4 block-contig (pattern C)", [&] { once<256, 4, false><<<(unsigned)(nvec / 1024), 256>>>(s, d); });
    report("once  256x4 warp-contig", [&] { once<256, 4, true><<<(unsigned)(nvec / 1024), 256>>>(s, d); });
    report("once  1024x4 block-contig", [&] { once<1024, 4, false><<<(unsigned)(nvec / 4096), 1024>>>(s, d); });
    report("once  1024x4 warp-contig", [&] { once<1024, 4, true><<<(unsigned)(nvec / 4096), 1024>>>(s, d); });
    report("once  256x8 block-contig", [&] { once<256, 8, false><<<(unsigned)(nvec / 2048), 256>>>(s, d); });
    report("persist 256x4 block-contig grid 148*8", [&] { persist<256, 4, false><<<148 * 8, 256>>>(s, d, nvec / 1024); });
    report("persist 256x4 warp-contig  grid 148*8", [&] { persist<256, 4, true><<<148 * 8, 256>>>(s, d, nvec / 1024); });
    report("persist 1024x4 block-contig grid 148*2", [&] { persist<1024, 4, false><<<148 * 2, 1024>>>(s, d, nvec / 4096); });
    report("persist 1024x4 block-contig grid 148", [&] { persist<1024, 4, false><<<148, 1024>>>(s, d, nvec / 4096); });
    report("persist 1024x4 warp-contig  grid 148", [&] { persist<1024, 4, true><<<148, 1024>>>(s, d, nvec / 4096); });
    report("persist 256x4 block-contig grid 148*32", [&] { persist<256, 4, false><<<148 * 32, 256>>>(s, d, nvec / 1024); });
    return 0;
}
