// Micro-benchmark (not product code): which global access pattern gets closest to the HBM copy peak?
//   A: each warp streams its own contiguous TILE (what pack_kernel does), tiles dealt round-robin
//   B: the warps of a CTA take adjacent 512 B rows of one CTA-wide region (interleaved rows)
//   C: classic grid-stride vector copy with a huge grid (what the library copy kernels do)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint4 ld(const uint4* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st(uint4* p, uint4 v) { asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
template <int U> __device__ __forceinline__ void copy_rows(const uint4* s, uint4* d, size_t v0, size_t stride_vec, int nrows, int lane) {
    for (int r = 0; r + U <= nrows; r += U) {
        uint4 x[U];
#pragma unroll
        for (int u = 0; u < U; u++) x[u] = ld(s + v0 + (size_t)(r + u) * stride_vec + lane);
#pragma unroll
        for (int u = 0; u < U; u++) st(d + v0 + (size_t)(r + u) * stride_vec + lane, x[u]);
    }
}
template <int ROWS> __global__ void __launch_bounds__(1024, 1) pat_a(const uint4* s, uint4* d, size_t ntiles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t t = blockIdx.x + (size_t)warp * gridDim.x; t < ntiles; t += (size_t)gridDim.x * 32) copy_rows<4>(s, d, t * ROWS * 32, 32, ROWS, lane);
}
template <int ROWS> __global__ void __launch_bounds__(1024, 1) pat_b(const uint4* s, uint4* d, size_t nregions) {  // region = 32 warps x ROWS rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t t = blockIdx.x; t < nregions; t += gridDim.x) copy_rows<4>(s, d, t * ROWS * 32 * 32 + (size_t)warp * 32, 32 * 32, ROWS, lane);
}
__global__ void __launch_bounds__(256) pat_c(const uint4* s, uint4* d, size_t nvec) {
    size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x);
    const size_t base = (size_t)blockIdx.x * 256 * 4 + threadIdx.x;
    uint4 x[4];
#pragma unroll
    for (int u = 0; u < 4; u++) x[u] = base + u * 256 < nvec ? ld(s + base + u * 256) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; u++) if (base + u * 256 < nvec) st(d + base + u * 256, x[u]);
    (void)i;
}
int main() {
    const size_t bytes = 8ull << 30, nvec = bytes / 16;
    uint4 *s, *d; cudaMalloc(&s, bytes); cudaMalloc(&d, bytes); cudaMemset(s, 1, bytes); cudaMemset(d, 0, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto report = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaEventRecord(a); for (int i = 0; i < 10; i++) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
        printf("%-34s %7.3f ms  %7.1f GB/s  %5.1f%% of 6538.9\n", name, ms, 2.0 * bytes / ms / 1e6, 2.0 * bytes / ms / 1e6 / 6538.9 * 100);
    };
    report("cudaMemcpy D2D", [&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); });
    report("A per-warp 16 KiB tiles", [&] { pat_a<32><<<148, 1024>>>(s, d, nvec / (32 * 32)); });
    report("A per-warp 32 KiB tiles", [&] { pat_a<64><<<148, 1024>>>(s, d, nvec / (64 * 32)); });
    report("A per-warp 128 KiB tiles", [&] { pat_a<256><<<148, 1024>>>(s, d, nvec / (256 * 32)); });
    report("B CTA-interleaved, 4 rows/warp", [&] { pat_b<4><<<148, 1024>>>(s, d, nvec / (4 * 32 * 32)); });
    report("B CTA-interleaved, 16 rows/warp", [&] { pat_b<16><<<148, 1024>>>(s, d, nvec / (16 * 32 * 32)); });
    report("B CTA-interleaved, 64 rows/warp", [&] { pat_b<64><<<148, 1024>>>(s, d, nvec / (64 * 32 * 32)); });
    report("B x2 CTAs/SM? (grid 296)", [&] { pat_b<16><<<296, 1024>>>(s, d, nvec / (16 * 32 * 32)); });
    report("C grid-stride 256 thr, huge grid", [&] { pat_c<<<(unsigned)(nvec / 1024), 256>>>(s, d, nvec); });
    return 0;
}
