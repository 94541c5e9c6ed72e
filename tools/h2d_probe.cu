// h2d_probe.cu -- what can cudaMemcpyAsync H2D deliver with the pipeline's slot geometry?
// T host threads, each cycling through K pinned slots of S KiB: wait for the slot's previous copy (event), optionally
// refill the slot with a memcpy from a per-thread source buffer (stands in for the read), enqueue the copy, record.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/h2d_probe.cu -o tools/h2d_probe
//   tools/h2d_probe <threads> <copy_kib> <slots> <total_gib> <shared_stream 0|1> <refill 0|1> [wait: 0 = spin on cudaEventQuery, 1 = sleep 50us]
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

// mode 2: the copy is a kernel that pulls the slot over PCIe through its mapped address (no copy engine): every warp
// takes 32 KiB tiles, `U` rows of 512 B in flight per warp
template <int U>
__global__ void __launch_bounds__(256) pull_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n_tiles) {
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < n_tiles; t += gridDim.x * 8) {
        const uint4* s = src + (size_t)t * 2048;
        uint4* d = dst + (size_t)t * 2048;
#pragma unroll 1
        for (int r = 0; r < 64; r += U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(s + (r + u) * 32 + lane));
#pragma unroll
            for (int u = 0; u < U; u++) d[(r + u) * 32 + lane] = v[u];
        }
    }
}

// mode 5: the same pull through the TMA engine: one thread per CTA streams the slot through a ring of shared-memory
// buffers with cp.async.bulk (global/sysmem -> shared, mbarrier complete_tx) and back out with cp.async.bulk (shared -> global)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int CH, int NB>
__global__ void __launch_bounds__(32) tma_pull_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[NB];
    if (threadIdx.x != 0) return;
    for (int i = 0; i < NB; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t n_all = bytes / CH, per = (n_all + gridDim.x - 1) / gridDim.x;
    const size_t c0 = blockIdx.x * per, c1 = c0 + per < n_all ? c0 + per : n_all;
    if (c0 >= c1) return;
    const size_t n = c1 - c0;
    auto load = [&](size_t i) {
        const int b = (int)(i % NB);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)b * CH)),
                     "l"(src + (c0 + i) * CH), "r"(CH), "r"(smem_u32(&bar[b]))
                     : "memory");
    };
    for (size_t i = 0; i < n && i < NB; i++) load(i);
    for (size_t i = 0; i < n; i++) {
        const int b = (int)(i % NB);
        const uint32_t parity = (uint32_t)((i / NB) & 1);
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(parity) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (c0 + i) * CH), "r"(smem_u32(smem + (size_t)b * CH)), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (i + NB < n) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            load(i + NB);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// mode 2 with wait = 9: loads carry the L2::256B prefetch-size hint (larger fills from system memory?)
__global__ void __launch_bounds__(256) pull_kernel_256(const uint4* __restrict__ src, uint4* __restrict__ dst, int n_tiles) {
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < n_tiles; t += gridDim.x * 8) {
        const uint4* s = src + (size_t)t * 2048;
        uint4* d = dst + (size_t)t * 2048;
#pragma unroll 1
        for (int r = 0; r < 64; r += 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(s + (r + u) * 32 + lane));
#pragma unroll
            for (int u = 0; u < 4; u++) d[(r + u) * 32 + lane] = v[u];
        }
    }
}

int main(int argc, char** argv) {
    if (argc < 7) return fprintf(stderr, "usage: threads copy_kib slots total_gib shared_stream refill [wait]\n"), 2;
    const int T = atoi(argv[1]);
    const size_t S = (size_t)atoi(argv[2]) << 10;
    const int K = atoi(argv[3]);
    const double total = atof(argv[4]) * (1ull << 30);
    const int shared = atoi(argv[5]), refill = atoi(argv[6]), wait_mode = argc > 7 ? atoi(argv[7]) : 0;
    uint8_t *host, *dev;
    CK(cudaHostAlloc((void**)&host, S * K * T, cudaHostAllocPortable | cudaHostAllocMapped));
    memset(host, 1, S * K * T);
    const size_t dev_bytes = 4ull << 30;
    CK(cudaMalloc((void**)&dev, dev_bytes));
    std::vector<uint8_t> src(refill ? (size_t)T * (64 << 20) : 1, 7);
    CK(cudaFuncSetAttribute(tma_pull_kernel<16384, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(tma_pull_kernel<32768, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(tma_pull_kernel<8192, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaStream_t ss;
    CK(cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking));
    std::mutex mu;
    std::atomic<uint64_t> next{0};
    const uint64_t n_copies = (uint64_t)(total / S);
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            CK(cudaSetDevice(0));
            cudaStream_t st = ss;
            if (shared != 1) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            std::vector<cudaEvent_t> ev(K);
            std::vector<char> busy(K, 0);
            for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ready++;
            while (!go.load()) std::this_thread::yield();
            size_t srcoff = 0;
            for (int k = 0;; k = (k + 1) % K) {
                const uint64_t i = next.fetch_add(1);
                if (i >= n_copies) break;
                if (busy[k]) {
                    while (cudaEventQuery(ev[k]) == cudaErrorNotReady) {
                        if (wait_mode) std::this_thread::sleep_for(std::chrono::microseconds(50));
                    }
                }
                uint8_t* slot = host + ((size_t)t * K + k) * S;
                if (refill) {
                    memcpy(slot, src.data() + (size_t)t * (64 << 20) + srcoff, S);
                    srcoff = (srcoff + S) % ((64 << 20) - S + 1);
                }
                const size_t doff = (i * S) % (dev_bytes - S + 1) / 512 * 512;
                if (shared >= 5) {
                    uint8_t* mapped;
                    CK(cudaHostGetDevicePointer((void**)&mapped, slot, 0));
                    const int grid = shared == 5 ? (int)(S >> 17) : (shared == 6 ? (int)(S >> 16) : (int)(S >> 18));  // 128 / 64 / 256 KiB per CTA
                    if (wait_mode == 16) tma_pull_kernel<16384, 4><<<grid < 1 ? 1 : grid, 32, 65536, st>>>(mapped, dev + doff, S);
                    else if (wait_mode == 32) tma_pull_kernel<32768, 4><<<grid < 1 ? 1 : grid, 32, 131072, st>>>(mapped, dev + doff, S);
                    else tma_pull_kernel<8192, 8><<<grid < 1 ? 1 : grid, 32, 65536, st>>>(mapped, dev + doff, S);
                    CK(cudaEventRecord(ev[k], st));
                } else if (shared >= 2) {
                    const int n_tiles = (int)(S >> 15);
                    const int grid = shared == 2 ? (n_tiles + 7) / 8 : (shared == 3 ? (n_tiles + 15) / 16 : (n_tiles + 31) / 32);
                    uint8_t* mapped;
                    CK(cudaHostGetDevicePointer((void**)&mapped, slot, 0));
                    if (wait_mode == 9) pull_kernel_256<<<grid, 256, 0, st>>>((const uint4*)mapped, (uint4*)(dev + doff), n_tiles);
                    else if (wait_mode >= 8) pull_kernel<8><<<grid, 256, 0, st>>>((const uint4*)mapped, (uint4*)(dev + doff), n_tiles);
                    else pull_kernel<4><<<grid, 256, 0, st>>>((const uint4*)mapped, (uint4*)(dev + doff), n_tiles);
                    CK(cudaEventRecord(ev[k], st));
                } else if (shared) {
                    std::lock_guard<std::mutex> lk(mu);
                    CK(cudaMemcpyAsync(dev + doff, slot, S, cudaMemcpyHostToDevice, st));
                    CK(cudaEventRecord(ev[k], st));
                } else {
                    CK(cudaMemcpyAsync(dev + doff, slot, S, cudaMemcpyHostToDevice, st));
                    CK(cudaEventRecord(ev[k], st));
                }
                busy[k] = 1;
            }
            if (shared != 1) CK(cudaStreamSynchronize(st));
        });
    while (ready.load() < T) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true);
    for (auto& x : th) x.join();
    CK(cudaDeviceSynchronize());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    {   // the bytes must have arrived (the slots hold 1s, or 7s with refill)
        std::vector<uint8_t> back(1 << 20);
        CK(cudaMemcpy(back.data(), dev, back.size(), cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (uint8_t b : back) bad += b != (refill ? 7 : 1);
        if (bad) fprintf(stderr, "h2d_probe: %zu wrong bytes in the first MiB\n", bad);
    }
    printf("{\"threads\": %d, \"copy_kib\": %zu, \"slots\": %d, \"shared_stream\": %d, \"refill\": %d, \"wait\": %d, \"gib_s\": %.2f}\n", T, S >> 10, K, shared, refill, wait_mode,
           n_copies * (double)S / secs / (1ull << 30));
    return 0;
}
