// misc_kernels.cu -- xxh64_chunks and verify_pattern kernels for sm_100a.
//
// xxh64_chunks: XXH64 is a sequential recurrence per accumulator, so a chunk offers only 4-way
// parallelism (the four 8-byte lanes of each 32-byte stripe).  The kernel therefore runs 4
// threads per chunk and parallelises ACROSS chunks; it is HBM-bound only when a launch holds
// >= ~75k chunks and is latency-bound otherwise, which is why crc32c is the default proof.
// Algorithmic bytes: sum(len) (read only).
//
// verify_pattern: recomputes the synthetic content byte(file_id, off) (generator
// restated in tests, never linked from oracle/) and counts mismatching bytes.  Test aid.
#include "common.hpp"
#include "lsio_cuda.h"

#include <atomic>
#include <cstdlib>
#include <vector>

namespace lsio {
namespace {

constexpr uint64_t XP1 = 0x9E3779B185EBCA87ull, XP2 = 0xC2B2AE3D27D4EB4Full, XP3 = 0x165667B19E3779F9ull,
                   XP4 = 0x85EBCA77C2B2AE63ull, XP5 = 0x27D4EB2F165667C5ull;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xround(uint64_t acc, uint64_t in) { return rotl64(acc + in * XP2, 31) * XP1; }
__device__ __forceinline__ uint64_t xmerge(uint64_t acc, uint64_t v) { return (acc ^ xround(0, v)) * XP1 + XP4; }

// 8 bytes at an arbitrary address, via aligned 8-byte loads of the aligned superset.
// Streaming load that bypasses L1: a warp walks 8 chunks at once, and chunks that sit a power of two apart
// (64 KiB: measured 4.5 % of HBM peak with L1-allocating loads) thrash the same L1 sets; L2 serves 32-byte
// sectors, exactly one per 4-lane group and stripe.
__device__ __forceinline__ uint64_t ld64_stream(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t ld64_any(uint64_t addr) {
    const uint64_t* q = (const uint64_t*)(addr & ~7ull);
    const int s = (int)(addr & 7) * 8;
    const uint64_t lo = ld64_stream(q);
    if (s == 0) return lo;
    const uint64_t hi = ld64_stream(q + 1);
    return (lo >> s) | (hi << (64 - s));
}
__device__ __forceinline__ uint32_t ld8(uint64_t addr) { return *(const uint8_t*)addr; }

__global__ void __launch_bounds__(256) xxh64_kernel(const uint64_t* __restrict__ ptrs, const uint64_t* __restrict__ lens,
                                                    uint64_t n, uint64_t* __restrict__ out) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t chunk = gid >> 2;
    const int j = (int)(gid & 3);
    const bool live = chunk < n;
    const uint64_t p = live ? ptrs[chunk] : 0, len = live ? lens[chunk] : 0;
    const uint64_t seed = 0;
    uint64_t v = j == 0 ? seed + XP1 + XP2 : j == 1 ? seed + XP2 : j == 2 ? seed : seed - XP1;
    const uint64_t nstripes = len >> 5;
    uint64_t a = p + 8 * j;
    uint64_t i = 0;
    for (; i + 8 <= nstripes; i += 8, a += 256) {  // eight loads in flight per thread (16: worse, registers)
        uint64_t w[8];
#pragma unroll
        for (int u = 0; u < 8; u++) w[u] = ld64_any(a + 32 * u);
#pragma unroll
        for (int u = 0; u < 8; u++) v = xround(v, w[u]);
    }
    for (; i < nstripes; i++, a += 32) v = xround(v, ld64_any(a));
    // gather the four accumulators of this chunk into its lane 0 (groups of 4 lanes)
    const unsigned full = 0xFFFFFFFFu;
    const int base = (threadIdx.x & 31) & ~3;
    const uint64_t v1 = __shfl_sync(full, v, base), v2 = __shfl_sync(full, v, base + 1), v3 = __shfl_sync(full, v, base + 2),
                   v4 = __shfl_sync(full, v, base + 3);
    if (!live || j != 0) return;
    uint64_t h;
    if (len >= 32) {
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else {
        h = seed + XP5;
    }
    h += len;
    uint64_t q = p + (nstripes << 5);
    const uint64_t end = p + len;
    for (; q + 8 <= end; q += 8) {
        uint64_t w = 0;
        for (int b = 0; b < 8; b++) w |= (uint64_t)ld8(q + b) << (8 * b);
        h ^= xround(0, w);
        h = rotl64(h, 27) * XP1 + XP4;
    }
    if (q + 4 <= end) {
        uint32_t w = 0;
        for (int b = 0; b < 4; b++) w |= ld8(q + b) << (8 * b);
        h ^= (uint64_t)w * XP1;
        h = rotl64(h, 23) * XP2 + XP3;
        q += 4;
    }
    for (; q < end; q++) {
        h ^= (uint64_t)ld8(q) * XP5;
        h = rotl64(h, 11) * XP1;
    }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    out[chunk] = h;
}

// ---- xxh64, staged through shared memory by the bulk-copy (TMA) engine ------------------------------------------------------
// Same arithmetic, different data movement: an A/B that LOST (numbers at xxh64_launch below) and is kept opt-in.  The plain kernel above
// walks thousands of chunks at once with 8-byte loads, 256 B per chunk per round: DRAM sees short bursts scattered over as many pages
// (still 81-88 % of the HBM peak with >= 131k chunks, kernel-only).  Here every team of four
// lanes owns two 1 KiB buffers in shared memory and its leader streams the chunk into them with cp.async.bulk (one contiguous
// 1 KiB + 16 B burst per copy, completion on an mbarrier), while the four lanes run the accumulators out of shared memory.
// A chunk may start at any address: staging starts at the 16-byte boundary below it (`lead` bytes early) and every piece carries
// 16 extra bytes, so that each of a piece's 32 stripes is whole inside its buffer; lanes read their 8-byte words with two aligned
// LDS and a funnel shift when the chunk is not 8-byte aligned.
constexpr int kXPiece = 1024;                    // bytes of stripes per bulk copy (32 stripes)
constexpr int kXBuf = kXPiece + 32;              // + 16 carried bytes + 16 of slack for the second aligned word of the last lane
constexpr int kXTeamBytes = 2 * kXBuf + 96;      // 2208 B = 552 words = 8 (mod 32): eight teams of a warp tile the banks exactly twice
constexpr int kXTeams = 32;                      // chunks per CTA (128 threads)
constexpr size_t kXSmem = (size_t)kXTeams * kXTeamBytes + 2 * kXTeams * sizeof(uint64_t);

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(4 * kXTeams, 3) xxh64_tma_kernel(const uint64_t* __restrict__ ptrs, const uint64_t* __restrict__ lens, uint64_t n,
                                                                  uint64_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t xsm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(xsm + (size_t)kXTeams * kXTeamBytes);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * kXTeams; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int team = threadIdx.x >> 2, j = threadIdx.x & 3;
    const uint64_t chunk = (uint64_t)blockIdx.x * kXTeams + team;
    const bool live = chunk < n;
    const uint64_t p = live ? ptrs[chunk] : 0, len = live ? lens[chunk] : 0;
    const uint64_t seed = 0;
    uint64_t v = j == 0 ? seed + XP1 + XP2 : j == 1 ? seed + XP2 : j == 2 ? seed : seed - XP1;
    const uint64_t nstripes = len >> 5;
    const uint32_t lead = (uint32_t)(p & 15);
    const uint8_t* abase = reinterpret_cast<const uint8_t*>(p - lead);
    const uint64_t t16 = ((uint64_t)lead + (nstripes << 5) + 15) & ~15ull;  // staged bytes, from abase
    const uint64_t n_pieces = (nstripes + kXPiece / 32 - 1) / (kXPiece / 32);
    uint8_t* const buf0 = xsm + (size_t)team * kXTeamBytes;
    const uint32_t bar0 = smem_addr(&bars[2 * team]);
    const unsigned tmask = 0xFu << ((threadIdx.x & 31) & ~3);
    auto piece_bytes = [&](uint64_t i) { return (uint32_t)((t16 - i * kXPiece) < (uint64_t)(kXPiece + 16) ? (t16 - i * kXPiece) : (uint64_t)(kXPiece + 16)); };
    if (j == 0)
        for (uint64_t i = 0; i < n_pieces && i < 2; i++) bulk_load(smem_addr(buf0 + i * kXBuf), abase + i * kXPiece, piece_bytes(i), bar0 + 8 * (uint32_t)i);
    const uint32_t sh = (lead & 7) * 8;
    for (uint64_t i = 0; i < n_pieces; i++) {
        const uint32_t b = (uint32_t)(i & 1);
        bar_wait(bar0 + 8 * b, (uint32_t)((i >> 1) & 1));
        const uint8_t* buf = buf0 + b * kXBuf;
        const uint32_t cnt = (uint32_t)((nstripes - i * (kXPiece / 32)) < (uint64_t)(kXPiece / 32) ? (nstripes - i * (kXPiece / 32)) : (uint64_t)(kXPiece / 32));
        // this lane's word of stripe s sits at lead + 32 s + 8 j: two aligned 8-byte words and a funnel shift when lead is not a multiple of 8
        const uint64_t* q = reinterpret_cast<const uint64_t*>(buf + ((lead + 8 * j) & ~7u));
        if (sh == 0) {
#pragma unroll 8
            for (uint32_t s = 0; s < cnt; s++) v = xround(v, q[4 * s]);
        } else {
#pragma unroll 8
            for (uint32_t s = 0; s < cnt; s++) v = xround(v, (q[4 * s] >> sh) | (q[4 * s + 1] << (64 - sh)));
        }
        __syncwarp(tmask);  // all four lanes are done with this buffer
        if (j == 0 && i + 2 < n_pieces) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(smem_addr(buf0 + b * kXBuf), abase + (i + 2) * kXPiece, piece_bytes(i + 2), bar0 + 8 * b);
        }
    }
    // gather the four accumulators of this chunk into its lane 0 (groups of 4 lanes) and finish as the plain kernel does
    const unsigned full = 0xFFFFFFFFu;
    const int base = (threadIdx.x & 31) & ~3;
    const uint64_t v1 = __shfl_sync(full, v, base), v2 = __shfl_sync(full, v, base + 1), v3 = __shfl_sync(full, v, base + 2),
                   v4 = __shfl_sync(full, v, base + 3);
    if (!live || j != 0) return;
    uint64_t h;
    if (len >= 32) {
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else {
        h = seed + XP5;
    }
    h += len;
    uint64_t qa = p + (nstripes << 5);
    const uint64_t end = p + len;
    for (; qa + 8 <= end; qa += 8) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w |= (uint64_t)ld8(qa + k) << (8 * k);
        h ^= xround(0, w);
        h = rotl64(h, 27) * XP1 + XP4;
    }
    if (qa + 4 <= end) {
        uint32_t w = 0;
        for (int k = 0; k < 4; k++) w |= ld8(qa + k) << (8 * k);
        h ^= (uint64_t)w * XP1;
        h = rotl64(h, 23) * XP2 + XP3;
        qa += 4;
    }
    for (; qa < end; qa++) {
        h ^= (uint64_t)ld8(qa) * XP5;
        h = rotl64(h, 11) * XP1;
    }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    out[chunk] = h;
}

// ---- verify_pattern ------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t pattern_word(uint64_t file_id, uint64_t word_idx) {
    return splitmix64(0x4C53494Full ^ (file_id * 0x9E3779B97F4A7C15ull) ^ word_idx);
}
struct Piece {  // <= 256 KiB of one span
    uint64_t ptr, len, file_id, file_off;
};
__global__ void __launch_bounds__(256) verify_kernel(const Piece* __restrict__ pieces, uint32_t npieces,
                                                     unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    for (uint32_t pi = blockIdx.x; pi < npieces; pi += gridDim.x) {
        const Piece pc = pieces[pi];
        // iterate over the file's 8-byte words that intersect [file_off, file_off + len)
        const uint64_t w0 = pc.file_off >> 3, w1 = (pc.file_off + pc.len + 7) >> 3;
        for (uint64_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
            const uint64_t expect = pattern_word(pc.file_id, w);
            const int64_t rel = (int64_t)(w << 3) - (int64_t)pc.file_off;  // may be negative for the first word
            const uint64_t addr = pc.ptr + (uint64_t)rel;
            if (rel >= 0 && (uint64_t)rel + 8 <= pc.len && (addr & 7) == 0) {
                uint64_t x = *(const uint64_t*)addr ^ expect;
                // count non-zero bytes
                x |= x >> 4; x |= x >> 2; x |= x >> 1;
                bad += __popcll(x & 0x0101010101010101ull);
            } else {
                for (int b = 0; b < 8; b++) {
                    const int64_t r = rel + b;
                    if (r >= 0 && (uint64_t)r < pc.len) bad += ld8(pc.ptr + (uint64_t)r) != (uint8_t)(expect >> (8 * b));
                }
            }
        }
    }
    // warp then block reduce, one atomic per block
    for (int o = 16; o; o >>= 1) bad += __shfl_down_sync(0xFFFFFFFFu, bad, o);
    __shared__ unsigned long long sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = bad;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += sh[i];
        if (t) atomicAdd(mismatches, t);
    }
}

__global__ void __launch_bounds__(256) fill_kernel(uint8_t* __restrict__ dst, uint64_t len, uint64_t file_id, uint64_t file_off) {
    // one 8-byte file word per thread-iteration
    const uint64_t w0 = file_off >> 3, w1 = (file_off + len + 7) >> 3;
    for (uint64_t w = w0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < w1; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = pattern_word(file_id, w);
        const int64_t rel = (int64_t)(w << 3) - (int64_t)file_off;
        if (rel >= 0 && (uint64_t)rel + 8 <= len && (((uint64_t)dst + (uint64_t)rel) & 7) == 0) {
            *(uint64_t*)(dst + rel) = v;
        } else {
            for (int b = 0; b < 8; b++) {
                const int64_t r = rel + b;
                if (r >= 0 && (uint64_t)r < len) dst[r] = (uint8_t)(v >> (8 * b));
            }
        }
    }
}

}  // namespace
}  // namespace lsio

namespace lsio {
// Asynchronous launch for the pipeline: ptrs/lens/out are device-accessible (mapped pinned) arrays.
// Variant.  Kernel-only durations (ncu, one launch each, profiles/r02_xxh64_kernels.txt): the PLAIN kernel wins everywhere --
// 524,288 x 16 KiB: 1497 us (5.74 TB/s, 88 % of the measured HBM peak) vs 2203 us staged; 131,072 x 16 KiB: 405 vs 598 us; 131,072 x 64 KiB:
// 1573 vs 2435 us; 32,768 x 256 KiB: 2359 vs 3057 us.  The staged kernel runs 12 warps per SM (70 KiB of shared memory per CTA) and its
// 1 KiB bulk copies do not make up for that; the plain kernel's 64 warps per SM hide the scattered 256-byte bursts well enough.  The
// staged kernel therefore stays an opt-in A/B (variant 2 through lsio_cuda_xxh64_variant or LSIO_XXH64_VARIANT=2), never the default.
static void xxh64_launch(const uint64_t* ptrs_dev, const uint64_t* lens_dev, size_t n, uint64_t* out_dev, cudaStream_t st, int variant) {
    static const int forced = [] { const char* e = getenv("LSIO_XXH64_VARIANT"); return e ? atoi(e) : 0; }();
    static std::atomic<int> ready[64] = {};  // per device: 0 unknown, 1 attribute set, -1 refused
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (ready[dev].load(std::memory_order_acquire) == 0)
        ready[dev].store(cudaFuncSetAttribute(xxh64_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXSmem) == cudaSuccess ? 1 : -1, std::memory_order_release);
    if (!variant) variant = forced;
    const bool staged = ready[dev].load(std::memory_order_acquire) == 1 && variant == 2;
    if (staged) xxh64_tma_kernel<<<(unsigned)((n + kXTeams - 1) / kXTeams), 4 * kXTeams, kXSmem, st>>>(ptrs_dev, lens_dev, n, out_dev);
    else xxh64_kernel<<<(unsigned)((4 * (uint64_t)n + 255) / 256), 256, 0, st>>>(ptrs_dev, lens_dev, n, out_dev);
}
int xxh64_launch_async(const uint64_t* ptrs_dev, const uint64_t* lens_dev, size_t n, uint64_t* out_dev, cudaStream_t st) {
    if (!n) return 0;
    xxh64_launch(ptrs_dev, lens_dev, n, out_dev, st, 0);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    LSIO_CUDA_TRY(cudaGetLastError());
    return 0;
}
}  // namespace lsio

extern "C" {

int lsio_cuda_fill_pattern(int device, void* dst_dev, uint64_t len, uint64_t file_id, uint64_t file_off, void* stream) {
    using namespace lsio;
    if (!dst_dev && len) return -EINVAL;
    if (!len) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    const uint64_t words = (len + 15) / 8;
    const unsigned grid = (unsigned)std::min<uint64_t>((words + 255) / 256, 148 * 16);
    fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint8_t*)dst_dev, len, file_id, file_off);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    LSIO_CUDA_TRY(cudaGetLastError());
    return 0;
}

int lsio_cuda_xxh64_variant(int device, const uint64_t* ptrs, const uint64_t* lens, size_t n, uint64_t* out_dev, void* stream, int variant) {
    using namespace lsio;
    if ((!ptrs || !lens || !out_dev) && n) return -EINVAL;
    if (n == 0) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t* d = nullptr;
    LSIO_CUDA_TRY(cudaMallocAsync(&d, 2 * n * sizeof(uint64_t), st));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d, ptrs, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d + n, lens, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    xxh64_launch(d, d + n, n, out_dev, st, variant);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    LSIO_CUDA_TRY(cudaGetLastError());
    LSIO_CUDA_TRY(cudaFreeAsync(d, st));
    LSIO_CUDA_TRY(cudaStreamSynchronize(st));  // ptrs/lens are pageable host memory: keep the call self-contained
    return 0;
}

int lsio_cuda_xxh64(int device, const uint64_t* ptrs, const uint64_t* lens, size_t n, uint64_t* out_dev, void* stream) {
    return lsio_cuda_xxh64_variant(device, ptrs, lens, n, out_dev, stream, 0);
}

int lsio_cuda_verify_pattern(int device, const uint64_t* ptrs, const uint64_t* lens, const uint64_t* file_ids,
                             const uint64_t* file_offs, size_t n, uint64_t* mismatch_dev, void* stream) {
    using namespace lsio;
    if ((!ptrs || !lens || !file_ids || !file_offs) && n) return -EINVAL;
    if (!mismatch_dev) return -EINVAL;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    cudaStream_t st = (cudaStream_t)stream;
    constexpr uint64_t kPiece = 256 << 10;
    std::vector<Piece> pieces;
    for (size_t i = 0; i < n; i++)
        for (uint64_t o = 0; o < lens[i]; o += kPiece)
            pieces.push_back(Piece{ptrs[i] + o, std::min<uint64_t>(kPiece, lens[i] - o), file_ids[i], file_offs[i] + o});
    LSIO_CUDA_TRY(cudaMemsetAsync(mismatch_dev, 0, sizeof(uint64_t), st));
    if (pieces.empty()) return (int)cudaStreamSynchronize(st);
    Piece* d = nullptr;
    LSIO_CUDA_TRY(cudaMallocAsync(&d, pieces.size() * sizeof(Piece), st));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d, pieces.data(), pieces.size() * sizeof(Piece), cudaMemcpyHostToDevice, st));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const unsigned grid = (unsigned)std::min<size_t>(pieces.size(), (size_t)sms * 8);
    verify_kernel<<<grid, 256, 0, st>>>(d, (uint32_t)pieces.size(), (unsigned long long*)mismatch_dev);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    LSIO_CUDA_TRY(cudaGetLastError());
    LSIO_CUDA_TRY(cudaFreeAsync(d, st));
    LSIO_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
