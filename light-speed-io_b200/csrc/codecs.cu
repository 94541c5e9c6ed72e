// codecs.cu -- sm_100a kernels of the codec layer (include/lsio_codecs.h): chunks landed in HBM by the
// I/O path are decoded in HBM.  The reference plans this layer (planned_design.md:83-90 `lsio_codecs`,
// README.md:68-70) and holds no code for it; formats follow the published LZ4 block format and
// numcodecs.Shuffle.  Product code: no oracle, no CPU fallback.
//
//   lz4_kernel        one warp per block.  LZ4 is a serial token stream, so the parallelism is (a) across
//                     blocks and (b) inside each copy: every lane parses the same tokens from a shared-memory
//                     window of the compressed input (refilled with coalesced 16-byte loads), literals and
//                     matches are copied by the 32 lanes.  A match that overlaps its own output
//                     (offset < length) is periodic, so byte i is out[op - offset + i % offset]: every byte
//                     of a match is independent of the others and only depends on bytes before `op`.
//   shuffle_kernel    byte (un)shuffle as a tiled gather: each thread assembles one 16-byte output vector from
//                     16/E bytes of each of the E byte planes (register transpose, PRMT), warp accesses are
//                     contiguous in every plane and in the output.  HBM-bound: 2 bytes moved per byte.
#include <algorithm>
#include <vector>

#include "common.hpp"
#include "lsio_codecs.h"

namespace lsio {
namespace {

// ---------------------------------------------------------------------------------------------------------
// LZ4 block decode
// ---------------------------------------------------------------------------------------------------------
constexpr int kLzWarps = 8;     // blocks decoded per CTA
constexpr int kLzWin = 1024;    // bytes of compressed input staged per warp

struct LzIn {
    const uint8_t* src;  // compressed bytes
    uint32_t n;          // their count
    uint32_t lead;       // src - (src rounded down to 16): window coordinates are relative to the aligned base
    uint8_t* win;        // this warp's shared-memory window
    uint32_t ws;         // window start (multiple of 16, aligned-base coordinates)
    int lane;
};

// Stage [ws, ws + kLzWin) of the aligned-base coordinates: 64 vectors, two per lane.
__device__ __forceinline__ void lz_fill(LzIn& in) {
    const uint8_t* abase = in.src - in.lead;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kLzWin / 16 / 32; r++) {
        const uint32_t voff = in.ws + 16u * (uint32_t)(in.lane + 32 * r);  // aligned-base coordinate of this vector
        uint4 v = make_uint4(0, 0, 0, 0);
        if (voff >= in.lead && voff + 16 <= in.lead + in.n) {
            v = *reinterpret_cast<const uint4*>(abase + voff);
        } else if (voff + 16 > in.lead && voff < in.lead + in.n) {  // edge vector: only bytes inside [src, src+n)
            uint8_t b[16];
#pragma unroll
            for (int k = 0; k < 16; k++) b[k] = (voff + k >= in.lead && voff + k < in.lead + in.n) ? abase[voff + k] : (uint8_t)0;
            v.x = b[0] | b[1] << 8 | b[2] << 16 | (uint32_t)b[3] << 24;
            v.y = b[4] | b[5] << 8 | b[6] << 16 | (uint32_t)b[7] << 24;
            v.z = b[8] | b[9] << 8 | b[10] << 16 | (uint32_t)b[11] << 24;
            v.w = b[12] | b[13] << 8 | b[14] << 16 | (uint32_t)b[15] << 24;
        }
        *reinterpret_cast<uint4*>(in.win + 16 * (in.lane + 32 * r)) = v;
    }
    __syncwarp();
}
// Byte `ip` of the compressed stream (ip < n, uniform across the warp).
__device__ __forceinline__ uint32_t lz_byte(LzIn& in, uint32_t ip) {
    const uint32_t p = ip + in.lead;
    if (p - in.ws >= (uint32_t)kLzWin) {  // also true when p < ws (never happens: ip only grows)
        in.ws = p & ~15u;
        lz_fill(in);
    }
    return in.win[p - in.ws];
}

__global__ void __launch_bounds__(kLzWarps * 32) lz4_kernel(const lsio_codec_item* __restrict__ items, uint32_t n_items,
                                                            lsio_codec_result* __restrict__ results) {
    __shared__ __align__(16) uint8_t s_win[kLzWarps][kLzWin];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t idx = blockIdx.x * kLzWarps + warp;
    if (idx >= n_items) return;
    const lsio_codec_item it = items[idx];
    LzIn in;
    in.src = reinterpret_cast<const uint8_t*>(it.src);
    in.n = it.src_len;
    in.lead = (uint32_t)(it.src & 15);
    in.win = s_win[warp];
    in.ws = 0;
    in.lane = lane;
    uint8_t* dst = reinterpret_cast<uint8_t*>(it.dst);
    const uint32_t n = it.src_len, cap = it.dst_cap;
    uint32_t ip = 0, op = 0;
    int status = LSIO_CODEC_OK;
    if (n == 0) status = LSIO_CODEC_ERR_INPUT;
    else lz_fill(in);
    while (status == LSIO_CODEC_OK) {
        if (ip >= n) { status = LSIO_CODEC_ERR_INPUT; break; }
        const uint32_t token = lz_byte(in, ip++);
        uint32_t lit = token >> 4;
        if (lit == 15) {
            uint32_t b;
            do {
                if (ip >= n) { status = LSIO_CODEC_ERR_INPUT; break; }
                b = lz_byte(in, ip++);
                lit += b;
            } while (b == 255);
            if (status) break;
        }
        if (lit > n - ip) { status = LSIO_CODEC_ERR_INPUT; break; }
        if (lit > cap - op) { status = LSIO_CODEC_ERR_OUTPUT; break; }
        if (lit) {
            const uint32_t p = ip + in.lead;
            if (p >= in.ws && p + lit <= in.ws + (uint32_t)kLzWin) {  // short runs: straight from the window
                const uint8_t* w = in.win + (p - in.ws);
                for (uint32_t i = lane; i < lit; i += 32) dst[op + i] = w[i];
            } else {
                const uint8_t* s = in.src + ip;
                for (uint32_t i = lane; i < lit; i += 32) dst[op + i] = s[i];
            }
            ip += lit;
            op += lit;
        }
        if (ip == n) break;  // the last sequence has literals only
        if (n - ip < 2) { status = LSIO_CODEC_ERR_INPUT; break; }
        uint32_t offset = lz_byte(in, ip);
        offset |= lz_byte(in, ip + 1) << 8;
        ip += 2;
        uint32_t ml = token & 15;
        if (ml == 15) {
            uint32_t b;
            do {
                if (ip >= n) { status = LSIO_CODEC_ERR_INPUT; break; }
                b = lz_byte(in, ip++);
                ml += b;
            } while (b == 255);
            if (status) break;
        }
        ml += 4;
        if (offset == 0 || offset > op) { status = LSIO_CODEC_ERR_OFFSET; break; }
        if (ml > cap - op) { status = LSIO_CODEC_ERR_OUTPUT; break; }
        __syncwarp();  // the literals above (and everything before them) are visible to every lane's match reads
        const uint8_t* m = dst + op - offset;
        if (offset >= ml) {
            for (uint32_t i = lane; i < ml; i += 32) dst[op + i] = m[i];
        } else {  // overlapping match = periodic extension of the last `offset` bytes
            for (uint32_t i = lane; i < ml; i += 32) dst[op + i] = m[i % offset];
        }
        op += ml;
        __syncwarp();
    }
    if (lane == 0) {
        results[idx].out_len = op;
        results[idx].status = status;
    }
}

// ---------------------------------------------------------------------------------------------------------
// byte shuffle / unshuffle
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kShufTile = 16384;  // bytes of one buffer handled by one CTA
constexpr int kShufThreads = 256;

template <int E> struct PlaneT;
template <> struct PlaneT<2> { using T = uint64_t; };
template <> struct PlaneT<4> { using T = uint32_t; };
template <> struct PlaneT<8> { using T = uint16_t; };
template <> struct PlaneT<16> { using T = uint8_t; };

// Output vector `e0 / V` of an unshuffle: elements [e0, e0 + V), V = 16 / E, gathered from the E planes.
template <int E>
__device__ __forceinline__ uint4 unshuffle_vec(const uint8_t* __restrict__ src, uint32_t cnt, uint32_t e0) {
    using T = typename PlaneT<E>::T;
    uint64_t pl[E];
#pragma unroll
    for (int j = 0; j < E; j++) pl[j] = (uint64_t)__ldg(reinterpret_cast<const T*>(src + (size_t)j * cnt + e0));
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int i = k / E, j = k % E;  // output byte k = byte j of element e0 + i = byte i of plane j's word
        w[k / 4] |= ((uint32_t)(pl[j] >> (8 * i)) & 255u) << (8 * (k % 4));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
// Input vector of a shuffle scattered to the E planes.
template <int E>
__device__ __forceinline__ void shuffle_vec(uint4 v, uint8_t* __restrict__ dst, uint32_t cnt, uint32_t e0) {
    using T = typename PlaneT<E>::T;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint64_t pl[E];
#pragma unroll
    for (int j = 0; j < E; j++) pl[j] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int i = k / E, j = k % E;
        pl[j] |= (uint64_t)((w[k / 4] >> (8 * (k % 4))) & 255u) << (8 * i);
    }
#pragma unroll
    for (int j = 0; j < E; j++) *reinterpret_cast<T*>(dst + (size_t)j * cnt + e0) = (T)pl[j];
}

template <int E, bool kForward>
__device__ __forceinline__ void shuffle_tile_fast(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t cnt, uint32_t v0, uint32_t v1) {
    constexpr int V = 16 / E;
    constexpr int U = 4;  // vectors in flight per thread
    uint32_t v = v0 + threadIdx.x;
    for (; v + (U - 1) * kShufThreads < v1; v += U * kShufThreads) {
        uint4 r[U];
        if (kForward) {
#pragma unroll
            for (int u = 0; u < U; u++) r[u] = __ldg(reinterpret_cast<const uint4*>(src) + v + u * kShufThreads);
#pragma unroll
            for (int u = 0; u < U; u++) shuffle_vec<E>(r[u], dst, cnt, (v + u * kShufThreads) * V);
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) r[u] = unshuffle_vec<E>(src, cnt, (v + u * kShufThreads) * V);
#pragma unroll
            for (int u = 0; u < U; u++) reinterpret_cast<uint4*>(dst)[v + u * kShufThreads] = r[u];
        }
    }
    for (; v < v1; v += kShufThreads) {
        if (kForward) shuffle_vec<E>(__ldg(reinterpret_cast<const uint4*>(src) + v), dst, cnt, v * V);
        else reinterpret_cast<uint4*>(dst)[v] = unshuffle_vec<E>(src, cnt, v * V);
    }
}

// Bytes [b0, b1) of the contiguous side (the output of an unshuffle, the input of a shuffle), one at a time.
template <bool kForward>
__device__ __forceinline__ void shuffle_bytes(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t len, uint32_t elem, uint32_t cnt,
                                              uint32_t b0, uint32_t b1) {
    const uint32_t body = cnt * elem;
    for (uint32_t k = b0 + threadIdx.x; k < b1; k += kShufThreads) {
        const uint32_t planar = k < body ? (k % elem) * cnt + k / elem : k;  // the len % elem trailing bytes stay put
        if (kForward) dst[planar] = src[k];
        else dst[k] = src[planar];
    }
}

template <bool kForward>
__global__ void __launch_bounds__(kShufThreads) shuffle_kernel(const lsio_codec_item* __restrict__ items, const uint32_t* __restrict__ tile_prefix,
                                                               uint32_t n_items, uint32_t elem) {
    __shared__ uint32_t s_item;
    if (threadIdx.x == 0) {  // the item whose tile range holds blockIdx.x
        uint32_t lo = 0, hi = n_items;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (tile_prefix[mid] <= blockIdx.x) lo = mid;
            else hi = mid;
        }
        s_item = lo;
    }
    __syncthreads();
    const uint32_t item = s_item;
    const lsio_codec_item it = items[item];
    const uint8_t* src = reinterpret_cast<const uint8_t*>(it.src);
    uint8_t* dst = reinterpret_cast<uint8_t*>(it.dst);
    const uint32_t len = it.src_len, cnt = elem ? len / elem : 0;
    const uint32_t tile = blockIdx.x - tile_prefix[item];
    const uint32_t b0 = tile * kShufTile, b1 = min(len, b0 + kShufTile);
    // vector path: whole 16-byte vectors of the contiguous side that lie inside the cnt*elem body, when every
    // plane access is naturally aligned
    const uint32_t V = (elem == 2 || elem == 4 || elem == 8 || elem == 16) ? 16 / elem : 0;
    const uint64_t contig = kForward ? it.src : it.dst, planar = kForward ? it.dst : it.src;
    uint32_t vec_end = b0;  // bytes [b0, vec_end) go through the vector path
    if (V && (contig & 15) == 0 && (planar % V) == 0 && (cnt % V) == 0) {
        vec_end = min(b1, (cnt * elem) & ~15u);
        if (vec_end < b0) vec_end = b0;
        const uint32_t v0 = b0 / 16, v1 = vec_end / 16;
        switch (elem) {
            case 2: shuffle_tile_fast<2, kForward>(src, dst, cnt, v0, v1); break;
            case 4: shuffle_tile_fast<4, kForward>(src, dst, cnt, v0, v1); break;
            case 8: shuffle_tile_fast<8, kForward>(src, dst, cnt, v0, v1); break;
            default: shuffle_tile_fast<16, kForward>(src, dst, cnt, v0, v1); break;
        }
    }
    if (vec_end < b1) shuffle_bytes<kForward>(src, dst, len, elem, cnt, vec_end, b1);
}

int check_device(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount (the codec layer has no CPU fallback)", __FILE__, __LINE__);
    if (device < 0 || device >= ndev) return -ENODEV;
    return 0;
}

}  // namespace
}  // namespace lsio

using namespace lsio;

extern "C" {

int lsio_codecs_lz4_decompress_async(int device, const lsio_codec_item* items_dev, size_t n, lsio_codec_result* results_dev, void* stream) {
    if ((!items_dev || !results_dev) && n) return -EINVAL;
    if (n > 0xFFFFFFF0u) return -E2BIG;
    if (!n) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    lz4_kernel<<<(unsigned)((n + kLzWarps - 1) / kLzWarps), kLzWarps * 32, 0, (cudaStream_t)stream>>>(items_dev, (uint32_t)n, results_dev);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    LSIO_CUDA_TRY(cudaGetLastError());
    return 0;
}

int lsio_codecs_lz4_decompress(int device, const lsio_codec_item* items, size_t n, lsio_codec_result* results, void* stream) {
    if ((!items || !results) && n) return -EINVAL;
    if (int rc = check_device(device)) return rc;
    if (!n) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    cudaStream_t st = (cudaStream_t)stream;
    char* d = nullptr;
    LSIO_CUDA_TRY(cudaMallocAsync((void**)&d, n * (sizeof(lsio_codec_item) + sizeof(lsio_codec_result)), st));
    lsio_codec_item* d_items = reinterpret_cast<lsio_codec_item*>(d);
    lsio_codec_result* d_res = reinterpret_cast<lsio_codec_result*>(d + n * sizeof(lsio_codec_item));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d_items, items, n * sizeof(lsio_codec_item), cudaMemcpyHostToDevice, st));
    int rc = lsio_codecs_lz4_decompress_async(device, d_items, n, d_res, stream);
    if (rc) return rc;
    LSIO_CUDA_TRY(cudaMemcpyAsync(results, d_res, n * sizeof(lsio_codec_result), cudaMemcpyDeviceToHost, st));
    LSIO_CUDA_TRY(cudaFreeAsync(d, st));
    LSIO_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

uint32_t lsio_codecs_shuffle_tiles(uint32_t src_len) { return (uint32_t)(((uint64_t)src_len + kShufTile - 1) / kShufTile); }

int lsio_codecs_unshuffle_async(int device, const lsio_codec_item* items_dev, const uint32_t* tile_prefix_dev, size_t n, uint32_t n_tiles,
                                uint32_t elem_size, int forward, void* stream) {
    if ((!items_dev || !tile_prefix_dev) && n) return -EINVAL;
    if (elem_size == 0 || n > 0xFFFFFFF0u) return -EINVAL;
    if (!n || !n_tiles) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    if (forward) shuffle_kernel<true><<<n_tiles, kShufThreads, 0, (cudaStream_t)stream>>>(items_dev, tile_prefix_dev, (uint32_t)n, elem_size);
    else shuffle_kernel<false><<<n_tiles, kShufThreads, 0, (cudaStream_t)stream>>>(items_dev, tile_prefix_dev, (uint32_t)n, elem_size);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    LSIO_CUDA_TRY(cudaGetLastError());
    return 0;
}

static int shuffle_host_items(int device, const lsio_codec_item* items, size_t n, uint32_t elem_size, int forward, void* stream) {
    if (!items && n) return -EINVAL;
    if (elem_size == 0) return -EINVAL;
    if (int rc = check_device(device)) return rc;
    if (!n) return 0;
    std::vector<uint32_t> prefix(n + 1, 0);
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        prefix[i] = (uint32_t)total;
        total += lsio_codecs_shuffle_tiles(items[i].src_len);
        if (total > 0x7FFFFFFFu) return -E2BIG;
    }
    prefix[n] = (uint32_t)total;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    cudaStream_t st = (cudaStream_t)stream;
    char* d = nullptr;
    LSIO_CUDA_TRY(cudaMallocAsync((void**)&d, n * sizeof(lsio_codec_item) + (n + 1) * sizeof(uint32_t), st));
    lsio_codec_item* d_items = reinterpret_cast<lsio_codec_item*>(d);
    uint32_t* d_prefix = reinterpret_cast<uint32_t*>(d + n * sizeof(lsio_codec_item));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d_items, items, n * sizeof(lsio_codec_item), cudaMemcpyHostToDevice, st));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d_prefix, prefix.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    int rc = lsio_codecs_unshuffle_async(device, d_items, d_prefix, n, (uint32_t)total, elem_size, forward, stream);
    if (rc) return rc;
    LSIO_CUDA_TRY(cudaFreeAsync(d, st));
    LSIO_CUDA_TRY(cudaStreamSynchronize(st));  // `prefix` and `items` are pageable host memory: keep the call self-contained
    return 0;
}

int lsio_codecs_unshuffle(int device, const lsio_codec_item* items, size_t n, uint32_t elem_size, void* stream) {
    return shuffle_host_items(device, items, n, elem_size, 0, stream);
}
int lsio_codecs_shuffle(int device, const lsio_codec_item* items, size_t n, uint32_t elem_size, void* stream) {
    return shuffle_host_items(device, items, n, elem_size, 1, stream);
}

int lsio_codecs_crc32c_trailer_ok(uint32_t crc32c_of_chunk_with_trailer) { return crc32c_of_chunk_with_trailer == LSIO_CRC32C_RESIDUE; }

}  // extern "C"
