// pack_crc.cu -- pack_strip / pack_strip_crc32c / crc32c kernels for sm_100a.
//
// What it computes (per segment i): dst[i][0..len) = src[i][0..len) -- i.e. the view
// [start_off - aligned_start, end_off - aligned_start) that the reference takes of each
// padded O_DIRECT read (crates/lsio_uring/src/sqe.rs:106-119), materialised at the request's
// prefix-sum offset in HBM -- and crc32c(dst[i]) fused into the same pass.
//
// Roofline: HBM.  Algorithmic bytes = 2*sum(len) (read once, write once; the CRC adds none).
//
// Design (B200-first, nothing here is a contraction so no tensor cores):
//  * Work is cut into 16-byte vectors aligned in DESTINATION space (so every interior store
//    is an aligned STG.128) and grouped in 512-byte rows (one warp-wide access) and tiles of
//    kTileRows rows (64 rows = 32 KiB), both counted from the END of the segment's 16-byte-aligned superset.
//    Counting from the end makes every tile "full" as a polynomial (leading zeros are free in
//    a CRC), so no variable-length shift operator is ever needed in the hot loop.
//  * src and dst may differ mod 16: each thread loads the two aligned source vectors that
//    cover its destination vector and funnel-shifts them (warp-uniform shift per segment).
//  * CRC-32C is linear over GF(2).  Each lane keeps 4 accumulators (one per 32-bit word of
//    its vector) and does Horner with the fixed stride operator G = multiply by x^4096
//    (512 bytes) mod P:  acc = G(acc) ^ word.  G is applied with four byte-sliced table
//    lookups.  The tables live in shared memory in a LANE-PRIVATE layout
//    (entry e of lane l at word e*32 + l -> bank l): every LDS is conflict free, which is what
//    lets one lookup per byte keep up with HBM.  128 KiB of the SM's 227 KiB go to this.
//  * Per tile the 4 accumulators and the 32 lanes are merged with fixed operators
//    (x^32, x^128 ... x^2048: small shared tables + warp shuffles), giving one 32-bit partial.
//    Single-tile segments finish in place; for longer ones a second, tiny kernel
//    (fold_kernel, one warp per segment) folds the partials with the fixed tile operator:
//    lanes run strided Horner chains with Z^32 tables, then a warp-cooperative GF(2) multiply
//    (one lane per bit + redux.xor) joins the lanes, undoes the trailing zero padding
//    (multiply by x^(32-8z)) and writes the CRC.  A kernel boundary instead of
//    __threadfence + atomics: the fence drained every warp's stores once per tile and cost
//    ~14 % of HBM bandwidth when measured.
//  * init/xorout: the 0xFFFFFFFF init is folded in by complementing the first 4 bytes of the
//    message; xorout is a final NOT.  Segments shorter than 4 bytes take a scalar path.
//  * Grid: one 1024-thread CTA per SM (persistent); warps take tiles IN ORDER from a global counter
//    (dynamic tile scheduler), which keeps the concurrently touched window dense.
#include "common.hpp"
#include "gf2.hpp"
#include "lsio_cuda.h"
#include "pack_batch.hpp"

#include <mutex>
#include <vector>

namespace lsio {
namespace {

constexpr uint32_t kPoly = gf2::kPoly;  // CRC-32C, reflected
constexpr int kThreads = 1024;
#ifndef LSIO_TILE_ROWS
#define LSIO_TILE_ROWS 64  // 32 KiB tiles: measured 88 -> 92 % of HBM peak vs 16 KiB (per-tile overhead, DRAM locality)
#endif
#ifndef LSIO_UNROLL
#define LSIO_UNROLL 4          // rows in flight per warp, src/dst congruent mod 16
#endif
#ifndef LSIO_UNROLL_SHIFTED
#define LSIO_UNROLL_SHIFTED 2  // ... when every vector needs two source loads (measured: 2 beats 4, registers)
#endif
#ifndef LSIO_PREFETCH_ROWS
#define LSIO_PREFETCH_ROWS 0   // rows of 512 B to prefetch into L2 ahead of the load stream (0 = off)
#endif
#ifndef LSIO_LD_L2_256
#define LSIO_LD_L2_256 0       // 1: ld...L2::256B prefetch-size hint on the streaming loads
#endif
#ifndef LSIO_ST_HINT
#define LSIO_ST_HINT 1   // 1: st.global.L1::no_allocate, 0: plain st.global, 2: st.global.cs
#endif
constexpr int kTileRows = LSIO_TILE_ROWS;     // rows of 512 B per tile
constexpr int kTileVecs = kTileRows * 32;     // 16-byte vectors per tile
constexpr int kTileBytes = kTileVecs * 16;    // 32 KiB

// ---- table block (uint32 words), built once on the host and copied to each device ----------
constexpr int kG = 0;                   // [4][256]  multiply by x^4096 (one row = 512 bytes)
constexpr int kH = kG + 1024;           // [6][4][256]: x^32, x^128, x^256, x^512, x^1024, x^2048
constexpr int kZ = kH + 6 * 1024;       // [32]      Z*x^i, Z = x^(8*kTileBytes)
constexpr int kF = kZ + 32;             // [16][32]  x^(32-8z)*x^i, z = trailing zero bytes
constexpr int kZ32 = kF + 16 * 32;      // [4][256]  multiply by Z^32 (fold_kernel's lane stride)
constexpr int kTableWords = kZ32 + 1024;

// shared memory layout (words)
constexpr int kSmG = 0;                 // lane-private G: [4][256][32]
constexpr int kSmH = kSmG + 4 * 256 * 32;
constexpr int kSmZ = kSmH + 6 * 1024;
constexpr int kSmF = kSmZ + 32;
constexpr int kSmWords = kSmF + 16 * 32;
constexpr size_t kSmBytes = (size_t)kSmWords * 4;

using gf2::mul_x;
using gf2::multmodp;
using gf2::xpow;
void fill_byte_tables(uint32_t* t, uint32_t k) {  // t[b][v] = (v << 8b) * k
    for (int b = 0; b < 4; b++)
        for (uint32_t v = 0; v < 256; v++) t[b * 256 + v] = multmodp(v << (8 * b), k);
}
void fill_bit_row(uint32_t* t, uint32_t k) {  // t[i] = k * x^i
    for (int i = 0; i < 32; i++) {
        t[i] = k;
        k = mul_x(k);
    }
}
const std::vector<uint32_t>& host_tables() {
    static std::vector<uint32_t> t = [] {
        std::vector<uint32_t> v(kTableWords);
        fill_byte_tables(&v[kG], xpow(8 * 512));
        const int hbytes[6] = {4, 16, 32, 64, 128, 256};
        for (int i = 0; i < 6; i++) fill_byte_tables(&v[kH + i * 1024], xpow(8 * hbytes[i]));
        fill_bit_row(&v[kZ], xpow(8 * (int64_t)kTileBytes));
        for (int z = 0; z < 16; z++) fill_bit_row(&v[kF + z * 32], xpow(32 - 8 * z));
        fill_byte_tables(&v[kZ32], xpow(8 * (int64_t)kTileBytes * 32));
        return v;
    }();
    return t;
}

struct DeviceTables {
    std::mutex mu;
    uint32_t* dev[64] = {};
    int sms[64] = {};
};
DeviceTables& device_tables() {
    static DeviceTables d;
    return d;
}

// ---- device descriptors ------------------------------------------------------------------
struct Seg {  // 32 bytes
    uint64_t src, dst, len;
    uint32_t tile_base;  // index of this segment's first partial / tile-table entry
    uint32_t ntiles;
};
struct Params {
    const Seg* segs;
    const uint2* tiles;  // {segment, tile index counted from the segment's end}
    uint32_t ntiles;
    uint32_t* partials;  // one per tile
    const uint32_t* multi;  // indices of the segments with more than one tile
    uint32_t nmulti;
    uint32_t* crc_out;   // one per segment
    const uint32_t* tables;
    uint32_t* sched;     // [0] next tile, [1] warps finished: the dynamic tile scheduler (zero between launches)
    uint32_t* seg_done;  // light variant: tiles finished per segment (zero between launches); the last one folds
};

// ---- device helpers ----------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
#if LSIO_LD_L2_256
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
#endif
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint4 ld_cached(const uint4* p) { return __ldg(p); }
__device__ __forceinline__ void st_vec(uint4* p, const uint4& v) {
#if LSIO_ST_HINT == 1
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
#elif LSIO_ST_HINT == 2
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
    *p = v;
#endif
}

// acc * x^4096 with the lane-private tables; ltb = (char*)gtab + lane*4
__device__ __forceinline__ uint32_t mul_g_private(uint32_t a, const char* ltb) {
    const uint32_t o0 = (a << 7) & 0x7F80u;
    const uint32_t o1 = (a >> 1) & 0x7F80u;
    const uint32_t o2 = (a >> 9) & 0x7F80u;
    const uint32_t o3 = (a >> 17) & 0x7F80u;
    return *(const uint32_t*)(ltb + o0) ^ *(const uint32_t*)(ltb + 32768 + o1) ^ *(const uint32_t*)(ltb + 65536 + o2) ^
           *(const uint32_t*)(ltb + 98304 + o3);
}
// Light variant: one shared [4][256] copy of G (4 KiB per CTA instead of 128 KiB; bank conflicts accepted:
// its launches are small pipeline batches bound by latency and by PCIe, not by shared-memory throughput).
__device__ __forceinline__ uint32_t mul_g_shared(uint32_t a, const char* tb) {
    const uint32_t* t = (const uint32_t*)tb;
    return t[a & 0xFF] ^ t[256 + ((a >> 8) & 0xFF)] ^ t[512 + ((a >> 16) & 0xFF)] ^ t[768 + (a >> 24)];
}
template <bool kLight>
__device__ __forceinline__ uint32_t mul_g(uint32_t a, const char* tb) {
    return kLight ? mul_g_shared(a, tb) : mul_g_private(a, tb);
}
// a * K with an ordinary [4][256] byte-sliced table (rare path: bank conflicts tolerated)
__device__ __forceinline__ uint32_t mul_t(uint32_t a, const uint32_t* t) {
    return t[a & 0xFF] ^ t[256 + ((a >> 8) & 0xFF)] ^ t[512 + ((a >> 16) & 0xFF)] ^ t[768 + (a >> 24)];
}
// c * K where row[i] = K*x^i; c is warp-uniform; one lane per bit, xor-reduced in hardware.
__device__ __forceinline__ uint32_t warp_mul(uint32_t c, const uint32_t* row, int lane) {
    const uint32_t term = ((c >> (31 - lane)) & 1u) ? row[lane] : 0u;
    return __reduce_xor_sync(0xFFFFFFFFu, term);
}
// bytes [lo, hi) of a 16-byte vector as four word masks
__device__ __forceinline__ uint32_t word_mask(int lo, int hi) {  // lo, hi in word-local units, any int
    lo = max(0, min(4, lo));
    hi = max(0, min(4, hi));
    const uint32_t mh = hi >= 4 ? 0xFFFFFFFFu : ((1u << (8 * hi)) - 1u);
    const uint32_t ml = lo >= 4 ? 0xFFFFFFFFu : ((1u << (8 * lo)) - 1u);
    return mh & ~ml;
}
__device__ __forceinline__ uint4 vec_mask(int lo, int hi) {
    return make_uint4(word_mask(lo, hi), word_mask(lo - 4, hi - 4), word_mask(lo - 8, hi - 8), word_mask(lo - 12, hi - 12));
}
__device__ __forceinline__ uint32_t crc_bitwise_dev(const uint8_t* p, int n) {
    uint32_t c = 0xFFFFFFFFu;
    for (int i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (kPoly & (0u - (c & 1u)));
    }
    return ~c;
}

// Per-segment geometry, warp-uniform.
struct Geo {
    uint64_t dst, len;
    uint64_t a0;         // 16-aligned start of the destination superset
    const uint4* sbase;  // 16-aligned source address of destination vector 0's low half
    uint64_t src_lo, src_hi;  // [src, src+len)
    int nvec;            // vectors in the superset
    int delta;           // (src - dst) mod 16
};

// Fetch destination-aligned vector v.  WS = word shift (delta >> 2), or -1 for delta == 0.
// kGuard: predicate the two source loads on intersecting [src, src+len) (edge vectors only).
template <int WS, bool kGuard>
__device__ __forceinline__ uint4 fetch(const Geo& g, int v, int bs) {
    const uint4* q = g.sbase + v;
    if (WS < 0) {
        if (kGuard) {
            const uint64_t qa = (uint64_t)q;
            if (!(qa + 16 > g.src_lo && qa < g.src_hi)) return make_uint4(0, 0, 0, 0);
            return ld_cached(q);
        }
        return ld_stream(q);
    } else {
        uint4 lo, hi;
        if (kGuard) {
            const uint64_t qa = (uint64_t)q;
            lo = (qa + 16 > g.src_lo && qa < g.src_hi) ? ld_cached(q) : make_uint4(0, 0, 0, 0);
            hi = (qa + 32 > g.src_lo && qa + 16 < g.src_hi) ? ld_cached(q + 1) : make_uint4(0, 0, 0, 0);
        } else {
            lo = ld_cached(q);
            hi = ld_cached(q + 1);
        }
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        uint4 r;
        r.x = __funnelshift_r(w[WS + 0], w[WS + 1], bs);
        r.y = __funnelshift_r(w[WS + 1], w[WS + 2], bs);
        r.z = __funnelshift_r(w[WS + 2], w[WS + 3], bs);
        r.w = __funnelshift_r(w[WS + 3], w[(WS + 4) & 7], bs);
        return r;
    }
}

__device__ __forceinline__ void store_bytes(uint64_t addr, const uint4& d, int lo, int hi) {
    const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 16; j++)
        if (j >= lo && j < hi) *(uint8_t*)(addr + j) = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
}

// One tile (kTileRows rows ending at vector `vend`, exclusive) of one segment, by one warp.
// Returns the tile partial (valid in every lane): the tile's bytes as a polynomial mod P,
// positioned at the tile's last word.
// `h`: the six [4][256] merge tables (shared memory in the roofline variant, global/L1 in the light one).
template <int WS, bool kStore, bool kCrc, bool kLight>
__device__ __forceinline__ uint32_t run_tile(const Geo& g, int vend, int lane, const char* ltb, const uint32_t* h) {
    // the light variant runs few warps per SM with registers to spare: twice the rows in flight per warp
    constexpr int kUnroll = (WS < 0 ? LSIO_UNROLL : LSIO_UNROLL_SHIFTED) * (kLight ? 2 : 1);
    const int bs = (g.delta & 3) * 8;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    uint4* const dvec = (uint4*)g.a0;
    const int head = (int)(g.dst - g.a0);                    // bytes before dst in vector 0
    const int64_t end_rel = (int64_t)(g.dst + g.len - g.a0);  // dst end relative to a0
    int vr = vend - kTileVecs;
    // skip rows that lie wholly before the segment (leading zeros: free)
    if (vr < 0) vr += ((-vr) >> 5) << 5;
    constexpr int kPf = LSIO_PREFETCH_ROWS;
    const int pf_end = min(vend, g.nvec);  // prefetch only source vectors of this tile
    if (kPf > 0) {  // prologue: fill the L2 prefetch window
#pragma unroll 1
        for (int r = kUnroll; r < kPf; r++) {
            const int v = vr + 32 * r + lane;
            if (v < pf_end) prefetch_l2(g.sbase + v);
        }
    }
    for (; vr < vend;) {
        // interior rows: every vector lies wholly inside the segment and past the init bytes
        const bool interior4 = vr >= 2 && vr + 32 * kUnroll <= min(vend, g.nvec - 1);
        if (interior4) {
            if (kPf > 0) {
#pragma unroll
                for (int u = 0; u < kUnroll; u++) {
                    const int v = vr + 32 * (kPf + u) + lane;
                    if (v < pf_end) prefetch_l2(g.sbase + v);
                }
            }
            uint4 d[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) d[u] = fetch<WS, false>(g, vr + 32 * u + lane, bs);
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                if (kStore) st_vec(dvec + vr + 32 * u + lane, d[u]);
                if (kCrc) {
                    a0 = mul_g<kLight>(a0, ltb) ^ d[u].x;
                    a1 = mul_g<kLight>(a1, ltb) ^ d[u].y;
                    a2 = mul_g<kLight>(a2, ltb) ^ d[u].z;
                    a3 = mul_g<kLight>(a3, ltb) ^ d[u].w;
                }
            }
            vr += 32 * kUnroll;
            continue;
        }
        if (vr >= 2 && vr + 32 <= g.nvec - 1) {  // a single interior row (next to a segment edge or the tile end)
            const uint4 d = fetch<WS, false>(g, vr + lane, bs);
            if (kStore) st_vec(dvec + vr + lane, d);
            if (kCrc) {
                a0 = mul_g<kLight>(a0, ltb) ^ d.x;
                a1 = mul_g<kLight>(a1, ltb) ^ d.y;
                a2 = mul_g<kLight>(a2, ltb) ^ d.z;
                a3 = mul_g<kLight>(a3, ltb) ^ d.w;
            }
            vr += 32;
            continue;
        }
        // generic row: may contain the segment's first/last vector or sit before its start
        const int v = vr + lane;
        uint4 d = make_uint4(0, 0, 0, 0);
        if (v >= 0 && v < g.nvec) {
            d = fetch<WS, true>(g, v, bs);
            const int lo = v == 0 ? head : 0;
            const int64_t hi64 = end_rel - 16 * (int64_t)v;
            const int hi = hi64 >= 16 ? 16 : (int)hi64;
            if (kStore) {
                if (lo == 0 && hi == 16) st_vec(dvec + v, d);
                else store_bytes(g.a0 + 16 * (uint64_t)v, d, lo, hi);
            }
            if (kCrc) {
                uint4 m = vec_mask(lo, hi);
                d.x &= m.x; d.y &= m.y; d.z &= m.z; d.w &= m.w;
                if (v <= 1) {  // fold the 0xFFFFFFFF init into the first 4 message bytes
                    const int i0 = head - 16 * v;
                    m = vec_mask(i0, i0 + 4);
                    d.x ^= m.x; d.y ^= m.y; d.z ^= m.z; d.w ^= m.w;
                }
            }
        }
        if (kCrc) {
            a0 = mul_g<kLight>(a0, ltb) ^ d.x;
            a1 = mul_g<kLight>(a1, ltb) ^ d.y;
            a2 = mul_g<kLight>(a2, ltb) ^ d.z;
            a3 = mul_g<kLight>(a3, ltb) ^ d.w;
        }
        vr += 32;
    }
    if (!kCrc) return 0;
    // merge the four word accumulators (x^32 apart), then the 32 lanes (x^128 apart)
    uint32_t V = mul_t(mul_t(mul_t(a0, h) ^ a1, h) ^ a2, h) ^ a3;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, V, 1 << j);
        const uint32_t m = (2u << j) - 1u;
        const uint32_t t = mul_t(up, h + (j + 1) * 1024);
        if ((lane & m) == m) V ^= t;
    }
    return __shfl_sync(0xFFFFFFFFu, V, 31);
}

// One warp folds a multi-tile segment: V = sum_k partial[k] * Z^k (k counted from the segment's end).
// Lane t folds k = t, t+32, ... with Z^32 (byte tables), then the lanes are joined with Z; the trailing zero
// padding is undone (x^(32-8z)) and the final NOT applied.  `parts` is read with L1 bypassed (other SMs wrote it).
__device__ __forceinline__ uint32_t fold_segment(const uint32_t* parts, int n, uint64_t dst, uint64_t len, const uint32_t* z32, const uint32_t* zrow,
                                                 const uint32_t* frow, int lane) {
    uint32_t L = 0;
    if (lane < n) {
        int k = lane + ((n - 1 - lane) & ~31);  // highest k congruent to lane
        L = __ldcg(parts + k);
        for (k -= 32; k >= 0; k -= 32) L = mul_t(L, z32) ^ __ldcg(parts + k);
    }
    uint32_t V = 0;
    for (int t = min(31, n - 1); t >= 0; t--) V = warp_mul(V, zrow, lane) ^ __shfl_sync(0xFFFFFFFFu, L, t);
    const int z = (int)(((dst + len + 15ull) & ~15ull) - (dst + len));
    return ~warp_mul(V, frow + z * 32, lane);
}

// Roofline variant (kLight = false): one 1024-thread CTA per SM, 128 KiB of lane-private tables, multi-tile
// segments finished by fold_kernel.  Light variant: 256-thread CTAs that need 4 KiB of shared memory and fit
// next to a consumer's kernels; merge tables come from global memory through L1; the warp that finishes a
// segment's last tile folds the partials itself (one launch per batch); a segment with src == dst (landed
// directly by DMA) is checksummed in place.
// The tables a warp needs, wherever the variant keeps them.
struct TileTables {
    const char* ltb;       // G: lane-private (roofline variant) or one shared copy (light / pull)
    const uint32_t* htab;  // six [4][256] merge tables
    const uint32_t* ftab;  // [16][32] trailing-zero / final operators
};

// One tile of one segment by one warp: move (unless in place / crc-only), checksum, and finish the segment's CRC if
// this was its only or -- light variant -- its last tile.  `si`: the segment's index (its CRC slot), `k`: the tile's
// index counted from the segment's end.
template <bool kStore, bool kCrc, bool kLight>
__device__ __forceinline__ void do_tile(const Seg& s, uint32_t si, uint32_t k, int lane, const TileTables& tb, const uint32_t* tables, uint32_t* partials,
                                        uint32_t* seg_done, uint32_t* crc_out) {
    const uint64_t dst = kStore ? s.dst : s.src;  // crc-only: checksum in place
    const bool in_place = kStore && s.src == s.dst;  // landed by DMA: nothing to move
    if (!kCrc && in_place) return;
    if (s.len < 4) {  // scalar path (covers empty segments): ntiles == 1 by construction
        if (lane == 0) {
            uint8_t b[4];
            for (int i = 0; i < (int)s.len; i++) b[i] = *(const uint8_t*)(s.src + i);
            if (kStore && !in_place) for (int i = 0; i < (int)s.len; i++) *(uint8_t*)(s.dst + i) = b[i];
            if (kCrc) crc_out[si] = crc_bitwise_dev(b, (int)s.len);
        }
        return;
    }
    Geo g;
    g.dst = dst; g.len = s.len;
    g.a0 = dst & ~15ull;
    const uint64_t a1 = (dst + s.len + 15ull) & ~15ull;
    g.nvec = (int)((a1 - g.a0) >> 4);
    g.delta = (int)((s.src - dst) & 15ull);
    g.sbase = (const uint4*)(g.a0 + (s.src - dst) - (uint64_t)g.delta);
    g.src_lo = s.src; g.src_hi = s.src + s.len;
    const int vend = g.nvec - (int)k * kTileVecs;
    uint32_t part;
    if (in_place) part = run_tile<-1, false, kCrc, kLight>(g, vend, lane, tb.ltb, tb.htab);
    else if (g.delta == 0) part = run_tile<-1, kStore, kCrc, kLight>(g, vend, lane, tb.ltb, tb.htab);
    else switch (g.delta >> 2) {
        case 0: part = run_tile<0, kStore, kCrc, kLight>(g, vend, lane, tb.ltb, tb.htab); break;
        case 1: part = run_tile<1, kStore, kCrc, kLight>(g, vend, lane, tb.ltb, tb.htab); break;
        case 2: part = run_tile<2, kStore, kCrc, kLight>(g, vend, lane, tb.ltb, tb.htab); break;
        default: part = run_tile<3, kStore, kCrc, kLight>(g, vend, lane, tb.ltb, tb.htab); break;
    }
    if (!kCrc) return;

    if (s.ntiles > 1) {
        if (!kLight) {  // fold_kernel finishes multi-tile segments
            if (lane == 0) partials[s.tile_base + k] = part;
            return;
        }
        // light: publish the partial; whoever completes the segment folds it.  The fence orders the partial
        // (and, harmlessly, this warp's data stores) before the count; it costs nothing at pipeline rates.
        uint32_t before = 0;
        if (lane == 0) {
            partials[s.tile_base + k] = part;
            __threadfence();
            before = atomicAdd(seg_done + si, 1u);
        }
        before = __shfl_sync(0xFFFFFFFFu, before, 0);
        if (before != s.ntiles - 1) return;
        __threadfence();
        const uint32_t crc = fold_segment(partials + s.tile_base, (int)s.ntiles, dst, s.len, tables + kZ32, tables + kZ, tables + kF, lane);
        if (lane == 0) {
            crc_out[si] = crc;
            seg_done[si] = 0;  // re-armed for the next launch of this batch / plan
        }
        return;
    }
    // undo the z trailing zero bytes of the aligned superset and apply the final x^32
    const int z = (int)(a1 - (dst + s.len));
    const uint32_t crc = ~warp_mul(part, tb.ftab + z * 32, lane);
    if (lane == 0) crc_out[si] = crc;
}

template <bool kStore, bool kCrc, bool kLight>
__global__ void __launch_bounds__(kLight ? 256 : kThreads, kLight ? 3 : 1) pack_kernel(const Params p) {
    extern __shared__ __align__(16) uint32_t sh[];
    const int tid = threadIdx.x, lane = tid & 31;
    if (kCrc) {
        if (kLight) {
            for (int i = tid; i < 1024; i += blockDim.x) sh[i] = __ldg(p.tables + kG + i);
        } else {
            // replicate the 4 KiB G table into the lane-private layout; copy the small tables
            for (int i = tid; i < 1024 * 32; i += blockDim.x) sh[kSmG + i] = __ldg(p.tables + kG + (i >> 5));
            for (int i = tid; i < kSmWords - kSmH; i += blockDim.x) sh[kSmH + i] = __ldg(p.tables + kH + i);
        }
        __syncthreads();
    }
    const TileTables tb{kLight ? (const char*)sh : (const char*)(sh + kSmG) + lane * 4, kLight ? p.tables + kH : sh + kSmH, kLight ? p.tables + kF : sh + kSmF};

    // Dynamic tile scheduler: every warp takes the NEXT tile from a global counter, like the hardware
    // block scheduler would.  Measured on a plain copy (tools/copy_pattern_bench.cu): static striding lets
    // warps drift apart until the concurrently touched window spans hundreds of MB (94 % of the HBM copy
    // peak); taking tiles in order keeps a dense sliding window (101 %).
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(p.sched, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= p.ntiles) break;
        const uint2 te = __ldg(p.tiles + t);
        const Seg s = *(p.segs + te.x);  // 32 B, warp-uniform broadcast
        do_tile<kStore, kCrc, kLight>(s, te.x, te.y, lane, tb, p.tables, p.partials, p.seg_done, p.crc_out);
    }
    // the last warp to run dry re-arms the scheduler for the next launch of this batch / plan
    if (lane == 0) {
        const uint32_t done = atomicAdd(p.sched + 1, 1u);
        if (done == gridDim.x * (blockDim.x >> 5) - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
        }
    }
}

// ---- pull kernel: one extent straight from its pinned host slot into the request buffers ----------------------------
// The segments' `src` are MAPPED HOST addresses: the warps read the slot over PCIe themselves (no copy engine, no device
// staging ring, no second pass), checksum what they move and write it at its final place: PCIe traffic L, HBM traffic L.
// Why a kernel and not cudaMemcpyAsync: the copy engine loses ~3.8 us per copy whatever its size, so cache-sized slots
// cost bandwidth (512 KiB copies 36.5 GiB/s, 2 MiB 46.9 of 51.6); SM loads reach 47-47.8 GiB/s from 512 KiB up
// (tools/h2d_probe.cu), which lets the slots shrink until the worker's whole window stays in its cache.
// The descriptors travel in the kernel parameters (no upload); partials / counters are per-slot device scratch.
struct PullParams {
    Seg segs[kPullMaxSegs];
    uint32_t nsegs, ntiles;
    uint32_t* scratch;  // [0] next tile, [1] warps finished, [2..2+kPullMaxSegs) tiles done per segment, then partials
    uint32_t* crc_out;  // one per segment (mapped host memory)
    const uint32_t* tables;
};
template <bool kCrc>
__global__ void __launch_bounds__(256, 3) pull_kernel(const __grid_constant__ PullParams p) {
    __shared__ __align__(16) uint32_t sh[1024];
    const int tid = threadIdx.x, lane = tid & 31;
    if (kCrc) {
        for (int i = tid; i < 1024; i += blockDim.x) sh[i] = __ldg(p.tables + kG + i);
        __syncthreads();
    }
    const TileTables tb{(const char*)sh, p.tables + kH, p.tables + kF};
    uint32_t* const seg_done = p.scratch + 2;
    uint32_t* const partials = p.scratch + 2 + kPullMaxSegs;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(p.scratch, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= p.ntiles) break;
        uint32_t si = 0;
        while (si + 1 < p.nsegs && t >= p.segs[si + 1].tile_base) si++;  // warp-uniform scan of at most kPullMaxSegs entries
        const Seg s = p.segs[si];
        do_tile<true, kCrc, true>(s, si, s.ntiles - 1 - (t - s.tile_base), lane, tb, p.tables, partials, seg_done, p.crc_out);
    }
    if (lane == 0) {
        const uint32_t done = atomicAdd(p.scratch + 1, 1u);
        if (done == gridDim.x * (blockDim.x >> 5) - 1) {
            p.scratch[0] = 0;
            p.scratch[1] = 0;
        }
    }
}

// One warp per multi-tile segment (roofline variant only; a kernel boundary instead of fences in the hot loop).
template <bool kStore>
__global__ void __launch_bounds__(256) fold_kernel(const Params p) {
    __shared__ uint32_t z32[1024];
    __shared__ uint32_t zrow[32];
    __shared__ uint32_t frow[16 * 32];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) z32[i] = __ldg(p.tables + kZ32 + i);
    for (int i = threadIdx.x; i < 32; i += blockDim.x) zrow[i] = __ldg(p.tables + kZ + i);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) frow[i] = __ldg(p.tables + kF + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= p.nmulti) return;
    const uint32_t si = __ldg(p.multi + w);
    const Seg s = p.segs[si];
    const uint32_t crc = fold_segment(p.partials + s.tile_base, (int)s.ntiles, kStore ? s.dst : s.src, s.len, z32, zrow, frow, lane);
    if (lane == 0) p.crc_out[si] = crc;
}

}  // namespace

// ---- host side: batches and plans ------------------------------------------------------------
}  // namespace lsio

namespace lsio {

static int get_device_tables(int device, const uint32_t** out, int* sms) {
    if (device < 0 || device >= 64) return -EINVAL;
    DeviceTables& dt = device_tables();
    std::lock_guard<std::mutex> lk(dt.mu);
    if (!dt.dev[device]) {
        DeviceGuard dg(device);
        if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
        const std::vector<uint32_t>& h = host_tables();
        uint32_t* d = nullptr;
        LSIO_CUDA_TRY(cudaMalloc(&d, h.size() * 4));
        LSIO_CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        int n = 0;
        LSIO_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
        LSIO_CUDA_TRY(cudaFuncSetAttribute(pack_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmBytes));
        LSIO_CUDA_TRY(cudaFuncSetAttribute(pack_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmBytes));
        dt.dev[device] = d;
        dt.sms[device] = n;
    }
    *out = dt.dev[device];
    *sms = dt.sms[device];
    return 0;
}

int pack_prepare_device(int device) {
    const uint32_t* t;
    int sms;
    return get_device_tables(device, &t, &sms);
}

size_t pack_tile_count(uint64_t dst, uint64_t len) {
    if (len < 4) return 1;
    const uint64_t a0 = dst & ~15ull, a1 = (dst + len + 15ull) & ~15ull;
    const uint64_t nvec = (a1 - a0) >> 4;
    return (size_t)((nvec + kTileVecs - 1) / kTileVecs);
}

PackBatchLayout pack_batch_layout(size_t max_segs, size_t max_tiles) {
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    PackBatchLayout l;
    l.max_segs = max_segs;
    l.max_tiles = max_tiles;
    l.off_tiles = up(max_segs * sizeof(Seg));
    l.off_multi = l.off_tiles + up(max_tiles * sizeof(uint2));
    l.upload_bytes = l.off_multi + up(max_segs * 4);
    l.off_partials = l.upload_bytes;
    l.off_sched = l.off_partials + up(max_tiles * 4);
    l.off_segdone = l.off_sched + 256;
    l.total_bytes = l.off_segdone + up(max_segs * 4);
    return l;
}

bool pack_batch_append(const PackBatchLayout& lay, void* host_block, PackBatchCounts* c, const lsio_cuda_segment& seg, bool store) {
    const uint64_t d = store ? seg.dst : seg.src;
    const size_t nt = pack_tile_count(d, seg.len);
    if (c->nsegs + 1 > lay.max_segs || c->ntiles + nt > lay.max_tiles) return false;
    char* base = (char*)host_block;
    Seg* segs = (Seg*)base;
    uint2* tiles = (uint2*)(base + lay.off_tiles);
    uint32_t* multi = (uint32_t*)(base + lay.off_multi);
    segs[c->nsegs] = Seg{seg.src, seg.dst, seg.len, c->ntiles, (uint32_t)nt};
    // first tile in memory order first, so neighbouring warps touch neighbouring memory
    for (size_t k = nt; k-- > 0;) tiles[c->ntiles++] = make_uint2(c->nsegs, (unsigned)k);
    if (nt > 1) multi[c->nmulti++] = c->nsegs;
    c->nsegs++;
    c->payload += seg.len;
    return true;
}

// Tiles one pull launch may carry: the segments of an extent can overlap in the SLOT (many ranges over the same bytes), so the payload
// -- and with it the number of tiles and partials -- is not bounded by the slot size.  Four times the slot plus the per-segment edge
// tiles, at least 1024; a launch that needs more is refused (-E2BIG) and the caller takes the ring path.
uint32_t pack_pull_max_tiles(size_t slot_bytes) { return (uint32_t)std::max<size_t>(1024, 4 * (slot_bytes / kTileBytes) + 2 * kPullMaxSegs); }
size_t pack_pull_scratch_bytes(size_t slot_bytes) { return (2 + kPullMaxSegs + (size_t)pack_pull_max_tiles(slot_bytes)) * sizeof(uint32_t); }

int pack_pull_launch(int device, const lsio_cuda_segment* segs, uint32_t n, bool crc, void* scratch_dev, uint32_t max_tiles, uint32_t* crc_out_dev, cudaStream_t st) {
    if (!segs || n == 0 || n > kPullMaxSegs || !scratch_dev) return -EINVAL;
    uint64_t want_tiles = 0;
    for (uint32_t i = 0; i < n; i++) want_tiles += pack_tile_count(segs[i].dst, segs[i].len);
    if (want_tiles > max_tiles) return -E2BIG;
    const uint32_t* tables;
    int sms;
    int rc = get_device_tables(device, &tables, &sms);
    if (rc) return rc;
    PullParams p;
    uint32_t ntiles = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t nt = (uint32_t)pack_tile_count(segs[i].dst, segs[i].len);
        p.segs[i] = Seg{segs[i].src, segs[i].dst, segs[i].len, ntiles, nt};
        ntiles += nt;
    }
    p.nsegs = n;
    p.ntiles = ntiles;
    p.scratch = (uint32_t*)scratch_dev;
    p.crc_out = crc_out_dev;
    p.tables = tables;
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)sms * 3, (ntiles + 7) / 8);
    if (crc) pull_kernel<true><<<grid, 256, 0, st>>>(p);
    else pull_kernel<false><<<grid, 256, 0, st>>>(p);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    LSIO_CUDA_TRY(cudaGetLastError());
    return 0;
}

int pack_batch_upload(const PackBatchLayout& lay, const void* host_block, void* dev_block, const PackBatchCounts& c, cudaStream_t st) {
    const char* h = (const char*)host_block;
    char* d = (char*)dev_block;
    LSIO_CUDA_TRY(cudaMemcpyAsync(d, h, (size_t)c.nsegs * sizeof(Seg), cudaMemcpyHostToDevice, st));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d + lay.off_tiles, h + lay.off_tiles, (size_t)c.ntiles * sizeof(uint2), cudaMemcpyHostToDevice, st));
    if (c.nmulti) LSIO_CUDA_TRY(cudaMemcpyAsync(d + lay.off_multi, h + lay.off_multi, (size_t)c.nmulti * 4, cudaMemcpyHostToDevice, st));
    return 0;
}

int pack_batch_launch(int device, const PackBatchLayout& lay, void* dev_block, const PackBatchCounts& c, uint32_t flags,
                      uint32_t* crc_out, cudaStream_t st, unsigned* launches, unsigned threads) {
    const bool light = flags & LSIO_CUDA_PACK_LIGHT;
    if (threads == 0 || light) threads = light ? 256 : kThreads;
    if (threads % 32 || threads > (unsigned)kThreads) return -EINVAL;
    const bool store = flags & LSIO_CUDA_PACK_STORE, crc = flags & LSIO_CUDA_PACK_CRC32C;
    if (launches) *launches = 0;
    if (c.ntiles == 0) return 0;
    const uint32_t* tables = nullptr;
    int sms = 0;
    int rc = get_device_tables(device, &tables, &sms);
    if (rc) return rc;
    char* base = (char*)dev_block;
    Params p{(const Seg*)base, (const uint2*)(base + lay.off_tiles), c.ntiles, (uint32_t*)(base + lay.off_partials),
             (const uint32_t*)(base + lay.off_multi), c.nmulti, crc_out, tables, (uint32_t*)(base + lay.off_sched), (uint32_t*)(base + lay.off_segdone)};
    if (light) {
        // small CTAs, several per SM; every warp takes tiles from the counter, so any grid size is correct
        const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)sms * 6, (c.ntiles + 7) / 8);
        if (store && crc) pack_kernel<true, true, true><<<grid, 256, 4096, st>>>(p);
        else if (crc) pack_kernel<false, true, true><<<grid, 256, 4096, st>>>(p);
        else pack_kernel<true, false, true><<<grid, 256, 0, st>>>(p);
        launch_counter().fetch_add(1, std::memory_order_relaxed);
        if (launches) ++*launches;
        LSIO_CUDA_TRY(cudaGetLastError());
        return 0;
    }
    // one CTA per SM; `threads` < 1024 is the small-footprint launch for a GPU shared with the consumer's kernels
    // (a 1024-thread CTA with 64 registers per thread needs a whole SM to itself)
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)sms, (c.ntiles + threads / 32 - 1) / (threads / 32));
    if (store && crc) pack_kernel<true, true, false><<<grid, threads, kSmBytes, st>>>(p);
    else if (crc) pack_kernel<false, true, false><<<grid, threads, kSmBytes, st>>>(p);
    else pack_kernel<true, false, false><<<grid, threads, 0, st>>>(p);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    if (launches) ++*launches;
    LSIO_CUDA_TRY(cudaGetLastError());
    if (crc && c.nmulti) {
        const unsigned fgrid = (c.nmulti + 7) / 8;
        if (store) fold_kernel<true><<<fgrid, 256, 0, st>>>(p);
        else fold_kernel<false><<<fgrid, 256, 0, st>>>(p);
        launch_counter().fetch_add(1, std::memory_order_relaxed);
        if (launches) ++*launches;
        LSIO_CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

}  // namespace lsio

struct lsio_cuda_pack_plan {
    int device = 0;
    uint32_t flags = 0;
    lsio::PackBatchLayout lay;
    lsio::PackBatchCounts counts;
    void* dev_block = nullptr;
};

extern "C" {

int lsio_cuda_pack_plan_create(int device, const lsio_cuda_segment* segs, size_t n, uint32_t flags, lsio_cuda_pack_plan** out) {
    using namespace lsio;
    if (!out || (!segs && n) || !(flags & (LSIO_CUDA_PACK_STORE | LSIO_CUDA_PACK_CRC32C)) || (flags & ~(LSIO_CUDA_PACK_STORE | LSIO_CUDA_PACK_CRC32C | LSIO_CUDA_PACK_LIGHT)))
        return -EINVAL;
    const bool store = flags & LSIO_CUDA_PACK_STORE;
    size_t ntiles = 0;
    for (size_t i = 0; i < n; i++) {
        if (segs[i].len >= (1ull << 34)) return -EFBIG;  // vector indices are 32-bit in the kernel
        ntiles += pack_tile_count(store ? segs[i].dst : segs[i].src, segs[i].len);
    }
    if (ntiles > 0xFFFFFFF0ull || n > 0xFFFFFFF0ull) return -EFBIG;
    int rc = pack_prepare_device(device);
    if (rc) return rc;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    auto* plan = new lsio_cuda_pack_plan();
    plan->device = device;
    plan->flags = flags;
    plan->lay = pack_batch_layout(n ? n : 1, ntiles ? ntiles : 1);
    std::vector<char> host(plan->lay.upload_bytes);
    for (size_t i = 0; i < n; i++) pack_batch_append(plan->lay, host.data(), &plan->counts, segs[i], store);
    cudaError_t e = cudaMalloc(&plan->dev_block, plan->lay.total_bytes);
    if (e == cudaSuccess) e = cudaMemset(plan->dev_block, 0, plan->lay.total_bytes);  // scheduler words start at zero
    if (e == cudaSuccess) e = cudaMemcpy(plan->dev_block, host.data(), plan->lay.upload_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (plan->dev_block) cudaFree(plan->dev_block);
        delete plan;
        return cuda_fail(e, "plan upload", __FILE__, __LINE__);
    }
    *out = plan;
    return 0;
}

int lsio_cuda_pack_plan_launch(lsio_cuda_pack_plan* plan, uint32_t* crc_out_dev, void* stream) {
    using namespace lsio;
    if (!plan) return -EINVAL;
    if ((plan->flags & LSIO_CUDA_PACK_CRC32C) && !crc_out_dev && plan->counts.nsegs) return -EINVAL;
    DeviceGuard dg(plan->device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    return pack_batch_launch(plan->device, plan->lay, plan->dev_block, plan->counts, plan->flags, crc_out_dev, (cudaStream_t)stream, nullptr, 0);
}

uint64_t lsio_cuda_pack_plan_algorithmic_bytes(const lsio_cuda_pack_plan* plan) {
    if (!plan) return 0;
    return (plan->flags & LSIO_CUDA_PACK_STORE) ? 2 * plan->counts.payload : plan->counts.payload;
}
uint64_t lsio_cuda_pack_plan_num_tiles(const lsio_cuda_pack_plan* plan) { return plan ? plan->counts.ntiles : 0; }

void lsio_cuda_pack_plan_destroy(lsio_cuda_pack_plan* plan) {
    if (!plan) return;
    lsio::DeviceGuard dg(plan->device);
    if (plan->dev_block) cudaFree(plan->dev_block);
    delete plan;
}

int lsio_cuda_pack_strip_crc32c(int device, const lsio_cuda_segment* segs, size_t n, uint32_t flags, uint32_t* crc_out_dev, void* stream) {
    lsio_cuda_pack_plan* plan = nullptr;
    int rc = lsio_cuda_pack_plan_create(device, segs, n, flags, &plan);
    if (rc) return rc;
    rc = lsio_cuda_pack_plan_launch(plan, crc_out_dev, stream);
    if (rc == 0) {
        lsio::DeviceGuard dg(device);
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) rc = lsio::cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    lsio_cuda_pack_plan_destroy(plan);
    return rc;
}

int lsio_cuda_pull_strip_crc32c(int device, const lsio_cuda_segment* segs, size_t n, uint32_t* crc_out_dev, void* stream) {
    using namespace lsio;
    static_assert(LSIO_CUDA_PULL_MAX_SEGMENTS == kPullMaxSegs, "header and kernel disagree");
    if (!segs || n == 0 || n > kPullMaxSegs) return -EINVAL;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    lsio_cuda_segment dev_view[kPullMaxSegs];
    uint64_t bytes = 0;
    for (size_t i = 0; i < n; i++) {
        void* alias = nullptr;
        LSIO_CUDA_TRY(cudaHostGetDevicePointer(&alias, (void*)(uintptr_t)segs[i].src, 0));  // fails for memory that is not pinned + mapped
        dev_view[i] = lsio_cuda_segment{(uint64_t)(uintptr_t)alias, segs[i].dst, segs[i].len};
        bytes += segs[i].len;
    }
    void* scratch = nullptr;
    const size_t sb = pack_pull_scratch_bytes(bytes + 64 * 32768);
    LSIO_CUDA_TRY(cudaMalloc(&scratch, sb));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = (int)cudaMemsetAsync(scratch, 0, sb, st);
    if (!rc) rc = pack_pull_launch(device, dev_view, (uint32_t)n, crc_out_dev != nullptr, scratch, pack_pull_max_tiles(bytes + 64 * 32768), crc_out_dev, st);
    const cudaError_t se = cudaStreamSynchronize(st);
    cudaFree(scratch);
    if (rc) return rc;
    return se == cudaSuccess ? 0 : cuda_fail(se, "pull_kernel", __FILE__, __LINE__);
}

uint64_t lsio_cuda_launch_count(void) { return lsio::launch_counter().load(std::memory_order_relaxed); }
const char* lsio_cuda_last_error(void) { return lsio::last_error_text().c_str(); }

}  // extern "C"
