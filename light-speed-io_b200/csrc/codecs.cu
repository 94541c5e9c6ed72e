// codecs.cu -- sm_100a kernels of the codec layer (include/lsio_codecs.h): chunks landed in HBM by the
// I/O path are decoded in HBM.  The reference plans this layer (planned_design.md:83-90 `lsio_codecs`,
// README.md:68-70) and holds no code for it; formats follow the published LZ4 block format and
// numcodecs.Shuffle.  Product code: no oracle, no CPU fallback.
//
//   lz4_kernel        one team of 8 (or 32) lanes per block.  LZ4 is a serial token stream, so the parallelism
//                     is (a) across blocks and (b) inside each copy: every lane of the team parses the same
//                     tokens from a shared-memory window of the compressed input (refilled with coalesced
//                     16-byte loads), literals and matches are copied by the team's lanes.  A match that overlaps its own output
//                     (offset < length) is periodic, so byte i is out[op - offset + i % offset]: every byte
//                     of a match is independent of the others and only depends on bytes before `op`.
//   shuffle_kernel    byte (un)shuffle as a tiled gather: each thread assembles one 16-byte output vector from
//                     16/E bytes of each of the E byte planes (register transpose, PRMT), warp accesses are
//                     contiguous in every plane and in the output.  HBM-bound: 2 bytes moved per byte.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.hpp"
#include "lsio_codecs.h"

namespace lsio {
namespace {

// ---------------------------------------------------------------------------------------------------------
// LZ4 block decode
// ---------------------------------------------------------------------------------------------------------
// Real chunks decode to very short sequences (measured on shuffled float32 data: 7.7 output bytes per
// sequence, 1.6 literal + 6 match bytes), so the kernel is bound by instructions ISSUED per sequence, not by
// bytes: a TEAM of T lanes (8 by default, four blocks per warp) decodes one block, every lane of the team
// parsing the same tokens.  T = 32 is kept for launches with few blocks.
constexpr int kLzThreads = 256;
constexpr int kLzWin = 512;  // bytes of compressed input staged per team

template <int T>
struct LzIn {
    const uint8_t* src;  // compressed bytes
    uint32_t n;          // their count
    uint32_t lead;       // src - (src rounded down to 16): window coordinates are relative to the aligned base
    uint8_t* win;        // this team's shared-memory window
    uint32_t ws;         // window start (multiple of 16, aligned-base coordinates)
    unsigned mask;       // the team's lanes
    int l;               // lane within the team
};

// Stage [ws, ws + kLzWin) of the aligned-base coordinates with coalesced 16-byte loads.
// Out of line and by value: keeps the per-sequence loop compact and the caller's state in registers.
template <int T>
__device__ __noinline__ void lz_fill_window(const uint8_t* abase, uint32_t lead, uint32_t n, uint8_t* win, uint32_t ws, unsigned mask, int l) {
    __syncwarp(mask);
#pragma unroll
    for (int r = 0; r < kLzWin / 16 / T; r++) {
        const uint32_t voff = ws + 16u * (uint32_t)(l + T * r);  // aligned-base coordinate of this vector
        uint4 v = make_uint4(0, 0, 0, 0);
        if (voff >= lead && voff + 16 <= lead + n) {
            v = *reinterpret_cast<const uint4*>(abase + voff);
        } else if (voff + 16 > lead && voff < lead + n) {  // edge vector: only bytes inside [src, src+n)
            uint32_t w[4] = {0, 0, 0, 0};
            for (int k = 0; k < 16; k++)
                if (voff + k >= lead && voff + k < lead + n) w[k >> 2] |= (uint32_t)abase[voff + k] << (8 * (k & 3));
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *reinterpret_cast<uint4*>(win + 16 * (l + T * r)) = v;
    }
    __syncwarp(mask);
}
template <int T>
__device__ __forceinline__ void lz_fill(LzIn<T>& in) {
    lz_fill_window<T>(in.src - in.lead, in.lead, in.n, in.win, in.ws, in.mask, in.l);
}
// Byte `ip` of the compressed stream (uniform across the team), refilling the window when it has moved on.
template <int T>
__device__ __forceinline__ uint32_t lz_byte(LzIn<T>& in, uint32_t ip) {
    const uint32_t p = ip + in.lead;
    if (p - in.ws >= (uint32_t)kLzWin) {
        in.ws = p & ~15u;
        lz_fill(in);
    }
    return in.win[p - in.ws];
}

// Cooperative copy of a long run (literals from the compressed stream, or a match that does not overlap its
// own output): destination-aligned 32-bit stores, source words realigned with a funnel shift.
template <int T>
__device__ __forceinline__ void lz_copy_long(uint8_t* d, const uint8_t* s, uint32_t len, int l) {
    const uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(d) & 3)) & 3u;  // len >= 64 > head
    if ((uint32_t)l < head) d[l] = s[l];
    d += head;
    s += head;
    len -= head;
    const uint32_t words = len >> 2, sh = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
    uint32_t* dw = reinterpret_cast<uint32_t*>(d);
    if (sh == 0) {
        for (uint32_t i = l; i < words; i += T) dw[i] = sw[i];
    } else {
        // sw[i + 1] is the aligned word that holds the last byte needed: never beyond the word of a valid byte
        for (uint32_t i = l; i < words; i += T) dw[i] = __funnelshift_r(sw[i], sw[i + 1], sh);
    }
    const uint32_t done = words << 2;
    if ((uint32_t)l < len - done) d[done + l] = s[done + l];
}

// One block advances by one sequence per ~900 cycles whatever the team size (measured: 16 MB/s per block on
// shuffled float32 data; a dependent chain of ~80 instructions and four shared-memory reads per sequence), so
// throughput comes from blocks in flight.  Tried and rejected: mirroring the last 16/32 KiB of output in shared
// memory so that matches never wait for L2 -- no gain per block (4.0 -> 4.2 GB/s at 256 blocks of 1 MiB; the
// match reads are not on the critical path), and the lost occupancy cost 2-5x at larger launches.
// Also tried and rejected: deferring a short match's stores by one sequence so that its loads overlap the next
// parse (software pipelining; +7 % at 16,384 blocks, but 8 more live registers: -20 % wherever occupancy counts).
// kMinCtas: 8 CTAs per SM (32 registers, a few spilled bytes) pays once a launch has enough blocks to fill them.
template <int T, int kMinCtas>
__global__ void __launch_bounds__(kLzThreads, kMinCtas) lz4_kernel(const lsio_codec_item* __restrict__ items, uint32_t n_items,
                                                                   lsio_codec_result* __restrict__ results) {
    constexpr int kTeams = kLzThreads / T;
    __shared__ __align__(16) uint8_t s_lz[kTeams * kLzWin];
    const int team = threadIdx.x / T, l = threadIdx.x % T;
    const uint32_t idx = blockIdx.x * kTeams + team;
    if (idx >= n_items) return;  // whole teams leave: every later sync is masked to one team
    const lsio_codec_item it = items[idx];
    LzIn<T> in;
    in.src = reinterpret_cast<const uint8_t*>(it.src);
    in.n = it.src_len;
    in.lead = (uint32_t)(it.src & 15);
    in.win = s_lz + team * kLzWin;
    in.ws = 0;
    in.mask = T == 32 ? 0xFFFFFFFFu : (((1u << (T & 31)) - 1u) << ((threadIdx.x & 31) / T * T));
    in.l = l;
    uint8_t* dst = reinterpret_cast<uint8_t*>(it.dst);
    __builtin_assume(__isGlobal(dst));  // LDG/STG instead of generic loads and stores
    __builtin_assume(__isGlobal(in.src));
    const uint32_t n = it.src_len, cap = it.dst_cap;
    uint32_t ip = 0, op = 0;
    int status = LSIO_CODEC_OK;
    if (n == 0) status = LSIO_CODEC_ERR_INPUT;
    else lz_fill(in);
    while (status == LSIO_CODEC_OK) {
        if (ip >= n) { status = LSIO_CODEC_ERR_INPUT; break; }
        // one check per sequence: token + up to 14 literals + offset + one length byte inside the window
        uint32_t q = ip + in.lead - in.ws;
        if (q + 20 > (uint32_t)kLzWin) {
            in.ws = (ip + in.lead) & ~15u;
            lz_fill(in);
            q = ip + in.lead - in.ws;
        }
        const uint8_t* w = in.win + q;
        const uint32_t token = w[0];
        uint32_t lit = token >> 4, ml = token & 15, offset;
        if (lit < 15) {  // the common case: everything up to the offset sits in the window
            if (lit > n - ip - 1) { status = LSIO_CODEC_ERR_INPUT; break; }
            if (lit > cap - op) { status = LSIO_CODEC_ERR_OUTPUT; break; }
            if ((uint32_t)l < lit) dst[op + l] = w[1 + l];
            if (T < 16 && (uint32_t)l + T < lit) dst[op + l + T] = w[1 + l + T];
            ip += 1 + lit;
            op += lit;
            if (ip == n) break;  // the last sequence has literals only
            if (n - ip < 2) { status = LSIO_CODEC_ERR_INPUT; break; }
            offset = w[1 + lit] | (uint32_t)w[2 + lit] << 8;
            ip += 2;
        } else {
            ip++;
            uint32_t b;
            do {
                if (ip >= n) { status = LSIO_CODEC_ERR_INPUT; break; }
                b = lz_byte(in, ip++);
                lit += b;
            } while (b == 255);
            if (status) break;
            if (lit > n - ip) { status = LSIO_CODEC_ERR_INPUT; break; }
            if (lit > cap - op) { status = LSIO_CODEC_ERR_OUTPUT; break; }
            if (lit >= 64) {
                lz_copy_long<T>(dst + op, in.src + ip, lit, l);
            } else {
                const uint8_t* s = in.src + ip;
                for (uint32_t i = l; i < lit; i += T) dst[op + i] = s[i];
            }
            ip += lit;
            op += lit;
            if (ip == n) break;
            if (n - ip < 2) { status = LSIO_CODEC_ERR_INPUT; break; }
            offset = lz_byte(in, ip);
            offset |= lz_byte(in, ip + 1) << 8;
            ip += 2;
        }
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
        __syncwarp(in.mask);  // the literals above (and everything before them) are visible to the team's match reads
        const uint8_t* m = dst + op - offset;
        if (offset >= ml) {
            if (ml >= 64) {
                lz_copy_long<T>(dst + op, m, ml, l);
            } else {
                for (uint32_t i = l; i < ml; i += T) dst[op + i] = m[i];
            }
        } else {  // a match overlapping its own output is the periodic extension of the last `offset` bytes
            for (uint32_t i = l; i < ml; i += T) dst[op + i] = m[i % offset];
        }
        op += ml;
        __syncwarp(in.mask);
    }
    if (l == 0) {
        results[idx].out_len = op;
        results[idx].status = status;
    }
}

// ---------------------------------------------------------------------------------------------------------
// LZ4, launches with few blocks: parse in parallel, copy in parallel
// ---------------------------------------------------------------------------------------------------------
// With few blocks in flight the kernel above is bound by the latency of ONE block: ~900 cycles per sequence,
// because every sequence waits for its own copies (a store, a team sync, a load of bytes that were just written
// and therefore come from L2, a store, a sync) and because a single warp retires a dependent instruction only
// every few cycles (a plain serial parse alone measured ~350 cycles per sequence).  This kernel gives a warp one
// block and works on the next 32 bytes of the compressed stream at a time:
//   parse   lane j assumes a token at byte ip + j and decodes it in parallel with the others (literal length,
//           match length, offset, where the next token would be); the real tokens are then found by following the
//           "next" links from lane 0 with one warp shuffle per sequence, and a warp prefix sum over the real lanes
//           gives every sequence its output position.  Tokens with length extensions, the end of the block and
//           anything invalid fall back to a fully checked serial parse of one sequence.
//   copy    every real lane copies its own literals and its own match, all loads before all stores, so the batch
//           pays one memory latency instead of one per sequence.  A match may read bytes that another match of the
//           same batch produces; matches therefore retire in rounds: in each round every pending match whose source
//           ends at or below the destination of the FIRST pending one is ready (all bytes below that point are
//           final), the rest wait for the next round.  Runs longer than kWideBytes are copied by the whole warp.
// Measured: 2x per block (32 instead of 16 MB/s; 1,024 blocks of 1 MiB decode at 32 GB/s); what remains is the copy
// phase's two to three dependent L2 round trips per batch of ~7 sequences.
constexpr int kWideBytes = 18;  // bytes one lane moves per sequence and round trip (a match without length extension is <= 18)
constexpr int kWideLits = 14;   // literals of a token without length extension

// Lane-private copy of n <= N bytes in two steps, so that a lane can issue the loads of all its bytes (and of several
// sequences) before the first store.  The copy phase is bound by the instructions of these unrolled loops (mirroring the
// batch in shared memory to save L2 round trips halved the speed), so they are kept minimal: no wrap handling here,
// N = 14 for literals.
template <int N>
__device__ __forceinline__ void lane_load(uint32_t (&v)[kWideBytes], const uint8_t* s, uint32_t n) {
#pragma unroll
    for (int i = 0; i < N; i++)
        if ((uint32_t)i < n) v[i] = s[i];
}
// A match overlapping its own output: the source is the `period` bytes at s, repeated.
__device__ __forceinline__ void lane_load_periodic(uint32_t (&v)[kWideBytes], const uint8_t* s, uint32_t n, uint32_t period) {
    uint32_t j = 0;
#pragma unroll
    for (int i = 0; i < kWideBytes; i++) {
        if ((uint32_t)i < n) v[i] = s[j];
        if (++j == period) j = 0;
    }
}
template <int N>
__device__ __forceinline__ void lane_store(uint8_t* d, const uint32_t (&v)[kWideBytes], uint32_t n) {
#pragma unroll
    for (int i = 0; i < N; i++)
        if ((uint32_t)i < n) d[i] = (uint8_t)v[i];
}
// Whole-warp copy of one long run (uniform arguments).
__device__ __forceinline__ void warp_copy(uint8_t* d, const uint8_t* s, uint32_t n, uint32_t period, int lane) {
    if (period) {
        for (uint32_t i = lane; i < n; i += 32) d[i] = s[i % period];
    } else if (n >= 64) {
        lz_copy_long<32>(d, s, n, lane);
    } else {
        for (uint32_t i = lane; i < n; i += 32) d[i] = s[i];
    }
}

// H: token positions examined per batch = 32 * H (lane j looks at positions j, j + 32, ...).  H = 2 spreads the memory
// round trips of the copy phase over twice the sequences but needs more rounds per batch (longer dependence chains):
// measured slower (24.6 vs 31.6 GB/s at 1,024 blocks of 1 MiB), so H = 1 is what launches use.
template <int H>
__global__ void __launch_bounds__(kLzThreads) lz4_wide_kernel(const lsio_codec_item* __restrict__ items, uint32_t n_items,
                                                              lsio_codec_result* __restrict__ results) {
    constexpr int T = 32, kTeams = kLzThreads / T;
    constexpr unsigned kAll = 0xFFFFFFFFu;
    constexpr uint32_t kSpan = 32 * H;       // token positions per batch
    constexpr uint32_t kNeed = kSpan + 20;   // compressed bytes a batch may look at: the last position's token, 14 literals, offset
    __shared__ __align__(16) uint8_t s_lz[kTeams * kLzWin];
    const int team = threadIdx.x / T, lane = threadIdx.x % T;
    const uint32_t idx = blockIdx.x * kTeams + team;
    if (idx >= n_items) return;
    const lsio_codec_item it = items[idx];
    LzIn<T> in;
    in.src = reinterpret_cast<const uint8_t*>(it.src);
    in.n = it.src_len;
    in.lead = (uint32_t)(it.src & 15);
    in.win = s_lz + team * kLzWin;
    in.ws = 0;
    in.mask = kAll;
    in.l = lane;
    uint8_t* dst = reinterpret_cast<uint8_t*>(it.dst);
    __builtin_assume(__isGlobal(dst));
    __builtin_assume(__isGlobal(in.src));
    const uint32_t n = it.src_len, cap = it.dst_cap;
    uint32_t ip = 0, op = 0;
    int status = LSIO_CODEC_OK;
    bool done = false;
    if (n == 0) status = LSIO_CODEC_ERR_INPUT;
    else lz_fill(in);
    while (status == LSIO_CODEC_OK && !done) {
        // what this lane copies in this batch, per slot h: literals [lit_ip, +lit) -> dst + lop, then a match of ml bytes
        uint32_t my_lit_ip[H], my_lit[H], my_lop[H], my_off[H], my_ml[H];
        unsigned real[H];  // real[h] bit j: slot h of lane j holds a sequence
#pragma unroll
        for (int h = 0; h < H; h++) my_lit_ip[h] = my_lit[h] = my_lop[h] = my_off[h] = my_ml[h] = 0, real[h] = 0;
        bool any = false;
        // ---- parse, in parallel: slot h of lane j decodes the token that would start at ip + 32 h + j ----
        if (n - ip >= kNeed) {  // far enough from the end: every field looked at is inside the block
            uint32_t q = ip + in.lead - in.ws;
            if (q + kNeed > (uint32_t)kLzWin) {
                in.ws = (ip + in.lead) & ~15u;
                lz_fill(in);
                q = ip + in.lead - in.ws;
            }
            uint32_t lit[H], mlc[H], offset[H], nxt[H];
#pragma unroll
            for (int h = 0; h < H; h++) {
                const uint8_t* w = in.win + q + 32 * h + lane;
                const uint32_t token = w[0];
                lit[h] = token >> 4;
                mlc[h] = token & 15;
                offset[h] = w[1 + lit[h]] | (uint32_t)w[2 + lit[h]] << 8;
                // where the next token would start; 0: length extensions, not handled here
                nxt[h] = (lit[h] < 15 && mlc[h] < 15) ? 32u * h + (uint32_t)lane + 3 + lit[h] : 0;
            }
            uint32_t cur = 0;
            while (cur < kSpan) {  // follow the links from the token at ip (cur is uniform)
                uint32_t nx = 0;
#pragma unroll
                for (int h = 0; h < H; h++)
                    if ((cur >> 5) == (uint32_t)h) nx = __shfl_sync(kAll, nxt[h], cur & 31);
                if (nx == 0) break;
#pragma unroll
                for (int h = 0; h < H; h++)
                    if ((cur >> 5) == (uint32_t)h) real[h] |= 1u << (cur & 31);
                cur = nx;
            }
            // output positions: prefix sum of the output bytes of the real sequences, in position order
            uint32_t v[H], incl[H], carry = 0;
            unsigned badmask[H];
#pragma unroll
            for (int h = 0; h < H; h++) {
                v[h] = ((real[h] >> lane) & 1u) ? lit[h] + mlc[h] + 4 : 0;
                uint32_t x = v[h];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t up = __shfl_up_sync(kAll, x, d);
                    if (lane >= d) x += up;
                }
                incl[h] = x + carry;
                carry = __shfl_sync(kAll, incl[h], 31);
                const uint32_t mop = op + incl[h] - v[h] + lit[h];
                // a sequence that does not fit the output or whose offset reaches before the output ends the batch:
                // the checked parse below takes it (and reports the error if it is one)
                badmask[h] = __ballot_sync(kAll, ((real[h] >> lane) & 1u) && (incl[h] > cap - op || offset[h] == 0 || offset[h] > mop));
            }
            bool cut = false;
#pragma unroll
            for (int h = 0; h < H; h++) {
                if (cut) real[h] = 0;
                else if (badmask[h]) {
                    const uint32_t pos = (uint32_t)__ffs(badmask[h]) - 1;
                    real[h] &= (1u << pos) - 1u;
                    cur = 32u * h + pos;
                    cut = true;
                }
            }
            uint32_t total = 0;
#pragma unroll
            for (int h = 0; h < H; h++) {
                if (real[h]) {
                    any = true;
                    total = __shfl_sync(kAll, incl[h], 31 - __clz(real[h]));  // the batch ends after its last real sequence
                    if ((real[h] >> lane) & 1u) {
                        my_lit_ip[h] = ip + 32u * h + lane + 1;
                        my_lit[h] = lit[h];
                        my_lop[h] = op + incl[h] - v[h];
                        my_off[h] = offset[h];
                        my_ml[h] = mlc[h] + 4;
                    }
                }
            }
            if (any) {
                op += total;
                ip += cur;  // the first token not taken: the bad one, or where the last link points
            }
        }
        if (!any) {
            // ---- parse, serially: ONE sequence, fully checked (length extensions, the end of the block, errors) ----
            if (ip >= n) { status = LSIO_CODEC_ERR_INPUT; break; }
            uint32_t q = ip + in.lead - in.ws;
            if (q + 20 > (uint32_t)kLzWin) {
                in.ws = (ip + in.lead) & ~15u;
                lz_fill(in);
                q = ip + in.lead - in.ws;
            }
            const uint8_t* w = in.win + q;
            const uint32_t token = w[0];
            uint32_t lit = token >> 4, ml = token & 15, offset = 0, lit_ip = 0;
            bool last = false, have = true;
            if (lit < 15) {  // everything up to the offset sits in the window
                if (lit > n - ip - 1) { status = LSIO_CODEC_ERR_INPUT; break; }
                if (lit > cap - op) { status = LSIO_CODEC_ERR_OUTPUT; break; }
                lit_ip = ip + 1;
                ip += 1 + lit;
                if (ip == n) last = true;
                else if (n - ip < 2) status = LSIO_CODEC_ERR_INPUT;  // the literals are still valid output
                else {
                    offset = w[1 + lit] | (uint32_t)w[2 + lit] << 8;
                    ip += 2;
                }
            } else {
                ip++;
                uint32_t b = 0;
                do {
                    if (ip >= n) { status = LSIO_CODEC_ERR_INPUT; have = false; break; }
                    b = lz_byte(in, ip++);
                    lit += b;
                } while (b == 255);
                if (have && lit > n - ip) { status = LSIO_CODEC_ERR_INPUT; have = false; }
                if (have && lit > cap - op) { status = LSIO_CODEC_ERR_OUTPUT; have = false; }
                if (have) {
                    lit_ip = ip;
                    ip += lit;
                    if (ip == n) last = true;
                    else if (n - ip < 2) status = LSIO_CODEC_ERR_INPUT;
                    else {
                        offset = lz_byte(in, ip);
                        offset |= lz_byte(in, ip + 1) << 8;
                        ip += 2;
                    }
                }
            }
            if (!have) break;
            const uint32_t lop = op;
            op += lit;
            if (!last && status == LSIO_CODEC_OK) {
                if (ml == 15) {
                    uint32_t b;
                    do {
                        if (ip >= n) { status = LSIO_CODEC_ERR_INPUT; break; }
                        b = lz_byte(in, ip++);
                        ml += b;
                    } while (b == 255);
                }
                ml += 4;
                if (status == LSIO_CODEC_OK && (offset == 0 || offset > op)) status = LSIO_CODEC_ERR_OFFSET;
                if (status == LSIO_CODEC_OK && ml > cap - op) status = LSIO_CODEC_ERR_OUTPUT;
            }
            if (last || status != LSIO_CODEC_OK) ml = 0;  // literals only: the end of the block, or a match that cannot be used
            if (lane == 0) {
                my_lit_ip[0] = lit_ip;
                my_lit[0] = lit;
                my_lop[0] = lop;
                my_off[0] = offset;
                my_ml[0] = ml;
            }
            real[0] = 1;
            op += ml;
            if (last) done = true;
        }
        // ---- copy: literals (no dependencies), then matches in rounds ----
        uint32_t buf[H][kWideBytes];
        bool act[H];
#pragma unroll
        for (int h = 0; h < H; h++) {
            act[h] = (real[h] >> lane) & 1u;
            if (act[h] && my_lit[h] && my_lit[h] <= (uint32_t)kWideLits) lane_load<kWideLits>(buf[h], in.src + my_lit_ip[h], my_lit[h]);
        }
#pragma unroll
        for (int h = 0; h < H; h++) {
            if (act[h] && my_lit[h] && my_lit[h] <= (uint32_t)kWideLits) lane_store<kWideLits>(dst + my_lop[h], buf[h], my_lit[h]);
            for (unsigned m = __ballot_sync(kAll, act[h] && my_lit[h] > (uint32_t)kWideLits); m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                warp_copy(dst + __shfl_sync(kAll, my_lop[h], j), in.src + __shfl_sync(kAll, my_lit_ip[h], j), __shfl_sync(kAll, my_lit[h], j), 0, lane);
            }
        }
        __syncwarp();
        uint32_t my_mop[H], my_src_hi[H];
        unsigned pending[H], any_pending = 0;
#pragma unroll
        for (int h = 0; h < H; h++) {
            my_mop[h] = my_lop[h] + my_lit[h];  // where this slot's match goes
            my_src_hi[h] = my_mop[h] - my_off[h] + (my_ml[h] < my_off[h] ? my_ml[h] : my_off[h]);
            pending[h] = __ballot_sync(kAll, act[h] && my_ml[h] != 0);
            any_pending |= pending[h];
        }
        while (any_pending) {
            // every byte below the destination of the first pending match (in position order) is final
            uint32_t floor_op = 0;
            bool found = false;
#pragma unroll
            for (int h = 0; h < H; h++)
                if (!found && pending[h]) {
                    floor_op = __shfl_sync(kAll, my_mop[h], __ffs(pending[h]) - 1);
                    found = true;
                }
            bool ready[H];
#pragma unroll
            for (int h = 0; h < H; h++) {
                ready[h] = ((pending[h] >> lane) & 1u) && my_src_hi[h] <= floor_op;
                if (ready[h] && my_ml[h] <= (uint32_t)kWideBytes) {
                    if (my_off[h] < my_ml[h]) lane_load_periodic(buf[h], dst + my_mop[h] - my_off[h], my_ml[h], my_off[h]);
                    else lane_load<kWideBytes>(buf[h], dst + my_mop[h] - my_off[h], my_ml[h]);
                }
            }
            any_pending = 0;
#pragma unroll
            for (int h = 0; h < H; h++) {
                if (ready[h] && my_ml[h] <= (uint32_t)kWideBytes) lane_store<kWideBytes>(dst + my_mop[h], buf[h], my_ml[h]);
                for (unsigned m = __ballot_sync(kAll, ready[h] && my_ml[h] > (uint32_t)kWideBytes); m; m &= m - 1) {
                    const int j = __ffs(m) - 1;
                    const uint32_t mop = __shfl_sync(kAll, my_mop[h], j), off = __shfl_sync(kAll, my_off[h], j), ml = __shfl_sync(kAll, my_ml[h], j);
                    warp_copy(dst + mop, dst + mop - off, ml, off < ml ? off : 0, lane);
                }
                pending[h] &= ~__ballot_sync(kAll, ready[h]);
                any_pending |= pending[h];
            }
            __syncwarp();
        }
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

template <int V> struct UIntOf;
template <> struct UIntOf<4> { using T = uint32_t; };
template <> struct UIntOf<8> { using T = uint64_t; };

// One GROUP = V consecutive elements: V bytes of each of the E planes on the planar side, E*V contiguous bytes
// (NV 16-byte vectors) on the other.  (E, V) = (2, 8), (4, 4): one vector per group; (8, 4): two; (16, 4): four --
// plane accesses are never narrower than 32 bits (2-byte plane loads made E = 8 LSU-bound: 58-77 % of HBM peak).
template <int E, int V>
__device__ __forceinline__ void unshuffle_group(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t cnt, uint32_t g) {
    using T = typename UIntOf<V>::T;
    constexpr int NV = E * V / 16;
    uint64_t pl[E];
#pragma unroll
    for (int j = 0; j < E; j++) pl[j] = (uint64_t)__ldg(reinterpret_cast<const T*>(src + (size_t)j * cnt + (size_t)g * V));
    uint32_t w[NV * 4];
#pragma unroll
    for (int k = 0; k < NV * 4; k++) w[k] = 0;
#pragma unroll
    for (int k = 0; k < E * V; k++) {
        const int i = k / E, j = k % E;  // contiguous byte k = byte j of element i of the group = byte i of plane j's word
        w[k / 4] |= ((uint32_t)(pl[j] >> (8 * i)) & 255u) << (8 * (k % 4));
    }
    uint4* o = reinterpret_cast<uint4*>(dst + (size_t)g * (E * V));
#pragma unroll
    for (int v = 0; v < NV; v++) o[v] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
}
template <int E, int V>
__device__ __forceinline__ void shuffle_group(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t cnt, uint32_t g) {
    using T = typename UIntOf<V>::T;
    constexpr int NV = E * V / 16;
    uint32_t w[NV * 4];
    const uint4* in = reinterpret_cast<const uint4*>(src + (size_t)g * (E * V));
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const uint4 x = __ldg(in + v);
        w[4 * v] = x.x, w[4 * v + 1] = x.y, w[4 * v + 2] = x.z, w[4 * v + 3] = x.w;
    }
    uint64_t pl[E];
#pragma unroll
    for (int j = 0; j < E; j++) pl[j] = 0;
#pragma unroll
    for (int k = 0; k < E * V; k++) {
        const int i = k / E, j = k % E;
        pl[j] |= (uint64_t)((w[k / 4] >> (8 * (k % 4))) & 255u) << (8 * i);
    }
#pragma unroll
    for (int j = 0; j < E; j++) *reinterpret_cast<T*>(dst + (size_t)j * cnt + (size_t)g * V) = (T)pl[j];
}

template <int E, int V, bool kForward>
__device__ __forceinline__ void shuffle_tile_fast(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t cnt, uint32_t g0, uint32_t g1) {
    constexpr int U = 64 / (E * V) > 1 ? 64 / (E * V) : 1;  // groups per thread per pass: 64 bytes of each side in flight
    uint32_t g = g0 + threadIdx.x;
    for (; g + (U - 1) * kShufThreads < g1; g += U * kShufThreads) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (kForward) shuffle_group<E, V>(src, dst, cnt, g + u * kShufThreads);
            else unshuffle_group<E, V>(src, dst, cnt, g + u * kShufThreads);
        }
    }
    for (; g < g1; g += kShufThreads) {
        if (kForward) shuffle_group<E, V>(src, dst, cnt, g);
        else unshuffle_group<E, V>(src, dst, cnt, g);
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

// tile_map[blockIdx.x] = {item, tile within the item}: one load instead of a search over the items' prefix sums
// (a 17-step binary search per CTA cost 25-30 % of the HBM roofline at 16 KiB chunks).
template <bool kForward>
__global__ void __launch_bounds__(kShufThreads) shuffle_kernel(const lsio_codec_item* __restrict__ items, const uint2* __restrict__ tile_map,
                                                               uint32_t elem) {
    const uint2 tm = __ldg(tile_map + blockIdx.x);
    const lsio_codec_item it = items[tm.x];
    const uint8_t* src = reinterpret_cast<const uint8_t*>(it.src);
    uint8_t* dst = reinterpret_cast<uint8_t*>(it.dst);
    __builtin_assume(__isGlobal(src));
    __builtin_assume(__isGlobal(dst));
    const uint32_t len = it.src_len, cnt = elem ? len / elem : 0;
    const uint32_t b0 = tm.y * kShufTile, b1 = min(len, b0 + kShufTile);
    // vector path: whole groups of the contiguous side that lie inside the cnt*elem body, when every plane access
    // is naturally aligned
    const uint32_t V = elem == 2 ? 8 : (elem == 4 || elem == 8 || elem == 16) ? 4 : 0;
    const uint64_t contig = kForward ? it.src : it.dst, planar = kForward ? it.dst : it.src;
    uint32_t vec_end = b0;  // bytes [b0, vec_end) go through the vector path
    if (V && (contig & 15) == 0 && (planar % V) == 0 && (cnt % V) == 0) {
        const uint32_t gb = elem * V;  // bytes per group: 16, 16, 32, 64 (all divide the tile size)
        vec_end = min(b1, cnt * elem / gb * gb);
        if (vec_end < b0) vec_end = b0;
        const uint32_t g0 = b0 / gb, g1 = vec_end / gb;
        switch (elem) {
            case 2: shuffle_tile_fast<2, 8, kForward>(src, dst, cnt, g0, g1); break;
            case 4: shuffle_tile_fast<4, 4, kForward>(src, dst, cnt, g0, g1); break;
            case 8: shuffle_tile_fast<8, 4, kForward>(src, dst, cnt, g0, g1); break;
            default: shuffle_tile_fast<16, 4, kForward>(src, dst, cnt, g0, g1); break;
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

namespace {
// A cudaMallocAsync block that is returned to the pool on every path out of the synchronous helpers.
struct AsyncBlock {
    void* p;
    cudaStream_t st;
    ~AsyncBlock() {
        if (p) cudaFreeAsync(p, st);
    }
};
}  // namespace

extern "C" {

int lsio_codecs_lz4_decompress_async(int device, const lsio_codec_item* items_dev, size_t n, lsio_codec_result* results_dev, void* stream) {
    if ((!items_dev || !results_dev) && n) return -EINVAL;
    if (n > 0xFFFFFFF0u) return -E2BIG;
    if (!n) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    // Many blocks: four blocks per warp (4x fewer instructions issued per sequence, which is what bounds the kernel
    // then).  Few blocks: a warp per block (more warps in flight, 4x wider copies).  Crossover measured between
    // 4096 and 8192 blocks of 64 KiB (tools/codec_bench.py).
    // Measured on shuffled float32 chunks of 64 KiB (GB/s decoded): 16,384 blocks 122 (5 CTAs/SM) vs 116 (8 CTAs/SM);
    // 32,768 blocks 174 with 8; 65,536 blocks of 16 KiB 167 vs 193.
    // Below ~6k blocks the parallel-parse kernel replaces a warp per block (GB/s decoded, shuffled float32: 8.9 vs 4.2 at
    // 256 x 1 MiB, 34.6 vs 16.4 at 1,024 x 1 MiB, 53.4 vs 31.8 at 2,048 x 64 KiB, 56.0 vs 53.1 at 4,096 x 64 KiB); the
    // warp-per-block variant stays selectable (32).
    int variant = n > 24576 ? 88 : n > 6144 ? 8 : 1;  // 88 / 8: teams of 8 at 8 / 5 CTAs per SM; 1 (2): parallel parse, 32 (64) positions per batch; 32: a warp per block
    static const int forced = [] { const char* e = getenv("LSIO_LZ4_VARIANT"); return e ? atoi(e) : 0; }();  // measurement knob (tools/codec_bench.py), read once
    if (forced) variant = forced;
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == 88) lz4_kernel<8, 8><<<(unsigned)((n + 31) / 32), kLzThreads, 0, st>>>(items_dev, (uint32_t)n, results_dev);
    else if (variant == 8) lz4_kernel<8, 5><<<(unsigned)((n + 31) / 32), kLzThreads, 0, st>>>(items_dev, (uint32_t)n, results_dev);
    else if (variant == 1) lz4_wide_kernel<1><<<(unsigned)((n + 7) / 8), kLzThreads, 0, st>>>(items_dev, (uint32_t)n, results_dev);
    else if (variant == 2) lz4_wide_kernel<2><<<(unsigned)((n + 7) / 8), kLzThreads, 0, st>>>(items_dev, (uint32_t)n, results_dev);
    else lz4_kernel<32, 1><<<(unsigned)((n + 7) / 8), kLzThreads, 0, st>>>(items_dev, (uint32_t)n, results_dev);
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
    AsyncBlock guard{d, st};  // freed on every return path
    lsio_codec_item* d_items = reinterpret_cast<lsio_codec_item*>(d);
    lsio_codec_result* d_res = reinterpret_cast<lsio_codec_result*>(d + n * sizeof(lsio_codec_item));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d_items, items, n * sizeof(lsio_codec_item), cudaMemcpyHostToDevice, st));
    int rc = lsio_codecs_lz4_decompress_async(device, d_items, n, d_res, stream);
    if (rc) return rc;
    LSIO_CUDA_TRY(cudaMemcpyAsync(results, d_res, n * sizeof(lsio_codec_result), cudaMemcpyDeviceToHost, st));
    LSIO_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

uint32_t lsio_codecs_shuffle_tiles(uint32_t src_len) { return (uint32_t)(((uint64_t)src_len + kShufTile - 1) / kShufTile); }

int64_t lsio_codecs_shuffle_tile_map(const uint32_t* src_lens, size_t n, uint32_t* tile_map, size_t cap_tiles) {
    if (!src_lens && n) return -EINVAL;
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        const uint32_t t = lsio_codecs_shuffle_tiles(src_lens[i]);
        if (tile_map && total + t <= cap_tiles)
            for (uint32_t k = 0; k < t; k++) {
                tile_map[2 * (total + k)] = (uint32_t)i;
                tile_map[2 * (total + k) + 1] = k;
            }
        total += t;
    }
    if (total > 0x7FFFFFFFu) return -E2BIG;
    return (int64_t)total;
}

int lsio_codecs_unshuffle_async(int device, const lsio_codec_item* items_dev, const uint32_t* tile_map_dev, uint32_t n_tiles, uint32_t elem_size,
                                int forward, void* stream) {
    if ((!items_dev || !tile_map_dev) && n_tiles) return -EINVAL;
    if (elem_size == 0) return -EINVAL;
    if (!n_tiles) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    const uint2* tm = reinterpret_cast<const uint2*>(tile_map_dev);
    if (forward) shuffle_kernel<true><<<n_tiles, kShufThreads, 0, (cudaStream_t)stream>>>(items_dev, tm, elem_size);
    else shuffle_kernel<false><<<n_tiles, kShufThreads, 0, (cudaStream_t)stream>>>(items_dev, tm, elem_size);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    LSIO_CUDA_TRY(cudaGetLastError());
    return 0;
}

static int shuffle_host_items(int device, const lsio_codec_item* items, size_t n, uint32_t elem_size, int forward, void* stream) {
    if (!items && n) return -EINVAL;
    if (elem_size == 0) return -EINVAL;
    if (int rc = check_device(device)) return rc;
    if (!n) return 0;
    std::vector<uint32_t> lens(n);
    for (size_t i = 0; i < n; i++) lens[i] = items[i].src_len;
    const int64_t total = lsio_codecs_shuffle_tile_map(lens.data(), n, nullptr, 0);
    if (total < 0) return (int)total;
    if (total == 0) return 0;
    std::vector<uint32_t> map(2 * (size_t)total);
    lsio_codecs_shuffle_tile_map(lens.data(), n, map.data(), (size_t)total);
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    cudaStream_t st = (cudaStream_t)stream;
    char* d = nullptr;
    LSIO_CUDA_TRY(cudaMallocAsync((void**)&d, n * sizeof(lsio_codec_item) + map.size() * sizeof(uint32_t), st));
    AsyncBlock guard{d, st};  // freed on every return path
    lsio_codec_item* d_items = reinterpret_cast<lsio_codec_item*>(d);
    uint32_t* d_map = reinterpret_cast<uint32_t*>(d + n * sizeof(lsio_codec_item));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d_items, items, n * sizeof(lsio_codec_item), cudaMemcpyHostToDevice, st));
    LSIO_CUDA_TRY(cudaMemcpyAsync(d_map, map.data(), map.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    int rc = lsio_codecs_unshuffle_async(device, d_items, d_map, (uint32_t)total, elem_size, forward, stream);
    if (rc) return rc;
    LSIO_CUDA_TRY(cudaStreamSynchronize(st));  // `map` and `items` are pageable host memory: keep the call self-contained
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
