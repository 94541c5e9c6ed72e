/*
 * lsio_oracle.c -- CPU ORACLE (test infrastructure only; see lsio_oracle.h header note).
 * Each function cites the reference file:line it restates (relative to /root/reference).
 */
#define _GNU_SOURCE
#include "lsio_oracle.h"

#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#if defined(__SSE4_2__)
#include <nmmintrin.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Range arithmetic.  crates/lsio_uring/src/sqe.rs:59-122 (build_read_range_sqe) and
 * crates/lsio_uring/src/get_range.rs:23-30 (2 GiB guard uses Range::len(), which is 0 when
 * end < start, so relative ends skip the guard -- kept).
 * ---------------------------------------------------------------------------------------- */
int lsio_oracle_resolve_range(int64_t filesize, int64_t start, int64_t end, int strict,
                              lsio_oracle_range* out) {
    const int64_t ALIGN = LSIO_ORACLE_ALIGN;
    int64_t range_len = end > start ? end - start : 0; /* Range<isize>::len() */
    if (range_len > 2147479552LL) return -4;           /* get_range.rs:24 */
    int64_t start_off = start >= 0 ? start : filesize + start; /* sqe.rs:64-70 */
    if (start_off < 0) return -1;                               /* sqe.rs:71 */
    int64_t end_off = end >= 0 ? end : filesize + end + 1;      /* sqe.rs:73-80 */
    if (end_off < 0) return -2;                                 /* sqe.rs:81 */
    int64_t aligned_start = (start_off / ALIGN) * ALIGN;        /* sqe.rs:83 */
    int64_t buf_len = end_off - aligned_start;                  /* sqe.rs:87 */
    if (buf_len <= 0) return -3;                                /* sqe.rs:88 */
    if (strict && aligned_start != start_off) return -5;        /* sqe.rs:109-113 -> lib.rs:82-86 */
    out->start_off = start_off;
    out->end_off = end_off;
    out->aligned_start = aligned_start;
    out->buf_len = buf_len;
    out->padded_len = (int64_t)lsio_oracle_padded_size((size_t)buf_len, (size_t)ALIGN);
    out->view_start = start_off - aligned_start; /* sqe.rs:116 */
    out->view_end = end_off - aligned_start;     /* sqe.rs:117 */
    return 0;
}

/* crates/lsio_uring/src/user_data.rs:36-41 (Into<u64>) and :28-34 (From<u64>). */
uint64_t lsio_oracle_user_data_pack(uint32_t index_of_op, uint8_t opcode) {
    return ((uint64_t)index_of_op << 32) | (uint64_t)opcode;
}
void lsio_oracle_user_data_unpack(uint64_t v, uint32_t* index_of_op, uint8_t* opcode) {
    *index_of_op = (uint32_t)(v >> 32);
    *opcode = (uint8_t)(v & 0xFF);
}

/* ------------------------------------------------------------------------------------------
 * Tracker.  crates/lsio_uring/src/tracker.rs:3-68.  Vec<Option<T>> + VecDeque free list that
 * starts 0..n (tracker.rs:13), pop_front on get_next_index (:19), push_back on remove (:58).
 * ---------------------------------------------------------------------------------------- */
struct lsio_oracle_tracker {
    size_t n, len;
    uint64_t* ops;
    uint8_t* present;
    size_t* fifo; /* circular, capacity n */
    size_t fifo_head, fifo_len;
};
lsio_oracle_tracker* lsio_oracle_tracker_new(size_t n) {
    lsio_oracle_tracker* t = calloc(1, sizeof *t);
    t->n = n;
    t->ops = calloc(n, sizeof(uint64_t));
    t->present = calloc(n, 1);
    t->fifo = calloc(n, sizeof(size_t));
    for (size_t i = 0; i < n; i++) t->fifo[i] = i;
    t->fifo_len = n;
    return t;
}
void lsio_oracle_tracker_free(lsio_oracle_tracker* t) {
    if (!t) return;
    free(t->ops); free(t->present); free(t->fifo); free(t);
}
int64_t lsio_oracle_tracker_get_next_index(lsio_oracle_tracker* t) {
    if (t->fifo_len == 0) return -1;
    size_t idx = t->fifo[t->fifo_head];
    t->fifo_head = (t->fifo_head + 1) % t->n;
    t->fifo_len--;
    return (int64_t)idx;
}
void lsio_oracle_tracker_put(lsio_oracle_tracker* t, size_t idx, uint64_t op) {
    t->ops[idx] = op; t->present[idx] = 1; t->len++; /* tracker.rs:22-25 */
}
int lsio_oracle_tracker_get(lsio_oracle_tracker* t, size_t idx, uint64_t* op) {
    if (idx >= t->n) return -1; /* "index out of bounds" panic, tracker.rs:108-113 */
    if (!t->present[idx]) return 0;
    *op = t->ops[idx];
    return 1;
}
int lsio_oracle_tracker_remove(lsio_oracle_tracker* t, size_t idx, uint64_t* op) {
    int r = lsio_oracle_tracker_get(t, idx, op);
    if (r != 1) return r;
    t->fifo[(t->fifo_head + t->fifo_len) % t->n] = idx; /* tracker.rs:58 push_back */
    t->fifo_len++;
    t->len--;
    t->present[idx] = 0;
    return 1;
}
int lsio_oracle_tracker_is_empty(const lsio_oracle_tracker* t) { return t->len == 0; }
int lsio_oracle_tracker_is_full(const lsio_oracle_tracker* t) { return t->fifo_len == 0; }

/* crates/lsio_aligned_bytes/src/lib.rs:178-181: Layout::from_size_align(len, align).pad_to_align() */
size_t lsio_oracle_padded_size(size_t len, size_t align) {
    return (len + align - 1) / align * align;
}
/* crates/lsio_aligned_bytes/src/lib.rs:74-95 (error order: contains, zero, alignment). */
int lsio_oracle_split_to_check(size_t range_start, size_t range_end, size_t idx, size_t align) {
    if (!(idx >= range_start && idx < range_end)) return -1;
    if (idx == 0) return -2;
    if (idx % align != 0) return -3;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic content.  Position-keyed so any chunk can be proven from
 * (file_id, offset) alone; the reference's own pattern (lsio_bench main.rs:102) cannot.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t pattern_word(uint64_t file_id, uint64_t word_idx) {
    return splitmix64(LSIO_ORACLE_SEED ^ (file_id * 0x9E3779B97F4A7C15ULL) ^ word_idx);
}
void lsio_oracle_pattern_fill(uint64_t file_id, uint64_t off, uint8_t* dst, size_t len) {
    size_t i = 0;
    while (i < len && ((off + i) & 7)) {
        uint64_t o = off + i;
        dst[i++] = (uint8_t)(pattern_word(file_id, o >> 3) >> (8 * (o & 7)));
    }
    for (; i + 8 <= len; i += 8) {
        uint64_t w = pattern_word(file_id, (off + i) >> 3); /* little-endian byte order */
        memcpy(dst + i, &w, 8);
    }
    for (; i < len; i++) {
        uint64_t o = off + i;
        dst[i] = (uint8_t)(pattern_word(file_id, o >> 3) >> (8 * (o & 7)));
    }
}
void lsio_oracle_ref_pattern_fill(uint64_t off, uint8_t* dst, size_t len) {
    for (size_t i = 0; i < len; i++) dst[i] = (uint8_t)(off + i); /* lsio_bench main.rs:102 */
}
int lsio_oracle_write_file(const char* path, uint64_t file_id, uint64_t size, int ref_pattern) {
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return -errno;
    const size_t BUF = 4u << 20;
    uint8_t* buf = malloc(BUF);
    int rc = 0;
    for (uint64_t off = 0; off < size;) {
        size_t n = size - off < BUF ? (size_t)(size - off) : BUF;
        if (ref_pattern) lsio_oracle_ref_pattern_fill(off, buf, n);
        else lsio_oracle_pattern_fill(file_id, off, buf, n);
        size_t done = 0;
        while (done < n) {
            ssize_t w = write(fd, buf + done, n - done);
            if (w < 0) { if (errno == EINTR) continue; rc = -errno; goto out; }
            done += (size_t)w;
        }
        off += n;
    }
out:
    free(buf);
    if (close(fd) < 0 && rc == 0) rc = -errno;
    return rc;
}

/* Ground truth: file bytes [start_off, end_off), offsets by the sqe.rs:64-81 rule. */
int64_t lsio_oracle_pread_range(const char* path, int64_t start, int64_t end, uint8_t* dst,
                                size_t cap) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -errno;
    struct stat st;
    if (fstat(fd, &st) < 0) { int e = errno; close(fd); return -e; }
    lsio_oracle_range r;
    if (lsio_oracle_resolve_range(st.st_size, start, end, 0, &r) != 0) { close(fd); return -EINVAL; }
    int64_t want = r.end_off - r.start_off;
    if (want < 0 || (size_t)want > cap || r.end_off > st.st_size) { close(fd); return -EINVAL; }
    int64_t done = 0;
    while (done < want) {
        ssize_t n = pread(fd, dst + done, (size_t)(want - done), r.start_off + done);
        if (n < 0) { if (errno == EINTR) continue; int e = errno; close(fd); return -e; }
        if (n == 0) break;
        done += n;
    }
    close(fd);
    return done;
}

/* ------------------------------------------------------------------------------------------
 * CRC-32C (Castagnoli), RFC 3720 B.4: reflected polynomial 0x82F63B78, init and xorout
 * 0xFFFFFFFF.  `crc` argument/return are the finalised value (0 to start), zlib style.
 * ---------------------------------------------------------------------------------------- */
#define CRC32C_POLY_REFLECTED 0x82F63B78u
uint32_t lsio_oracle_crc32c_bitwise(uint32_t crc, const uint8_t* p, size_t n) {
    uint32_t c = ~crc;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (CRC32C_POLY_REFLECTED & (0u - (c & 1u)));
    }
    return ~c;
}
static uint32_t crc_tab[8][256];
static int crc_tab_ready;
static void crc_tab_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (CRC32C_POLY_REFLECTED & (0u - (c & 1u)));
        crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++)
            crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xFF];
    __atomic_store_n(&crc_tab_ready, 1, __ATOMIC_RELEASE);
}
uint32_t lsio_oracle_crc32c(uint32_t crc, const uint8_t* p, size_t n) {
    uint32_t c = ~crc;
#if defined(__SSE4_2__)
    while (n && ((uintptr_t)p & 7)) { c = _mm_crc32_u8(c, *p++); n--; }
    uint64_t c64 = c;
    for (; n >= 8; n -= 8, p += 8) { uint64_t w; memcpy(&w, p, 8); c64 = _mm_crc32_u64(c64, w); }
    c = (uint32_t)c64;
    while (n--) c = _mm_crc32_u8(c, *p++);
#else
    if (!__atomic_load_n(&crc_tab_ready, __ATOMIC_ACQUIRE)) crc_tab_init();
    for (; n >= 8; n -= 8, p += 8) {
        uint32_t lo, hi; memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
        lo ^= c;
        c = crc_tab[7][lo & 0xFF] ^ crc_tab[6][(lo >> 8) & 0xFF] ^ crc_tab[5][(lo >> 16) & 0xFF] ^
            crc_tab[4][lo >> 24] ^ crc_tab[3][hi & 0xFF] ^ crc_tab[2][(hi >> 8) & 0xFF] ^
            crc_tab[1][(hi >> 16) & 0xFF] ^ crc_tab[0][hi >> 24];
    }
    while (n--) c = (c >> 8) ^ crc_tab[0][(c ^ *p++) & 0xFF];
#endif
    return ~c;
}
/* CRC-32C of consecutive blocks of a file: out[k] = crc32c(file[k*blocksize, min((k+1)*blocksize, size))), k < n_blocks.
 * Buffered pread, one block at a time (bench.py builds its expected-CRC tables with this: the landed chunks'
 * device CRCs are compared with values the product never computed).  Returns blocks done or -errno. */
int64_t lsio_oracle_crc32c_file_blocks(const char* path, uint64_t blocksize, uint64_t n_blocks, uint32_t* out) {
    if (!path || !blocksize || !out) return -EINVAL;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -errno;
    uint8_t* buf = (uint8_t*)malloc(blocksize);
    if (!buf) { close(fd); return -ENOMEM; }
    uint64_t k = 0;
    for (; k < n_blocks; k++) {
        uint64_t done = 0;
        while (done < blocksize) {
            ssize_t n = pread(fd, buf + done, blocksize - done, (off_t)(k * blocksize + done));
            if (n < 0) { if (errno == EINTR) continue; int e = errno; free(buf); close(fd); return -e; }
            if (n == 0) break;
            done += (uint64_t)n;
        }
        if (!done) break;
        out[k] = lsio_oracle_crc32c(0, buf, done);
        if (done < blocksize) { k++; break; }
    }
    free(buf);
    close(fd);
    return (int64_t)k;
}
/* table-only variant kept callable so the SSE path can be cross-checked in tests */
uint32_t lsio_oracle_crc32c_table(uint32_t crc, const uint8_t* p, size_t n) {
    if (!__atomic_load_n(&crc_tab_ready, __ATOMIC_ACQUIRE)) crc_tab_init();
    uint32_t c = ~crc;
    while (n--) c = (c >> 8) ^ crc_tab[0][(c ^ *p++) & 0xFF];
    return ~c;
}

/* crc(A || B) from crc(A), crc(B), len(B): multiply crc(A) by x^(8 len_b) mod P in GF(2)
 * (zlib's crc32_combine construction, restated for the Castagnoli polynomial). */
static uint32_t gf2_matrix_times(const uint32_t* mat, uint32_t vec) {
    uint32_t sum = 0;
    while (vec) { if (vec & 1) sum ^= *mat; vec >>= 1; mat++; }
    return sum;
}
static void gf2_matrix_square(uint32_t* square, const uint32_t* mat) {
    for (int n = 0; n < 32; n++) square[n] = gf2_matrix_times(mat, mat[n]);
}
uint32_t lsio_oracle_crc32c_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
    uint32_t even[32], odd[32];
    if (len2 == 0) return crc1;
    odd[0] = CRC32C_POLY_REFLECTED; /* operator for one zero bit */
    uint32_t row = 1;
    for (int n = 1; n < 32; n++) { odd[n] = row; row <<= 1; }
    gf2_matrix_square(even, odd); /* two zero bits */
    gf2_matrix_square(odd, even); /* four zero bits */
    do {
        gf2_matrix_square(even, odd);
        if (len2 & 1) crc1 = gf2_matrix_times(even, crc1);
        len2 >>= 1;
        if (len2 == 0) break;
        gf2_matrix_square(odd, even);
        if (len2 & 1) crc1 = gf2_matrix_times(odd, crc1);
        len2 >>= 1;
    } while (len2 != 0);
    return crc1 ^ crc2;
}

/* ------------------------------------------------------------------------------------------
 * XXH64 (xxHash specification, 64-bit variant), scalar.
 * ---------------------------------------------------------------------------------------- */
#define XP1 0x9E3779B185EBCA87ULL
#define XP2 0xC2B2AE3D27D4EB4FULL
#define XP3 0x165667B19E3779F9ULL
#define XP4 0x85EBCA77C2B2AE63ULL
#define XP5 0x27D4EB2F165667C5ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t xxh_round(uint64_t acc, uint64_t in) { return rotl64(acc + in * XP2, 31) * XP1; }
static inline uint64_t xxh_merge(uint64_t acc, uint64_t v) { return (acc ^ xxh_round(0, v)) * XP1 + XP4; }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
uint64_t lsio_oracle_xxh64(const uint8_t* p, size_t n, uint64_t seed) {
    const uint8_t* end = p + n;
    uint64_t h;
    if (n >= 32) {
        uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        const uint8_t* limit = end - 32;
        do {
            v1 = xxh_round(v1, rd64(p)); v2 = xxh_round(v2, rd64(p + 8));
            v3 = xxh_round(v3, rd64(p + 16)); v4 = xxh_round(v4, rd64(p + 24));
            p += 32;
        } while (p <= limit);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    } else {
        h = seed + XP5;
    }
    h += (uint64_t)n;
    while (p + 8 <= end) { h ^= xxh_round(0, rd64(p)); h = rotl64(h, 27) * XP1 + XP4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * XP1; h = rotl64(h, 23) * XP2 + XP3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p++) * XP5; h = rotl64(h, 11) * XP1; }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}

/* What the pack/strip kernel must produce: the bytes of each padded read's view
 * (sqe.rs:116-119: [start_off - aligned_start, end_off - aligned_start)) gathered at the
 * request's prefix-sum offsets, with per-chunk checksums. */
void lsio_oracle_pack(const uint8_t* src, uint8_t* dst, const uint64_t* src_off,
                      const uint64_t* dst_off, const uint64_t* len, size_t n, uint32_t* crc,
                      uint64_t* xxh) {
    for (size_t i = 0; i < n; i++) {
        memcpy(dst + dst_off[i], src + src_off[i], (size_t)len[i]);
        if (crc) crc[i] = lsio_oracle_crc32c(0, dst + dst_off[i], (size_t)len[i]);
        if (xxh) xxh[i] = lsio_oracle_xxh64(dst + dst_off[i], (size_t)len[i], 0);
    }
}
