/*
 * lsio_oracle.h -- CPU ORACLE for the LSIO hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the algorithm of JackKelly/light-speed-io's
 * `lsio_uring` batched range read (reference files cited per function below, relative to
 * /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library (light-speed-io_b200/csrc) never
 * links, imports or calls anything in this directory.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - The reference is pure Rust; no cargo/rustc in this image, so oracle/_ref cannot be
 *     built ("unbuildable": needs cargo + un-vendored crates io-uring 0.6.4, crossbeam-*,
 *     nix 0.28, libc, anyhow).
 *   - Pinned by the reference's own tests, restated in tests/test_oracle_*.py:
 *       crates/lsio_uring/tests/integration_test.rs:13-120   (aligned 4 KiB ranges -> file)
 *       crates/lsio_uring/src/tracker.rs:74-114              (slot order 0,1,0)
 *       crates/lsio_uring/src/user_data.rs:47-56             (u64 packing round trip)
 *       crates/lsio_aligned_bytes/src/lib.rs:215-251 + README.md doctests (split/freeze/slice)
 *       crates/lsio_threadpool/src/threadpool.rs:186-268     (exactly-once delivery)
 *   - Unaligned starts/lengths, negative offsets, EOF tails, error paths: PARITY UNPINNED by
 *     any reference test (the reference panics or gets EINVAL there); ground
 *     truth is the file's bytes [start_off, end_off) read with buffered pread.
 *   - crc32c / xxh64 do not exist in the reference; pinned by RFC 3720 B.4 vectors and
 *     python-xxhash vectors committed under tests/golden/.
 */
#ifndef LSIO_ORACLE_H
#define LSIO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSIO_ORACLE_ALIGN 512 /* crates/lsio_uring/src/sqe.rs:12 */

/* ---- range arithmetic: crates/lsio_uring/src/sqe.rs:59-122 ------------------------------ */
typedef struct {
    int64_t start_off;     /* sqe.rs:64-71  */
    int64_t end_off;       /* sqe.rs:73-81  */
    int64_t aligned_start; /* sqe.rs:83     */
    int64_t buf_len;       /* sqe.rs:87: end_off - aligned_start (un-padded, what the reference reads) */
    int64_t padded_len;    /* buf_len rounded up to ALIGN: size of the allocation (lsio_aligned_bytes lib.rs:178-181) */
    int64_t view_start;    /* sqe.rs:116 */
    int64_t view_end;      /* sqe.rs:117 */
} lsio_oracle_range;

/* Returns 0, or a negative code where the reference would panic:
 *  -1 start_off < 0 (sqe.rs:71)   -2 end_off < 0 (sqe.rs:81)   -3 buf_len <= 0 (sqe.rs:88)
 *  -4 range.len() > 2_147_479_552 (get_range.rs:23-30)
 *  -5 start not ALIGN-aligned: split_to() error unwrap (sqe.rs:109-113, lib.rs:82-86);
 *     only reported when strict != 0.  With strict == 0 the evident intent
 *     (view = file[start_off, end_off)) is returned instead.                               */
int lsio_oracle_resolve_range(int64_t filesize, int64_t start, int64_t end, int strict,
                              lsio_oracle_range* out);

/* ---- user_data packing: crates/lsio_uring/src/user_data.rs:28-41 ------------------------ */
uint64_t lsio_oracle_user_data_pack(uint32_t index_of_op, uint8_t opcode);
void lsio_oracle_user_data_unpack(uint64_t v, uint32_t* index_of_op, uint8_t* opcode);

/* ---- tracker: crates/lsio_uring/src/tracker.rs:3-68 -------------------------------------- */
typedef struct lsio_oracle_tracker lsio_oracle_tracker;
lsio_oracle_tracker* lsio_oracle_tracker_new(size_t n);
void lsio_oracle_tracker_free(lsio_oracle_tracker*);
int64_t lsio_oracle_tracker_get_next_index(lsio_oracle_tracker*);      /* -1 == None */
void lsio_oracle_tracker_put(lsio_oracle_tracker*, size_t idx, uint64_t op);
int lsio_oracle_tracker_get(lsio_oracle_tracker*, size_t idx, uint64_t* op); /* 0 == None, -1 == out of bounds (panic) */
int lsio_oracle_tracker_remove(lsio_oracle_tracker*, size_t idx, uint64_t* op);
int lsio_oracle_tracker_is_empty(const lsio_oracle_tracker*);
int lsio_oracle_tracker_is_full(const lsio_oracle_tracker*);

/* ---- aligned bytes model: crates/lsio_aligned_bytes/src/lib.rs -------------------------- */
size_t lsio_oracle_padded_size(size_t len, size_t align);              /* lib.rs:178-181 pad_to_align */
/* split_to validity (lib.rs:74-95): 0 ok, -1 idx not in range, -2 idx == 0, -3 idx not aligned */
int lsio_oracle_split_to_check(size_t range_start, size_t range_end, size_t idx, size_t align);

/* ---- synthetic content (not from the reference except the ref pattern) ------- */
#define LSIO_ORACLE_SEED 0x4C53494FULL /* "LSIO" */
void lsio_oracle_pattern_fill(uint64_t file_id, uint64_t off, uint8_t* dst, size_t len);
/* crates/lsio_bench/src/main.rs:102: byte[i] = i as u8 */
void lsio_oracle_ref_pattern_fill(uint64_t off, uint8_t* dst, size_t len);
/* Write a whole file with either pattern. Returns 0 or -errno. */
int lsio_oracle_write_file(const char* path, uint64_t file_id, uint64_t size, int ref_pattern);

/* ---- ground truth read: buffered pread of [start_off, end_off) -------------------------- */
/* Resolves (start, end) with the sqe.rs rule against the file's size, then preads.
 * Returns number of bytes written to dst (== end_off - start_off), or -errno / -EINVAL.   */
int64_t lsio_oracle_pread_range(const char* path, int64_t start, int64_t end, uint8_t* dst,
                                size_t cap);

/* ---- checksums (not in the reference; RFC 3720 / xxHash spec) ---------------------------- */
uint32_t lsio_oracle_crc32c_bitwise(uint32_t crc, const uint8_t* p, size_t n); /* definition */
uint32_t lsio_oracle_crc32c(uint32_t crc, const uint8_t* p, size_t n);         /* slice-by-8 / SSE4.2 */
uint32_t lsio_oracle_crc32c_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);
int64_t lsio_oracle_crc32c_file_blocks(const char* path, uint64_t blocksize, uint64_t n_blocks, uint32_t* out);
uint64_t lsio_oracle_xxh64(const uint8_t* p, size_t n, uint64_t seed);

/* ---- pack/strip restatement: what the pack kernel must produce -------------------------- */
/* Gathers n segments src[src_off[i] .. +len[i]) into dst[dst_off[i] ..) and returns per-segment
 * crc32c / xxh64 (either may be NULL).                                                     */
void lsio_oracle_pack(const uint8_t* src, uint8_t* dst, const uint64_t* src_off,
                      const uint64_t* dst_off, const uint64_t* len, size_t n, uint32_t* crc,
                      uint64_t* xxh);

#ifdef __cplusplus
}
#endif
#endif
