/*
 * lsio_cuda.h -- C ABI of the `lsio_cuda` layer: the hand-written sm_100a kernels of the LSIO
 * hot path (BASELINE.json north_star: "a new lsio_cuda crate (Rust over a thin C-ABI to .cu
 * files built for sm_100a ...) holds the hand-written kernels").  There is no counterpart in
 * the reference (it has no GPU code); what these entry points replace is the
 * host-side view arithmetic the reference does per read:
 *     crates/lsio_uring/src/sqe.rs:106-119     split_to()/set_slice(): strip the O_DIRECT
 *                                              head/tail padding from the read buffer
 *     crates/lsio_aligned_bytes/src/lib.rs:137-165  set_slice/as_slice: the per-range view
 * re-targeted so the view is materialised as packed bytes in HBM and proven by a checksum.
 *
 * Plain pointers and sizes only; all `dev` pointers are CUDA device addresses on `device`.
 * Every function returns 0 on success, a negative errno-style code for bad arguments, or a
 * positive cudaError_t value when the CUDA runtime reported an error.  There is no CPU
 * fallback: without a usable CUDA device these functions fail.
 */
#ifndef LSIO_CUDA_H
#define LSIO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One padded read (or any byte span) to gather: copy `len` bytes from device address `src`
 * to device address `dst`.  Neither needs any alignment.  The 16-byte-aligned superset of
 * [src, src+len) must be readable (true for any span inside a cudaMalloc'd buffer whose size
 * is a multiple of 16 -- the staging slots are 512-aligned and 512-padded). */
typedef struct {
    uint64_t src;
    uint64_t dst;
    uint64_t len;
} lsio_cuda_segment;

#define LSIO_CUDA_PACK_STORE 1u  /* write the packed bytes (strip + gather)            */
#define LSIO_CUDA_PACK_CRC32C 2u /* fused per-segment CRC-32C of the bytes as landed   */
#define LSIO_CUDA_PACK_LIGHT 4u  /* small-footprint variant for launches that share the GPU with other kernels or carry
                                    little data: 256-thread CTAs, 4 KiB of shared memory, one launch (the fold is fused).
                                    Same results; the default variant (one 1024-thread CTA per SM, 128 KiB of tables) is
                                    the one that reaches the HBM roofline on large launches.
                                    With STORE, a segment whose src == dst is checksummed in place (nothing is moved). */

typedef struct lsio_cuda_pack_plan lsio_cuda_pack_plan;

/* Builds the device-resident descriptor and tile tables for `n` segments.  `flags` is a
 * combination of LSIO_CUDA_PACK_*; STORE|CRC32C is the fused pack_strip_crc32c kernel,
 * CRC32C alone checksums in place (dst is ignored, src is read), STORE alone is pack_strip. */
int lsio_cuda_pack_plan_create(int device, const lsio_cuda_segment* segs, size_t n, uint32_t flags,
                               lsio_cuda_pack_plan** out);
/* Enqueues ONE kernel launch on `stream` (a cudaStream_t, may be NULL for the default
 * stream).  `crc_out_dev` receives n uint32 results (may be NULL without CRC32C). Asynchronous. */
int lsio_cuda_pack_plan_launch(lsio_cuda_pack_plan* plan, uint32_t* crc_out_dev, void* stream);
/* Algorithmic HBM bytes of one launch: 2*sum(len) with STORE, sum(len) without. */
uint64_t lsio_cuda_pack_plan_algorithmic_bytes(const lsio_cuda_pack_plan* plan);
uint64_t lsio_cuda_pack_plan_num_tiles(const lsio_cuda_pack_plan* plan);
void lsio_cuda_pack_plan_destroy(lsio_cuda_pack_plan* plan);

/* One-shot convenience: plan + launch + synchronise + destroy. */
int lsio_cuda_pack_strip_crc32c(int device, const lsio_cuda_segment* segs, size_t n, uint32_t flags,
                                uint32_t* crc_out_dev, void* stream);

/* The PULL form of pack_strip_crc32c (pack_crc.cu pull_kernel): `segs[i].src` are HOST addresses inside one
 * pinned allocation (cudaHostAlloc / cudaHostRegister with the mapped flag): the kernel's warps read them over PCIe
 * through the mapped alias, checksum what they move and store it at `segs[i].dst` in HBM -- one pass, no copy engine, no
 * device staging copy.  This is how the backend empties a pinned slot when `direct_landing` = 2.  At most
 * LSIO_CUDA_PULL_MAX_SEGMENTS segments per call (the descriptors travel as kernel parameters).  `crc_out_dev` may be
 * NULL (no checksum) or any device-visible address of n words.  Synchronous one-shot (scratch allocated per call); the
 * backend itself keeps per-slot scratch and launches asynchronously on its workers' streams. */
#define LSIO_CUDA_PULL_MAX_SEGMENTS 64
int lsio_cuda_pull_strip_crc32c(int device, const lsio_cuda_segment* segs_host_src, size_t n, uint32_t* crc_out_dev, void* stream);

/* XXH64 (seed 0) of n spans; one 64-bit result per span.  Spans are device addresses. */
int lsio_cuda_xxh64(int device, const uint64_t* ptrs, const uint64_t* lens, size_t n,
                    uint64_t* out_dev, void* stream);
/* Same, choosing the kernel: 0 = default (the plain kernel: four lanes per chunk, 8-byte loads), 1 = plain, 2 = the chunk staged through
 * shared memory by the bulk-copy engine (cp.async.bulk): an A/B that measured slower at every shape and is kept for that record. */
int lsio_cuda_xxh64_variant(int device, const uint64_t* ptrs, const uint64_t* lens, size_t n,
                            uint64_t* out_dev, void* stream, int variant);

/* Test aid (SURVEY 2.2 verify_pattern): recompute the synthetic content
 * byte(file_id, off) in-kernel for n spans and count mismatching bytes into *mismatch_dev
 * (one uint64 total).  Lets large runs prove the landed bytes without a D2H copy. */
int lsio_cuda_verify_pattern(int device, const uint64_t* ptrs, const uint64_t* lens,
                             const uint64_t* file_ids, const uint64_t* file_offs, size_t n,
                             uint64_t* mismatch_dev, void* stream);

/* Test aid: the generator side of verify_pattern -- writes byte(file_id, file_off + i) to
 * dst_dev[i] for i in [0, len). Used to create the synthetic files and staging images. */
int lsio_cuda_fill_pattern(int device, void* dst_dev, uint64_t len, uint64_t file_id, uint64_t file_off, void* stream);

/* Number of kernel launches this library has issued in this process (bench `gpu_launches`). */
uint64_t lsio_cuda_launch_count(void);
/* Human-readable text for the last CUDA error seen by this library on the calling thread. */
const char* lsio_cuda_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
