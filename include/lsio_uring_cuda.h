/*
 * lsio_uring_cuda.h -- C ABI of the B200 backend: a drop-in for the reference's
 * `lsio_uring::IoUring` (crates/lsio_uring/src/io_uring.rs:9-49) whose chunks land in GPU HBM.
 *
 *   reference item                                   (file:line)              here
 *   IoUring::new(n_worker_threads)                   io_uring.rs:15-27        lsio_uring_cuda_new
 *   impl Reader for IoUring :: get_ranges            io_uring.rs:36-49        lsio_uring_cuda_get_ranges
 *   impl Completion for IoUring :: completion        io_uring.rs:30-34        lsio_uring_cuda_completion
 *   Drop (ThreadPool joins workers)                  threadpool.rs:140-165    lsio_uring_cuda_free
 *   compile-time knobs SQ_RING_SIZE, ALIGN, channel  worker.rs:14-22,         lsio_uring_cuda_config
 *     bound, SQPOLL idle                             sqe.rs:12, io_uring.rs:16
 *
 * Submit/complete semantics are the reference's: get_ranges copies its inputs and returns
 * immediately; exactly one Ok chunk per successful range, `user_data` echoed, completions
 * unordered; failures arrive in-band as Err items; the channel is bounded (back-pressure) and
 * disconnects once the backend is freed.  Differences, all deliberate:
 * chunk.buffer is a DEVICE view; unaligned starts/lengths work (the reference panics / gets
 * EINVAL); short reads are retried; bad ranges are Err items, not panics.
 */
#ifndef LSIO_URING_CUDA_H
#define LSIO_URING_CUDA_H

#include <stddef.h>
#include <stdint.h>

#include "lsio_io.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsio_uring_cuda lsio_uring_cuda;

#define LSIO_MAX_DEVICES 16

typedef struct lsio_uring_cuda_config {
    uint32_t n_worker_threads;  /* io_uring worker threads PER DEVICE group (IoUring::new's argument)      */
    uint32_t n_devices;         /* number of entries used in devices[]                                      */
    int32_t devices[LSIO_MAX_DEVICES]; /* CUDA ordinals; files are dealt to groups round-robin unless routed */
    uint32_t sq_entries;        /* per-worker SQ size (reference: 64, worker.rs:17)                          */
    uint32_t sqpoll;            /* 1 = IORING_SETUP_SQPOLL with 1000 ms idle as the reference (worker.rs:40) */
    uint32_t completion_bound;  /* bounded channel capacity (reference: 1024, io_uring.rs:16)                */
    uint32_t checksum;          /* LSIO_CHECKSUM_* bits computed on the device per chunk                     */
    uint32_t pack_align;        /* each chunk starts at a multiple of this in its request buffer (1 = tight) */
    uint32_t pin_threads;       /* 1 = pin each group's threads to the CPUs of its GPU's NUMA node           */
    uint32_t o_direct;          /* 1 = O_RDONLY|O_DIRECT as the reference (sqe.rs:25); 0 = buffered          */
    uint64_t staging_bytes;     /* pinned + io_uring-registered staging ring per device                      */
    uint64_t arena_bytes;       /* HBM arena per device that request buffers are carved from                 */
    uint64_t max_read_bytes;    /* reads (merged or not) are cut at this size; default 512 KiB, <= 1 GiB      */
    uint64_t min_copy_bytes;    /* a pack launch (staging_mode 0: and its coalesced H2D copy) is batched up to at least this
                                   many landed bytes (default 64 MiB) ...                                                     */
    uint32_t flush_us;          /* ...or until the oldest landed extent has waited this long (default 2000), or the burst is
                                   over (nothing being read and 100 us without a landing)                                   */
    uint32_t merge_ranges;      /* 1 = ranges of one request whose padded spans touch/overlap share one read    */
    uint64_t merge_gap_bytes;   /* ...also when separated by at most this many bytes (the gap is read and dropped) */
    uint32_t statx_via_ring;    /* 1 = the reference's two unlinked SQEs per file, openat + statx (get_ranges.rs:49-65):
                                   a missing file yields two Err items; the kernel always runs statx on an io-wq
                                   worker.  0 (default) = openat SQE, then fstat() on the new fd when it completes
                                   (never blocks, no second SQE): a missing file yields one Err item, as the
                                   reference's own doc comment promises (lsio_io lib.rs:30-35) */
    uint32_t poll_sleep_us;     /* device thread: 0 = spin-poll while work is in flight (lowest latency, one core per GPU);
                                   n > 0 = sleep n microseconds between polls (for boxes with few CPUs per GPU) */
    uint32_t zero_copy;         /* 1 = no H2D memcpy: the pack kernel reads the mapped pinned ring over PCIe and writes
                                   HBM once (fuses H2D + strip + checksum); 0 (default) = cudaMemcpyAsync + device ring */
    uint32_t staging_mode;      /* 1 (default) = every io worker owns `host_slots_per_worker` pinned slots
                                   (a window that stays in its core's cache; slots are `host_slot_kib` each) and copies each landed read to the
                                   device log itself on a private stream: the host is DRAM-bandwidth bound and a
                                   cache-resident window avoids the write-allocate/write-back traffic of a big ring
                                   (measured: tmpfs reads 69 GiB/s into <= 2 MiB/thread buffers vs 40 into a big one).
                                   0 = one `staging_bytes` pinned ring mirrored on the device, copied in coalesced
                                   runs by the device thread (needed for zero_copy) */
    uint32_t host_slots_per_worker; /* staging_mode 1: slots per worker; 0 (default) = 2..4, as many as window_budget_kib allows */
    uint32_t host_slot_kib;     /* staging_mode 1: slot (extent) size in KiB, >= max_read_bytes; 0 (default) = 512, or 256 when two of those per worker
                                   would not fit window_budget_kib: the unit that
                                   leaves the host in one pull launch / H2D copy; several reads of max_read_bytes can be in
                                   flight in one.  With the pull kernel (direct_landing 2) small slots cost no PCIe bandwidth,
                                   so the window (slots x slot size) is sized to stay in the worker's L2                  */
    uint32_t pack_threads;      /* threads per CTA of the pipeline's pack launches: 0 = 1024 (one CTA fills an SM: the
                                   configuration the kernel reaches the HBM roofline with); 256 = small footprint for a GPU
                                   shared with the consumer's own kernels (a 1024-thread CTA waits for a whole SM to drain,
                                   a 256-thread one for a single slot); the pipeline is PCIe-bound either way              */
    /* ---- round 2 ---- */
    uint32_t inline_reads;      /* what a worker does with a file the kernel cannot read asynchronously (no FMODE_NOWAIT:
                                   tmpfs, hugetlbfs, some FUSE -- probed once per filesystem with preadv2(RWF_NOWAIT)):
                                   io_uring would hand every such read to an io-wq kernel thread (a thread hop per read and
                                   twice the threads; measured 3.3 vs 4.6 GiB/s per thread on tmpfs).  1 (default) = read those
                                   files with preadv2() on the worker itself, into the same registered slot; 0 = always through
                                   the ring (the reference's behaviour); 2 = always inline.  Files that can be read
                                   asynchronously (O_DIRECT on a block device) always go through READ_FIXED unless 2.        */
    uint32_t force_align;       /* 0 (default) = per-file O_DIRECT alignment max(512, stx_dio_offset_align, stx_dio_mem_align)
                                   from statx(STATX_DIOALIGN) on the opened file (reference: fetched, unused, 512 hard-coded:
                                   open_file.rs:23-25, sqe.rs:12); n = use n instead (power of two >= 512; test knob for 4Kn)  */
    uint32_t debug_short_read_every; /* test knob: every k-th read is issued with a truncated length so that the short-read
                                   retry path (get_range.rs:66-68 TODO in the reference) runs; 0 = off                        */
    uint32_t null_source;       /* diagnostic: 1 = workers skip the read itself (slots keep whatever they held; files are still
                                   opened and sized): measures the device stage alone -- bytes delivered are NOT the files'   */
    uint32_t direct_landing;    /* how an extent leaves its pinned slot.  0 = always H2D copy into the device staging ring, then the
                                   pack kernel strips / scatters / checksums (HBM traffic 3L).  1 = an extent whose payload forms
                                   at most 4 runs contiguous both in the slot and in the request buffers is copied by the copy
                                   engine straight into the arena and checksummed in place (traffic 2L), consecutive slots
                                   coalesced (copy_coalesce_slots).  2 = an extent of at most 64 segments whose slot addresses are
                                   congruent to their destinations mod 16 (every 512-aligned block read) is moved by the PULL
                                   KERNEL: SM loads over PCIe from the mapped slot, CRC fused, stored at the final place -- one
                                   pass, traffic L, no copy-engine gap per copy, so slots can stay cache-sized.  Extents that do
                                   not qualify fall back to 1's rule, then to the ring (small unaligned ranges).              */
    uint32_t cpu_first;         /* with cpu_count > 0: this backend's threads (workers + device thread of every group) are     */
    uint32_t cpu_count;         /* pinned to CPUs [cpu_first, cpu_first + cpu_count) instead of the GPU's NUMA node            */
    uint32_t steal_batch;       /* tasks a worker takes from the BACK of a peer's deque in one steal, kept in order (default 8 =
                                   two extents' worth of reads: the owner's own sequence keeps no holes and both sides
                                   fill their slots with consecutive bytes; 1 = the reference's Stealer::steal(), one task
                                   from the front, worker.rs:58-62)                                                          */
    uint32_t pack_light_below_mib; /* pipeline launches over batches smaller than this use the light kernel variant
                                   (256-thread CTAs, 5 KiB of shared memory, co-resident with a consumer's kernels, fold fused);
                                   larger ones the roofline variant.  default 512; 0 = never light                          */
    uint32_t copy_coalesce_slots; /* directly landed extents that fill consecutive slots of a worker AND consecutive bytes of one
                                   request buffer are copied with one cudaMemcpyAsync of up to this many slots (default 2,
                                   at most 8; 1 = one copy per extent).  The copy engine loses ~3.8 us per copy whatever its
                                   size (2 MiB copies: 46.9 GiB/s, 4 MiB: 49.2, 8 MiB: 50.4 of 51.6 on this pool's B200s).     */
    uint32_t window_budget_kib; /* staging_mode 1, automatic slot geometry: the pinned windows of ALL workers of one device group
                                   together stay within this many KiB, so that a slot is still in the last-level cache when the
                                   GPU reads it back (otherwise each payload byte costs three extra DRAM transfers).
                                   0 (default) = 60 % of the L3 reported by sysfs, divided by the number of devices of this
                                   backend; one process per GPU on a shared host: pass this process's share                  */
    uint32_t pull_any_shift;    /* direct_landing 2: 1 (default) = extents whose segments sit at a different offset mod 16 in the slot than
                                   in the request buffer (unaligned ranges) are pulled too: each destination vector then needs two
                                   source loads, but the second one hits the L2 line the first one brought in, so the bytes still
                                   cross PCIe once; 0 = only congruent segments are pulled, the rest goes ring -> pack kernel          */
} lsio_uring_cuda_config;

typedef struct lsio_uring_cuda_stats {
    uint64_t requests, ranges, reads, short_reads, read_bytes;
    uint64_t h2d_copies, h2d_bytes, pack_launches, kernel_launches, payload_bytes;
    uint64_t chunks_ok, errors;
    uint64_t staging_stalls, arena_stalls;
    /* ---- round 2 ---- */
    uint64_t inline_reads;      /* reads done with preadv2 on the worker (see lsio_uring_cuda_config.inline_reads)  */
    uint64_t direct_extents;    /* extents landed by H2D DMA straight into the arena                                 */
    uint64_t direct_copies;     /* cudaMemcpyAsync calls of those extents (one per run)                              */
    uint64_t read_ns;           /* ns workers spent inside inline reads (sum over workers)                           */
    uint64_t copy_enqueue_ns;   /* ns workers spent enqueueing H2D copies                                            */
    uint64_t worker_stall_ns;   /* ns workers slept because no slot / no log space was free                          */
    uint64_t worker_busy_ns;    /* ns workers were not parked (sum over workers)                                     */
    uint64_t pack_kernel_ns;    /* device time of the pipeline's pack/crc launches (cudaEvent pairs on the pack stream) */
    uint64_t pack_kernel_bytes; /* algorithmic bytes of those launches (2L for ring segments, L for in-place ones)   */
    uint64_t first_chunk_ns;    /* ns from the first get_ranges after construction / reset to the first chunk sent   */
    uint64_t max_dio_align;     /* largest per-file O_DIRECT alignment seen                                         */
    uint64_t pull_launches;     /* extents moved by the pull kernel (direct_landing 2), one launch each             */
    uint64_t slot_bytes;        /* staging_mode 1: the slot geometry in use (after automatic sizing)                */
    uint64_t slots_per_worker;
} lsio_uring_cuda_stats;

/* Per-chunk latency histogram (submit of the request -> chunk in the channel), log2 buckets of microseconds:
 * bucket i counts latencies in [2^i, 2^(i+1)) us, i = 0..31.  Reset together with the stats' first_chunk_ns. */
int lsio_uring_cuda_get_latency_histogram(const lsio_uring_cuda* self, uint64_t buckets[32]);
void lsio_uring_cuda_reset_latency(lsio_uring_cuda* self);

void lsio_uring_cuda_config_default(lsio_uring_cuda_config* cfg);
/* IoUring::new(n_worker_threads): device 0, default config. Returns 0 or an error code. */
int lsio_uring_cuda_new(size_t n_worker_threads, lsio_uring_cuda** out);
int lsio_uring_cuda_new_with_config(const lsio_uring_cuda_config* cfg, lsio_uring_cuda** out);

/* Reader::get_ranges (lsio_io lib.rs:36-44).  `location` must be absolute (the reference opens
 * with dirfd -1, sqe.rs:22).  range i is [range_start[i], range_end[i]); negative values are
 * relative to the file size exactly as in the reference (lib.rs:19-23, sqe.rs:64-81).
 * Returns 0 after enqueueing; -EINVAL if n_ranges != n_user_data (the reference asserts,
 * get_ranges.rs:28) or location is NULL/relative; -ESHUTDOWN after lsio_uring_cuda_free. */
int lsio_uring_cuda_get_ranges(lsio_uring_cuda* self, const char* location, const int64_t* range_start,
                               const int64_t* range_end, size_t n_ranges, const uint64_t* user_data,
                               size_t n_user_data);
/* Same, routed to devices[group] instead of round-robin (file -> GPU affinity map). */
int lsio_uring_cuda_get_ranges_on(lsio_uring_cuda* self, uint32_t group, const char* location,
                                  const int64_t* range_start, const int64_t* range_end, size_t n_ranges,
                                  const uint64_t* user_data, size_t n_user_data);

/* Ordered groups of operations -- planned in the reference, not built there (lsio_io lib.rs:6-8,
 * planned_design.md:71-72: "the IO crate will guarantee that all operations in group n are completed
 * before any IO operations in group n+1 are started"; outputs "will also be grouped").
 * Every get_ranges call belongs to the backend's current group (0 at construction).  next_group()
 * closes the current group and returns the id of the new one: requests submitted from then on are
 * held back -- not even opened -- until every request of all earlier groups has delivered all of its
 * items (chunks and errors) to the completion channel, so items arrive in non-decreasing group order.
 * Within a group completions stay unordered.  Thread-safe; never blocks. */
uint64_t lsio_uring_cuda_next_group(lsio_uring_cuda* self);
/* The group new submissions are tagged with. */
uint64_t lsio_uring_cuda_current_group(const lsio_uring_cuda* self);

/* Completion::completion(): a new handle on the backend's receiver (drop it yourself). */
lsio_completion* lsio_uring_cuda_completion(lsio_uring_cuda* self);

int lsio_uring_cuda_get_stats(const lsio_uring_cuda* self, lsio_uring_cuda_stats* out);
/* Text of the last construction error on this thread. */
const char* lsio_uring_cuda_last_error(void);

/* Measurement aid (stands in for crates/lsio_uring/benches/fio.ini; fio is not installed):
 * the read ceiling of the storage under `paths` with no GPU work at all -- n_threads threads,
 * each with a private ring, issue `block_bytes` READ_FIXED reads at queue depth `queue_depth`
 * into a registered pinned buffer, striding over the files, until every file has been read
 * once. *gib_per_s receives aggregate GiB/s. */
int lsio_probe_read_ceiling(const char* const* paths, size_t n_paths, uint64_t block_bytes, uint32_t n_threads,
                            uint32_t queue_depth, uint32_t o_direct, double* gib_per_s);

/* Same, reading the way the backend itself would with `inline_reads` (lsio_uring_cuda_config): 0 = always READ_FIXED
 * through the ring, 1 = files the kernel cannot read asynchronously are read with preadv2 on the probing thread (cycling
 * through the same queue_depth * block_bytes window), 2 = always preadv2.  The I/O ceiling bench.py reports is the best
 * of several shapes of this probe, so that it bounds what the pipeline -- which does strictly more work per byte --
 * can reach on the same files.  `files_in_flight`: files each thread keeps open at once, their blocks issued round robin
 * (the queue stays full across file boundaries, as the backend keeps it); 1 = one file after the other. */
int lsio_probe_read_ceiling_ex(const char* const* paths, size_t n_paths, uint64_t block_bytes, uint32_t n_threads,
                               uint32_t queue_depth, uint32_t o_direct, uint32_t inline_reads, uint32_t files_in_flight,
                               double* gib_per_s);

/* Measurement aid: the minimum any host-staged path must do for a memory-backed source, and nothing else -- `n_threads` threads read
 * the files block by block with preadv2 into a cycle of `slots_per_thread` pinned slots of `slot_kib` and ship every filled slot to
 * `device` (pull kernel without checksum; use_copy_engine = 1: one cudaMemcpyAsync) into scratch memory.  No requests, planning, arena,
 * channel or checksums.  cpu_count > 0 pins the threads to [cpu_first, cpu_first + cpu_count).  Returns the wall time and the bytes
 * shipped; several ranks on one host call it between the same two barriers and add up: what the host can stage AND ship at once. */
int lsio_probe_staged_ceiling(const char* const* paths, size_t n_paths, int device, uint32_t n_threads, uint32_t slot_kib,
                              uint32_t slots_per_thread, uint32_t use_copy_engine, uint32_t cpu_first, uint32_t cpu_count, double* seconds,
                              uint64_t* bytes_out);

/* Test aid (no device, no I/O): the request planner's arithmetic on its own -- range resolution (sqe.rs:64-81), the prefix-sum layout of
 * the request's buffer, and the read plan (merge / split / per-file alignment; sqe.rs:83-119 re-stated for a whole request).
 * ranges_out: 4 int64 per range {valid, start_off, end_off, dst_off}; reads_out: 6 uint64 per read {file_off, read_len, needed, first_seg,
 * nsegs, ntiles}; segs_out: 4 uint64 per segment {range, head, payload_len, dst_off}.  *n_reads / *n_segs receive the full counts even
 * when they exceed the caps.  `buffer_address`: where the request's buffer would start (kernel tiles are counted in destination space). */
int lsio_uring_cuda_plan_reads(int64_t filesize, const int64_t* starts, const int64_t* ends, size_t n, uint32_t align, uint64_t max_read_bytes,
                               uint32_t merge_ranges, uint64_t merge_gap_bytes, uint32_t pack_align, uint64_t buffer_address, int64_t* ranges_out,
                               uint64_t* reads_out, size_t cap_reads, size_t* n_reads, uint64_t* segs_out, size_t cap_segs, size_t* n_segs,
                               uint64_t* total_bytes);

/* Test aid: host-logic self tests of the shipped components (tracker, user_data, channel, staging
 * log, arena, GF(2) combine), no GPU needed; see csrc/selftest.cpp. 0 = pass. */
int lsio_uring_cuda_selftest(int which);

/* Drop: finishes in-flight work, joins every thread, disconnects the channel.  Chunks already
 * received stay valid until dropped. */
void lsio_uring_cuda_free(lsio_uring_cuda* self);

#ifdef __cplusplus
}
#endif
#endif
