// backend.hpp -- shared types of the B200 backend (split along the reference's own modules):
//   planner.cpp        sqe.rs (range arithmetic) + get_ranges.rs (fan-out), plus the read planner
//   worker.cpp         worker.rs (UringWorker::run) + get_range.rs / close.rs state machines
//   device_stage.cpp   the stage the reference does not have: batching, pack+crc launches, chunk emission
//   io_uring_cuda.cpp  io_uring.rs (IoUring::new / get_ranges / completion / Drop) as a C ABI
#pragma once

#include <fcntl.h>
#include <linux/stat.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <deque>
#include <fstream>
#include <sstream>
#include <thread>

#include "aligned_bytes.hpp"
#include "channel.hpp"
#include "common.hpp"
#include "gf2.hpp"
#include "lsio_uring_cuda.h"
#include "op_groups.hpp"
#include "pack_batch.hpp"
#include "staging.hpp"
#include "threadpool.hpp"
#include "tracker.hpp"
#include "uring.hpp"

namespace lsio {

inline uint64_t now_ns() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr uint64_t kAlign = 512;  // crates/lsio_uring/src/sqe.rs:12 -- the MINIMUM; the per-file value is Request::align

// Limits that keep every unit of work inside one pack batch (device_stage.cpp flushes whole extents):
// a read carries at most kMaxSegsPerRead descriptors / kMaxTilesPerRead kernel tiles (the planner cuts there),
// an extent at most kMaxSegsPerExtent / kMaxTilesPerExtent (the worker closes it there), a batch holds
// kBatchMaxSegs / kBatchMaxTiles.
constexpr uint32_t kMaxSegsPerRead = 1024, kMaxTilesPerRead = 16384;
constexpr uint32_t kMaxSegsPerExtent = 4096, kMaxTilesPerExtent = 32768;
constexpr size_t kBatchMaxSegs = 8192, kBatchMaxTiles = 1 << 16;

struct Group;
struct Backend;
struct Request;
using StagingLog = StagingLogT<Request>;
using StageEntry = StageEntryT<Request>;

// One read serves one or more segments; a segment is the part of one range that one read carries.
// (The reference issues exactly one read per range; merging adjacent ranges into one read is its
// stated plan: lsio_aligned_bytes/README.md:43-57, README.md:78.)
struct ReadPlan {
    uint64_t file_off = 0;   // 512-aligned
    uint32_t read_len = 0;   // multiple of 512
    uint32_t needed = 0;     // bytes that must land: the furthest payload byte of any of its segments
    uint32_t first_seg = 0, nsegs = 0;  // its segments are consecutive in Request::segs
    uint32_t ntiles = 0;     // kernel tiles its segments need (pack_tile_count over their destinations)
};
struct SegPlan {
    uint32_t range = 0;
    uint32_t head = 0;         // offset of the payload inside the read's slot (padding / other ranges before it)
    uint64_t payload_len = 0;  // bytes of the range this segment carries
    uint64_t dst_off = 0;      // where they go in the request buffer
    uint32_t crc = 0;
};
struct RangePlan {
    int64_t start_off = 0, end_off = 0;
    uint64_t dst_off = 0, len = 0;
    std::vector<uint32_t> seg_ids;  // this range's segments in file order (for the CRC combine)
    uint64_t user_data = 0;
    bool valid = false;
    std::atomic<uint32_t> segs_left{0};
    std::atomic<bool> failed{false};
};

// One get_ranges call (GetRanges, get_ranges.rs:15-25 + OpenFile, open_file.rs:4-12).
struct Request {
    std::string path;
    std::vector<int64_t> starts, ends;
    std::vector<uint64_t> user_data;
    Group* grp = nullptr;
    Backend* be = nullptr;   // set once the request is counted in its ordered group (see OpGroups)
    uint64_t op_group = 0;   // ordered group it was submitted in (lsio_io lib.rs:6-8)
    struct statx stx;
    int fd = -1;
    bool have_fd = false, have_statx = false, planned = false;
    uint8_t n_cqes_received = 0;
    uint64_t total_bytes = 0;
    uint64_t arena_epoch_seen = 0;  // arena epoch when this request last found the arena full
    uint32_t align = kAlign;        // O_DIRECT alignment of this file: max(512, stx_dio_offset_align, stx_dio_mem_align) (open_file.rs:23-25)
    bool inline_io = false;         // the kernel cannot read this file asynchronously: the worker reads it with preadv2 itself
    uint64_t submit_ns = 0;         // when get_ranges was called (chunk latency)
    std::shared_ptr<InnerBuffer> buffer;  // the request's HBM buffer
    std::unique_ptr<RangePlan[]> ranges;
    std::vector<ReadPlan> reads;
    std::vector<SegPlan> segs;
    std::atomic<uint32_t> reads_left{0};  // Arc<OpenFile> strong count, get_range.rs:78
    ~Request();  // io_uring_cuda.cpp: closes a left-over fd, reports the request finished to its ordered group
};

// operation.rs:7-12, plus Plan (a request that is open but still waits for HBM).
struct Op {
    enum Kind : uint8_t { GetRanges, Plan, Read, Close } kind = GetRanges;
    std::shared_ptr<Request> req;
    uint32_t read = 0;  // index into req->reads
    StageEntry* entry = nullptr;
    uint32_t done = 0;  // bytes of this read that have landed (short-read continuation)
    int slot = -1;      // staging_mode 1: the worker's host slot (extent) this read lands in
    uint32_t ref = 0;   // index of this read in entry->reads
};

struct Batch {
    PackBatchLayout lay;
    PackBatchCounts counts;
    char* host_desc = nullptr;  // pinned
    char* dev_desc = nullptr;
    uint32_t* crc_host = nullptr;  // pinned + mapped: the kernel writes results straight to the host
    uint32_t* crc_dev = nullptr;   // device alias of crc_host
    cudaEvent_t copied = nullptr, done = nullptr;
    cudaEvent_t k0 = nullptr, k1 = nullptr;  // timing pair around this batch's kernel launches (stats.pack_kernel_ns)
    bool timed = false;
    uint64_t algo_bytes = 0;
    uint64_t seq_begin = 0, seq_end = 0;
    struct DescRef {
        uint64_t seq;  // staging extent (holds the request alive until the batch retires)
        uint32_t ref;  // read inside the extent
        uint32_t seg;
        uint32_t kidx;  // index of its descriptor in the kernel batch (its CRC slot), or UINT32_MAX: nothing launched for it
        uint32_t pull_idx;  // index of its CRC in Group::pull_crc_host (moved and checksummed by the pull kernel), or UINT32_MAX
    };
    std::vector<DescRef> desc;  // segment of each descriptor
    std::vector<cudaEvent_t> copy_events;
};

// Second stage when LSIO_CHECKSUM_XXH64 is requested: ranges whose bytes are packed wait here for
// one xxh64_chunks launch (arrays are mapped pinned memory: no copies in either direction).
struct XxhBatch {
    static constexpr size_t kMax = 8192;
    uint64_t* ptrs = nullptr;  // [kMax] ptrs | [kMax] lens | [kMax] out, mapped pinned
    uint64_t* dev = nullptr;   // device alias of ptrs
    cudaEvent_t done = nullptr;
    struct Item {
        std::shared_ptr<Request> req;
        uint32_t range;
        uint32_t crc;
    };
    std::vector<Item> items;
};

// staging_mode 1: the private staging window of one io worker.
struct WorkerSlots {
    uint8_t* base = nullptr;  // host_slots_per_worker slots of slot_bytes, inside Group::slot_block
    uint8_t* base_dev = nullptr;  // the same memory as the GPU sees it (mapped): the pull kernel reads the slots in place
    std::vector<void*> scratch;   // per slot: device scratch of the pull kernel (scheduler words, counters, partials)
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> ev;
    std::vector<uint8_t> busy;  // 0 free, 1 read in flight, 2 H2D copy / pull in flight, 3 filled, copy deferred (touched by the owning worker only)
};

struct Stats {
    std::atomic<uint64_t> requests{0}, ranges{0}, reads{0}, short_reads{0}, read_bytes{0}, h2d_copies{0}, h2d_bytes{0}, pack_launches{0},
        kernel_launches{0}, payload_bytes{0}, chunks_ok{0}, errors{0}, staging_stalls{0}, arena_stalls{0}, inline_reads{0}, direct_extents{0},
        direct_copies{0}, read_ns{0}, copy_enqueue_ns{0}, worker_stall_ns{0}, worker_busy_ns{0}, pack_kernel_ns{0}, pack_kernel_bytes{0},
        first_submit_ns{0}, first_chunk_ns{0}, max_dio_align{0}, read_seq{0}, pull_launches{0};
    std::atomic<uint64_t> latency[32] = {};  // log2(us) buckets: request submitted -> chunk in the channel
};

struct Backend;

struct Group {
    Backend* be = nullptr;
    int index = 0, device = 0;
    std::vector<int> cpus;
    std::shared_ptr<DeviceArena> arena;
    uint8_t* host_ring = nullptr;
    uint8_t* dev_ring = nullptr;
    uint8_t* host_ring_dev = nullptr;  // device alias of the mapped pinned ring (zero_copy mode)
    uint8_t* slot_block = nullptr;     // staging_mode 1: all workers' slots, one pinned allocation
    uint64_t slot_bytes = 0;
    std::vector<WorkerSlots> wslots;
    uint8_t* scratch_block = nullptr;   // all slots' pull scratch, one device allocation
    uint32_t* pull_crc_host = nullptr;  // pinned + mapped: [n_pull][kPullMaxSegs] CRCs written by the pull kernel, row = extent seq % n_pull
    uint32_t* pull_crc_dev = nullptr;
    size_t n_pull = 0;
    StagingLog log;
    cudaStream_t s_copy = nullptr, s_pack = nullptr;
    std::vector<Batch> batches;
    std::deque<Batch*> free_batches, inflight;
    std::vector<XxhBatch> xxh_batches;
    std::deque<XxhBatch*> xxh_free, xxh_inflight;
    std::vector<XxhBatch::Item> xxh_pending;
    std::vector<lsio_output> out_pending;  // device thread: chunks of the batch being retired, sent in one go
    std::unique_ptr<ThreadPool<Op>> pool;
    std::thread device_thread;
    std::atomic<bool> stop_device{false};
    std::atomic<uint32_t> workers_alive{0};
    uint64_t copy_seq = 0;  // device thread: next staging entry to copy
    Batch* last_flush_batch = nullptr;  // the batch flush_run() was filling when it failed
};

using OutChannel = Channel<lsio_output>;

// Ordered groups of operations (op_groups.hpp).  A request counts as finished when the last reference to it is
// released, which the device stage does only after the request's items are in the completion channel.
using OpGroups = OrderedGroups<std::pair<size_t, Op>>;  // (device group, op)

struct Backend {
    lsio_uring_cuda_config cfg;
    std::shared_ptr<OutChannel> out;
    OpGroups op_groups;  // declared before `groups`: requests destroyed with a Group still find it
    std::vector<std::unique_ptr<Group>> groups;
    std::atomic<uint64_t> rr{0};
    std::atomic<bool> closing{false};
    Stats stats;
};

// ---- planner.cpp -------------------------------------------------------------------------------------
enum class Start { Started, Finished, Retry };
void send_output(Backend* be, lsio_output& o);
// operation.rs:82-100 maybe_send_error: one Err per failed CQE, with context.
void send_error(Backend* be, int status, uint8_t opcode, const Request& r, int range, const char* what);
Start plan_request(Group& g, const std::shared_ptr<Request>& rp, WorkerThread<Op>& wt);
void resolve_ranges(Request& r, int64_t filesize, uint64_t pack_align, std::vector<uint32_t>* invalid);
void build_read_plan(Request& r, const lsio_uring_cuda_config& cfg, const uint8_t* block);

// ---- worker.cpp: body of one io worker thread (UringWorker::new + run, worker.rs:32-125) -----------------
void worker_main(Group& g, WorkerThread<Op>& wt);

// ---- device_stage.cpp: body of the per-GPU device thread ---------------------------------------------------
void device_main(Group& g);

}  // namespace lsio
