// io_uring_cuda.cpp -- the B200 backend: a from-scratch re-design of the reference's
// `lsio_uring` hot path (crates/lsio_uring/src/{io_uring,worker,operation,get_ranges,get_range,
// close,sqe,open_file}.rs) whose per-range chunks land in GPU HBM.
//
// Per device group (one group per GPU, threads pinned to that GPU's NUMA node):
//
//   caller ── get_ranges ─▶ injector ─▶ k io workers (work-stealing pool, private io_uring each)
//        openat(O_DIRECT) [+statx] ─▶ plan: resolve ranges (sqe.rs:64-83), lay the request out in one HBM
//        buffer (arena), merge/split into reads ─▶ READ_FIXED into the worker's own pinned slots, packed
//        back to back as 2 MiB EXTENTS that stay in the core's cache
//                                   │ all reads of an extent landed (CQEs, short reads retried)
//                                   ▼
//        the worker copies the extent with ONE cudaMemcpyAsync on its own stream into the device log
//   device thread:  batch published extents in log order ─▶ wait on their copy events
//        ─▶ pack_strip_crc32c over the batch (pack stream) ─▶ event
//        ─▶ one Chunk{device view, user_data, crc32c} per range on the bounded channel
//   (staging_mode 0: one big pinned ring instead; the device thread copies coalesced runs of landed reads)
//
// What is kept from the reference: submit/complete semantics, one chunk per range, unordered
// completions, in-band errors (one per failed CQE: operation.rs:82-100), user_data packing,
// tracker slots, openat+statx per file, close by the last reader, work stealing of per-range
// reads between workers, bounded(1024) completion channel.  What changes: no allocation per
// read, registered buffers, batched submission, the device boundary after the read CQE.
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
#include "pack_batch.hpp"
#include "staging.hpp"
#include "threadpool.hpp"
#include "tracker.hpp"
#include "uring.hpp"

namespace lsio {

static inline uint64_t now_ns() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr uint64_t kAlign = 512;  // crates/lsio_uring/src/sqe.rs:12

struct Group;
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
    struct statx stx;
    int fd = -1;
    bool have_fd = false, have_statx = false, planned = false;
    uint8_t n_cqes_received = 0;
    uint64_t total_bytes = 0;
    std::shared_ptr<InnerBuffer> buffer;  // the request's HBM buffer
    std::unique_ptr<RangePlan[]> ranges;
    std::vector<ReadPlan> reads;
    std::vector<SegPlan> segs;
    std::atomic<uint32_t> reads_left{0};  // Arc<OpenFile> strong count, get_range.rs:78
    ~Request() {
        if (fd >= 0) close(fd);  // only reached when the backend is torn down with reads still queued
    }
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
    uint64_t seq_begin = 0, seq_end = 0;
    struct DescRef {
        uint64_t seq;  // staging extent (holds the request alive until the batch retires)
        uint32_t ref;  // read inside the extent
        uint32_t seg;
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
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> ev;
    std::vector<uint8_t> busy;  // 0 free, 1 read in flight, 2 H2D copy in flight (touched by the owning worker only)
};

struct Stats {
    std::atomic<uint64_t> requests{0}, ranges{0}, reads{0}, short_reads{0}, read_bytes{0}, h2d_copies{0}, h2d_bytes{0}, pack_launches{0},
        kernel_launches{0}, payload_bytes{0}, chunks_ok{0}, errors{0}, staging_stalls{0}, arena_stalls{0};
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
    std::string init_error;
};

using OutChannel = Channel<lsio_output>;

struct Backend {
    lsio_uring_cuda_config cfg;
    std::shared_ptr<OutChannel> out;
    std::vector<std::unique_ptr<Group>> groups;
    std::atomic<uint64_t> rr{0};
    std::atomic<bool> closing{false};
    Stats stats;
};

static thread_local std::string g_last_error;
static void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

// ---- completion items ----------------------------------------------------------------------
static void send_output(Backend* be, lsio_output& o) {
    if (o.status != 0) be->stats.errors++;
    else be->stats.chunks_ok++;
    be->out->send(o);
}

// operation.rs:82-100 maybe_send_error: one Err per failed CQE, with context.
static void send_error(Backend* be, int status, uint8_t opcode, const Request& r, int range, const char* what) {
    lsio_output o;
    memset(&o, 0, sizeof o);
    o.status = status;
    o.opcode = opcode;
    o.device = r.grp ? r.grp->device : -1;
    if (range >= 0) {
        o.chunk.user_data = r.user_data[range];
        snprintf(o.context, sizeof o.context, "%s: %s (io_uring opcode %u, result %d). path=%s range=%lld..%lld user_data=%llu", what,
                 status < 0 ? strerror(-status) : "CUDA error", opcode, status, r.path.c_str(), (long long)r.starts[range], (long long)r.ends[range],
                 (unsigned long long)r.user_data[range]);
    } else {
        snprintf(o.context, sizeof o.context, "%s: %s (io_uring opcode %u, result %d). path=%s n_ranges=%zu", what,
                 status < 0 ? strerror(-status) : "CUDA error", opcode, status, r.path.c_str(), r.starts.size());
    }
    send_output(be, o);
}

// ---- planning: sqe.rs:59-122 re-stated for a whole request -----------------------------------
enum class Start { Started, Finished, Retry };

static Start plan_request(Group& g, const std::shared_ptr<Request>& rp, WorkerThread<Op>& wt) {
    Request& r = *rp;
    Backend* be = g.be;
    const lsio_uring_cuda_config& cfg = be->cfg;
    const int64_t filesize = (int64_t)r.stx.stx_size;
    const size_t n = r.starts.size();
    if (!r.planned) {
        r.ranges.reset(new RangePlan[n ? n : 1]);
        uint64_t dst = 0;
        const uint64_t pa = cfg.pack_align ? cfg.pack_align : 1;
        for (size_t i = 0; i < n; i++) {
            RangePlan& p = r.ranges[i];
            p.user_data = r.user_data[i];
            const int64_t s = r.starts[i], e = r.ends[i];
            const int64_t start_off = s >= 0 ? s : filesize + s;  // sqe.rs:64-70
            const int64_t end_off = e >= 0 ? e : filesize + e + 1;  // sqe.rs:73-80
            // The reference asserts (panics) on these; here they are in-band errors (Appendix A).
            if (start_off < 0 || end_off < 0 || end_off <= start_off || end_off > filesize) {
                send_error(be, -EINVAL, 0, r, (int)i, "invalid byte range for this file");
                continue;
            }
            p.valid = true;
            p.start_off = start_off;
            p.end_off = end_off;
            p.len = (uint64_t)(end_off - start_off);
            dst = (dst + pa - 1) / pa * pa;
            p.dst_off = dst;
            dst += p.len;
        }
        r.total_bytes = dst;
        r.planned = true;
        be->stats.ranges += n;
    }
    if (r.total_bytes == 0) {  // nothing to read
        close(r.fd);
        r.fd = -1;
        return Start::Finished;
    }
    uint8_t* block = g.arena->alloc(r.total_bytes + 16);
    if (!block) {
        if (r.total_bytes + 16 > g.arena->capacity()) {
            for (size_t i = 0; i < n; i++)
                if (r.ranges[i].valid) send_error(be, -ENOMEM, 0, r, (int)i, "request larger than the HBM arena (arena_bytes)");
            close(r.fd);
            r.fd = -1;
            return Start::Finished;
        }
        be->stats.arena_stalls++;
        return Start::Retry;  // wait for the consumer to drop chunks
    }
    auto buf = std::make_shared<InnerBuffer>();
    buf->buf = block;
    buf->size = r.total_bytes;
    buf->align = kAlign;
    buf->kind = MemKind::Device;
    buf->device = g.device;
    buf->release = [arena = g.arena, block](InnerBuffer&) { arena->release(block); };
    r.buffer = std::move(buf);

    // Read plan: sort the ranges' padded spans, merge spans that touch / overlap / sit within
    // merge_gap_bytes of each other into runs, cut every run into reads of at most max_read_bytes
    // (reference: one read per range, panics above 2 GiB, get_range.rs:23-30).
    const uint64_t max_read = std::max<uint64_t>(kAlign, cfg.max_read_bytes / kAlign * kAlign);
    std::vector<uint32_t> order;
    for (size_t i = 0; i < n; i++)
        if (r.ranges[i].valid) order.push_back((uint32_t)i);
    auto span_lo = [&](uint32_t i) { return (uint64_t)r.ranges[i].start_off / kAlign * kAlign; };                   // sqe.rs:83
    auto span_hi = [&](uint32_t i) { return ((uint64_t)r.ranges[i].end_off + kAlign - 1) / kAlign * kAlign; };     // padded length
    if (cfg.merge_ranges) std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return span_lo(a) < span_lo(b); });
    for (size_t k = 0; k < order.size();) {
        uint64_t run_lo = span_lo(order[k]), run_hi = span_hi(order[k]);
        size_t k_end = k + 1;
        if (cfg.merge_ranges)
            while (k_end < order.size() && k_end - k < 1024 && span_lo(order[k_end]) <= run_hi + cfg.merge_gap_bytes) {  // <= 1024 segments per read: a read must fit one batch
                run_hi = std::max(run_hi, span_hi(order[k_end]));
                k_end++;
            }
        for (uint64_t fo = run_lo; fo < run_hi; fo += max_read) {
            ReadPlan rd;
            rd.file_off = fo;
            rd.read_len = (uint32_t)std::min<uint64_t>(max_read, run_hi - fo);
            rd.first_seg = (uint32_t)r.segs.size();
            for (size_t j = k; j < k_end; j++) {
                RangePlan& p = r.ranges[order[j]];
                const uint64_t lo = std::max<uint64_t>(fo, (uint64_t)p.start_off), hi = std::min<uint64_t>(fo + rd.read_len, (uint64_t)p.end_off);
                if (lo >= hi) continue;
                SegPlan sp;
                sp.range = order[j];
                sp.head = (uint32_t)(lo - fo);  // sqe.rs:116: the view starts at start_off - aligned_start
                sp.payload_len = hi - lo;
                sp.dst_off = p.dst_off + (lo - (uint64_t)p.start_off);
                p.seg_ids.push_back((uint32_t)r.segs.size());
                rd.needed = std::max<uint32_t>(rd.needed, (uint32_t)(hi - fo));
                r.segs.push_back(sp);
            }
            rd.nsegs = (uint32_t)r.segs.size() - rd.first_seg;
            if (rd.nsegs) r.reads.push_back(rd);  // a read that carries no payload (pure gap) is skipped
        }
        k = k_end;
    }
    for (size_t i = 0; i < n; i++)
        if (r.ranges[i].valid) r.ranges[i].segs_left.store((uint32_t)r.ranges[i].seg_ids.size(), std::memory_order_relaxed);
    r.reads_left.store((uint32_t)r.reads.size(), std::memory_order_release);
    // get_ranges.rs:39-45: one read op per segment onto THIS worker's deque; peers steal them.
    for (uint32_t s = 0; s < r.reads.size(); s++) {
        Op op;
        op.kind = Op::Read;
        op.req = rp;
        op.read = s;
        wt.push(std::move(op));
    }
    return Start::Finished;
}

// ---- io worker (UringWorker, worker.rs:24-135) ------------------------------------------------
struct Worker {
    Group& g;
    WorkerThread<Op>& wt;
    Ring ring;
    Tracker<Op> tracker;
    // Ops that could not start: reads waiting for staging space (strictly FIFO; nothing new is
    // pulled while one waits) and plans waiting for HBM (must NOT hold back reads: it is the
    // completion and consumption of earlier reads that frees the arena).
    std::deque<Op> stalled_reads, stalled_plans;
    WorkerSlots* slots = nullptr;  // staging_mode 1
    Worker(Group& grp, WorkerThread<Op>& w) : g(grp), wt(w), tracker(grp.be->cfg.sq_entries) {
        if (grp.be->cfg.staging_mode == 1) {
            slots = &grp.wslots[w.id()];
            extents.resize(slots->busy.size());
        }
    }

    // staging_mode 1: per-slot extent bookkeeping (this worker only)
    struct Extent {
        StageEntry* entry = nullptr;
        uint32_t pending = 0;  // reads submitted but not yet landed/failed
        bool closed = false;   // no more reads will be added
    };
    std::vector<Extent> extents;  // one per slot
    int cur = -1;                 // slot of the extent being filled, or -1

    // The extent's reads have all landed (or failed) and nothing more will be added: copy it to the device log
    // with ONE cudaMemcpyAsync while it is still in this core's cache, then publish it to the device thread.
    void finish_extent(int slot) {
        Extent& x = extents[slot];
        StageEntry* ent = x.entry;
        Backend* be = g.be;
        bool any = false;
        for (auto& r : ent->reads) any |= r.state == kStageLanded;
        if (any) {
            const uint64_t phys = ent->log_off % g.log.capacity();
            cudaError_t ce = cudaMemcpyAsync(g.dev_ring + phys, slots->base + (uint64_t)slot * g.slot_bytes, ent->used, cudaMemcpyHostToDevice, slots->stream);
            if (ce == cudaSuccess) ce = cudaEventRecord(slots->ev[slot], slots->stream);
            if (ce != cudaSuccess) {
                cudaGetLastError();
                for (auto& r : ent->reads) {
                    if (r.state != kStageLanded) continue;
                    r.state = kStageFailed;
                    const ReadPlan& rd = r.req->reads[r.read];
                    for (uint32_t k = 0; k < rd.nsegs; k++) {
                        const uint32_t range = r.req->segs[rd.first_seg + k].range;
                        if (!r.req->ranges[range].failed.exchange(true)) send_error(be, (int)ce, 0, *r.req, (int)range, "cudaMemcpyAsync (H2D) failed");
                    }
                }
                slots->busy[slot] = 0;
            } else {
                slots->busy[slot] = 2;
                ent->copy_event = slots->ev[slot];
                ent->owner = (uint32_t)wt.id();
                be->stats.h2d_copies++;
                be->stats.h2d_bytes += ent->used;
            }
        } else {
            slots->busy[slot] = 0;
        }
        x.entry = nullptr;
        ent->landed_ns = now_ns();
        ent->state.store(kStageLanded, std::memory_order_release);
    }
    void close_current_extent() {
        if (cur < 0) return;
        const int slot = cur;
        cur = -1;
        extents[slot].closed = true;
        if (extents[slot].pending == 0) finish_extent(slot);
    }

    // staging_mode 1: a free slot of this worker, recycling slots whose H2D copy has finished; -1 if none
    int acquire_slot() {
        for (size_t i = 0; i < slots->busy.size(); i++)
            if (slots->busy[i] == 2 && cudaEventQuery(slots->ev[i]) != cudaErrorNotReady) slots->busy[i] = 0;
        for (size_t i = 0; i < slots->busy.size(); i++)
            if (slots->busy[i] == 0) return (int)i;
        return -1;
    }

    // An SQE slot, flushing the SQ first if a burst of completions filled it.
    io_uring_sqe* must_get_sqe() {
        io_uring_sqe* e;
        while (!(e = ring.get_sqe())) {
            ring.submit();
            __builtin_ia32_pause();
        }
        return e;
    }

    void push_read_sqe(size_t idx, Op& op) {
        const ReadPlan& sp = op.req->reads[op.read];
        io_uring_sqe* e = must_get_sqe();
        e->opcode = IORING_OP_READ_FIXED;  // north star: READ_FIXED into registered pinned memory (reference: Read, sqe.rs:95)
        e->fd = op.req->fd;
        if (slots) {
            e->addr = (uint64_t)(uintptr_t)(slots->base + (uint64_t)op.slot * g.slot_bytes + op.entry->reads[op.ref].off + op.done);
            e->buf_index = 0;
        } else {
            const uint64_t phys = op.entry->log_off % g.log.capacity() + op.done;
            e->addr = (uint64_t)(uintptr_t)(g.host_ring + phys);
            e->buf_index = (uint16_t)(phys / StagingLog::kRegBoundary);
        }
        e->len = sp.read_len - op.done;
        e->off = sp.file_off + op.done;
        e->user_data = UringUserData::pack((uint32_t)idx, IORING_OP_READ_FIXED);
    }

    Start start(Op& op) {
        Backend* be = g.be;
        switch (op.kind) {
        case Op::GetRanges: {  // get_ranges.rs:49-65: two unlinked SQEs, same slot
            auto idx = tracker.get_next_index();
            Request& r = *op.req;
            memset(&r.stx, 0, sizeof r.stx);
            io_uring_sqe* e = must_get_sqe();  // sqe.rs:17-28
            e->opcode = IORING_OP_OPENAT;
            e->fd = -1;  // dirfd ignored for absolute paths
            e->addr = (uint64_t)(uintptr_t)r.path.c_str();
            e->open_flags = O_RDONLY | (be->cfg.o_direct ? O_DIRECT : 0);
            e->user_data = UringUserData::pack((uint32_t)*idx, IORING_OP_OPENAT);
            if (be->cfg.statx_via_ring) {
                e = must_get_sqe();  // sqe.rs:40-57
                e->opcode = IORING_OP_STATX;
                e->fd = -1;
                e->addr = (uint64_t)(uintptr_t)r.path.c_str();
                e->off = (uint64_t)(uintptr_t)&r.stx;
                e->len = STATX_SIZE | STATX_DIOALIGN;
                e->user_data = UringUserData::pack((uint32_t)*idx, IORING_OP_STATX);
            }
            tracker.put(*idx, std::move(op));
            be->stats.requests++;
            return Start::Started;
        }
        case Op::Plan:
            return plan_request(g, op.req, wt);
        case Op::Read: {  // get_range.rs:43-51 + sqe.rs:59-122
            const ReadPlan& sp = op.req->reads[op.read];
            StageEntry* ent;
            uint32_t off = 0;
            int slot = -1;
            if (slots) {
                // pack this read behind the previous ones of the current extent, or open a new extent
                if (cur >= 0 && extents[cur].entry->used + sp.read_len > g.slot_bytes) close_current_extent();
                if (cur < 0) {
                    const int free_slot = acquire_slot();
                    if (free_slot < 0) {  // all of this worker's slots are still being read or copied
                        be->stats.staging_stalls++;
                        return Start::Retry;
                    }
                    StageEntry* fresh = g.log.alloc((uint32_t)g.slot_bytes);
                    if (!fresh) {
                        be->stats.staging_stalls++;
                        return Start::Retry;
                    }
                    slots->busy[free_slot] = 1;
                    extents[free_slot] = Extent{fresh, 0, false};
                    cur = free_slot;
                }
                slot = cur;
                ent = extents[cur].entry;
                off = ent->used;
                ent->used += sp.read_len;
                extents[cur].pending++;
            } else {
                ent = g.log.alloc(sp.read_len);
                if (!ent) {
                    be->stats.staging_stalls++;
                    return Start::Retry;
                }
            }
            op.ref = (uint32_t)ent->reads.size();
            ent->reads.push_back(ReadRefT<Request>{op.req, op.read, off, kStageReading});
            op.entry = ent;
            op.done = 0;
            op.slot = slot;
            auto idx = tracker.get_next_index();
            push_read_sqe(*idx, op);
            tracker.put(*idx, std::move(op));
            be->stats.reads++;
            return Start::Started;
        }
        case Op::Close:
            break;
        }
        return Start::Finished;
    }

    // A read has landed or failed: record it in its extent and publish the extent when it is complete.
    void read_done(Op& op, uint32_t state) {
        StageEntry* ent = op.entry;
        ent->reads[op.ref].state = state;
        if (slots) {
            Extent& x = extents[op.slot];
            x.pending--;
            if (x.closed && x.pending == 0) finish_extent(op.slot);
        } else {
            ent->landed_ns = now_ns();
            ent->state.store(kStageLanded, std::memory_order_release);  // staging_mode 0: one read per extent
        }
    }

    void finish_read_accounting(size_t idx, Op& op) {
        // get_range.rs:78-90: the last reader of the file closes it, re-using this tracker slot
        Request& r = *op.req;
        if (r.reads_left.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            io_uring_sqe* e = must_get_sqe();  // sqe.rs:126-133
            e->opcode = IORING_OP_CLOSE;
            e->fd = r.fd;
            e->user_data = UringUserData::pack((uint32_t)idx, IORING_OP_CLOSE);
            r.fd = -1;
            Op c;
            c.kind = Op::Close;
            c.req = op.req;
            tracker.replace(idx, std::move(c));  // NextStep::ReplaceWith
            return;
        }
        tracker.remove(idx);  // NextStep::Done
    }

    void handle(const io_uring_cqe& cqe) {
        Backend* be = g.be;
        const UringUserData ud = UringUserData::unpack(cqe.user_data);
        Op* op = tracker.get(ud.index_of_op);
        if (!op) return;
        Request& r = *op->req;
        switch (op->kind) {
        case Op::GetRanges: {  // get_ranges.rs:67-110
            if (cqe.res < 0) send_error(be, cqe.res, ud.opcode, r, -1, ud.opcode == IORING_OP_OPENAT ? "openat failed" : "statx failed");
            r.n_cqes_received++;
            if (cqe.res >= 0) {
                if (ud.opcode == IORING_OP_OPENAT) {
                    r.fd = cqe.res;
                    r.have_fd = true;
                } else if (ud.opcode == IORING_OP_STATX) {
                    r.have_statx = true;
                }
            }
            if (!be->cfg.statx_via_ring) {  // fast metadata path: size from fstat on the fd we just opened
                if (r.have_fd) {
                    struct stat st;
                    if (fstat(r.fd, &st) == 0) {
                        r.stx.stx_size = (uint64_t)st.st_size;
                        r.have_statx = true;
                    } else {
                        send_error(be, -errno, IORING_OP_STATX, r, -1, "fstat failed");
                    }
                }
            } else if (r.n_cqes_received < 2) {
                return;  // NextStep::Pending
            }
            Op done = tracker.remove(ud.index_of_op);
            if (r.have_fd && r.have_statx) {
                done.kind = Op::Plan;
                if (plan_request(g, done.req, wt) == Start::Retry) stalled_plans.push_back(std::move(done));
            } else if (r.have_fd) {
                close(r.fd);  // the reference leaks this fd (Appendix A)
                r.fd = -1;
            }
            return;
        }
        case Op::Read: {  // get_range.rs:53-91
            const ReadPlan& rd = r.reads[op->read];
            auto fail = [&](int status, const char* what) {
                // one Err per range that this CQE was carrying (reference: one Err per failed CQE = per range)
                uint32_t last = UINT32_MAX;
                for (uint32_t k = 0; k < rd.nsegs; k++) {
                    const uint32_t range = r.segs[rd.first_seg + k].range;
                    if (range == last) continue;
                    last = range;
                    if (!r.ranges[range].failed.exchange(true, std::memory_order_relaxed)) send_error(be, status, IORING_OP_READ_FIXED, r, (int)range, what);
                }
                read_done(*op, kStageFailed);
                finish_read_accounting(ud.index_of_op, *op);
            };
            if (cqe.res < 0) {
                fail(cqe.res, "read failed");
                return;
            }
            op->done += (uint32_t)cqe.res;
            be->stats.read_bytes += (uint64_t)cqe.res;
            if (op->done < rd.needed) {  // the reference does not check (get_range.rs:66-68, issue #100)
                if (cqe.res == 0 || (op->done % kAlign) != 0) {
                    fail(-EIO, "short read (unexpected end of file)");
                    return;
                }
                be->stats.short_reads++;
                push_read_sqe(ud.index_of_op, *op);  // read the remainder into the same slot
                return;
            }
            read_done(*op, kStageLanded);
            finish_read_accounting(ud.index_of_op, *op);
            return;
        }
        case Op::Close:  // close.rs:33-44
            tracker.remove(ud.index_of_op);
            return;
        case Op::Plan:
            return;
        }
    }

    void run() {
        Backend* be = g.be;
        const lsio_uring_cuda_config& cfg = be->cfg;
        int rc = ring.init(cfg.sq_entries, cfg.sqpoll ? 1000 : 0);
        if (rc == 0) {
            std::vector<iovec> iov;
            if (slots) {
                cudaSetDevice(g.device);  // this worker issues its own H2D copies
                iov.push_back(iovec{slots->base, (size_t)(g.slot_bytes * slots->busy.size())});
            } else {
                for (uint64_t o = 0; o < g.log.capacity(); o += StagingLog::kRegBoundary)
                    iov.push_back(iovec{g.host_ring + o, (size_t)std::min<uint64_t>(StagingLog::kRegBoundary, g.log.capacity() - o)});
            }
            rc = ring.register_buffers(iov.data(), (unsigned)iov.size());
        }
        if (rc != 0) {
            fprintf(stderr, "lsio_b200: io_uring set-up failed on worker %zu: %s\n", wt.id(), strerror(-rc));
            // keep consuming tasks so callers see errors rather than a hang
        }
        const bool ring_ok = rc == 0;
        const unsigned sq = ring_ok ? ring.sq_entries() : 0;
        while (true) {
            const bool running = wt.keep_running();
            if (!running && tracker.is_empty()) break;
            unsigned started = 0;
            if (running) {
                auto room = [&] { return !ring_ok || !(tracker.is_full() || ring.sq_len() + 2 > sq || ring.sq_len() + ring.cq_len() + 2 > sq); };
                // 1. reads that waited for staging space, in order
                while (!stalled_reads.empty() && room()) {
                    const Start s = start(stalled_reads.front());
                    if (s == Start::Retry) break;
                    stalled_reads.pop_front();
                    if (s == Start::Started) started++;
                }
                // 2. plans that waited for HBM, each once
                for (size_t n = stalled_plans.size(); n-- > 0;) {
                    Op op = std::move(stalled_plans.front());
                    stalled_plans.pop_front();
                    if (start(op) == Start::Retry) stalled_plans.push_back(std::move(op));
                    else started++;
                }
                // 3. top up the SQ with new tasks (worker.rs:65-88), two SQEs of headroom (MAX_SQ_ENTRIES_PER_ITERATION)
                bool ran_dry = false;
                while (stalled_reads.empty() && room()) {
                    auto t = wt.find_task();
                    if (!t) {
                        ran_dry = true;
                        break;
                    }
                    Op op = std::move(*t);
                    if (!ring_ok) {
                        send_error(be, rc, 0, *op.req, -1, "io_uring unavailable");
                        continue;
                    }
                    const Start s = start(op);
                    if (s == Start::Retry) (op.kind == Op::Read ? stalled_reads : stalled_plans).push_back(std::move(op));
                    else if (s == Start::Started) started++;
                }
                // nothing more to add right now: let the partly filled extent go (bounds latency, ends a workload)
                if (slots && ran_dry && stalled_reads.empty()) close_current_extent();
            } else if (slots) {
                close_current_extent();
            }
            if (ring_ok && ring.unsubmitted()) ring.submit();  // one enter per batch (reference: per op, worker.rs:79)
            bool got = false;
            io_uring_cqe cqe;
            while (ring_ok && ring.pop_cqe(&cqe)) {  // worker.rs:104-122
                got = true;
                handle(cqe);
            }
            if (ring_ok && ring.unsubmitted()) ring.submit();
            if (got || started) continue;
            if (tracker.is_empty() && stalled_reads.empty() && stalled_plans.empty()) {
                if (running) wt.park();  // worker.rs:93-99
                continue;
            }
            if (!tracker.is_empty()) ring.submit(1, 200);  // submit_and_wait(1), bounded so new tasks are noticed
            else std::this_thread::sleep_for(std::chrono::microseconds(50));  // only waiting for staging / HBM space
        }
    }
};

// ---- device thread -----------------------------------------------------------------------------
static uint32_t range_crc(const Request& r, const RangePlan& p) {
    uint32_t crc = 0;
    for (size_t s = 0; s < p.seg_ids.size(); s++) {
        const SegPlan& sp = r.segs[p.seg_ids[s]];
        crc = s == 0 ? sp.crc : gf2::crc32c_combine(crc, sp.crc, sp.payload_len);
    }
    return crc;
}

static void send_chunk(Group& g, const std::shared_ptr<Request>& rp, uint32_t range_idx, uint32_t crc, bool have_xxh, uint64_t xxh) {
    Backend* be = g.be;
    Request& r = *rp;
    RangePlan& p = r.ranges[range_idx];
    lsio_output o;
    memset(&o, 0, sizeof o);
    o.status = 0;
    o.kind = LSIO_OUTPUT_CHUNK;
    o.device = g.device;
    if (be->cfg.checksum & LSIO_CHECKSUM_CRC32C) {
        o.checksums |= LSIO_CHECKSUM_CRC32C;
        o.crc32c = crc;
    }
    if (have_xxh) {
        o.checksums |= LSIO_CHECKSUM_XXH64;
        o.xxh64 = xxh;
    }
    auto* view = new lsio_aligned_bytes();
    view->buf = r.buffer;
    view->start = p.dst_off;
    view->end = p.dst_off + p.len;
    o.chunk.buffer = view;
    o.chunk.user_data = p.user_data;
    be->stats.payload_bytes += p.len;
    g.out_pending.push_back(o);  // flushed by flush_outputs(): one lock + one wake-up per retired batch, not per chunk
}

static void flush_outputs(Group& g) {
    if (g.out_pending.empty()) return;
    g.be->stats.chunks_ok += g.out_pending.size();
    g.be->out->send_many(g.out_pending.data(), g.out_pending.size());
    g.out_pending.clear();
}

// All segments of a range are packed: emit its chunk, or queue it for the XXH64 stage.
static void emit_range(Group& g, const std::shared_ptr<Request>& rp, uint32_t range_idx) {
    RangePlan& p = rp->ranges[range_idx];
    if (p.failed.load(std::memory_order_relaxed)) return;  // its errors were already sent; no Chunk (get_range.rs:63)
    const uint32_t crc = (g.be->cfg.checksum & LSIO_CHECKSUM_CRC32C) ? range_crc(*rp, p) : 0;
    if (g.be->cfg.checksum & LSIO_CHECKSUM_XXH64) g.xxh_pending.push_back(XxhBatch::Item{rp, range_idx, crc});
    else send_chunk(g, rp, range_idx, crc, false, 0);
}

static int flush_xxh(Group& g) {
    while (!g.xxh_pending.empty() && !g.xxh_free.empty()) {
        XxhBatch* xb = g.xxh_free.front();
        g.xxh_free.pop_front();
        const size_t n = std::min(g.xxh_pending.size(), XxhBatch::kMax);
        xb->items.assign(std::make_move_iterator(g.xxh_pending.begin()), std::make_move_iterator(g.xxh_pending.begin() + n));
        g.xxh_pending.erase(g.xxh_pending.begin(), g.xxh_pending.begin() + n);
        for (size_t i = 0; i < n; i++) {
            const RangePlan& p = xb->items[i].req->ranges[xb->items[i].range];
            xb->ptrs[i] = (uint64_t)(uintptr_t)(xb->items[i].req->buffer->buf + p.dst_off);
            xb->ptrs[XxhBatch::kMax + i] = p.len;
        }
        int rc = xxh64_launch_async(xb->dev, xb->dev + XxhBatch::kMax, n, xb->dev + 2 * XxhBatch::kMax, g.s_pack);
        if (rc) return rc;
        g.be->stats.kernel_launches++;
        LSIO_CUDA_TRY(cudaEventRecord(xb->done, g.s_pack));
        g.xxh_inflight.push_back(xb);
    }
    return 0;
}

static void retire_batch(Group& g, Batch& b) {
    // failed reads carry no descriptor but their segments still count towards their ranges
    for (uint64_t s = b.seq_begin; s < b.seq_end; s++) {
        StageEntry& e = g.log.at(s);
        for (auto& ref : e.reads) {
            if (ref.state != kStageFailed) continue;
            const ReadPlan& rd = ref.req->reads[ref.read];
            for (uint32_t k = 0; k < rd.nsegs; k++) ref.req->ranges[ref.req->segs[rd.first_seg + k].range].segs_left.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    for (uint32_t i = 0; i < b.counts.nsegs; i++) {
        const std::shared_ptr<Request>& rp = g.log.at(b.desc[i].seq).reads[b.desc[i].ref].req;
        SegPlan& sp = rp->segs[b.desc[i].seg];
        sp.crc = b.crc_host[i];
        RangePlan& p = rp->ranges[sp.range];
        if (p.segs_left.fetch_sub(1, std::memory_order_acq_rel) == 1) emit_range(g, rp, sp.range);
    }
    flush_outputs(g);
    g.log.retire_through(b.seq_end);
}

static int flush_run(Group& g, uint64_t seq_begin, uint64_t seq_end) {
    Backend* be = g.be;
    Batch* b = g.free_batches.front();
    g.free_batches.pop_front();
    g.last_flush_batch = b;
    b->counts = PackBatchCounts();
    b->desc.clear();
    b->seq_begin = seq_begin;
    b->seq_end = seq_end;
    const uint64_t cap = g.log.capacity();
    const StageEntry& first = g.log.at(seq_begin);
    const StageEntry& last = g.log.at(seq_end - 1);
    const uint64_t phys = first.log_off % cap, bytes = last.log_off + last.len - first.log_off;
    const bool store = true;
    const uint32_t flags = LSIO_CUDA_PACK_STORE | ((be->cfg.checksum & LSIO_CHECKSUM_CRC32C) ? LSIO_CUDA_PACK_CRC32C : 0);
    uint64_t landed_bytes = 0;
    std::vector<cudaEvent_t>& copy_events = b->copy_events;  // staging_mode 1: the H2D copy of every extent in this run
    copy_events.clear();
    for (uint64_t s = seq_begin; s < seq_end; s++) {
        StageEntry& e = g.log.at(s);
        for (uint32_t ri = 0; ri < e.reads.size(); ri++) {
            auto& ref = e.reads[ri];
            if (ref.state != kStageLanded) continue;
            const ReadPlan& rd = ref.req->reads[ref.read];
            for (uint32_t k = 0; k < rd.nsegs; k++) {
                const SegPlan& sp = ref.req->segs[rd.first_seg + k];
                lsio_cuda_segment seg;
                seg.src = (uint64_t)(uintptr_t)((be->cfg.zero_copy ? g.host_ring_dev : g.dev_ring) + e.log_off % cap + ref.off + sp.head);
                seg.dst = (uint64_t)(uintptr_t)(ref.req->buffer->buf + sp.dst_off);
                seg.len = sp.payload_len;
                if (!pack_batch_append(b->lay, b->host_desc, &b->counts, seg, store)) {
                    // cannot happen: device_main sizes every run with the same tile arithmetic
                    RangePlan& p = ref.req->ranges[sp.range];
                    if (!p.failed.exchange(true)) send_error(be, -EOVERFLOW, 0, *ref.req, (int)sp.range, "internal: pack batch overflow");
                    p.segs_left.fetch_sub(1, std::memory_order_acq_rel);
                    continue;
                }
                b->desc.push_back(Batch::DescRef{s, ri, rd.first_seg + k});
            }
            landed_bytes += rd.read_len;
        }
        if (e.copy_event) copy_events.push_back((cudaEvent_t)e.copy_event);
    }
    if (landed_bytes && be->cfg.staging_mode == 1) {
        // The extents of this run were copied by their workers: the pack stream waits for EVERY one of them.
        // (Not just the last extent per worker in log order: with asynchronous reads an extent allocated
        // earlier can complete -- and be copied -- after a later one of the same worker.  A slot's event is
        // re-recorded only after its previous copy finished, so a wait never under-covers.)
        for (cudaEvent_t ev : copy_events) LSIO_CUDA_TRY(cudaStreamWaitEvent(g.s_pack, ev, 0));
    }
    if (landed_bytes) {
        if (!be->cfg.zero_copy && be->cfg.staging_mode == 0) {
            LSIO_CUDA_TRY(cudaMemcpyAsync(g.dev_ring + phys, g.host_ring + phys, bytes, cudaMemcpyHostToDevice, g.s_copy));
            LSIO_CUDA_TRY(cudaEventRecord(b->copied, g.s_copy));
            be->stats.h2d_copies++;
            be->stats.h2d_bytes += bytes;
            LSIO_CUDA_TRY(cudaStreamWaitEvent(g.s_pack, b->copied, 0));
        }
        LSIO_CUDA_TRY(cudaMemcpyAsync(b->dev_desc, b->host_desc, b->lay.upload_bytes, cudaMemcpyHostToDevice, g.s_pack));
        unsigned launches = 0;
        int rc = pack_batch_launch(g.device, b->lay, b->dev_desc, b->counts, flags, b->crc_dev, g.s_pack, &launches);
        if (rc) return rc;
        be->stats.pack_launches++;
        be->stats.kernel_launches += launches;
    }
    LSIO_CUDA_TRY(cudaEventRecord(b->done, g.s_pack));
    g.inflight.push_back(b);
    return 0;
}

static void device_main(Group& g) {
    Backend* be = g.be;
    const lsio_uring_cuda_config& cfg = be->cfg;
    if (!g.cpus.empty()) {
        cpu_set_t set;
        CPU_ZERO(&set);
        for (int c : g.cpus) CPU_SET(c, &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set);
    }
    cudaSetDevice(g.device);
    const uint64_t cap = g.log.capacity();
    const size_t max_segs = g.batches[0].lay.max_segs, max_tiles = g.batches[0].lay.max_tiles;
    const uint64_t max_batch_bytes = std::max<uint64_t>(cfg.min_copy_bytes * 4, 64ull << 20);
    unsigned idle_spins = 0;
    while (true) {
        bool progressed = false;
        // 1. retire finished batches in order
        while (!g.inflight.empty()) {
            Batch* b = g.inflight.front();
            cudaError_t q = cudaEventQuery(b->done);
            if (q == cudaErrorNotReady) break;
            if (q != cudaSuccess) {
                cuda_fail(q, "pack batch", __FILE__, __LINE__);
                fprintf(stderr, "lsio_b200: %s\n", last_error_text().c_str());
                for (uint32_t i = 0; i < b->counts.nsegs; i++) {
                    Request& rq = *g.log.at(b->desc[i].seq).reads[b->desc[i].ref].req;
                    const uint32_t range = rq.segs[b->desc[i].seg].range;
                    if (!rq.ranges[range].failed.exchange(true)) send_error(be, (int)q, 0, rq, (int)range, "CUDA error in the H2D/pack stage");
                }
            }
            g.inflight.pop_front();
            retire_batch(g, *b);
            g.free_batches.push_back(b);
            progressed = true;
        }
        // 1b. XXH64 stage: finished launches emit their chunks; packed ranges waiting for one are launched
        while (!g.xxh_inflight.empty()) {
            XxhBatch* xb = g.xxh_inflight.front();
            cudaError_t q = cudaEventQuery(xb->done);
            if (q == cudaErrorNotReady) break;
            g.xxh_inflight.pop_front();
            for (size_t i = 0; i < xb->items.size(); i++) {
                if (q == cudaSuccess) send_chunk(g, xb->items[i].req, xb->items[i].range, xb->items[i].crc, true, xb->ptrs[2 * XxhBatch::kMax + i]);
                else send_error(be, (int)q, 0, *xb->items[i].req, (int)xb->items[i].range, "CUDA error in the xxh64 stage");
            }
            flush_outputs(g);
            xb->items.clear();
            g.xxh_free.push_back(xb);
            progressed = true;
        }
        if (!g.xxh_pending.empty() && !g.xxh_free.empty()) {
            if (flush_xxh(g)) fprintf(stderr, "lsio_b200: xxh64 launch failed: %s\n", last_error_text().c_str());
            progressed = true;
        }
        // 2. extend the copy frontier over landed reads
        const uint64_t avail = g.log.alloc_seq();
        if (g.copy_seq < avail && !g.free_batches.empty()) {
            uint64_t s = g.copy_seq, bytes = 0, run_phys0 = g.log.at(s).log_off % cap;
            size_t nsegs = 0, ntiles = 0;
            uint64_t oldest = 0;
            bool hit_limit = false;
            for (; s < avail; s++) {
                StageEntry& e = g.log.at(s);
                const uint32_t st = e.state.load(std::memory_order_acquire);
                if (st == kStageReading) break;
                if (e.log_off % cap < run_phys0) {  // the ring wrapped: a copy must be physically contiguous
                    hit_limit = true;
                    break;
                }
                size_t t = 0;
                size_t ns = 0;
                for (auto& ref : e.reads) {
                    if (ref.state != kStageLanded) continue;
                    const ReadPlan& rd = ref.req->reads[ref.read];
                    for (uint32_t k = 0; k < rd.nsegs; k++) {
                        const SegPlan& sp = ref.req->segs[rd.first_seg + k];
                        t += pack_tile_count((uint64_t)(uintptr_t)(ref.req->buffer->buf + sp.dst_off), sp.payload_len);
                    }
                    ns += rd.nsegs;
                    if (!oldest) oldest = e.landed_ns;
                }
                if (nsegs + ns > max_segs || ntiles + t > max_tiles || (bytes && bytes + e.len > max_batch_bytes)) {
                    hit_limit = true;
                    break;
                }
                nsegs += ns;
                ntiles += t;
                bytes = e.log_off + e.len - g.log.at(g.copy_seq).log_off;
            }
            if (s > g.copy_seq) {
                const bool draining = g.stop_device.load(std::memory_order_relaxed);
                const bool timed_out = oldest && now_ns() - oldest > (uint64_t)cfg.flush_us * 1000;
                if (hit_limit || bytes >= cfg.min_copy_bytes || timed_out || draining || (!oldest)) {
                    int rc = flush_run(g, g.copy_seq, s);
                    if (rc) {
                        // A CUDA call failed while enqueueing: report every range of the run in-band and
                        // retire it, so that callers see errors instead of a hang.
                        fprintf(stderr, "lsio_b200: flush failed: %s\n", last_error_text().c_str());
                        Batch* b = g.last_flush_batch;  // popped from the free list by flush_run, never reached `inflight`
                        for (uint64_t q = g.copy_seq; q < s; q++) {
                            StageEntry& e = g.log.at(q);
                            for (auto& ref : e.reads) {
                                if (ref.state != kStageLanded) continue;
                                const ReadPlan& rd = ref.req->reads[ref.read];
                                for (uint32_t k = 0; k < rd.nsegs; k++) {
                                    const uint32_t range = ref.req->segs[rd.first_seg + k].range;
                                    if (!ref.req->ranges[range].failed.exchange(true)) send_error(be, rc, 0, *ref.req, (int)range, "CUDA error while enqueueing the H2D/pack stage");
                                }
                            }
                        }
                        cudaStreamSynchronize(g.s_copy);
                        cudaStreamSynchronize(g.s_pack);
                        cudaGetLastError();
                        retire_batch(g, *b);  // its ranges are marked failed: nothing is emitted, the slots are freed
                        g.free_batches.push_back(b);
                    }
                    g.copy_seq = s;
                    progressed = true;
                }
            }
        }
        if (progressed) {
            idle_spins = 0;
            continue;
        }
        if (g.stop_device.load(std::memory_order_acquire) && g.workers_alive.load(std::memory_order_acquire) == 0 && g.inflight.empty() &&
            g.xxh_inflight.empty() && g.xxh_pending.empty() && g.copy_seq == g.log.alloc_seq())
            break;
        // 3. wait: spin briefly while work is in flight, sleep when idle
        if (g.log.empty() && g.inflight.empty() && g.xxh_inflight.empty()) std::this_thread::sleep_for(std::chrono::microseconds(100));
        else if (cfg.poll_sleep_us) std::this_thread::sleep_for(std::chrono::microseconds(cfg.poll_sleep_us));
        else if (++idle_spins > 200) std::this_thread::sleep_for(std::chrono::microseconds(20));
        else
            for (int i = 0; i < 50; i++) __builtin_ia32_pause();
    }
}

// ---- construction / teardown ----------------------------------------------------------------------
static std::vector<int> numa_cpus_of_device(int device) {
    std::vector<int> cpus;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return cpus;
    std::string b(bus);
    std::transform(b.begin(), b.end(), b.begin(), ::tolower);
    int node = -1;
    {
        std::ifstream f("/sys/bus/pci/devices/" + b + "/numa_node");
        if (!(f >> node)) node = -1;
    }
    if (node < 0) return cpus;
    std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string list;
    if (!std::getline(f, list)) return cpus;
    std::stringstream ss(list);
    std::string part;
    while (std::getline(ss, part, ',')) {
        int a, z;
        if (sscanf(part.c_str(), "%d-%d", &a, &z) == 2) for (int c = a; c <= z; c++) cpus.push_back(c);
        else if (sscanf(part.c_str(), "%d", &a) == 1) cpus.push_back(a);
    }
    return cpus;
}

static void free_group_cuda(Group& g) {
    DeviceGuard dg(g.device);
    for (Batch& b : g.batches) {
        if (b.host_desc) cudaFreeHost(b.host_desc);
        if (b.dev_desc) cudaFree(b.dev_desc);
        if (b.crc_host) cudaFreeHost(b.crc_host);
        if (b.copied) cudaEventDestroy(b.copied);
        if (b.done) cudaEventDestroy(b.done);
    }
    g.batches.clear();
    for (XxhBatch& x : g.xxh_batches) {
        if (x.ptrs) cudaFreeHost(x.ptrs);
        if (x.done) cudaEventDestroy(x.done);
    }
    g.xxh_batches.clear();
    if (g.s_copy) cudaStreamDestroy(g.s_copy);
    if (g.s_pack) cudaStreamDestroy(g.s_pack);
    for (WorkerSlots& ws : g.wslots) {
        for (auto ev : ws.ev)
            if (ev) cudaEventDestroy(ev);
        if (ws.stream) cudaStreamDestroy(ws.stream);
    }
    g.wslots.clear();
    if (g.slot_block) cudaFreeHost(g.slot_block);
    g.slot_block = nullptr;
    if (g.host_ring) cudaFreeHost(g.host_ring);
    if (g.dev_ring) cudaFree(g.dev_ring);
    g.s_copy = g.s_pack = nullptr;
    g.host_ring = g.dev_ring = nullptr;
}

static int init_group(Group& g) {
    const lsio_uring_cuda_config& cfg = g.be->cfg;
    DeviceGuard dg(g.device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    int rc = pack_prepare_device(g.device);
    if (rc) return rc;
    if (cfg.pin_threads) g.cpus = numa_cpus_of_device(g.device);
    const uint64_t ring_bytes = cfg.staging_bytes / 4096 * 4096;
    if (cfg.staging_mode == 1) {
        // per-worker cache-sized windows of pinned slots; the ordered log lives on the device only
        g.slot_bytes = std::max<uint64_t>((uint64_t)cfg.host_slot_kib << 10, (cfg.max_read_bytes + 4095) / 4096 * 4096);
        const uint64_t per_worker = g.slot_bytes * cfg.host_slots_per_worker;
        LSIO_CUDA_TRY(cudaHostAlloc((void**)&g.slot_block, per_worker * cfg.n_worker_threads, cudaHostAllocPortable));
        g.wslots.resize(cfg.n_worker_threads);
        for (uint32_t w = 0; w < cfg.n_worker_threads; w++) {
            WorkerSlots& ws = g.wslots[w];
            ws.base = g.slot_block + per_worker * w;
            LSIO_CUDA_TRY(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
            ws.ev.resize(cfg.host_slots_per_worker);
            ws.busy.assign(cfg.host_slots_per_worker, 0);
            for (auto& ev : ws.ev) LSIO_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        }
        LSIO_CUDA_TRY(cudaMalloc((void**)&g.dev_ring, ring_bytes));
    } else {
        // pinned (page-locked, DMA-able) and later registered with every worker's io_uring
        LSIO_CUDA_TRY(cudaHostAlloc((void**)&g.host_ring, ring_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
        LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&g.host_ring_dev, g.host_ring, 0));
        if (!cfg.zero_copy) LSIO_CUDA_TRY(cudaMalloc((void**)&g.dev_ring, ring_bytes));
    }
    g.log.init(ring_bytes, 1 << 16);
    uint8_t* arena_base = nullptr;
    LSIO_CUDA_TRY(cudaMalloc((void**)&arena_base, cfg.arena_bytes));
    g.arena = std::make_shared<DeviceArena>(g.device, arena_base, cfg.arena_bytes, [](int dev, void* p) {
        DeviceGuard d(dev);
        cudaFree(p);
    });
    LSIO_CUDA_TRY(cudaStreamCreateWithFlags(&g.s_copy, cudaStreamNonBlocking));
    LSIO_CUDA_TRY(cudaStreamCreateWithFlags(&g.s_pack, cudaStreamNonBlocking));
    const size_t max_segs = 8192, max_tiles = 1 << 16;
    g.batches.resize(8);
    for (Batch& b : g.batches) {
        b.lay = pack_batch_layout(max_segs, max_tiles);
        LSIO_CUDA_TRY(cudaHostAlloc((void**)&b.host_desc, b.lay.upload_bytes, cudaHostAllocPortable));
        LSIO_CUDA_TRY(cudaMalloc((void**)&b.dev_desc, b.lay.total_bytes));
        LSIO_CUDA_TRY(cudaMemset(b.dev_desc, 0, b.lay.total_bytes));  // the kernel's tile-scheduler words start at zero
        LSIO_CUDA_TRY(cudaHostAlloc((void**)&b.crc_host, max_segs * 4, cudaHostAllocPortable | cudaHostAllocMapped));
        LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&b.crc_dev, b.crc_host, 0));
        LSIO_CUDA_TRY(cudaEventCreateWithFlags(&b.copied, cudaEventDisableTiming));
        LSIO_CUDA_TRY(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
        b.desc.reserve(max_segs);
        g.free_batches.push_back(&b);
    }
    if (cfg.checksum & LSIO_CHECKSUM_XXH64) {
        g.xxh_batches.resize(4);
        for (XxhBatch& x : g.xxh_batches) {
            LSIO_CUDA_TRY(cudaHostAlloc((void**)&x.ptrs, 3 * XxhBatch::kMax * sizeof(uint64_t), cudaHostAllocPortable | cudaHostAllocMapped));
            LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&x.dev, x.ptrs, 0));
            LSIO_CUDA_TRY(cudaEventCreateWithFlags(&x.done, cudaEventDisableTiming));
            g.xxh_free.push_back(&x);
        }
    }
    return 0;
}

}  // namespace lsio

struct lsio_uring_cuda {
    std::unique_ptr<lsio::Backend> be;
};
struct lsio_completion {
    std::shared_ptr<lsio::OutChannel> ch;
};

using namespace lsio;

extern "C" {

void lsio_uring_cuda_config_default(lsio_uring_cuda_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->n_worker_threads = 4;
    cfg->n_devices = 1;
    cfg->devices[0] = 0;
    cfg->sq_entries = 64;          // worker.rs:17
    cfg->sqpoll = 0;               // the reference uses SQPOLL (one kernel thread per ring); off here: submissions are batched
    cfg->completion_bound = 1024;  // io_uring.rs:16
    cfg->checksum = LSIO_CHECKSUM_CRC32C;
    cfg->pack_align = 1;
    cfg->pin_threads = 1;
    cfg->o_direct = 1;             // sqe.rs:25
    cfg->staging_bytes = 1ull << 30;
    cfg->arena_bytes = 8ull << 30;
    // Reads (merged or split) are cut at 512 KiB and packed into 2 MiB extents (host_slot_kib).  Measured on this
    // pool: tmpfs-sourced C2 runs at 41.8 GiB/s for reads of 512 KiB..2 MiB, while the (virtio) disk gives 6.1 GiB/s
    // with 256-512 KiB reads and 3.8 with 2 MiB ones (more, smaller requests in flight).  In staging_mode 0 reads of
    // <= 1 MiB matter for another reason: the slot is handed to the H2D DMA while still in the 2 MiB per-core L2.
    cfg->max_read_bytes = 512ull << 10;
    cfg->min_copy_bytes = 8ull << 20;
    cfg->flush_us = 200;
    cfg->merge_ranges = 1;
    cfg->merge_gap_bytes = 0;
    cfg->statx_via_ring = 0;
    cfg->zero_copy = 0;
    cfg->staging_mode = 1;
    cfg->host_slots_per_worker = 4;
    cfg->host_slot_kib = 2048;
    cfg->poll_sleep_us = 50;  // measured: same throughput as spinning (33-34 GiB/s C2, 27.7 vs 28.1 C3) without burning a core per GPU
}

int lsio_uring_cuda_new_with_config(const lsio_uring_cuda_config* cfgp, lsio_uring_cuda** out) {
    if (!cfgp || !out) return -EINVAL;
    lsio_uring_cuda_config cfg = *cfgp;
    // small rings: keep a read (and a coalesced copy) to a quarter of the ring so the log keeps moving
    if (cfg.staging_bytes >= (4ull << 20)) {
        cfg.max_read_bytes = std::min<uint64_t>(cfg.max_read_bytes, cfg.staging_bytes / 4 / kAlign * kAlign);
        cfg.min_copy_bytes = std::min<uint64_t>(cfg.min_copy_bytes, cfg.staging_bytes / 4);
    }
    if (cfg.zero_copy) cfg.staging_mode = 0;  // the kernel needs the whole pinned ring mapped
    if (cfg.staging_mode == 1 && (cfg.host_slots_per_worker == 0 || cfg.host_slots_per_worker > 256 || cfg.n_worker_threads > 64)) {
        set_error("staging_mode 1 needs 1..256 host_slots_per_worker and at most 64 worker threads per device");
        return -EINVAL;
    }
    if (cfg.n_worker_threads == 0 || cfg.n_devices == 0 || cfg.n_devices > LSIO_MAX_DEVICES || cfg.sq_entries < 4 ||
        (cfg.sq_entries & (cfg.sq_entries - 1)) || cfg.staging_bytes < (4ull << 20) || cfg.arena_bytes < (1ull << 20) || cfg.max_read_bytes < kAlign ||
        cfg.max_read_bytes > (1ull << 30) || cfg.max_read_bytes > cfg.staging_bytes / 4) {
        set_error("invalid lsio_uring_cuda_config");
        return -EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s): this backend has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return e == cudaSuccess ? -ENODEV : (int)e;
    }
    auto be = std::make_unique<Backend>();
    be->cfg = cfg;
    be->out = std::make_shared<OutChannel>(cfg.completion_bound ? cfg.completion_bound : 1024);
    for (uint32_t i = 0; i < cfg.n_devices; i++) {
        if (cfg.devices[i] < 0 || cfg.devices[i] >= ndev) {
            set_error("device %d not present (%d devices)", cfg.devices[i], ndev);
            for (auto& g : be->groups) free_group_cuda(*g);
            return -ENODEV;
        }
        auto g = std::make_unique<Group>();
        g->be = be.get();
        g->index = (int)i;
        g->device = cfg.devices[i];
        int rc = init_group(*g);
        if (rc) {
            set_error("device %d set-up failed: %s", g->device, last_error_text().c_str());
            free_group_cuda(*g);
            for (auto& o : be->groups) free_group_cuda(*o);
            return rc;
        }
        be->groups.push_back(std::move(g));
    }
    // senders: every io worker and every device thread
    for (auto& gp : be->groups) {
        Group* g = gp.get();
        for (uint32_t w = 0; w < cfg.n_worker_threads + 1; w++) be->out->add_sender();
        g->workers_alive.store(cfg.n_worker_threads);
        g->device_thread = std::thread([g] {
            device_main(*g);
            g->be->out->remove_sender();
        });
        g->pool = std::make_unique<ThreadPool<Op>>(cfg.n_worker_threads, [g](WorkerThread<Op>& wt) {
            {
                Worker w(*g, wt);
                w.run();
            }
            g->workers_alive.fetch_sub(1, std::memory_order_acq_rel);
            g->be->out->remove_sender();
        }, g->cpus);
    }
    auto* h = new lsio_uring_cuda();
    h->be = std::move(be);
    *out = h;
    return 0;
}

int lsio_uring_cuda_new(size_t n_worker_threads, lsio_uring_cuda** out) {
    lsio_uring_cuda_config cfg;
    lsio_uring_cuda_config_default(&cfg);
    cfg.n_worker_threads = (uint32_t)n_worker_threads;
    return lsio_uring_cuda_new_with_config(&cfg, out);
}

static int submit(lsio_uring_cuda* self, int group, const char* location, const int64_t* rs, const int64_t* re, size_t n, const uint64_t* ud,
                  size_t nud) {
    if (!self || !location || ((!rs || !re || !ud) && n)) return -EINVAL;
    if (n != nud) return -EINVAL;           // assert_eq!(ranges.len(), user_data.len()), get_ranges.rs:28
    if (location[0] != '/') return -EINVAL;  // dirfd = -1: the path must be absolute (sqe.rs:22)
    Backend* be = self->be.get();
    if (be->closing.load()) return -ESHUTDOWN;
    auto r = std::make_shared<Request>();
    r->path = location;  // the reference takes ownership of its inputs; we copy (lsio_io lib.rs:38-40)
    r->starts.assign(rs, rs + n);
    r->ends.assign(re, re + n);
    r->user_data.assign(ud, ud + n);
    const size_t gi = group >= 0 ? (size_t)group : (size_t)(be->rr.fetch_add(1) % be->groups.size());
    if (gi >= be->groups.size()) return -EINVAL;
    r->grp = be->groups[gi].get();
    Op op;
    op.kind = Op::GetRanges;
    op.req = std::move(r);
    be->groups[gi]->pool->push(std::move(op));  // threadpool.rs:134-137
    return 0;
}

int lsio_uring_cuda_get_ranges(lsio_uring_cuda* self, const char* location, const int64_t* range_start, const int64_t* range_end, size_t n_ranges,
                               const uint64_t* user_data, size_t n_user_data) {
    return submit(self, -1, location, range_start, range_end, n_ranges, user_data, n_user_data);
}
int lsio_uring_cuda_get_ranges_on(lsio_uring_cuda* self, uint32_t group, const char* location, const int64_t* range_start, const int64_t* range_end,
                                  size_t n_ranges, const uint64_t* user_data, size_t n_user_data) {
    return submit(self, (int)group, location, range_start, range_end, n_ranges, user_data, n_user_data);
}

lsio_completion* lsio_uring_cuda_completion(lsio_uring_cuda* self) {
    if (!self) return nullptr;
    auto* c = new lsio_completion();
    c->ch = self->be->out;
    return c;
}

int lsio_uring_cuda_get_stats(const lsio_uring_cuda* self, lsio_uring_cuda_stats* o) {
    if (!self || !o) return -EINVAL;
    const Stats& s = self->be->stats;
    o->requests = s.requests; o->ranges = s.ranges; o->reads = s.reads; o->short_reads = s.short_reads; o->read_bytes = s.read_bytes;
    o->h2d_copies = s.h2d_copies; o->h2d_bytes = s.h2d_bytes; o->pack_launches = s.pack_launches; o->kernel_launches = s.kernel_launches;
    o->payload_bytes = s.payload_bytes; o->chunks_ok = s.chunks_ok; o->errors = s.errors; o->staging_stalls = s.staging_stalls;
    o->arena_stalls = s.arena_stalls;
    return 0;
}

const char* lsio_uring_cuda_last_error(void) { return g_last_error.c_str(); }

void lsio_uring_cuda_free(lsio_uring_cuda* self) {
    if (!self) return;
    Backend* be = self->be.get();
    be->closing.store(true);
    // A consumer that stopped reading must not wedge the teardown: drain undelivered items while joining.
    std::atomic<bool> joined{false};
    std::thread drainer([&] {
        lsio_output o;
        while (!joined.load()) {
            if (be->out->len() >= be->cfg.completion_bound && be->out->try_pop(&o)) {
                if (o.status == 0 && o.chunk.buffer) delete o.chunk.buffer;
            } else {
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
        }
    });
    for (auto& g : be->groups) g->pool->shutdown();  // Drop for ThreadPool, threadpool.rs:140-165
    for (auto& g : be->groups) {
        g->stop_device.store(true, std::memory_order_release);
        if (g->device_thread.joinable()) g->device_thread.join();
    }
    joined.store(true);
    drainer.join();
    for (auto& g : be->groups) free_group_cuda(*g);
    delete self;  // chunks already handed out keep the arena alive through their InnerBuffer
}

int lsio_probe_read_ceiling(const char* const* paths, size_t n_paths, uint64_t block_bytes, uint32_t n_threads, uint32_t queue_depth,
                            uint32_t o_direct, double* gib_per_s) {
    if (!paths || !n_paths || !gib_per_s || !n_threads || !queue_depth || block_bytes < kAlign || block_bytes % kAlign) return -EINVAL;
    const uint64_t slot_bytes = block_bytes * queue_depth;
    uint8_t* pinned = nullptr;
    LSIO_CUDA_TRY(cudaHostAlloc((void**)&pinned, slot_bytes * n_threads, cudaHostAllocPortable));
    std::atomic<size_t> next_file{0};
    std::atomic<uint64_t> total{0};
    std::atomic<int> err{0};
    const uint64_t t0 = now_ns();
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; t++) {
        th.emplace_back([&, t] {
            Ring ring;
            unsigned sq = 4;
            while (sq < queue_depth) sq <<= 1;
            int rc = ring.init(sq, 0);
            iovec iv{pinned + (uint64_t)t * slot_bytes, (size_t)slot_bytes};
            if (rc == 0) rc = ring.register_buffers(&iv, 1);
            if (rc) { err.store(rc); return; }
            uint64_t bytes = 0;
            for (size_t f; (f = next_file.fetch_add(1)) < n_paths;) {
                int fd = open(paths[f], O_RDONLY | (o_direct ? O_DIRECT : 0));
                if (fd < 0) { err.store(-errno); break; }
                struct stat st;
                fstat(fd, &st);
                const uint64_t nblk = ((uint64_t)st.st_size + block_bytes - 1) / block_bytes;
                uint64_t issued = 0, done = 0;
                std::vector<unsigned> free_slots;
                for (unsigned i = 0; i < queue_depth; i++) free_slots.push_back(i);
                while (done < nblk) {
                    while (issued < nblk && !free_slots.empty()) {
                        io_uring_sqe* e = ring.get_sqe();
                        if (!e) break;
                        const unsigned slot = free_slots.back();
                        free_slots.pop_back();
                        e->opcode = IORING_OP_READ_FIXED;
                        e->fd = fd;
                        e->addr = (uint64_t)(uintptr_t)(pinned + (uint64_t)t * slot_bytes + (uint64_t)slot * block_bytes);
                        e->len = (uint32_t)block_bytes;
                        e->off = issued * block_bytes;
                        e->buf_index = 0;
                        e->user_data = slot;
                        issued++;
                    }
                    ring.submit(1);
                    io_uring_cqe c;
                    while (ring.pop_cqe(&c)) {
                        if (c.res < 0) err.store(c.res);
                        else bytes += (uint64_t)c.res;
                        free_slots.push_back((unsigned)c.user_data);
                        done++;
                    }
                }
                close(fd);
            }
            total.fetch_add(bytes);
        });
    }
    for (auto& x : th) x.join();
    const double secs = (double)(now_ns() - t0) * 1e-9;
    cudaFreeHost(pinned);
    if (err.load()) return err.load();
    *gib_per_s = (double)total.load() / secs / (double)(1ull << 30);
    return 0;
}

lsio_completion* lsio_completion_clone(const lsio_completion* self) { return self ? new lsio_completion(*self) : nullptr; }
int lsio_completion_recv(lsio_completion* self, lsio_output* out, int timeout_ms) {
    if (!self || !out) return -EINVAL;
    return (int)self->ch->recv(out, timeout_ms);
}
int lsio_completion_try_recv(lsio_completion* self, lsio_output* out) {
    if (!self || !out) return -EINVAL;
    return (int)self->ch->recv(out, 0);
}
int lsio_completion_recv_batch(lsio_completion* self, lsio_output* out, size_t max, int timeout_ms, size_t* n_out) {
    if (!self || !out || !n_out) return -EINVAL;
    return (int)self->ch->recv_batch(out, max, timeout_ms, n_out);
}
size_t lsio_completion_len(const lsio_completion* self) { return self ? self->ch->len() : 0; }
void lsio_completion_drop(lsio_completion* self) {
    if (!self) return;
    // the last receiver releases chunks nobody will ever read
    if (self->ch.use_count() == 1) {
        lsio_output o;
        while (self->ch->try_pop(&o))
            if (o.status == 0 && o.chunk.buffer) delete o.chunk.buffer;
    }
    delete self;
}

}  // extern "C"
