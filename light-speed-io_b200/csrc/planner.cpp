// planner.cpp -- completion items and the request planner: the reference's range arithmetic
// (crates/lsio_uring/src/sqe.rs:59-122) and fan-out (get_ranges.rs:39-45) re-stated for a whole request,
// plus what the reference only plans: merging adjacent ranges and splitting long reads.
#include "backend.hpp"

namespace lsio {

// ---- completion items ----------------------------------------------------------------------
void send_output(Backend* be, lsio_output& o) {
    if (o.status != 0) be->stats.errors++;
    else be->stats.chunks_ok++;
    be->out->send(o);
}

// operation.rs:82-100 maybe_send_error: one Err per failed CQE, with context.
void send_error(Backend* be, int status, uint8_t opcode, const Request& r, int range, const char* what) {
    lsio_output o;
    memset(&o, 0, sizeof o);
    o.status = status;
    o.opcode = opcode;
    o.device = r.grp ? r.grp->device : -1;
    o.group = r.op_group;
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
// Range resolution and layout of the request's buffer (sqe.rs:64-81 + the prefix-sum layout): pure arithmetic, no device, no
// I/O.  Fills r.ranges and r.total_bytes; `invalid` receives the indices the reference would have panicked on.
void resolve_ranges(Request& r, int64_t filesize, uint64_t pack_align, std::vector<uint32_t>* invalid) {
    const size_t n = r.starts.size();
    r.ranges.reset(new RangePlan[n ? n : 1]);
    uint64_t dst = 0;
    const uint64_t pa = pack_align ? pack_align : 1;
    for (size_t i = 0; i < n; i++) {
        RangePlan& p = r.ranges[i];
        p.user_data = r.user_data[i];
        const int64_t s = r.starts[i], e = r.ends[i];
        const int64_t start_off = s >= 0 ? s : filesize + s;  // sqe.rs:64-70
        const int64_t end_off = e >= 0 ? e : filesize + e + 1;  // sqe.rs:73-80
        if (start_off < 0 || end_off < 0 || end_off <= start_off || end_off > filesize) {
            if (invalid) invalid->push_back((uint32_t)i);
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
}

// The read plan of a resolved request whose buffer starts at `block` (only its address matters: kernel tiles are counted in
// destination space): pure arithmetic.  Fills r.reads, r.segs and every range's seg_ids.
void build_read_plan(Request& r, const lsio_uring_cuda_config& cfg, const uint8_t* block) {
    const size_t n = r.starts.size();
    // Read plan: sort the ranges' padded spans, merge spans that touch / overlap / sit within
    // merge_gap_bytes of each other into runs, cut every run into reads of at most max_read_bytes
    // (reference: one read per range, panics above 2 GiB, get_range.rs:23-30).  Offsets and lengths are
    // multiples of the FILE's O_DIRECT alignment (reference: 512 hard-coded, sqe.rs:12, with the statx value
    // fetched and unused, open_file.rs:23-25).
    const uint64_t al = r.align;
    const uint64_t max_read = std::max<uint64_t>(al, cfg.max_read_bytes / al * al);
    std::vector<uint32_t> order;
    for (size_t i = 0; i < n; i++)
        if (r.ranges[i].valid) order.push_back((uint32_t)i);
    auto span_lo = [&](uint32_t i) { return (uint64_t)r.ranges[i].start_off / al * al; };              // sqe.rs:83
    auto span_hi = [&](uint32_t i) { return ((uint64_t)r.ranges[i].end_off + al - 1) / al * al; };    // padded length
    if (cfg.merge_ranges) std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return span_lo(a) < span_lo(b); });
    for (size_t k = 0; k < order.size();) {
        uint64_t run_lo = span_lo(order[k]), run_hi = span_hi(order[k]);
        size_t k_end = k + 1;
        if (cfg.merge_ranges)
            while (k_end < order.size() && span_lo(order[k_end]) <= run_hi + cfg.merge_gap_bytes) {
                run_hi = std::max(run_hi, span_hi(order[k_end]));
                k_end++;
            }
        size_t j_lo = k;  // ranges before j_lo end at or before the current read: skipped
        for (uint64_t fo = run_lo; fo < run_hi; fo += max_read) {
            ReadPlan rd;
            rd.file_off = fo;
            rd.read_len = (uint32_t)std::min<uint64_t>(max_read, run_hi - fo);
            rd.first_seg = (uint32_t)r.segs.size();
            auto close_read = [&] {
                rd.nsegs = (uint32_t)r.segs.size() - rd.first_seg;
                if (rd.nsegs) r.reads.push_back(rd);  // a read that carries no payload (pure gap) is skipped
                rd.first_seg = (uint32_t)r.segs.size();
                rd.nsegs = rd.ntiles = rd.needed = 0;
            };
            while (j_lo < k_end && (uint64_t)r.ranges[order[j_lo]].end_off <= fo && span_hi(order[j_lo]) <= fo) j_lo++;
            for (size_t j = j_lo; j < k_end; j++) {
                RangePlan& p = r.ranges[order[j]];
                if ((uint64_t)p.start_off >= fo + rd.read_len) break;  // sorted by span_lo: nothing further starts inside this read
                const uint64_t lo = std::max<uint64_t>(fo, (uint64_t)p.start_off), hi = std::min<uint64_t>(fo + rd.read_len, (uint64_t)p.end_off);
                if (lo >= hi) continue;
                SegPlan sp;
                sp.range = order[j];
                sp.head = (uint32_t)(lo - fo);  // sqe.rs:116: the view starts at start_off - aligned_start
                sp.payload_len = hi - lo;
                sp.dst_off = p.dst_off + (lo - (uint64_t)p.start_off);
                const uint32_t t = (uint32_t)pack_tile_count((uint64_t)(uintptr_t)(block + sp.dst_off), sp.payload_len);
                // A read must fit one pack batch with room to spare: many (duplicate / overlapping) ranges over the same
                // bytes are served by several reads of the same span.
                if (rd.first_seg != r.segs.size() && ((uint32_t)r.segs.size() - rd.first_seg >= kMaxSegsPerRead || rd.ntiles + t > kMaxTilesPerRead)) close_read();
                p.seg_ids.push_back((uint32_t)r.segs.size());
                rd.needed = std::max<uint32_t>(rd.needed, (uint32_t)(hi - fo));
                rd.ntiles += t;
                r.segs.push_back(sp);
            }
            close_read();
        }
        k = k_end;
    }
}

Start plan_request(Group& g, const std::shared_ptr<Request>& rp, WorkerThread<Op>& wt) {
    Request& r = *rp;
    Backend* be = g.be;
    const lsio_uring_cuda_config& cfg = be->cfg;
    const int64_t filesize = (int64_t)r.stx.stx_size;
    const size_t n = r.starts.size();
    if (!r.planned) {
        std::vector<uint32_t> invalid;
        resolve_ranges(r, filesize, cfg.pack_align, &invalid);
        // The reference asserts (panics) on these; here they are in-band errors (Appendix A).
        for (uint32_t i : invalid) send_error(be, -EINVAL, 0, r, (int)i, "invalid byte range for this file");
        r.planned = true;
        be->stats.ranges += n;
    }
    if (r.total_bytes == 0) {  // nothing to read
        close(r.fd);
        r.fd = -1;
        return Start::Finished;
    }
    const uint64_t epoch_before = g.arena->epoch();
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
        r.arena_epoch_seen = epoch_before;
        return Start::Retry;  // wait for the consumer to drop chunks
    }
    auto buf = std::make_shared<InnerBuffer>();
    buf->buf = block;
    buf->size = r.total_bytes;
    buf->align = r.align;
    buf->kind = MemKind::Device;
    buf->device = g.device;
    buf->release = [arena = g.arena, block](InnerBuffer&) { arena->release(block); };
    r.buffer = std::move(buf);

    build_read_plan(r, cfg, block);
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

}  // namespace lsio

// Test aid (include/lsio_uring_cuda.h): the planner's arithmetic on its own, no device, no I/O.
extern "C" int lsio_uring_cuda_plan_reads(int64_t filesize, const int64_t* starts, const int64_t* ends, size_t n, uint32_t align, uint64_t max_read_bytes,
                                          uint32_t merge_ranges, uint64_t merge_gap_bytes, uint32_t pack_align, uint64_t buffer_address, int64_t* ranges_out,
                                          uint64_t* reads_out, size_t cap_reads, size_t* n_reads, uint64_t* segs_out, size_t cap_segs, size_t* n_segs,
                                          uint64_t* total_bytes) {
    using namespace lsio;
    if ((!starts || !ends || !ranges_out) && n) return -EINVAL;
    if (!n_reads || !n_segs || !total_bytes || filesize < 0 || align < kAlign || (align & (align - 1)) || max_read_bytes < align) return -EINVAL;
    Request r;
    r.starts.assign(starts, starts + n);
    r.ends.assign(ends, ends + n);
    r.user_data.assign(n, 0);
    r.align = align;
    lsio_uring_cuda_config cfg;
    lsio_uring_cuda_config_default(&cfg);
    cfg.max_read_bytes = max_read_bytes;
    cfg.merge_ranges = merge_ranges;
    cfg.merge_gap_bytes = merge_gap_bytes;
    cfg.pack_align = pack_align;
    resolve_ranges(r, filesize, pack_align, nullptr);
    if (r.total_bytes) build_read_plan(r, cfg, reinterpret_cast<const uint8_t*>((uintptr_t)buffer_address));
    for (size_t i = 0; i < n; i++) {
        const RangePlan& p = r.ranges[i];
        ranges_out[4 * i] = p.valid;
        ranges_out[4 * i + 1] = p.start_off;
        ranges_out[4 * i + 2] = p.end_off;
        ranges_out[4 * i + 3] = (int64_t)p.dst_off;
    }
    for (size_t k = 0; k < r.reads.size() && k < cap_reads && reads_out; k++) {
        const ReadPlan& rd = r.reads[k];
        uint64_t* o = reads_out + 6 * k;
        o[0] = rd.file_off; o[1] = rd.read_len; o[2] = rd.needed; o[3] = rd.first_seg; o[4] = rd.nsegs; o[5] = rd.ntiles;
    }
    for (size_t k = 0; k < r.segs.size() && k < cap_segs && segs_out; k++) {
        const SegPlan& sp = r.segs[k];
        uint64_t* o = segs_out + 4 * k;
        o[0] = sp.range; o[1] = sp.head; o[2] = sp.payload_len; o[3] = sp.dst_off;
    }
    *n_reads = r.reads.size();
    *n_segs = r.segs.size();
    *total_bytes = r.total_bytes;
    return 0;
}
