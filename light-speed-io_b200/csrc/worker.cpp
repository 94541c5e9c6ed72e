// worker.cpp -- one io worker: private io_uring, tracker of in-flight ops, the GetRanges / Read / Close state
// machines and the staging of landed reads (crates/lsio_uring/src/worker.rs:24-135, get_ranges.rs, get_range.rs,
// close.rs, sqe.rs builders, open_file.rs).
#include <sys/prctl.h>
#include <sys/uio.h>

#include <unordered_map>

#include "backend.hpp"

namespace lsio {

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
    // Can the kernel read files of this filesystem asynchronously (FMODE_NOWAIT)?  One probe per st_dev.
    std::unordered_map<uint64_t, bool> async_capable;
    // thread-local counters, folded into Backend::stats when the worker parks / exits
    uint64_t c_read_ns = 0, c_copy_ns = 0, c_stall_ns = 0, c_busy_ns = 0, c_inline = 0;
    static constexpr size_t kNoIdx = ~size_t(0);

    Worker(Group& grp, WorkerThread<Op>& w) : g(grp), wt(w), tracker(grp.be->cfg.sq_entries) {
        if (grp.be->cfg.staging_mode == 1) {
            slots = &grp.wslots[w.id()];
            extents.resize(slots->busy.size());
            slot_ev.assign(slots->busy.size(), 0);
        }
    }

    void fold_counters() {
        Stats& s = g.be->stats;
        s.read_ns += c_read_ns;
        s.copy_enqueue_ns += c_copy_ns;
        s.worker_stall_ns += c_stall_ns;
        s.worker_busy_ns += c_busy_ns;
        s.inline_reads += c_inline;
        c_read_ns = c_copy_ns = c_stall_ns = c_busy_ns = c_inline = 0;
    }

    // staging_mode 1: per-slot extent bookkeeping (this worker only)
    struct Extent {
        StageEntry* entry = nullptr;
        uint32_t pending = 0;  // reads submitted but not yet landed/failed
        bool closed = false;   // no more reads will be added
        uint32_t nsegs = 0, ntiles = 0;  // descriptors / kernel tiles of the reads placed so far (bounded: one extent fits one pack batch)
    };
    std::vector<Extent> extents;  // one per slot
    int cur = -1;                 // slot of the extent being filled, or -1

    // Direct single-run extents that fill consecutive slots and consecutive bytes of one request buffer are copied with
    // ONE cudaMemcpyAsync (copy_coalesce_slots): the copy engine loses ~3.8 us per copy whatever its size (measured,
    // tools/h2d_probe.cu: 2 MiB copies 46.9 GiB/s, 4 MiB 49.2, 8 MiB 50.4, 64 MiB 51.6), the slots stay cache-sized.
    struct PendingCopy {
        int first_slot = -1, last_slot = -1, n = 0;
        uint32_t src = 0;  // offset of the run in the first slot
        uint8_t* dst = nullptr;
        uint64_t len = 0;
        StageEntry* ents[8] = {};
        int slot_of[8] = {};
    } pend;
    std::vector<int> slot_ev;  // event that frees each slot (a coalesced copy records one event for all of its slots)
    int next_slot = 0;

    void fail_extent(StageEntry* ent, cudaError_t ce) {
        for (auto& r : ent->reads) {
            if (r.state != kStageLanded) continue;
            r.state = kStageFailed;
            const ReadPlan& rd = r.req->reads[r.read];
            for (uint32_t k = 0; k < rd.nsegs; k++) {
                const uint32_t range = r.req->segs[rd.first_seg + k].range;
                if (!r.req->ranges[range].failed.exchange(true)) send_error(g.be, (int)ce, 0, *r.req, (int)range, "cudaMemcpyAsync (H2D) failed");
            }
        }
        ent->nsegs = ent->ntiles = 0;
    }
    void publish(StageEntry* ent) {
        ent->landed_ns = now_ns();
        ent->state.store(kStageLanded, std::memory_order_release);
    }

    // Issue the deferred copy (if any) and publish its extents.  Called whenever the worker is about to wait for
    // anything (a free slot, a completion, new work) -- a deferred slot is never held across a wait.
    void flush_pending() {
        if (pend.n == 0) return;
        Backend* be = g.be;
        const uint64_t t0 = now_ns();
        cudaError_t ce = cudaMemcpyAsync(pend.dst, slots->base + (uint64_t)pend.first_slot * g.slot_bytes + pend.src, pend.len, cudaMemcpyHostToDevice, slots->stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(slots->ev[pend.last_slot], slots->stream);
        c_copy_ns += now_ns() - t0;
        if (ce != cudaSuccess) cudaGetLastError();
        for (int i = 0; i < pend.n; i++) {
            StageEntry* ent = pend.ents[i];
            const int slot = pend.slot_of[i];
            if (ce != cudaSuccess) {
                fail_extent(ent, ce);
                slots->busy[slot] = 0;
            } else {
                slots->busy[slot] = 2;
                slot_ev[slot] = pend.last_slot;
                ent->copy_event = slots->ev[pend.last_slot];
                ent->owner = (uint32_t)wt.id();
                ent->direct = true;
                be->stats.direct_extents++;
            }
            publish(ent);
        }
        if (ce == cudaSuccess) {
            be->stats.direct_copies++;
            be->stats.h2d_copies++;
            be->stats.h2d_bytes += pend.len;
        }
        pend = PendingCopy{};
    }

    // direct_landing 2: the extent is moved by the pull kernel, launched here on this worker's stream: its warps read the
    // pinned slot over PCIe through the mapped address, checksum the bytes and write them at their final place in the
    // request buffers -- one pass, no copy engine, no device staging ring (pack_crc.cu pull_kernel).  Needs at most kPullMaxSegs
    // segments (the descriptors travel as kernel parameters) and a payload the slot's scratch can hold (overlapping ranges share
    // a read's bytes, so payload > slot is possible); without pull_any_shift also every segment's slot address congruent to its
    // destination mod 16.  Anything else takes the other paths.  (With a shift every destination vector needs two source loads;
    // the second hits the L2 line the first brought in, so the bytes still cross PCIe once: measured)
    bool try_pull(int slot, StageEntry* ent, uint32_t nsegs) {
        Backend* be = g.be;
        lsio_cuda_segment segs[kPullMaxSegs];
        uint32_t n = 0;
        uint64_t payload = 0, tiles = 0;
        const uint32_t max_tiles = pack_pull_max_tiles(g.slot_bytes);
        uint8_t* sdev = slots->base_dev + (uint64_t)slot * g.slot_bytes;
        for (auto& r : ent->reads) {
            if (r.state != kStageLanded) continue;
            const ReadPlan& rd = r.req->reads[r.read];
            for (uint32_t k = 0; k < rd.nsegs; k++) {
                const SegPlan& sp = r.req->segs[rd.first_seg + k];
                lsio_cuda_segment& sg = segs[n++];
                sg.src = (uint64_t)(uintptr_t)(sdev + r.off + sp.head);
                sg.dst = (uint64_t)(uintptr_t)(r.req->buffer->buf + sp.dst_off);
                sg.len = sp.payload_len;
                if (((sg.src - sg.dst) & 15) && !be->cfg.pull_any_shift) return false;
                payload += sp.payload_len;
                // ranges that overlap in the file share the slot's bytes: the extent's payload can be many times the slot;
                // beyond what the slot's scratch holds it goes ring -> pack kernel (whose batches are sized for it)
                tiles += pack_tile_count(sg.dst, sg.len);
                if (tiles > max_tiles) return false;
            }
        }
        flush_pending();
        const bool want_crc = be->cfg.checksum & LSIO_CHECKSUM_CRC32C;
        const uint64_t t0 = now_ns();
        int rc = pack_pull_launch(g.device, segs, n, want_crc, slots->scratch[slot], max_tiles, g.pull_crc_dev + (ent->seq % g.n_pull) * kPullMaxSegs, slots->stream);
        cudaError_t ce = rc ? cudaErrorLaunchFailure : cudaEventRecord(slots->ev[slot], slots->stream);
        c_copy_ns += now_ns() - t0;
        if (ce != cudaSuccess) {
            cudaGetLastError();
            fail_extent(ent, ce);
            slots->busy[slot] = 0;
        } else {
            slots->busy[slot] = 2;
            slot_ev[slot] = slot;
            ent->copy_event = slots->ev[slot];
            ent->owner = (uint32_t)wt.id();
            ent->direct = ent->pulled = true;
            be->stats.direct_extents++;
            be->stats.pull_launches++;
            be->stats.kernel_launches++;
            be->stats.h2d_bytes += payload;
        }
        publish(ent);
        return true;
    }

    // The extent's reads have all landed (or failed) and nothing more will be added: copy it to the device
    // while it is still in this core's cache, then publish it to the device thread.  Two ways:
    //   direct  -- its payload forms a few runs contiguous both in the slot and in the request buffers: one
    //              cudaMemcpyAsync per run straight into the arena; the kernel later only checksums in place
    //   ring    -- ONE cudaMemcpyAsync of the whole extent into its reservation in the device log; the pack
    //              kernel strips the padding and scatters the ranges (small / unaligned / interleaved ranges)
    void finish_extent(int slot) {
        Extent& x = extents[slot];
        StageEntry* ent = x.entry;
        Backend* be = g.be;
        const lsio_uring_cuda_config& cfg = be->cfg;
        struct Run {
            uint32_t src;
            uint8_t* dst;
            uint64_t len;
        };
        constexpr int kMaxRuns = 4;
        Run runs[kMaxRuns];
        int nruns = 0;
        bool direct = cfg.direct_landing != 0;
        uint64_t payload = 0;
        uint32_t nsegs = 0, ntiles = 0;
        for (auto& r : ent->reads) {
            if (r.state != kStageLanded) continue;
            const ReadPlan& rd = r.req->reads[r.read];
            nsegs += rd.nsegs;
            ntiles += rd.ntiles;
            for (uint32_t k = 0; direct && k < rd.nsegs; k++) {
                const SegPlan& sp = r.req->segs[rd.first_seg + k];
                const uint32_t src = r.off + sp.head;
                uint8_t* dst = r.req->buffer->buf + sp.dst_off;
                payload += sp.payload_len;
                if (nruns && runs[nruns - 1].src + runs[nruns - 1].len == src && runs[nruns - 1].dst + runs[nruns - 1].len == dst) runs[nruns - 1].len += sp.payload_len;
                else if (nruns == kMaxRuns) direct = false;
                else runs[nruns++] = Run{src, dst, sp.payload_len};
            }
        }
        // several small copies cost more than one large copy plus the kernel's scatter
        direct = direct && nruns > 0 && (nruns == 1 || payload >= (uint64_t)nruns * (128u << 10));
        ent->nsegs = nsegs;
        ent->ntiles = ntiles;
        x.entry = nullptr;
        if (nsegs && cfg.direct_landing == 2 && nsegs <= kPullMaxSegs && try_pull(slot, ent, nsegs)) return;
        const uint32_t coalesce = std::min<uint32_t>(cfg.copy_coalesce_slots, 8);
        if (nsegs && direct && nruns == 1 && coalesce > 1) {
            const bool to_slot_end = runs[0].src + runs[0].len == g.slot_bytes;
            if (pend.n && slot == pend.last_slot + 1 && runs[0].src == 0 && runs[0].dst == pend.dst + pend.len) {
                pend.ents[pend.n] = ent;
                pend.slot_of[pend.n++] = slot;
                pend.last_slot = slot;
                pend.len += runs[0].len;
                slots->busy[slot] = 3;
                if (!to_slot_end || pend.n == (int)coalesce || slot + 1 == (int)slots->busy.size()) flush_pending();
                return;
            }
            flush_pending();
            if (to_slot_end && slot + 1 < (int)slots->busy.size()) {  // may be continued by the next slot
                pend.first_slot = pend.last_slot = slot;
                pend.n = 1;
                pend.src = runs[0].src;
                pend.dst = runs[0].dst;
                pend.len = runs[0].len;
                pend.ents[0] = ent;
                pend.slot_of[0] = slot;
                slots->busy[slot] = 3;
                return;
            }
        } else {
            flush_pending();
        }
        if (nsegs) {
            const uint64_t t0 = now_ns();
            uint8_t* sbase = slots->base + (uint64_t)slot * g.slot_bytes;
            cudaError_t ce = cudaSuccess;
            if (direct) {
                for (int i = 0; i < nruns && ce == cudaSuccess; i++) ce = cudaMemcpyAsync(runs[i].dst, sbase + runs[i].src, runs[i].len, cudaMemcpyHostToDevice, slots->stream);
            } else {
                const uint64_t phys = ent->log_off % g.log.capacity();
                ce = cudaMemcpyAsync(g.dev_ring + phys, sbase, ent->used, cudaMemcpyHostToDevice, slots->stream);
            }
            if (ce == cudaSuccess) ce = cudaEventRecord(slots->ev[slot], slots->stream);
            c_copy_ns += now_ns() - t0;
            if (ce != cudaSuccess) {
                cudaGetLastError();
                fail_extent(ent, ce);
                slots->busy[slot] = 0;
            } else {
                slots->busy[slot] = 2;
                slot_ev[slot] = slot;
                ent->copy_event = slots->ev[slot];
                ent->owner = (uint32_t)wt.id();
                ent->direct = direct;
                if (direct) {
                    be->stats.direct_extents++;
                    be->stats.direct_copies += (uint64_t)nruns;
                    be->stats.h2d_copies += (uint64_t)nruns;
                    be->stats.h2d_bytes += payload;
                } else {
                    be->stats.h2d_copies++;
                    be->stats.h2d_bytes += ent->used;
                }
            }
        } else {
            slots->busy[slot] = 0;
        }
        publish(ent);
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
        const int n = (int)slots->busy.size();
        for (int i = 0; i < n; i++)
            if (slots->busy[i] == 2 && cudaEventQuery(slots->ev[slot_ev[i]]) != cudaErrorNotReady) slots->busy[i] = 0;
        // consecutive extents go to consecutive slots, so that their copies can be coalesced
        for (int k = 0; k < n; k++) {
            const int i = (next_slot + k) % n;
            if (slots->busy[i] == 0) {
                next_slot = (i + 1) % n;
                return i;
            }
        }
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

    // Host address this read lands at (+ the bytes already landed).
    uint8_t* read_dst(const Op& op) const {
        if (slots) return slots->base + (uint64_t)op.slot * g.slot_bytes + op.entry->reads[op.ref].off + op.done;
        return g.host_ring + op.entry->log_off % g.log.capacity() + op.done;
    }

    // Test knob (debug_short_read_every): every k-th read is issued shorter than it should be, so that the
    // continuation path below runs (the reference leaves short reads as a TODO: get_range.rs:66-68).
    uint32_t maybe_truncate(const Request& r, uint32_t len) {
        const uint32_t k = g.be->cfg.debug_short_read_every;
        if (!k || len <= r.align) return len;
        if (g.be->stats.read_seq.fetch_add(1, std::memory_order_relaxed) % k != k - 1) return len;
        return std::max<uint32_t>(r.align, len / 2 / r.align * r.align);
    }

    void push_read_sqe(size_t idx, Op& op) {
        const ReadPlan& sp = op.req->reads[op.read];
        io_uring_sqe* e = must_get_sqe();
        e->opcode = IORING_OP_READ_FIXED;  // north star: READ_FIXED into registered pinned memory (reference: Read, sqe.rs:95)
        e->fd = op.req->fd;
        e->addr = (uint64_t)(uintptr_t)read_dst(op);
        e->buf_index = slots ? 0 : (uint16_t)((op.entry->log_off % g.log.capacity() + op.done) / StagingLog::kRegBoundary);
        e->len = maybe_truncate(*op.req, sp.read_len - op.done);
        e->off = sp.file_off + op.done;
        e->user_data = UringUserData::pack((uint32_t)idx, IORING_OP_READ_FIXED);
    }

    // Per-file set-up once fd and statx are known (OpenFileBuilder::build, open_file.rs:69-78): the O_DIRECT
    // alignment the reference fetches and never uses, and how this file is read.
    bool configure_request(Request& r) {
        const lsio_uring_cuda_config& cfg = g.be->cfg;
        uint32_t al = (uint32_t)kAlign;
        if (cfg.force_align) al = cfg.force_align;
        else if (cfg.o_direct && (r.stx.stx_mask & STATX_DIOALIGN)) al = std::max<uint32_t>(al, std::max(r.stx.stx_dio_offset_align, r.stx.stx_dio_mem_align));
        const uint64_t room = slots ? g.slot_bytes : std::min<uint64_t>(g.log.capacity() / 4, 1u << 16);
        if ((al & (al - 1)) || al > room || al > 65536) {
            send_error(g.be, -EINVAL, IORING_OP_STATX, r, -1, "unsupported O_DIRECT alignment reported for this file");
            return false;
        }
        r.align = al;
        uint64_t seen = g.be->stats.max_dio_align.load(std::memory_order_relaxed);
        while (al > seen && !g.be->stats.max_dio_align.compare_exchange_weak(seen, al)) {}
        if (cfg.inline_reads == 2) r.inline_io = true;
        else if (cfg.inline_reads == 1) {
            const uint64_t dev = ((uint64_t)r.stx.stx_dev_major << 32) | r.stx.stx_dev_minor;
            auto it = async_capable.find(dev);
            if (it == async_capable.end()) {
                // RWF_NOWAIT is refused with EOPNOTSUPP exactly when the file lacks FMODE_NOWAIT -- the same test io_uring
                // applies before it hands a read to an io-wq thread instead of issuing it from the submitter.
                uint8_t byte;
                iovec iv{&byte, 1};
                const ssize_t rc = preadv2(r.fd, &iv, 1, 0, RWF_NOWAIT);
                it = async_capable.emplace(dev, !(rc < 0 && errno == EOPNOTSUPP)).first;
            }
            r.inline_io = !it->second;
        }
        return true;
    }

    // One Err per range that this read was carrying (reference: one Err per failed CQE = per range), then the
    // read counts as done.
    void fail_read(size_t idx, Op& op, int status, const char* what) {
        Request& r = *op.req;
        const ReadPlan& rd = r.reads[op.read];
        uint32_t last = UINT32_MAX;
        for (uint32_t k = 0; k < rd.nsegs; k++) {
            const uint32_t range = r.segs[rd.first_seg + k].range;
            if (range == last) continue;
            last = range;
            if (!r.ranges[range].failed.exchange(true, std::memory_order_relaxed)) send_error(g.be, status, IORING_OP_READ_FIXED, r, (int)range, what);
        }
        read_done(op, kStageFailed);
        finish_read_accounting(idx, op);
    }

    // The kernel cannot issue this file's reads asynchronously (or null_source): the worker performs the read
    // itself, straight into the slot, handling short reads and errors exactly like the CQE path.
    void read_inline(Op& op) {
        Backend* be = g.be;
        Request& r = *op.req;
        const ReadPlan& rd = r.reads[op.read];
        const uint64_t t0 = now_ns();
        int err = 0;
        const char* what = nullptr;
        while (op.done < rd.needed) {
            const uint32_t want = maybe_truncate(r, rd.read_len - op.done);
            ssize_t n;
            if (be->cfg.null_source) {
                n = want;
            } else {
                iovec iv{read_dst(op), want};
                n = preadv2(r.fd, &iv, 1, (off_t)(rd.file_off + op.done), 0);
                if (n < 0 && errno == EINTR) continue;
            }
            if (n < 0) {
                err = -errno;
                what = "read failed";
                break;
            }
            op.done += (uint32_t)n;
            be->stats.read_bytes += (uint64_t)n;
            if (op.done < rd.needed) {
                if (n == 0 || (op.done % r.align) != 0) {
                    err = -EIO;
                    what = "short read (unexpected end of file)";
                    break;
                }
                be->stats.short_reads++;
            }
        }
        c_read_ns += now_ns() - t0;
        c_inline++;
        if (err) {
            fail_read(kNoIdx, op, err, what);
            return;
        }
        read_done(op, kStageLanded);
        finish_read_accounting(kNoIdx, op);
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
            Request& r = *op.req;
            const ReadPlan& sp = r.reads[op.read];
            StageEntry* ent;
            uint32_t off = 0;
            int slot = -1;
            if (slots) {
                // pack this read behind the previous ones of the current extent (at the file's alignment), or open a new extent;
                // an extent never outgrows a pack batch (segments / kernel tiles)
                if (cur >= 0) {
                    const Extent& x = extents[cur];
                    off = (x.entry->used + r.align - 1) / r.align * r.align;
                    if (off + sp.read_len > g.slot_bytes || x.nsegs + sp.nsegs > kMaxSegsPerExtent || x.ntiles + sp.ntiles > kMaxTilesPerExtent) close_current_extent();
                }
                if (cur < 0) {
                    const int free_slot = acquire_slot();
                    if (free_slot < 0) {  // all of this worker's slots are still being read or copied
                        flush_pending();  // a deferred copy holds slots: it must be under way before this worker waits for one
                        be->stats.staging_stalls++;
                        return Start::Retry;
                    }
                    StageEntry* fresh = g.log.alloc((uint32_t)g.slot_bytes);
                    if (!fresh) {
                        flush_pending();
                        next_slot = free_slot;
                        be->stats.staging_stalls++;
                        return Start::Retry;
                    }
                    slots->busy[free_slot] = 1;
                    extents[free_slot] = Extent{fresh, 0, false, 0, 0};
                    cur = free_slot;
                    off = 0;
                }
                slot = cur;
                Extent& x = extents[cur];
                ent = x.entry;
                ent->used = off + sp.read_len;
                x.pending++;
                x.nsegs += sp.nsegs;
                x.ntiles += sp.ntiles;
            } else {
                ent = g.log.alloc(sp.read_len, r.align);
                if (!ent) {
                    be->stats.staging_stalls++;
                    return Start::Retry;
                }
                ent->nsegs = sp.nsegs;
                ent->ntiles = sp.ntiles;
            }
            op.ref = (uint32_t)ent->reads.size();
            ent->reads.push_back(ReadRefT<Request>{op.req, op.read, off, kStageReading});
            op.entry = ent;
            op.done = 0;
            op.slot = slot;
            be->stats.reads++;
            if (r.inline_io || be->cfg.null_source) {
                read_inline(op);
                return Start::Finished;
            }
            auto idx = tracker.get_next_index();
            push_read_sqe(*idx, op);
            tracker.put(*idx, std::move(op));
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
            if (state != kStageLanded) ent->nsegs = ent->ntiles = 0;
            ent->landed_ns = now_ns();
            ent->state.store(kStageLanded, std::memory_order_release);  // staging_mode 0: one read per extent
        }
    }

    void finish_read_accounting(size_t idx, Op& op) {
        // get_range.rs:78-90: the last reader of the file closes it, re-using this tracker slot
        Request& r = *op.req;
        if (r.reads_left.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            if (idx == kNoIdx) {  // inline read: no tracker slot to re-use, no reason to go through the ring
                close(r.fd);
                r.fd = -1;
                return;
            }
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
        if (idx != kNoIdx) tracker.remove(idx);  // NextStep::Done
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
            if (!be->cfg.statx_via_ring) {
                // fast metadata path: statx on the fd we just opened (size + STATX_DIOALIGN, as open_file.rs:23-25
                // wants them); as cheap as an fstat and never blocks
                if (r.have_fd) {
                    if (statx(r.fd, "", AT_EMPTY_PATH, STATX_SIZE | STATX_DIOALIGN, &r.stx) == 0) r.have_statx = true;
                    else send_error(be, -errno, IORING_OP_STATX, r, -1, "statx failed");
                }
            } else if (r.n_cqes_received < 2) {
                return;  // NextStep::Pending
            }
            Op done = tracker.remove(ud.index_of_op);
            if (r.have_fd && r.have_statx && configure_request(r)) {
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
            if (cqe.res < 0) {
                fail_read(ud.index_of_op, *op, cqe.res, "read failed");
                return;
            }
            op->done += (uint32_t)cqe.res;
            be->stats.read_bytes += (uint64_t)cqe.res;
            if (op->done < rd.needed) {  // the reference does not check (get_range.rs:66-68, issue #100)
                if (cqe.res == 0 || (op->done % r.align) != 0) {
                    fail_read(ud.index_of_op, *op, -EIO, "short read (unexpected end of file)");
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
        prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);  // short sleeps stay short (default slack: 50 us on top of every sleep)
        uint64_t t_awake = now_ns();
        while (true) {
            const bool running = wt.keep_running();
            if (!running) {
                // shutting down: the partly filled extent must still be published (the device thread waits for every
                // extent that was reserved in the log) -- with inline reads it can be open while the tracker is empty
                if (slots) {
                    close_current_extent();
                    flush_pending();
                }
                if (tracker.is_empty()) break;
            }
            unsigned started = 0, inline_done = 0;
            if (running) {
                auto room = [&] { return !ring_ok || !(tracker.is_full() || ring.sq_len() + 2 > sq || ring.sq_len() + ring.cq_len() + 2 > sq); };
                // Reads done inline block this thread: after an extent's worth the loop goes on to submit and reap what
                // is queued in the ring (the openat of the next file, closes) before taking more.
                auto inline_budget = [&] { return inline_done < 8; };
                // 1. reads that waited for staging space, in order
                while (!stalled_reads.empty() && room() && inline_budget()) {
                    const uint64_t before = c_inline;
                    const Start s = start(stalled_reads.front());
                    if (s == Start::Retry) break;
                    stalled_reads.pop_front();
                    if (s == Start::Started) started++;
                    inline_done += (unsigned)(c_inline - before);
                }
                // 2. plans that waited for HBM, each once
                for (size_t n = stalled_plans.size(); n-- > 0;) {
                    Op op = std::move(stalled_plans.front());
                    stalled_plans.pop_front();
                    // nothing has been released since it last failed: do not even take the arena lock
                    if (op.req->planned && op.req->arena_epoch_seen == g.arena->epoch()) {
                        stalled_plans.push_back(std::move(op));
                        continue;
                    }
                    if (start(op) == Start::Retry) stalled_plans.push_back(std::move(op));
                    else started++;
                }
                // 3. top up the SQ with new tasks (worker.rs:65-88), two SQEs of headroom (MAX_SQ_ENTRIES_PER_ITERATION)
                bool ran_dry = false;
                while (stalled_reads.empty() && room() && inline_budget()) {
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
                    const uint64_t before = c_inline;
                    const Start s = start(op);
                    if (s == Start::Retry) (op.kind == Op::Read ? stalled_reads : stalled_plans).push_back(std::move(op));
                    else if (s == Start::Started) started++;
                    inline_done += (unsigned)(c_inline - before);
                }
                // nothing more to add right now: let the partly filled extent go (bounds latency, ends a workload)
                if (slots && ran_dry && stalled_reads.empty()) close_current_extent();
            }
            if (ring_ok && ring.unsubmitted()) ring.submit();  // one enter per batch (reference: per op, worker.rs:79)
            bool got = false;
            io_uring_cqe cqe;
            while (ring_ok && ring.pop_cqe(&cqe)) {  // worker.rs:104-122
                got = true;
                handle(cqe);
            }
            if (ring_ok && ring.unsubmitted()) ring.submit();
            if (got || started || inline_done) continue;
            if (slots) flush_pending();  // about to wait (for work, a completion, a slot, HBM): nothing stays deferred
            if (tracker.is_empty() && stalled_reads.empty() && stalled_plans.empty()) {
                if (running) {
                    c_busy_ns += now_ns() - t_awake;
                    fold_counters();
                    wt.park();  // worker.rs:93-99
                    t_awake = now_ns();
                }
                continue;
            }
            if (!tracker.is_empty()) {
                ring.submit(1, 200);  // submit_and_wait(1), bounded so new tasks are noticed
            } else {  // only waiting for staging / HBM space
                const uint64_t t0 = now_ns();
                std::this_thread::sleep_for(std::chrono::microseconds(20));  // precise: the thread's timer slack is 1 us (run())
                c_stall_ns += now_ns() - t0;
            }
        }
        c_busy_ns += now_ns() - t_awake;
        fold_counters();
    }
};

void worker_main(Group& g, WorkerThread<Op>& wt) {
    Worker w(g, wt);
    w.run();
}

}  // namespace lsio
