// worker.cpp -- one io worker: private io_uring, tracker of in-flight ops, the GetRanges / Read / Close state
// machines and the staging of landed reads (crates/lsio_uring/src/worker.rs:24-135, get_ranges.rs, get_range.rs,
// close.rs, sqe.rs builders).
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

void worker_main(Group& g, WorkerThread<Op>& wt) {
    Worker w(g, wt);
    w.run();
}

}  // namespace lsio
