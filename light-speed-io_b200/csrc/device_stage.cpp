// device_stage.cpp -- the per-GPU device thread: batches published extents in log order, launches
// pack_strip_crc32c (+ the optional XXH64 stage), retires batches and emits one Chunk per range.
// No counterpart in the reference: this is the stage inserted after the read CQE.
#include <sys/prctl.h>

#include "backend.hpp"

namespace lsio {

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
    o.group = r.op_group;
    be->stats.payload_bytes += p.len;
    {   // chunk latency: request submitted -> chunk handed to the channel (log2 buckets of microseconds)
        const uint64_t now = now_ns();
        const uint64_t us = (now - r.submit_ns) / 1000;
        be->stats.latency[us ? std::min(31, 63 - __builtin_clzll(us)) : 0].fetch_add(1, std::memory_order_relaxed);
        if (!be->stats.first_chunk_ns.load(std::memory_order_relaxed)) {
            uint64_t zero = 0;
            const uint64_t first = be->stats.first_submit_ns.load(std::memory_order_relaxed);
            be->stats.first_chunk_ns.compare_exchange_strong(zero, now > first ? now - first : 1);
        }
    }
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
        if (rc == 0) {
            g.be->stats.kernel_launches++;
            const cudaError_t ce = cudaEventRecord(xb->done, g.s_pack);
            if (ce != cudaSuccess) rc = cuda_fail(ce, "cudaEventRecord (xxh64 stage)", __FILE__, __LINE__);
        }
        if (rc) {
            // errors instead of a hang: every range of this batch gets its Err item, the batch goes back to the pool
            cudaStreamSynchronize(g.s_pack);
            cudaGetLastError();
            for (auto& it : xb->items) send_error(g.be, rc, 0, *it.req, (int)it.range, "CUDA error while enqueueing the xxh64 stage");
            xb->items.clear();
            g.xxh_free.push_back(xb);
            return rc;
        }
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
    if (b.timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, b.k0, b.k1) == cudaSuccess) {
            g.be->stats.pack_kernel_ns += (uint64_t)(ms * 1e6);
            g.be->stats.pack_kernel_bytes += b.algo_bytes;
        } else {
            cudaGetLastError();
        }
        b.timed = false;
    }
    for (size_t i = 0; i < b.desc.size(); i++) {
        const std::shared_ptr<Request>& rp = g.log.at(b.desc[i].seq).reads[b.desc[i].ref].req;
        SegPlan& sp = rp->segs[b.desc[i].seg];
        sp.crc = b.desc[i].kidx != UINT32_MAX ? b.crc_host[b.desc[i].kidx] : (b.desc[i].pull_idx != UINT32_MAX ? g.pull_crc_host[b.desc[i].pull_idx] : 0);
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
    const bool want_crc = be->cfg.checksum & LSIO_CHECKSUM_CRC32C;
    const uint32_t flags = LSIO_CUDA_PACK_STORE | (want_crc ? LSIO_CUDA_PACK_CRC32C : 0);
    uint64_t landed_bytes = 0;
    b->algo_bytes = 0;
    std::vector<cudaEvent_t>& copy_events = b->copy_events;  // staging_mode 1: the H2D copy of every extent in this run
    copy_events.clear();
    for (uint64_t s = seq_begin; s < seq_end; s++) {
        StageEntry& e = g.log.at(s);
        uint32_t pull_j = 0;  // the pull kernel numbered this extent's segments in this same order
        for (uint32_t ri = 0; ri < e.reads.size(); ri++) {
            auto& ref = e.reads[ri];
            if (ref.state != kStageLanded) continue;
            const ReadPlan& rd = ref.req->reads[ref.read];
            for (uint32_t k = 0; k < rd.nsegs; k++) {
                const SegPlan& sp = ref.req->segs[rd.first_seg + k];
                if (e.pulled) {  // moved and checksummed already: only its completion is awaited (copy_event below)
                    b->desc.push_back(Batch::DescRef{s, ri, rd.first_seg + k, UINT32_MAX, want_crc ? (uint32_t)((e.seq % g.n_pull) * kPullMaxSegs + pull_j) : UINT32_MAX});
                    pull_j++;
                    continue;
                }
                lsio_cuda_segment seg;
                seg.dst = (uint64_t)(uintptr_t)(ref.req->buffer->buf + sp.dst_off);
                // an extent that landed directly is already in place: src == dst tells the kernel to checksum only
                seg.src = e.direct ? seg.dst : (uint64_t)(uintptr_t)((be->cfg.zero_copy ? g.host_ring_dev : g.dev_ring) + e.log_off % cap + ref.off + sp.head);
                seg.len = sp.payload_len;
                if (e.direct && !want_crc) {  // nothing left to do on the device for this segment
                    b->desc.push_back(Batch::DescRef{s, ri, rd.first_seg + k, UINT32_MAX, UINT32_MAX});
                    continue;
                }
                b->algo_bytes += e.direct ? sp.payload_len : 2 * sp.payload_len;
                const uint32_t kidx = b->counts.nsegs;
                if (!pack_batch_append(b->lay, b->host_desc, &b->counts, seg, store)) {
                    // cannot happen: device_main sizes every run with the same tile arithmetic
                    RangePlan& p = ref.req->ranges[sp.range];
                    if (!p.failed.exchange(true)) send_error(be, -EOVERFLOW, 0, *ref.req, (int)sp.range, "internal: pack batch overflow");
                    p.segs_left.fetch_sub(1, std::memory_order_acq_rel);
                    continue;
                }
                b->desc.push_back(Batch::DescRef{s, ri, rd.first_seg + k, kidx, UINT32_MAX});
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
    }
    if (b->counts.ntiles) {
        // only the part of the descriptor block this batch filled travels (Seg | tiles | multi are separate arrays)
        int rc = pack_batch_upload(b->lay, b->host_desc, b->dev_desc, b->counts, g.s_pack);
        if (rc) return rc;
        unsigned launches = 0;
        const bool light = be->cfg.pack_light_below_mib && b->counts.payload < ((uint64_t)be->cfg.pack_light_below_mib << 20);
        LSIO_CUDA_TRY(cudaEventRecord(b->k0, g.s_pack));
        rc = pack_batch_launch(g.device, b->lay, b->dev_desc, b->counts, flags | (light ? LSIO_CUDA_PACK_LIGHT : 0), b->crc_dev, g.s_pack, &launches,
                               be->cfg.pack_threads);
        if (rc) return rc;
        LSIO_CUDA_TRY(cudaEventRecord(b->k1, g.s_pack));
        b->timed = true;
        be->stats.pack_launches++;
        be->stats.kernel_launches += launches;
    }
    LSIO_CUDA_TRY(cudaEventRecord(b->done, g.s_pack));
    g.inflight.push_back(b);
    return 0;
}

void device_main(Group& g) {
    Backend* be = g.be;
    const lsio_uring_cuda_config& cfg = be->cfg;
    if (!g.cpus.empty()) {
        cpu_set_t set;
        CPU_ZERO(&set);
        for (int c : g.cpus) CPU_SET(c, &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set);
    }
    cudaSetDevice(g.device);
    prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);  // poll_sleep_us means what it says (default slack: +50 us per sleep)
    const uint64_t cap = g.log.capacity();
    const size_t max_segs = g.batches[0].lay.max_segs, max_tiles = g.batches[0].lay.max_tiles;  // >= 2 extents' worth: see backend.hpp
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
                for (size_t i = 0; i < b->desc.size(); i++) {
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
            uint64_t oldest = 0, newest = 0;
            bool hit_limit = false, needs_kernel = false;
            for (; s < avail; s++) {
                StageEntry& e = g.log.at(s);
                const uint32_t st = e.state.load(std::memory_order_acquire);
                if (st == kStageReading) break;
                if (e.log_off % cap < run_phys0) {  // the ring wrapped: a copy must be physically contiguous
                    hit_limit = true;
                    break;
                }
                // descriptors and kernel tiles of its landed reads, counted by the worker that published it
                const size_t t = e.ntiles, ns = e.nsegs;
                if (ns && !oldest) oldest = e.landed_ns;
                newest = std::max(newest, e.landed_ns);
                if (nsegs + ns > max_segs || ntiles + t > max_tiles || (bytes && bytes + e.len > max_batch_bytes)) {
                    hit_limit = true;
                    break;
                }
                nsegs += ns;
                ntiles += t;
                needs_kernel |= ns && !e.pulled && !(e.direct && !(cfg.checksum & LSIO_CHECKSUM_CRC32C));
                bytes = e.log_off + e.len - g.log.at(g.copy_seq).log_off;
            }
            if (s > g.copy_seq) {
                const bool draining = g.stop_device.load(std::memory_order_relaxed);
                const uint64_t now = now_ns();
                // a run that launches nothing (every extent was moved and checksummed by the pull kernel) gains nothing from
                // growing: it only waits long enough to share its event waits and its wake-up of the consumer
                const uint64_t wait_us = needs_kernel ? cfg.flush_us : std::min<uint32_t>(cfg.flush_us, 200);
                const bool timed_out = oldest && now - oldest > wait_us * 1000;
                // Nothing else is being read and nothing has landed for a moment: the burst is over, do not make a
                // lone request wait out flush_us (batches are kept large only while extents keep arriving).
                const bool idle = s == avail && newest && now - newest > (uint64_t)std::min<uint32_t>(cfg.flush_us, 100) * 1000;
                if (hit_limit || bytes >= cfg.min_copy_bytes || timed_out || idle || draining || (!oldest)) {
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

}  // namespace lsio
