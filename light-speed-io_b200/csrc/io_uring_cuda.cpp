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
#include "backend.hpp"

namespace lsio {

static thread_local std::string g_last_error;
static void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
// ---- ordered groups (lsio_io/src/lib.rs:6-8; bookkeeping in op_groups.hpp) --------------------------------
Request::~Request() {
    if (fd >= 0) close(fd);  // only reached when the backend is torn down with reads still queued
    if (!be) return;
    std::vector<std::pair<size_t, Op>> go;
    be->op_groups.finished(op_group, !be->closing.load(), &go);
    for (auto& [gi, op] : go) be->groups[gi]->pool->push(std::move(op));
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

// Size of the last-level cache shared by this process's CPUs (sysfs), or 32 MiB if it cannot be read.
static uint64_t llc_bytes() {
    uint64_t best = 0;
    for (int idx = 2; idx <= 4; idx++) {
        char path[128];
        snprintf(path, sizeof path, "/sys/devices/system/cpu/cpu0/cache/index%d/size", idx);
        FILE* f = fopen(path, "r");
        if (!f) continue;
        unsigned long long v = 0;
        char unit = 0;
        if (fscanf(f, "%llu%c", &v, &unit) >= 1) {
            if (unit == 'K') v <<= 10;
            else if (unit == 'M') v <<= 20;
            if (idx >= 3 || !best) best = std::max<uint64_t>(best, v);
        }
        fclose(f);
    }
    return best ? best : (32ull << 20);
}

static void free_group_cuda(Group& g) {
    DeviceGuard dg(g.device);
    for (Batch& b : g.batches) {
        if (b.k0) cudaEventDestroy(b.k0);
        if (b.k1) cudaEventDestroy(b.k1);
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
    if (g.scratch_block) cudaFree(g.scratch_block);
    g.scratch_block = nullptr;
    if (g.pull_crc_host) cudaFreeHost(g.pull_crc_host);
    g.pull_crc_host = nullptr;
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
    if (cfg.cpu_count) {  // an explicit CPU slice (one process per GPU on a shared box: disjoint slices per rank)
        for (uint32_t c = 0; c < cfg.cpu_count; c++) g.cpus.push_back((int)(cfg.cpu_first + c));
    } else if (cfg.pin_threads) {
        g.cpus = numa_cpus_of_device(g.device);
    }
    const uint64_t ring_bytes = cfg.staging_bytes / 4096 * 4096;
    if (cfg.staging_mode == 1) {
        // per-worker cache-sized windows of pinned slots; the ordered log lives on the device only
        g.slot_bytes = std::max<uint64_t>((uint64_t)cfg.host_slot_kib << 10, (cfg.max_read_bytes + 4095) / 4096 * 4096);
        const uint64_t per_worker = g.slot_bytes * cfg.host_slots_per_worker;
        // + 64 KiB: slots start at a 64 KiB boundary whatever the allocator returns (O_DIRECT memory alignment up to 64 KiB)
        LSIO_CUDA_TRY(cudaHostAlloc((void**)&g.slot_block, per_worker * cfg.n_worker_threads + 65536, cudaHostAllocPortable | cudaHostAllocMapped));
        uint8_t* slots0 = (uint8_t*)(((uintptr_t)g.slot_block + 65535) & ~(uintptr_t)65535);
        uint8_t* block_dev = nullptr;
        LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&block_dev, g.slot_block, 0));
        // pull kernel: per-slot device scratch, and one row of CRC slots per extent that can be live in the log
        const size_t scratch_bytes = (pack_pull_scratch_bytes(g.slot_bytes) + 255) & ~size_t(255);
        const size_t n_slots = (size_t)cfg.host_slots_per_worker * cfg.n_worker_threads;
        LSIO_CUDA_TRY(cudaMalloc((void**)&g.scratch_block, scratch_bytes * n_slots));
        LSIO_CUDA_TRY(cudaMemset(g.scratch_block, 0, scratch_bytes * n_slots));
        g.n_pull = (size_t)std::min<uint64_t>(1 << 16, ring_bytes / g.slot_bytes + 64);
        LSIO_CUDA_TRY(cudaHostAlloc((void**)&g.pull_crc_host, g.n_pull * kPullMaxSegs * sizeof(uint32_t), cudaHostAllocMapped));
        LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&g.pull_crc_dev, g.pull_crc_host, 0));
        g.wslots.resize(cfg.n_worker_threads);
        for (uint32_t w = 0; w < cfg.n_worker_threads; w++) {
            WorkerSlots& ws = g.wslots[w];
            ws.base = slots0 + per_worker * w;
            ws.base_dev = block_dev + (ws.base - g.slot_block);
            for (uint32_t k = 0; k < cfg.host_slots_per_worker; k++) ws.scratch.push_back(g.scratch_block + scratch_bytes * ((size_t)w * cfg.host_slots_per_worker + k));
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
    // The landing path runs at the highest stream priority: a consumer's own kernels on the same GPU (the codec
    // layer decoding the previous batch, a training step) fill every SM for milliseconds, and a pack launch queued
    // behind them stalls the staging log and with it the reads.  With priority the block scheduler places the pack
    // CTAs first as SMs drain (measured with LZ4 decode running beside the I/O: see DESIGN 3.3).
    int prio_lo = 0, prio_hi = 0;
    LSIO_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    LSIO_CUDA_TRY(cudaStreamCreateWithPriority(&g.s_copy, cudaStreamNonBlocking, prio_hi));
    LSIO_CUDA_TRY(cudaStreamCreateWithPriority(&g.s_pack, cudaStreamNonBlocking, prio_hi));
    const size_t max_segs = kBatchMaxSegs, max_tiles = kBatchMaxTiles;
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
        LSIO_CUDA_TRY(cudaEventCreate(&b.k0));
        LSIO_CUDA_TRY(cudaEventCreate(&b.k1));
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
    // A pack launch is worth batching: every launch re-stages 148 x 128 KiB of CRC tables and competes with the H2D
    // DMA.  Measured on C2 (alternating runs on one box): 8 MiB / 200 us 41.7 GiB/s, 32 MiB / 1 ms 44.7, 64 MiB / 2 ms
    // 45.5, 128 MiB / 4 ms (2 GiB log) 45.6.  A burst that ends is flushed after 100 us of silence (device_stage.cpp).
    cfg->min_copy_bytes = 64ull << 20;
    cfg->flush_us = 2000;
    cfg->merge_ranges = 1;
    cfg->merge_gap_bytes = 0;
    cfg->statx_via_ring = 0;
    cfg->zero_copy = 0;
    cfg->staging_mode = 1;
    cfg->host_slots_per_worker = 0;  // auto: see window_budget_kib
    cfg->host_slot_kib = 0;          // auto
    cfg->pack_threads = 0;
    cfg->poll_sleep_us = 50;  // measured: same throughput as spinning (33-34 GiB/s C2, 27.7 vs 28.1 C3) without burning a core per GPU
    cfg->inline_reads = 1;
    cfg->force_align = 0;
    cfg->debug_short_read_every = 0;
    cfg->null_source = 0;
    cfg->direct_landing = 2;
    cfg->cpu_first = cfg->cpu_count = 0;
    cfg->steal_batch = 8;
    cfg->pack_light_below_mib = 512;
    cfg->copy_coalesce_slots = 2;
    cfg->window_budget_kib = 0;
    cfg->pull_any_shift = 1;
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
    if (cfg.staging_mode == 1 && (cfg.host_slots_per_worker == 0 || cfg.host_slot_kib == 0)) {
        // Size the pinned windows so that ALL of them stay in the last-level cache: a slot is written by the reading core
        // and read back by the GPU a moment later; if it has been evicted in between, every payload byte costs three more
        // DRAM transfers (write-back, re-fetch for ownership, DMA read) on top of the read of the source.  Measured on
        // this pool (60 MiB L3): 24 workers x 2 MiB windows 91.6 GiB/s, x 4 MiB 59.4; 32 workers x 2 MiB 77, x 1.5 MiB 80.6.
        uint64_t budget = (uint64_t)cfg.window_budget_kib << 10;
        if (!budget) budget = llc_bytes() * 6 / 10 / std::max<uint32_t>(1, cfg.n_devices);
        const uint64_t per_worker = budget / std::max<uint32_t>(1, cfg.n_worker_threads);
        uint64_t slot = cfg.host_slot_kib ? (uint64_t)cfg.host_slot_kib << 10 : 512u << 10;
        if (!cfg.host_slot_kib && per_worker < 2 * slot) slot = 256u << 10;  // two slots must fit: halve them
        if (!cfg.host_slot_kib) {
            cfg.host_slot_kib = (uint32_t)(slot >> 10);
            cfg.max_read_bytes = std::min<uint64_t>(cfg.max_read_bytes, slot);
        }
        if (!cfg.host_slots_per_worker) cfg.host_slots_per_worker = (uint32_t)std::min<uint64_t>(4, std::max<uint64_t>(2, per_worker / slot));
    }
    if (cfg.staging_mode == 1) {
        // every worker must be able to hold two extents in the device log at once, and a slot holds at least one read
        const uint64_t slot_bytes = std::max<uint64_t>((uint64_t)cfg.host_slot_kib << 10, (cfg.max_read_bytes + 4095) / 4096 * 4096);
        if (cfg.host_slot_kib == 0 || slot_bytes > cfg.staging_bytes / 2) {
            set_error("staging_mode 1: host_slot_kib must be > 0 and a slot (max(host_slot_kib KiB, max_read_bytes) = %llu B) at most staging_bytes / 2",
                      (unsigned long long)slot_bytes);
            return -EINVAL;
        }
    }
    if (cfg.force_align && ((cfg.force_align & (cfg.force_align - 1)) || cfg.force_align < kAlign || cfg.force_align > 65536 || cfg.force_align > cfg.max_read_bytes)) {
        set_error("force_align must be a power of two in [512, 65536] and at most max_read_bytes");
        return -EINVAL;
    }
    if (cfg.copy_coalesce_slots == 0) cfg.copy_coalesce_slots = 1;
    if (cfg.copy_coalesce_slots > 8) {
        set_error("copy_coalesce_slots must be at most 8");
        return -EINVAL;
    }
    if (cfg.direct_landing > 2) {
        set_error("direct_landing must be 0 (ring), 1 (copy engine) or 2 (pull kernel)");
        return -EINVAL;
    }
    if (cfg.inline_reads > 2) {
        set_error("inline_reads must be 0, 1 or 2");
        return -EINVAL;
    }
    if (cfg.staging_mode == 1 && (cfg.host_slots_per_worker == 0 || cfg.host_slots_per_worker > 256 || cfg.n_worker_threads > 64)) {
        set_error("staging_mode 1 needs 1..256 host_slots_per_worker and at most 64 worker threads per device");
        return -EINVAL;
    }
    if (cfg.n_worker_threads == 0 || cfg.n_devices == 0 || cfg.n_devices > LSIO_MAX_DEVICES || cfg.sq_entries < 4 ||
        (cfg.sq_entries & (cfg.sq_entries - 1)) || cfg.staging_bytes < (4ull << 20) || cfg.arena_bytes < (1ull << 20) || cfg.max_read_bytes < kAlign ||
        cfg.max_read_bytes > (1ull << 30) || cfg.max_read_bytes > cfg.staging_bytes / 4 || cfg.pack_threads % 32 || cfg.pack_threads > 1024) {
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
            worker_main(*g, wt);
            g->workers_alive.fetch_sub(1, std::memory_order_acq_rel);
            g->be->out->remove_sender();
        }, g->cpus, cfg.steal_batch);
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
    r->submit_ns = now_ns();
    {
        uint64_t zero = 0;
        be->stats.first_submit_ns.compare_exchange_strong(zero, r->submit_ns);
    }
    const size_t gi = group >= 0 ? (size_t)group : (size_t)(be->rr.fetch_add(1) % be->groups.size());
    if (gi >= be->groups.size()) return -EINVAL;
    r->grp = be->groups[gi].get();
    std::pair<size_t, Op> item;
    item.first = gi;
    item.second.kind = Op::GetRanges;
    item.second.req = r;
    r->be = be;  // from here on ~Request reports to the group bookkeeping
    if (!be->op_groups.submit(item, &r->op_group)) return 0;  // an earlier group is still in flight: held back, unopened
    r.reset();
    be->groups[gi]->pool->push(std::move(item.second));  // threadpool.rs:134-137
    return 0;
}

uint64_t lsio_uring_cuda_next_group(lsio_uring_cuda* self) { return self ? self->be->op_groups.next_group() : 0; }
uint64_t lsio_uring_cuda_current_group(const lsio_uring_cuda* self) { return self ? self->be->op_groups.current() : 0; }

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
    o->inline_reads = s.inline_reads; o->direct_extents = s.direct_extents; o->direct_copies = s.direct_copies; o->read_ns = s.read_ns;
    o->copy_enqueue_ns = s.copy_enqueue_ns; o->worker_stall_ns = s.worker_stall_ns; o->worker_busy_ns = s.worker_busy_ns;
    o->pack_kernel_ns = s.pack_kernel_ns; o->pack_kernel_bytes = s.pack_kernel_bytes; o->first_chunk_ns = s.first_chunk_ns;
    o->max_dio_align = s.max_dio_align; o->pull_launches = s.pull_launches;
    o->slot_bytes = self->be->groups.empty() ? 0 : self->be->groups[0]->slot_bytes; o->slots_per_worker = self->be->cfg.staging_mode == 1 ? self->be->cfg.host_slots_per_worker : 0;
    return 0;
}

int lsio_uring_cuda_get_latency_histogram(const lsio_uring_cuda* self, uint64_t buckets[32]) {
    if (!self || !buckets) return -EINVAL;
    for (int i = 0; i < 32; i++) buckets[i] = self->be->stats.latency[i].load(std::memory_order_relaxed);
    return 0;
}
void lsio_uring_cuda_reset_latency(lsio_uring_cuda* self) {
    if (!self) return;
    Stats& s = self->be->stats;
    for (auto& b : s.latency) b.store(0, std::memory_order_relaxed);
    s.first_submit_ns.store(0);
    s.first_chunk_ns.store(0);
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
    be->op_groups.take_held();  // requests of later groups that were never admitted: dropped unopened (outside the lock)
    for (auto& g : be->groups) free_group_cuda(*g);
    delete self;  // chunks already handed out keep the arena alive through their InnerBuffer
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
