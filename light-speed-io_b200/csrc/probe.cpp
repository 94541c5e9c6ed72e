// probe.cpp -- lsio_probe_read_ceiling: raw READ_FIXED read rate of a file set with no GPU work
// (measurement aid; stands in for crates/lsio_uring/benches/fio.ini).
#include <sys/uio.h>

#include "backend.hpp"

using namespace lsio;

extern "C" {

int lsio_probe_read_ceiling(const char* const* paths, size_t n_paths, uint64_t block_bytes, uint32_t n_threads, uint32_t queue_depth,
                            uint32_t o_direct, double* gib_per_s) {
    return lsio_probe_read_ceiling_ex(paths, n_paths, block_bytes, n_threads, queue_depth, o_direct, 0, 1, gib_per_s);
}

int lsio_probe_read_ceiling_ex(const char* const* paths, size_t n_paths, uint64_t block_bytes, uint32_t n_threads, uint32_t queue_depth,
                               uint32_t o_direct, uint32_t inline_reads, uint32_t files_in_flight, double* gib_per_s) {
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
            // the product's own rule (worker.cpp configure_request): a file the kernel cannot read asynchronously is
            // read with preadv2 on this thread instead of being handed to an io-wq thread by io_uring
            auto is_inline = [&](int fd) {
                if (inline_reads == 2) return true;
                if (inline_reads != 1) return false;
                uint8_t byte;
                iovec iv{&byte, 1};
                return preadv2(fd, &iv, 1, 0, RWF_NOWAIT) < 0 && errno == EOPNOTSUPP;
            };
            // Asynchronous reads: up to `files_in_flight` files open at once, their blocks issued round robin, so that the
            // queue stays full across file boundaries (as the backend keeps it).
            struct OpenFile {
                int fd = -1;
                uint64_t nblk = 0, issued = 0, done = 0;
            };
            std::vector<OpenFile> of(files_in_flight ? files_in_flight : 1);
            std::vector<unsigned> free_slots;
            for (unsigned i = 0; i < queue_depth; i++) free_slots.push_back(i);
            bool more = true;
            unsigned inflight = 0, rr = 0;
            uint8_t* base = pinned + (uint64_t)t * slot_bytes;
            for (;;) {
                for (auto& o : of) {
                    while (o.fd < 0 && more) {
                        const size_t f = next_file.fetch_add(1);
                        if (f >= n_paths) { more = false; break; }
                        int fd = open(paths[f], O_RDONLY | (o_direct ? O_DIRECT : 0));
                        if (fd < 0) { err.store(-errno); more = false; break; }
                        struct stat st;
                        fstat(fd, &st);
                        const uint64_t nblk = ((uint64_t)st.st_size + block_bytes - 1) / block_bytes;
                        if (is_inline(fd)) {
                            for (uint64_t k = 0; k < nblk; k++) {
                                iovec iv{base + (k % queue_depth) * block_bytes, (size_t)block_bytes};
                                const ssize_t n = preadv2(fd, &iv, 1, (off_t)(k * block_bytes), 0);
                                if (n < 0) { err.store(-errno); break; }
                                bytes += (uint64_t)n;
                            }
                            close(fd);
                            continue;
                        }
                        o = OpenFile{fd, nblk, 0, 0};
                        if (nblk == 0) { close(fd); o.fd = -1; }
                    }
                }
                bool any_open = false;
                for (auto& o : of) any_open |= o.fd >= 0;
                if (!any_open && inflight == 0) break;
                while (!free_slots.empty()) {
                    OpenFile* pick = nullptr;
                    for (size_t k = 0; k < of.size(); k++) {
                        OpenFile& o = of[(rr + k) % of.size()];
                        if (o.fd >= 0 && o.issued < o.nblk) { pick = &o; rr = (unsigned)((rr + k + 1) % of.size()); break; }
                    }
                    if (!pick) break;
                    io_uring_sqe* e = ring.get_sqe();
                    if (!e) break;
                    const unsigned slot = free_slots.back();
                    free_slots.pop_back();
                    e->opcode = IORING_OP_READ_FIXED;
                    e->fd = pick->fd;
                    e->addr = (uint64_t)(uintptr_t)(base + (uint64_t)slot * block_bytes);
                    e->len = (uint32_t)block_bytes;
                    e->off = pick->issued * block_bytes;
                    e->buf_index = 0;
                    e->user_data = ((uint64_t)(pick - of.data()) << 32) | slot;
                    pick->issued++;
                    inflight++;
                }
                if (inflight == 0) continue;
                ring.submit(1);
                io_uring_cqe c;
                while (ring.pop_cqe(&c)) {
                    if (c.res < 0) err.store(c.res);
                    else bytes += (uint64_t)c.res;
                    free_slots.push_back((unsigned)(c.user_data & 0xFFFFFFFFu));
                    OpenFile& o = of[c.user_data >> 32];
                    inflight--;
                    if (++o.done == o.nblk) { close(o.fd); o.fd = -1; }
                }
                if (err.load()) {
                    // drain what is in flight, then stop
                    more = false;
                    for (auto& o : of) o.issued = o.nblk;
                    if (inflight == 0) { for (auto& o : of) if (o.fd >= 0) { close(o.fd); o.fd = -1; } }
                }
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

// The minimum any host-staged path must do for a memory-backed source, and nothing else: every thread reads its files block by block
// with preadv2 into a small cycle of pinned, mapped slots and sends each filled slot to the GPU (pull kernel without checksum, or one
// cudaMemcpyAsync) into a scratch buffer, waiting only for the slot's previous transfer.  No requests, no planning, no arena, no
// completion channel, no checksums, nothing is kept.  With several ranks on one host every rank runs it at the same time: the sum is what
// the host can stage AND ship at once -- the ceiling the pipeline is compared with where the source probe alone (no DMA sharing the
// memory system) over-promises.
int lsio_probe_staged_ceiling(const char* const* paths, size_t n_paths, int device, uint32_t n_threads, uint32_t slot_kib, uint32_t slots_per_thread,
                              uint32_t use_copy_engine, uint32_t cpu_first, uint32_t cpu_count, double* seconds, uint64_t* bytes_out) {
    if (!paths || !n_paths || !seconds || !bytes_out || !n_threads || !slot_kib || !slots_per_thread || slots_per_thread > 64) return -EINVAL;
    DeviceGuard dg(device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    int rc = pack_prepare_device(device);
    if (rc) return rc;
    const uint64_t slot = (uint64_t)slot_kib << 10;
    const size_t n_slots = (size_t)n_threads * slots_per_thread;
    uint8_t *pinned = nullptr, *pinned_dev = nullptr, *dev = nullptr, *scratch = nullptr;
    const size_t sb = (pack_pull_scratch_bytes(slot) + 255) & ~size_t(255);
    LSIO_CUDA_TRY(cudaHostAlloc((void**)&pinned, slot * n_slots, cudaHostAllocPortable | cudaHostAllocMapped));
    LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&pinned_dev, pinned, 0));
    LSIO_CUDA_TRY(cudaMalloc((void**)&dev, slot * n_slots));
    LSIO_CUDA_TRY(cudaMalloc((void**)&scratch, sb * n_slots));
    LSIO_CUDA_TRY(cudaMemset(scratch, 0, sb * n_slots));
    std::atomic<size_t> next_file{0};
    std::atomic<uint64_t> total{0};
    std::atomic<int> err{0};
    std::atomic<uint32_t> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; t++)
        th.emplace_back([&, t] {
            if (cpu_count) {
                cpu_set_t set;
                CPU_ZERO(&set);
                for (uint32_t c = 0; c < cpu_count; c++) CPU_SET((int)(cpu_first + c), &set);
                pthread_setaffinity_np(pthread_self(), sizeof set, &set);
            }
            cudaSetDevice(device);
            cudaStream_t st;
            std::vector<cudaEvent_t> ev(slots_per_thread);
            std::vector<char> busy(slots_per_thread, 0);
            bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
            for (auto& e : ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
            if (!ok) err.store(-EIO);
            ready.fetch_add(1);
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            uint64_t bytes = 0;
            uint32_t k = 0;
            for (size_t f; ok && !err.load() && (f = next_file.fetch_add(1)) < n_paths;) {
                int fd = open(paths[f], O_RDONLY);
                if (fd < 0) { err.store(-errno); break; }
                for (uint64_t off = 0;; off += slot, k = (k + 1) % slots_per_thread) {
                    const size_t si = (size_t)t * slots_per_thread + k;
                    if (busy[k]) while (cudaEventQuery(ev[k]) == cudaErrorNotReady) __builtin_ia32_pause();
                    iovec iv{pinned + si * slot, (size_t)slot};
                    const ssize_t n = preadv2(fd, &iv, 1, (off_t)off, 0);
                    if (n < 0) { err.store(-errno); break; }
                    if (n == 0) break;
                    cudaError_t ce = cudaSuccess;
                    int prc = 0;
                    if (use_copy_engine) {
                        ce = cudaMemcpyAsync(dev + si * slot, pinned + si * slot, (size_t)n, cudaMemcpyHostToDevice, st);
                    } else {
                        lsio_cuda_segment sg{(uint64_t)(uintptr_t)(pinned_dev + si * slot), (uint64_t)(uintptr_t)(dev + si * slot), (uint64_t)n};
                        prc = pack_pull_launch(device, &sg, 1, false, scratch + si * sb, pack_pull_max_tiles(slot), nullptr, st);
                    }
                    if (ce == cudaSuccess && !prc) ce = cudaEventRecord(ev[k], st);
                    if (ce != cudaSuccess || prc) { err.store(prc ? prc : (int)ce); break; }
                    busy[k] = 1;
                    bytes += (uint64_t)n;
                    if ((uint64_t)n < slot) { k = (k + 1) % slots_per_thread; break; }
                }
                close(fd);
            }
            if (ok) cudaStreamSynchronize(st);
            for (auto& e : ev) cudaEventDestroy(e);
            cudaStreamDestroy(st);
            total.fetch_add(bytes);
        });
    while (ready.load() < n_threads) std::this_thread::yield();
    const uint64_t t0 = now_ns();
    go.store(true, std::memory_order_release);
    for (auto& x : th) x.join();
    cudaDeviceSynchronize();
    *seconds = (double)(now_ns() - t0) * 1e-9;
    *bytes_out = total.load();
    cudaFreeHost(pinned);
    cudaFree(dev);
    cudaFree(scratch);
    return err.load();
}

}  // extern "C"
