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

}  // extern "C"
