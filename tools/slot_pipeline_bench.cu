// Micro-benchmark (not product code): per-thread cache-sized slot rings.  N threads, each with a private
// io_uring, K pinned slots of S bytes and a private CUDA stream: read (tmpfs, READ_FIXED) into a slot,
// cudaMemcpyAsync the slot to HBM, reuse the slot when its copy event is done.  Reports GiB/s into HBM.
#include <cuda_runtime.h>
#include <linux/io_uring.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <sys/uio.h>
#include <fcntl.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "../light-speed-io_b200/csrc/uring.hpp"
using namespace lsio;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    const char* dir = argc > 1 ? argv[1] : "/dev/shm/lsio_bench";
    const int nfiles = argc > 2 ? atoi(argv[2]) : 200;
    const bool noread = getenv("NOREAD") != nullptr;  // pure H2D copy pattern, no io_uring reads
    for (int nthreads : {2, 4, 8}) for (size_t S : {1ul << 20, 2ul << 20, 4ul << 20}) for (int K : {2, 4}) {
        
        uint8_t* dev = nullptr; const size_t devbytes = 1ull << 30; cudaMalloc(&dev, devbytes);
        std::atomic<int> next_file{0}, ready{0}; std::atomic<uint64_t> total{0}; std::atomic<bool> go{false};
        double t0 = 0;
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] {
            cudaSetDevice(0);
            uint8_t* slots = nullptr; cudaHostAlloc((void**)&slots, (size_t)K * S, cudaHostAllocPortable);
            cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            std::vector<cudaEvent_t> ev(K); for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            std::vector<char> busy(K, 0);
            Ring ring; ring.init(8, 0); iovec iv{slots, (size_t)K * S}; ring.register_buffers(&iv, 1);
            uint64_t bytes = 0, devoff = 0; int slot = 0; io_uring_sqe dummy;
            ready++; while (!go.load()) std::this_thread::yield();
            for (int f; (f = next_file.fetch_add(1)) < nfiles * 3;) {
                std::string p = std::string(dir) + "/lsio_bench_" + std::to_string(f % nfiles);
                int fd = open(p.c_str(), O_RDONLY | O_DIRECT); if (fd < 0) break;
                struct stat stt; fstat(fd, &stt);
                for (uint64_t off = 0; off + S <= (uint64_t)stt.st_size; off += S) {
                    if (busy[slot]) { cudaEventSynchronize(ev[slot]); busy[slot] = 0; }
                    io_uring_sqe* e = noread ? &dummy : ring.get_sqe(); e->opcode = IORING_OP_READ_FIXED; e->fd = fd; e->addr = (uint64_t)(uintptr_t)(slots + (size_t)slot * S);
                    e->len = (uint32_t)S; e->off = off; e->buf_index = 0; e->user_data = slot;
                    io_uring_cqe c; c.res = (int)S;
                    if (!noread) { ring.submit(1); while (!ring.pop_cqe(&c)) ring.submit(1); } 
                    if (c.res != (int)S) { fprintf(stderr, "short %d\n", c.res); break; }
                    cudaMemcpyAsync(dev + (uint64_t)t * (devbytes / nthreads) + (devoff % (devbytes / nthreads)), slots + (size_t)slot * S, S, cudaMemcpyHostToDevice, st);
                    cudaEventRecord(ev[slot], st); busy[slot] = 1; devoff += S; bytes += S; slot = (slot + 1) % K;
                }
                close(fd);
            }
            cudaStreamSynchronize(st); total += bytes; cudaFreeHost(slots);
        });
        while (ready.load() < nthreads) std::this_thread::yield();
        t0 = now(); go = true;
        for (auto& x : th) x.join();
        const double dt = now() - t0;
        printf("threads=%d slot=%zuK slots/thread=%d window/thread=%zuK total=%zuMiB: %.1f GiB/s into HBM\n", nthreads, S >> 10, K, K * S >> 10, nthreads * K * S >> 20, total.load() / dt / (1 << 30));
        fflush(stdout); cudaFree(dev);
    }
    return 0;
}
