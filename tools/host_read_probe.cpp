// host_read_probe.cpp -- CPU-only: what does one thread pay to "read" a memory-backed file?
// (tmpfs reads are kernel memcpys; this is the source ceiling of the e2e bench on this pool.)
//   g++ -O2 -std=c++17 -pthread -Ilight-speed-io_b200/csrc tools/host_read_probe.cpp -o /tmp/host_read_probe
//   /tmp/host_read_probe <dir> <n_files> <file_bytes> <threads> <mode> <block_kib> <qd> <window_kib> [o_direct]
// mode: 0 = io_uring READ_FIXED, 1 = pread, 2 = io_uring READ (not fixed), 3 = mmap + memcpy
// window_kib: bytes of destination buffer each thread cycles through (cache-resident or not)
#include <fcntl.h>
#include <sys/stat.h>
#include <dirent.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "uring.hpp"

using namespace lsio;
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int count_iowq() {
    int n = 0;
    DIR* d = opendir("/proc/self/task");
    if (!d) return -1;
    while (dirent* e = readdir(d)) {
        std::string p = std::string("/proc/self/task/") + e->d_name + "/comm";
        FILE* f = fopen(p.c_str(), "r");
        if (!f) continue;
        char buf[64] = {0};
        if (fgets(buf, sizeof buf, f) && !strncmp(buf, "iou-wrk", 7)) n++;
        fclose(f);
    }
    closedir(d);
    return n;
}

int main(int argc, char** argv) {
    if (argc < 9) return fprintf(stderr, "usage\n"), 2;
    std::string dir = argv[1];
    const int n_files = atoi(argv[2]);
    const uint64_t fsize = strtoull(argv[3], 0, 10);
    const int threads = atoi(argv[4]), mode = atoi(argv[5]);
    const uint64_t block = strtoull(argv[6], 0, 10) << 10;
    const unsigned qd = atoi(argv[7]);
    const uint64_t window = std::max<uint64_t>(strtoull(argv[8], 0, 10) << 10, block * qd);
    const int odirect = argc > 9 ? atoi(argv[9]) : 1;
    std::vector<std::string> paths;
    for (int i = 0; i < n_files; i++) {
        paths.push_back(dir + "/probe_" + std::to_string(i));
        struct stat st;
        if (stat(paths[i].c_str(), &st) || (uint64_t)st.st_size != fsize) {
            int fd = open(paths[i].c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0644);
            std::vector<char> b(1 << 20, (char)i);
            for (uint64_t o = 0; o < fsize; o += b.size()) if (write(fd, b.data(), std::min<uint64_t>(b.size(), fsize - o)) < 0) return 3;
            close(fd);
        }
    }
    std::atomic<int> next{0};
    std::atomic<uint64_t> total{0};
    std::atomic<int> iowq_max{0};
    std::vector<double> busy(threads);
    const double t0 = now_s();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t] {
            uint8_t* buf = (uint8_t*)aligned_alloc(4096, window);
            memset(buf, 1, window);
            Ring ring;
            unsigned sq = 4;
            while (sq < qd) sq <<= 1;
            if (mode == 0 || mode == 2) {
                if (ring.init(sq, 0)) { perror("ring"); exit(1); }
                iovec iv{buf, (size_t)window};
                if (mode == 0 && ring.register_buffers(&iv, 1)) { perror("reg"); exit(1); }
            }
            uint64_t bytes = 0, wpos = 0;
            const double b0 = now_s();
            for (int f; (f = next.fetch_add(1)) < n_files;) {
                int fd = open(paths[f].c_str(), O_RDONLY | (odirect ? O_DIRECT : 0));
                if (fd < 0) { perror("open"); exit(1); }
                const uint64_t nblk = (fsize + block - 1) / block;
                if (mode == 1) {
                    for (uint64_t k = 0; k < nblk; k++) {
                        if (wpos + block > window) wpos = 0;
                        ssize_t r = pread(fd, buf + wpos, block, k * block);
                        if (r < 0) { perror("pread"); exit(1); }
                        bytes += r; wpos += block;
                    }
                } else if (mode == 3) {
                    uint8_t* m = (uint8_t*)mmap(nullptr, fsize, PROT_READ, MAP_SHARED, fd, 0);
                    for (uint64_t k = 0; k < nblk; k++) {
                        if (wpos + block > window) wpos = 0;
                        const uint64_t n = std::min(block, fsize - k * block);
                        memcpy(buf + wpos, m + k * block, n);
                        bytes += n; wpos += block;
                    }
                    munmap(m, fsize);
                } else {
                    uint64_t issued = 0, done = 0;
                    while (done < nblk) {
                        while (issued < nblk && issued - done < qd) {
                            io_uring_sqe* e = ring.get_sqe();
                            if (!e) break;
                            if (wpos + block > window) wpos = 0;
                            e->opcode = mode == 0 ? IORING_OP_READ_FIXED : IORING_OP_READ;
                            e->fd = fd;
                            e->addr = (uint64_t)(uintptr_t)(buf + wpos);
                            e->len = (uint32_t)block;
                            e->off = issued * block;
                            e->buf_index = 0;
                            wpos += block;
                            issued++;
                        }
                        ring.submit(1);
                        io_uring_cqe c;
                        while (ring.pop_cqe(&c)) {
                            if (c.res < 0) { fprintf(stderr, "cqe %d\n", c.res); exit(1); }
                            bytes += c.res;
                            done++;
                        }
                    }
                    if (t == 0) { int n = count_iowq(); if (n > iowq_max.load()) iowq_max.store(n); }
                }
                close(fd);
            }
            busy[t] = now_s() - b0;
            total.fetch_add(bytes);
            free(buf);
        });
    for (auto& x : th) x.join();
    const double secs = now_s() - t0;
    printf("mode=%d threads=%d block=%lluK qd=%u window=%lluK odirect=%d: %.2f GiB/s total, %.2f GiB/s per thread, iowq_threads=%d\n", mode, threads,
           (unsigned long long)(block >> 10), qd, (unsigned long long)(window >> 10), odirect, total.load() / secs / (1 << 30), total.load() / secs / (1 << 30) / threads,
           iowq_max.load());
    return 0;
}
