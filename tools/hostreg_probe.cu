// hostreg_probe.cu -- how fast can a memory-backed (tmpfs) file be pinned IN PLACE for the GPU (mmap + cudaHostRegister) and
// pulled from there, compared with copying it through a pinned slot?  Decides whether a "no bounce buffer" landing for
// memory-backed files is worth building.   usage: hostreg_probe <dir with files> <n_files> <threads>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void pull(const uint4* s, uint4* d, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string dir = argv[1];
    const int n_files = atoi(argv[2]), T = atoi(argv[3]);
    cudaFree(0);
    uint8_t* dev;
    cudaMalloc(&dev, (size_t)T * (64 << 20));
    for (unsigned flags : {(unsigned)cudaHostRegisterMapped, (unsigned)(cudaHostRegisterMapped | cudaHostRegisterReadOnly)}) {
        for (int populate = 0; populate < 2; populate++) {
            std::atomic<int> next{0};
            std::atomic<uint64_t> bytes{0}, reg_ns{0}, unreg_ns{0}, map_ns{0}, pull_ns{0};
            std::atomic<int> fails{0};
            const double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    cudaSetDevice(0);
                    cudaStream_t st;
                    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
                    for (int f; (f = next.fetch_add(1)) < n_files;) {
                        const std::string p = dir + "/lsio_bench_" + std::to_string(f);
                        int fd = open(p.c_str(), O_RDWR);
                        if (fd < 0) { fails++; continue; }
                        struct stat sb;
                        fstat(fd, &sb);
                        const size_t len = (size_t)sb.st_size / 4096 * 4096;
                        double a = now();
                        void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED | (populate ? MAP_POPULATE : 0), fd, 0);
                        double b = now();
                        if (m == MAP_FAILED) { fails++; close(fd); continue; }
                        cudaError_t e = cudaHostRegister(m, len, flags);
                        double c = now();
                        if (e != cudaSuccess) {
                            if (fails.fetch_add(1) == 0) fprintf(stderr, "cudaHostRegister: %s\n", cudaGetErrorString(e));
                            cudaGetLastError();
                            munmap(m, len);
                            close(fd);
                            continue;
                        }
                        void* alias = nullptr;
                        cudaHostGetDevicePointer(&alias, m, 0);
                        pull<<<64, 256, 0, st>>>((const uint4*)alias, (uint4*)(dev + (size_t)t * (64 << 20)), len / 16);
                        cudaStreamSynchronize(st);
                        double d = now();
                        cudaHostUnregister(m);
                        double g = now();
                        munmap(m, len);
                        close(fd);
                        bytes += len;
                        map_ns += (uint64_t)((b - a) * 1e9);
                        reg_ns += (uint64_t)((c - b) * 1e9);
                        pull_ns += (uint64_t)((d - c) * 1e9);
                        unreg_ns += (uint64_t)((g - d) * 1e9);
                    }
                });
            for (auto& x : th) x.join();
            const double secs = now() - t0;
            const double gb = bytes.load() / 1e9;
            printf("{\"flags\": %u, \"populate\": %d, \"threads\": %d, \"files\": %d, \"fails\": %d, \"gib_s\": %.2f, \"per_thread_GBps\": {\"mmap\": %.2f, \"register\": %.2f, \"pull\": %.2f, \"unregister\": %.2f}}\n", flags, populate, T,
                   n_files, fails.load(), bytes.load() / secs / (1 << 30), gb / (map_ns.load() * 1e-9 + 1e-12), gb / (reg_ns.load() * 1e-9 + 1e-12), gb / (pull_ns.load() * 1e-9 + 1e-12),
                   gb / (unreg_ns.load() * 1e-9 + 1e-12));
            fflush(stdout);
        }
    }
    return 0;
}
