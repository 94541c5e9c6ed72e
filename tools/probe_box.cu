// Box probe: what the GPU box offers the LSIO hot path (cores, storage, io_uring, pinned memory, PCIe).
// Test aid only; not part of the product path.
#include <cuda_runtime.h>
#include <linux/io_uring.h>
#include <sys/syscall.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <fcntl.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <chrono>
#include <vector>
#include <thread>
#include <atomic>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA ERR %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); } } while (0)

struct Ring {
  int fd = -1; io_uring_params p{}; void *sq_ptr = nullptr, *cq_ptr = nullptr; io_uring_sqe* sqes = nullptr;
  unsigned *sq_head, *sq_tail, *sq_mask, *sq_array, *sq_flags, *cq_head, *cq_tail, *cq_mask; io_uring_cqe* cqes;
  int init(unsigned entries, unsigned flags) {
    memset(&p, 0, sizeof p); p.flags = flags; if (flags & IORING_SETUP_SQPOLL) p.sq_thread_idle = 1000;
    fd = syscall(__NR_io_uring_setup, entries, &p); if (fd < 0) return -errno;
    size_t sqsz = p.sq_off.array + p.sq_entries * sizeof(unsigned), cqsz = p.cq_off.cqes + p.cq_entries * sizeof(io_uring_cqe);
    sq_ptr = mmap(0, sqsz, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_SQ_RING);
    cq_ptr = mmap(0, cqsz, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_CQ_RING);
    sqes = (io_uring_sqe*)mmap(0, p.sq_entries * sizeof(io_uring_sqe), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_SQES);
    char* s = (char*)sq_ptr; char* c = (char*)cq_ptr;
    sq_head = (unsigned*)(s + p.sq_off.head); sq_tail = (unsigned*)(s + p.sq_off.tail); sq_mask = (unsigned*)(s + p.sq_off.ring_mask);
    sq_array = (unsigned*)(s + p.sq_off.array); sq_flags = (unsigned*)(s + p.sq_off.flags);
    cq_head = (unsigned*)(c + p.cq_off.head); cq_tail = (unsigned*)(c + p.cq_off.tail); cq_mask = (unsigned*)(c + p.cq_off.ring_mask);
    cqes = (io_uring_cqe*)(c + p.cq_off.cqes); return 0;
  }
};

static double disk_probe(const char* path, size_t fsize, size_t bs, int qd, bool fixed, void* buf, size_t buflen) {
  int fd = open(path, O_RDONLY | O_DIRECT); if (fd < 0) { printf("  open O_DIRECT %s failed errno=%d\n", path, errno); return 0; }
  Ring r; int rc = r.init(256, 0); if (rc) { printf("  ring init failed %d\n", rc); return 0; }
  if (fixed) { iovec iv{buf, buflen}; rc = syscall(__NR_io_uring_register, r.fd, IORING_REGISTER_BUFFERS, &iv, 1); if (rc < 0) { printf("  register buffers failed errno=%d\n", errno); fixed = false; } }
  size_t nblk = fsize / bs, next = 0, done = 0, inflight = 0; double t0 = now(); long long bytes = 0;
  while (done < nblk) {
    unsigned tail = *r.sq_tail; int pushed = 0;
    while (inflight < (size_t)qd && next < nblk) {
      unsigned idx = tail & *r.sq_mask; io_uring_sqe* e = &r.sqes[idx]; memset(e, 0, sizeof *e);
      e->opcode = fixed ? IORING_OP_READ_FIXED : IORING_OP_READ; e->fd = fd; e->off = next * bs;
      e->addr = (unsigned long long)((char*)buf + (next % (buflen / bs)) * bs); e->len = bs; e->buf_index = 0; e->user_data = next;
      r.sq_array[idx] = idx; tail++; next++; inflight++; pushed++;
    }
    __atomic_store_n(r.sq_tail, tail, __ATOMIC_RELEASE);
    int ret = syscall(__NR_io_uring_enter, r.fd, pushed, 1, IORING_ENTER_GETEVENTS, 0, 0);
    if (ret < 0) { printf("  enter failed errno=%d\n", errno); break; }
    unsigned head = *r.cq_head, ctail = __atomic_load_n(r.cq_tail, __ATOMIC_ACQUIRE);
    while (head != ctail) { io_uring_cqe* c = &r.cqes[head & *r.cq_mask]; if (c->res != (int)bs) { printf("  cqe res=%d\n", c->res); } else bytes += c->res; head++; done++; inflight--; }
    __atomic_store_n(r.cq_head, head, __ATOMIC_RELEASE);
  }
  double dt = now() - t0; close(fd); close(r.fd);
  return bytes / dt / (1 << 30);
}

static void make_file(const char* path, size_t fsize) {
  int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644); std::vector<char> b(8 << 20); for (size_t i = 0; i < b.size(); i++) b[i] = (char)(i * 2654435761u >> 13);
  double t0 = now(); for (size_t o = 0; o < fsize; o += b.size()) { if (write(fd, b.data(), b.size()) < 0) { printf("write failed errno=%d\n", errno); break; } }
  fsync(fd); close(fd); printf("  wrote %s %.1f GiB at %.2f GiB/s\n", path, fsize / 1073741824.0, fsize / (now() - t0) / (1 << 30));
  int f2 = open(path, O_RDONLY); posix_fadvise(f2, 0, 0, POSIX_FADV_DONTNEED); close(f2);
}

int main(int argc, char** argv) {
  const char* diskdir = argc > 1 ? argv[1] : "/tmp"; const char* shmdir = "/dev/shm";
  int ndev = 0; CK(cudaGetDeviceCount(&ndev)); printf("cuda devices: %d\n", ndev);
  for (int d = 0; d < ndev; d++) { cudaDeviceProp p; cudaGetDeviceProperties(&p, d); char bus[32]; cudaDeviceGetPCIBusId(bus, 32, d);
    printf(" dev %d %s sm=%d smem/blk optin=%zu L2=%d MB mem=%.1f GB bus=%s canMapHost=%d asyncEngines=%d\n", d, p.name, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.l2CacheSize >> 20, p.totalGlobalMem / 1e9, bus, p.canMapHostMemory, p.asyncEngineCount); }
  // io_uring features
  { Ring r; int rc = r.init(64, 0); printf("io_uring plain: rc=%d sq=%u cq=%u features=0x%x\n", rc, r.p.sq_entries, r.p.cq_entries, r.p.features); if (!rc) close(r.fd);
    Ring q; rc = q.init(64, IORING_SETUP_SQPOLL); printf("io_uring sqpoll: rc=%d\n", rc); if (!rc) close(q.fd); }
  // pinned memory
  size_t big = 1ull << 30; void* hp = nullptr; double t0 = now(); cudaError_t e = cudaHostAlloc(&hp, big, cudaHostAllocDefault); printf("cudaHostAlloc 1GiB: %s %.3fs\n", cudaGetErrorString(e), now() - t0);
  void* mp = nullptr; posix_memalign(&mp, 2 << 20, big); memset(mp, 1, big); t0 = now(); e = cudaHostRegister(mp, big, cudaHostRegisterDefault); printf("cudaHostRegister 1GiB: %s %.3fs\n", cudaGetErrorString(e), now() - t0);
  if (e != cudaSuccess) cudaGetLastError();
  { Ring r; r.init(64, 0); iovec iv{hp, big}; t0 = now(); int rc = syscall(__NR_io_uring_register, r.fd, IORING_REGISTER_BUFFERS, &iv, 1); printf("io_uring register 1GiB cudaHostAlloc'd: rc=%d errno=%d %.3fs\n", rc, rc < 0 ? errno : 0, now() - t0); close(r.fd); }
  // H2D
  void* dp = nullptr; CK(cudaMalloc(&dp, big)); cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (size_t sz : {64ul << 10, 256ul << 10, 1ul << 20, 4ul << 20, 64ul << 20}) {
    for (void* src : {hp, mp}) {
      int n = (int)std::min<size_t>(2000, (4ull << 30) / sz); size_t slots = big / sz; double h0 = now(); cudaEventRecord(a, st);
      for (int i = 0; i < n; i++) cudaMemcpyAsync((char*)dp + (i % slots) * sz, (char*)src + (i % slots) * sz, sz, cudaMemcpyHostToDevice, st);
      double h1 = now(); cudaEventRecord(b, st); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
      printf("H2D %s sz=%zuK n=%d: %.2f GB/s  issue %.2f us/copy\n", src == hp ? "hostalloc" : "registered", sz >> 10, n, n * (double)sz / ms / 1e6, (h1 - h0) / n * 1e6);
    }
  }
  // D2D copy rate (sanity of HBM)
  { void* d2; CK(cudaMalloc(&d2, big)); cudaEventRecord(a, st); for (int i = 0; i < 20; i++) cudaMemcpyAsync(d2, dp, big, cudaMemcpyDeviceToDevice, st); cudaEventRecord(b, st); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); printf("D2D memcpy: %.1f GB/s (r+w)\n", 20 * 2.0 * big / ms / 1e6); cudaFree(d2); }
  // storage
  char path[512];
  for (const char* dir : {diskdir, shmdir}) {
    snprintf(path, sizeof path, "%s/lsio_probe.bin", dir); size_t fsz = 4ull << 30; printf("storage probe in %s\n", dir); make_file(path, fsz);
    for (size_t bs : {256ul << 10, 1ul << 20}) for (int fixed = 0; fixed < 2; fixed++) { double g = disk_probe(path, fsz, bs, 64, fixed, hp, 256ul << 20); printf("  1 thread bs=%zuK qd=64 %s: %.2f GiB/s\n", bs >> 10, fixed ? "READ_FIXED" : "READ", g); }
    // multi-thread
    for (int nt : {4, 8, 16}) { std::vector<std::thread> th; std::vector<double> res(nt); double t1 = now();
      for (int t = 0; t < nt; t++) th.emplace_back([&, t] { res[t] = disk_probe(path, fsz, 1 << 20, 32, true, (char*)hp + (size_t)t * (64ul << 20), 64ul << 20); });
      for (auto& x : th) x.join(); double dt = now() - t1; printf("  %d threads bs=1M qd=32 READ_FIXED (each reads whole file): aggregate %.2f GiB/s\n", nt, nt * (double)fsz / dt / (1 << 30)); }
    unlink(path);
  }
  return 0;
}
