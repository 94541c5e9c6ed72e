#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include "lsio_threadpool.h"
#include "lsio_uring_cuda.h"
static std::atomic<int> seen{0};
static uint64_t on_task(void*, size_t, uint64_t task) { seen++; return task > 1 && task < 100000 ? task - 1 : 0; }
int main() {
    for (int c = 1; c <= 7; c++) { int rc = lsio_uring_cuda_selftest(c); printf("selftest %d -> %d\n", c, rc); if (rc) return 1; }
    for (int round = 0; round < 3; round++) {
        seen = 0;
        lsio_threadpool* p = lsio_threadpool_new(4, on_task, nullptr);
        for (int i = 0; i < 8; i++) { lsio_threadpool_push(p, 500); if (i % 3 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(5)); }
        while (seen.load() < 4000) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        lsio_threadpool_free(p);
        printf("pool round %d: %d tasks\n", round, seen.load());
    }
    return 0;
}
