// threadpool_capi.cpp -- C ABI over threadpool.hpp (include/lsio_threadpool.h).
#include "lsio_threadpool.h"
#include "threadpool.hpp"

struct lsio_threadpool {
    std::unique_ptr<lsio::ThreadPool<uint64_t>> pool;
};

extern "C" {

lsio_threadpool* lsio_threadpool_new(size_t n_worker_threads, lsio_threadpool_task_fn fn, void* ctx) {
    if (!n_worker_threads || !fn) return nullptr;
    auto* p = new lsio_threadpool();
    p->pool = std::make_unique<lsio::ThreadPool<uint64_t>>(n_worker_threads, [fn, ctx](lsio::WorkerThread<uint64_t>& wt) {
        while (wt.keep_running()) {  // threadpool.rs:52-67
            if (auto t = wt.find_task()) {
                const uint64_t follow = fn(ctx, wt.id(), *t);
                if (follow) wt.push(follow);
            } else {
                wt.park();
            }
        }
    });
    return p;
}
void lsio_threadpool_push(lsio_threadpool* self, uint64_t task) {
    if (self) self->pool->push(task);
}
uint64_t lsio_threadpool_n_parks(const lsio_threadpool* self, size_t worker_id) {
    return self && worker_id < self->pool->size() ? self->pool->n_parks(worker_id) : 0;
}
void lsio_threadpool_free(lsio_threadpool* self) {
    if (!self) return;
    self->pool->shutdown();
    delete self;
}

}  // extern "C"
