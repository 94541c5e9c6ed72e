/*
 * lsio_threadpool.h -- C ABI of the work-stealing pool (csrc/threadpool.hpp), the B200 build's
 * `lsio_threadpool` (crates/lsio_threadpool/src/threadpool.rs:20-165, worker.rs:17-102).
 * The backend uses the C++ template directly; this binding exists so the reference's own
 * pool test (threadpool.rs:186-268: exactly-once delivery, every thread works, park counts,
 * shutdown on drop) can be restated against the shipped code.
 *
 * Each worker runs the reference's canonical loop (threadpool.rs:46-76):
 *     while keep_running { match find_task() { Some(t) => fn(ctx, worker, t), None => park() } }
 */
#ifndef LSIO_THREADPOOL_H
#define LSIO_THREADPOOL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsio_threadpool lsio_threadpool;
/* Called on a worker thread for every task. Return a non-zero task id to push a follow-up
 * task onto THIS worker's local deque (WorkerThread::push, worker.rs:91-94), or 0 for none. */
typedef uint64_t (*lsio_threadpool_task_fn)(void* ctx, size_t worker_id, uint64_t task);

lsio_threadpool* lsio_threadpool_new(size_t n_worker_threads, lsio_threadpool_task_fn fn, void* ctx); /* ThreadPool::new, threadpool.rs:78 */
void lsio_threadpool_push(lsio_threadpool* self, uint64_t task);                                      /* ThreadPool::push, threadpool.rs:134 */
uint64_t lsio_threadpool_n_parks(const lsio_threadpool* self, size_t worker_id);
void lsio_threadpool_free(lsio_threadpool* self);                                                     /* Drop, threadpool.rs:140-165 */

#ifdef __cplusplus
}
#endif
#endif
