/*
 * lsio_uring_ref.c -- CPU ORACLE, part 2: restatement of the reference's lsio_uring backend
 * (the timed CPU baseline, cpu_baseline.kind = "port").  TEST INFRASTRUCTURE ONLY: loaded by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never
 * by the product library.
 *
 * The reference is Rust and cannot be built here (no cargo/rustc; un-vendored crates
 * io-uring 0.6.4, crossbeam-deque 0.8.5, crossbeam-channel 0.5.12, nix 0.28), so this file
 * restates it over raw io_uring syscalls (no liburing in the image).  What it follows:
 *   IoUring::new / get_ranges / completion        crates/lsio_uring/src/io_uring.rs:15-49
 *   UringWorker consts, new, run                  crates/lsio_uring/src/worker.rs:14-135
 *   Operation dispatch, maybe_send_error          crates/lsio_uring/src/operation.rs:7-107
 *   GetRanges (openat + statx, fan-out)           crates/lsio_uring/src/get_ranges.rs:15-111
 *   GetRange (one READ per range, emit, close)    crates/lsio_uring/src/get_range.rs:14-92
 *   Close                                         crates/lsio_uring/src/close.rs:12-45
 *   SQE builders                                  crates/lsio_uring/src/sqe.rs:17-133
 *   Tracker / user_data                           tracker.rs:3-68, user_data.rs:6-41
 *   ThreadPool / WorkerThread / ParkManager       crates/lsio_threadpool/src/ (all .rs files)
 *   one fresh aligned allocation per read         crates/lsio_aligned_bytes/src/lib.rs:169-213
 * Deliberate deviations, all documented in DESIGN.md:
 *   - read length is the padded allocation size (multiple of 512) so O_DIRECT on ext4 does
 *     not return EINVAL; the reference passes the un-padded length (sqe.rs:95-102);
 *   - an unaligned start yields the evident intent (view = file[start, end)) instead of the
 *     split_to() panic (sqe.rs:109-113);
 *   - the last-reader close uses an atomic counter instead of Arc::strong_count (racy).
 * mode 1 is the "object_store-like" baseline: the same pool, but each range is a blocking
 * buffered open + pread + close into a fresh allocation (crates/lsio_uring/benches/get.rs:118-123;
 * object_store 0.10.1's source is not in /root/reference, hence "-like").
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <linux/io_uring.h>
#include <linux/stat.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include "lsio_oracle.h"

/* worker.rs:14-22 */
#define MAX_SQ_ENTRIES_PER_ITERATION 2
#define SQ_RING_SIZE 64
#define HIGH_WATER_LINE (SQ_RING_SIZE / 2)
#define OUTPUT_CHANNEL_BOUND 1024 /* io_uring.rs:16 */
#define SQPOLL_IDLE_MS 1000       /* worker.rs:40 */

/* ---- public record -------------------------------------------------------------------- */
typedef struct {
    int32_t status;     /* 0 = Ok(Output::Chunk), <0 = Err(-errno) from a CQE (operation.rs:82-100) */
    uint8_t opcode;     /* io_uring opcode the CQE belonged to */
    uint64_t user_data; /* echoed (lsio_io lib.rs:50-56); 0 for errors of a whole request */
    uint8_t* ptr;       /* view start = base + (start_off - aligned_start) (sqe.rs:116-119) */
    uint64_t len;       /* end_off - start_off */
    void* base;         /* the allocation; free with lsio_ref_chunk_free */
} lsio_ref_output;

/* ---- raw ring ------------------------------------------------------------------------- */
typedef struct {
    int fd;
    unsigned setup_flags, sq_entries, cq_entries;
    unsigned *sq_head, *sq_tail, *sq_mask, *sq_array, *sq_flags;
    unsigned *cq_head, *cq_tail, *cq_mask;
    struct io_uring_sqe* sqes;
    struct io_uring_cqe* cqes;
    unsigned local_tail; /* SQEs written but tail not yet published */
    unsigned unsubmitted; /* published but not yet handed to the kernel (non-SQPOLL) */
    void *sq_map, *cq_map; size_t sq_map_sz, cq_map_sz;
} ring_t;

static int ring_init(ring_t* r, unsigned entries, int sqpoll) {
    struct io_uring_params p;
    memset(&p, 0, sizeof p);
    if (sqpoll) { p.flags |= IORING_SETUP_SQPOLL; p.sq_thread_idle = SQPOLL_IDLE_MS; }
    int fd = (int)syscall(__NR_io_uring_setup, entries, &p);
    if (fd < 0) return -errno;
    memset(r, 0, sizeof *r);
    r->fd = fd; r->setup_flags = p.flags; r->sq_entries = p.sq_entries; r->cq_entries = p.cq_entries;
    r->sq_map_sz = p.sq_off.array + p.sq_entries * sizeof(unsigned);
    r->cq_map_sz = p.cq_off.cqes + p.cq_entries * sizeof(struct io_uring_cqe);
    r->sq_map = mmap(0, r->sq_map_sz, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_SQ_RING);
    r->cq_map = mmap(0, r->cq_map_sz, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_CQ_RING);
    r->sqes = mmap(0, p.sq_entries * sizeof(struct io_uring_sqe), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_SQES);
    if (r->sq_map == MAP_FAILED || r->cq_map == MAP_FAILED || r->sqes == MAP_FAILED) { close(fd); return -ENOMEM; }
    char* s = r->sq_map; char* c = r->cq_map;
    r->sq_head = (unsigned*)(s + p.sq_off.head); r->sq_tail = (unsigned*)(s + p.sq_off.tail);
    r->sq_mask = (unsigned*)(s + p.sq_off.ring_mask); r->sq_array = (unsigned*)(s + p.sq_off.array);
    r->sq_flags = (unsigned*)(s + p.sq_off.flags);
    r->cq_head = (unsigned*)(c + p.cq_off.head); r->cq_tail = (unsigned*)(c + p.cq_off.tail);
    r->cq_mask = (unsigned*)(c + p.cq_off.ring_mask); r->cqes = (struct io_uring_cqe*)(c + p.cq_off.cqes);
    r->local_tail = *r->sq_tail;
    return 0;
}
static void ring_close(ring_t* r) {
    munmap(r->sqes, r->sq_entries * sizeof(struct io_uring_sqe));
    munmap(r->sq_map, r->sq_map_sz); munmap(r->cq_map, r->cq_map_sz); close(r->fd);
}
static unsigned ring_sq_len(ring_t* r) { return r->local_tail - __atomic_load_n(r->sq_head, __ATOMIC_ACQUIRE); }
static unsigned ring_cq_len(ring_t* r) { return __atomic_load_n(r->cq_tail, __ATOMIC_ACQUIRE) - *r->cq_head; }
static int ring_push(ring_t* r, const struct io_uring_sqe* e) {
    if (ring_sq_len(r) >= r->sq_entries) return -1; /* PushError */
    unsigned idx = r->local_tail & *r->sq_mask;
    r->sqes[idx] = *e; r->sq_array[idx] = idx; r->local_tail++; r->unsubmitted++;
    __atomic_store_n(r->sq_tail, r->local_tail, __ATOMIC_RELEASE);
    return 0;
}
static int ring_enter(ring_t* r, unsigned min_complete) {
    unsigned flags = min_complete ? IORING_ENTER_GETEVENTS : 0; unsigned to_submit = r->unsubmitted;
    if (r->setup_flags & IORING_SETUP_SQPOLL) {
        /* io-uring crate Submitter::submit_and_wait: with SQPOLL only enter when the kernel
         * thread sleeps or when we must wait for completions. */
        if (__atomic_load_n(r->sq_flags, __ATOMIC_ACQUIRE) & IORING_SQ_NEED_WAKEUP) flags |= IORING_ENTER_SQ_WAKEUP;
        else if (!min_complete) { r->unsubmitted = 0; return 0; }
    }
    for (;;) {
        int rc = (int)syscall(__NR_io_uring_enter, r->fd, to_submit, min_complete, flags, NULL, 0);
        if (rc < 0 && errno == EINTR) continue;
        if (rc < 0) return -errno;
        r->unsubmitted = 0;
        return rc;
    }
}

/* ---- operations (operation.rs:7-12) ----------------------------------------------------- */
typedef struct open_file {
    char* location; int fd; uint64_t size; uint32_t alignment; /* open_file.rs:4-12 */
    atomic_int refs;                                            /* Arc<OpenFile> */
} open_file;

enum { OP_GETRANGES, OP_GETRANGE, OP_CLOSE, OP_BLOCKING_RANGE };
typedef struct operation {
    int kind;
    /* GetRanges (get_ranges.rs:15-25) */
    char* location; int64_t *starts, *ends; uint64_t* user_datas; size_t n;
    struct statx stx; int fd; int have_fd, have_statx; uint8_t n_cqes_received;
    /* GetRange (get_range.rs:14-19) */
    open_file* file; int64_t start, end; uint64_t user_data; void* base; uint8_t* view; uint64_t view_len;
    struct operation* next; /* queue link */
} operation;

/* ---- queues: crossbeam Injector + per-thread FIFO Worker/Stealer, mutex-backed ---------- */
typedef struct { pthread_mutex_t mu; operation *head, *tail; size_t len; } opqueue;
static void q_init(opqueue* q) { pthread_mutex_init(&q->mu, NULL); q->head = q->tail = NULL; q->len = 0; }
static void q_push(opqueue* q, operation* op) {
    op->next = NULL; pthread_mutex_lock(&q->mu);
    if (q->tail) q->tail->next = op; else q->head = op;
    q->tail = op; q->len++; pthread_mutex_unlock(&q->mu);
}
static operation* q_pop(opqueue* q) { /* FIFO (Worker::new_fifo, threadpool.rs:99) */
    pthread_mutex_lock(&q->mu); operation* op = q->head;
    if (op) { q->head = op->next; if (!q->head) q->tail = NULL; q->len--; }
    pthread_mutex_unlock(&q->mu); return op;
}

/* ---- bounded channel (crossbeam_channel::bounded(1024), io_uring.rs:16) ----------------- */
typedef struct {
    pthread_mutex_t mu; pthread_cond_t not_empty, not_full;
    lsio_ref_output buf[OUTPUT_CHANNEL_BOUND]; size_t head, len; int senders;
} channel;
static void ch_send(channel* c, const lsio_ref_output* o) {
    pthread_mutex_lock(&c->mu);
    while (c->len == OUTPUT_CHANNEL_BOUND) pthread_cond_wait(&c->not_full, &c->mu); /* back-pressure */
    c->buf[(c->head + c->len) % OUTPUT_CHANNEL_BOUND] = *o; c->len++;
    pthread_cond_signal(&c->not_empty); pthread_mutex_unlock(&c->mu);
}

typedef struct lsio_ref_backend lsio_ref_backend;
typedef struct {
    lsio_ref_backend* b; int id; pthread_t th; opqueue local;
    pthread_mutex_t park_mu; pthread_cond_t park_cv; int park_token;
} worker;

struct lsio_ref_backend {
    int n_threads, sqpoll, mode; atomic_int keep_running;
    opqueue injector; worker* workers; channel out;
    /* ParkManager (park_manager.rs:18-71): FIFO of parked threads */
    pthread_mutex_t pm_mu; int* parked; int parked_head, parked_len; atomic_int at_least_one_parked;
};

/* shared_state.rs:26-32 + park_manager.rs:60-70 */
static void unpark_at_most_n(lsio_ref_backend* b, unsigned n) {
    if (!atomic_load_explicit(&b->at_least_one_parked, memory_order_relaxed)) return;
    pthread_mutex_lock(&b->pm_mu);
    while (n-- && b->parked_len) {
        int id = b->parked[b->parked_head]; b->parked_head = (b->parked_head + 1) % b->n_threads; b->parked_len--;
        worker* w = &b->workers[id];
        pthread_mutex_lock(&w->park_mu); w->park_token = 1; pthread_cond_signal(&w->park_cv); pthread_mutex_unlock(&w->park_mu);
    }
    if (!b->parked_len) atomic_store_explicit(&b->at_least_one_parked, 0, memory_order_relaxed);
    pthread_mutex_unlock(&b->pm_mu);
}
/* lsio_threadpool worker.rs:75-86 */
static void worker_park(worker* w) {
    lsio_ref_backend* b = w->b;
    pthread_mutex_lock(&b->pm_mu);
    b->parked[(b->parked_head + b->parked_len) % b->n_threads] = w->id; b->parked_len++;
    atomic_store_explicit(&b->at_least_one_parked, 1, memory_order_relaxed);
    pthread_mutex_unlock(&b->pm_mu);
    pthread_mutex_lock(&w->park_mu);
    if (!w->park_token) { /* like thread::park: may also wake spuriously / on timeout */
        struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts); ts.tv_nsec += 2000000; /* 2 ms safety net for the lost-wakeup window */
        if (ts.tv_nsec >= 1000000000L) { ts.tv_sec++; ts.tv_nsec -= 1000000000L; }
        pthread_cond_timedwait(&w->park_cv, &w->park_mu, &ts);
    }
    w->park_token = 0; pthread_mutex_unlock(&w->park_mu);
    /* remove ourselves if still listed (timed-out wake) */
    pthread_mutex_lock(&b->pm_mu);
    for (int i = 0; i < b->parked_len; i++) {
        int pos = (b->parked_head + i) % b->n_threads;
        if (b->parked[pos] == w->id) {
            for (int j = i; j + 1 < b->parked_len; j++)
                b->parked[(b->parked_head + j) % b->n_threads] = b->parked[(b->parked_head + j + 1) % b->n_threads];
            b->parked_len--; break;
        }
    }
    pthread_mutex_unlock(&b->pm_mu);
}
/* lsio_threadpool worker.rs:45-64: local pop -> injector batch -> steal one from a peer */
static operation* find_task(worker* w) {
    operation* op = q_pop(&w->local); if (op) return op;
    lsio_ref_backend* b = w->b;
    op = q_pop(&b->injector);
    if (op) { /* steal_batch_and_pop: move up to half of the injector (<= 32) to the local queue */
        pthread_mutex_lock(&b->injector.mu); size_t take = b->injector.len / 2; if (take > 32) take = 32; pthread_mutex_unlock(&b->injector.mu);
        while (take--) { operation* more = q_pop(&b->injector); if (!more) break; q_push(&w->local, more); }
        return op;
    }
    for (int i = 0; i < b->n_threads; i++) { if (i == w->id) continue; op = q_pop(&b->workers[i].local); if (op) return op; }
    return NULL;
}
/* lsio_threadpool worker.rs:91-101 */
static void worker_push(worker* w, operation* op) {
    q_push(&w->local, op);
    pthread_mutex_lock(&w->local.mu); size_t n = w->local.len; pthread_mutex_unlock(&w->local.mu);
    if (n > 1) unpark_at_most_n(w->b, (unsigned)n);
}

/* ---- SQE builders (sqe.rs) -------------------------------------------------------------- */
static void sqe_openat(struct io_uring_sqe* e, size_t idx, const char* location) { /* sqe.rs:17-28 */
    memset(e, 0, sizeof *e); e->opcode = IORING_OP_OPENAT; e->fd = -1; e->addr = (uint64_t)(uintptr_t)location;
    e->open_flags = O_RDONLY | O_DIRECT; e->user_data = lsio_oracle_user_data_pack((uint32_t)idx, IORING_OP_OPENAT);
}
static void sqe_statx(struct io_uring_sqe* e, size_t idx, const char* location, struct statx* stx) { /* sqe.rs:40-57 */
    memset(e, 0, sizeof *e); e->opcode = IORING_OP_STATX; e->fd = -1; e->addr = (uint64_t)(uintptr_t)location;
    e->off = (uint64_t)(uintptr_t)stx; e->len = STATX_SIZE | STATX_DIOALIGN; e->user_data = lsio_oracle_user_data_pack((uint32_t)idx, IORING_OP_STATX);
}
static void sqe_close(struct io_uring_sqe* e, size_t idx, int fd) { /* sqe.rs:126-133 */
    memset(e, 0, sizeof *e); e->opcode = IORING_OP_CLOSE; e->fd = fd; e->user_data = lsio_oracle_user_data_pack((uint32_t)idx, IORING_OP_CLOSE);
}
/* sqe.rs:59-122; returns 0 or negative (reference: panic) */
static int sqe_read_range(struct io_uring_sqe* e, size_t idx, operation* op) {
    lsio_oracle_range r;
    int rc = lsio_oracle_resolve_range((int64_t)op->file->size, op->start, op->end, 0, &r);
    if (rc) return rc;
    void* base = NULL;
    if (posix_memalign(&base, LSIO_ORACLE_ALIGN, (size_t)r.padded_len)) return -ENOMEM; /* fresh allocation per read */
    op->base = base; op->view = (uint8_t*)base + r.view_start; op->view_len = (uint64_t)(r.view_end - r.view_start);
    memset(e, 0, sizeof *e); e->opcode = IORING_OP_READ; e->fd = op->file->fd; e->addr = (uint64_t)(uintptr_t)base;
    e->len = (uint32_t)r.padded_len; /* deviation F3: padded length */
    e->off = (uint64_t)r.aligned_start; e->user_data = lsio_oracle_user_data_pack((uint32_t)idx, IORING_OP_READ);
    return 0;
}

typedef struct {
    worker* w; ring_t ring; lsio_oracle_tracker* tracker; operation* slots[SQ_RING_SIZE];
} uring_worker;

static void send_error(uring_worker* uw, operation* op, uint8_t opcode, int res) { /* operation.rs:77-100 */
    lsio_ref_output o; memset(&o, 0, sizeof o); o.status = res; o.opcode = opcode;
    o.user_data = op->kind == OP_GETRANGE ? op->user_data : 0;
    ch_send(&uw->w->b->out, &o);
}
static void op_free(operation* op) {
    if (op->kind == OP_GETRANGES) { free(op->location); free(op->starts); free(op->ends); free(op->user_datas); }
    free(op);
}
static void file_unref_free(open_file* f) { free(f->location); free(f); }

/* get_ranges.rs:39-45 */
static void submit_get_range_ops(uring_worker* uw, operation* g) {
    open_file* f = calloc(1, sizeof *f);
    f->location = g->location; g->location = NULL; f->fd = g->fd; f->size = g->stx.stx_size; f->alignment = g->stx.stx_dio_mem_align;
    atomic_init(&f->refs, (int)g->n);
    if (g->n == 0) { close(f->fd); file_unref_free(f); return; }
    for (size_t i = 0; i < g->n; i++) {
        operation* op = calloc(1, sizeof *op);
        op->kind = OP_GETRANGE; op->file = f; op->start = g->starts[i]; op->end = g->ends[i]; op->user_data = g->user_datas[i];
        worker_push(uw->w, op);
    }
}

enum { NEXT_PENDING, NEXT_DONE, NEXT_REPLACED };

static int submit_first_step(uring_worker* uw, size_t idx, operation* op) {
    struct io_uring_sqe e;
    switch (op->kind) {
    case OP_GETRANGES: /* get_ranges.rs:49-65: two unlinked SQEs */
        sqe_openat(&e, idx, op->location); if (ring_push(&uw->ring, &e)) return -1;
        sqe_statx(&e, idx, op->location, &op->stx); if (ring_push(&uw->ring, &e)) return -1;
        return 0;
    case OP_GETRANGE: { /* get_range.rs:43-51 */
        int rc = sqe_read_range(&e, idx, op); if (rc) return rc;
        return ring_push(&uw->ring, &e); }
    case OP_CLOSE: /* close.rs:24-31 */
        sqe_close(&e, idx, op->file->fd); return ring_push(&uw->ring, &e);
    }
    return -1;
}

static int process_cqe(uring_worker* uw, size_t idx, operation* op, uint8_t opcode, int res) {
    if (res < 0) send_error(uw, op, opcode, res); /* operation.rs:47 maybe_send_error first */
    switch (op->kind) {
    case OP_GETRANGES: /* get_ranges.rs:67-110 */
        op->n_cqes_received++;
        if (res >= 0) {
            if (opcode == IORING_OP_OPENAT) { op->fd = res; op->have_fd = 1; }
            else if (opcode == IORING_OP_STATX) op->have_statx = 1;
        }
        if (op->n_cqes_received == 2) {
            if (op->have_fd && op->have_statx) submit_get_range_ops(uw, op);
            else if (op->have_fd) close(op->fd); /* reference leaks the fd here (Appendix A) */
            return NEXT_DONE;
        }
        return NEXT_PENDING;
    case OP_GETRANGE: { /* get_range.rs:53-91 */
        if (res >= 0) { /* no short-read check in the reference (get_range.rs:66-68) */
            lsio_ref_output o; memset(&o, 0, sizeof o); o.status = 0; o.opcode = opcode; o.user_data = op->user_data;
            o.ptr = op->view; o.len = op->view_len; o.base = op->base; op->base = NULL;
            ch_send(&uw->w->b->out, &o);
        } else if (op->base) { free(op->base); op->base = NULL; }
        if (atomic_fetch_sub(&op->file->refs, 1) == 1) { /* last reader closes (get_range.rs:78-87) */
            op->kind = OP_CLOSE; /* NextStep::ReplaceWith(Close) reusing this slot */
            if (submit_first_step(uw, idx, op) == 0) return NEXT_REPLACED;
            close(op->file->fd); file_unref_free(op->file); return NEXT_DONE;
        }
        return NEXT_DONE; }
    case OP_CLOSE: /* close.rs:33-44 */
        file_unref_free(op->file); return NEXT_DONE;
    }
    return NEXT_DONE;
}

/* mode 1: blocking "object_store-like" range read (benches/get.rs:118-123) */
static void blocking_range(worker* w, operation* op) {
    lsio_ref_output o; memset(&o, 0, sizeof o); o.user_data = op->user_data; o.opcode = IORING_OP_READ;
    int fd = open(op->location, O_RDONLY);
    if (fd < 0) { o.status = -errno; ch_send(&w->b->out, &o); return; }
    struct stat st; lsio_oracle_range r;
    if (fstat(fd, &st) < 0 || lsio_oracle_resolve_range(st.st_size, op->start, op->end, 0, &r)) { o.status = -EINVAL; close(fd); ch_send(&w->b->out, &o); return; }
    size_t want = (size_t)(r.end_off - r.start_off); uint8_t* buf = malloc(want ? want : 1); size_t done = 0;
    while (done < want) { ssize_t n = pread(fd, buf + done, want - done, r.start_off + (int64_t)done); if (n <= 0) { if (n < 0 && errno == EINTR) continue; break; } done += (size_t)n; }
    close(fd);
    if (done != want) { o.status = -EIO; free(buf); } else { o.ptr = buf; o.len = want; o.base = buf; }
    ch_send(&w->b->out, &o);
}
static void run_blocking(worker* w) {
    lsio_ref_backend* b = w->b;
    while (atomic_load_explicit(&b->keep_running, memory_order_relaxed)) {
        operation* op = find_task(w);
        if (!op) { worker_park(w); continue; }
        if (op->kind == OP_GETRANGES) {
            for (size_t i = 0; i < op->n; i++) {
                operation* r = calloc(1, sizeof *r); r->kind = OP_BLOCKING_RANGE; r->location = strdup(op->location);
                r->start = op->starts[i]; r->end = op->ends[i]; r->user_data = op->user_datas[i]; worker_push(w, r);
            }
            op_free(op);
        } else { blocking_range(w, op); free(op->location); free(op); }
    }
}

/* worker.rs:55-125 */
static void* worker_main(void* arg) {
    worker* w = arg; lsio_ref_backend* b = w->b;
    if (b->mode == 1) { run_blocking(w); goto out; }
    uring_worker uw; memset(&uw, 0, sizeof uw); uw.w = w;
    if (ring_init(&uw.ring, SQ_RING_SIZE, b->sqpoll) != 0) { fprintf(stderr, "lsio_ref: Failed to initialise io_uring.\n"); abort(); }
    if (uw.ring.cq_entries != uw.ring.sq_entries * 2) abort(); /* worker.rs:44 */
    uw.tracker = lsio_oracle_tracker_new(SQ_RING_SIZE);
    for (;;) {
        int running = atomic_load_explicit(&b->keep_running, memory_order_relaxed);
        if (!running && lsio_oracle_tracker_is_empty(uw.tracker)) break;
        unsigned sqcq = ring_sq_len(&uw.ring) + ring_cq_len(&uw.ring);
        if (!running || lsio_oracle_tracker_is_full(uw.tracker) || sqcq >= SQ_RING_SIZE - MAX_SQ_ENTRIES_PER_ITERATION) {
            if (ring_cq_len(&uw.ring) == 0) ring_enter(&uw.ring, 1); /* submit_and_wait(1), worker.rs:58-62 */
        } else {
            operation* op = find_task(w);
            if (op) {
                int64_t idx = lsio_oracle_tracker_get_next_index(uw.tracker);
                int rc = submit_first_step(&uw, (size_t)idx, op);
                if (rc) { /* reference: expect()/assert panic; here: in-band error */
                    send_error(&uw, op, IORING_OP_READ, rc < -1 ? -EINVAL : -EBUSY);
                    lsio_oracle_tracker_put(uw.tracker, (size_t)idx, 0); uint64_t d; lsio_oracle_tracker_remove(uw.tracker, (size_t)idx, &d);
                    if (op->kind == OP_GETRANGE && atomic_fetch_sub(&op->file->refs, 1) == 1) { close(op->file->fd); file_unref_free(op->file); }
                    op_free(op); continue;
                }
                ring_enter(&uw.ring, 0); /* submit() on every op, worker.rs:79 */
                lsio_oracle_tracker_put(uw.tracker, (size_t)idx, (uint64_t)(uintptr_t)op); uw.slots[idx] = op;
                if (ring_sq_len(&uw.ring) + ring_cq_len(&uw.ring) < HIGH_WATER_LINE) continue; /* top up, worker.rs:81-88 */
            } else if (lsio_oracle_tracker_is_empty(uw.tracker)) {
                worker_park(w); continue; /* worker.rs:93-99 */
            } /* else fall through: busy-poll the CQ */
        }
        /* drain CQ, worker.rs:104-122 */
        unsigned head = *uw.ring.cq_head, tail = __atomic_load_n(uw.ring.cq_tail, __ATOMIC_ACQUIRE);
        while (head != tail) {
            struct io_uring_cqe cqe = uw.ring.cqes[head & *uw.ring.cq_mask]; head++;
            __atomic_store_n(uw.ring.cq_head, head, __ATOMIC_RELEASE);
            uint32_t idx; uint8_t opcode; lsio_oracle_user_data_unpack(cqe.user_data, &idx, &opcode);
            operation* op = uw.slots[idx];
            int next = process_cqe(&uw, idx, op, opcode, cqe.res);
            if (next == NEXT_DONE) { uint64_t d; lsio_oracle_tracker_remove(uw.tracker, idx, &d); uw.slots[idx] = NULL; op_free(op); }
            else if (next == NEXT_REPLACED) ring_enter(&uw.ring, 0);
        }
    }
    lsio_oracle_tracker_free(uw.tracker); ring_close(&uw.ring);
out:
    pthread_mutex_lock(&b->out.mu); b->out.senders--; pthread_cond_broadcast(&b->out.not_empty); pthread_mutex_unlock(&b->out.mu);
    return NULL;
}

/* ---- public API (io_uring.rs:15-49; lsio_io lib.rs:11-45) -------------------------------- */
lsio_ref_backend* lsio_ref_new(int n_worker_threads, int sqpoll, int mode) {
    lsio_ref_backend* b = calloc(1, sizeof *b);
    b->n_threads = n_worker_threads; b->sqpoll = sqpoll; b->mode = mode; atomic_init(&b->keep_running, 1);
    q_init(&b->injector); pthread_mutex_init(&b->pm_mu, NULL); b->parked = calloc((size_t)n_worker_threads, sizeof(int));
    pthread_mutex_init(&b->out.mu, NULL); pthread_cond_init(&b->out.not_empty, NULL); pthread_cond_init(&b->out.not_full, NULL);
    b->out.senders = n_worker_threads;
    b->workers = calloc((size_t)n_worker_threads, sizeof(worker));
    for (int i = 0; i < n_worker_threads; i++) {
        worker* w = &b->workers[i]; w->b = b; w->id = i; q_init(&w->local);
        pthread_mutex_init(&w->park_mu, NULL); pthread_cond_init(&w->park_cv, NULL);
    }
    for (int i = 0; i < n_worker_threads; i++) pthread_create(&b->workers[i].th, NULL, worker_main, &b->workers[i]);
    return b;
}
/* Returns 0; -EINVAL when the reference would panic (length mismatch is impossible here: one n). */
int lsio_ref_get_ranges(lsio_ref_backend* b, const char* location, const int64_t* starts, const int64_t* ends,
                        const uint64_t* user_data, size_t n) {
    operation* op = calloc(1, sizeof *op);
    op->kind = OP_GETRANGES; op->location = strdup(location); op->n = n;
    op->starts = malloc(n * sizeof(int64_t) + 1); op->ends = malloc(n * sizeof(int64_t) + 1); op->user_datas = malloc(n * sizeof(uint64_t) + 1);
    memcpy(op->starts, starts, n * sizeof(int64_t)); memcpy(op->ends, ends, n * sizeof(int64_t)); memcpy(op->user_datas, user_data, n * sizeof(uint64_t));
    q_push(&b->injector, op); unpark_at_most_n(b, 1); /* threadpool.rs:134-137 */
    return 0;
}
/* 0 = got one, 1 = timeout, 2 = disconnected (all senders gone and queue empty) */
int lsio_ref_recv(lsio_ref_backend* b, lsio_ref_output* out, int timeout_ms) {
    channel* c = &b->out; struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_sec += timeout_ms / 1000; ts.tv_nsec += (long)(timeout_ms % 1000) * 1000000L;
    if (ts.tv_nsec >= 1000000000L) { ts.tv_sec++; ts.tv_nsec -= 1000000000L; }
    pthread_mutex_lock(&c->mu);
    while (c->len == 0) {
        if (c->senders == 0) { pthread_mutex_unlock(&c->mu); return 2; }
        if (timeout_ms < 0) pthread_cond_wait(&c->not_empty, &c->mu);
        else if (pthread_cond_timedwait(&c->not_empty, &c->mu, &ts) == ETIMEDOUT && c->len == 0) {
            int gone = c->senders == 0; pthread_mutex_unlock(&c->mu); return gone ? 2 : 1;
        }
    }
    *out = c->buf[c->head]; c->head = (c->head + 1) % OUTPUT_CHANNEL_BOUND; c->len--;
    pthread_cond_signal(&c->not_full); pthread_mutex_unlock(&c->mu);
    return 0;
}
void lsio_ref_chunk_free(void* base) { free(base); } /* drop of the last Arc<InnerBuffer>, lib.rs:209-213 */

/* Drop for ThreadPool (threadpool.rs:140-165): stop, unpark, join.  The handle stays valid for
 * lsio_ref_recv (a cloned Receiver, integration_test.rs:92-96) until lsio_ref_free. */
void lsio_ref_shutdown(lsio_ref_backend* b) {
    if (!atomic_exchange(&b->keep_running, 0)) return;
    for (int i = 0; i < b->n_threads; i++) {
        worker* w = &b->workers[i];
        pthread_mutex_lock(&w->park_mu); w->park_token = 1; pthread_cond_signal(&w->park_cv); pthread_mutex_unlock(&w->park_mu);
    }
    for (int i = 0; i < b->n_threads; i++) pthread_join(b->workers[i].th, NULL);
}
void lsio_ref_free(lsio_ref_backend* b) {
    lsio_ref_shutdown(b);
    for (operation* op; (op = q_pop(&b->injector));) op_free(op);
    while (b->out.len) { lsio_ref_output* o = &b->out.buf[b->out.head]; if (o->base) free(o->base); b->out.head = (b->out.head + 1) % OUTPUT_CHANNEL_BOUND; b->out.len--; }
    free(b->workers); free(b->parked); free(b);
}

/* ---- bench driver in C so the timed baseline has no Python in its loop --------------------
 * Restates lsio_bench::read_files (crates/lsio_bench/src/main.rs:124-190): submit one
 * get_ranges per file, then drain n_files*n_chunks completions (drop each chunk on receipt),
 * wall-clock from before the first submit to the last completion.  Returns seconds, or <0. */
double lsio_ref_bench_read_files(const char* const* paths, size_t n_files, const int64_t* starts, const int64_t* ends,
                                 size_t n_chunks, int n_worker_threads, int sqpoll, int mode, uint64_t* bytes_out, uint64_t* errors_out) {
    uint64_t* ud = malloc(n_chunks * sizeof(uint64_t) + 1);
    lsio_ref_backend* b = lsio_ref_new(n_worker_threads, sqpoll, mode);
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    uint64_t bytes = 0, errors = 0, got = 0, total = (uint64_t)n_files * n_chunks;
    size_t submitted = 0;
    /* lsio_bench submits everything first (the injector is unbounded); same here. */
    for (; submitted < n_files; submitted++) {
        for (size_t i = 0; i < n_chunks; i++) ud[i] = ((uint64_t)submitted << 32) | i;
        lsio_ref_get_ranges(b, paths[submitted], starts, ends, ud, n_chunks);
    }
    while (got < total) {
        lsio_ref_output o; int rc = lsio_ref_recv(b, &o, 10000); /* main.rs:171 10 s timeout */
        if (rc != 0) { errors += total - got; break; }
        if (o.status == 0) { bytes += o.len; got++; lsio_ref_chunk_free(o.base); }
        else { errors++; if (o.opcode == IORING_OP_READ) got++; else { got = total; } }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    lsio_ref_free(b); free(ud);
    if (bytes_out) *bytes_out = bytes;
    if (errors_out) *errors_out = errors;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Same driver for arbitrary per-file range lists (BASELINE configs C3-C5): file i owns
 * counts[i] consecutive entries of starts[]/ends[]. */
double lsio_ref_bench_read_ranges(const char* const* paths, size_t n_files, const int64_t* starts, const int64_t* ends,
                                  const uint64_t* counts, int n_worker_threads, int sqpoll, int mode, uint64_t* bytes_out,
                                  uint64_t* errors_out) {
    size_t max_n = 0; uint64_t total = 0;
    for (size_t i = 0; i < n_files; i++) { if (counts[i] > max_n) max_n = counts[i]; total += counts[i]; }
    uint64_t* ud = malloc(max_n * sizeof(uint64_t) + 8);
    lsio_ref_backend* b = lsio_ref_new(n_worker_threads, sqpoll, mode);
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    uint64_t bytes = 0, errors = 0, got = 0, off = 0;
    for (size_t i = 0; i < n_files; i++) {
        for (size_t k = 0; k < counts[i]; k++) ud[k] = off + k;
        lsio_ref_get_ranges(b, paths[i], starts + off, ends + off, ud, counts[i]);
        off += counts[i];
    }
    while (got < total) {
        lsio_ref_output o; int rc = lsio_ref_recv(b, &o, 20000);
        if (rc != 0) { errors += total - got; break; }
        if (o.status == 0) { bytes += o.len; got++; lsio_ref_chunk_free(o.base); }
        else { errors++; if (o.opcode == IORING_OP_READ) got++; else got = total; }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    lsio_ref_free(b); free(ud);
    if (bytes_out) *bytes_out = bytes;
    if (errors_out) *errors_out = errors;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
