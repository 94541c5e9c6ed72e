// uring.hpp -- raw io_uring over syscalls (no liburing in the image).  Product code.
// Replaces what the reference gets from the `io-uring` crate (crates/lsio_uring/src/worker.rs:38-42,
// :61, :79, :104-111): ring set-up, SQE push, submit / submit_and_wait, CQ iteration; plus what the
// north star adds: IORING_REGISTER_BUFFERS for the pinned staging ring and READ_FIXED SQEs.
#pragma once

#include <linux/io_uring.h>
#include <linux/time_types.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <sys/uio.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstring>

namespace lsio {

class Ring {
  public:
    Ring() = default;
    Ring(const Ring&) = delete;
    Ring& operator=(const Ring&) = delete;
    ~Ring() { close_ring(); }

    // Returns 0 or -errno.  sqpoll_idle_ms > 0 enables IORING_SETUP_SQPOLL (worker.rs:40).
    int init(unsigned entries, unsigned sqpoll_idle_ms) {
        io_uring_params p;
        memset(&p, 0, sizeof p);
        if (sqpoll_idle_ms) {
            p.flags |= IORING_SETUP_SQPOLL;
            p.sq_thread_idle = sqpoll_idle_ms;
        }
        fd_ = (int)syscall(__NR_io_uring_setup, entries, &p);
        if (fd_ < 0) return -errno;
        flags_ = p.flags;
        sq_entries_ = p.sq_entries;
        cq_entries_ = p.cq_entries;
        sq_sz_ = p.sq_off.array + p.sq_entries * sizeof(unsigned);
        cq_sz_ = p.cq_off.cqes + p.cq_entries * sizeof(io_uring_cqe);
        sq_map_ = mmap(nullptr, sq_sz_, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd_, IORING_OFF_SQ_RING);
        cq_map_ = mmap(nullptr, cq_sz_, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd_, IORING_OFF_CQ_RING);
        sqes_ = (io_uring_sqe*)mmap(nullptr, p.sq_entries * sizeof(io_uring_sqe), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd_,
                                    IORING_OFF_SQES);
        if (sq_map_ == MAP_FAILED || cq_map_ == MAP_FAILED || (void*)sqes_ == MAP_FAILED) {
            int e = errno;
            close_ring();
            return -e;
        }
        char* s = (char*)sq_map_;
        char* c = (char*)cq_map_;
        sq_head_ = (unsigned*)(s + p.sq_off.head);
        sq_tail_ = (unsigned*)(s + p.sq_off.tail);
        sq_mask_ = *(unsigned*)(s + p.sq_off.ring_mask);
        sq_array_ = (unsigned*)(s + p.sq_off.array);
        sq_flags_ = (unsigned*)(s + p.sq_off.flags);
        cq_head_ = (unsigned*)(c + p.cq_off.head);
        cq_tail_ = (unsigned*)(c + p.cq_off.tail);
        cq_mask_ = *(unsigned*)(c + p.cq_off.ring_mask);
        cqes_ = (io_uring_cqe*)(c + p.cq_off.cqes);
        local_tail_ = *sq_tail_;
        for (unsigned i = 0; i < sq_entries_; i++) sq_array_[i] = i;
        return 0;
    }

    unsigned sq_entries() const { return sq_entries_; }
    unsigned cq_entries() const { return cq_entries_; }
    unsigned sq_len() const { return local_tail_ - __atomic_load_n(sq_head_, __ATOMIC_ACQUIRE); }
    unsigned cq_len() const { return __atomic_load_n(cq_tail_, __ATOMIC_ACQUIRE) - *cq_head_; }
    unsigned unsubmitted() const { return unsubmitted_; }

    // Next free SQE (zeroed) or nullptr when the SQ is full (the crate's PushError).
    io_uring_sqe* get_sqe() {
        if (sq_len() >= sq_entries_) return nullptr;
        io_uring_sqe* e = &sqes_[local_tail_ & sq_mask_];
        memset(e, 0, sizeof *e);
        local_tail_++;
        unsubmitted_++;
        return e;
    }

    // submit(): publish the tail and, unless an SQPOLL thread is awake, enter the kernel.
    int submit(unsigned wait_nr = 0, unsigned timeout_us = 0) {
        __atomic_store_n(sq_tail_, local_tail_, __ATOMIC_RELEASE);
        unsigned flags = wait_nr ? IORING_ENTER_GETEVENTS : 0;
        unsigned to_submit = unsubmitted_;
        if (flags_ & IORING_SETUP_SQPOLL) {
            if (__atomic_load_n(sq_flags_, __ATOMIC_ACQUIRE) & IORING_SQ_NEED_WAKEUP) flags |= IORING_ENTER_SQ_WAKEUP;
            else if (!wait_nr) {
                unsubmitted_ = 0;
                return 0;
            }
        } else if (!to_submit && !wait_nr) {
            return 0;
        }
        io_uring_getevents_arg arg;
        __kernel_timespec ts;
        void* argp = nullptr;
        size_t argsz = 0;
        if (wait_nr && timeout_us) {
            memset(&arg, 0, sizeof arg);
            ts.tv_sec = timeout_us / 1000000;
            ts.tv_nsec = (long long)(timeout_us % 1000000) * 1000;
            arg.ts = (uint64_t)(uintptr_t)&ts;
            flags |= IORING_ENTER_EXT_ARG;
            argp = &arg;
            argsz = sizeof arg;
        }
        for (;;) {
            int rc = (int)syscall(__NR_io_uring_enter, fd_, to_submit, wait_nr, flags, argp, argsz);
            if (rc < 0 && errno == EINTR) continue;
            if (rc < 0 && errno == ETIME) {  // the wait timed out; with EXT_ARG the submission part has already happened
                unsubmitted_ = 0;
                return 0;
            }
            if (rc < 0) return -errno;
            // the kernel may consume fewer SQEs than offered: the rest stay counted and go with the next enter
            unsubmitted_ -= (unsigned)rc < to_submit ? (unsigned)rc : to_submit;
            return rc;
        }
    }

    // Pop one CQE; false when the CQ is empty.
    bool pop_cqe(io_uring_cqe* out) {
        unsigned head = *cq_head_;
        if (head == __atomic_load_n(cq_tail_, __ATOMIC_ACQUIRE)) return false;
        *out = cqes_[head & cq_mask_];
        __atomic_store_n(cq_head_, head + 1, __ATOMIC_RELEASE);
        return true;
    }

    int register_buffers(const iovec* iov, unsigned n) {
        int rc = (int)syscall(__NR_io_uring_register, fd_, IORING_REGISTER_BUFFERS, iov, n);
        return rc < 0 ? -errno : 0;
    }

  private:
    void close_ring() {
        if (sqes_ && (void*)sqes_ != MAP_FAILED) munmap(sqes_, sq_entries_ * sizeof(io_uring_sqe));
        if (sq_map_ && sq_map_ != MAP_FAILED) munmap(sq_map_, sq_sz_);
        if (cq_map_ && cq_map_ != MAP_FAILED) munmap(cq_map_, cq_sz_);
        if (fd_ >= 0) close(fd_);
        sqes_ = nullptr; sq_map_ = cq_map_ = nullptr; fd_ = -1;
    }
    int fd_ = -1;
    unsigned flags_ = 0, sq_entries_ = 0, cq_entries_ = 0, sq_mask_ = 0, cq_mask_ = 0;
    unsigned *sq_head_ = nullptr, *sq_tail_ = nullptr, *sq_array_ = nullptr, *sq_flags_ = nullptr;
    unsigned *cq_head_ = nullptr, *cq_tail_ = nullptr;
    io_uring_sqe* sqes_ = nullptr;
    io_uring_cqe* cqes_ = nullptr;
    void *sq_map_ = nullptr, *cq_map_ = nullptr;
    size_t sq_sz_ = 0, cq_sz_ = 0;
    unsigned local_tail_ = 0, unsubmitted_ = 0;
};

// u64 user_data = slot index << 32 | opcode (crates/lsio_uring/src/user_data.rs:28-41).
struct UringUserData {
    uint32_t index_of_op;
    uint8_t opcode;
    static uint64_t pack(uint32_t index_of_op, uint8_t opcode) { return ((uint64_t)index_of_op << 32) | opcode; }
    static UringUserData unpack(uint64_t v) { return UringUserData{(uint32_t)(v >> 32), (uint8_t)(v & 0xFF)}; }
};

}  // namespace lsio
