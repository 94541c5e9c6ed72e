// channel.hpp -- bounded MPMC channel with disconnect, standing in for
// crossbeam_channel::bounded(1_024) (crates/lsio_uring/src/io_uring.rs:16): producers block when
// it is full (back-pressure, get_range.rs:70-75), receivers see "disconnected" once every
// sender is gone and the queue is drained (integration_test.rs:92-96).
#pragma once

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>

namespace lsio {

enum class RecvStatus { Ok = 0, Timeout = 1, Disconnected = 2 };

template <class T>
class Channel {
  public:
    explicit Channel(size_t bound) : bound_(bound ? bound : 1) {}
    void add_sender() {
        std::lock_guard<std::mutex> lk(mu_);
        senders_++;
    }
    void remove_sender() {
        std::lock_guard<std::mutex> lk(mu_);
        if (--senders_ == 0) not_empty_.notify_all();
    }
    // Blocks while full. Returns false if there are no receivers left to ever read it (never here).
    void send(T v) {
        std::unique_lock<std::mutex> lk(mu_);
        not_full_.wait(lk, [&] { return q_.size() < bound_; });
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    // Sends n items with as few lock round trips and wake-ups as the bound allows (order kept).
    void send_many(T* items, size_t n) {
        size_t i = 0;
        while (i < n) {
            std::unique_lock<std::mutex> lk(mu_);
            not_full_.wait(lk, [&] { return q_.size() < bound_; });
            while (i < n && q_.size() < bound_) q_.push_back(std::move(items[i++]));
            not_empty_.notify_all();
        }
    }
    RecvStatus recv(T* out, int timeout_ms) {
        std::unique_lock<std::mutex> lk(mu_);
        auto ready = [&] { return !q_.empty() || senders_ == 0; };
        if (timeout_ms < 0) not_empty_.wait(lk, ready);
        else if (!not_empty_.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready)) return RecvStatus::Timeout;
        if (q_.empty()) return RecvStatus::Disconnected;
        *out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return RecvStatus::Ok;
    }
    // First item as recv(); then whatever else is queued, up to max.
    RecvStatus recv_batch(T* out, size_t max, int timeout_ms, size_t* n_out) {
        *n_out = 0;
        if (!max) return RecvStatus::Timeout;
        std::unique_lock<std::mutex> lk(mu_);
        auto ready = [&] { return !q_.empty() || senders_ == 0; };
        if (timeout_ms < 0) not_empty_.wait(lk, ready);
        else if (!not_empty_.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready)) return RecvStatus::Timeout;
        if (q_.empty()) return RecvStatus::Disconnected;
        size_t n = 0;
        while (n < max && !q_.empty()) {
            out[n++] = std::move(q_.front());
            q_.pop_front();
        }
        *n_out = n;
        not_full_.notify_all();
        return RecvStatus::Ok;
    }
    size_t len() {
        std::lock_guard<std::mutex> lk(mu_);
        return q_.size();
    }
    // Drain without blocking (used at teardown to release undelivered chunks).
    bool try_pop(T* out) {
        std::lock_guard<std::mutex> lk(mu_);
        if (q_.empty()) return false;
        *out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }

  private:
    std::mutex mu_;
    std::condition_variable not_empty_, not_full_;
    std::deque<T> q_;
    size_t bound_;
    int senders_ = 0;
};

}  // namespace lsio
