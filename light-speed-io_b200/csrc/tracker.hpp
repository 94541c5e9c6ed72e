// tracker.hpp -- fixed slab of in-flight operations with a FIFO free list.
// Same behaviour as crates/lsio_uring/src/tracker.rs:3-68: indices are handed out 0..n-1 in
// order, a removed index goes to the back of the queue (so a 2-slot tracker yields 0, 1, 0).
#pragma once

#include <cstddef>
#include <deque>
#include <optional>
#include <stdexcept>
#include <utility>
#include <vector>

namespace lsio {

template <class T>
class Tracker {
  public:
    explicit Tracker(size_t n) : ops_(n) {
        for (size_t i = 0; i < n; i++) next_.push_back(i);
    }
    std::optional<size_t> get_next_index() {  // tracker.rs:18-20
        if (next_.empty()) return std::nullopt;
        size_t i = next_.front();
        next_.pop_front();
        return i;
    }
    void put(size_t index, T op) {  // tracker.rs:22-25
        ops_.at(index) = std::move(op);
        len_++;
    }
    T* get(size_t index) {  // tracker.rs:27-36; out of range throws (the reference panics)
        auto& o = ops_.at(index);
        return o ? &*o : nullptr;
    }
    T remove(size_t index) {  // TrackerGuard::remove, tracker.rs:57-61
        auto& o = ops_.at(index);
        if (!o) throw std::logic_error("Tracker::remove on an empty slot");
        T v = std::move(*o);
        o.reset();
        next_.push_back(index);
        len_--;
        return v;
    }
    void replace(size_t index, T op) { ops_.at(index) = std::move(op); }  // tracker.rs:63-65
    bool is_empty() const { return len_ == 0; }
    bool is_full() const { return next_.empty(); }
    size_t len() const { return len_; }
    size_t capacity() const { return ops_.size(); }

  private:
    std::vector<std::optional<T>> ops_;
    std::deque<size_t> next_;
    size_t len_ = 0;
};

}  // namespace lsio
