// op_groups.hpp -- ordered groups of operations: "all operations in group n are completed before any IO
// operations in group n+1 are started" (planned in the reference, not built there: crates/lsio_io/src/lib.rs:6-8,
// planned_design.md:71-72).  Pure bookkeeping, no CUDA and no I/O, so it is unit-tested without a GPU
// (csrc/selftest.cpp case 7).  Item = whatever a submission is (the backend: device group + Op).
#pragma once

#include <cstdint>
#include <deque>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace lsio {

template <class Item>
class OrderedGroups {
  public:
    // Registers one submission in the current group.  Returns true if it may start now (the caller keeps `item`
    // and starts it); false if an earlier group is still unfinished: `item` has been moved into the hold queue and
    // will come back from finished() later.  *group receives the group it belongs to.
    bool submit(Item& item, uint64_t* group) {
        std::lock_guard<std::mutex> lk(mu_);
        *group = current_;
        alive_[current_]++;
        if (current_ > running_) {
            held_.emplace_back(current_, std::move(item));
            return false;
        }
        return true;
    }
    // One submission of `group` has finished for good (every item delivered).  Appends to `out`, in submission
    // order, the held submissions that may start now.  `release == false` (teardown) only does the accounting.
    void finished(uint64_t group, bool release, std::vector<Item>* out) {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = alive_.find(group);
        if (it == alive_.end() || --it->second != 0) return;
        alive_.erase(it);
        if (release) admit(out);
    }
    // Closes the current group and opens the next; returns its id.
    uint64_t next_group() {
        std::lock_guard<std::mutex> lk(mu_);
        current_++;
        if (alive_.empty()) running_ = current_;  // nothing in flight: the new group may start at once
        return current_;
    }
    uint64_t current() {
        std::lock_guard<std::mutex> lk(mu_);
        return current_;
    }
    // Teardown: the submissions that were never admitted, to be dropped unstarted by the caller (outside the lock).
    std::deque<std::pair<uint64_t, Item>> take_held() {
        std::lock_guard<std::mutex> lk(mu_);
        std::deque<std::pair<uint64_t, Item>> h;
        h.swap(held_);
        return h;
    }
    size_t held_count() {
        std::lock_guard<std::mutex> lk(mu_);
        return held_.size();
    }

  private:
    void admit(std::vector<Item>* out) {
        while (true) {
            // running = the oldest group with unfinished submissions, else the current group
            running_ = alive_.empty() ? current_ : std::min(alive_.begin()->first, current_);
            bool any = false;
            while (!held_.empty() && held_.front().first <= running_) {
                out->push_back(std::move(held_.front().second));
                held_.pop_front();
                any = true;
            }
            if (!any) break;
        }
    }
    std::mutex mu_;
    uint64_t current_ = 0;                 // group that new submissions join
    uint64_t running_ = 0;                 // newest group that may issue I/O
    std::map<uint64_t, uint64_t> alive_;   // unfinished submissions per group
    std::deque<std::pair<uint64_t, Item>> held_;  // submissions of groups > running_, in submission order
};

}  // namespace lsio
