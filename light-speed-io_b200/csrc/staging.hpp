// staging.hpp -- the pinned staging ring, run as a LOG, and the HBM arena.  Product code.
//
// North star: "lsio_aligned_bytes gains cudaHostRegister'd pinned staging rings and
// device-buffer views".  The reference allocates one fresh 512-aligned heap buffer per read
// (crates/lsio_uring/src/sqe.rs:92, lsio_aligned_bytes lib.rs:178-188: mmap/munmap churn,
// 11 %+5 % of its profile).  Here every read lands in a slot of one pinned,
// io_uring-registered ring; slots are handed out in log order so that the device thread can
// copy whole contiguous runs of landed reads with ONE cudaMemcpyAsync (a 256 KiB copy reaches
// 29 GB/s on this box, a 4 MiB one 52.6 GB/s -- measured).  The device mirror ring
// has identical offsets.
#pragma once

#include <atomic>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace lsio {

enum : uint32_t { kStageReading = 0, kStageLanded = 1, kStageFailed = 2 };

// One read that lands inside an extent.  R = the request type it keeps alive (the backend's Request;
// a dummy in the self test).
template <class R>
struct ReadRefT {
    std::shared_ptr<R> req;
    uint32_t read = 0;   // index into req->reads
    uint32_t off = 0;    // byte offset of the read's buffer inside the extent (512-aligned)
    uint32_t state = kStageReading;  // written by the owning io worker before the extent is published
};

// A staging EXTENT: one contiguous reservation of the (device) log carrying one or more reads.
// staging_mode 0: one read per extent, host memory at the same ring offset.  staging_mode 1: an io worker
// packs consecutive reads into one of its cache-sized pinned slots and copies the extent with one H2D.
template <class R>
struct StageEntryT {
    uint64_t log_off = 0;  // monotonically increasing logical byte offset; phys = log_off % capacity
    uint32_t len = 0;      // bytes reserved
    std::atomic<uint32_t> state{kStageReading};  // kStageLanded once every read has landed/failed (and its copy is enqueued)
    std::vector<ReadRefT<R>> reads;              // owned by the allocating worker until `state` is published
    uint64_t landed_ns = 0;
    void* copy_event = nullptr;  // staging_mode 1: the cudaEvent_t recorded after this extent's H2D copy
    uint32_t owner = 0;          // staging_mode 1: the io worker (copy stream) that issued it
    uint32_t used = 0;           // staging_mode 1: bytes of the extent actually filled (= copied)
    uint32_t nsegs = 0, ntiles = 0;  // pack descriptors / kernel tiles its landed reads need (set before `state` is published)
    bool direct = false;         // its payload was copied straight into the request buffers (the kernel checksums in place)
    bool pulled = false;         // ... by the pull kernel, which also checksummed it: CRC of its j-th segment in Group::pull_crc_host
    uint64_t seq = 0;            // its sequence number in the log
};

template <class R>
class StagingLogT {
  public:
    using StageEntry = StageEntryT<R>;
    static constexpr uint64_t kRegBoundary = 1ull << 30;  // one io_uring registered buffer is <= 1 GiB

    void init(uint64_t capacity, size_t max_entries) {
        cap_ = capacity;
        entries_ = std::vector<StageEntry>(max_entries);
    }
    uint64_t capacity() const { return cap_; }

    // Reserve `len` bytes (multiple of 512) at a multiple of `align` (a power of two dividing the capacity).
    // Thread-safe.  nullptr = ring or entry table full.
    StageEntry* alloc(uint32_t len, uint32_t align = 512) {
        std::lock_guard<std::mutex> lk(mu_);
        if (alloc_seq_ - retire_seq_.load(std::memory_order_acquire) >= entries_.size()) return nullptr;
        uint64_t off = (alloc_off_ + align - 1) / align * align;
        const uint64_t phys = off % cap_;
        const uint64_t boundary = std::min<uint64_t>(cap_, (phys / kRegBoundary + 1) * kRegBoundary);
        if (phys + len > boundary) off += boundary - phys;  // never straddle the ring end or a registration boundary
        if (off + len - free_off_.load(std::memory_order_acquire) > cap_) return nullptr;
        StageEntry& e = entries_[alloc_seq_ % entries_.size()];
        e.log_off = off;
        e.len = len;
        e.state.store(kStageReading, std::memory_order_relaxed);
        e.reads.clear();
        e.used = 0;
        e.nsegs = e.ntiles = 0;
        e.direct = false;
        e.pulled = false;
        e.seq = alloc_seq_;
        e.copy_event = nullptr;
        alloc_off_ = off + len;
        alloc_seq_pub_.store(++alloc_seq_, std::memory_order_release);
        return &e;
    }

    // ---- device thread only ----
    uint64_t alloc_seq() const { return alloc_seq_pub_.load(std::memory_order_acquire); }
    uint64_t retire_seq() const { return retire_seq_.load(std::memory_order_relaxed); }
    StageEntry& at(uint64_t seq) { return entries_[seq % entries_.size()]; }
    // Entries [retire_seq, seq_end) are finished: drop their references and free their bytes.
    void retire_through(uint64_t seq_end) {
        uint64_t s = retire_seq_.load(std::memory_order_relaxed);
        uint64_t free_to = free_off_.load(std::memory_order_relaxed);
        for (; s < seq_end; s++) {
            StageEntry& e = at(s);
            free_to = e.log_off + e.len;
            e.reads.clear();  // drops the requests' references
        }
        free_off_.store(free_to, std::memory_order_release);
        retire_seq_.store(seq_end, std::memory_order_release);
    }
    bool empty() const { return alloc_seq() == retire_seq(); }

  private:
    std::mutex mu_;
    uint64_t cap_ = 0;
    std::vector<StageEntry> entries_;
    uint64_t alloc_off_ = 0, alloc_seq_ = 0;
    std::atomic<uint64_t> alloc_seq_pub_{0}, retire_seq_{0}, free_off_{0};
};

// First-fit arena over one cudaMalloc'd region: the per-request HBM buffers ("contiguous
// per-request HBM buffers", north star) are carved from it and return to it when the last
// chunk view of the request is dropped.
class DeviceArena {
  public:
    static constexpr uint64_t kGrain = 512;
    DeviceArena(int device, uint8_t* base, uint64_t bytes, void (*deleter)(int, void*)) : device_(device), base_(base), cap_(bytes), deleter_(deleter) {
        free_[0] = bytes;
    }
    ~DeviceArena() {
        if (deleter_) deleter_(device_, base_);
    }
    uint64_t capacity() const { return cap_; }
    int device() const { return device_; }
    uint8_t* alloc(uint64_t bytes) {
        bytes = (bytes + kGrain - 1) / kGrain * kGrain;
        std::lock_guard<std::mutex> lk(mu_);
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second >= bytes) {
                const uint64_t off = it->first, len = it->second;
                free_.erase(it);
                if (len > bytes) free_[off + bytes] = len - bytes;
                used_[off] = bytes;
                in_use_ += bytes;
                return base_ + off;
            }
        }
        return nullptr;
    }
    void release(uint8_t* p) {
        std::lock_guard<std::mutex> lk(mu_);
        uint64_t off = (uint64_t)(p - base_);
        auto u = used_.find(off);
        if (u == used_.end()) return;
        uint64_t len = u->second;
        used_.erase(u);
        in_use_ -= len;
        epoch_.fetch_add(1, std::memory_order_release);
        auto next = free_.lower_bound(off);
        if (next != free_.end() && off + len == next->first) {
            len += next->second;
            next = free_.erase(next);
        }
        if (next != free_.begin()) {
            auto prev = std::prev(next);
            if (prev->first + prev->second == off) {
                prev->second += len;
                return;
            }
        }
        free_[off] = len;
    }
    uint64_t in_use() {
        std::lock_guard<std::mutex> lk(mu_);
        return in_use_;
    }
    // Bumped on every release: a request that found the arena full retries only after it moved.
    uint64_t epoch() const { return epoch_.load(std::memory_order_acquire); }

  private:
    int device_;
    uint8_t* base_;
    uint64_t cap_;
    void (*deleter_)(int, void*);
    std::mutex mu_;
    std::map<uint64_t, uint64_t> free_, used_;
    uint64_t in_use_ = 0;
    std::atomic<uint64_t> epoch_{0};
};

}  // namespace lsio
