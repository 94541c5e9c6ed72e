// selftest.cpp -- host-logic self tests of the shipped C++ components, callable without a GPU
// (lsio_uring_cuda_selftest in include/lsio_uring_cuda.h).  Each case restates a reference unit
// test against the PRODUCT code (not the oracle):
//   1 tracker    crates/lsio_uring/src/tracker.rs:74-114
//   2 user_data  crates/lsio_uring/src/user_data.rs:47-56
//   3 channel    bounded(…) back-pressure + disconnect (io_uring.rs:16, integration_test.rs:92-96)
//   4 staging    the pinned ring run as a log: contiguity, wrap, registration boundary, in-order retire
//   5 arena      first-fit + coalescing of the per-request HBM buffers
//   6 gf2        crc32c_combine against a direct bitwise CRC
//   7 groups     ordered groups of operations (planned: lsio_io/src/lib.rs:6-8): admission order, held submissions,
//                empty groups, teardown, and a multi-threaded run checked for "group n finishes before n+1 starts"
// Returns 0 on success, or (case * 100 + failed check).
#include <linux/io_uring.h>

#include <atomic>
#include <map>
#include <mutex>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

#include "channel.hpp"
#include "gf2.hpp"
#include "lsio_uring_cuda.h"
#include "op_groups.hpp"
#include "staging.hpp"
#include "tracker.hpp"
#include "uring.hpp"

using namespace lsio;
namespace {
struct DummyRequest {};
using Request = DummyRequest;
using StagingLog = StagingLogT<DummyRequest>;
using StageEntry = StageEntryT<DummyRequest>;
}  // namespace

#define CHECK(n, cond) \
    do {               \
        if (!(cond)) return which * 100 + (n); \
    } while (0)

static uint32_t crc_bitwise(const uint8_t* p, size_t n) {
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (gf2::kPoly & (0u - (c & 1u)));
    }
    return ~c;
}

extern "C" int lsio_uring_cuda_selftest(int which) {
    switch (which) {
    case 1: {  // tracker.rs::test_op_tracker + test_panic_if_wrong_index
        Tracker<std::string> t(2);
        CHECK(1, t.get(0) == nullptr);
        auto i0 = t.get_next_index();
        CHECK(2, i0 && *i0 == 0);
        t.put(*i0, "string0");
        CHECK(3, t.remove(*i0) == "string0");
        auto i1 = t.get_next_index();
        CHECK(4, i1 && *i1 == 1);
        t.put(*i1, "string1");
        auto i2 = t.get_next_index();
        CHECK(5, i2 && *i2 == 0);  // FIFO free list
        t.put(*i2, "string2");
        CHECK(6, !t.get_next_index() && t.is_full());
        CHECK(7, t.remove(*i1) == "string1" && t.remove(*i2) == "string2" && t.is_empty());
        bool threw = false;
        try {
            t.get(100);
        } catch (const std::out_of_range&) {
            threw = true;  // "index out of bounds"
        }
        CHECK(8, threw);
        t.put(0, "a");
        t.replace(0, "b");  // NextStep::ReplaceWith keeps the slot
        CHECK(9, *t.get(0) == "b" && t.len() == 1);
        return 0;
    }
    case 2: {  // user_data.rs::test_uring_user_data_round_trip
        const uint64_t v = UringUserData::pack(100, IORING_OP_READ);
        CHECK(1, v == ((100ull << 32) | IORING_OP_READ));
        const UringUserData u = UringUserData::unpack(v);
        CHECK(2, u.index_of_op == 100 && u.opcode == IORING_OP_READ);
        CHECK(3, UringUserData::unpack(0xFFFFFFFF000000FFull).index_of_op == 0xFFFFFFFFu);
        return 0;
    }
    case 3: {
        Channel<int> ch(4);
        ch.add_sender();
        std::atomic<int> sent{0};
        std::thread prod([&] {
            for (int i = 0; i < 20; i++) {
                ch.send(i);
                sent++;
            }
            int more[7] = {20, 21, 22, 23, 24, 25, 26};
            ch.send_many(more, 7);
            sent += 7;
            ch.remove_sender();
        });
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
        CHECK(1, sent.load() == 4 && ch.len() == 4);  // producers block at the bound (back-pressure)
        int v = -1, expect = 0;
        while (ch.recv(&v, 1000) == RecvStatus::Ok) CHECK(2, v == expect++);
        prod.join();
        CHECK(3, expect == 27);
        CHECK(4, ch.recv(&v, 10) == RecvStatus::Disconnected);  // all senders gone and drained
        Channel<int> idle(2);
        idle.add_sender();
        CHECK(5, idle.recv(&v, 10) == RecvStatus::Timeout);
        int many[3];
        size_t n = 9;
        CHECK(6, idle.recv_batch(many, 3, 10, &n) == RecvStatus::Timeout && n == 0);
        idle.send(7);
        idle.send(8);
        CHECK(7, idle.recv_batch(many, 3, 10, &n) == RecvStatus::Ok && n == 2 && many[0] == 7 && many[1] == 8);
        return 0;
    }
    case 4: {
        StagingLog log;
        const uint64_t cap = 1 << 20;
        log.init(cap, 8);
        auto req = std::make_shared<Request>();
        std::vector<StageEntry*> e;
        auto alloc = [&](StagingLog& l, uint32_t len, uint32_t idx) {
            StageEntry* x = l.alloc(len);
            if (x) x->reads.push_back(ReadRefT<DummyRequest>{req, idx, 0, kStageReading});
            return x;
        };
        for (int i = 0; i < 3; i++) e.push_back(alloc(log, 256 << 10, (uint32_t)i));
        CHECK(1, e[0] && e[1] && e[2] && e[0]->log_off == 0 && e[1]->log_off == (256 << 10) && e[2]->log_off == (512 << 10));
        CHECK(2, alloc(log, 512 << 10, 3) == nullptr);  // only 256 KiB left before the wrap and nothing is free yet
        StageEntry* last = alloc(log, 256 << 10, 3);
        CHECK(3, last && last->log_off == (768 << 10));
        CHECK(4, alloc(log, 512, 4) == nullptr);  // ring full
        CHECK(5, log.alloc_seq() == 4 && log.retire_seq() == 0 && req.use_count() == 5);
        log.retire_through(2);  // the device thread retires in order
        CHECK(6, req.use_count() == 3);
        StageEntry* w = alloc(log, 384 << 10, 5);
        CHECK(7, w && w->log_off == cap && w->log_off % cap == 0);  // wrapped: logical offsets keep growing
        CHECK(8, alloc(log, 256 << 10, 6) == nullptr);  // [384K, 512K) is free but too small
        log.retire_through(4);
        StageEntry* x = alloc(log, 512 << 10, 7);
        CHECK(9, x && x->log_off == cap + (384 << 10));
        // an allocation never straddles the wrap: it skips to the start of the ring
        log.retire_through(6);
        StageEntry* y = alloc(log, 256 << 10, 8);
        CHECK(10, y && y->log_off % cap == 0 && y->log_off == 2 * cap);
        // the entry table bounds the number of live slots too
        log.retire_through(7);
        StagingLog small;
        small.init(1 << 20, 2);
        CHECK(11, alloc(small, 512, 0) && alloc(small, 512, 1) && alloc(small, 512, 2) == nullptr);
        return 0;
    }
    case 5: {
        uint8_t* base = (uint8_t*)0x100000000ull;  // never dereferenced
        DeviceArena a(0, base, 1 << 20, nullptr);
        uint8_t* p1 = a.alloc(1000);  // rounded to the 512 B grain
        uint8_t* p2 = a.alloc(512 << 10);
        uint8_t* p3 = a.alloc(256 << 10);
        CHECK(1, p1 == base && p2 == base + 1024 && p3 == base + 1024 + (512 << 10));
        CHECK(2, a.alloc(512 << 10) == nullptr && a.in_use() == 1024 + (768 << 10));
        a.release(p2);
        uint8_t* p4 = a.alloc(300 << 10);
        CHECK(3, p4 == p2);  // first fit re-uses the hole
        a.release(p1);
        a.release(p4);       // coalesces with both neighbours' free space
        uint8_t* p5 = a.alloc(513 << 10);
        CHECK(4, p5 == base);
        a.release(p3);
        a.release(p5);
        CHECK(5, a.in_use() == 0 && a.alloc(1 << 20) == base);
        return 0;
    }
    case 6: {
        std::vector<uint8_t> buf(70000);
        uint32_t x = 12345;
        for (auto& b : buf) {
            x = x * 1664525u + 1013904223u;
            b = (uint8_t)(x >> 24);
        }
        const size_t cuts[] = {0, 1, 4, 511, 512, 4096, 65536, 69999, 70000};
        const uint32_t whole = crc_bitwise(buf.data(), buf.size());
        for (size_t c : cuts) {
            const uint32_t a = crc_bitwise(buf.data(), c), b = crc_bitwise(buf.data() + c, buf.size() - c);
            CHECK(1, gf2::crc32c_combine(a, b, buf.size() - c) == whole);
        }
        CHECK(2, crc_bitwise((const uint8_t*)"123456789", 9) == 0xE3069283u);
        CHECK(3, gf2::multmodp(gf2::xpow(100), gf2::xpow(-100)) == 0x80000000u);  // x^100 * x^-100 == 1
        return 0;
    }
    case 7: {
        OrderedGroups<int> og;
        uint64_t g = 99;
        int a = 1, b = 2, c = 3, d = 4, e = 5;
        std::vector<int> out;
        CHECK(1, og.current() == 0 && og.submit(a, &g) && g == 0);  // group 0 starts at once
        CHECK(2, og.submit(b, &g) && g == 0);
        CHECK(3, og.next_group() == 1);
        CHECK(4, !og.submit(c, &g) && g == 1 && og.held_count() == 1);  // group 1 waits for group 0
        CHECK(5, og.next_group() == 2 && og.next_group() == 3);        // group 2 stays empty
        CHECK(6, !og.submit(d, &g) && g == 3 && !og.submit(e, &g) && og.held_count() == 3);
        og.finished(0, true, &out);
        CHECK(7, out.empty());  // one submission of group 0 is still running
        og.finished(0, true, &out);
        CHECK(8, out.size() == 1 && out[0] == 3);  // group 1 admitted, group 3 not yet
        out.clear();
        og.finished(1, true, &out);
        CHECK(9, out.size() == 2 && out[0] == 4 && out[1] == 5);  // the empty group 2 is skipped, order kept
        out.clear();
        og.finished(3, true, &out);
        og.finished(3, true, &out);
        int f = 6;
        CHECK(10, out.empty() && og.submit(f, &g) && g == 3);  // nothing in flight: the current group runs at once
        CHECK(11, og.next_group() == 4);
        int h = 7;
        CHECK(12, !og.submit(h, &g) && g == 4);
        og.finished(3, false, &out);  // teardown: accounting only, nothing is released
        CHECK(13, out.empty() && og.held_count() == 1 && og.take_held().size() == 1 && og.held_count() == 0);
        // next_group with nothing in flight admits the new group immediately
        OrderedGroups<int> og2;
        CHECK(14, og2.next_group() == 1 && og2.submit(a, &g) && g == 1);
        // concurrent: 4 submitters x 200 submissions with group changes, 4 finisher threads; every submission of
        // group n must finish before any of group n+1 starts
        struct Sub {
            uint64_t group;
        };
        OrderedGroups<int> og3;
        std::mutex mu;
        std::vector<int> runnable;  // ids that may start
        std::vector<uint64_t> group_of(800);
        std::atomic<int> started{0}, finished{0}, violations{0};
        std::map<uint64_t, int> running_per_group;
        uint64_t max_started_group = 0;
        std::atomic<bool> submit_done{false};
        auto start = [&](int id) {  // caller holds mu
            const uint64_t gg = group_of[id];
            for (auto& kv : running_per_group)
                if (kv.first < gg && kv.second > 0) violations++;  // an older group still has running work
            if (gg < max_started_group) violations++;
            max_started_group = std::max(max_started_group, gg);
            running_per_group[gg]++;
            runnable.push_back(id);
            started++;
        };
        std::thread submitter([&] {
            for (int id = 0; id < 800; id++) {
                int item = id;
                uint64_t gg;
                {
                    // group_of must be known before the item can be released by a finisher: submit under mu
                    std::lock_guard<std::mutex> lk(mu);
                    const bool now = og3.submit(item, &gg);
                    group_of[id] = gg;
                    if (now) start(id);
                }
                if (id % 7 == 6) og3.next_group();
            }
            submit_done = true;
        });
        std::vector<std::thread> finishers;
        for (int t = 0; t < 4; t++)
            finishers.emplace_back([&] {
                while (finished.load() < 800) {
                    int id = -1;
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        if (!runnable.empty()) {
                            id = runnable.back();
                            runnable.pop_back();
                        }
                    }
                    if (id < 0) {
                        std::this_thread::yield();
                        continue;
                    }
                    std::vector<int> rel;
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        running_per_group[group_of[id]]--;
                        og3.finished(group_of[id], true, &rel);
                        for (int r : rel) start(r);
                    }
                    finished++;
                }
            });
        submitter.join();
        for (auto& t : finishers) t.join();
        CHECK(15, started.load() == 800 && finished.load() == 800);
        CHECK(16, violations.load() == 0);
        CHECK(17, og3.held_count() == 0);
        return 0;
    }
    default:
        return -22;
    }
}
