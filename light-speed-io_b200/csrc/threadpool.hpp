// threadpool.hpp -- work-stealing pool: the B200 build's `lsio_threadpool`
// (crates/lsio_threadpool/src/threadpool.rs:20-165, worker.rs:17-102, park_manager.rs:11-71,
// shared_state.rs:12-47), re-designed for "one worker group per GPU pinned to that GPU's
// NUMA/PCIe root" (BASELINE.json north_star).
//
// Same contract as the reference: a global injector queue fed by ThreadPool::push, one FIFO
// deque per worker, find_task() = local pop -> batch from the injector -> steal one from a
// peer (worker.rs:45-64); WorkerThread::push puts follow-up tasks on the local deque where
// peers can steal them and wakes up to `len` parked peers (worker.rs:91-101); park()/unpark
// through a parked-thread FIFO (the reference's ParkManager thread is folded into a mutex:
// same order of wake-ups, one thread fewer); destruction sets keep_running = false, unparks
// and joins everyone (threadpool.rs:140-165).  Tasks here are coarse (one file or one >= 4 KiB
// read), so the deques are mutex-protected rather than lock-free.
#pragma once

#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <thread>
#include <vector>

namespace lsio {

template <class T>
class ThreadPool;

template <class T>
class WorkerThread {
  public:
    // worker.rs:45-64
    std::optional<T> find_task() {
        if (auto t = pop(local())) return t;
        {
            std::lock_guard<std::mutex> lk(s_->injector_mu);
            if (!s_->injector.empty()) {
                T first = std::move(s_->injector.front());
                s_->injector.pop_front();
                // steal_batch_and_pop: move up to half of the rest (bounded) to the local deque
                size_t take = std::min<size_t>(s_->injector.size() / 2, 32);
                if (take) {
                    std::lock_guard<std::mutex> lk2(local().mu);
                    for (size_t i = 0; i < take; i++) {
                        local().q.push_back(std::move(s_->injector.front()));
                        s_->injector.pop_front();
                    }
                }
                return first;
            }
        }
        const size_t n = s_->locals.size();
        for (size_t k = 1; k < n; k++) {
            Local& victim = *s_->locals[(id_ + k) % n];
            if (s_->steal_batch <= 1) {  // the reference's Stealer::steal(): one task (worker.rs:58-62)
                if (auto t = pop(victim)) return t;
                continue;
            }
            // take the LAST steal_batch tasks, in their order: the owner keeps walking its queue from the front without
            // holes (consecutive reads of a request fill consecutive extents of one worker, which is what lets their
            // copies be coalesced), the thief gets a run of consecutive reads of its own
            std::vector<T> got;
            {
                std::lock_guard<std::mutex> lk(victim.mu);
                const size_t take = std::min<size_t>(victim.q.size(), s_->steal_batch);
                for (size_t i = victim.q.size() - take; i < victim.q.size(); i++) got.push_back(std::move(victim.q[i]));
                victim.q.erase(victim.q.end() - (std::ptrdiff_t)take, victim.q.end());
            }
            if (got.empty()) continue;
            if (got.size() > 1) {
                std::lock_guard<std::mutex> lk(local().mu);
                for (size_t i = 1; i < got.size(); i++) local().q.push_back(std::move(got[i]));
            }
            return std::move(got[0]);
        }
        return std::nullopt;
    }
    bool keep_running() const { return s_->keep_running.load(std::memory_order_relaxed); }  // worker.rs:67-69
    // worker.rs:75-86. Returns after an unpark (or a short safety timeout: the reference has a
    // lost-wake-up window between the flag and park(); the timeout closes it).
    void park() {
        auto& me = *s_->parkers[id_];
        {
            std::lock_guard<std::mutex> lk(s_->park_mu);
            s_->parked.push_back(id_);
            s_->at_least_one_parked.store(true, std::memory_order_relaxed);
        }
        s_->n_parks[id_].fetch_add(1, std::memory_order_relaxed);
        {
            std::unique_lock<std::mutex> lk(me.mu);
            me.cv.wait_for(lk, std::chrono::milliseconds(s_->park_timeout_ms), [&] { return me.token; });
            me.token = false;
        }
        std::lock_guard<std::mutex> lk(s_->park_mu);
        for (auto it = s_->parked.begin(); it != s_->parked.end(); ++it)
            if (*it == id_) {
                s_->parked.erase(it);
                break;
            }
        if (s_->parked.empty()) s_->at_least_one_parked.store(false, std::memory_order_relaxed);
    }
    // worker.rs:91-101
    void push(T task) {
        size_t n;
        {
            std::lock_guard<std::mutex> lk(local().mu);
            local().q.push_back(std::move(task));
            n = local().q.size();
        }
        if (n > 1) s_->unpark_at_most_n_threads((unsigned)n);
    }
    size_t id() const { return id_; }
    size_t local_len() {
        std::lock_guard<std::mutex> lk(local().mu);
        return local().q.size();
    }

  private:
    friend class ThreadPool<T>;
    struct Local {
        std::mutex mu;
        std::deque<T> q;
    };
    struct Parker {
        std::mutex mu;
        std::condition_variable cv;
        bool token = false;
    };
    struct Shared {  // shared_state.rs:12-24
        std::mutex injector_mu;
        std::deque<T> injector;
        std::atomic<bool> keep_running{true};
        std::atomic<bool> at_least_one_parked{false};
        std::mutex park_mu;
        std::deque<size_t> parked;  // FIFO, park_manager.rs:21
        std::vector<std::unique_ptr<Local>> locals;
        std::vector<std::unique_ptr<Parker>> parkers;
        std::unique_ptr<std::atomic<uint64_t>[]> n_parks;
        int park_timeout_ms = 50;
        size_t steal_batch = 1;
        void unpark(size_t id) {
            auto& p = *parkers[id];
            std::lock_guard<std::mutex> lk(p.mu);
            p.token = true;
            p.cv.notify_one();
        }
        // shared_state.rs:26-32 + park_manager.rs:60-70
        void unpark_at_most_n_threads(unsigned n) {
            if (!at_least_one_parked.load(std::memory_order_relaxed)) return;
            std::lock_guard<std::mutex> lk(park_mu);
            while (n-- && !parked.empty()) {
                size_t id = parked.front();
                parked.pop_front();
                unpark(id);
            }
            if (parked.empty()) at_least_one_parked.store(false, std::memory_order_relaxed);
        }
    };
    WorkerThread(std::shared_ptr<Shared> s, size_t id) : s_(std::move(s)), id_(id) {}
    Local& local() { return *s_->locals[id_]; }
    static std::optional<T> pop(Local& l) {
        std::lock_guard<std::mutex> lk(l.mu);
        if (l.q.empty()) return std::nullopt;
        T t = std::move(l.q.front());  // FIFO (Worker::new_fifo, threadpool.rs:99)
        l.q.pop_front();
        return t;
    }
    std::shared_ptr<Shared> s_;
    size_t id_;
};

template <class T>
class ThreadPool {
  public:
    using Op = std::function<void(WorkerThread<T>&)>;
    // threadpool.rs:78-127.  `cpus`: if non-empty every worker is pinned to this CPU set
    // (the NUMA node of the group's GPU).
    ThreadPool(size_t n_worker_threads, Op op, const std::vector<int>& cpus = {}, size_t steal_batch = 1)
        : s_(std::make_shared<typename WorkerThread<T>::Shared>()) {
        s_->steal_batch = steal_batch ? steal_batch : 1;
        s_->n_parks.reset(new std::atomic<uint64_t>[n_worker_threads]);
        for (size_t i = 0; i < n_worker_threads; i++) {
            s_->locals.emplace_back(new typename WorkerThread<T>::Local());
            s_->parkers.emplace_back(new typename WorkerThread<T>::Parker());
            s_->n_parks[i].store(0);
        }
        for (size_t i = 0; i < n_worker_threads; i++) {
            threads_.emplace_back([s = s_, i, op, cpus] {
                if (!cpus.empty()) {
                    cpu_set_t set;
                    CPU_ZERO(&set);
                    for (int c : cpus) CPU_SET(c, &set);
                    pthread_setaffinity_np(pthread_self(), sizeof set, &set);
                }
                WorkerThread<T> w(s, i);
                op(w);
            });
        }
    }
    // threadpool.rs:134-137
    void push(T task) {
        {
            std::lock_guard<std::mutex> lk(s_->injector_mu);
            s_->injector.push_back(std::move(task));
        }
        s_->unpark_at_most_n_threads(1);
    }
    size_t injector_len() {
        std::lock_guard<std::mutex> lk(s_->injector_mu);
        return s_->injector.size();
    }
    uint64_t n_parks(size_t worker) const { return s_->n_parks[worker].load(); }
    size_t size() const { return threads_.size(); }
    void wake_all() {
        for (size_t i = 0; i < threads_.size(); i++) s_->unpark(i);
    }
    // Drop, threadpool.rs:140-165
    void shutdown() {
        if (joined_) return;
        s_->keep_running.store(false, std::memory_order_relaxed);
        for (size_t i = 0; i < threads_.size(); i++) s_->unpark(i);
        for (auto& t : threads_) t.join();
        joined_ = true;
    }
    ~ThreadPool() { shutdown(); }

  private:
    std::shared_ptr<typename WorkerThread<T>::Shared> s_;
    std::vector<std::thread> threads_;
    bool joined_ = false;
};

}  // namespace lsio
