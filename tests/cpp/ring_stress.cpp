// Stress test of csrc/ring_book.h: producers reserve worst-case regions, commit less, a consumer releases in
// order after a delay.  Detects overlap (pattern corruption) and deadlock (watchdog).  Built and run by
// tests/test_ring_book.py with g++; no CUDA.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "../../jpeg_decoder_b200/csrc/ring_book.h"

using b200jpg::RingBook;

struct Item {
    int thread;
    size_t pos, len;
    uint64_t ticket;
    unsigned char tag;
};

int main(int argc, char** argv) {
    const int nthreads = argc > 1 ? atoi(argv[1]) : 4;
    const int per_thread = argc > 2 ? atoi(argv[2]) : 20000;
    const size_t cap = argc > 3 ? (size_t)atol(argv[3]) : 4096;
    std::vector<RingBook> books((size_t)nthreads);
    std::vector<std::vector<unsigned char>> mem((size_t)nthreads, std::vector<unsigned char>(cap));
    for (auto& b : books) b.reset(cap);
    std::mutex mu;
    std::condition_variable space_cv, items_cv;
    std::deque<Item> queue;
    std::atomic<int> active(nthreads);
    std::atomic<long> progress(0);
    std::atomic<bool> corrupt(false);

    auto producer = [&](int tid) {
        std::mt19937 rng(1234u + (unsigned)tid);
        RingBook& rb = books[(size_t)tid];
        for (int i = 0; i < per_thread; i++) {
            // mostly mid-size requests, sometimes nearly the whole ring (the case an empty ring must accept)
            const size_t need = rng() % 16 == 0 ? cap - rng() % 8 : 1 + rng() % (cap / 2);
            size_t pos = 0;
            if (!rb.try_reserve(need, &pos)) {
                std::unique_lock<std::mutex> lk(mu);
                space_cv.wait(lk, [&] { return rb.try_reserve(need, &pos); });
            }
            if (rng() % 10 == 0) continue;  // abandoned reservation ("decode failed")
            const size_t used = 1 + rng() % need;
            const unsigned char tag = (unsigned char)(rng() & 0xff);
            memset(mem[(size_t)tid].data() + pos, tag, used);
            Item it{tid, pos, used, rb.commit(used), tag};
            {
                std::lock_guard<std::mutex> lk(mu);
                queue.push_back(it);
            }
            items_cv.notify_one();
            progress++;
        }
        active--;
        items_cv.notify_one();
    };
    auto consumer = [&] {
        std::mt19937 rng(99);
        for (;;) {
            std::vector<Item> batch;
            {
                std::unique_lock<std::mutex> lk(mu);
                items_cv.wait_for(lk, std::chrono::microseconds(200), [&] { return !queue.empty() || active.load() == 0; });
                while (!queue.empty() && batch.size() < 8) {
                    batch.push_back(queue.front());
                    queue.pop_front();
                }
                if (batch.empty() && active.load() == 0) return;
            }
            if (rng() % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 50));
            for (const Item& it : batch) {
                const unsigned char* p = mem[(size_t)it.thread].data() + it.pos;
                for (size_t k = 0; k < it.len; k++)
                    if (p[k] != it.tag) corrupt = true;  // somebody wrote over a region still in flight
                books[(size_t)it.thread].release(it.ticket);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
            }
            space_cv.notify_all();
        }
    };
    std::thread watchdog([&] {
        long last = -1;
        for (;;) {
            std::this_thread::sleep_for(std::chrono::seconds(5));
            if (active.load() == 0) return;
            const long now = progress.load();
            if (now == last) {
                fprintf(stderr, "DEADLOCK: no progress for 5 s at %ld items\n", now);
                _Exit(3);
            }
            last = now;
        }
    });
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(producer, t);
    std::thread cons(consumer);
    for (auto& t : th) t.join();
    cons.join();
    watchdog.detach();
    if (corrupt) {
        fprintf(stderr, "CORRUPTION: a region was overwritten while in flight\n");
        return 2;
    }
    printf("ok %ld items\n", progress.load());
    return 0;
}
