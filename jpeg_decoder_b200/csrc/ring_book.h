// ring_book.h -- bookkeeping of a single-producer ring of contiguous, variable-size regions (no memory of its own).
// The producer reserves a worst-case region, fills it, commits the bytes actually used; regions are released in
// commit order by whoever consumed them (another thread).  head/tail are monotonic byte counters.
//   * a region never straddles the end: a request that would skips to offset 0; the skipped bytes return with
//     the next release -- except when the ring is empty, where nothing is in flight to return them, so both
//     counters jump (without this an empty ring can refuse a request that fits: (cap - pos) + need > cap);
//   * a reservation that is never committed leaks nothing.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <atomic>

namespace b200jpg {

class RingBook {
public:
    void reset(size_t cap) {
        cap_ = cap;
        head_ = pending_ = 0;
        tail_.store(0);
    }
    size_t cap() const { return cap_; }
    bool empty() const { return tail_.load() == head_; }  // producer thread only
    // Producer: where would `need` contiguous bytes go?  false while regions still in flight are in the way
    // (wait for a release and ask again); need must be <= cap().
    bool try_reserve(size_t need, size_t* pos) {
        uint64_t h = head_;
        size_t p = (size_t)(h % cap_);
        if (p + need > cap_) {
            h += cap_ - p;
            p = 0;
        }
        const uint64_t t = tail_.load();
        if (t == head_) {  // empty: nobody else touches tail now
            head_ = h;
            tail_.store(h);
        } else if (h + need - t > cap_) {
            return false;
        }
        pending_ = h;
        *pos = p;
        return true;
    }
    // Producer: the reserved region holds `used` bytes.  Returns the ticket for release().
    uint64_t commit(size_t used) {
        head_ = pending_ + used;
        return head_;
    }
    // Consumer, in commit order: everything up to this ticket is free again.
    void release(uint64_t ticket) {
        if (ticket > tail_.load()) tail_.store(ticket);  // one consumer; never moves backwards
    }

private:
    size_t cap_ = 0;
    uint64_t head_ = 0, pending_ = 0;
    std::atomic<uint64_t> tail_{0};
};

}  // namespace b200jpg
