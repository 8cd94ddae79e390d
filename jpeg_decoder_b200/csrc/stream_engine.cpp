// stream_engine.cpp -- see stream_engine.h.
#include "stream_engine.h"

#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "context.h"
#include "ring_book.h"

namespace b200jpg {

namespace {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// One host thread's page-locked ring: memory + the bookkeeping of ring_book.h (regions are released by the
// submitter once their upload has completed).
struct Ring {
    uint8_t* base = nullptr;
    RingBook book;
};

// Persistent host threads: spawning 64-128 threads per call costs milliseconds, a call is tens of them.
class Pool {
public:
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void start(int n, std::function<void(int)> fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = std::move(fn);
            want_ = n;
            running_ = n;
            gen_++;
            while ((int)th_.size() < n) {
                const int tid = (int)th_.size();
                th_.emplace_back([this, tid] { body(tid); });
            }
        }
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return running_ == 0; });
    }

private:
    void body(int tid) {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            if (tid >= want_) continue;
            std::function<void(int)> fn = fn_;
            lk.unlock();
            fn(tid);
            lk.lock();
            if (--running_ == 0) done_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> th_;
    std::function<void(int)> fn_;
    uint64_t gen_ = 0;
    int want_ = 0, running_ = 0;
    bool stop_ = false;
};

// Persistent per-context state: rings (pinning memory costs ~0.3 s per GB), threads, the device pipeline.
struct Engine {
    b200jpg_ctx* ctx;
    std::vector<std::unique_ptr<Ring>> rings;
    std::unique_ptr<SbsPipeline> pipe;
    Pool pool;
    std::mutex call_mu;  // one run at a time per context
    explicit Engine(b200jpg_ctx* c) : ctx(c) {}
    ~Engine() {
        pipe.reset();
        for (auto& r : rings)
            if (r && r->base) cudaFreeHost(r->base);
    }
};

void engine_free(void* p) { delete (Engine*)p; }

Engine* get_engine(b200jpg_ctx* ctx) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->files_engine) {
        ctx->files_engine = new Engine(ctx);
        ctx->files_engine_free = engine_free;
    }
    return (Engine*)ctx->files_engine;
}

constexpr size_t kMinRing = (size_t)12 << 20;

}  // namespace

int stream_engine_default_threads() {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) return std::max(1, CPU_COUNT(&set));
    return (int)std::max(1u, std::thread::hardware_concurrency());
}

int stream_engine_run(b200jpg_ctx* ctx, JobSource& src, int nthreads) {
    const size_t n = src.size();
    if (!ctx) return B200JPG_ERR_INTERNAL;
    if (n == 0) return B200JPG_OK;
    if (nthreads < 1) nthreads = stream_engine_default_threads();
    nthreads = (int)std::min<size_t>((size_t)nthreads, n);
    const bool trace = getenv("B200JPG_TRACE") != nullptr;
    const double t_start = now_ms();
    Engine* eng = get_engine(ctx);
    std::lock_guard<std::mutex> call_lock(eng->call_mu);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return B200JPG_ERR_INTERNAL;
    if (!eng->pipe) {
        eng->pipe.reset(new SbsPipeline(ctx, 8));
        if (!eng->pipe->ok()) {
            eng->pipe.reset();
            return b200jpg_fail(ctx, B200JPG_ERR_INTERNAL, "internal: could not create the device pipeline (streams / events)");
        }
    }
    while (eng->rings.size() < (size_t)nthreads) eng->rings.emplace_back(new Ring());

    std::atomic<size_t> next{0};
    std::mutex mu;  // queue, workers_active; also the mutex of both condition variables
    std::condition_variable items_cv, space_cv;
    std::deque<SbsItem> queue;
    int workers_active = nthreads;
    std::mutex gpu_mu;  // sources that fall back to the dense path share the context with the submitter
    std::atomic<uint64_t> ring_wait_us{0}, produce_us{0};
    std::atomic<int> host_error{B200JPG_OK};

    auto worker = [&](int tid) {
        cudaSetDevice(ctx->device);
        Ring& ring = *eng->rings[(size_t)tid];
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            const double t0 = now_ms();
            size_t need = 0;
            void* state = nullptr;
            int st = src.prepare(i, &need, &state, &gpu_mu);
            if (st != B200JPG_OK || need == 0) continue;  // the source has recorded the job's status
            need = (need + 512 + 255) / 256 * 256;        // + room for the quantisation tables
            if (ring.book.cap() < need) {  // (re)allocate once everything handed out earlier has been uploaded
                if (!ring.book.empty()) {
                    const double w0 = now_ms();
                    std::unique_lock<std::mutex> lk(mu);
                    space_cv.wait(lk, [&] { return ring.book.empty(); });
                    ring_wait_us += (uint64_t)((now_ms() - w0) * 1e3);
                }
                if (ring.base) cudaFreeHost(ring.base);
                ring.base = nullptr;
                ring.book.reset(0);
                const size_t want = std::max(kMinRing, need * 2 + need / 2);
                void* p = nullptr;
                if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
                    cudaGetLastError();
                    host_error = B200JPG_ERR_INTERNAL;
                    SbsItem dummy;
                    src.produce(i, state, nullptr, 0, &dummy);  // lets the source free its state and record the failure
                    continue;
                }
                ring.base = (uint8_t*)p;
                ring.book.reset(want);
            }
            size_t pos = 0;
            if (!ring.book.try_reserve(need, &pos)) {
                const double w0 = now_ms();
                std::unique_lock<std::mutex> lk(mu);
                space_cv.wait(lk, [&] { return ring.book.try_reserve(need, &pos); });
                ring_wait_us += (uint64_t)((now_ms() - w0) * 1e3);
            }
            SbsItem item;
            st = src.produce(i, state, ring.base + pos, need, &item);
            if (st != B200JPG_OK) continue;  // nothing was committed: the region is simply reused
            item.stream = ring.base + pos;
            item.job = i;
            item.thread = tid;
            item.mapped = true;  // the rings are cudaHostAlloc'ed
            item.ring_end = ring.book.commit((item.len + 512 + 255) / 256 * 256);
            produce_us += (uint64_t)((now_ms() - t0) * 1e3);
            {
                std::lock_guard<std::mutex> lk(mu);
                queue.push_back(item);
            }
            items_cv.notify_one();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            workers_active--;
        }
        items_cv.notify_one();
    };

    SbsPipeline& pipe = *eng->pipe;
    // (a call with a handful of images -- a lone Decoder::decode -- takes what it needs instead of 0.7-2 GB per slot)
    pipe.reserve_for_group_bytes(n >= 32 ? (size_t)(src.outputs_on_device() ? 640 : 160) << 20 : 0);
    pipe.grow_ms = pipe.enqueue_ms = pipe.retire_wait_ms = 0;
    pipe.grows = pipe.h2d_copies = pipe.d2h_copies = 0;
    pipe.timeline_begin();
    auto release = [&](const std::vector<SbsItem>& items) {
        for (const SbsItem& it : items) eng->rings[(size_t)it.thread]->book.release(it.ring_end);
        {
            std::lock_guard<std::mutex> lk(mu);
        }
        space_cv.notify_all();
    };
    pipe.on_h2d = [&](const SbsPipeline::Group& g) { release(g.items); };
    pipe.on_done = [&](const SbsPipeline::Group& g) {
        for (size_t k = 0; k < g.items.size(); k++) src.finish(g.items[k].job, g.statuses[k]);
    };

    eng->pool.start(nthreads, worker);

    // the submitter: group whatever is ready (bounded by bytes and count) and push it to the device
    // groups: small enough that the first download starts early and the last one is short when pixels go back over PCIe,
    // larger when they stay on the device (per-group latencies amortise over more images)
    const bool dev_out = src.outputs_on_device();
    const size_t max_items = dev_out ? 96 : 32, max_bytes = (size_t)(dev_out ? 640 : 160) << 20;
    // A short bounded wait lets a few more images join a very small group.  Waiting for LARGE groups when the pixels stay
    // on the device (to amortise the latency-bound synchronisation rounds of the entropy kernels) was measured and loses:
    // 54.6 GP/s with min 1, 50.4 / 47.8 / 45.7 with min 24 / 48 / 96 images (profiles/r02_files_group_size.jsonl) --
    // small groups taking turns on several compute streams overlap better than large ones amortise.  B200JPG_GROUP_MIN: knob.
    static const size_t min_items_env = getenv("B200JPG_GROUP_MIN") ? (size_t)atoi(getenv("B200JPG_GROUP_MIN")) : 0;
    const size_t min_items = min_items_env ? min_items_env : 4;
    const int fill_us = min_items_env > 4 ? 1500 : 150;
    size_t ngroups = 0, nitems = 0;
    double idle_ms = 0, submit_ms = 0;
    int result = B200JPG_OK;
    for (;;) {
        std::vector<SbsItem> group;
        bool finished = false;
        {
            std::unique_lock<std::mutex> lk(mu);
            const double w0 = now_ms();
            if (queue.empty() && workers_active > 0) items_cv.wait_for(lk, std::chrono::microseconds(200));
            // a bounded second wait lets more images join a small group (fewer, larger launches).  With pixels leaving over PCIe
            // and two groups already queued behind the download there is no hurry at all: small groups download at 25-46 GB/s
            // where groups of 16+ images reach 50 (per-group timeline), so wait up to a millisecond for 16: 16.4-16.7 GP/s in every
            // run, where eager grouping gave 16.5 or -- one run in three, and with 4 CPUs -- 14.5 (profiles/r02_files_group_relax_ab.jsonl).
            static const bool relax_off = getenv("B200JPG_GROUP_RELAX") && atoi(getenv("B200JPG_GROUP_RELAX")) == 0;
            const bool relaxed = !relax_off && !dev_out && pipe.in_flight() >= 2;
            const size_t want_items = relaxed ? std::max<size_t>(min_items, 16) : min_items;
            if (!queue.empty() && queue.size() < want_items && workers_active > 0) {
                const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(relaxed ? std::max(fill_us, 1000) : fill_us);
                while (queue.size() < want_items && workers_active > 0 && (ngroups > 0 || !dev_out) &&
                       items_cv.wait_until(lk, deadline) != std::cv_status::timeout) {
                }
            }
            idle_ms += now_ms() - w0;
            // lowest job numbers first: a group is then (nearly) a run of consecutive jobs, whose pixel buffers callers usually
            // lay out back to back -- they leave the device as a few large copies instead of one per image.  (Every thread
            // takes its jobs in ascending order, so this never overtakes an earlier region of a thread's ring.)
            std::sort(queue.begin(), queue.end(), [](const SbsItem& a, const SbsItem& b) { return a.job < b.job; });
            size_t bytes = 0;
            while (!queue.empty() && group.size() < max_items && bytes < max_bytes) {
                const SbsItem& it = queue.front();
                for (int k = 0; k < it.desc.ncomp; k++) bytes += (size_t)it.desc.comps[k].block_w * it.desc.comps[k].block_h * 128;
                group.push_back(it);
                queue.pop_front();
            }
            finished = queue.empty() && workers_active == 0 && group.empty();
        }
        const double s0 = now_ms();
        {
            std::lock_guard<std::mutex> g(gpu_mu);
            pipe.poll();
            if (!group.empty()) {
                ngroups++;
                nitems += group.size();
                std::vector<SbsItem> copy = group;
                const int rc = pipe.submit(std::move(group));
                if (rc != B200JPG_OK) {  // device-level failure: these images fail, their ring space is released
                    result = rc;
                    for (const SbsItem& it : copy) src.finish(it.job, rc);
                    release(copy);
                }
            }
        }
        submit_ms += now_ms() - s0;
        if (finished) break;
    }
    eng->pool.wait();
    {
        std::lock_guard<std::mutex> g(gpu_mu);
        const int rc = pipe.drain();
        if (rc != B200JPG_OK) result = rc;
    }
    pipe.on_h2d = nullptr;
    pipe.on_done = nullptr;
    if (result == B200JPG_OK && host_error.load() != B200JPG_OK) result = host_error.load();
    if (trace)
        fprintf(stderr,
                "[b200jpg] %s: %zu images, %d host threads, %.1f ms; %zu groups (%.1f images each; %u upload and %u download copies); submitter idle %.1f ms, "
                "busy %.1f ms (enqueue %.1f, of it %u buffer regrowths %.1f; waiting for a free slot %.1f); host threads: %.2f ms/image, "
                "waiting for ring space %.1f ms in total\n",
                src.name(), n, nthreads, now_ms() - t_start, ngroups, ngroups ? (double)nitems / ngroups : 0.0, pipe.h2d_copies, pipe.d2h_copies, idle_ms, submit_ms,
                pipe.enqueue_ms, pipe.grows, pipe.grow_ms, pipe.retire_wait_ms, nitems ? produce_us.load() / 1e3 / nitems : 0.0,
                ring_wait_us.load() / 1e3);
    return result;
}

// ---------------------------------------------------------------------------------------------------------
// dense coefficient buffers -> streams (host-side compaction for b200jpg_batch_run_host)
// ---------------------------------------------------------------------------------------------------------
namespace {

constexpr size_t kMaxDenseSbsImage = (size_t)1 << 30;

class DenseSource : public JobSource {
public:
    DenseSource(b200jpg_ctx* ctx, const b200jpg_image_desc* imgs, size_t n, const int* plan_status, uint8_t* const* outs,
                const size_t* out_caps, int* statuses)
        : ctx_(ctx), imgs_(imgs), n_(n), plan_(plan_status), outs_(outs), caps_(out_caps), st_(statuses) {}
    size_t size() const override { return n_; }
    const char* name() const override { return "run_host(compacted)"; }
    int prepare(size_t i, size_t* need, void** state, std::mutex* gpu_mu) override {
        *state = nullptr;
        *need = 0;
        const b200jpg_image_desc& d = imgs_[i];
        if (plan_ && plan_[i] != B200JPG_OK) {  // rejected by the planner: reported per image only, like the dense path
            if (st_) st_[i] = plan_[i];
            return plan_[i];
        }
        if (d.ncomp < 1 || d.ncomp > 4) return set(i, B200JPG_ERR_INTERNAL);
        size_t nb = 0;
        for (int k = 0; k < d.ncomp; k++) {
            if (!d.coefs[k]) return set(i, B200JPG_ERR_FORMAT);  // "not all components have data", src/decoder.rs:1306-1308
            nb += (size_t)d.comps[k].block_w * d.comps[k].block_h;
        }
        const size_t worst = SbsLayout::make(nb).worst_bytes();
        if (worst > kMaxDenseSbsImage) {  // too large for a ring: the plain dense path, one image at a time
            std::lock_guard<std::mutex> g(*gpu_mu);
            int st = B200JPG_OK;
            const size_t cap = caps_ ? caps_[i] : (size_t)d.width * d.height * d.ncomp;
            const int rc = b200jpg_decode_batch(ctx_, &d, 1, &outs_[i], &cap, &st);
            set(i, rc != B200JPG_OK && st == B200JPG_OK ? rc : st);
            return B200JPG_OK;
        }
        *need = worst;
        return set(i, B200JPG_OK);
    }
    int produce(size_t i, void*, uint8_t* dst, size_t, SbsItem* item) override {
        if (!dst) return set(i, B200JPG_ERR_INTERNAL);
        const b200jpg_image_desc& d = imgs_[i];
        size_t nb = 0;
        for (int k = 0; k < d.ncomp; k++) nb += (size_t)d.comps[k].block_w * d.comps[k].block_h;
        SbsWriter w;
        w.begin(dst, nb);
        for (int k = 0; k < d.ncomp; k++) {
            const size_t cnt = (size_t)d.comps[k].block_w * d.comps[k].block_h;
            w.put_dense_natural_run(d.coefs[k], cnt);
        }
        item->desc = d;
        item->len = w.finish();
        item->order = SBS_PLANAR | SBS_NATURAL;
        item->out = outs_[i];
        item->out_cap = caps_ ? caps_[i] : (size_t)d.width * d.height * d.ncomp;
        return B200JPG_OK;
    }
    void finish(size_t i, int status) override { set(i, status); }

private:
    int set(size_t i, int st) {
        if (st_) st_[i] = st;
        if (st != B200JPG_OK) any_error_ = st;
        return st;
    }
    b200jpg_ctx* ctx_;
    const b200jpg_image_desc* imgs_;
    size_t n_;
    const int* plan_;
    uint8_t* const* outs_;
    const size_t* caps_;
    int* st_;

public:
    std::atomic<int> any_error_{B200JPG_OK};
};

}  // namespace

int stream_engine_run_dense(b200jpg_ctx* ctx, const b200jpg_image_desc* imgs, size_t n, const int* plan_status, uint8_t* const* outs,
                            const size_t* out_caps, int* statuses, int nthreads) {
    DenseSource src(ctx, imgs, n, plan_status, outs, out_caps, statuses);
    const int rc = stream_engine_run(ctx, src, nthreads);
    if (rc != B200JPG_OK) return rc;
    return src.any_error_.load();
}

double stream_engine_sample_density(const b200jpg_image_desc* imgs, size_t n) {
    size_t nz = 0, total = 0;
    const size_t step = std::max<size_t>(1, n / 8);
    for (size_t i = 0; i < n; i += step) {
        const b200jpg_image_desc& d = imgs[i];
        for (int k = 0; k < d.ncomp && k < 4; k++) {
            if (!d.coefs[k]) continue;
            const size_t cnt = (size_t)d.comps[k].block_w * d.comps[k].block_h;
            const size_t bstep = std::max<size_t>(1, cnt / 64);
            for (size_t b = 0; b < cnt; b += bstep) {
                const int16_t* c = d.coefs[k] + 64 * b;
                for (int j = 0; j < 64; j++) nz += c[j] != 0;
                total += 64;
            }
        }
    }
    return total ? (double)nz / (double)total : 1.0;
}

}  // namespace b200jpg
