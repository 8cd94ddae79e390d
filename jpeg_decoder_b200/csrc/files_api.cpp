// files_api.cpp -- whole-file batch decoding: n JPEG files in, n pixel buffers out.
//
// What a user of the reference gets from an outer `par_iter` over `Decoder::decode()` (one image per host
// thread), re-cut for the GPU worker.  Host threads do only what is inherently serial per image
// -- marker parsing and Huffman decoding (HostDecoder, csrc/host_decoder.cpp) -- and write what they decode as
// a sparse block stream (sbs.h) into their own page-locked ring.  Finished images go onto a queue; the calling
// thread drains it in small groups into the three-stream device pipeline of sbs_pipeline.h (H2D | K0+K1+K2 |
// D2H).  There is no barrier anywhere: host threads keep decoding while earlier images upload, compute and
// download; a ring region is recycled as soon as its upload has finished.
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
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200jpg.h"
#include "context.h"
#include "host_decoder.h"
#include "ring_book.h"
#include "sbs_pipeline.h"

using b200jpg::HostDecoder;
using b200jpg::SbsItem;
using b200jpg::SbsLayout;
using b200jpg::SbsPipeline;

namespace {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <typename F>
void parallel_for(size_t begin, size_t end, int nthreads, F&& fn) {
    if (end <= begin) return;
    nthreads = (int)std::min<size_t>((size_t)std::max(1, nthreads), end - begin);
    if (nthreads == 1) {
        for (size_t i = begin; i < end; i++) fn(i);
        return;
    }
    std::atomic<size_t> next(begin);
    std::vector<std::thread> th;
    th.reserve((size_t)nthreads);
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&] {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= end) break;
                fn(i);
            }
        });
    for (auto& t : th) t.join();
}

void fill_info(const HostDecoder& hd, b200jpg_file_job* job) {
    const auto& f = hd.frame();
    job->info.width = f.output_w;
    job->info.height = f.output_h;
    job->info.pixel_format = hd.pixel_format();
    job->info.coding_process = f.coding_process;
    job->out_len = (size_t)f.output_w * f.output_h * f.comps.size();
}

int default_threads() {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) return std::max(1, CPU_COUNT(&set));
    return (int)std::max(1u, std::thread::hardware_concurrency());
}

// One host thread's page-locked ring: memory + the bookkeeping of ring_book.h (regions are released by the
// submitter once their upload has completed).
struct Ring {
    uint8_t* base = nullptr;
    b200jpg::RingBook book;
};

// Persistent per-context state: the rings (pinning memory costs ~0.3 s per GB) and the device pipeline.
struct FilesEngine {
    b200jpg_ctx* ctx;
    std::vector<std::unique_ptr<Ring>> rings;
    std::unique_ptr<SbsPipeline> pipe;
    std::mutex call_mu;  // one decode_files call at a time per context
    explicit FilesEngine(b200jpg_ctx* c) : ctx(c) {}
    ~FilesEngine() {
        pipe.reset();
        for (auto& r : rings)
            if (r && r->base) cudaFreeHost(r->base);
    }
};

void engine_free(void* p) { delete (FilesEngine*)p; }

FilesEngine* get_engine(b200jpg_ctx* ctx) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->files_engine) {
        ctx->files_engine = new FilesEngine(ctx);
        ctx->files_engine_free = engine_free;
    }
    return (FilesEngine*)ctx->files_engine;
}

constexpr size_t kMinRing = (size_t)12 << 20;
constexpr size_t kMaxSbsImage = (size_t)1 << 30;  // larger images take the dense path, one at a time

struct CallState {
    b200jpg_file_job* jobs = nullptr;
    size_t n = 0;
    std::atomic<size_t> next{0};
    std::mutex mu;
    std::condition_variable items_cv, space_cv;
    std::deque<SbsItem> queue;
    int workers_active = 0;
    std::mutex gpu_mu;  // the dense fallback and the submitter share the context
    std::atomic<int> device_error{B200JPG_OK};
    // tracing
    std::atomic<uint64_t> ring_wait_us{0}, decode_us{0};
};

}  // namespace

extern "C" {

int b200jpg_read_info_files(b200jpg_file_job* jobs, size_t n, int nthreads) {
    if (!jobs && n) return B200JPG_ERR_INTERNAL;
    if (nthreads < 1) nthreads = default_threads();
    parallel_for(0, n, nthreads, [&](size_t i) {
        HostDecoder hd(jobs[i].data, jobs[i].len);
        jobs[i].status = hd.read_info();
        jobs[i].out_len = 0;
        if (jobs[i].status == B200JPG_OK) fill_info(hd, &jobs[i]);
    });
    return B200JPG_OK;
}

int b200jpg_decode_files(b200jpg_ctx* ctx, b200jpg_file_job* jobs, size_t n, int nthreads) {
    if (!ctx || (!jobs && n)) return B200JPG_ERR_INTERNAL;
    if (n == 0) return B200JPG_OK;
    if (nthreads < 1) nthreads = default_threads();
    nthreads = (int)std::min<size_t>((size_t)nthreads, n);
    const bool trace = getenv("B200JPG_TRACE") != nullptr;
    const double t_start = now_ms();
    FilesEngine* eng = get_engine(ctx);
    std::lock_guard<std::mutex> call_lock(eng->call_mu);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return B200JPG_ERR_INTERNAL;
    if (!eng->pipe) {
        eng->pipe.reset(new SbsPipeline(ctx, 4));
        if (!eng->pipe->ok()) {
            eng->pipe.reset();
            ctx->err = "internal: could not create the device pipeline (streams / events)";
            return B200JPG_ERR_INTERNAL;
        }
    }
    while (eng->rings.size() < (size_t)nthreads) eng->rings.emplace_back(new Ring());

    CallState cs;
    cs.jobs = jobs;
    cs.n = n;
    cs.workers_active = nthreads;

    auto worker = [&](int tid) {
        cudaSetDevice(ctx->device);
        Ring& ring = *eng->rings[(size_t)tid];
        for (;;) {
            const size_t i = cs.next.fetch_add(1);
            if (i >= n) break;
            b200jpg_file_job& job = jobs[i];
            const double t0 = now_ms();
            HostDecoder hd(job.data, job.len);
            job.status = hd.read_info();
            job.out_len = 0;
            if (job.status != B200JPG_OK) continue;
            fill_info(hd, &job);
            if (!job.out || job.out_cap < job.out_len) {
                job.status = B200JPG_ERR_INTERNAL;
                continue;
            }
            const size_t nb = hd.total_blocks();
            const SbsLayout lay = SbsLayout::make(nb);
            const size_t need = (lay.worst_bytes() + 4 * 128 + 255) / 256 * 256;  // stream + the quantisation tables
            if (need > kMaxSbsImage) {
                // dense path, synchronously (the image alone is a GPU-sized batch)
                job.status = hd.entropy_decode();
                if (job.status != B200JPG_OK) continue;
                b200jpg_image_desc d;
                memset(&d, 0, sizeof d);
                const auto& f = hd.frame();
                d.width = f.output_w;
                d.height = f.output_h;
                d.ncomp = (uint8_t)f.comps.size();
                d.color_transform = (uint8_t)hd.determine_color_transform();
                bool complete = true;
                for (size_t k = 0; k < f.comps.size() && k < 4; k++) {
                    complete = complete && hd.component_has_data((int)k);
                    d.comps[k] = f.comps[k];
                    d.qt[k] = hd.component_qtable((int)k);
                    d.coefs[k] = hd.coefficients((int)k);
                }
                if (!complete) {
                    job.status = B200JPG_ERR_FORMAT;
                    continue;
                }
                std::lock_guard<std::mutex> g(cs.gpu_mu);
                int st = B200JPG_OK;
                const int rc = b200jpg_decode_batch(ctx, &d, 1, &job.out, &job.out_cap, &st);
                job.status = rc != B200JPG_OK && st == B200JPG_OK ? rc : st;
                continue;
            }
            // a contiguous worst-case region in this thread's ring
            if (ring.book.cap() < need) {  // (re)allocate once everything handed out earlier has been uploaded
                if (!ring.book.empty()) {
                    const double w0 = now_ms();
                    std::unique_lock<std::mutex> lk(cs.mu);
                    cs.space_cv.wait(lk, [&] { return ring.book.empty(); });
                    cs.ring_wait_us += (uint64_t)((now_ms() - w0) * 1e3);
                }
                if (ring.base) cudaFreeHost(ring.base);
                ring.base = nullptr;
                ring.book.reset(0);
                const size_t want = std::max(kMinRing, need * 2 + need / 2);
                void* p = nullptr;
                if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
                    cudaGetLastError();
                    job.status = B200JPG_ERR_INTERNAL;
                    continue;
                }
                ring.base = (uint8_t*)p;
                ring.book.reset(want);
            }
            size_t pos = 0;
            if (!ring.book.try_reserve(need, &pos)) {
                const double w0 = now_ms();
                std::unique_lock<std::mutex> lk(cs.mu);
                cs.space_cv.wait(lk, [&] { return ring.book.try_reserve(need, &pos); });
                cs.ring_wait_us += (uint64_t)((now_ms() - w0) * 1e3);
            }
            hd.set_sbs_sink(ring.base + pos);
            job.status = hd.entropy_decode();
            if (job.status != B200JPG_OK) continue;
            const auto& f = hd.frame();
            bool complete = true;
            for (size_t k = 0; k < f.comps.size(); k++) complete = complete && hd.component_has_data((int)k);
            if (!complete || hd.sbs_length() == 0) {  // "not all components have data", src/decoder.rs:1306-1308
                job.status = B200JPG_ERR_FORMAT;
                continue;
            }
            SbsItem item;
            memset(&item.desc, 0, sizeof item.desc);
            item.desc.width = f.output_w;
            item.desc.height = f.output_h;
            item.desc.ncomp = (uint8_t)f.comps.size();
            item.desc.color_transform = (uint8_t)hd.determine_color_transform();
            // the quantisation tables must outlive the decoder: they ride at the end of the ring region
            const size_t len = hd.sbs_length();
            uint16_t* qcopy = (uint16_t*)(ring.base + pos + len);
            size_t extra = 0;
            for (size_t k = 0; k < f.comps.size() && k < 4; k++) {
                item.desc.comps[k] = f.comps[k];
                memcpy(qcopy + 64 * k, hd.component_qtable((int)k), 128);
                item.desc.qt[k] = qcopy + 64 * k;
                extra += 128;
            }
            item.stream = ring.base + pos;
            item.len = len;
            item.order = hd.sbs_order();
            item.out = job.out;
            item.out_cap = job.out_cap;
            item.job = i;
            item.thread = tid;
            item.ring_end = ring.book.commit((len + extra + 255) / 256 * 256);
            cs.decode_us += (uint64_t)((now_ms() - t0) * 1e3);
            {
                std::lock_guard<std::mutex> lk(cs.mu);
                cs.queue.push_back(item);
            }
            cs.items_cv.notify_one();
        }
        {
            std::lock_guard<std::mutex> lk(cs.mu);
            cs.workers_active--;
        }
        cs.items_cv.notify_one();
    };

    SbsPipeline& pipe = *eng->pipe;
    pipe.on_h2d = [&](const SbsPipeline::Group& g) {
        for (const SbsItem& it : g.items) eng->rings[(size_t)it.thread]->book.release(it.ring_end);
        {
            std::lock_guard<std::mutex> lk(cs.mu);
        }
        cs.space_cv.notify_all();
    };
    pipe.on_done = [&](const SbsPipeline::Group& g) {
        for (size_t k = 0; k < g.items.size(); k++) jobs[g.items[k].job].status = g.statuses[k];
    };

    std::vector<std::thread> threads;
    threads.reserve((size_t)nthreads);
    for (int t = 0; t < nthreads; t++) threads.emplace_back(worker, t);

    // the submitter: group whatever has been decoded (bounded by bytes and count) and push it to the device
    const size_t max_items = 48, max_bytes = (size_t)192 << 20;
    size_t ngroups = 0, nitems = 0;
    double idle_ms = 0, submit_ms = 0;
    int result = B200JPG_OK;
    for (;;) {
        std::vector<SbsItem> group;
        bool finished = false;
        {
            std::unique_lock<std::mutex> lk(cs.mu);
            const double w0 = now_ms();
            if (cs.queue.empty() && cs.workers_active > 0) cs.items_cv.wait_for(lk, std::chrono::microseconds(200));
            // a short second wait lets a few more images join a very small group (fewer, larger launches)
            if (!cs.queue.empty() && cs.queue.size() < 4 && cs.workers_active > 0) cs.items_cv.wait_for(lk, std::chrono::microseconds(150));
            idle_ms += now_ms() - w0;
            size_t bytes = 0;
            while (!cs.queue.empty() && group.size() < max_items && bytes < max_bytes) {
                const SbsItem& it = cs.queue.front();
                for (int k = 0; k < it.desc.ncomp; k++) bytes += (size_t)it.desc.comps[k].block_w * it.desc.comps[k].block_h * 128;
                group.push_back(it);
                cs.queue.pop_front();
            }
            finished = cs.queue.empty() && cs.workers_active == 0 && group.empty();
        }
        const double s0 = now_ms();
        {
            std::lock_guard<std::mutex> g(cs.gpu_mu);
            pipe.poll();
            if (!group.empty()) {
                ngroups++;
                nitems += group.size();
                std::vector<SbsItem> copy = group;
                const int rc = pipe.submit(std::move(group));
                if (rc != B200JPG_OK) {  // device-level failure: these images fail, their ring space is released
                    result = rc;
                    for (const SbsItem& it : copy) {
                        jobs[it.job].status = rc;
                        eng->rings[(size_t)it.thread]->book.release(it.ring_end);
                    }
                    {
                        std::lock_guard<std::mutex> lk(cs.mu);
                    }
                    cs.space_cv.notify_all();
                }
            }
        }
        submit_ms += now_ms() - s0;
        if (finished) break;
    }
    for (auto& t : threads) t.join();
    {
        std::lock_guard<std::mutex> g(cs.gpu_mu);
        const int rc = pipe.drain();
        if (rc != B200JPG_OK) result = rc;
    }
    pipe.on_h2d = nullptr;
    pipe.on_done = nullptr;
    if (trace)
        fprintf(stderr,
                "[b200jpg] decode_files: %zu images, %d host threads, %.1f ms; %zu groups (%.1f images each); submitter idle %.1f ms, "
                "busy %.1f ms; host threads: decode %.1f ms/image, waiting for ring space %.1f ms in total\n",
                n, nthreads, now_ms() - t_start, ngroups, ngroups ? (double)nitems / ngroups : 0.0, idle_ms, submit_ms,
                nitems ? cs.decode_us.load() / 1e3 / nitems : 0.0, cs.ring_wait_us.load() / 1e3);
    return result;
}

}  // extern "C"
