// files_api.cpp -- whole-file batch decoding: n JPEG files in, n pixel buffers out.
//
// What a user of the reference gets from an outer `par_iter` over `Decoder::decode()` (one image per
// host thread), re-cut for the GPU worker: host threads do only what is inherently
// serial per image -- marker parsing and Huffman decoding (HostDecoder, csrc/host_decoder.cpp) -- writing
// the dense coefficients straight into one page-locked arena per chunk; a submitter thread then pushes the
// chunk through the batch path (one H2D, K1, K2, D2H into the callers' buffers) while the host threads
// already decode the next chunk.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200jpg.h"
#include "context.h"
#include "host_decoder.h"

using b200jpg::HostDecoder;

namespace {

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

struct Slot {  // one chunk in flight
    int16_t* arena = nullptr;  // page-locked, owned by the context
    std::vector<std::unique_ptr<HostDecoder>> decs;
    std::vector<b200jpg_image_desc> descs;
    std::vector<size_t> job_of_desc;
    std::thread gpu;
    int rc = B200JPG_OK;
    std::string err;
};

}  // namespace

extern "C" {

int b200jpg_read_info_files(b200jpg_file_job* jobs, size_t n, int nthreads) {
    if (!jobs && n) return B200JPG_ERR_INTERNAL;
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
    if (nthreads < 1) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    const size_t chunk = (size_t)std::max(8, 2 * nthreads);
    const bool trace = getenv("B200JPG_TRACE") != nullptr;  // phase timings on stderr
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    constexpr size_t NSLOTS = 3;
    Slot slots[NSLOTS];
    std::mutex gpu_mutex;  // one chunk at a time on the context's streams (the batch path pipelines internally)
    int result = B200JPG_OK;
    size_t which = 0;
    for (size_t i0 = 0; i0 < n; i0 += chunk, which = (which + 1) % NSLOTS) {
        const size_t i1 = std::min(n, i0 + chunk);
        Slot& s = slots[which];
        const double t_c0 = now();
        if (s.gpu.joinable()) s.gpu.join();  // the slot's previous chunk has left the GPU
        const double t_c1 = now();
        if (s.rc != B200JPG_OK) result = s.rc;
        s.decs.clear();
        s.decs.resize(i1 - i0);
        // phase 0: headers (sizes of the coefficient buffers)
        parallel_for(i0, i1, nthreads, [&](size_t i) {
            auto hd = std::make_unique<HostDecoder>(jobs[i].data, jobs[i].len);
            jobs[i].status = hd->read_info();
            jobs[i].out_len = 0;
            if (jobs[i].status == B200JPG_OK) {
                fill_info(*hd, &jobs[i]);
                // (lossless files are rejected by entropy_decode at their first scan, after the same header checks
                //  the reference performs, so the error class matches)
                if (jobs[i].out && jobs[i].out_cap < jobs[i].out_len) jobs[i].status = B200JPG_ERR_INTERNAL;
                else if (!jobs[i].out) jobs[i].status = B200JPG_ERR_INTERNAL;
            }
            s.decs[i - i0] = std::move(hd);
        });
        const double t_c2 = now();
        // arena layout: components of one image back to back, 1 KiB aligned like the device slab
        std::vector<size_t> off(i1 - i0 + 1, 0);
        size_t total = 0;
        for (size_t i = i0; i < i1; i++) {
            off[i - i0] = total;
            if (jobs[i].status != B200JPG_OK) continue;
            for (const auto& c : s.decs[i - i0]->frame().comps) total += ((size_t)c.block_w * c.block_h * 64 + 511) / 512 * 512;
        }
        off[i1 - i0] = total;
        {   // the page-locked arena lives in the context (grow-only): pinning memory costs ~0.3 s per GB
            b200jpg_ctx::Buf& pa = ctx->pinned[which];
            if (total * sizeof(int16_t) > pa.cap) {
                if (pa.p) cudaFreeHost(pa.p);
                pa.p = nullptr;
                pa.cap = 0;
                const size_t want = (total + total / 4) * sizeof(int16_t);
                if (cudaHostAlloc(&pa.p, want, cudaHostAllocDefault) != cudaSuccess) {
                    cudaGetLastError();
                    result = B200JPG_ERR_INTERNAL;
                    break;
                }
                pa.cap = want;
            }
            s.arena = (int16_t*)pa.p;
        }
        const double t_c3 = now();
        // phase 1: entropy decoding straight into the arena
        parallel_for(i0, i1, nthreads, [&](size_t i) {
            if (jobs[i].status != B200JPG_OK) return;
            HostDecoder& hd = *s.decs[i - i0];
            size_t o = off[i - i0];
            int k = 0;
            for (const auto& c : hd.frame().comps) {
                hd.set_external_buffer(k++, s.arena + o);
                o += ((size_t)c.block_w * c.block_h * 64 + 511) / 512 * 512;
            }
            jobs[i].status = hd.entropy_decode();
        });
        const double t_c4 = now();
        if (trace)
            fprintf(stderr, "[b200jpg] chunk %zu..%zu: wait-gpu %.1f ms, headers %.1f ms, arena %.1f ms, huffman %.1f ms (t=%.1f)\n", i0, i1,
                    t_c1 - t_c0, t_c2 - t_c1, t_c3 - t_c2, t_c4 - t_c3, t_c4 - t_start);
        // phase 2: hand the chunk to the GPU on a submitter thread
        s.descs.clear();
        s.job_of_desc.clear();
        for (size_t i = i0; i < i1; i++) {
            if (jobs[i].status != B200JPG_OK) continue;
            HostDecoder& hd = *s.decs[i - i0];
            const auto& f = hd.frame();
            b200jpg_image_desc d;
            memset(&d, 0, sizeof d);
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
            if (!complete) {  // "not all components have data", src/decoder.rs:1306-1308
                jobs[i].status = B200JPG_ERR_FORMAT;
                continue;
            }
            s.descs.push_back(d);
            s.job_of_desc.push_back(i);
        }
        s.rc = B200JPG_OK;
        s.gpu = std::thread([&s, &gpu_mutex, ctx, jobs, trace, now, t_start] {
            const size_t m = s.descs.size();
            if (m == 0) return;
            std::lock_guard<std::mutex> lock(gpu_mutex);
            const double t_g0 = now();
            std::vector<uint8_t*> outs(m);
            std::vector<size_t> caps(m);
            std::vector<int> st(m, 0);
            for (size_t k = 0; k < m; k++) {
                outs[k] = jobs[s.job_of_desc[k]].out;
                caps[k] = jobs[s.job_of_desc[k]].out_cap;
            }
            const int rc = b200jpg_decode_batch(ctx, s.descs.data(), m, outs.data(), caps.data(), st.data());
            if (trace) fprintf(stderr, "[b200jpg] gpu chunk of %zu images: %.1f ms (t=%.1f)\n", m, now() - t_g0, now() - t_start);
            for (size_t k = 0; k < m; k++) jobs[s.job_of_desc[k]].status = st[k];
            bool any_image_error = false;
            for (int v : st) any_image_error = any_image_error || v != 0;
            if (rc != B200JPG_OK && !any_image_error) {  // a device-level failure, not a per-image one
                s.rc = rc;
                for (size_t k = 0; k < m; k++) jobs[s.job_of_desc[k]].status = rc;
            }
        });
    }
    for (auto& s : slots) {
        if (s.gpu.joinable()) s.gpu.join();
        if (s.rc != B200JPG_OK) result = s.rc;
    }
    return result;
}

}  // extern "C"
