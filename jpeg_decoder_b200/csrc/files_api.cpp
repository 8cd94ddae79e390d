// files_api.cpp -- whole-file batch decoding: n JPEG files in, n pixel buffers out.
//
// What a user of the reference gets from an outer `par_iter` over `Decoder::decode()` (one image per host
// thread), re-cut for the GPU worker.  Host threads do only what is inherently serial per image
// -- marker parsing and Huffman decoding (HostDecoder, csrc/host_decoder.cpp) -- and write what they decode as
// a sparse block stream (sbs.h); the stream engine (stream_engine.h) moves the streams through the device
// pipeline H2D | K0+K1+K2 | D2H while the host threads are already on the next images.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/b200jpg.h"
#include "context.h"
#include "entropy_host.h"
#include "host_decoder.h"
#include "stream_engine.h"

using b200jpg::HostDecoder;
using b200jpg::SbsItem;
using b200jpg::SbsLayout;

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

void fill_desc(const HostDecoder& hd, b200jpg_image_desc* d) {
    memset(d, 0, sizeof *d);
    const auto& f = hd.frame();
    d->width = f.output_w;
    d->height = f.output_h;
    d->ncomp = (uint8_t)f.comps.size();
    d->color_transform = (uint8_t)hd.determine_color_transform();
    for (size_t k = 0; k < f.comps.size() && k < 4; k++) {
        d->comps[k] = f.comps[k];
        d->qt[k] = hd.component_qtable((int)k);
    }
}

constexpr size_t kMaxSbsImage = (size_t)1 << 30;  // larger images take the dense path, one at a time

class FileSource : public b200jpg::JobSource {
public:
    // index: optional subset of jobs (the second wave: images the device sent back); device_entropy: let qualifying
    // scans be Huffman-decoded on the GPU
    FileSource(b200jpg_ctx* ctx, b200jpg_file_job* jobs, size_t n, const std::vector<size_t>* index, bool device_entropy)
        : ctx_(ctx), jobs_(jobs), n_(index ? index->size() : n), index_(index), device_entropy_(device_entropy) {}
    size_t size() const override { return n_; }
    std::vector<size_t> take_retries() {
        std::lock_guard<std::mutex> g(mu_);
        return std::move(retry_);
    }
    size_t device_scans() const { return device_scans_.load(); }
    bool outputs_on_device() const override {
        if (n_ == 0) return false;
        const b200jpg_file_job& job = jobs_[index_ ? (*index_)[0] : 0];
        cudaPointerAttributes pa;
        const bool dev = job.out && cudaPointerGetAttributes(&pa, job.out) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        return dev;
    }
    const char* name() const override { return "decode_files"; }

    int prepare(size_t i, size_t* need, void** state, std::mutex* gpu_mu) override {
        *need = 0;
        *state = nullptr;
        b200jpg_file_job& job = jobs_[index_ ? (*index_)[i] : i];
        std::unique_ptr<HostDecoder> hd(new HostDecoder(job.data, job.len));
        job.status = hd->read_info();
        job.out_len = 0;
        if (job.status != B200JPG_OK) return job.status;
        fill_info(*hd, &job);
        if (!job.out || job.out_cap < job.out_len) return job.status = B200JPG_ERR_INTERNAL;
        const size_t worst = SbsLayout::make(hd->total_blocks()).worst_bytes();
        if (worst > kMaxSbsImage) {
            // dense path, synchronously (the image alone is a GPU-sized batch)
            job.status = hd->entropy_decode();
            if (job.status != B200JPG_OK) return job.status;
            b200jpg_image_desc d;
            fill_desc(*hd, &d);
            for (int k = 0; k < d.ncomp; k++) {
                if (!hd->component_has_data(k)) return job.status = B200JPG_ERR_FORMAT;
                d.coefs[k] = hd->coefficients(k);
            }
            std::lock_guard<std::mutex> g(*gpu_mu);
            int st = B200JPG_OK;
            const int rc = b200jpg_decode_batch(ctx_, &d, 1, &job.out, &job.out_cap, &st);
            job.status = rc != B200JPG_OK && st == B200JPG_OK ? rc : st;
            return job.status;
        }
        *need = device_entropy_ ? std::max(worst, b200jpg::ent_payload_bound(job.len)) : worst;
        *state = hd.release();
        return B200JPG_OK;
    }

    int produce(size_t i, void* state, uint8_t* dst, size_t need, SbsItem* item) override {
        std::unique_ptr<HostDecoder> hd((HostDecoder*)state);
        b200jpg_file_job& job = jobs_[index_ ? (*index_)[i] : i];
        if (!dst) return job.status = B200JPG_ERR_INTERNAL;
        hd->set_sbs_sink(dst);
        hd->probe_device_scan(device_entropy_);
        job.status = hd->entropy_decode();
        if (job.status == b200jpg::B200JPG_INTERNAL_DEVICE_SCAN) {
            // a complete baseline scan: ship the entropy-coded bytes, the GPU does the Huffman decoding
            const size_t len = b200jpg::ent_build_payload(*hd, job.data, job.len, dst, need);
            if (len) {
                job.status = B200JPG_OK;
                fill_desc(*hd, &item->desc);
                item->len = len;
                item->order = b200jpg::SBS_ENTROPY;
                item->out = job.out;
                item->out_cap = job.out_cap;
                uint16_t* qcopy = (uint16_t*)(dst + item->len);
                for (int k = 0; k < item->desc.ncomp && k < 4; k++) {
                    memcpy(qcopy + 64 * k, item->desc.qt[k], 128);
                    item->desc.qt[k] = qcopy + 64 * k;
                }
                device_scans_++;
                return B200JPG_OK;
            }
            // markers inside the scan, a truncated file, ...: the host loop mirrors the reference there
            hd.reset(new HostDecoder(job.data, job.len));
            hd->set_sbs_sink(dst);
            job.status = hd->entropy_decode();
        }
        if (job.status != B200JPG_OK) return job.status;
        bool complete = hd->sbs_length() != 0;
        for (size_t k = 0; k < hd->frame().comps.size(); k++) complete = complete && hd->component_has_data((int)k);
        if (!complete) return job.status = B200JPG_ERR_FORMAT;  // "not all components have data", src/decoder.rs:1306-1308
        fill_desc(*hd, &item->desc);
        item->len = hd->sbs_length();
        item->order = hd->sbs_order();
        item->out = job.out;
        item->out_cap = job.out_cap;
        // the quantisation tables die with *hd: they ride behind the stream
        uint16_t* qcopy = (uint16_t*)(dst + item->len);
        for (int k = 0; k < item->desc.ncomp && k < 4; k++) {
            memcpy(qcopy + 64 * k, item->desc.qt[k], 128);
            item->desc.qt[k] = qcopy + 64 * k;
        }
        return B200JPG_OK;
    }
    void finish(size_t i, int status) override {
        const size_t j = index_ ? (*index_)[i] : i;
        if (status == b200jpg::B200JPG_INTERNAL_RETRY_HOST) {  // the device flagged the scan: second wave, host Huffman
            std::lock_guard<std::mutex> g(mu_);
            retry_.push_back(j);
            jobs_[j].status = B200JPG_ERR_INTERNAL;  // overwritten by the second wave
            return;
        }
        jobs_[j].status = status;
    }

private:
    b200jpg_ctx* ctx_;
    b200jpg_file_job* jobs_;
    size_t n_;
    const std::vector<size_t>* index_;
    bool device_entropy_;
    std::mutex mu_;
    std::vector<size_t> retry_;
    std::atomic<size_t> device_scans_{0};
};

}  // namespace

extern "C" {

int b200jpg_read_info_files(b200jpg_file_job* jobs, size_t n, int nthreads) {
    if (!jobs && n) return B200JPG_ERR_INTERNAL;
    if (nthreads < 1) nthreads = b200jpg::stream_engine_default_threads();
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
    const bool device = b200jpg_device_entropy_enabled(ctx);
    FileSource src(ctx, jobs, n, nullptr, device);
    int rc = b200jpg::stream_engine_run(ctx, src, nthreads);
    const std::vector<size_t> retry = src.take_retries();
    {
        std::lock_guard<std::mutex> g(ctx->mu);  // calls on one context may come from several threads
        ctx->device_scans += src.device_scans();
        ctx->device_scan_retries += retry.size();
    }
    if (!retry.empty()) {  // scans the device flagged (malformed streams, or no convergence): the reference's own loop decides
        FileSource again(ctx, jobs, n, &retry, false);
        const int rc2 = b200jpg::stream_engine_run(ctx, again, nthreads);
        if (rc == B200JPG_OK) rc = rc2;
    }
    return rc;
}

}  // extern "C"
