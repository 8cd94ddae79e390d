// stream_engine.h -- the host side of every host-fed path: a pool of host threads turns jobs into sparse block
// streams (sbs.h) inside their own page-locked rings (ring_book.h); the calling thread groups finished jobs and
// pushes them through the three-stream device pipeline (sbs_pipeline.h).  No barrier anywhere: producers keep
// going while earlier images upload, compute and download; a ring region is recycled when its upload is done.
// Two producers exist: JPEG files (Huffman decoding, files_api.cpp) and dense coefficient buffers (compaction,
// b200jpg_batch_run_host).  Not part of the C ABI.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <mutex>

#include "../../include/b200jpg.h"
#include "sbs_pipeline.h"

namespace b200jpg {

class JobSource {
public:
    virtual ~JobSource() {}
    virtual size_t size() const = 0;
    // Host thread, step 1: look at job i.  Returns its status; on B200JPG_OK with *need > 0 the engine reserves
    // *need bytes of page-locked memory and calls produce(); *need == 0 means the job is already finished (the
    // source handled it itself, e.g. an image too large for a ring).  *state travels to produce()/drop().
    virtual int prepare(size_t i, size_t* need, void** state, std::mutex* gpu_mu) = 0;
    // Host thread, step 2: write the stream into dst and fill item->{desc, len, order, out, out_cap}; frees *state.
    // Everything item->desc points to must stay valid until finish(): 512 bytes behind the stream (dst + len) are
    // reserved for quantisation tables that would otherwise die with *state.  dst == nullptr: only free *state.
    virtual int produce(size_t i, void* state, uint8_t* dst, size_t need, SbsItem* item) = 0;
    // Submitter thread: final status of a job that went through the device.
    virtual void finish(size_t i, int status) = 0;
    virtual const char* name() const = 0;
    // hint: the jobs' pixel buffers are device memory (no download to wait for: larger groups amortise better)
    virtual bool outputs_on_device() const { return false; }
};

// Runs every job of `src` with `nthreads` host threads (< 1: one per CPU this process may run on).
int stream_engine_run(b200jpg_ctx* ctx, JobSource& src, int nthreads);
int stream_engine_default_threads();

// b200jpg_batch_run_host with host-side compaction: dense coefficient buffers -> streams -> device.
int stream_engine_run_dense(b200jpg_ctx* ctx, const b200jpg_image_desc* imgs, size_t n, const int* plan_status, uint8_t* const* outs,
                            const size_t* out_caps, int* statuses, int nthreads);
// fraction of non-zero coefficients in a sample of the images' blocks (decides whether compaction pays)
double stream_engine_sample_density(const b200jpg_image_desc* imgs, size_t n);

}  // namespace b200jpg
