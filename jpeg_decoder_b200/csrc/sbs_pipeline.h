// sbs_pipeline.h -- the sparse-stream host pipeline: groups of images whose coefficients arrive as sparse block
// streams (sbs.h) flow through three CUDA streams,
//     copy-in : H2D of the streams + the plan's tables          (event h2d)
//     compute : K0 expand / device entropy decoding -> K1 dequant+IDCT -> K2 colour   (event comp)
//     copy-out: D2H of the pixels into the callers' buffers      (event done)
// so that group k+1 uploads while group k computes and group k-1 downloads.  A ring of slots owns the device
// buffers (grow-only: nothing is cudaMalloc'ed or cudaFree'd in steady state, both would serialise the device).
// Single-threaded: one submitter thread drives submit()/poll()/drain().  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <functional>
#include <vector>

#include "../../include/b200jpg.h"
#include "batch_internal.h"
#include "sbs.h"

namespace b200jpg {

// positive, internal (never leaves the library): the device flagged this image's entropy-coded scan; decode it on the host
enum { B200JPG_INTERNAL_RETRY_HOST = 2 };

struct SbsItem {
    b200jpg_image_desc desc;  // geometry, tables, colour transform; desc.coefs is ignored
    const uint8_t* stream = nullptr;  // host (ideally page-locked) sparse stream, valid until on_h2d
    size_t len = 0;
    unsigned order = SBS_PLANAR;  // SBS_ENTROPY: `stream` is an entropy payload (entropy_dev.h), not a block stream
    uint8_t* out = nullptr;  // host pixels (ideally page-locked)
    size_t out_cap = 0;
    size_t job = 0;        // caller's cookie
    int thread = 0;        // caller's cookie
    uint64_t ring_end = 0; // caller's cookie
    bool mapped = false;   // `stream` is page-locked memory the device can read in place (cudaHostAlloc): it may be uploaded by kernel
};

class SbsPipeline {
public:
    struct Group {
        std::vector<SbsItem> items;
        std::vector<int> statuses;
    };
    SbsPipeline(b200jpg_ctx* ctx, int nslots);
    ~SbsPipeline();
    bool ok() const { return ok_; }
    // called from poll()/drain()/submit() on the submitter thread, in submission order
    std::function<void(const Group&)> on_h2d;   // the streams of this group have left host memory
    std::function<void(const Group&)> on_done;  // pixels are in the callers' buffers, statuses are final

    int submit(std::vector<SbsItem>&& items);  // B200JPG_OK or a device-level error
    void poll();                               // fire callbacks of whatever has completed
    int drain();                               // wait for everything in flight
    // optional: keep dense slab copies for debugging (tests): after drain(), the last group's coefficient slab
    const void* last_coef_slab() const { return last_coefs_; }
    // groups submitted and not yet retired (as of the last poll)
    size_t in_flight() const { return next_ - oldest_; }

    // The caller's bound on the dense coefficient bytes of one group (the stream engine's group cap): slot buffers are sized
    // for it the first time they are needed, so that nothing regrows in steady state.  A regrowth of a few hundred MB costs
    // 15-50 ms even from the stream-ordered allocator (measured: profiles/r02_files_regrowth_trace.txt); group sizes depend on
    // timing, so without the reservation a new maximum -- and a stall -- can turn up many calls into a run.
    void reserve_for_group_bytes(size_t coef_bytes) { reserve_ = coef_bytes; }

    // trace counters (B200JPG_TRACE): time and count of buffer regrowths, time inside enqueue, time waiting for a slot
    double grow_ms = 0, enqueue_ms = 0, retire_wait_ms = 0;
    // B200JPG_TIMELINE=1: one stderr line per group with the device times of its upload, kernels and download (ms since the
    // call began) -- the poor man's timeline where no system profiler is installed
    void timeline_begin();
    unsigned grows = 0, h2d_copies = 0, d2h_copies = 0;  // copies = cudaMemcpyAsync calls for streams / pixels

private:
    struct Buf {
        void* p = nullptr;
        size_t cap = 0;
    };
    struct Slot {
        Buf d_streams, d_coefs, d_planes, d_out, d_tables, h_tables, d_ent, h_status;
        std::vector<size_t> ent_items;  // group index of the image of every device-decoded restart interval
        size_t ent_images = 0;          // images of the group whose scan is decoded on the device
        cudaEvent_t e_h2d = nullptr, e_comp = nullptr, e_done = nullptr;
        cudaEvent_t e_t0[3] = {nullptr, nullptr, nullptr};  // B200JPG_TIMELINE: when the upload / the kernels / the download of the group began
        double t_enq = 0;  // host clock when the group was handed to the streams
        bool busy = false, h2d_reported = false;
        b200jpg_batch* batch = nullptr;
        Group group;
    };
    int grow_device(Buf& b, size_t need, size_t hint = 0);
    int grow_pinned(Buf& b, size_t need, size_t hint = 0);
    void retire(Slot& s, bool wait);
    int enqueue(Slot& s);

    b200jpg_ctx* ctx_;
    bool ok_ = false;
    // up to four compute streams, taken in turn by the groups: the synchronisation rounds of device entropy decoding are latency-bound
    // (a few lanes busy per round), so the next group's throughput-bound kernels fill the machine meanwhile
    cudaStream_t s_in_ = nullptr, s_comp2_[4] = {nullptr, nullptr, nullptr, nullptr}, s_out_ = nullptr;
    unsigned ncomp_streams_ = 4;  // compute streams device-output groups rotate over (B200JPG_COMP_STREAMS, 1..4; measured 53.5 / 59.8 / 64.3 / 65.7 GP/s: profiles/r02_files_streams_ab.jsonl)
    std::vector<Slot> slots_;
    size_t next_ = 0, oldest_ = 0;  // tickets: slot = ticket % nslots
    bool last_device_outs_ = false;
    size_t reserve_ = 0;
    bool timeline_ = false;
    cudaEvent_t e_base_ = nullptr;
    double t_base_ = 0;
    bool async_alloc_ = false;  // slot buffers come from the stream-ordered allocator (cudaMallocAsync on s_in_)
    const void* last_coefs_ = nullptr;
    int error_ = B200JPG_OK;
};

}  // namespace b200jpg
