// sbs_pipeline.cpp -- see sbs_pipeline.h.
#include "sbs_pipeline.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "entropy_host.h"

namespace b200jpg {

static int ent_max_passes() {
    static const int v = [] {
        const char* e = getenv("B200JPG_ENT_PASSES");
        const int n = e ? atoi(e) : 0;
        // A launch settles everything inside its CTAs of 256 subsequences (up to 32 rounds); one more launch per CTA boundary
        // that a stretch of unsynchronised subsequences crosses.  Real files need 2 (tests/cpp/entropy_emul.cpp prints them);
        // a launch with nothing to do still costs ~2.5 us of a group's ~900.  Whatever 6 do not settle the write pass flags
        // and the host decodes.
        return n > 0 && n <= 1024 ? n : 6;
    }();
    return v;
}

static inline size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

SbsPipeline::SbsPipeline(b200jpg_ctx* ctx, int nslots) : ctx_(ctx) {
    if (cudaSetDevice(ctx->device) != cudaSuccess) return;
    if (cudaStreamCreateWithFlags(&s_in_, cudaStreamNonBlocking) != cudaSuccess) return;
    for (auto& sc : s_comp2_)
        if (cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking) != cudaSuccess) return;
    if (cudaStreamCreateWithFlags(&s_out_, cudaStreamNonBlocking) != cudaSuccess) return;
    // device buffers of the slots regrow while a run is in flight (group sizes depend on timing): stream-ordered
    // allocation from a pool that keeps what it is given back, so a regrowth costs microseconds instead of the
    // device-wide synchronisation of cudaFree / cudaMalloc (measured: 40-70 ms outliers of a 10 ms call at 2 GPUs)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        async_alloc_ = true;
    }
    cudaGetLastError();
    if (const char* e = getenv("B200JPG_COMP_STREAMS")) ncomp_streams_ = (unsigned)std::min(4, std::max(1, atoi(e)));
    if (const char* e = getenv("B200JPG_SLOTS")) nslots = std::min(8, std::max(2, atoi(e)));
    slots_.resize((size_t)std::max(2, nslots));
    timeline_ = getenv("B200JPG_TIMELINE") != nullptr;
    const unsigned ev_flags = timeline_ ? cudaEventDefault : cudaEventDisableTiming;
    for (auto& s : slots_) {
        if (cudaEventCreateWithFlags(&s.e_h2d, ev_flags) != cudaSuccess) return;
        if (cudaEventCreateWithFlags(&s.e_comp, ev_flags) != cudaSuccess) return;
        if (cudaEventCreateWithFlags(&s.e_done, ev_flags) != cudaSuccess) return;
        if (timeline_)
            for (auto& e : s.e_t0)
                if (cudaEventCreate(&e) != cudaSuccess) return;
    }
    if (timeline_ && cudaEventCreate(&e_base_) != cudaSuccess) return;
    ok_ = true;
}

SbsPipeline::~SbsPipeline() {
    cudaSetDevice(ctx_->device);
    drain();
    cudaDeviceSynchronize();
    for (auto& s : slots_) {
        cudaFree(s.d_streams.p);
        cudaFree(s.d_coefs.p);
        cudaFree(s.d_planes.p);
        cudaFree(s.d_out.p);
        cudaFree(s.d_tables.p);
        cudaFree(s.d_ent.p);
        if (s.h_tables.p) cudaFreeHost(s.h_tables.p);
        if (s.h_status.p) cudaFreeHost(s.h_status.p);
        if (s.e_h2d) cudaEventDestroy(s.e_h2d);
        if (s.e_comp) cudaEventDestroy(s.e_comp);
        if (s.e_done) cudaEventDestroy(s.e_done);
        for (auto& e : s.e_t0)
            if (e) cudaEventDestroy(e);
    }
    if (e_base_) cudaEventDestroy(e_base_);
    if (s_in_) cudaStreamDestroy(s_in_);
    for (auto& sc : s_comp2_)
        if (sc) cudaStreamDestroy(sc);
    if (s_out_) cudaStreamDestroy(s_out_);
}

int SbsPipeline::grow_device(Buf& b, size_t need, size_t hint) {
    // With a reservation the buffer goes to its full size the FIRST time the slot is used, whatever that group needs: a call
    // that switches from host to device outputs raises the group cap fourfold, and a buffer that only grew when a group
    // finally exceeded the old size stalled a call in mid-run (2 GPUs x 256 images per call: 15 instead of 60-97 GP/s).
    if (std::max(need, hint) <= b.cap) return B200JPG_OK;
    const double t0 = now_ms();
    struct Tally { SbsPipeline* p; double t0; ~Tally() { p->grow_ms += now_ms() - t0; p->grows++; } } tally{this, t0};
    // doubling, so that a buffer regrows a handful of times in its life.  The slot is idle (its previous group has
    // retired), and every stream that touches the buffer afterwards first waits for work enqueued on s_in_.
    // (group sizes depend on timing -- how many images were ready when the submitter came round -- so a new maximum can
    // turn up many calls into a run: twice the need, so that the buffers settle after two or three regrowths)
    const size_t want = up(need <= hint ? hint : std::max(2 * need, 2 * b.cap), 1 << 20);   // within the reservation: once and for all
    if (async_alloc_) {
        if (b.p) CU_TRY(ctx_, cudaFreeAsync(b.p, s_in_));
        b.p = nullptr;
        b.cap = 0;
        CU_TRY(ctx_, cudaMallocAsync(&b.p, want, s_in_));
    } else {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
        CU_TRY(ctx_, cudaMalloc(&b.p, want));
    }
    b.cap = want;
    return B200JPG_OK;
}
int SbsPipeline::grow_pinned(Buf& b, size_t need, size_t hint) {
    if (std::max(need, hint) <= b.cap) return B200JPG_OK;
    const double t0 = now_ms();
    struct Tally { SbsPipeline* p; double t0; ~Tally() { p->grow_ms += now_ms() - t0; p->grows++; } } tally{this, t0};
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = up(need <= hint ? hint : need + need / 4, 1 << 16);
    CU_TRY(ctx_, cudaHostAlloc(&b.p, want, cudaHostAllocDefault));
    b.cap = want;
    return B200JPG_OK;
}

// Fires the callbacks of a slot whose work has completed (or waits for it) and frees its plan.
void SbsPipeline::retire(Slot& s, bool wait) {
    if (!s.busy) return;
    if (!s.h2d_reported) {
        if (wait) cudaEventSynchronize(s.e_h2d);
        else if (cudaEventQuery(s.e_h2d) != cudaSuccess) return;
        s.h2d_reported = true;
        if (on_h2d) on_h2d(s.group);
    }
    if (wait) {
        if (cudaEventSynchronize(s.e_done) != cudaSuccess && error_ == B200JPG_OK)
            error_ = b200jpg_cuda_fail(ctx_, cudaGetLastError(), "sparse-stream pipeline");
    } else if (cudaEventQuery(s.e_done) != cudaSuccess) {
        return;
    }
    // images whose scan the device flagged (entropy_dev.h) go back to the caller for a host decode
    for (size_t k = 0; k < s.ent_items.size(); k++) {
        const unsigned* st = (const unsigned*)s.h_status.p + 2 * k;
        const size_t i = s.ent_items[k];
        if (s.group.statuses[i] == B200JPG_OK && (st[0] != 0 || st[1] != 1)) {
            s.group.statuses[i] = B200JPG_INTERNAL_RETRY_HOST;
            static const bool trace = getenv("B200JPG_TRACE") != nullptr;
            if (trace)
                fprintf(stderr, "[b200jpg] device entropy: job %zu (%ux%u) flagged 0x%x, completed %u -> host\n", s.group.items[i].job,
                        s.group.items[i].desc.width, s.group.items[i].desc.height, st[0], st[1]);
        }
    }
    if (timeline_) {
        float t[6] = {0, 0, 0, 0, 0, 0};
        cudaEvent_t ev[6] = {s.e_t0[0], s.e_h2d, s.e_t0[1], s.e_comp, s.e_t0[2], s.e_done};
        for (int k = 0; k < 6; k++) cudaEventElapsedTime(&t[k], e_base_, ev[k]);
        cudaGetLastError();
        fprintf(stderr, "[b200jpg] timeline: group of %3zu enqueued %7.2f | upload %7.2f-%7.2f | kernels %7.2f-%7.2f | download %7.2f-%7.2f | retired %7.2f\n",
                s.group.items.size(), s.t_enq - t_base_, t[0], t[1], t[2], t[3], t[4], t[5], now_ms() - t_base_);
    }
    if (on_done) on_done(s.group);
    if (s.batch) {
        batch_release_device(s.batch);
        delete s.batch;
        s.batch = nullptr;
    }
    s.group.items.clear();
    s.busy = false;
}

void SbsPipeline::poll() {
    while (oldest_ < next_) {
        Slot& s = slots_[oldest_ % slots_.size()];
        retire(s, false);
        if (s.busy) break;  // callbacks fire in submission order
        oldest_++;
    }
}

int SbsPipeline::drain() {
    while (oldest_ < next_) {
        retire(slots_[oldest_ % slots_.size()], true);
        oldest_++;
    }
    const int e = error_;
    error_ = B200JPG_OK;
    return e;
}

static void fill_k0(const b200jpg_image_desc& d, const ImageLayout& L, unsigned order, size_t stream_off, K0Image* k) {
    memset(k, 0, sizeof *k);
    k->stream_off = stream_off;
    k->order = order;
    unsigned nb = 0, j = 0;
    for (int c = 0; c < d.ncomp; c++) {
        k->slab_row[c] = (unsigned)(L.coef_off[c] / 128);
        k->block_w[c] = d.comps[c].block_w;
        k->first[c] = nb;
        k->h[c] = d.comps[c].h;
        k->v[c] = d.comps[c].v;
        nb += (unsigned)d.comps[c].block_w * d.comps[c].block_h;
        if (order & SBS_INTERLEAVED)
            for (unsigned vy = 0; vy < d.comps[c].v; vy++)
                for (unsigned hx = 0; hx < d.comps[c].h; hx++)
                    if (j < 12) {
                        k->mcu_comp[j] = (unsigned char)c;
                        k->mcu_hx[j] = (unsigned char)hx;
                        k->mcu_vy[j] = (unsigned char)vy;
                        j++;
                    }
    }
    for (int c = d.ncomp; c < 4; c++) k->first[c] = 0xffffffffu;  // never reached by the planar search
    k->nb = nb;
    k->bpm = j ? j : 1;
    k->mcu_w = d.comps[0].h ? d.comps[0].block_w / d.comps[0].h : 1;
    if (k->mcu_w == 0) k->mcu_w = 1;
}

void SbsPipeline::timeline_begin() {
    if (!timeline_ || !ok_) return;
    cudaSetDevice(ctx_->device);
    cudaEventRecord(e_base_, s_in_);
    cudaEventSynchronize(e_base_);
    t_base_ = now_ms();
}

int SbsPipeline::submit(std::vector<SbsItem>&& items) {
    if (!ok_) return b200jpg_fail(ctx_, B200JPG_ERR_INTERNAL, "sparse-stream pipeline could not be created");
    if (items.empty()) return B200JPG_OK;
    CU_TRY(ctx_, cudaSetDevice(ctx_->device));
    if (next_ - oldest_ >= slots_.size()) {  // the slot we are about to reuse is still in flight
        const double t_w = now_ms();
        struct Tally { SbsPipeline* p; double t0; ~Tally() { p->retire_wait_ms += now_ms() - t0; } } tally{this, t_w};
        retire(slots_[oldest_ % slots_.size()], true);
        oldest_++;
    }
    Slot& s = slots_[next_ % slots_.size()];
    s.group.items = std::move(items);
    const double t_enq = now_ms();
    const int rc = enqueue(s);
    enqueue_ms += now_ms() - t_enq;
    if (rc != B200JPG_OK) {  // leave the slot reusable: nothing of this group may still be running
        cudaStreamSynchronize(s_in_);
        for (auto& sc : s_comp2_) cudaStreamSynchronize(sc);
        cudaStreamSynchronize(s_out_);
        if (s.batch) {
            batch_release_device(s.batch);
            delete s.batch;
            s.batch = nullptr;
        }
        s.group.items.clear();
        return rc;
    }
    last_coefs_ = s.d_coefs.p;
    s.busy = true;
    s.h2d_reported = false;
    next_++;
    return B200JPG_OK;
}

int SbsPipeline::enqueue(Slot& s) {
    const size_t n = s.group.items.size();
    s.group.statuses.assign(n, B200JPG_OK);
    const std::vector<SbsItem>& it = s.group.items;

    // bound of the plan's tables (see batch_create_impl): per component 32 B + two tables, per 128 blocks 16 B, ...
    std::vector<b200jpg_image_desc> descs(n);
    size_t tbound = 10 * 256 + n * (sizeof(DevImage) + sizeof(K0Image) + 64), max_ent = 0;
    for (size_t i = 0; i < n; i++) {
        if (it[i].order == SBS_ENTROPY && it[i].stream && it[i].len >= sizeof(EntHeader)) {  // one descriptor per restart interval
            EntHeader h;
            memcpy(&h, it[i].stream, sizeof h);
            max_ent += std::min<size_t>(h.nintervals, (size_t)1 << 24);
        }
        descs[i] = it[i].desc;
        for (int c = 0; c < 4; c++) descs[i].coefs[c] = nullptr;
        for (int c = 0; c < descs[i].ncomp && c < 4; c++)
            tbound += sizeof(DevComp) + 96 * sizeof(unsigned) + ((size_t)descs[i].comps[c].block_w * descs[i].comps[c].block_h / K1_TILE + 1) * sizeof(DevTile);
        tbound += ((size_t)descs[i].width / 2048 + 1) * sizeof(K2Strip);
        tbound += ((size_t)descs[i].width / 960 + 2) * sizeof(FColumn);
    }
    tbound += max_ent * sizeof(EntImage) + n * sizeof(GatherItem) + 1024 + 256;
    // Pixel buffers in device memory (every image of the group): the kernels write them directly -- no pixel slab, no
    // device-to-device copy afterwards (unified addressing tells host from device pointers).
    bool device_outs = n > 0;
    std::vector<unsigned long long> out_addr(n, 0ull);
    for (size_t i = 0; i < n && device_outs; i++) {
        cudaPointerAttributes pa;
        if (!it[i].out || cudaPointerGetAttributes(&pa, it[i].out) != cudaSuccess || pa.type != cudaMemoryTypeDevice) device_outs = false;
        out_addr[i] = (unsigned long long)(uintptr_t)it[i].out;
    }
    cudaGetLastError();
    int rc = grow_device(s.d_tables, tbound, reserve_ ? (size_t)8 << 20 : 0);
    if (rc == B200JPG_OK) rc = grow_pinned(s.h_tables, tbound, reserve_ ? (size_t)8 << 20 : 0);
    if (rc) return rc;

    TableArena arena;
    arena.d = (char*)s.d_tables.p;
    arena.h = (char*)s.h_tables.p;
    arena.bytes = tbound - n * sizeof(K0Image) - max_ent * sizeof(EntImage) - n * sizeof(GatherItem) - 4 * 256;
    PlanOverrides ov;
    ov.arena = &arena;
    ov.upload_stream = s_in_;
    if (device_outs) ov.out_addr = out_addr.data();
    rc = batch_create_impl(ctx_, descs.data(), n, s.group.statuses.data(), ov, &s.batch);
    if (rc) return rc;
    b200jpg_batch* b = s.batch;

    // K0 / entropy descriptors + stream placement for the images the planner accepted
    const size_t k0_at = up(b->table_bytes, 256);
    K0Image* h_k0 = (K0Image*)(arena.h + k0_at);
    const K0Image* d_k0 = (const K0Image*)(arena.d + k0_at);
    const size_t ent_at = up(k0_at + n * sizeof(K0Image), 256);
    EntImage* h_ent = (EntImage*)(arena.h + ent_at);
    const EntImage* d_ent = (const EntImage*)(arena.d + ent_at);
    const size_t gat_at = up(ent_at + max_ent * sizeof(EntImage), 256);
    GatherItem* h_gat = (GatherItem*)(arena.h + gat_at);
    const GatherItem* d_gat = (const GatherItem*)(arena.d + gat_at);
    size_t nk0 = 0, nent = 0, ngat = 0, stream_bytes = 0;
    unsigned max_nb = 0, max_nsub = 0, total_sub = 0, max_comp_blocks = 0;
    std::vector<size_t> soff(n, 0);
    s.ent_items.clear();
    s.ent_images = 0;
    // the compact streams the device writes for entropy images live behind everything that is uploaded
    size_t cs_at = 0;
    for (size_t i = 0; i < n; i++)
        if (!s.group.statuses[i] && it[i].stream) cs_at += it[i].len;
    cs_at = up(cs_at, 256);
    for (size_t i = 0; i < n; i++) {
        if (s.group.statuses[i]) continue;
        if (!it[i].out || it[i].out_cap < b->layout[i].out_len || !it[i].stream) {
            s.group.statuses[i] = B200JPG_ERR_INTERNAL;
            b200jpg_fail(ctx_, B200JPG_ERR_INTERNAL, "output buffer too small");
            continue;
        }
        if (it[i].order == SBS_ENTROPY) {  // an entropy-coded scan: Huffman decoding happens on the device, interval by interval
            unsigned nsub = 0;
            const unsigned k = ent_fill_images(it[i].stream, it[i].len, descs[i], b->layout[i].coef_off, stream_bytes, cs_at, total_sub,
                                               h_ent + nent, max_ent - nent, &nsub);
            if (k == 0) {
                s.group.statuses[i] = B200JPG_ERR_INTERNAL;
                b200jpg_fail(ctx_, B200JPG_ERR_INTERNAL, "malformed entropy payload");
                continue;
            }
            for (unsigned r = 0; r < k; r++) {
                max_nsub = std::max(max_nsub, h_ent[nent + r].nsub);
                for (int c = 0; c < 4; c++) max_comp_blocks = std::max(max_comp_blocks, h_ent[nent + r].comp_blocks[c]);
                s.ent_items.push_back(i);
            }
            total_sub += nsub;
            nent += k;
            // K0 expands the compact stream like a host-made one (scan order = MCU order; a lone component: raster order)
            fill_k0(descs[i], b->layout[i], (descs[i].ncomp > 1 ? SBS_INTERLEAVED : SBS_PLANAR) | SBS_BLOCK_OFFSETS, cs_at, &h_k0[nk0]);
            max_nb = std::max(max_nb, h_k0[nk0].nb);
            cs_at += ent_cs_bytes(h_k0[nk0].nb);
            nk0++;
            s.ent_images++;
        } else {
            fill_k0(descs[i], b->layout[i], it[i].order, stream_bytes, &h_k0[nk0]);
            const SbsLayout lay = SbsLayout::make(h_k0[nk0].nb);
            if (it[i].len < lay.off_vals || it[i].len > lay.worst_bytes() || it[i].len % 16 != 0) {
                s.group.statuses[i] = B200JPG_ERR_INTERNAL;
                b200jpg_fail(ctx_, B200JPG_ERR_INTERNAL, "malformed sparse block stream");
                continue;
            }
            max_nb = std::max(max_nb, h_k0[nk0].nb);
            nk0++;
        }
        soff[i] = stream_bytes;
        stream_bytes += it[i].len;
    }
    const int passes = ent_max_passes();
    // reservation: one more image than the cap may join a group; planes are half the coefficient bytes, pixels at most as
    // many, the compact streams of device-decoded scans 140 B per block against 128 B dense, plus the uploaded payloads
    const size_t R = reserve_ ? reserve_ + reserve_ / 8 : 0;
    rc = grow_device(s.d_streams, std::max(stream_bytes, cs_at) + 256, R + R / 4 + (R ? (size_t)32 << 20 : 0));
    if (rc == B200JPG_OK) rc = grow_device(s.d_coefs, b->info.coef_bytes + K1_TILE * 128, R);
    if (rc == B200JPG_OK) rc = grow_device(s.d_planes, b->info.plane_bytes + 256, R / 2);
    if (rc == B200JPG_OK) rc = grow_device(s.d_out, b->info.out_bytes + 256, device_outs ? 0 : R);
    if (rc == B200JPG_OK && nent) rc = grow_device(s.d_ent, ent_work_bytes(total_sub, (unsigned)nent, max_comp_blocks, passes), R / 16);
    if (rc == B200JPG_OK && nent) rc = grow_pinned(s.h_status, nent * 8, reserve_ ? (size_t)1 << 20 : 0);
    if (rc) return rc;

    // copy-in
    s.t_enq = now_ms();
    if (timeline_) CU_TRY(ctx_, cudaEventRecord(s.e_t0[0], s_in_));
    // Streams in page-locked rings are uploaded by ONE kernel per group (k_gather_streams, k0_expand.cu: why); anything else
    // (b200jpg_decode_batch_sbs takes whatever memory the caller has) goes through the copy engine, runs merged.
    {
        const char* run_src = nullptr;
        size_t run_dst = 0, run_bytes = 0;
        for (size_t i = 0; i < n; i++) {
            if (s.group.statuses[i]) continue;
            const size_t dst = soff[i];
            const char* src = (const char*)it[i].stream;
            if (it[i].mapped && it[i].len % 16 == 0 && ((uintptr_t)src & 15u) == 0 && dst % 16 == 0) {
                h_gat[ngat].src = src;
                h_gat[ngat].dst_off = dst;
                h_gat[ngat].n16 = (unsigned)(it[i].len / 16);
                h_gat[ngat].pad_ = 0;
                ngat++;
                continue;
            }
            if (run_bytes && run_src + run_bytes == src && run_dst + run_bytes == dst) {
                run_bytes += it[i].len;
            } else {
                if (run_bytes) CU_TRY(ctx_, cudaMemcpyAsync((char*)s.d_streams.p + run_dst, run_src, run_bytes, cudaMemcpyHostToDevice, s_in_));
                run_src = src;
                run_dst = dst;
                run_bytes = it[i].len;
                h2d_copies++;
            }
        }
        if (run_bytes) CU_TRY(ctx_, cudaMemcpyAsync((char*)s.d_streams.p + run_dst, run_src, run_bytes, cudaMemcpyHostToDevice, s_in_));
    }
    // the K0, entropy and gather descriptors lie behind each other in the table arena: one copy
    if (nk0 || nent || ngat) {
        const size_t end = ngat ? gat_at + ngat * sizeof(GatherItem) : (nent ? ent_at + nent * sizeof(EntImage) : k0_at + nk0 * sizeof(K0Image));
        CU_TRY(ctx_, cudaMemcpyAsync((void*)d_k0, h_k0, end - k0_at, cudaMemcpyHostToDevice, s_in_));
    }
    if (ngat) {
        CU_TRY(ctx_, launch_gather_streams(d_gat, (unsigned)ngat, (uint8_t*)s.d_streams.p, s_in_));
        h2d_copies++;
        ctx_->launches++;
    }
    CU_TRY(ctx_, cudaEventRecord(s.e_h2d, s_in_));
    // compute
    // Pixels that leave over PCIe: one compute stream, first in first out, so that the oldest group finishes (and starts
    // downloading) as early as possible.  Pixels that stay in device memory: nothing downstream is waiting, and rotating
    // over four streams lets the following groups' kernels fill the SMs during the latency-bound synchronisation rounds.
    cudaStream_t s_comp_ = s_comp2_[device_outs ? next_ % ncomp_streams_ : 0];
    if (device_outs != last_device_outs_) {  // switching modes: do not let the two streams' groups race each other's events
        for (auto& sc : s_comp2_) CU_TRY(ctx_, cudaStreamSynchronize(sc));
        last_device_outs_ = device_outs;
    }
    CU_TRY(ctx_, cudaStreamWaitEvent(s_comp_, s.e_h2d, 0));
    if (timeline_) CU_TRY(ctx_, cudaEventRecord(s.e_t0[1], s_comp_));
    unsigned* d_status = nullptr;
    if (nent) {
        // Huffman decoding on the device: payloads -> compact streams (bitmaps are OR-ed into: zero them first) ...
        CU_TRY(ctx_, launch_k0_zero_headers(d_k0, (unsigned)nk0, max_nb, (uint8_t*)s.d_streams.p, s_comp_));
        uint64_t ent_launches = 0;
        CU_TRY(ctx_, launch_entropy(d_ent, (unsigned)nent, max_nsub, total_sub, max_comp_blocks, (uint8_t*)s.d_streams.p, s.d_ent.p, passes, &d_status,
                                    s_comp_, &ent_launches));
        ctx_->launches += ent_launches + 1;
    }
    // ... and K0 expands them, like the streams the host made: the dense slab K1 reads, zeros included
    if (nent) {
        CU_TRY(ctx_, launch_k0_expand_blocks(d_k0, (unsigned)nk0, max_nb, (const uint8_t*)s.d_streams.p, (short*)s.d_coefs.p, s_comp_));
        ctx_->launches++;
    }
    if (nk0 > s.ent_images) {
        CU_TRY(ctx_, launch_k0_expand(d_k0, (unsigned)nk0, max_nb, (const uint8_t*)s.d_streams.p, (short*)s.d_coefs.p, s_comp_));
        ctx_->launches++;
    }
    if (nk0 || nent) {
        rc = batch_launch(b, s.d_coefs.p, s.d_planes.p, device_outs ? nullptr : s.d_out.p, 3, 0, (unsigned)b->tiles.size(), 0, (unsigned)n, s_comp_);
        if (rc) return rc;
    }
    CU_TRY(ctx_, cudaEventRecord(s.e_comp, s_comp_));
    // copy-out (cudaMemcpyDefault: the callers' pixel buffers may be host OR device memory -- unified addressing tells)
    CU_TRY(ctx_, cudaStreamWaitEvent(s_out_, s.e_comp, 0));
    if (timeline_) CU_TRY(ctx_, cudaEventRecord(s.e_t0[2], s_out_));
    if (nent) CU_TRY(ctx_, cudaMemcpyAsync(s.h_status.p, d_status, nent * 8, cudaMemcpyDeviceToHost, s_out_));
    if (!device_outs) {
        char* out_dst = nullptr;
        size_t out_src = 0, out_bytes = 0;
        for (size_t i = 0; i < n; i++) {
            if (s.group.statuses[i]) continue;
            const ImageLayout& L = b->layout[i];
            if (out_bytes && out_dst + out_bytes == (char*)it[i].out && out_src + out_bytes == L.out_off) {
                out_bytes += L.out_len;
            } else {
                if (out_bytes) CU_TRY(ctx_, cudaMemcpyAsync(out_dst, (char*)s.d_out.p + out_src, out_bytes, cudaMemcpyDefault, s_out_));
                out_dst = (char*)it[i].out;
                out_src = L.out_off;
                out_bytes = L.out_len;
                d2h_copies++;
            }
        }
        if (out_bytes) CU_TRY(ctx_, cudaMemcpyAsync(out_dst, (char*)s.d_out.p + out_src, out_bytes, cudaMemcpyDefault, s_out_));
    }
    CU_TRY(ctx_, cudaEventRecord(s.e_done, s_out_));
    return B200JPG_OK;
}

}  // namespace b200jpg
