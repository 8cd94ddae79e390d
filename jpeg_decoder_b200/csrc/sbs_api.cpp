// sbs_api.cpp -- C ABI around the sparse block stream (sbs.h): the host half of a decode that ends in a stream
// instead of dense buffers, and the batch entry point that takes streams (H2D -> K0 -> K1 -> K2 -> D2H).
#include <cuda_runtime.h>
#include <string.h>

#include <memory>
#include <mutex>
#include <vector>

#include "../../include/b200jpg.h"
#include "batch_internal.h"
#include "sbs_pipeline.h"

using namespace b200jpg;

namespace {

void pipeline_free(void* p) { delete (SbsPipeline*)p; }

SbsPipeline* get_pipeline(b200jpg_ctx* ctx) {
    if (!ctx->sbs_pipeline) {
        SbsPipeline* p = new SbsPipeline(ctx, 2);
        if (!p->ok()) {
            delete p;
            return nullptr;
        }
        ctx->sbs_pipeline = p;
        ctx->sbs_pipeline_free = pipeline_free;
    }
    return (SbsPipeline*)ctx->sbs_pipeline;
}

// A stream from outside the library is untrusted: K0 follows its offsets, so they must be exactly what the
// bitmaps imply and stay inside the buffer.
bool sbs_valid(const uint8_t* s, size_t len, size_t nb) {
    const SbsLayout lay = SbsLayout::make(nb);
    if (!s || nb == 0 || len < lay.off_vals || len % 16 != 0) return false;
    const uint64_t* bm = (const uint64_t*)s;
    const uint32_t* voff = (const uint32_t*)(s + lay.off_voff);
    uint64_t at = 0;
    for (size_t g = 0; g < lay.nb_pad / 32; g++) {
        if (voff[g] != at) return false;
        for (size_t t = 32 * g; t < 32 * g + 32; t++) {
            if (t >= nb && bm[t] != 0) return false;
            at += (uint64_t)__builtin_popcountll(bm[t] >> 1) << (bm[t] & 1);
        }
        if (at > 0xffffffffull) return false;
    }
    return lay.off_vals + at <= len;
}

size_t desc_blocks(const b200jpg_image_desc& d) {
    size_t nb = 0;
    for (int c = 0; c < d.ncomp && c < 4; c++) nb += (size_t)d.comps[c].block_w * d.comps[c].block_h;
    return nb;
}

}  // namespace

extern "C" {

size_t b200jpg_sbs_worst_bytes(size_t nblocks) { return SbsLayout::make(nblocks).worst_bytes(); }

int b200jpg_sbs_from_dense(const b200jpg_image_desc* img, uint8_t* buf, size_t cap, b200jpg_sbs_stream* stream) {
    if (!img || !buf || !stream || img->ncomp < 1 || img->ncomp > 4) return B200JPG_ERR_INTERNAL;
    const size_t nb = desc_blocks(*img);
    if (nb == 0 || cap < SbsLayout::make(nb).worst_bytes()) return B200JPG_ERR_INTERNAL;
    for (int c = 0; c < img->ncomp; c++)
        if (!img->coefs[c]) return B200JPG_ERR_FORMAT;
    SbsWriter w;
    w.begin(buf, nb);
    for (int c = 0; c < img->ncomp; c++) {
        const size_t cnt = (size_t)img->comps[c].block_w * img->comps[c].block_h;
        w.put_dense_natural_run(img->coefs[c], cnt);
    }
    stream->data = buf;
    stream->len = w.finish();
    stream->order = (int)(SBS_PLANAR | SBS_NATURAL);
    return B200JPG_OK;
}

}  // extern "C"

// K0 places block j of MCU (mx, my) of an interleaved stream at row (my*v+vy)*block_w + mx*h+hx of its component: that
// lands inside the component's slab only when every block grid is exactly mcu_w*h x mcu_h*v (parser.rs:306-307) and an MCU
// has at most 12 blocks (the descriptor's table size).
static bool interleaved_geometry_ok(const b200jpg_image_desc& d) {
    if (d.ncomp < 1 || d.ncomp > 4) return true;  // rejected by the planner with its own message
    unsigned bpm = 0, mcu_w = 0, mcu_h = 0;
    for (int c = 0; c < d.ncomp; c++) {
        const b200jpg_component& k = d.comps[c];
        if (k.h == 0 || k.v == 0 || k.block_w % k.h != 0 || k.block_h % k.v != 0) return false;
        const unsigned mw = k.block_w / k.h, mh = k.block_h / k.v;
        if (c == 0) { mcu_w = mw; mcu_h = mh; }
        else if (mw != mcu_w || mh != mcu_h) return false;
        bpm += (unsigned)k.h * k.v;
    }
    return bpm <= 12 && mcu_w > 0 && mcu_h > 0;
}

extern "C" {

int b200jpg_decode_batch_sbs(b200jpg_ctx* ctx, const b200jpg_image_desc* imgs, const b200jpg_sbs_stream* streams, size_t n,
                             uint8_t* const* outs, const size_t* out_caps, int* statuses) {
    if (!ctx || (n && (!imgs || !streams || !outs || !out_caps))) return B200JPG_ERR_INTERNAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    SbsPipeline* pipe = get_pipeline(ctx);
    if (!pipe) return b200jpg_fail(ctx, B200JPG_ERR_INTERNAL, "internal: could not create the device pipeline");
    int result = B200JPG_OK;
    std::vector<int> local(n, B200JPG_OK);
    pipe->on_h2d = nullptr;
    pipe->on_done = [&](const SbsPipeline::Group& g) {
        for (size_t k = 0; k < g.items.size(); k++) local[g.items[k].job] = g.statuses[k];
    };
    const size_t group_max = 32;
    for (size_t i0 = 0; i0 < n; i0 += group_max) {
        std::vector<SbsItem> items;
        for (size_t i = i0; i < n && i < i0 + group_max; i++) {
            if (streams[i].order < 0 || streams[i].order > (int)(SBS_INTERLEAVED | SBS_NATURAL)) {
                local[i] = B200JPG_ERR_INTERNAL;
                continue;
            }
            if ((streams[i].order & SBS_INTERLEAVED) && !interleaved_geometry_ok(imgs[i])) {
                local[i] = b200jpg_fail(ctx, B200JPG_ERR_INTERNAL, "interleaved stream: block grids are not mcu_w*h x mcu_h*v, or more than 12 blocks per MCU");
                continue;
            }
            if (imgs[i].ncomp >= 1 && imgs[i].ncomp <= 4 && !sbs_valid(streams[i].data, streams[i].len, desc_blocks(imgs[i]))) {
                local[i] = b200jpg_fail(ctx, B200JPG_ERR_INTERNAL, "malformed sparse block stream");
                continue;
            }
            SbsItem it;
            it.desc = imgs[i];
            it.stream = streams[i].data;
            it.len = streams[i].len;
            it.order = (unsigned)streams[i].order;
            it.out = outs[i];
            it.out_cap = out_caps[i];
            it.job = i;
            items.push_back(it);
        }
        const int rc = pipe->submit(std::move(items));
        if (rc != B200JPG_OK) {
            result = rc;
            for (size_t i = i0; i < n && i < i0 + group_max; i++)
                if (local[i] == B200JPG_OK) local[i] = rc;
        }
    }
    const int rc = pipe->drain();
    pipe->on_done = nullptr;
    if (rc != B200JPG_OK) result = rc;
    for (size_t i = 0; i < n; i++) {
        if (statuses) statuses[i] = local[i];
        if (result == B200JPG_OK && local[i] != B200JPG_OK) result = local[i];
    }
    return result;
}

int b200jpg_debug_expand_sbs(b200jpg_ctx* ctx, const b200jpg_image_desc* img, const b200jpg_sbs_stream* stream,
                             int16_t* const dense_out[4]) {
    if (!ctx || !img || !stream || !dense_out) return B200JPG_ERR_INTERNAL;
    const size_t npx = (size_t)img->width * img->height * img->ncomp;
    std::vector<uint8_t> px(npx ? npx : 1);
    uint8_t* out = px.data();
    size_t cap = px.size();
    int st = B200JPG_OK;
    int rc = b200jpg_decode_batch_sbs(ctx, img, stream, 1, &out, &cap, &st);
    if (rc != B200JPG_OK) return rc;
    // the planner is deterministic: a second plan of the same image has the same slab offsets
    b200jpg_batch* b = nullptr;
    rc = b200jpg_batch_create(ctx, img, 1, nullptr, &b);
    if (rc != B200JPG_OK) return rc;
    size_t coef_off[4];
    b200jpg_batch_image_layout(b, 0, coef_off, nullptr, nullptr, nullptr);
    b200jpg_batch_free(b);
    std::lock_guard<std::mutex> lock(ctx->mu);
    SbsPipeline* pipe = get_pipeline(ctx);
    if (!pipe || !pipe->last_coef_slab()) return B200JPG_ERR_INTERNAL;
    for (int c = 0; c < img->ncomp && c < 4; c++) {
        const size_t bytes = (size_t)img->comps[c].block_w * img->comps[c].block_h * 128;
        if (dense_out[c])
            CU_TRY(ctx, cudaMemcpy(dense_out[c], (const char*)pipe->last_coef_slab() + coef_off[c], bytes, cudaMemcpyDeviceToHost));
    }
    return B200JPG_OK;
}

}  // extern "C"
