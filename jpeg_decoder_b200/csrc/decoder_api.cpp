// decoder_api.cpp -- C ABI of the whole-file decoder (include/b200jpg.h, "Whole-file decoder"):
// Decoder::{new, read_info, info, scale, decode, icc_profile, exif_data, xmp_data,
// set_color_transform, set_max_decoding_buffer_size}, reference src/decoder.rs:132-295.
// Host half: HostDecoder (host_decoder.cpp).  Worker half: the GPU batch path (pipeline.cu).
#include <string.h>

#include <string>
#include <vector>

#include <memory>
#include <mutex>

#include "../../include/b200jpg.h"
#include "context.h"
#include "host_decoder.h"

using b200jpg::HostDecoder;

// Pixel buffers of decoders: page-locked (the download runs at link speed instead of through the driver's staging
// buffer) and recycled through a small process-wide pool -- cudaHostAlloc costs about a millisecond, a 1080p decode two.
namespace {
struct PinnedPool {
    struct Buf {
        uint8_t* p;
        size_t cap;
    };
    std::mutex mu;
    std::vector<Buf> free_list;
    static constexpr size_t kMaxBuffers = 8, kMaxBytes = (size_t)512 << 20;
    uint8_t* get(size_t need, size_t* cap) {
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = free_list.size();
            for (size_t i = 0; i < free_list.size(); i++)
                if (free_list[i].cap >= need && (best == free_list.size() || free_list[i].cap < free_list[best].cap)) best = i;
            if (best != free_list.size()) {
                const Buf b = free_list[best];
                free_list.erase(free_list.begin() + (long)best);
                *cap = b.cap;
                return b.p;
            }
        }
        size_t want = (size_t)1 << 16;
        while (want < need) want <<= 1;
        if (want - need > need / 2) want = (need + 65535) / 65536 * 65536;  // large images: not the next power of two
        void* p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        *cap = want;
        return (uint8_t*)p;
    }
    void put(uint8_t* p, size_t cap) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> g(mu);
            size_t total = cap;
            for (const Buf& b : free_list) total += b.cap;
            if (free_list.size() < kMaxBuffers && total <= kMaxBytes) {
                free_list.push_back(Buf{p, cap});
                return;
            }
        }
        cudaFreeHost(p);
    }
};
PinnedPool g_pixel_pool;
}  // namespace

struct b200jpg_decoder {
    b200jpg_ctx* ctx;
    const uint8_t* data;
    size_t len;
    HostDecoder host;
    std::unique_ptr<HostDecoder> again;  // a second, complete host decode when the device route hands the image back
    uint8_t* pixels = nullptr;  // page-locked, from g_pixel_pool
    size_t pixels_cap = 0, pixels_len = 0;
    std::vector<uint8_t> icc;
    std::string err;
    b200jpg_decoder(b200jpg_ctx* c, const uint8_t* d, size_t n) : ctx(c), data(d), len(n), host(d, n) {}
    ~b200jpg_decoder() { g_pixel_pool.put(pixels, pixels_cap); }
    // a buffer of at least n bytes (contents undefined); false when page-locked memory cannot be had
    bool reserve_pixels(size_t n) {
        if (pixels_cap < n || !pixels) {
            g_pixel_pool.put(pixels, pixels_cap);
            pixels = g_pixel_pool.get(n ? n : 1, &pixels_cap);
            if (!pixels) pixels_cap = 0;
        }
        pixels_len = pixels ? n : 0;
        return pixels != nullptr;
    }
};

// Fills `desc` with what decode_planes hands to compute_image (src/decoder.rs:617-696).
static int fill_desc(b200jpg_decoder* d, const HostDecoder& host, b200jpg_image_desc* desc) {
    const auto& f = host.frame();
    memset(desc, 0, sizeof *desc);
    desc->width = f.output_w;
    desc->height = f.output_h;
    desc->ncomp = (uint8_t)f.comps.size();
    desc->color_transform = (uint8_t)host.determine_color_transform();
    for (size_t i = 0; i < f.comps.size() && i < 4; i++) {
        // "not all components have data", src/decoder.rs:1306-1308
        if (!host.component_has_data((int)i)) {
            d->err = "invalid JPEG format: not all components have data";
            return B200JPG_ERR_FORMAT;
        }
        desc->comps[i] = f.comps[i];
        desc->qt[i] = host.component_qtable((int)i);
        desc->coefs[i] = host.coefficients((int)i);
    }
    return B200JPG_OK;
}

extern "C" {

int b200jpg_decoder_new(b200jpg_ctx* ctx, const uint8_t* data, size_t len, b200jpg_decoder** d) {
    if (!d || (!data && len)) return B200JPG_ERR_INTERNAL;
    *d = new b200jpg_decoder(ctx, data, len);
    return B200JPG_OK;
}
void b200jpg_decoder_free(b200jpg_decoder* d) { delete d; }
void b200jpg_decoder_set_color_transform(b200jpg_decoder* d, int ct) {
    if (d) d->host.set_color_transform(ct);
}
void b200jpg_decoder_set_max_decoding_buffer_size(b200jpg_decoder* d, size_t max) {
    if (d) d->host.set_max_decoding_buffer_size(max);
}
int b200jpg_decoder_read_info(b200jpg_decoder* d) {
    if (!d) return B200JPG_ERR_INTERNAL;
    int rc = d->host.read_info();
    if (rc) d->err = d->host.error();
    return rc;
}
int b200jpg_decoder_info(const b200jpg_decoder* d, b200jpg_image_info* info) {
    if (!d || !info || !d->host.has_frame()) return 0;
    const auto& f = d->host.frame();
    info->width = f.output_w;
    info->height = f.output_h;
    info->pixel_format = d->host.pixel_format();
    info->coding_process = f.coding_process;
    return 1;
}
int b200jpg_decoder_scale(b200jpg_decoder* d, uint16_t req_w, uint16_t req_h, uint16_t* w, uint16_t* h) {
    if (!d || !w || !h) return B200JPG_ERR_INTERNAL;
    int rc = d->host.scale(req_w, req_h, w, h);
    if (rc) d->err = d->host.error();
    return rc;
}
// the host decoder to run a complete entropy decode on: d->host, unless decode() already stopped it at the scan to send
// the scan to the GPU (then a fresh one)
static HostDecoder& full_host(b200jpg_decoder* d) {
    if (!d->host.device_scan().eligible) return d->host;
    d->again.reset(new HostDecoder(d->data, d->len));
    return *d->again;
}

int b200jpg_decoder_entropy_decode(b200jpg_decoder* d, b200jpg_image_desc* desc) {
    if (!d || !desc) return B200JPG_ERR_INTERNAL;
    HostDecoder& host = full_host(d);
    int rc = host.entropy_decode();
    if (rc) {
        d->err = host.error();
        return rc;
    }
    return fill_desc(d, host, desc);
}
int b200jpg_decoder_total_blocks(b200jpg_decoder* d, size_t* nblocks) {
    if (!d || !nblocks) return B200JPG_ERR_INTERNAL;
    const int rc = d->host.read_info();
    if (rc) {
        d->err = d->host.error();
        return rc;
    }
    *nblocks = d->host.total_blocks();
    return B200JPG_OK;
}
int b200jpg_decoder_entropy_decode_sbs(b200jpg_decoder* d, uint8_t* buf, size_t cap, b200jpg_image_desc* desc,
                                       b200jpg_sbs_stream* stream) {
    if (!d || !buf || !desc || !stream) return B200JPG_ERR_INTERNAL;
    int rc = d->host.read_info();
    if (rc) {
        d->err = d->host.error();
        return rc;
    }
    if (cap < b200jpg_sbs_worst_bytes(d->host.total_blocks())) {
        d->err = "internal: sparse block stream buffer is smaller than b200jpg_sbs_worst_bytes()";
        return B200JPG_ERR_INTERNAL;
    }
    HostDecoder& host = full_host(d);
    if (&host != &d->host) {
        rc = host.read_info();
        if (rc) {
            d->err = host.error();
            return rc;
        }
    }
    host.set_sbs_sink(buf);
    rc = host.entropy_decode();
    if (rc) {
        d->err = host.error();
        return rc;
    }
    rc = fill_desc(d, host, desc);
    if (rc) return rc;
    for (int i = 0; i < 4; i++) desc->coefs[i] = nullptr;  // the coefficients are in the stream
    stream->data = buf;
    stream->len = host.sbs_length();
    stream->order = (int)host.sbs_order();
    return B200JPG_OK;
}
// Worker half of decode(): the dense coefficients of `host` through the GPU batch path.
static int run_worker_path(b200jpg_decoder* d, const HostDecoder& host, const uint8_t** pixels, size_t* len) {
    b200jpg_image_desc desc;
    int rc = fill_desc(d, host, &desc);
    if (rc) return rc;
    if (!d->reserve_pixels((size_t)desc.width * desc.height * desc.ncomp)) {
        d->err = "internal: no page-locked memory for the pixels";
        return B200JPG_ERR_INTERNAL;
    }
    uint8_t* out = d->pixels;
    size_t cap = d->pixels_len;
    int status = B200JPG_OK;
    rc = b200jpg_decode_batch(d->ctx, &desc, 1, &out, &cap, &status);
    if (rc == B200JPG_OK) rc = status;
    if (rc) {
        d->err = b200jpg_last_error(d->ctx);
        return rc;
    }
    *pixels = d->pixels;
    *len = d->pixels_len;
    return B200JPG_OK;
}

// Decoder::decode, src/decoder.rs:292-295.  A complete baseline scan (entropy_dev.h) does not wait for the host's
// sequential Huffman loop -- 6.6 ms for a 1080p image: the host parses the markers up to the scan (metadata, tables),
// the scan itself goes through the whole-file engine and is Huffman-decoded on the GPU (< 1 ms).  Everything else, and
// anything the device hands back, takes the host loop, whose pixels and errors are the reference's.
int b200jpg_decoder_decode(b200jpg_decoder* d, const uint8_t** pixels, size_t* len) {
    if (!d || !pixels || !len) return B200JPG_ERR_INTERNAL;
    if (!d->ctx) {  // no CPU fallback for the worker path
        const int rc0 = d->host.entropy_decode();
        if (rc0) {
            d->err = d->host.error();
            return rc0;
        }
        d->err = "internal: decoder was created without a device context; the worker path only exists on the GPU";
        return B200JPG_ERR_INTERNAL;
    }
    const bool device_route = b200jpg_device_entropy_enabled(d->ctx) && d->host.default_config();
    d->host.probe_device_scan(device_route);
    int rc = d->host.device_scan().eligible ? (int)b200jpg::B200JPG_INTERNAL_DEVICE_SCAN : d->host.entropy_decode();
    if (rc == b200jpg::B200JPG_INTERNAL_DEVICE_SCAN) {
        const auto& f = d->host.frame();
        if (!d->reserve_pixels((size_t)f.output_w * f.output_h * f.comps.size())) {
            d->err = "internal: no page-locked memory for the pixels";
            return B200JPG_ERR_INTERNAL;
        }
        b200jpg_file_job job;
        memset(&job, 0, sizeof job);
        job.data = d->data;
        job.len = d->len;
        job.out = d->pixels;
        job.out_cap = d->pixels_len;
        rc = b200jpg_decode_files(d->ctx, &job, 1, 1);
        if (rc == B200JPG_OK && job.status == B200JPG_OK && job.out_len == d->pixels_len) {
            *pixels = d->pixels;
            *len = d->pixels_len;
            return B200JPG_OK;
        }
        // an error, or a partial image: the complete host decode words it (and decides, should the two ever differ)
        d->again.reset(new HostDecoder(d->data, d->len));
        rc = d->again->entropy_decode();
        if (rc) {
            d->err = d->again->error();
            return rc;
        }
        return run_worker_path(d, *d->again, pixels, len);
    }
    if (rc) {
        d->err = d->host.error();
        return rc;
    }
    return run_worker_path(d, d->host, pixels, len);
}
const char* b200jpg_decoder_error(const b200jpg_decoder* d) { return d ? d->err.c_str() : "no decoder"; }
int b200jpg_decoder_icc_profile(b200jpg_decoder* d, const uint8_t** data, size_t* len) {
    if (!d || !data || !len || !d->host.icc_profile(&d->icc)) return 0;
    *data = d->icc.data();
    *len = d->icc.size();
    return 1;
}
int b200jpg_decoder_exif_data(const b200jpg_decoder* d, const uint8_t** data, size_t* len) {
    if (!d || !data || !len || !d->host.exif()) return 0;
    *data = d->host.exif()->data();
    *len = d->host.exif()->size();
    return 1;
}
int b200jpg_decoder_xmp_data(const b200jpg_decoder* d, const uint8_t** data, size_t* len) {
    if (!d || !data || !len || !d->host.xmp()) return 0;
    *data = d->host.xmp()->data();
    *len = d->host.xmp()->size();
    return 1;
}

}  // extern "C"
