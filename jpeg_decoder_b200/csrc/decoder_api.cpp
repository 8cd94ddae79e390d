// decoder_api.cpp -- C ABI of the whole-file decoder (include/b200jpg.h, "Whole-file decoder"):
// Decoder::{new, read_info, info, scale, decode, icc_profile, exif_data, xmp_data,
// set_color_transform, set_max_decoding_buffer_size}, reference src/decoder.rs:132-295.
// Host half: HostDecoder (host_decoder.cpp).  Worker half: the GPU batch path (pipeline.cu).
#include <string.h>

#include <string>
#include <vector>

#include <memory>

#include "../../include/b200jpg.h"
#include "context.h"
#include "host_decoder.h"

using b200jpg::HostDecoder;

struct b200jpg_decoder {
    b200jpg_ctx* ctx;
    const uint8_t* data;
    size_t len;
    HostDecoder host;
    std::unique_ptr<HostDecoder> again;  // a second, complete host decode when the device route hands the image back
    std::vector<uint8_t> pixels;
    std::vector<uint8_t> icc;
    std::string err;
    b200jpg_decoder(b200jpg_ctx* c, const uint8_t* d, size_t n) : ctx(c), data(d), len(n), host(d, n) {}
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
    d->pixels.assign((size_t)desc.width * desc.height * desc.ncomp, 0);
    uint8_t* out = d->pixels.data();
    size_t cap = d->pixels.size();
    int status = B200JPG_OK;
    rc = b200jpg_decode_batch(d->ctx, &desc, 1, &out, &cap, &status);
    if (rc == B200JPG_OK) rc = status;
    if (rc) {
        d->err = b200jpg_last_error(d->ctx);
        return rc;
    }
    *pixels = d->pixels.data();
    *len = d->pixels.size();
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
        d->pixels.resize((size_t)f.output_w * f.output_h * f.comps.size());  // every byte is written by the download
        b200jpg_file_job job;
        memset(&job, 0, sizeof job);
        job.data = d->data;
        job.len = d->len;
        job.out = d->pixels.data();
        job.out_cap = d->pixels.size();
        rc = b200jpg_decode_files(d->ctx, &job, 1, 1);
        if (rc == B200JPG_OK && job.status == B200JPG_OK && job.out_len == d->pixels.size()) {
            *pixels = d->pixels.data();
            *len = d->pixels.size();
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
