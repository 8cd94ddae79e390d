// pipeline.cu -- host side of the C ABI (include/b200jpg.h): context, batch planner, worker-shaped
// API and the host<->device pipelines around the kernels of k1_idct.cu / k2_color.cu.
//
// Reference interfaces mirrored here: trait Worker (src/worker/mod.rs:24-35) and its immediate
// implementation (src/worker/immediate.rs), compute_image (src/decoder.rs:1300-1336),
// choose_color_convert_func (src/decoder.rs:1339-1389), Upsampler::new / choose_upsampler
// (src/upsampler.rs:20-45, 76-105), update_component_sizes (src/parser.rs:292-310).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "batch_internal.h"
#include "stream_engine.h"

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
int b200jpg_fail(b200jpg_ctx* ctx, int code, const std::string& msg) {
    if (ctx) {
        std::lock_guard<std::mutex> lock(ctx->err_mu);
        ctx->err = msg;
    }
    return code;
}
int b200jpg_cuda_fail(b200jpg_ctx* ctx, cudaError_t e, const char* what) {
    return b200jpg_fail(ctx, B200JPG_ERR_INTERNAL, std::string(what) + ": " + cudaGetErrorString(e));
}
static inline int fail(b200jpg_ctx* ctx, int code, const std::string& msg) { return b200jpg_fail(ctx, code, msg); }
static inline int cuda_fail(b200jpg_ctx* ctx, cudaError_t e, const char* what) { return b200jpg_cuda_fail(ctx, e, what); }

extern "C" {

void b200jpg_default_options(b200jpg_options* opt) {
    memset(opt, 0, sizeof *opt);
    opt->device = 0;
    opt->arith = B200JPG_ARITH_SCALAR;
}

const char* b200jpg_version(void) { return "b200jpg 0.1 (sm_100a)"; }

int b200jpg_create(const b200jpg_options* opt, b200jpg_ctx** out) {
    if (!out) return B200JPG_ERR_INTERNAL;
    *out = nullptr;
    b200jpg_options o;
    if (opt) o = *opt; else b200jpg_default_options(&o);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || o.device < 0 || o.device >= ndev) return B200JPG_ERR_INTERNAL;  // no CPU fallback
    b200jpg_ctx* ctx = new b200jpg_ctx();
    ctx->device = o.device;
    ctx->arith = o.arith;
    ctx->k1_kernel = o.k1_kernel;
    ctx->k2_kernel = o.k2_kernel;
    ctx->host_compact = o.host_compact;
    ctx->host_threads = o.host_threads;
    ctx->entropy = o.entropy;
    ctx->fuse = o.fuse;
    if (cudaSetDevice(o.device) != cudaSuccess) { delete ctx; return B200JPG_ERR_INTERNAL; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, o.device) != cudaSuccess) { delete ctx; return B200JPG_ERR_INTERNAL; }
    ctx->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {  // kernels are built for sm_100a only
        delete ctx;
        return B200JPG_ERR_INTERNAL;
    }
    if (o.stream) {
        ctx->stream = (cudaStream_t)o.stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return B200JPG_ERR_INTERNAL; }
        ctx->own_stream = true;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return B200JPG_ERR_INTERNAL; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        ctx->encode = (PFN_tensorMapEncodeTiled)fn;
    *out = ctx;
    return B200JPG_OK;
}

void b200jpg_destroy(b200jpg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->files_engine && ctx->files_engine_free) ctx->files_engine_free(ctx->files_engine);
    if (ctx->sbs_pipeline && ctx->sbs_pipeline_free) ctx->sbs_pipeline_free(ctx->sbs_pipeline);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    for (auto& sc : ctx->scratch) cudaFree(sc.p);
    delete ctx;
}
const char* b200jpg_last_error(const b200jpg_ctx* ctx) {
    if (!ctx) return "no context";
    // a per-thread copy: the pointer stays valid while other threads keep failing into ctx->err
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lock(const_cast<b200jpg_ctx*>(ctx)->err_mu);
    copy = ctx->err;
    return copy.c_str();
}
uint64_t b200jpg_launch_count(const b200jpg_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
void b200jpg_device_scan_counts(const b200jpg_ctx* ctx, uint64_t* decoded, uint64_t* retried) {
    if (decoded) *decoded = ctx ? ctx->device_scans : 0;
    if (retried) *retried = ctx ? ctx->device_scan_retries : 0;
}
void b200jpg_set_fuse(b200jpg_ctx* ctx, int fuse) {
    if (ctx) ctx->fuse = fuse;
}
int b200jpg_synchronize(b200jpg_ctx* ctx) {
    if (!ctx) return B200JPG_ERR_INTERNAL;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream2));
    return B200JPG_OK;
}

void b200jpg_debug_set_kernel_modes(int k1_mode, int k2_mode) {
    g_k1_mode = k1_mode;
    g_k2_mode = k2_mode;
}

void* b200jpg_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void b200jpg_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// src/parser.rs:283-310
static int ceil_div_u16(uint32_t x, uint32_t y, uint16_t* r) {
    if (x == 0 || y == 0) return B200JPG_ERR_FORMAT;
    *r = (uint16_t)(1 + ((x - 1) / y));
    return B200JPG_OK;
}
int b200jpg_update_component_sizes(uint16_t width, uint16_t height, b200jpg_component* comps, int ncomp,
                                   uint16_t* mcu_w, uint16_t* mcu_h) {
    if (!comps || ncomp <= 0) return B200JPG_ERR_FORMAT;
    uint32_t h_max = 0, v_max = 0;
    for (int i = 0; i < ncomp; i++) {
        h_max = std::max<uint32_t>(h_max, comps[i].h);
        v_max = std::max<uint32_t>(v_max, comps[i].v);
    }
    uint16_t mw, mh;
    if (ceil_div_u16(width, h_max * 8, &mw) || ceil_div_u16(height, v_max * 8, &mh)) return B200JPG_ERR_FORMAT;
    for (int i = 0; i < ncomp; i++) {
        b200jpg_component* c = &comps[i];
        if (ceil_div_u16((uint32_t)width * c->h * c->dct_scale, h_max * 8, &c->size_w)) return B200JPG_ERR_FORMAT;
        if (ceil_div_u16((uint32_t)height * c->v * c->dct_scale, v_max * 8, &c->size_h)) return B200JPG_ERR_FORMAT;
        c->block_w = (uint16_t)(mw * c->h);
        c->block_h = (uint16_t)(mh * c->v);
    }
    if (mcu_w) *mcu_w = mw;
    if (mcu_h) *mcu_h = mh;
    return B200JPG_OK;
}
// src/idct.rs:14-28
int b200jpg_choose_idct_size(uint16_t full_w, uint16_t full_h, uint16_t req_w, uint16_t req_h) {
    static const uint32_t scales[3] = {1, 2, 4};
    for (int k = 0; k < 3; k++) {
        uint16_t sw = (uint16_t)(((uint32_t)full_w * scales[k] - 1) / 8 + 1);
        uint16_t sh = (uint16_t)(((uint32_t)full_h * scales[k] - 1) / 8 + 1);
        if (sw >= req_w || sh >= req_h) return (int)scales[k];
    }
    return 8;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// batch plan
// ---------------------------------------------------------------------------------------------
// choose_upsampler, src/upsampler.rs:76-105
static int choose_upsampler(uint8_t h, uint8_t v, uint8_t hmax, uint8_t vmax, uint16_t out_w, uint16_t out_h,
                            DevUpComp* u, std::string* err) {
    bool h1 = h == hmax || out_w == 1, v1 = v == vmax || out_h == 1;
    bool h2 = h * 2 == hmax, v2 = v * 2 == vmax;
    u->hs = u->vs = 1;
    if (h1 && v1) u->kind = UP_H1V1;
    else if (h2 && v1) u->kind = UP_H2V1;
    else if (h1 && v2) u->kind = UP_H1V2;
    else if (h2 && v2) u->kind = UP_H2V2;
    else if (h == 0 || v == 0 || hmax % h != 0 || vmax % v != 0) {
        *err = "unsupported JPEG feature: NonIntegerSubsamplingRatio";
        return B200JPG_ERR_UNSUPPORTED;
    } else {
        u->kind = UP_GENERIC;
        u->hs = hmax / h;
        u->vs = vmax / v;
    }
    return B200JPG_OK;
}

// choose_color_convert_func, src/decoder.rs:1339-1389
static int choose_color_convert(int ncomp, int ct, unsigned* cc, std::string* err) {
    auto bad = [&](const char* what) {
        *err = std::string("invalid JPEG format: Invalid number of channels (") + (ncomp == 3 ? "3" : "4") + ") for " + what + " data";
        return B200JPG_ERR_FORMAT;
    };
    if (ncomp != 3 && ncomp != 4) {
        *err = "internal: component count must be 1, 3 or 4 (the reference panics)";
        return B200JPG_ERR_INTERNAL;
    }
    switch (ct) {
    case B200JPG_CT_NONE: *cc = CC_NOCONVERT; return B200JPG_OK;
    case B200JPG_CT_GRAYSCALE: return bad("Grayscale");
    case B200JPG_CT_RGB: if (ncomp == 3) { *cc = CC_RGB; return B200JPG_OK; } return bad("RGB");
    case B200JPG_CT_YCBCR: if (ncomp == 3) { *cc = CC_YCBCR; return B200JPG_OK; } return bad("YCbCr");
    case B200JPG_CT_CMYK: if (ncomp == 4) { *cc = CC_CMYK; return B200JPG_OK; } return bad("CMYK");
    case B200JPG_CT_YCCK: if (ncomp == 4) { *cc = CC_YCCK; return B200JPG_OK; } return bad("YCCK");
    case B200JPG_CT_JCS_BG_YCC: *err = "unsupported JPEG feature: ColorTransform(JcsBgYcc)"; return B200JPG_ERR_UNSUPPORTED;
    case B200JPG_CT_JCS_BG_RGB: *err = "unsupported JPEG feature: ColorTransform(JcsBgRgb)"; return B200JPG_ERR_UNSUPPORTED;
    default: *err = "invalid JPEG format: Unknown colour transform"; return B200JPG_ERR_FORMAT;
    }
}

// Validates one image and fills its DevImage (without offsets).
static int plan_image(const b200jpg_ctx* ctx, const b200jpg_image_desc& d, DevImage* img, std::string* err) {
    memset(img, 0, sizeof *img);
    const int n = d.ncomp;
    if (n != 1 && n != 3 && n != 4) {
        *err = "internal: component count must be 1, 3 or 4";
        return B200JPG_ERR_INTERNAL;
    }
    if (d.width == 0 || d.height == 0) {
        *err = "invalid JPEG format: invalid dimensions";
        return B200JPG_ERR_FORMAT;
    }
    for (int i = 0; i < n; i++) {
        const b200jpg_component& c = d.comps[i];
        if (!(c.dct_scale == 1 || c.dct_scale == 2 || c.dct_scale == 4 || c.dct_scale == 8)) {
            *err = "internal: Unsupported IDCT scale";  // src/idct.rs:237
            return B200JPG_ERR_INTERNAL;
        }
        if (c.block_w == 0 || c.block_h == 0 || c.h == 0 || c.v == 0 || c.h > 4 || c.v > 4 || c.size_w == 0 || c.size_h == 0) {
            *err = "internal: component geometry not initialised";
            return B200JPG_ERR_INTERNAL;
        }
        if ((size_t)c.size_w > (size_t)c.block_w * c.dct_scale || (size_t)c.size_h > (size_t)c.block_h * c.dct_scale) {
            *err = "internal: component size exceeds its block grid";
            return B200JPG_ERR_INTERNAL;
        }
    }
    img->width = d.width;
    img->height = d.height;
    img->ncomp = (unsigned)n;
    uint8_t hmax = 0, vmax = 0;
    size_t wmax = 0;
    for (int i = 0; i < n; i++) {
        hmax = std::max(hmax, d.comps[i].h);
        vmax = std::max(vmax, d.comps[i].v);
        wmax = std::max<size_t>(wmax, d.comps[i].size_w);
    }
    if (n == 1) {
        // src/decoder.rs:1310-1332: crop only, the colour transform is not consulted
        img->cc = CC_GRAY;
        img->c[0].kind = UP_H1V1;
        img->c[0].hs = img->c[0].vs = 1;
        // output is component.size, which must agree with the requested output size
        if (d.comps[0].size_w != d.width || d.comps[0].size_h != d.height) {
            *err = "internal: single-component output size differs from component.size";
            return B200JPG_ERR_INTERNAL;
        }
    } else {
        int rc = choose_color_convert(n, d.color_transform, &img->cc, err);
        if (rc) return rc;
        for (int i = 0; i < n; i++) {
            rc = choose_upsampler(d.comps[i].h, d.comps[i].v, hmax, vmax, d.width, d.height, &img->c[i], err);
            if (rc) return rc;
        }
        if (img->cc == CC_NOCONVERT && wmax * hmax != d.width) {
            // color_no_convert unwrap()s past the end of the row when the line buffers are longer
            // than the row (src/decoder.rs:1476-1484)
            *err = "internal: ColorTransform::None needs line buffers of exactly the row width (the reference panics)";
            return B200JPG_ERR_INTERNAL;
        }
    }
    for (int i = 0; i < n; i++) {
        const b200jpg_component& c = d.comps[i];
        DevUpComp& u = img->c[i];
        u.stride = (unsigned)c.block_w * c.dct_scale;
        u.in_w = c.size_w;
        u.in_h = c.size_h;
        const size_t rows = (size_t)c.block_h * c.dct_scale;
        // every sample the upsampler can touch must exist (the reference would panic on the slice)
        bool ok = true;
        switch (u.kind) {
        case UP_H1V1: ok = d.width <= u.stride && d.height <= rows; break;
        case UP_H2V1: ok = (size_t)u.in_w * 2 >= d.width && d.height <= rows; break;
        case UP_H1V2: ok = d.width <= u.stride && (size_t)u.in_h * 2 >= d.height; break;
        case UP_H2V2: ok = (size_t)u.in_w * 2 >= d.width && (size_t)u.in_h * 2 >= d.height; break;
        default: ok = (size_t)u.in_w * u.hs >= d.width && ((size_t)d.height + u.vs - 1) / u.vs <= rows; break;
        }
        if (!ok) {
            *err = "internal: output size is not covered by the component planes";
            return B200JPG_ERR_INTERNAL;
        }
    }
    // SSSE3 colour path: first (W/8 - 1) * 8 pixels of each row, src/arch/ssse3.rs:206, only for YCbCr
    img->ssse3_pixels = 0;
    if (ctx->arith == B200JPG_ARITH_SSSE3 && img->cc == CC_YCBCR) {
        unsigned nv = d.width / 8;
        img->ssse3_pixels = nv ? (nv - 1) * 8 : 0;
    }
    // kernel choice
    img->path = K2_PATH_GENERIC;
    if (ctx->k2_kernel != B200JPG_KERNEL_GENERIC && img->cc == CC_GRAY && img->c[0].stride % 8 == 0) img->path = K2_PATH_GRAY;
    if (ctx->k2_kernel != B200JPG_KERNEL_GENERIC && img->cc == CC_YCBCR) {
        const DevUpComp* u = img->c;
        const unsigned groups = (d.width + 15u) / 16u;
        if (u[0].kind == UP_H1V1 && u[1].kind == UP_H2V2 && u[2].kind == UP_H2V2 && u[0].stride % 16 == 0 &&
            u[1].stride % 8 == 0 && u[2].stride % 8 == 0 && u[1].in_w == u[2].in_w && u[1].in_h == u[2].in_h &&
            u[1].in_w == (d.width + 1u) / 2u && groups * 16u <= u[0].stride && groups * 8u <= u[1].stride &&
            groups * 8u <= u[2].stride)
            img->path = u[1].in_w % 8u == 0 ? K2_PATH_420 : K2_PATH_420R;
        else if (u[0].kind == UP_H1V1 && u[1].kind == UP_H1V1 && u[2].kind == UP_H1V1 && u[0].stride % 8 == 0 &&
                 u[1].stride % 8 == 0 && u[2].stride % 8 == 0)
            img->path = K2_PATH_444;
        else if (u[0].kind == UP_H1V1 && u[1].kind == UP_H2V1 && u[2].kind == UP_H2V1 && u[0].stride % 8 == 0 &&
                 u[1].stride % 8 == 0 && u[2].stride % 8 == 0 && u[1].in_w == u[2].in_w && u[1].in_w == (d.width + 1u) / 2u &&
                 groups * 16u <= ((u[0].stride + 15u) & ~15u) && groups * 8u <= u[1].stride && groups * 8u <= u[2].stride)
            img->path = K2_PATH_422;
        else if (u[0].kind == UP_H1V1 && u[1].kind == UP_H1V2 && u[2].kind == UP_H1V2 && u[0].stride % 8 == 0 && u[1].stride % 8 == 0 &&
                 u[2].stride % 8 == 0 && u[1].stride >= d.width && u[2].stride >= d.width &&
                 (d.height + 1u) / 2u <= u[1].in_h && (d.height + 1u) / 2u <= u[2].in_h)
            img->path = K2_PATH_440;
    }
    // every component at full resolution, bytes only: RGB / CMYK / YCCK / None
    if (ctx->k2_kernel != B200JPG_KERNEL_GENERIC && (img->cc == CC_RGB || img->cc == CC_CMYK || img->cc == CC_YCCK || img->cc == CC_NOCONVERT)) {
        bool ok = true;
        for (int i = 0; i < n; i++) ok = ok && img->c[i].kind == UP_H1V1 && img->c[i].stride % 8 == 0 && img->c[i].stride >= d.width;
        if (ok) img->path = K2_PATH_BYTES;
    }
    return B200JPG_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused kernel KF: can this image take it, and if so its column strips (device_types.h: FColumn / FRun).
// Eligible: 3-component YCbCr in scalar arithmetic at full IDCT size whose K2 path is the 4:2:0 or the 4:4:4 one,
// with the standard block grids (luma 2x2 blocks per MCU over 1x1 chroma, or 1x1 throughout).
// ---------------------------------------------------------------------------------------------
// which sampling modes the fused kernel takes: bit m set = KF_MODE_m
static unsigned fuse_modes(const b200jpg_ctx* ctx) {
    int f = ctx->fuse;
    if (const char* e = getenv("B200JPG_FUSE")) f = atoi(e) == 0 ? B200JPG_FUSE_OFF : (atoi(e) == 2 ? B200JPG_FUSE_AUTO : B200JPG_FUSE_ON);  // profiling
    if (f == B200JPG_FUSE_OFF) return 0u;
    if (f == B200JPG_FUSE_ON) return (1u << KF_MODE_444) | (1u << KF_MODE_420);
    return 1u << KF_MODE_444;
}

static int fused_mode_of(const b200jpg_ctx* ctx, const b200jpg_image_desc& d, const DevImage& img) {
    if (ctx->arith != B200JPG_ARITH_SCALAR || ctx->k1_kernel == B200JPG_KERNEL_GENERIC || ctx->k2_kernel == B200JPG_KERNEL_GENERIC) return -1;
    if (d.ncomp != 3 || img.cc != CC_YCBCR || img.ssse3_pixels != 0) return -1;
    for (int k = 0; k < 3; k++)
        if (d.comps[k].dct_scale != 8) return -1;
    const b200jpg_component *y = &d.comps[0], *cb = &d.comps[1], *cr = &d.comps[2];
    if (cb->block_w != cr->block_w || cb->block_h != cr->block_h || cb->size_w != cr->size_w || cb->size_h != cr->size_h) return -1;
    if (img.path == K2_PATH_444) {
        if (y->h != 1 || y->v != 1 || cb->h != 1 || cb->v != 1 || cr->h != 1 || cr->v != 1) return -1;
        if (y->block_w != cb->block_w || y->block_h != cb->block_h) return -1;
        if ((unsigned)y->block_w * 8u < d.width || (unsigned)y->block_h != (d.height + 7u) / 8u) return -1;
        return KF_MODE_444;
    }
    if (img.path == K2_PATH_420 || img.path == K2_PATH_420R || img.path == K2_PATH_420T) {
        if (y->h != 2 || y->v != 2 || cb->h != 1 || cb->v != 1 || cr->h != 1 || cr->v != 1) return -1;
        if (y->block_w != 2 * cb->block_w || y->block_h != 2 * cb->block_h) return -1;
        if ((unsigned)cb->block_w * 16u < d.width || (unsigned)cb->block_h != (d.height + 15u) / 16u) return -1;
        if (cb->size_w != (d.width + 1u) / 2u || cb->size_h != (d.height + 1u) / 2u) return -1;
        return KF_MODE_420;
    }
    return -1;
}

static void plan_fused_columns(b200jpg_batch* b, unsigned image, int mode, const b200jpg_image_desc& d, const ImageLayout& L, unsigned comp0) {
    const unsigned mcu_px = mode == KF_MODE_420 ? 16u : 8u;
    const unsigned mcu_w = d.comps[1].block_w, nrows = d.comps[1].block_h;
    const unsigned ms_max = kf_strip_px() / mcu_px;
    // only the MCUs that hold visible pixels (block grids may be wider than the image)
    const unsigned mcu_vis = (d.width + mcu_px - 1u) / mcu_px;
    const unsigned nstrips = (mcu_vis + ms_max - 1u) / ms_max;
    const unsigned ms_each = (mcu_vis + nstrips - 1u) / nstrips;
    const unsigned slab[3] = {(unsigned)(L.coef_off[0] / 128), (unsigned)(L.coef_off[1] / 128), (unsigned)(L.coef_off[2] / 128)};
    for (unsigned m0 = 0; m0 < mcu_vis; m0 += ms_each) {
        const unsigned ms = std::min(ms_each, mcu_vis - m0);
        FColumn c;
        memset(&c, 0, sizeof c);
        c.image = image;
        c.comp0 = comp0;
        c.first_item = b->fitems[mode];
        c.nrows = nrows;
        c.x0 = m0 * mcu_px;
        c.wpx = std::min(ms * mcu_px, (unsigned)d.width - c.x0);
        c.ngroups = (c.wpx + 15u) / 16u;
        c.gmagic = c.ngroups > 1 ? (unsigned)((0x100000000ull + c.ngroups - 1u) / c.ngroups) : 0u;
        unsigned nr = 0, box = 0;
        auto add_run = [&](unsigned comp, unsigned row0, unsigned step, unsigned len, unsigned wrap, unsigned dst_x, unsigned dst_row) {
            FRun& r = c.run[nr++];
            r.slab_row0 = row0; r.step = step; r.len = len; r.wrap = wrap; r.comp = comp; r.dst_x = dst_x; r.dst_row = dst_row; r.box0 = box;
            box += (len + 31u) / 32u;
        };
        if (mode == KF_MODE_420) {
            const unsigned ybw = d.comps[0].block_w, cbw = mcu_w;
            if (m0 == 0 && ms == (unsigned)cbw) {  // the strip spans whole block rows: the two luma block rows are contiguous
                add_run(0, slab[0], 2u * ybw, 2u * ybw, ybw, 0, 0);
            } else {
                add_run(0, slab[0] + 2u * m0, 2u * ybw, 2u * ms, 2u * ms, 0, 0);
                add_run(0, slab[0] + ybw + 2u * m0, 2u * ybw, 2u * ms, 2u * ms, 0, 8);
            }
            const unsigned hl = m0 > 0 ? 1u : 0u, hr = m0 + ms < cbw ? 1u : 0u;  // horizontal halo of the triangle filter
            c.cx_base = 8u * (m0 - hl);
            add_run(1, slab[1] + m0 - hl, cbw, ms + hl + hr, ms + hl + hr, 16, 0);
            add_run(2, slab[2] + m0 - hl, cbw, ms + hl + hr, ms + hl + hr, 16, 0);
            b->f_ystride[mode] = std::max(b->f_ystride[mode], (16u * ms + 15u) & ~15u);
            b->f_cstride[mode] = std::max(b->f_cstride[mode], ((16u + 8u * (ms + hl + hr) + 15u) & ~15u) + 16u);
        } else {
            for (unsigned k = 0; k < 3; k++) add_run(k, slab[k] + m0, d.comps[k].block_w, ms, ms, 0, 0);
            b->f_ystride[mode] = std::max(b->f_ystride[mode], (8u * ms + 15u) & ~15u);
        }
        c.nruns = nr;
        c.nboxes = box;
        b->fcols[mode].push_back(c);
        b->fitems[mode] += nrows;
    }
}

void batch_release_device(b200jpg_batch* b) {
    if (!b) return;
    if (!b->tables_borrowed) {
        cudaFree(b->d_comps);
        cudaFree(b->d_tiles);
        cudaFree(b->d_images);
        cudaFree(b->d_qtabs);
        cudaFree(b->d_qpack);
        cudaFree(b->d_strips);
        for (auto& f : b->d_fcols) cudaFree(f);
    }
    if (b->slabs_borrowed) {
        std::lock_guard<std::mutex> lock(b->ctx->mu);
        b->ctx->scratch_busy = false;
        b->slabs_borrowed = false;
    } else {
        cudaFree(b->d_coefs);
        cudaFree(b->d_planes);
        cudaFree(b->d_out);
    }
    b->d_comps = nullptr; b->d_tiles = nullptr; b->d_images = nullptr; b->d_qtabs = nullptr; b->d_qpack = nullptr; b->d_strips = nullptr; b->d_fcols[0] = b->d_fcols[1] = nullptr;
    b->d_coefs = nullptr; b->d_planes = nullptr; b->d_out = nullptr;
}

int batch_create_impl(b200jpg_ctx* ctx, const b200jpg_image_desc* imgs, size_t n, int* statuses, const PlanOverrides& ov,
                      b200jpg_batch** out) {
    if (!ctx || !out || (!imgs && n)) return B200JPG_ERR_INTERNAL;
    *out = nullptr;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    b200jpg_batch* b = new b200jpg_batch();
    b->ctx = ctx;
    b->n = n;
    b->layout.resize(n);
    b->images.resize(n);
    b->planes_absolute = ov.plane_addr != nullptr;
    b->outs_absolute = ov.out_addr != nullptr;
    b->fmode.assign(n, -1);
    std::map<std::string, unsigned> qt_index;
    size_t coef_off = 0, plane_off = 0, out_off = 0;
    int first_error = B200JPG_OK;
    std::string first_msg;
    for (size_t i = 0; i < n; i++) {
        const b200jpg_image_desc& d = imgs[i];
        ImageLayout& L = b->layout[i];
        b->strip_first.push_back((unsigned)b->strips.size());
        for (unsigned m = 0; m < KF_NMODES; m++) b->fcol_first[m].push_back((unsigned)b->fcols[m].size());
        std::string err;
        int rc = plan_image(ctx, d, &b->images[i], &err);
        for (int k = 0; rc == B200JPG_OK && k < d.ncomp; k++)
            if (!d.qt[k]) {
                rc = B200JPG_ERR_FORMAT;
                err = "invalid JPEG format: use of unset quantization table";  // src/decoder.rs:810-815
            }
        L.status = rc;
        if (statuses) statuses[i] = rc;
        if (rc) {
            if (!first_error) { first_error = rc; first_msg = err; }
            b->images[i].ncomp = 0;  // skipped by every kernel
            b->images[i].width = b->images[i].height = 0;
            b->images[i].path = 0xffffffffu;
            continue;
        }
        DevImage& img = b->images[i];
        L.tile_first = (unsigned)b->tiles.size();
        const unsigned comp0 = (unsigned)b->comps.size();
        for (int k = 0; k < d.ncomp; k++) {
            const b200jpg_component& c = d.comps[k];
            const size_t nblocks = (size_t)c.block_w * c.block_h;
            const size_t plane_bytes = nblocks * c.dct_scale * c.dct_scale;
            // quantisation table, expanded to u32 and de-duplicated by content
            std::string key((const char*)d.qt[k], 128);
            auto it = qt_index.find(key);
            unsigned qi;
            if (it == qt_index.end()) {
                qi = (unsigned)(b->qtabs.size() / 64);
                bool is8 = true;
                for (int j = 0; j < 64; j++) {
                    b->qtabs.push_back(d.qt[k][j]);
                    is8 = is8 && d.qt[k][j] <= 255;
                }
                for (int j = 0; j < 32; j++)
                    b->qpack.push_back(is8 ? ((unsigned)d.qt[k][2 * j] | ((unsigned)d.qt[k][2 * j + 1] << 24)) : 0u);
                b->qt_is8.push_back(is8 ? 1 : 0);
                qt_index.emplace(key, qi);
            } else {
                qi = it->second;
            }
            coef_off = align_up(coef_off, 1024);
            plane_off = align_up(plane_off, 256);
            L.coef_off[k] = coef_off;
            L.coef_bytes[k] = nblocks * 128;
            L.plane_off[k] = ov.plane_addr ? (size_t)ov.plane_addr[i][k] : plane_off;
            DevComp dc;
            dc.plane_off = L.plane_off[k];
            dc.stride = (unsigned)c.block_w * c.dct_scale;
            dc.block_w = c.block_w;
            dc.qt_index = qi;
            dc.dct_scale = c.dct_scale;
            dc.nblocks = (unsigned)nblocks;
            // 8-bit tables take the IDP.2A path; the first four tables of the batch sit in the constant bank
            dc.qflags = (b->qt_is8[qi] ? 1u : 0u) | ((qi < 4 ? qi : 0xffu) << 8);
            const unsigned comp_index = (unsigned)b->comps.size();
            b->comps.push_back(dc);
            if (c.dct_scale != 8) b->all_scale8 = false;
            if (dc.plane_off % 8 != 0) b->k1_tma_aligned = false;
            for (size_t first = 0; first < nblocks; first += K1_TILE) {
                DevTile t;
                t.comp = comp_index;
                t.slab_row = (unsigned)(coef_off / 128 + first);
                t.bxy = (unsigned)(first % c.block_w) | ((unsigned)(first / c.block_w) << 16);
                t.nvalid = (unsigned)std::min<size_t>(K1_TILE, nblocks - first);
                b->tiles.push_back(t);
            }
            img.c[k].plane_off = L.plane_off[k];
            b->info.n_blocks += nblocks;
            b->info.k1_algorithmic_bytes += nblocks * (128 + (size_t)c.dct_scale * c.dct_scale);
            b->info.k2_algorithmic_bytes += plane_bytes;
            coef_off += nblocks * 128;
            plane_off += plane_bytes;
        }
        L.tile_count = (unsigned)b->tiles.size() - L.tile_first;
        out_off = align_up(out_off, 256);
        L.out_off = ov.out_addr ? (size_t)ov.out_addr[i] : out_off;   // absolute: the kernels write the caller's buffer directly
        L.out_len = (size_t)d.width * d.height * d.ncomp;
        img.out_off = L.out_off;
        if (!ov.out_addr) out_off += L.out_len;
        b->info.n_pixels += (size_t)d.width * d.height;
        b->info.k2_algorithmic_bytes += L.out_len;
        // bulk-copy fed 4:2:0 kernel: every row it copies must start on a 16-byte boundary (cp.async.bulk)
        if (img.path == K2_PATH_420 && img.c[1].stride % 16 == 0 && img.c[1].stride == img.c[2].stride &&
            img.c[0].plane_off % 16 == 0 && img.c[1].plane_off % 16 == 0 && img.c[2].plane_off % 16 == 0)
            img.path = K2_PATH_420T;
        if (img.path == K2_PATH_420T) {
            const unsigned npairs = img.height / 2u + 1u;
            for (unsigned x0 = 0; x0 < img.width; x0 += 2048u) {
                b->strips.push_back(K2Strip{(unsigned)i, x0, b->strip_items, npairs});
                b->strip_items += npairs;
            }
        }
        {
            const int fm = fused_mode_of(ctx, d, img);
            b->fmode[i] = (signed char)fm;
            if (fm >= 0) {
                plan_fused_columns(b, (unsigned)i, fm, d, L, comp0);
                b->info.n_fused++;
                for (int k = 0; k < d.ncomp; k++) b->info.kf_algorithmic_bytes += (size_t)d.comps[k].block_w * d.comps[k].block_h * 128;
                b->info.kf_algorithmic_bytes += L.out_len;
            }
        }
        if (img.path < K2_NPATHS) {
            b->path_used[img.path] = true;
            b->path_max_w[img.path] = std::max(b->path_max_w[img.path], img.width);
            b->path_max_h[img.path] = std::max(b->path_max_h[img.path], img.height);
        }
    }
    if (coef_off / 128 + K1_TILE >= 0xffffffffull) {
        delete b;
        return fail(ctx, B200JPG_ERR_INTERNAL, "batch too large: more than 2^32 blocks");
    }
    b->info.coef_bytes = align_up(coef_off, 1024);
    b->info.plane_bytes = ov.plane_addr ? 0 : align_up(plane_off, 256);
    b->info.out_bytes = align_up(out_off, 256);
    if (first_error) b200jpg_fail(ctx, first_error, first_msg);

    cudaStream_t up_stream = ov.upload_stream ? ov.upload_stream : ctx->stream;
    size_t arena_used = 0;
    b->tables_borrowed = ov.arena != nullptr;
    auto upload = [&](void** dptr, const void* src, size_t bytes) -> cudaError_t {
        *dptr = nullptr;
        if (bytes == 0) return cudaSuccess;
        if (ov.arena) {
            const size_t at = align_up(arena_used, 256);
            if (at + bytes > ov.arena->bytes) return cudaErrorMemoryAllocation;
            memcpy(ov.arena->h + at, src, bytes);
            *dptr = ov.arena->d + at;
            arena_used = at + bytes;
            return cudaSuccess;
        }
        cudaError_t e = cudaMalloc(dptr, bytes);
        if (e != cudaSuccess) return e;
        // pageable source: the copy is staged before the call returns, so the vectors may be reused
        return cudaMemcpyAsync(*dptr, src, bytes, cudaMemcpyHostToDevice, up_stream);
    };
    cudaError_t e = upload((void**)&b->d_comps, b->comps.data(), b->comps.size() * sizeof(DevComp));
    if (e == cudaSuccess) e = upload((void**)&b->d_tiles, b->tiles.data(), b->tiles.size() * sizeof(DevTile));
    if (e == cudaSuccess) e = upload((void**)&b->d_images, b->images.data(), b->images.size() * sizeof(DevImage));
    if (e == cudaSuccess) e = upload((void**)&b->d_qtabs, b->qtabs.data(), b->qtabs.size() * sizeof(unsigned));
    if (e == cudaSuccess) e = upload((void**)&b->d_qpack, b->qpack.data(), b->qpack.size() * sizeof(unsigned));
    if (e == cudaSuccess) e = upload((void**)&b->d_strips, b->strips.data(), b->strips.size() * sizeof(K2Strip));
    for (unsigned m = 0; m < KF_NMODES; m++)
        if (e == cudaSuccess) e = upload((void**)&b->d_fcols[m], b->fcols[m].data(), b->fcols[m].size() * sizeof(FColumn));
    if (e == cudaSuccess && ov.arena && arena_used)
        e = cudaMemcpyAsync(ov.arena->d, ov.arena->h, arena_used, cudaMemcpyHostToDevice, up_stream);
    b->table_bytes = arena_used;
    b->strip_first.push_back((unsigned)b->strips.size());
    for (unsigned m = 0; m < KF_NMODES; m++) b->fcol_first[m].push_back((unsigned)b->fcols[m].size());
    memset(&b->qcache, 0, sizeof b->qcache);
    for (size_t t = 0; t < 4 && t < b->qt_is8.size(); t++)
        memcpy(b->qcache.b[t], b->qpack.data() + 32 * t, 32 * sizeof(unsigned));
    if (e != cudaSuccess) {
        batch_release_device(b);
        delete b;
        return cuda_fail(ctx, e, "batch table upload");
    }
    *out = b;
    return B200JPG_OK;
}

static int ensure_tensor_map(b200jpg_batch* b, const void* d_coefs) {
    b200jpg_ctx* ctx = b->ctx;
    if (b->tmap_base == d_coefs) return B200JPG_OK;
    if (!ctx->encode) return fail(ctx, B200JPG_ERR_INTERNAL, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t rows = b->info.coef_bytes / 128;
    cuuint64_t gdim[2] = {64, rows};
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)K1_TILE};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = ctx->encode(&b->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(d_coefs), gdim, gstride, box,
                             estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[96];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return fail(ctx, B200JPG_ERR_INTERNAL, buf);
    }
    cuuint32_t box32[2] = {64, 32};
    r = ctx->encode(&b->tmap32, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(d_coefs), gdim, gstride, box32, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[96];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled (32-row box) failed with CUresult %d", (int)r);
        return fail(ctx, B200JPG_ERR_INTERNAL, buf);
    }
    b->tmap_base = d_coefs;
    return B200JPG_OK;
}

// K1 over tiles [tile_first, tile_first+tile_count), K2 over images [img_first, img_first+img_count)
int batch_launch(b200jpg_batch* b, const void* d_coefs, void* d_planes, void* d_out, int stages, unsigned tile_first,
                        unsigned tile_count, unsigned img_first, unsigned img_count, cudaStream_t stream) {
    b200jpg_ctx* ctx = b->ctx;
    // both stages in one call and every image of the range eligible: the fused kernel, planes never leave the SM
    const unsigned fmodes = fuse_modes(ctx);
    if (stages == 3 && img_count && fmodes && ((uintptr_t)d_coefs % 16 == 0) && (d_out || b->outs_absolute)) {
        bool all = true;
        for (unsigned i = img_first; i < img_first + img_count && all; i++)
            all = b->layout[i].status != B200JPG_OK || (b->fmode[i] >= 0 && ((fmodes >> b->fmode[i]) & 1u));
        if (all) {
            for (unsigned m = 0; m < KF_NMODES; m++) {
                const unsigned c0 = b->fcol_first[m][img_first], c1 = b->fcol_first[m][img_first + img_count];
                if (c1 == c0) continue;
                int rc = ensure_tensor_map(b, d_coefs);
                if (rc) return rc;
                KFParams p;
                p.cols = b->d_fcols[m] + c0;
                p.comps = b->d_comps;
                p.qtabs = b->d_qtabs;
                p.qpack = b->d_qpack;
                p.coefs = (const short*)d_coefs;
                p.images = b->d_images;
                p.out = (uint8_t*)d_out;
                p.ncols = c1 - c0;
                p.item_base = b->fcols[m][c0].first_item;
                p.total_items = (c1 < b->fcols[m].size() ? b->fcols[m][c1].first_item : b->fitems[m]) - p.item_base;
                p.ystride = b->f_ystride[m];
                p.cstride = b->f_cstride[m];
                p.sixteen = 16;
                cudaError_t e = launch_kf(m, b->tmap32, b->qcache, p, ctx->num_sms, stream);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "KF launch");
                ctx->launches++;
            }
            return B200JPG_OK;
        }
    }
    if ((stages & 1) && tile_count) {
        K1Params p;
        p.tiles = b->d_tiles + tile_first;
        p.comps = b->d_comps;
        p.qtabs = b->d_qtabs;
        p.qpack = b->d_qpack;
        p.coefs = (const short*)d_coefs;
        p.planes = (uint8_t*)d_planes;
        p.ntiles = tile_count;
        p.one = 1u;
        p.minus_one = 0xffffffffu;
        const bool tma_ok = b->k1_tma_aligned &&
                            ((uintptr_t)d_coefs % 16 == 0) && ((uintptr_t)d_planes % 8 == 0);
        if (ctx->k1_kernel == B200JPG_KERNEL_FAST && !tma_ok)
            return fail(ctx, B200JPG_ERR_INTERNAL, "k1_kernel=FAST requested but the batch is not eligible (plane offsets must be 8-byte aligned)");
        if (tma_ok && ctx->k1_kernel != B200JPG_KERNEL_GENERIC) {
            int rc = ensure_tensor_map(b, d_coefs);
            if (rc) return rc;
            CU_TRY(ctx, launch_k1_tma(b->tmap, b->qcache, p, ctx->arith == B200JPG_ARITH_SSSE3 ? 1 : 0, !b->all_scale8, ctx->num_sms, stream));
        } else {
            CU_TRY(ctx, launch_k1_generic(p, ctx->arith, stream));
        }
        ctx->launches++;
    }
    if ((stages & 2) && img_count) {
        K2Params p;
        p.images = b->d_images;
        p.planes = (const uint8_t*)d_planes;
        p.out = (uint8_t*)d_out;
        p.nimages = (unsigned)b->n;
        p.sixteen = make_int3(16, 16, 16);
        // (the bulk-copy fed 4:2:0 kernel exists in scalar arithmetic only; SSSE3 mode takes the load/store one)
        const bool bulk = k2_mode() == 0 && ((uintptr_t)d_planes % 16 == 0) && ctx->arith != B200JPG_ARITH_SSSE3;
        p.flags = (bulk ? 0u : K2_FLAG_LDG_TAKES_420T) | (ctx->arith == B200JPG_ARITH_SSSE3 ? K2_FLAG_SSSE3 : 0u);
        if (bulk && b->path_used[K2_PATH_420T]) {
            const unsigned s0 = b->strip_first[img_first], s1 = b->strip_first[img_first + img_count];
            if (s1 > s0) {
                const unsigned base = b->strips[s0].first_item;
                const unsigned end = s1 < b->strips.size() ? b->strips[s1].first_item : b->strip_items;
                cudaError_t e = launch_k2_420_tma(p, b->d_strips + s0, s1 - s0, base, end - base, ctx->num_sms, stream);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "K2 launch");
                ctx->launches++;
            }
        }
        if (ctx->k2_kernel == B200JPG_KERNEL_FAST && b->path_used[K2_PATH_GENERIC])
            return fail(ctx, B200JPG_ERR_INTERNAL, "k2_kernel=FAST requested but some image needs the generic kernel");
        for (unsigned first = img_first; first < img_first + img_count; first += 65535u) {
            const unsigned count = std::min(65535u, img_first + img_count - first);
            for (int path = 0; path < (int)K2_NPATHS; path++) {
                if (path == K2_PATH_420T) continue;
                unsigned max_w = b->path_used[path] ? b->path_max_w[path] : 0, max_h = b->path_used[path] ? b->path_max_h[path] : 0;
                if (path == K2_PATH_420 && !bulk && b->path_used[K2_PATH_420T]) {  // the load/store kernel takes those images too
                    max_w = std::max(max_w, b->path_max_w[K2_PATH_420T]);
                    max_h = std::max(max_h, b->path_max_h[K2_PATH_420T]);
                }
                if (max_w == 0 || max_h == 0) continue;
                cudaError_t e = cudaSuccess;
                if (path == K2_PATH_GRAY) e = launch_k2_gray(p, first, count, max_w, max_h, stream);
                else if (path == K2_PATH_GENERIC) e = launch_k2_generic(p, first, count, max_w, max_h, stream);
                else if (path == K2_PATH_420 || path == K2_PATH_420R)
                    e = launch_k2_420(p, first, count, max_w, max_h, path == K2_PATH_420R, stream);
                else e = launch_k2_rows16((unsigned)path, p, first, count, max_w, max_h, stream);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "K2 launch");
                ctx->launches++;
            }
        }
    }
    return B200JPG_OK;
}

extern "C" {

int b200jpg_batch_create(b200jpg_ctx* ctx, const b200jpg_image_desc* imgs, size_t n, int* statuses, b200jpg_batch** batch) {
    PlanOverrides ov;
    return batch_create_impl(ctx, imgs, n, statuses, ov, batch);
}

void b200jpg_batch_free(b200jpg_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaStreamSynchronize(b->ctx->stream2);
    batch_release_device(b);
    delete b;
}

int b200jpg_batch_get_info(const b200jpg_batch* b, b200jpg_batch_info* info) {
    if (!b || !info) return B200JPG_ERR_INTERNAL;
    *info = b->info;
    return B200JPG_OK;
}

int b200jpg_batch_image_layout(const b200jpg_batch* b, size_t i, size_t coef_off[4], size_t plane_off[4], size_t* out_off,
                               size_t* out_len) {
    if (!b || i >= b->n) return B200JPG_ERR_INTERNAL;
    const ImageLayout& L = b->layout[i];
    for (int k = 0; k < 4; k++) {
        if (coef_off) coef_off[k] = L.coef_off[k];
        if (plane_off) plane_off[k] = L.plane_off[k];
    }
    if (out_off) *out_off = L.out_off;
    if (out_len) *out_len = L.out_len;
    return L.status;
}

int b200jpg_batch_run_device(b200jpg_batch* b, const void* d_coefs, void* d_planes, void* d_out, int stages) {
    if (!b) return B200JPG_ERR_INTERNAL;
    b200jpg_ctx* ctx = b->ctx;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    return batch_launch(b, d_coefs, d_planes, d_out, stages, 0, (unsigned)b->tiles.size(), 0, (unsigned)b->n, ctx->stream);
}

int b200jpg_batch_format_device(b200jpg_batch* b, const void* d_out, int format, void* d_dst, const float scale[3], const float bias[3],
                                int* statuses) {
    if (!b || !d_out || !d_dst) return B200JPG_ERR_INTERNAL;
    b200jpg_ctx* ctx = b->ctx;
    if (format < B200JPG_FMT_RGB8_PLANAR || format > B200JPG_FMT_RGB_F32_NCHW) return fail(ctx, B200JPG_ERR_INTERNAL, "unknown output format");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    unsigned max_w = 0, max_h = 0;
    for (size_t i = 0; i < b->n; i++) {
        const bool ok = b->layout[i].status == B200JPG_OK && b->images[i].ncomp == 3;
        if (statuses) statuses[i] = ok ? B200JPG_OK : (b->layout[i].status ? b->layout[i].status : B200JPG_ERR_FORMAT);
        if (ok) {
            max_w = std::max(max_w, b->images[i].width);
            max_h = std::max(max_h, b->images[i].height);
        }
    }
    for (size_t first = 0; first < b->n; first += 65535u) {
        const unsigned count = (unsigned)std::min<size_t>(65535u, b->n - first);
        CU_TRY(ctx, launch_k3_format(b->d_images, (unsigned)first, count, max_w, max_h, d_out, d_dst, format, scale, bias, ctx->stream));
        ctx->launches++;
    }
    return B200JPG_OK;
}

// Host -> host: chunked, double-buffered over two streams (H2D | K1 | K2 | D2H overlap across chunks).
int b200jpg_batch_run_host(b200jpg_batch* b, const b200jpg_image_desc* imgs, uint8_t* const* outs, const size_t* out_caps,
                           int* statuses) {
    if (!b || !imgs || !outs) return B200JPG_ERR_INTERNAL;
    b200jpg_ctx* ctx = b->ctx;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    bool use_compaction = b->compact_decision == 1;
    bool trial = false;
    if (b->compact_decision < 0) {
        const int cpus = ctx->host_threads > 0 ? ctx->host_threads : stream_engine_default_threads();
        if (ctx->host_compact == B200JPG_COMPACT_ON) b->compact_decision = 1;
        else if (ctx->host_compact == B200JPG_COMPACT_OFF || b->n < 4 || cpus < 8 || b->d_coefs) b->compact_decision = 0;
        else if (b->compact_trials == 0 && stream_engine_sample_density(imgs, b->n) >= 0.35) b->compact_decision = 0;
        use_compaction = b->compact_decision == 1;
        if (b->compact_decision < 0) {  // AUTO: runs 1-2 dense, runs 3-4 compacted, then whichever second run was faster
            trial = true;
            use_compaction = b->compact_trials >= 2;
        }
    }
    const double t_run0 = now();
    if (use_compaction) {
        std::vector<int> plan(b->n);
        for (size_t m = 0; m < b->n; m++) plan[m] = b->layout[m].status;
        const int rc = stream_engine_run_dense(ctx, imgs, b->n, plan.data(), outs, out_caps, statuses, ctx->host_threads);
        if (trial && ++b->compact_trials == 4) {  // (the first run of each kind pays for allocations and page-locking)
            b->compact_ms[1] = now() - t_run0;
            b->compact_decision = b->compact_ms[1] < b->compact_ms[0] ? 1 : 0;
            if (getenv("B200JPG_TRACE"))
                fprintf(stderr, "[b200jpg] run_host auto: dense upload %.1f ms, host compaction %.1f ms -> %s\n", b->compact_ms[0],
                        b->compact_ms[1], b->compact_decision ? "compaction" : "dense upload");
        }
        return rc;
    }
    if (!b->d_coefs && !b->d_planes && !b->d_out) {
        // internal slabs: borrow the context's grow-only scratch buffers when they are free (repeated batches then
        // cost no cudaMalloc/cudaFree), otherwise allocate private ones
        const size_t need[3] = {b->info.coef_bytes, b->info.plane_bytes, b->info.out_bytes};
        bool borrowed = false;
        {
            std::lock_guard<std::mutex> lock(ctx->mu);
            if (!ctx->scratch_busy) {
                ctx->scratch_busy = true;
                borrowed = true;
            }
        }
        if (borrowed) {
            for (int k = 0; k < 3; k++)
                if (ctx->scratch[k].cap < need[k]) {
                    cudaFree(ctx->scratch[k].p);
                    ctx->scratch[k].p = nullptr;
                    ctx->scratch[k].cap = 0;
                    const size_t want = need[k] + need[k] / 8;
                    cudaError_t e = cudaMalloc(&ctx->scratch[k].p, want);
                    if (e != cudaSuccess) {
                        std::lock_guard<std::mutex> lock(ctx->mu);
                        ctx->scratch_busy = false;
                        return cuda_fail(ctx, e, "scratch slab allocation");
                    }
                    ctx->scratch[k].cap = want;
                }
            b->d_coefs = ctx->scratch[0].p;
            b->d_planes = ctx->scratch[1].p;
            b->d_out = ctx->scratch[2].p;
            b->slabs_borrowed = true;
            b->tmap_base = nullptr;
        }
    }
    if (!b->d_coefs && b->info.coef_bytes) CU_TRY(ctx, cudaMalloc(&b->d_coefs, b->info.coef_bytes));
    if (!b->d_planes && b->info.plane_bytes) CU_TRY(ctx, cudaMalloc(&b->d_planes, b->info.plane_bytes));
    if (!b->d_out && b->info.out_bytes) CU_TRY(ctx, cudaMalloc(&b->d_out, b->info.out_bytes));
    // the table uploads were enqueued on ctx->stream; make the second stream wait for them
    cudaEvent_t ready;
    CU_TRY(ctx, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CU_TRY(ctx, cudaEventRecord(ready, ctx->stream));
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream2, ready, 0));
    cudaEventDestroy(ready);

    const size_t chunk_target = (size_t)96 << 20;  // bytes of coefficients per chunk
    int result = B200JPG_OK;
    size_t i = 0;
    int which = 0;
    while (i < b->n) {
        size_t j = i, bytes = 0;
        while (j < b->n && (j == i || bytes < chunk_target)) {
            for (int k = 0; k < 4; k++) bytes += b->layout[j].coef_bytes[k];
            j++;
        }
        cudaStream_t s = which ? ctx->stream2 : ctx->stream;
        which ^= 1;
        unsigned tile_first = 0, tile_count = 0;
        bool have_tiles = false;
        const char* run_src = nullptr;
        size_t run_dst = 0, run_bytes = 0;
        for (size_t m = i; m < j; m++) {
            const ImageLayout& L = b->layout[m];
            if (statuses) statuses[m] = L.status;
            if (L.status) continue;
            if (out_caps && out_caps[m] < L.out_len) {
                if (statuses) statuses[m] = B200JPG_ERR_INTERNAL;
                result = fail(ctx, B200JPG_ERR_INTERNAL, "output buffer too small");
                continue;
            }
            bool have_all = true;
            for (int k = 0; k < imgs[m].ncomp; k++) have_all = have_all && imgs[m].coefs[k] != nullptr;
            if (!have_all) {  // "not all components have data", src/decoder.rs:1306-1308
                if (statuses) statuses[m] = B200JPG_ERR_FORMAT;
                result = fail(ctx, B200JPG_ERR_FORMAT, "invalid JPEG format: not all components have data");
                continue;
            }
            for (int k = 0; k < imgs[m].ncomp; k++) {
                // merge copies that are contiguous on both sides (one arena per batch is the common case)
                const char* src = (const char*)imgs[m].coefs[k];
                if (run_bytes && run_src + run_bytes == src && run_dst + run_bytes == L.coef_off[k]) {
                    run_bytes += L.coef_bytes[k];
                } else {
                    if (run_bytes) CU_TRY(ctx, cudaMemcpyAsync((char*)b->d_coefs + run_dst, run_src, run_bytes, cudaMemcpyHostToDevice, s));
                    run_src = src;
                    run_dst = L.coef_off[k];
                    run_bytes = L.coef_bytes[k];
                }
            }
            if (!have_tiles) { tile_first = L.tile_first; have_tiles = true; }
            tile_count = L.tile_first + L.tile_count - tile_first;
        }
        if (run_bytes) CU_TRY(ctx, cudaMemcpyAsync((char*)b->d_coefs + run_dst, run_src, run_bytes, cudaMemcpyHostToDevice, s));
        int rc = batch_launch(b, b->d_coefs, b->d_planes, b->d_out, 3, tile_first, tile_count, (unsigned)i, (unsigned)(j - i), s);
        if (rc) return rc;
        char* out_dst = nullptr;
        size_t out_src = 0, out_bytes = 0;
        for (size_t m = i; m < j; m++) {
            const ImageLayout& L = b->layout[m];
            if (L.status || (statuses && statuses[m])) continue;
            if (out_bytes && out_dst + out_bytes == (char*)outs[m] && out_src + out_bytes == L.out_off) {
                out_bytes += L.out_len;
            } else {
                if (out_bytes) CU_TRY(ctx, cudaMemcpyAsync(out_dst, (char*)b->d_out + out_src, out_bytes, cudaMemcpyDeviceToHost, s));
                out_dst = (char*)outs[m];
                out_src = L.out_off;
                out_bytes = L.out_len;
            }
        }
        if (out_bytes) CU_TRY(ctx, cudaMemcpyAsync(out_dst, (char*)b->d_out + out_src, out_bytes, cudaMemcpyDeviceToHost, s));
        i = j;
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream2));
    if (trial) {
        b->compact_ms[0] = now() - t_run0;
        b->compact_trials++;
    }
    return result;
}

int b200jpg_decode_batch(b200jpg_ctx* ctx, const b200jpg_image_desc* imgs, size_t n, uint8_t* const* outs,
                         const size_t* out_caps, int* statuses) {
    const bool trace = getenv("B200JPG_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    b200jpg_batch* b = nullptr;
    int rc = b200jpg_batch_create(ctx, imgs, n, statuses, &b);
    if (rc) return rc;
    const double t1 = now();
    rc = b200jpg_batch_run_host(b, imgs, outs, out_caps, statuses);
    const double t2 = now();
    b200jpg_batch_free(b);
    if (trace) fprintf(stderr, "[b200jpg] decode_batch(%zu): plan %.1f ms, run_host %.1f ms, free %.1f ms\n", n, t1 - t0, t2 - t1, now() - t2);
    return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// worker-shaped API (trait Worker, src/worker/mod.rs:24-35; ImmediateWorker, src/worker/immediate.rs)
// ---------------------------------------------------------------------------------------------
struct WorkerSlot {
    bool started = false;
    bool have_result = false;
    b200jpg_component comp{};
    uint16_t qt[64];
    size_t offset_i16 = 0;      // coefficients appended so far
    size_t rows = 0;            // MCU rows appended so far
    int16_t* h_coefs = nullptr; // pinned staging, block_w*block_h*64
    void* d_coefs = nullptr;
    void* d_plane = nullptr;
    size_t plane_len = 0;
    // The buffers and the launch plan outlive start(): a decoder that pushes image after image of one geometry through the
    // trait (a video stream, a directory of camera files) pays for page-locking, device allocations and the plan's tables
    // once, not three times per image.  Capacities in bytes; the plan is valid for (component, rows, table).
    size_t h_cap = 0, d_coefs_cap = 0, d_plane_cap = 0;
    b200jpg_batch* plan = nullptr;
    b200jpg_component plan_comp{};
    size_t plan_rows = 0;
    uint16_t plan_qt[64];
};

// what compute_image keeps between calls of one worker: the pixel buffer on the device and the launch plan, valid while the
// geometry, the colour transform and the addresses of the component planes stay what they were
struct ComputeCache {
    void* d_out = nullptr;
    size_t d_out_cap = 0;
    b200jpg_batch* plan = nullptr;
    b200jpg_image_desc desc;
    unsigned long long addr[4] = {0, 0, 0, 0};
};

struct b200jpg_worker {
    b200jpg_ctx* ctx = nullptr;
    WorkerSlot slot[4];
    ComputeCache cc;
};

static void slot_release(WorkerSlot& s) {
    if (s.h_coefs) cudaFreeHost(s.h_coefs);
    cudaFree(s.d_coefs);
    cudaFree(s.d_plane);
    if (s.plan) b200jpg_batch_free(s.plan);
    s = WorkerSlot();
}

extern "C" {

int b200jpg_worker_new(b200jpg_ctx* ctx, b200jpg_worker** w) {
    if (!ctx || !w) return B200JPG_ERR_INTERNAL;
    *w = new b200jpg_worker();
    (*w)->ctx = ctx;
    return B200JPG_OK;
}

void b200jpg_worker_free(b200jpg_worker* w) {
    if (!w) return;
    cudaSetDevice(w->ctx->device);
    cudaStreamSynchronize(w->ctx->stream);
    for (auto& s : w->slot) slot_release(s);
    cudaFree(w->cc.d_out);
    if (w->cc.plan) b200jpg_batch_free(w->cc.plan);
    delete w;
}

int b200jpg_worker_start(b200jpg_worker* w, int index, const b200jpg_component* c, const uint16_t qt_natural[64]) {
    if (!w || !c || !qt_natural) return B200JPG_ERR_INTERNAL;
    b200jpg_ctx* ctx = w->ctx;
    if (index < 0 || index >= 4) return fail(ctx, B200JPG_ERR_INTERNAL, "worker index out of range");
    WorkerSlot& s = w->slot[index];
    // assert!(self.results[data.index].is_empty()), src/worker/immediate.rs:31
    if (s.started && !s.have_result) return fail(ctx, B200JPG_ERR_INTERNAL, "Worker::start on a component whose result was not collected");
    if (!(c->dct_scale == 1 || c->dct_scale == 2 || c->dct_scale == 4 || c->dct_scale == 8))
        return fail(ctx, B200JPG_ERR_INTERNAL, "Unsupported IDCT scale");
    if (c->block_w == 0 || c->block_h == 0 || c->v == 0 || c->block_h % c->v != 0)
        return fail(ctx, B200JPG_ERR_INTERNAL, "component geometry not initialised");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    s.started = s.have_result = false;
    s.offset_i16 = s.rows = 0;
    s.comp = *c;
    memcpy(s.qt, qt_natural, 128);
    const size_t nblocks = (size_t)c->block_w * c->block_h;
    s.plane_len = nblocks * c->dct_scale * c->dct_scale;
    if (nblocks * 128 > s.h_cap) {
        if (s.h_coefs) cudaFreeHost(s.h_coefs);
        s.h_coefs = nullptr;
        s.h_cap = 0;
        CU_TRY(ctx, cudaHostAlloc((void**)&s.h_coefs, nblocks * 128, cudaHostAllocDefault));
        s.h_cap = nblocks * 128;
    }
    if (align_up(nblocks * 128, 1024) > s.d_coefs_cap) {  // the TMA tensor map spans whole KiB
        cudaFree(s.d_coefs);
        s.d_coefs = nullptr;
        s.d_coefs_cap = 0;
        CU_TRY(ctx, cudaMalloc(&s.d_coefs, align_up(nblocks * 128, 1024)));
        s.d_coefs_cap = align_up(nblocks * 128, 1024);
    }
    if (s.plane_len > s.d_plane_cap) {
        cudaFree(s.d_plane);
        s.d_plane = nullptr;
        s.d_plane_cap = 0;
        CU_TRY(ctx, cudaMalloc(&s.d_plane, s.plane_len));
        s.d_plane_cap = s.plane_len;
    }
    // the plane starts zeroed; MCU rows never appended stay 0 (src/worker/rayon.rs:46)
    CU_TRY(ctx, cudaMemsetAsync(s.d_plane, 0, s.plane_len, ctx->stream));
    s.started = true;
    return B200JPG_OK;
}

int b200jpg_worker_append_rows(b200jpg_worker* w, int index, const int16_t* coefs, size_t n_i16, size_t nrows) {
    if (!w || !coefs) return B200JPG_ERR_INTERNAL;
    b200jpg_ctx* ctx = w->ctx;
    if (index < 0 || index >= 4 || !w->slot[index].started || w->slot[index].have_result)
        return fail(ctx, B200JPG_ERR_INTERNAL, "Worker::append_row on a component that was not started");
    WorkerSlot& s = w->slot[index];
    const size_t per_row = (size_t)s.comp.block_w * s.comp.v * 64;
    // assert_eq!(data.len(), block_count * 64), src/worker/immediate.rs:47
    if (n_i16 != per_row * nrows) return fail(ctx, B200JPG_ERR_INTERNAL, "Worker::append_row: data.len() != block_count * 64");
    if (s.offset_i16 + n_i16 > (size_t)s.comp.block_w * s.comp.block_h * 64)
        return fail(ctx, B200JPG_ERR_INTERNAL, "Worker::append_row: more rows than the plane holds");
    memcpy(s.h_coefs + s.offset_i16, coefs, n_i16 * sizeof(int16_t));
    s.offset_i16 += n_i16;
    s.rows += nrows;
    return B200JPG_OK;
}

int b200jpg_worker_append_row(b200jpg_worker* w, int index, const int16_t* coefs, size_t n_i16) {
    return b200jpg_worker_append_rows(w, index, coefs, n_i16, 1);
}

int b200jpg_worker_get_result(b200jpg_worker* w, int index, uint8_t* plane_out, size_t cap, size_t* plane_len) {
    if (!w) return B200JPG_ERR_INTERNAL;
    b200jpg_ctx* ctx = w->ctx;
    if (index < 0 || index >= 4) return fail(ctx, B200JPG_ERR_INTERNAL, "worker index out of range");
    WorkerSlot& s = w->slot[index];
    if (!s.started) {  // mem::take of an empty Vec
        if (plane_len) *plane_len = 0;
        return B200JPG_OK;
    }
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (!s.have_result && s.rows > 0) {
        // one-component pseudo image covering exactly the appended MCU rows
        b200jpg_image_desc d;
        memset(&d, 0, sizeof d);
        d.ncomp = 1;
        d.comps[0] = s.comp;
        d.comps[0].block_h = (uint16_t)(s.rows * s.comp.v);
        d.comps[0].size_w = (uint16_t)std::min<size_t>(s.comp.size_w ? s.comp.size_w : 1, (size_t)s.comp.block_w * s.comp.dct_scale);
        d.comps[0].size_h = (uint16_t)std::max<size_t>(1, std::min<size_t>(s.comp.size_h, (size_t)d.comps[0].block_h * s.comp.dct_scale));
        d.width = d.comps[0].size_w;
        d.height = d.comps[0].size_h;
        d.color_transform = B200JPG_CT_GRAYSCALE;
        d.qt[0] = s.qt;
        int rc = B200JPG_OK;
        if (!s.plan || s.plan_rows != s.rows || memcmp(&s.plan_comp, &s.comp, sizeof s.comp) != 0 || memcmp(s.plan_qt, s.qt, 128) != 0) {
            if (s.plan) b200jpg_batch_free(s.plan);
            s.plan = nullptr;
            PlanOverrides ov;
            rc = batch_create_impl(ctx, &d, 1, nullptr, ov, &s.plan);
            if (rc) return rc;
            if (s.plan->layout[0].status) {
                rc = s.plan->layout[0].status;
                b200jpg_batch_free(s.plan);
                s.plan = nullptr;
                return rc;
            }
            s.plan_comp = s.comp;
            s.plan_rows = s.rows;
            memcpy(s.plan_qt, s.qt, 128);
        }
        b200jpg_batch* b = s.plan;
        cudaError_t e = cudaMemcpyAsync(s.d_coefs, s.h_coefs, s.offset_i16 * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            rc = batch_launch(b, s.d_coefs, s.d_plane, nullptr, 1, 0, (unsigned)b->tiles.size(), 0, 1, ctx->stream);
        if (e == cudaSuccess && rc == B200JPG_OK) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "worker get_result");
        if (rc) return rc;
    }
    s.have_result = true;
    if (plane_len) *plane_len = s.plane_len;
    if (plane_out) {
        if (cap < s.plane_len) return fail(ctx, B200JPG_ERR_INTERNAL, "plane buffer too small");
        CU_TRY(ctx, cudaMemcpyAsync(plane_out, s.d_plane, s.plane_len, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return B200JPG_OK;
}

static int compute_image_device(b200jpg_ctx* ctx, const b200jpg_component* comps, int ncomp, const void* const* d_planes,
                                uint16_t out_w, uint16_t out_h, int color_transform, uint8_t* out, size_t cap, size_t* out_len,
                                ComputeCache* cache = nullptr) {
    b200jpg_image_desc d;
    memset(&d, 0, sizeof d);
    if (ncomp < 1 || ncomp > 4) return fail(ctx, B200JPG_ERR_FORMAT, "invalid JPEG format: not all components have data");
    d.ncomp = (uint8_t)ncomp;
    d.width = out_w;
    d.height = out_h;
    d.color_transform = (uint8_t)color_transform;
    static const uint16_t dummy_qt[64] = {1};
    unsigned long long addr[1][4] = {{0, 0, 0, 0}};
    for (int i = 0; i < ncomp; i++) {
        d.comps[i] = comps[i];
        d.qt[i] = dummy_qt;
        addr[0][i] = (unsigned long long)(uintptr_t)d_planes[i];
    }
    if (ncomp == 1) {  // src/decoder.rs:1314-1316: the output is component.size
        d.width = comps[0].size_w;
        d.height = comps[0].size_h;
    }
    b200jpg_batch* b = nullptr;
    int rc = B200JPG_OK;
    for (int i = 0; i < 4; i++) d.qt[i] = i < ncomp ? dummy_qt : nullptr;  // (every byte of d is now a function of the arguments)
    const bool hit = cache && cache->plan && memcmp(&cache->desc, &d, sizeof d) == 0 && memcmp(cache->addr, addr[0], sizeof addr[0]) == 0;
    if (hit) {
        b = cache->plan;
    } else {
        PlanOverrides ov;
        ov.plane_addr = addr;
        rc = batch_create_impl(ctx, &d, 1, nullptr, ov, &b);
        if (rc) return rc;
        if (cache) {
            if (cache->plan) b200jpg_batch_free(cache->plan);
            cache->plan = b;
            cache->desc = d;
            memcpy(cache->addr, addr[0], sizeof addr[0]);
        }
    }
    rc = b->layout[0].status;
    const size_t len = b->layout[0].out_len;
    void* d_out = nullptr;
    if (rc == B200JPG_OK && cap < len) rc = fail(ctx, B200JPG_ERR_INTERNAL, "output buffer too small");
    cudaError_t e = cudaSuccess;
    if (rc == B200JPG_OK && cache) {
        if (cache->d_out_cap < (len ? len : 1)) {
            cudaFree(cache->d_out);
            cache->d_out = nullptr;
            cache->d_out_cap = 0;
            e = cudaMalloc(&cache->d_out, len ? len : 1);
            if (e == cudaSuccess) cache->d_out_cap = len ? len : 1;
        }
        d_out = cache->d_out;
    } else if (rc == B200JPG_OK) {
        e = cudaMalloc(&d_out, len ? len : 1);
    }
    if (rc == B200JPG_OK && e == cudaSuccess) rc = batch_launch(b, nullptr, nullptr, d_out, 2, 0, 0, 0, 1, ctx->stream);
    if (rc == B200JPG_OK && e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, len, cudaMemcpyDeviceToHost, ctx->stream);
    if (rc == B200JPG_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (!cache) {
        cudaFree(d_out);
        b200jpg_batch_free(b);
    }
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "compute_image");
    if (out_len) *out_len = len;
    return B200JPG_OK;
}

int b200jpg_worker_compute_image(b200jpg_worker* w, int ncomp, uint16_t out_w, uint16_t out_h, int color_transform,
                                 uint8_t* out, size_t cap, size_t* out_len) {
    if (!w || !out) return B200JPG_ERR_INTERNAL;
    b200jpg_ctx* ctx = w->ctx;
    if (ncomp < 1 || ncomp > 4) return fail(ctx, B200JPG_ERR_FORMAT, "invalid JPEG format: not all components have data");
    b200jpg_component comps[4];
    const void* planes[4];
    for (int i = 0; i < ncomp; i++) {
        // "not all components have data", src/decoder.rs:1306-1308
        if (!w->slot[i].started || !w->slot[i].have_result) return fail(ctx, B200JPG_ERR_FORMAT, "invalid JPEG format: not all components have data");
        comps[i] = w->slot[i].comp;
        planes[i] = w->slot[i].d_plane;
    }
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    return compute_image_device(ctx, comps, ncomp, planes, out_w, out_h, color_transform, out, cap, out_len, &w->cc);
}

int b200jpg_compute_image(b200jpg_ctx* ctx, const b200jpg_component* comps, int ncomp, const uint8_t* const* planes,
                          const size_t* plane_len, uint16_t out_w, uint16_t out_h, int color_transform, uint8_t* out, size_t cap,
                          size_t* out_len) {
    if (!ctx || !comps || !planes || !plane_len || !out) return B200JPG_ERR_INTERNAL;
    // data.is_empty() || data.iter().any(Vec::is_empty), src/decoder.rs:1306-1308
    if (ncomp < 1 || ncomp > 4) return fail(ctx, B200JPG_ERR_FORMAT, "invalid JPEG format: not all components have data");
    for (int i = 0; i < ncomp; i++)
        if (!planes[i] || plane_len[i] == 0) return fail(ctx, B200JPG_ERR_FORMAT, "invalid JPEG format: not all components have data");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    void* d[4] = {nullptr, nullptr, nullptr, nullptr};
    int rc = B200JPG_OK;
    for (int i = 0; i < ncomp && rc == B200JPG_OK; i++) {
        const size_t need = (size_t)comps[i].block_w * comps[i].block_h * comps[i].dct_scale * comps[i].dct_scale;
        if (plane_len[i] < need) {
            rc = fail(ctx, B200JPG_ERR_INTERNAL, "plane shorter than block_w*block_h*dct_scale^2 (the reference would panic on the slice)");
            break;
        }
        cudaError_t e = cudaMalloc(&d[i], need ? need : 1);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d[i], planes[i], need, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "compute_image upload");
    }
    if (rc == B200JPG_OK) rc = compute_image_device(ctx, comps, ncomp, d, out_w, out_h, color_transform, out, cap, out_len);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 4; i++) cudaFree(d[i]);
    return rc;
}

}  // extern "C"
