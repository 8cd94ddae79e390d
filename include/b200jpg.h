/*
 * b200jpg.h -- C ABI of the B200-native JPEG block pipeline.
 *
 * This is the drop-in boundary for the per-MCU worker path of image-rs/jpeg-decoder v0.3.2
 * (reference at /root/reference): everything after Huffman decoding -- dequantise + IDCT,
 * plane assembly, chroma upsampling, colour conversion -- runs as hand-written sm_100a CUDA
 * kernels.  Plain pointers and sizes only; no C++ or torch types.  Each entry point cites the
 * reference interface it replaces.  INTEGRATION.md shows the Rust `extern "C"` block a
 * maintainer of the reference would add to bind these.
 *
 * There is no CPU fallback: every function that needs the device returns
 * B200JPG_ERR_INTERNAL when CUDA is unavailable.
 */
#ifndef B200JPG_H
#define B200JPG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B200JPG_API
#else
#define B200JPG_API __attribute__((visibility("default")))
#endif

/* ---- status codes: src/error.rs:37-48 (Error::{Format, Unsupported, Io, Internal}) ------------ */
enum {
    B200JPG_OK = 0,
    B200JPG_ERR_FORMAT = -1,      /* Error::Format(String)                       */
    B200JPG_ERR_UNSUPPORTED = -2, /* Error::Unsupported(UnsupportedFeature)      */
    B200JPG_ERR_IO = -3,          /* Error::Io (truncated input)                 */
    B200JPG_ERR_INTERNAL = -4     /* Error::Internal: CUDA failure, contract violations the
                                     reference turns into assert!/panic (src/worker/rayon.rs:85,
                                     src/worker/immediate.rs:31,47)               */
};

/* ---- ColorTransform: src/decoder.rs:79-98, same order ---------------------------------------- */
enum {
    B200JPG_CT_NONE = 0,
    B200JPG_CT_UNKNOWN = 1,
    B200JPG_CT_GRAYSCALE = 2,
    B200JPG_CT_RGB = 3,
    B200JPG_CT_YCBCR = 4,
    B200JPG_CT_CMYK = 5,
    B200JPG_CT_YCCK = 6,
    B200JPG_CT_JCS_BG_YCC = 7,
    B200JPG_CT_JCS_BG_RGB = 8
};

/* ---- PixelFormat: src/decoder.rs:40-49 ; CodingProcess: src/parser.rs:26-33 ------------------- */
enum { B200JPG_PF_L8 = 0, B200JPG_PF_L16 = 1, B200JPG_PF_RGB24 = 2, B200JPG_PF_CMYK32 = 3 };
enum { B200JPG_CP_DCT_SEQUENTIAL = 0, B200JPG_CP_DCT_PROGRESSIVE = 1, B200JPG_CP_LOSSLESS = 2 };

/* ---- arithmetic variant of the block kernels (the reference ships two) --------- */
enum {
    B200JPG_ARITH_SCALAR = 0, /* src/idct.rs:260-370 + src/decoder.rs:1486-1508: the
                                 `platform_independent` i32 fixed-point path (canonical, default) */
    B200JPG_ARITH_SSSE3 = 1   /* src/arch/ssse3.rs: the 16-bit saturating path an x86-64 build of
                                 the reference runs by default, emulated bit-exactly            */
};

/* b200jpg_batch_run_host takes dense coefficient buffers (the reference's Worker::append_row payload).  They are
 * mostly zeros, so host threads can compact them into sparse block streams (see "Sparse block streams" below)
 * before the upload: PCIe then carries ~1/5 of the bytes and the download of the pixels gets the link to itself.
 * Whether that pays depends on what bounds the host side: PCIe (it does) or host DRAM bandwidth (it does not: the
 * CPU then reads what the DMA engine would have read).  AUTO therefore measures: a batch object that is run
 * repeatedly uploads densely on runs 1-2, compacts on runs 3-4 (only with >= 8 CPUs and a sample of the batch
 * < 35 % non-zero) and keeps whichever second run was faster. */
enum { B200JPG_COMPACT_AUTO = 0, B200JPG_COMPACT_OFF = 1, B200JPG_COMPACT_ON = 2 };

/* Where the Huffman entropy decoding of b200jpg_decode_files happens.  The reference decodes a scan with one
 * sequential loop per image (src/huffman.rs, src/decoder.rs:1086-1172).  DEVICE: complete baseline single-scan images
 * (with or without restart intervals) are decoded by the GPU (self-synchronising parallel Huffman decoding, csrc/entropy_dev.h);
 * host threads then only parse markers and copy the scan bytes.  Every other image, and every image whose scan the
 * device flags as irregular in any way, is decoded by the host loop, so results and errors never differ. */
enum { B200JPG_ENTROPY_AUTO = 0, B200JPG_ENTROPY_HOST = 1, B200JPG_ENTROPY_DEVICE = 2 };

/* Dense coefficients -> pixels in one kernel (KF, csrc/kf_fused.cu: the planes are staged in shared memory and never
 * written to HBM) for 3-component YCbCr 4:2:0 / 4:4:4 images at full IDCT size in scalar arithmetic, whenever both
 * stages are run in one call.  Same bytes either way; which route is faster was measured on B200:
 * AUTO fuses 4:4:4 (where K2 alone is HBM-bound) and runs K1 then K2 for 4:2:0 (both routes issue the same number of
 * instructions and are bound by integer issue, the fused one pays for its CTA-wide barriers); ON fuses everything
 * eligible; OFF never fuses. */
enum { B200JPG_FUSE_AUTO = 0, B200JPG_FUSE_OFF = 1, B200JPG_FUSE_ON = 2 };

/* kernel selection, for tests and profiling (0 = pick the fastest applicable kernel) */
enum { B200JPG_KERNEL_AUTO = 0, B200JPG_KERNEL_GENERIC = 1, B200JPG_KERNEL_FAST = 2 };

/* ---- parser::Component: src/parser.rs:77-89 (geometry from update_component_sizes, 292-310) --- */
typedef struct {
    uint8_t identifier;
    uint8_t h;  /* horizontal_sampling_factor 1..4 */
    uint8_t v;  /* vertical_sampling_factor 1..4   */
    uint8_t tq; /* quantization_table_index 0..3   */
    uint16_t dct_scale; /* 1, 2, 4 or 8 */
    uint16_t size_w, size_h;   /* component.size (samples)          */
    uint16_t block_w, block_h; /* component.block_size (8x8 blocks) */
} b200jpg_component;

typedef struct {
    int device;       /* CUDA device ordinal                                              */
    int arith;        /* B200JPG_ARITH_*                                                  */
    int k1_kernel;    /* B200JPG_KERNEL_* for dequant+IDCT                                 */
    int k2_kernel;    /* B200JPG_KERNEL_* for upsample+colour                              */
    void *stream;     /* cudaStream_t to enqueue on; NULL = the context creates its own    */
    int host_compact; /* b200jpg_batch_run_host: B200JPG_COMPACT_* (0 = decide per batch)   */
    int host_threads; /* host threads of the host-fed paths; 0 = one per CPU of the process */
    int entropy;      /* b200jpg_decode_files: B200JPG_ENTROPY_* (0 = device where it applies)    */
    int fuse;         /* B200JPG_FUSE_*: one fused kernel (planes never leave the SM) where it applies, or K1 + K2 */
} b200jpg_options;

typedef struct b200jpg_ctx b200jpg_ctx;

B200JPG_API void b200jpg_default_options(b200jpg_options *opt);
/* One context per device (and per host thread that wants its own stream). */
B200JPG_API int b200jpg_create(const b200jpg_options *opt, b200jpg_ctx **ctx);
B200JPG_API void b200jpg_destroy(b200jpg_ctx *ctx);
B200JPG_API const char *b200jpg_last_error(const b200jpg_ctx *ctx);
B200JPG_API const char *b200jpg_version(void);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
B200JPG_API uint64_t b200jpg_launch_count(const b200jpg_ctx *ctx);
B200JPG_API int b200jpg_synchronize(b200jpg_ctx *ctx);
/* changes b200jpg_options.fuse of a live context (B200JPG_FUSE_*): lets one plan be run both ways, e.g. to time the
 * two-kernel route next to the fused kernel on the same buffers */
B200JPG_API void b200jpg_set_fuse(b200jpg_ctx *ctx, int fuse);
/* b200jpg_decode_files since creation: scans whose Huffman decoding ran on the device, and how many of those the
 * device flagged and handed back to the host loop (see B200JPG_ENTROPY_*) */
B200JPG_API void b200jpg_device_scan_counts(const b200jpg_ctx *ctx, uint64_t *decoded, uint64_t *retried);

/* parser::update_component_sizes, src/parser.rs:292-310: fills size_* / block_* of comps[] from the
 * sampling factors and dct_scale already set in them. */
B200JPG_API int b200jpg_update_component_sizes(uint16_t width, uint16_t height, b200jpg_component *comps,
                                               int ncomp, uint16_t *mcu_w, uint16_t *mcu_h);
/* idct::choose_idct_size, src/idct.rs:14-28 */
B200JPG_API int b200jpg_choose_idct_size(uint16_t full_w, uint16_t full_h, uint16_t req_w, uint16_t req_h);

/* ============================================================================================
 * Worker-shaped API: trait Worker, src/worker/mod.rs:24-35 (impls src/worker/immediate.rs,
 * src/worker/rayon.rs, src/worker/multithreaded.rs), obtained per decode like WorkerScope
 * (src/worker/mod.rs:44-95).  Single-threaded use per worker, like the RefCell'd scope.
 * ========================================================================================== */
typedef struct b200jpg_worker b200jpg_worker;
B200JPG_API int b200jpg_worker_new(b200jpg_ctx *ctx, b200jpg_worker **w);
B200JPG_API void b200jpg_worker_free(b200jpg_worker *w);
/* Worker::start(RowData{index, component, quantization_table}), src/worker/mod.rs:18-25: allocates
 * the zeroed plane block_w*block_h*dct_scale^2 on the device and resets the row offset.
 * qt_natural: 64 entries in natural (de-zigzagged) order, src/decoder.rs:490-496. */
B200JPG_API int b200jpg_worker_start(b200jpg_worker *w, int index, const b200jpg_component *c,
                                     const uint16_t qt_natural[64]);
/* Worker::append_row((index, Vec<i16>)), src/worker/mod.rs:26: one MCU row of one component,
 * n_i16 must be block_w * v * 64 (assert at src/worker/immediate.rs:47).  The coefficients are
 * copied before returning; the IDCT itself is deferred to get_result (one launch per component). */
B200JPG_API int b200jpg_worker_append_row(b200jpg_worker *w, int index, const int16_t *coefs, size_t n_i16);
/* Worker::append_rows, src/worker/mod.rs:29-34: `nrows` consecutive MCU rows, contiguous. */
B200JPG_API int b200jpg_worker_append_rows(b200jpg_worker *w, int index, const int16_t *coefs, size_t n_i16,
                                           size_t nrows);
/* Worker::get_result(index) -> Vec<u8>, src/worker/mod.rs:27: runs the IDCT kernel over everything
 * appended and copies the plane to the host (plane_out may be NULL to leave it on the device
 * only; *plane_len receives the plane size either way).  The device plane stays alive for
 * b200jpg_worker_compute_image. */
B200JPG_API int b200jpg_worker_get_result(b200jpg_worker *w, int index, uint8_t *plane_out, size_t cap,
                                          size_t *plane_len);
/* decoder::compute_image over the planes still resident in the worker (indices 0..ncomp-1):
 * src/decoder.rs:1300-1336 without the host round trip of the planes. */
B200JPG_API int b200jpg_worker_compute_image(b200jpg_worker *w, int ncomp, uint16_t out_w, uint16_t out_h,
                                             int color_transform, uint8_t *out, size_t cap, size_t *out_len);

/* decoder::compute_image(components, data: Vec<Vec<u8>>, output_size, color_transform),
 * src/decoder.rs:1300-1336 -> worker::compute_image_parallel, src/worker/mod.rs:97-128 /
 * src/worker/rayon.rs:193-219: host planes in, interleaved pixels out. */
B200JPG_API int b200jpg_compute_image(b200jpg_ctx *ctx, const b200jpg_component *comps, int ncomp,
                                      const uint8_t *const *planes, const size_t *plane_len, uint16_t out_w,
                                      uint16_t out_h, int color_transform, uint8_t *out, size_t cap,
                                      size_t *out_len);

/* ============================================================================================
 * Batched hot path: n independent images, dense coefficients -> pixels.  New surface (the
 * reference has no batching); per image it is exactly start + append_rows +
 * get_result per component followed by compute_image.
 * ========================================================================================== */
typedef struct {
    uint16_t width, height; /* output size (frame.output_size, src/parser.rs:50-61) */
    uint8_t ncomp;          /* 1, 3 or 4 */
    uint8_t color_transform;
    uint16_t reserved;
    b200jpg_component comps[4];
    const uint16_t *qt[4];   /* natural order, 64 entries each */
    const int16_t *coefs[4]; /* host: block_w*block_h*64 i16 per component, blocks in raster order,
                                coefficients in natural order (src/decoder.rs:962-967).  May be NULL
                                when the plan is only used with device-resident coefficients. */
} b200jpg_image_desc;

typedef struct b200jpg_batch b200jpg_batch;
typedef struct {
    size_t coef_bytes;  /* size of the coefficient slab (device) */
    size_t plane_bytes; /* size of the plane slab (device)       */
    size_t out_bytes;   /* size of the pixel slab (device)       */
    size_t n_blocks;    /* 8x8 blocks over all components        */
    size_t n_pixels;    /* sum of width*height                   */
    size_t k1_algorithmic_bytes; /* 128 B read + dct_scale^2 B written per block */
    size_t k2_algorithmic_bytes; /* plane bytes read once + pixel bytes written  */
    size_t kf_algorithmic_bytes; /* fused kernel: 128 B read per block + pixel bytes written (planes stay on chip) */
    size_t n_fused;              /* images the fused kernel takes when both stages run in one call */
} b200jpg_batch_info;

/* Validates every image (same errors as the reference: NonIntegerSubsamplingRatio
 * src/upsampler.rs:93-98, invalid (ncomp, transform) src/decoder.rs:1344-1386), lays out the slabs
 * and uploads the per-component tables.  statuses[i] (optional) receives the per-image code; a bad
 * image does not poison the batch, it is skipped. */
B200JPG_API int b200jpg_batch_create(b200jpg_ctx *ctx, const b200jpg_image_desc *imgs, size_t n, int *statuses,
                                     b200jpg_batch **batch);
B200JPG_API void b200jpg_batch_free(b200jpg_batch *b);
B200JPG_API int b200jpg_batch_get_info(const b200jpg_batch *b, b200jpg_batch_info *info);
/* byte offsets of image i inside the slabs */
B200JPG_API int b200jpg_batch_image_layout(const b200jpg_batch *b, size_t i, size_t coef_off[4], size_t plane_off[4],
                                           size_t *out_off, size_t *out_len);
/* Device-resident run: K1 over d_coefs -> d_planes, then K2 over d_planes -> d_out, enqueued on the
 * context's stream (no synchronisation).  stages: bit0 = K1, bit1 = K2. */
B200JPG_API int b200jpg_batch_run_device(b200jpg_batch *b, const void *d_coefs, void *d_planes, void *d_out,
                                         int stages);
/* On-device output formats for GPU-side consumers: converts the interleaved RGB8 pixel slab a
 * device-resident run produced (d_out of b200jpg_batch_run_device, image i at its out_off) into the layout a training
 * or inference input pipeline wants, image by image at the same relative offsets (out_off for the 8-bit layout,
 * 4 * out_off bytes for the float layouts).  Float samples are value * scale[c] + bias[c] (NULL = 1 and 0): exact for
 * scale 1 / bias 0, otherwise one fp32 FMA per sample.  Only 3-component images; others are skipped (status ERR_FORMAT
 * in statuses[i] when given).  Enqueued on the context's stream, no synchronisation. */
enum {
    B200JPG_FMT_RGB8_PLANAR = 1,  /* CHW uint8: three W*H planes per image  */
    B200JPG_FMT_RGB_F32_NHWC = 2, /* HWC float32 (interleaved)              */
    B200JPG_FMT_RGB_F32_NCHW = 3  /* CHW float32: three W*H planes per image */
};
B200JPG_API int b200jpg_batch_format_device(b200jpg_batch *b, const void *d_out, int format, void *d_dst, const float scale[3],
                                            const float bias[3], int *statuses);
/* Host-to-host run: the coefficients named in the descs -> pixels in outs[i] (out_caps[i] bytes).  Either the
 * dense buffers are uploaded as they are (H2D, K1, K2, D2H, chunked and double-buffered over two streams through
 * internal device slabs), or -- b200jpg_options.host_compact -- host threads first compact them into sparse block
 * streams (H2D of the streams, K0, K1, K2, D2H over three streams).  Same pixels either way.  Page-locked caller
 * buffers make the copies truly asynchronous. */
B200JPG_API int b200jpg_batch_run_host(b200jpg_batch *b, const b200jpg_image_desc *imgs, uint8_t *const *outs,
                                       const size_t *out_caps, int *statuses);
/* convenience: create + run_host + free */
B200JPG_API int b200jpg_decode_batch(b200jpg_ctx *ctx, const b200jpg_image_desc *imgs, size_t n,
                                     uint8_t *const *outs, const size_t *out_caps, int *statuses);

/* Profiling knob, not part of the drop-in surface: selects bit-identical code-generation variants of
 * the hot kernels (K1: output-butterfly style 0 = IADD3 / 5 = IMAD; K2 4:2:0: row pairs per thread 1 / 4);
 * -1 = default.  Used by scripts/sweep_kernels.py to produce profiles/sweep_*.jsonl. */
B200JPG_API void b200jpg_debug_set_kernel_modes(int k1_mode, int k2_mode);

/* page-locked host memory helpers (cudaHostAlloc / cudaFreeHost) */
B200JPG_API void *b200jpg_host_alloc(size_t bytes);
B200JPG_API void b200jpg_host_free(void *p);

/* ============================================================================================
 * Whole-file decoder: Decoder<R>, src/decoder.rs:101-295.  Marker parsing and Huffman decoding
 * (src/parser.rs, src/huffman.rs, src/decoder.rs:297-1298) run on the host in C++; the worker
 * path runs on the GPU through the API above.
 * ========================================================================================== */
typedef struct b200jpg_decoder b200jpg_decoder;
typedef struct {
    uint16_t width, height;
    int pixel_format;   /* B200JPG_PF_* */
    int coding_process; /* B200JPG_CP_* */
} b200jpg_image_info; /* ImageInfo, src/decoder.rs:64-74 */

/* Decoder::new(reader), src/decoder.rs:134: `data` must stay valid for the decoder's lifetime. */
B200JPG_API int b200jpg_decoder_new(b200jpg_ctx *ctx, const uint8_t *data, size_t len, b200jpg_decoder **d);
B200JPG_API void b200jpg_decoder_free(b200jpg_decoder *d);
B200JPG_API void b200jpg_decoder_set_color_transform(b200jpg_decoder *d, int color_transform); /* :158 */
B200JPG_API void b200jpg_decoder_set_max_decoding_buffer_size(b200jpg_decoder *d, size_t max); /* :163 */
B200JPG_API int b200jpg_decoder_read_info(b200jpg_decoder *d);                                  /* :265 */
/* Decoder::info(), :171 -- returns 1 and fills *info once read_info/decode succeeded, else 0 */
B200JPG_API int b200jpg_decoder_info(const b200jpg_decoder *d, b200jpg_image_info *info);
B200JPG_API int b200jpg_decoder_scale(b200jpg_decoder *d, uint16_t req_w, uint16_t req_h, uint16_t *w,
                                      uint16_t *h); /* :278 */
/* Decoder::decode(), :293 -- pixels are owned by the decoder until free / next decode.  Complete baseline scans are
 * Huffman-decoded on the GPU (see B200JPG_ENTROPY_*) unless a colour transform, a size limit or a scale was set;
 * everything else, and anything the device hands back, takes the host loop -- same pixels, same errors. */
B200JPG_API int b200jpg_decoder_decode(b200jpg_decoder *d, const uint8_t **pixels, size_t *len);
B200JPG_API const char *b200jpg_decoder_error(const b200jpg_decoder *d);
/* icc_profile :211, exif_data :199, xmp_data :206 -- return 1 if present */
B200JPG_API int b200jpg_decoder_icc_profile(b200jpg_decoder *d, const uint8_t **data, size_t *len);
B200JPG_API int b200jpg_decoder_exif_data(const b200jpg_decoder *d, const uint8_t **data, size_t *len);
B200JPG_API int b200jpg_decoder_xmp_data(const b200jpg_decoder *d, const uint8_t **data, size_t *len);
/* Host half only (no GPU needed): marker parsing + entropy decoding into dense coefficient
 * buffers, i.e. everything decode() does before the worker boundary.  Fills *desc with pointers
 * owned by the decoder (valid until free / next call); desc->color_transform is
 * determine_color_transform() (src/decoder.rs:698-764). */
B200JPG_API int b200jpg_decoder_entropy_decode(b200jpg_decoder *d, b200jpg_image_desc *desc);

/* ============================================================================================
 * Sparse block streams ("sparse coefficient wire format").  The dense Vec<i16> the
 * reference pushes through Worker::append_row (src/worker/mod.rs:26, src/decoder.rs:962-983) is 128 B per
 * block of mostly zeros; a stream carries, per block, a 63-bit map of the non-zero AC coefficients (zig-zag
 * order), the DC coefficient, and the non-zero values as int8 / int16 -- what Huffman decoding produced
 * (csrc/sbs.h has the byte layout).  Kernel K0 rebuilds the dense slab on the device, bit for bit.
 * ========================================================================================== */
enum {
    B200JPG_SBS_PLANAR = 0,      /* blocks: component by component, raster order                          */
    B200JPG_SBS_INTERLEAVED = 1, /* blocks: MCU by MCU as in an interleaved scan (src/decoder.rs:978-983) */
    B200JPG_SBS_NATURAL = 2      /* flag: bitmaps / values in natural coefficient order, not zig-zag      */
};
typedef struct {
    const uint8_t *data; /* host memory, ideally page-locked */
    size_t len;
    int order;           /* B200JPG_SBS_* */
} b200jpg_sbs_stream;
/* upper bound of the stream length of an image with `nblocks` 8x8 blocks over all components */
B200JPG_API size_t b200jpg_sbs_worst_bytes(size_t nblocks);
/* number of 8x8 blocks over all components (sum of block_w*block_h); reads the headers if necessary */
B200JPG_API int b200jpg_decoder_total_blocks(b200jpg_decoder *d, size_t *nblocks);
/* Like b200jpg_decoder_entropy_decode, but the coefficients are written as a sparse stream into buf (cap >=
 * b200jpg_sbs_worst_bytes); desc->coefs stay NULL.  Host only. */
B200JPG_API int b200jpg_decoder_entropy_decode_sbs(b200jpg_decoder *d, uint8_t *buf, size_t cap,
                                                   b200jpg_image_desc *desc, b200jpg_sbs_stream *stream);
/* Host only: compacts the dense coefficients desc->coefs of one image into a stream (PLANAR | NATURAL) in buf
 * (cap >= b200jpg_sbs_worst_bytes) -- what b200jpg_batch_run_host does on its host threads when compacting. */
B200JPG_API int b200jpg_sbs_from_dense(const b200jpg_image_desc *img, uint8_t *buf, size_t cap,
                                       b200jpg_sbs_stream *stream);
/* b200jpg_decode_batch with the coefficients given as streams: H2D -> K0 expand -> K1 -> K2 -> D2H,
 * pipelined over three CUDA streams in groups of 32 images.  Streams are validated before use. */
B200JPG_API int b200jpg_decode_batch_sbs(b200jpg_ctx *ctx, const b200jpg_image_desc *imgs,
                                         const b200jpg_sbs_stream *streams, size_t n, uint8_t *const *outs,
                                         const size_t *out_caps, int *statuses);
/* test hook: runs one image through the stream path and copies the dense slab K0 produced back to the host
 * (dense_out[c]: block_w*block_h*64 int16 of component c, or NULL to skip) */
B200JPG_API int b200jpg_debug_expand_sbs(b200jpg_ctx *ctx, const b200jpg_image_desc *img,
                                         const b200jpg_sbs_stream *stream, int16_t *const dense_out[4]);

/* ============================================================================================
 * Whole-file batches: what an outer par_iter over Decoder::decode() gives a
 * user of the reference, with the host threads doing only marker parsing + Huffman decoding
 * (src/parser.rs, src/huffman.rs, src/decoder.rs:794-1298) and the worker path on the GPU.
 * ========================================================================================== */
typedef struct {
    const uint8_t *data; /* in: the JPEG file                                   */
    size_t len;
    uint8_t *out;        /* in: pixel buffer: host memory (ideally page-locked) or DEVICE memory of the context's
                            GPU (the pixels then never cross PCIe); unused by read_info_files                  */
    size_t out_cap;
    b200jpg_image_info info; /* out */
    size_t out_len;      /* out: width*height*ncomp                              */
    int status;          /* out: B200JPG_OK or the image's error                 */
} b200jpg_file_job;

/* Headers only (Decoder::read_info + info for each file), nthreads host threads (0 = all). */
B200JPG_API int b200jpg_read_info_files(b200jpg_file_job *jobs, size_t n, int nthreads);
/* Full decode of n files; per-image errors are reported in jobs[i].status and do not stop the batch. */
B200JPG_API int b200jpg_decode_files(b200jpg_ctx *ctx, b200jpg_file_job *jobs, size_t n, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
