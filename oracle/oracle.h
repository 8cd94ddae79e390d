/*
 * oracle.h -- CPU restatement of image-rs/jpeg-decoder v0.3.2 (reference @ /root/reference).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
 * call it, and only as the checker (or the timed CPU baseline), never as a fallback for the
 * CUDA path.  The product library (jpeg_decoder_b200/) does not link or include this directory.
 *
 * Parity status: PINNED against (a) the three in-source known-answer tests of the reference
 * (src/idct.rs:580-657), (b) the geometry test src/parser.rs:312-329, (c) every DCT-based
 * golden PNG under the reference's tests/reftest/images (tolerance +-3, the reference's own
 * bound, tests/reftest/mod.rs:99) -- see tests/test_oracle_*.py.  The reference itself (Rust)
 * cannot be compiled in this image (no cargo/rustc), so there is no oracle/_ref.
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef B200JPG_ORACLE_H
#define B200JPG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Arithmetic variants of the reference hot path. */
enum {
    ORC_ARITH_SCALAR = 0, /* src/idct.rs:260-370 + src/decoder.rs:1486-1508 (feature platform_independent) */
    ORC_ARITH_SSSE3 = 1,  /* src/arch/ssse3.rs (x86 default build), portable lane-by-lane emulation */
    ORC_ARITH_SSSE3_NATIVE = 2 /* the same arithmetic through <tmmintrin.h>, i.e. the instructions an x86-64 build
                                  of the reference executes (src/arch/mod.rs:13-57 selects them at run time);
                                  identical bytes to ORC_ARITH_SSSE3 (tests/test_oracle_kat.py); the timed CPU
                                  baseline.  Falls back to the emulation when built without -mssse3. */
};

/* src/decoder.rs:79-98, same order as the Rust enum */
enum {
    ORC_CT_NONE = 0,
    ORC_CT_UNKNOWN = 1,
    ORC_CT_GRAYSCALE = 2,
    ORC_CT_RGB = 3,
    ORC_CT_YCBCR = 4,
    ORC_CT_CMYK = 5,
    ORC_CT_YCCK = 6,
    ORC_CT_JCS_BG_YCC = 7,
    ORC_CT_JCS_BG_RGB = 8
};

/* src/error.rs:37-48 */
enum { ORC_OK = 0, ORC_ERR_FORMAT = 1, ORC_ERR_UNSUPPORTED = 2, ORC_ERR_IO = 3, ORC_ERR_INTERNAL = 4 };

/* src/parser.rs:77-89 */
typedef struct {
    uint8_t identifier;
    uint8_t h;  /* horizontal_sampling_factor */
    uint8_t v;  /* vertical_sampling_factor */
    uint8_t tq; /* quantization_table_index */
    uint16_t dct_scale;
    uint16_t size_w, size_h;   /* component.size */
    uint16_t block_w, block_h; /* component.block_size */
} orc_component;

/* ---- block kernels (src/idct.rs, src/arch/ssse3.rs) ------------------------------------ */
void orc_idct_block(int arith, int scale, const int16_t c[64], const uint16_t q[64],
                    size_t stride, uint8_t *out);
/* real <tmmintrin.h> implementation, only when compiled with -mssse3; returns 0 if unavailable */
int orc_idct8x8_ssse3_intrin(const int16_t c[64], const uint16_t q[64], size_t stride, uint8_t *out);
int orc_ycbcr_line_ssse3_intrin(const uint8_t *y, const uint8_t *cb, const uint8_t *cr, uint8_t *out,
                                size_t npix, size_t *done);

/* src/idct.rs:14-28 */
int orc_choose_idct_size(uint16_t full_w, uint16_t full_h, uint16_t req_w, uint16_t req_h);
/* src/parser.rs:292-310 ; returns ORC_OK / ORC_ERR_FORMAT */
int orc_update_component_sizes(uint16_t w, uint16_t h, orc_component *comps, int n, uint16_t *mcu_w,
                               uint16_t *mcu_h);

/* ---- worker (src/worker/mod.rs:24-35, src/worker/immediate.rs) ----------------------------- */
typedef struct orc_worker orc_worker;
orc_worker *orc_worker_new(int arith);
void orc_worker_free(orc_worker *w);
int orc_worker_start(orc_worker *w, int index, const orc_component *c, const uint16_t qt[64]);
int orc_worker_append_row(orc_worker *w, int index, const int16_t *coefs, size_t n);
/* moves the plane out (mem::take): caller owns *plane (free()) */
int orc_worker_get_result(orc_worker *w, int index, uint8_t **plane, size_t *len);

/* ---- image assembly (src/decoder.rs:1300-1508, src/upsampler.rs, src/worker/mod.rs:97-128) -- */
/* planes[i] has plane_len[i] bytes.  out must hold out_w*out_h*ncomp bytes.  */
int orc_compute_image(int arith, const orc_component *comps, int ncomp, const uint8_t *const *planes,
                      const size_t *plane_len, uint16_t out_w, uint16_t out_h, int color_transform,
                      uint8_t *out, size_t cap, size_t *out_len);
/* one output row: upsample every component then colour convert (src/upsampler.rs:47-63) */
int orc_upsample_and_interleave_row(int arith, const orc_component *comps, int ncomp,
                                    const uint8_t *const *planes, uint16_t out_w, uint16_t out_h,
                                    int color_transform, size_t row, uint8_t *out_row);
const char *orc_last_error(void);

/* ---- hot path, whole image from dense coefficients (start + append_row* + get_result +
 *      compute_image), used as the timed CPU baseline -------------------------------------- */
int orc_hotpath_image(int arith, const orc_component *comps, int ncomp, const uint16_t *const qts[4],
                      const int16_t *const coefs[4], uint16_t out_w, uint16_t out_h, int color_transform,
                      uint8_t *out, size_t cap);
/* Single-image latency shape of the reference's rayon build: IDCT inline on the calling thread
 * (src/worker/rayon.rs:121-132), then the output rows split over nthreads threads (one rayon task per
 * row, src/worker/rayon.rs:204-216). */
int orc_hotpath_image_mt(int arith, int nthreads, const orc_component *comps, int ncomp,
                         const uint16_t *const qts[4], const int16_t *const coefs[4], uint16_t out_w,
                         uint16_t out_h, int color_transform, uint8_t *out, size_t cap);
/* compute_image with the rows split over nthreads threads (compute_image_parallel, src/worker/rayon.rs:193-219) */
int orc_compute_image_mt(int arith, int nthreads, const orc_component *comps, int ncomp,
                         const uint8_t *const *planes, const size_t *plane_len, uint16_t out_w, uint16_t out_h,
                         int color_transform, uint8_t *out, size_t cap, size_t *out_len);
/* n images with identical geometry spread over nthreads pthreads (one image per thread at a
 * time: what an outer par_iter over Decoder::decode gives the reference). */
int orc_hotpath_batch(int arith, int nthreads, size_t n, const orc_component *comps, int ncomp,
                      const uint16_t *const qts[4], const int16_t *const *coefs /* n*ncomp */,
                      uint16_t out_w, uint16_t out_h, int color_transform, uint8_t *const *outs,
                      size_t cap);

/* ---- whole-file decoder (src/decoder.rs:101-1298, src/parser.rs, src/huffman.rs, src/marker.rs) */
typedef struct orc_decoder orc_decoder;
typedef struct {
    uint16_t width, height;
    int pixel_format;   /* 0 L8, 1 L16, 2 RGB24, 3 CMYK32 (src/decoder.rs:40-49) */
    int coding_process; /* 0 DctSequential, 1 DctProgressive, 2 Lossless (src/parser.rs:26-33) */
} orc_image_info;

orc_decoder *orc_decoder_new(const uint8_t *data, size_t len, int arith);
void orc_decoder_free(orc_decoder *d);
void orc_decoder_set_color_transform(orc_decoder *d, int ct);
void orc_decoder_set_max_decoding_buffer_size(orc_decoder *d, size_t max);
/* threads compute_image may use (default 1; > 1 = the rayon build's row-parallel colour stage) */
void orc_decoder_set_threads(orc_decoder *d, int nthreads);
/* test taps (coefficient / plane copies kept for orc_decoder_coefficients / _plane): on by default, off when timing */
void orc_decoder_set_taps(orc_decoder *d, int on);
int orc_decoder_read_info(orc_decoder *d);
int orc_decoder_info(const orc_decoder *d, orc_image_info *info); /* 1 if available */
int orc_decoder_scale(orc_decoder *d, uint16_t req_w, uint16_t req_h, uint16_t *w, uint16_t *h);
/* pixels owned by the decoder until free / next decode */
int orc_decoder_decode(orc_decoder *d, const uint8_t **pixels, size_t *len);
const char *orc_decoder_error(const orc_decoder *d);
int orc_decoder_color_transform(const orc_decoder *d); /* determine_color_transform, after decode */
/* test taps: what the decoder pushed through the worker boundary, per frame component */
int orc_decoder_ncomp(const orc_decoder *d);
int orc_decoder_component(const orc_decoder *d, int i, orc_component *c, uint16_t qt[64]);
int orc_decoder_coefficients(const orc_decoder *d, int i, const int16_t **coefs, size_t *n_i16);
int orc_decoder_plane(const orc_decoder *d, int i, const uint8_t **plane, size_t *len);
int orc_decoder_icc_profile(const orc_decoder *d, const uint8_t **data, size_t *len);
int orc_decoder_exif(const orc_decoder *d, const uint8_t **data, size_t *len);
int orc_decoder_xmp(const orc_decoder *d, const uint8_t **data, size_t *len);

#ifdef __cplusplus
}
#endif
#endif
