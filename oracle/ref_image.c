/*
 * ref_image.c -- oracle (test infrastructure, see oracle.h): worker planes -> interleaved pixels,
 * restating /root/reference/src/worker/immediate.rs, src/upsampler.rs, src/decoder.rs:1300-1508,
 * src/worker/mod.rs:97-128 and the SSSE3 colour path src/arch/ssse3.rs:196-288.
 */
#include "oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }
static int fail(int code, const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

/* ---------------------------------------------------------------------------------------------
 * Worker: src/worker/immediate.rs:11-64 (the rayon Scoped worker, src/worker/rayon.rs:71-132,
 * computes the same bytes through an 8x8 temporary).
 * ------------------------------------------------------------------------------------------ */
struct orc_worker {
    int arith;
    size_t offsets[4];
    uint8_t *results[4];
    size_t result_len[4];
    int has_component[4];
    orc_component components[4];
    uint16_t qt[4][64];
};

orc_worker *orc_worker_new(int arith) {
    orc_worker *w = (orc_worker *)calloc(1, sizeof *w);
    if (w) w->arith = arith;
    return w;
}
void orc_worker_free(orc_worker *w) {
    if (!w) return;
    for (int i = 0; i < 4; i++) free(w->results[i]);
    free(w);
}
/* src/worker/immediate.rs:30-37 */
int orc_worker_start(orc_worker *w, int index, const orc_component *c, const uint16_t qt[64]) {
    if (index < 0 || index >= 4) return fail(ORC_ERR_INTERNAL, "worker index out of range");
    if (w->results[index]) return fail(ORC_ERR_INTERNAL, "worker started twice (assert results.is_empty())");
    size_t len = (size_t)c->block_w * c->block_h * c->dct_scale * c->dct_scale;
    w->offsets[index] = 0;
    w->results[index] = (uint8_t *)calloc(len ? len : 1, 1);
    w->result_len[index] = len;
    w->components[index] = *c;
    w->has_component[index] = 1;
    memcpy(w->qt[index], qt, 128);
    return ORC_OK;
}
/* src/worker/immediate.rs:39-60 */
int orc_worker_append_row(orc_worker *w, int index, const int16_t *coefs, size_t n) {
    if (index < 0 || index >= 4 || !w->has_component[index] || !w->results[index])
        return fail(ORC_ERR_INTERNAL, "append_row on a component that was not started");
    const orc_component *c = &w->components[index];
    size_t block_count = (size_t)c->block_w * c->v;
    size_t line_stride = (size_t)c->block_w * c->dct_scale;
    if (n != block_count * 64) return fail(ORC_ERR_INTERNAL, "append_row: data.len() != block_count*64");
    size_t used = block_count * c->dct_scale * c->dct_scale;
    if (w->offsets[index] + used > w->result_len[index])
        return fail(ORC_ERR_INTERNAL, "append_row: more rows than the plane holds");
    for (size_t i = 0; i < block_count; i++) {
        size_t x = (i % c->block_w) * c->dct_scale;
        size_t y = (i / c->block_w) * c->dct_scale;
        orc_idct_block(w->arith, c->dct_scale, coefs + i * 64, w->qt[index], line_stride,
                       w->results[index] + w->offsets[index] + y * line_stride + x);
    }
    w->offsets[index] += used;
    return ORC_OK;
}
/* src/worker/immediate.rs:62-64 */
int orc_worker_get_result(orc_worker *w, int index, uint8_t **plane, size_t *len) {
    if (index < 0 || index >= 4) return fail(ORC_ERR_INTERNAL, "worker index out of range");
    *plane = w->results[index];
    *len = w->results[index] ? w->result_len[index] : 0;
    w->results[index] = NULL;
    w->result_len[index] = 0;
    return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Upsampling: src/upsampler.rs
 * ------------------------------------------------------------------------------------------ */
enum { UP_H1V1, UP_H2V1, UP_H1V2, UP_H2V2, UP_GENERIC };
typedef struct {
    int kind;
    int hs, vs; /* generic scaling factors */
    size_t width, height, row_stride;
} upcomp;

/* src/upsampler.rs:76-105 */
static int choose_upsampler(uint8_t h, uint8_t v, uint8_t hmax, uint8_t vmax, uint16_t out_w, uint16_t out_h,
                            upcomp *u) {
    int h1 = h == hmax || out_w == 1;
    int v1 = v == vmax || out_h == 1;
    int h2 = h * 2 == hmax;
    int v2 = v * 2 == vmax;
    u->hs = u->vs = 1;
    if (h1 && v1) u->kind = UP_H1V1;
    else if (h2 && v1) u->kind = UP_H2V1;
    else if (h1 && v2) u->kind = UP_H1V2;
    else if (h2 && v2) u->kind = UP_H2V2;
    else if (hmax % h != 0 || vmax % v != 0)
        return fail(ORC_ERR_UNSUPPORTED, "NonIntegerSubsamplingRatio");
    else {
        u->kind = UP_GENERIC;
        u->hs = hmax / h;
        u->vs = vmax / v;
    }
    return ORC_OK;
}

typedef struct {
    upcomp c[4];
    int n;
    size_t line_buffer_size;
} upsampler;

/* src/upsampler.rs:20-45 */
static int upsampler_new(const orc_component *comps, int n, uint16_t out_w, uint16_t out_h, upsampler *u) {
    uint8_t hmax = 0, vmax = 0;
    size_t wmax = 0;
    for (int i = 0; i < n; i++) {
        if (comps[i].h > hmax) hmax = comps[i].h;
        if (comps[i].v > vmax) vmax = comps[i].v;
        if (comps[i].size_w > wmax) wmax = comps[i].size_w;
    }
    for (int i = 0; i < n; i++) {
        int rc = choose_upsampler(comps[i].h, comps[i].v, hmax, vmax, out_w, out_h, &u->c[i]);
        if (rc) return rc;
        u->c[i].width = comps[i].size_w;
        u->c[i].height = comps[i].size_h;
        u->c[i].row_stride = (size_t)comps[i].block_w * comps[i].dct_scale;
    }
    u->n = n;
    u->line_buffer_size = wmax * hmax;
    return ORC_OK;
}

/* src/upsampler.rs:174-180 / 200-206 -- the f32 row selection, kept in f32 on purpose */
static void near_far(size_t row, size_t input_height, size_t *near, size_t *far) {
    float row_near = (float)row / 2.0f;
    float fr = row_near - (float)(size_t)row_near; /* fract() of a non-negative value */
    float row_far = row_near + fr * 3.0f - 0.25f;
    float lim = (float)(input_height - 1);
    if (row_far > lim) row_far = lim;
    *near = (size_t)row_near;
    *far = row_far < 0.0f ? 0 : (size_t)row_far; /* Rust float->usize cast saturates */
}

static void upsample_row(const upcomp *u, const uint8_t *input, size_t row, size_t output_width, uint8_t *out) {
    size_t iw = u->width;
    switch (u->kind) {
    case UP_H1V1: /* src/upsampler.rs:119-132 */
        memcpy(out, input + row * u->row_stride, output_width);
        break;
    case UP_H2V1: { /* src/upsampler.rs:134-163 */
        const uint8_t *in = input + row * u->row_stride;
        if (iw == 1) {
            out[0] = in[0];
            out[1] = in[0];
            return;
        }
        out[0] = in[0];
        out[1] = (uint8_t)(((uint32_t)in[0] * 3 + in[1] + 2) >> 2);
        for (size_t i = 1; i < iw - 1; i++) {
            uint32_t s = 3 * (uint32_t)in[i] + 2;
            out[i * 2] = (uint8_t)((s + in[i - 1]) >> 2);
            out[i * 2 + 1] = (uint8_t)((s + in[i + 1]) >> 2);
        }
        out[(iw - 1) * 2] = (uint8_t)(((uint32_t)in[iw - 1] * 3 + in[iw - 2] + 2) >> 2);
        out[(iw - 1) * 2 + 1] = in[iw - 1];
        break;
    }
    case UP_H1V2: { /* src/upsampler.rs:165-189 */
        size_t rn, rf;
        near_far(row, u->height, &rn, &rf);
        const uint8_t *n = input + rn * u->row_stride, *f = input + rf * u->row_stride;
        for (size_t i = 0; i < output_width; i++) out[i] = (uint8_t)((3 * (uint32_t)n[i] + f[i] + 2) >> 2);
        break;
    }
    case UP_H2V2: { /* src/upsampler.rs:191-228 */
        size_t rn, rf;
        near_far(row, u->height, &rn, &rf);
        const uint8_t *n = input + rn * u->row_stride, *f = input + rf * u->row_stride;
        if (iw == 1) {
            uint8_t v = (uint8_t)((3 * (uint32_t)n[0] + f[0] + 2) >> 2);
            out[0] = v;
            out[1] = v;
            return;
        }
        uint32_t t1 = 3 * (uint32_t)n[0] + f[0];
        out[0] = (uint8_t)((t1 + 2) >> 2);
        for (size_t i = 1; i < iw; i++) {
            uint32_t t0 = t1;
            t1 = 3 * (uint32_t)n[i] + f[i];
            out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        out[iw * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
        break;
    }
    default: { /* src/upsampler.rs:230-250 */
        const uint8_t *in = input + (row / (size_t)u->vs) * u->row_stride;
        size_t idx = 0;
        for (size_t i = 0; i < iw; i++)
            for (int k = 0; k < u->hs; k++) out[idx++] = in[i];
        break;
    }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Colour conversion: src/decoder.rs:1339-1508
 * ------------------------------------------------------------------------------------------ */
/* src/decoder.rs:1502-1504 : (x * (1<<20) as f32 + 0.5) as i32 in f32 */
static inline int32_t f2f20(float x) { return (int32_t)(x * 1048576.0f + 0.5f); }
static inline uint8_t clamp_fixed_point(int32_t v) {
    v >>= 20;
    return (uint8_t)(v > 255 ? 255 : (v < 0 ? 0 : v));
}
/* src/decoder.rs:1491-1500 (no overflow possible: |terms| < 2^29) */
static inline void ycbcr_to_rgb(uint8_t y8, uint8_t cb8, uint8_t cr8, uint8_t *o) {
    int32_t y = (int32_t)y8 * (1 << 20) + (1 << 19);
    int32_t cb = (int32_t)cb8 - 128, cr = (int32_t)cr8 - 128;
    o[0] = clamp_fixed_point(y + f2f20(1.40200f) * cr);
    o[1] = clamp_fixed_point(y - f2f20(0.34414f) * cb - f2f20(0.71414f) * cr);
    o[2] = clamp_fixed_point(y + f2f20(1.77200f) * cb);
}

static inline int16_t sat16(int32_t v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }
static inline int16_t mulhrs(int16_t a, int16_t b) { return (int16_t)(((((int32_t)a * b) >> 14) + 1) >> 1); }
static inline uint8_t packus(int32_t v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
/* src/arch/ssse3.rs:196-288, portable emulation; returns the number of pixels converted */
static size_t ycbcr_line_ssse3_emul(const uint8_t *y, const uint8_t *cb, const uint8_t *cr, uint8_t *out,
                                    size_t num) {
    size_t nv = num / 8;
    nv = nv ? nv - 1 : 0; /* (num / 8).saturating_sub(1), ssse3.rs:206 */
    for (size_t p = 0; p < nv * 8; p++) {
        int16_t yy = sat16(((int32_t)y[p] << 6) + 32);
        int16_t b = sat16(((int32_t)cb[p] << 6) - 8192);
        int16_t r = sat16(((int32_t)cr[p] << 6) - 8192);
        int16_t cr_140200 = sat16((int32_t)mulhrs(r, 13173) + r);
        int16_t cb_034414 = mulhrs(b, 11276);
        int16_t cr_071414 = mulhrs(r, 23401);
        int16_t cb_177200 = sat16((int32_t)mulhrs(b, 25297) + b);
        int16_t R = sat16((int32_t)yy + cr_140200);
        int16_t G = sat16((int32_t)yy - sat16((int32_t)cb_034414 + cr_071414));
        int16_t B = sat16((int32_t)yy + cb_177200);
        out[3 * p] = packus(R >> 6);
        out[3 * p + 1] = packus(G >> 6);
        out[3 * p + 2] = packus(B >> 6);
    }
    return nv * 8;
}

#if defined(__SSSE3__)
#include <tmmintrin.h>
int orc_ycbcr_line_ssse3_intrin(const uint8_t *y, const uint8_t *cb, const uint8_t *cr, uint8_t *out,
                                size_t npix, size_t *done) {
    size_t nv = npix / 8;
    nv = nv ? nv - 1 : 0;
    const __m128i shuf16 = _mm_setr_epi8(0, -0x7F, 1, -0x7F, 2, -0x7F, 3, -0x7F, 4, -0x7F, 5, -0x7F, 6, -0x7F, 7, -0x7F);
    const __m128i shufr = _mm_setr_epi8(0, -0x7F, -0x7F, 1, -0x7F, -0x7F, 2, -0x7F, -0x7F, 3, -0x7F, -0x7F, 4, -0x7F, -0x7F, 5);
    const __m128i shufg = _mm_setr_epi8(-0x7F, 0, -0x7F, -0x7F, 1, -0x7F, -0x7F, 2, -0x7F, -0x7F, 3, -0x7F, -0x7F, 4, -0x7F, -0x7F);
    const __m128i shufb = _mm_alignr_epi8(shufg, shufg, 15);
    const __m128i shufr1 = _mm_add_epi8(shufb, _mm_set1_epi8(6));
    const __m128i shufg1 = _mm_add_epi8(shufr, _mm_set1_epi8(5));
    const __m128i shufb1 = _mm_add_epi8(shufg, _mm_set1_epi8(5));
    for (size_t i = 0; i < nv; i++) {
        __m128i vy = _mm_loadl_epi64((const __m128i *)(y + i * 8));
        __m128i vb = _mm_loadl_epi64((const __m128i *)(cb + i * 8));
        __m128i vr = _mm_loadl_epi64((const __m128i *)(cr + i * 8));
        vy = _mm_slli_epi16(_mm_shuffle_epi8(vy, shuf16), 6);
        vb = _mm_slli_epi16(_mm_shuffle_epi8(vb, shuf16), 6);
        vr = _mm_slli_epi16(_mm_shuffle_epi8(vr, shuf16), 6);
        const __m128i c128 = _mm_set1_epi16(128 << 6);
        vy = _mm_adds_epi16(vy, _mm_set1_epi16(32));
        vb = _mm_subs_epi16(vb, c128);
        vr = _mm_subs_epi16(vr, c128);
        __m128i cr_140200 = _mm_adds_epi16(_mm_mulhrs_epi16(vr, _mm_set1_epi16(13173)), vr);
        __m128i cb_034414 = _mm_mulhrs_epi16(vb, _mm_set1_epi16(11276));
        __m128i cr_071414 = _mm_mulhrs_epi16(vr, _mm_set1_epi16(23401));
        __m128i cb_177200 = _mm_adds_epi16(_mm_mulhrs_epi16(vb, _mm_set1_epi16(25297)), vb);
        __m128i r = _mm_adds_epi16(vy, cr_140200);
        __m128i g = _mm_subs_epi16(vy, _mm_adds_epi16(cb_034414, cr_071414));
        __m128i b = _mm_adds_epi16(vy, cb_177200);
        const __m128i zero = _mm_setzero_si128();
        r = _mm_packus_epi16(_mm_srai_epi16(r, 6), zero);
        g = _mm_packus_epi16(_mm_srai_epi16(g, 6), zero);
        b = _mm_packus_epi16(_mm_srai_epi16(b, 6), zero);
        __m128i lo = _mm_or_si128(_mm_shuffle_epi8(r, shufr), _mm_or_si128(_mm_shuffle_epi8(g, shufg), _mm_shuffle_epi8(b, shufb)));
        __m128i hi = _mm_or_si128(_mm_shuffle_epi8(r, shufr1), _mm_or_si128(_mm_shuffle_epi8(g, shufg1), _mm_shuffle_epi8(b, shufb1)));
        uint8_t data[32];
        _mm_storeu_si128((__m128i *)data, lo);
        _mm_storeu_si128((__m128i *)(data + 16), hi);
        memcpy(out + 24 * i, data, 24);
    }
    *done = nv * 8;
    return 1;
}
#else
int orc_ycbcr_line_ssse3_intrin(const uint8_t *y, const uint8_t *cb, const uint8_t *cr, uint8_t *out,
                                size_t npix, size_t *done) {
    (void)y; (void)cb; (void)cr; (void)out; (void)npix; (void)done;
    return 0;
}
#endif

enum { CC_NOCONVERT, CC_RGB, CC_YCBCR, CC_CMYK, CC_YCCK };
/* src/decoder.rs:1339-1389 */
static int choose_color_convert(int ncomp, int ct, int *cc) {
    if (ncomp == 3) {
        switch (ct) {
        case ORC_CT_NONE: *cc = CC_NOCONVERT; return ORC_OK;
        case ORC_CT_GRAYSCALE: return fail(ORC_ERR_FORMAT, "Invalid number of channels (3) for Grayscale data");
        case ORC_CT_RGB: *cc = CC_RGB; return ORC_OK;
        case ORC_CT_YCBCR: *cc = CC_YCBCR; return ORC_OK;
        case ORC_CT_CMYK: return fail(ORC_ERR_FORMAT, "Invalid number of channels (3) for CMYK data");
        case ORC_CT_YCCK: return fail(ORC_ERR_FORMAT, "Invalid number of channels (3) for YCCK data");
        case ORC_CT_JCS_BG_YCC: return fail(ORC_ERR_UNSUPPORTED, "ColorTransform(JcsBgYcc)");
        case ORC_CT_JCS_BG_RGB: return fail(ORC_ERR_UNSUPPORTED, "ColorTransform(JcsBgRgb)");
        default: return fail(ORC_ERR_FORMAT, "Unknown colour transform");
        }
    } else if (ncomp == 4) {
        switch (ct) {
        case ORC_CT_NONE: *cc = CC_NOCONVERT; return ORC_OK;
        case ORC_CT_GRAYSCALE: return fail(ORC_ERR_FORMAT, "Invalid number of channels (4) for Grayscale data");
        case ORC_CT_RGB: return fail(ORC_ERR_FORMAT, "Invalid number of channels (4) for RGB data");
        case ORC_CT_YCBCR: return fail(ORC_ERR_FORMAT, "Invalid number of channels (4) for YCbCr data");
        case ORC_CT_CMYK: *cc = CC_CMYK; return ORC_OK;
        case ORC_CT_YCCK: *cc = CC_YCCK; return ORC_OK;
        case ORC_CT_JCS_BG_YCC: return fail(ORC_ERR_UNSUPPORTED, "ColorTransform(JcsBgYcc)");
        case ORC_CT_JCS_BG_RGB: return fail(ORC_ERR_UNSUPPORTED, "ColorTransform(JcsBgRgb)");
        default: return fail(ORC_ERR_FORMAT, "Unknown colour transform");
        }
    }
    return fail(ORC_ERR_INTERNAL, "panic: component count not 3 or 4");
}

/* line buffers hold `lbs` samples each; `out` is one row of W*ncomp bytes */
static int color_convert(int arith, int cc, int ncomp, uint8_t *const lb[4], size_t lbs, size_t w, uint8_t *out) {
    size_t n = w < lbs ? w : lbs; /* zip() stops at the shortest iterator */
    switch (cc) {
    case CC_RGB: /* src/decoder.rs:1391-1404 */
        for (size_t i = 0; i < n; i++) {
            out[3 * i] = lb[0][i];
            out[3 * i + 1] = lb[1][i];
            out[3 * i + 2] = lb[2][i];
        }
        return ORC_OK;
    case CC_YCBCR: { /* src/decoder.rs:1406-1437 */
        size_t done = 0;
        if (arith == ORC_ARITH_SSSE3_NATIVE && orc_ycbcr_line_ssse3_intrin(lb[0], lb[1], lb[2], out, w, &done)) {
        } else if (arith == ORC_ARITH_SSSE3 || arith == ORC_ARITH_SSSE3_NATIVE) {
            done = ycbcr_line_ssse3_emul(lb[0], lb[1], lb[2], out, w);
        }
        for (size_t i = done; i < n; i++) ycbcr_to_rgb(lb[0][i], lb[1][i], lb[2][i], out + 3 * i);
        return ORC_OK;
    }
    case CC_YCCK: /* src/decoder.rs:1439-1456 */
        for (size_t i = 0; i < n; i++) {
            ycbcr_to_rgb(lb[0][i], lb[1][i], lb[2][i], out + 4 * i);
            out[4 * i + 3] = (uint8_t)(255 - lb[3][i]);
        }
        return ORC_OK;
    case CC_CMYK: /* src/decoder.rs:1458-1474 */
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(255 - lb[k][i]);
        return ORC_OK;
    default: { /* color_no_convert, src/decoder.rs:1476-1484: planar dump, unwrap() panics on overrun */
        if (lbs * (size_t)ncomp > w * (size_t)ncomp)
            return fail(ORC_ERR_INTERNAL, "panic: color_no_convert line buffers longer than the output row");
        size_t k = 0;
        for (int c = 0; c < ncomp; c++)
            for (size_t i = 0; i < lbs; i++) out[k++] = lb[c][i];
        return ORC_OK;
    }
    }
}

typedef struct {
    upsampler up;
    int cc;
    int arith;
    int ncomp;
} assembler;

static int assembler_init(assembler *a, int arith, const orc_component *comps, int ncomp, uint16_t out_w,
                          uint16_t out_h, int ct) {
    int rc = choose_color_convert(ncomp, ct, &a->cc);
    if (rc) return rc;
    rc = upsampler_new(comps, ncomp, out_w, out_h, &a->up);
    if (rc) return rc;
    a->arith = arith;
    a->ncomp = ncomp;
    return ORC_OK;
}

/* src/upsampler.rs:47-63 with the per-row allocation hoisted into `scratch` (ncomp*lbs bytes, zeroed
 * per row like the reference's vec![0u8; ..]) */
static int assemble_row(const assembler *a, const uint8_t *const *planes, size_t row, size_t out_w,
                        uint8_t *scratch, uint8_t *out_row) {
    size_t lbs = a->up.line_buffer_size;
    uint8_t *lb[4] = {0, 0, 0, 0};
    memset(scratch, 0, lbs * (size_t)a->ncomp);
    for (int i = 0; i < a->ncomp; i++) {
        lb[i] = scratch + (size_t)i * lbs;
        upsample_row(&a->up.c[i], planes[i], row, out_w, lb[i]);
    }
    return color_convert(a->arith, a->cc, a->ncomp, lb, lbs, out_w, out_row);
}

int orc_upsample_and_interleave_row(int arith, const orc_component *comps, int ncomp,
                                    const uint8_t *const *planes, uint16_t out_w, uint16_t out_h,
                                    int color_transform, size_t row, uint8_t *out_row) {
    assembler a;
    int rc = assembler_init(&a, arith, comps, ncomp, out_w, out_h, color_transform);
    if (rc) return rc;
    uint8_t *scratch = (uint8_t *)malloc(a.up.line_buffer_size * (size_t)ncomp + 64);
    rc = assemble_row(&a, planes, row, out_w, scratch, out_row);
    free(scratch);
    return rc;
}

/* src/decoder.rs:1300-1336 + src/worker/mod.rs:97-128 */
int orc_compute_image(int arith, const orc_component *comps, int ncomp, const uint8_t *const *planes,
                      const size_t *plane_len, uint16_t out_w, uint16_t out_h, int color_transform,
                      uint8_t *out, size_t cap, size_t *out_len) {
    if (ncomp <= 0) return fail(ORC_ERR_FORMAT, "not all components have data");
    for (int i = 0; i < ncomp; i++)
        if (!planes[i] || plane_len[i] == 0) return fail(ORC_ERR_FORMAT, "not all components have data");
    if (ncomp == 1) {
        /* src/decoder.rs:1310-1332 */
        size_t width = comps[0].size_w, height = comps[0].size_h;
        size_t size = width * height;
        size_t line_stride = (size_t)comps[0].block_w * comps[0].dct_scale;
        if (cap < size) return fail(ORC_ERR_INTERNAL, "output buffer too small");
        if ((size_t)out_w != line_stride) {
            for (size_t y = 0; y < height; y++) memcpy(out + y * width, planes[0] + y * line_stride, width);
        } else {
            size_t n = size < plane_len[0] ? size : plane_len[0];
            memcpy(out, planes[0], n);
            if (n < size) memset(out + n, 0, size - n); /* decoded.resize(size, 0) */
        }
        if (out_len) *out_len = size;
        return ORC_OK;
    }
    assembler a;
    int rc = assembler_init(&a, arith, comps, ncomp, out_w, out_h, color_transform);
    if (rc) return rc;
    size_t line = (size_t)out_w * ncomp;
    if (cap < line * out_h) return fail(ORC_ERR_INTERNAL, "output buffer too small");
    uint8_t *scratch = (uint8_t *)malloc(a.up.line_buffer_size * (size_t)ncomp + 64);
    for (size_t row = 0; row < out_h && !rc; row++) rc = assemble_row(&a, planes, row, out_w, scratch, out + row * line);
    free(scratch);
    if (out_len) *out_len = line * out_h;
    return rc;
}

/* compute_image_parallel, src/worker/rayon.rs:193-219: the reference spawns one rayon task per output row; here
 * the rows are dealt to nthreads pthreads in interleaved order (same work, same per-row scratch). */
typedef struct {
    const assembler *a;
    const uint8_t *const *planes;
    size_t out_w, out_h, line, first, step;
    uint8_t *out;
    int rc;
} rows_job;

static void *rows_main(void *p) {
    rows_job *j = (rows_job *)p;
    uint8_t *scratch = (uint8_t *)malloc(j->a->up.line_buffer_size * (size_t)j->a->ncomp + 64);
    for (size_t row = j->first; row < j->out_h && !j->rc; row += j->step)
        j->rc = assemble_row(j->a, j->planes, row, j->out_w, scratch, j->out + row * j->line);
    free(scratch);
    return NULL;
}

int orc_compute_image_mt(int arith, int nthreads, const orc_component *comps, int ncomp,
                         const uint8_t *const *planes, const size_t *plane_len, uint16_t out_w, uint16_t out_h,
                         int color_transform, uint8_t *out, size_t cap, size_t *out_len) {
    if (nthreads <= 1 || ncomp <= 1)
        return orc_compute_image(arith, comps, ncomp, planes, plane_len, out_w, out_h, color_transform, out, cap, out_len);
    for (int i = 0; i < ncomp; i++)
        if (!planes[i] || plane_len[i] == 0) return fail(ORC_ERR_FORMAT, "not all components have data");
    assembler a;
    int rc = assembler_init(&a, arith, comps, ncomp, out_w, out_h, color_transform);
    if (rc) return rc;
    size_t line = (size_t)out_w * ncomp;
    if (cap < line * out_h) return fail(ORC_ERR_INTERNAL, "output buffer too small");
    if (nthreads > (int)out_h) nthreads = out_h;
    rows_job *jobs = (rows_job *)calloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
    for (int t = 0; t < nthreads; t++) {
        rows_job j = {&a, planes, out_w, out_h, line, (size_t)t, (size_t)nthreads, out, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, rows_main, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(jobs);
    free(th);
    if (out_len) *out_len = line * out_h;
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * Whole hot path for one image from dense coefficients: what decode_scan + decode_planes push
 * through the worker boundary for a baseline interleaved image (src/decoder.rs:848-861, 1058,
 * 1068-1078, 689-694).
 * ------------------------------------------------------------------------------------------ */
int orc_hotpath_image_mt(int arith, int nthreads, const orc_component *comps, int ncomp,
                         const uint16_t *const qts[4], const int16_t *const coefs[4], uint16_t out_w,
                         uint16_t out_h, int color_transform, uint8_t *out, size_t cap) {
    orc_worker *w = orc_worker_new(arith);
    uint8_t *planes[4] = {0, 0, 0, 0};
    size_t lens[4] = {0, 0, 0, 0};
    int rc = ORC_OK;
    for (int i = 0; i < ncomp && !rc; i++) rc = orc_worker_start(w, i, &comps[i], qts[i]);
    for (int i = 0; i < ncomp && !rc; i++) {
        size_t per_row = (size_t)comps[i].block_w * comps[i].v * 64;
        size_t rows = comps[i].block_h / comps[i].v; /* = mcu rows */
        for (size_t r = 0; r < rows && !rc; r++) rc = orc_worker_append_row(w, i, coefs[i] + r * per_row, per_row);
    }
    for (int i = 0; i < ncomp && !rc; i++) rc = orc_worker_get_result(w, i, &planes[i], &lens[i]);
    if (!rc)
        rc = orc_compute_image_mt(arith, nthreads, comps, ncomp, (const uint8_t *const *)planes, lens, out_w, out_h,
                                  color_transform, out, cap, NULL);
    for (int i = 0; i < 4; i++) free(planes[i]);
    orc_worker_free(w);
    return rc;
}

int orc_hotpath_image(int arith, const orc_component *comps, int ncomp, const uint16_t *const qts[4],
                      const int16_t *const coefs[4], uint16_t out_w, uint16_t out_h, int color_transform,
                      uint8_t *out, size_t cap) {
    return orc_hotpath_image_mt(arith, 1, comps, ncomp, qts, coefs, out_w, out_h, color_transform, out, cap);
}

typedef struct {
    int arith, ncomp, ct;
    size_t n;
    size_t *next;
    pthread_mutex_t *mu;
    const orc_component *comps;
    const uint16_t *const *qts;
    const int16_t *const *coefs;
    uint16_t out_w, out_h;
    uint8_t *const *outs;
    size_t cap;
    int rc;
} batch_job;

static void *batch_main(void *p) {
    batch_job *j = (batch_job *)p;
    for (;;) {
        pthread_mutex_lock(j->mu);
        size_t i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->n) break;
        const int16_t *c[4] = {0, 0, 0, 0};
        for (int k = 0; k < j->ncomp; k++) c[k] = j->coefs[i * (size_t)j->ncomp + k];
        int rc = orc_hotpath_image(j->arith, j->comps, j->ncomp, j->qts, c, j->out_w, j->out_h, j->ct,
                                   j->outs[i], j->cap);
        if (rc) j->rc = rc;
    }
    return NULL;
}

int orc_hotpath_batch(int arith, int nthreads, size_t n, const orc_component *comps, int ncomp,
                      const uint16_t *const qts[4], const int16_t *const *coefs, uint16_t out_w,
                      uint16_t out_h, int color_transform, uint8_t *const *outs, size_t cap) {
    if (nthreads < 1) nthreads = 1;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    size_t next = 0;
    batch_job *jobs = (batch_job *)calloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
    for (int t = 0; t < nthreads; t++) {
        batch_job j = {arith, ncomp, color_transform, n, &next, &mu, comps, qts, coefs, out_w, out_h, outs, cap, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, batch_main, &jobs[t]);
    }
    int rc = ORC_OK;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(jobs);
    free(th);
    return rc;
}
