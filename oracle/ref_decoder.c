/*
 * ref_decoder.c -- oracle (test infrastructure, see oracle.h): whole-file decoder, restating
 * /root/reference/src/decoder.rs:101-1298, src/parser.rs, src/huffman.rs, src/marker.rs.
 * It exists so that the oracle's hot path can be pinned against the reference's golden PNGs and so
 * that tests have coefficient buffers that really came out of JPEG files.  Lossless (SOF3) is not
 * restated: it bypasses the worker path entirely (src/decoder/lossless.rs) and is out of scope.
 */
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_COMPONENTS 4

/* src/decoder.rs:27-36 */
static const uint8_t UNZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* ---- markers (src/marker.rs) -------------------------------------------------------------- */
/* a marker is kept as its code byte; classification helpers below */
static int is_sof(uint8_t m) { return m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC; }
static int is_rst(uint8_t m) { return m >= 0xD0 && m <= 0xD7; }
static int is_app(uint8_t m) { return m >= 0xE0 && m <= 0xEF; }

/* ---- huffman tables (src/huffman.rs:181-285) ---------------------------------------------- */
typedef struct {
    int present;
    uint8_t values[256];
    int nvalues;
    int32_t delta[16];
    int32_t maxcode[16];
    uint8_t lut_value[256], lut_size[256];
    int has_ac_lut;
    int16_t ac_value[256];
    uint8_t ac_run_size[256];
} huff_table;

/* src/huffman.rs:165-173 */
static int16_t extend(uint16_t value, uint8_t count) {
    uint16_t vt = (uint16_t)(1u << (count - 1));
    if (value < vt) return (int16_t)((int32_t)value + (int32_t)((uint32_t)-1 << count) + 1);
    return (int16_t)value;
}

/* src/huffman.rs:191-285 ; returns 0 on "bad huffman code length" */
static int huff_table_new(huff_table *t, const uint8_t bits[16], const uint8_t *values, int nvalues, int is_ac) {
    uint8_t huffsize[256 + 16];
    uint16_t huffcode[256 + 16];
    int n = 0;
    for (int i = 0; i < 16; i++)
        for (int k = 0; k < bits[i]; k++) huffsize[n++] = (uint8_t)(i + 1);
    if (n == 0 || n > 256 || n != nvalues) return 0;
    uint8_t code_size = huffsize[0];
    uint32_t code = 0;
    for (int i = 0; i < n; i++) {
        while (code_size < huffsize[i]) {
            code <<= 1;
            code_size++;
        }
        if (code >= (1u << huffsize[i])) return 0;
        huffcode[i] = (uint16_t)code;
        code++;
    }
    memset(t, 0, sizeof *t);
    t->present = 1;
    memcpy(t->values, values, (size_t)nvalues);
    t->nvalues = nvalues;
    int j = 0;
    for (int i = 0; i < 16; i++) {
        t->maxcode[i] = -1;
        if (bits[i] != 0) {
            t->delta[i] = j - (int32_t)huffcode[j];
            j += bits[i];
            t->maxcode[i] = huffcode[j - 1];
        }
    }
    for (int i = 0; i < n; i++) {
        uint8_t size = huffsize[i];
        if (size > 8) continue;
        int rem = 8 - size;
        int start = huffcode[i] << rem;
        for (int b = 0; b < (1 << rem); b++) {
            t->lut_value[start + b] = values[i];
            t->lut_size[start + b] = size;
        }
    }
    if (is_ac) {
        t->has_ac_lut = 1;
        for (int i = 0; i < 256; i++) {
            uint8_t value = t->lut_value[i], size = t->lut_size[i];
            uint8_t run = value >> 4, mag = value & 0x0f;
            if (mag > 0 && size + mag <= 8) {
                uint16_t un = (uint16_t)((((unsigned)i << size) & 0xffu) >> (8 - mag));
                t->ac_value[i] = extend(un, mag);
                t->ac_run_size[i] = (uint8_t)((run << 4) | (size + mag));
            }
        }
    }
    return 1;
}

/* ---- decoder state ------------------------------------------------------------------------ */
typedef struct {
    uint8_t num_markers, seq_no;
    uint8_t *data;
    size_t len;
} icc_chunk;

typedef struct {
    int is_baseline, is_differential, coding_process, arithmetic;
    uint8_t precision;
    uint16_t image_w, image_h, output_w, output_h, mcu_w, mcu_h;
    int ncomp;
    orc_component comps[256];
} frame_info;

typedef struct {
    int n;
    int component_indices[4], dc_table_indices[4], ac_table_indices[4];
    uint8_t ss_start, ss_end; /* Range start..end (end exclusive) */
    uint8_t ah, al;
} scan_info;

struct orc_decoder {
    const uint8_t *data;
    size_t len, pos;
    int arith;
    int nthreads; /* compute_image: rows split over this many threads (rayon build), default 1 */
    int no_taps;  /* skip the test taps (copies of the planes) */

    int has_frame;
    frame_info frame;
    huff_table dc_tables[4], ac_tables[4];
    int has_qt[4];
    uint16_t qt[4][64]; /* natural order */
    uint16_t restart_interval;
    int has_adobe, adobe_transform; /* 0 Unknown 1 YCbCr 2 YCCK */
    int has_color_transform, color_transform;
    int is_jfif, is_mjpeg;
    icc_chunk *icc;
    size_t n_icc;
    uint8_t *exif, *xmp;
    size_t exif_len, xmp_len;
    int has_exif, has_xmp;
    int16_t *coefficients[MAX_COMPONENTS]; /* progressive store */
    size_t coefficients_len[MAX_COMPONENTS];
    int has_coefficients;
    uint64_t coefficients_finished[MAX_COMPONENTS];
    size_t buffer_limit;

    /* huffman bit reader (src/huffman.rs:14-18) */
    uint64_t bits;
    uint8_t num_bits;
    int has_marker;
    uint8_t marker;

    /* results */
    uint8_t *planes[MAX_COMPONENTS];
    size_t plane_len[MAX_COMPONENTS];
    uint8_t *pixels;
    size_t pixels_len;
    /* test taps */
    int16_t *tap_coefs[MAX_COMPONENTS];
    size_t tap_len[MAX_COMPONENTS], tap_cap[MAX_COMPONENTS];
    uint8_t *tap_plane[MAX_COMPONENTS];
    size_t tap_plane_len[MAX_COMPONENTS];
    uint8_t *icc_joined;
    size_t icc_joined_len;
    int final_ct;
    char err[256];
};

#define FAIL(d, code, ...)                                 \
    do {                                                   \
        snprintf((d)->err, sizeof(d)->err, __VA_ARGS__);   \
        return (code);                                     \
    } while (0)
#define TRY(x)            \
    do {                  \
        int rc_ = (x);    \
        if (rc_) return rc_; \
    } while (0)

/* ---- reader (src/lib.rs:56-66) ------------------------------------------------------------ */
static int read_u8(orc_decoder *d, uint8_t *b) {
    if (d->pos >= d->len) FAIL(d, ORC_ERR_IO, "failed to fill whole buffer");
    *b = d->data[d->pos++];
    return ORC_OK;
}
static int read_u16(orc_decoder *d, uint16_t *v) {
    if (d->pos + 2 > d->len) {
        d->pos = d->len;
        FAIL(d, ORC_ERR_IO, "failed to fill whole buffer");
    }
    *v = (uint16_t)((d->data[d->pos] << 8) | d->data[d->pos + 1]);
    d->pos += 2;
    return ORC_OK;
}
static int read_exact(orc_decoder *d, uint8_t *dst, size_t n) {
    if (d->pos + n > d->len) {
        d->pos = d->len;
        FAIL(d, ORC_ERR_IO, "failed to fill whole buffer");
    }
    memcpy(dst, d->data + d->pos, n);
    d->pos += n;
    return ORC_OK;
}
static int skip_bytes(orc_decoder *d, size_t n) {
    if (d->pos + n > d->len) {
        d->pos = d->len;
        FAIL(d, ORC_ERR_IO, "unexpected end of file");
    }
    d->pos += n;
    return ORC_OK;
}
/* src/parser.rs:136-147 */
static int read_length(orc_decoder *d, size_t *len) {
    uint16_t l;
    TRY(read_u16(d, &l));
    if (l < 2) FAIL(d, ORC_ERR_FORMAT, "encountered marker with invalid length %u", l);
    *len = (size_t)l - 2;
    return ORC_OK;
}

/* ---- huffman bit reader (src/huffman.rs:20-161) -------------------------------------------- */
static int read_bits(orc_decoder *d) {
    while (d->num_bits <= 56) {
        uint8_t byte = 0;
        if (!d->has_marker) TRY(read_u8(d, &byte));
        if (byte == 0xFF) {
            uint8_t next;
            TRY(read_u8(d, &next));
            if (next != 0x00) {
                while (next == 0xFF) TRY(read_u8(d, &next));
                if (next == 0x00) FAIL(d, ORC_ERR_FORMAT, "FF 00 found where marker was expected");
                d->has_marker = 1;
                d->marker = next;
                continue;
            }
        }
        d->bits |= (uint64_t)byte << (56 - d->num_bits);
        d->num_bits = (uint8_t)(d->num_bits + 8);
    }
    return ORC_OK;
}
static inline uint16_t peek_bits(orc_decoder *d, uint8_t count) {
    if (count == 0) return 0;
    return (uint16_t)((d->bits >> (64 - count)) & ((1u << count) - 1));
}
static inline void consume_bits(orc_decoder *d, uint8_t count) {
    d->bits = count >= 64 ? 0 : d->bits << count;
    d->num_bits = (uint8_t)(d->num_bits - count);
}
static int get_bits(orc_decoder *d, uint8_t count, uint16_t *v) {
    if (d->num_bits < count) TRY(read_bits(d));
    *v = peek_bits(d, count);
    consume_bits(d, count);
    return ORC_OK;
}
static int receive_extend(orc_decoder *d, uint8_t count, int16_t *v) {
    uint16_t u;
    TRY(get_bits(d, count, &u));
    *v = extend(u, count);
    return ORC_OK;
}
/* src/huffman.rs:31-58 */
static int huff_decode(orc_decoder *d, const huff_table *t, uint8_t *out) {
    if (d->num_bits < 16) TRY(read_bits(d));
    uint16_t idx = peek_bits(d, 8);
    if (t->lut_size[idx] > 0) {
        consume_bits(d, t->lut_size[idx]);
        *out = t->lut_value[idx];
        return ORC_OK;
    }
    uint16_t bits = peek_bits(d, 16);
    for (int i = 8; i < 16; i++) {
        int32_t code = bits >> (15 - i);
        if (code <= t->maxcode[i]) {
            consume_bits(d, (uint8_t)(i + 1));
            int32_t index = code + t->delta[i];
            if (index < 0 || index >= t->nvalues) FAIL(d, ORC_ERR_INTERNAL, "panic: huffman value index out of range");
            *out = t->values[index];
            return ORC_OK;
        }
    }
    FAIL(d, ORC_ERR_FORMAT, "failed to decode huffman code");
}
/* src/huffman.rs:60-78 ; *hit = 0 when the fast path does not apply */
static int decode_fast_ac(orc_decoder *d, const huff_table *t, int *hit, int16_t *value, uint8_t *run) {
    *hit = 0;
    if (t->has_ac_lut) {
        if (d->num_bits < 8) TRY(read_bits(d));
        uint16_t idx = peek_bits(d, 8);
        uint8_t rs = t->ac_run_size[idx];
        if (rs != 0) {
            *run = rs >> 4;
            consume_bits(d, rs & 0x0f);
            *value = t->ac_value[idx];
            *hit = 1;
        }
    }
    return ORC_OK;
}
/* src/huffman.rs:103-105 */
static int take_marker(orc_decoder *d, int *has, uint8_t *m) {
    TRY(read_bits(d));
    *has = d->has_marker;
    *m = d->marker;
    d->has_marker = 0;
    return ORC_OK;
}

/* ---- Annex K tables for MJPEG (src/huffman.rs:295-346) -------------------------------------- */
static const uint8_t K3_BITS[16] = {0x00, 0x01, 0x05, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00};
static const uint8_t K4_BITS[16] = {0x00, 0x03, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x00, 0x00, 0x00, 0x00, 0x00};
static const uint8_t K34_VALS[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t K5_BITS[16] = {0x00, 0x02, 0x01, 0x03, 0x03, 0x02, 0x04, 0x03, 0x05, 0x05, 0x04, 0x04, 0x00, 0x00, 0x01, 0x7D};
static const uint8_t K5_VALS[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3,
    0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3,
    0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2,
    0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};
static const uint8_t K6_BITS[16] = {0x00, 0x02, 0x01, 0x02, 0x04, 0x04, 0x03, 0x04, 0x07, 0x05, 0x04, 0x04, 0x00, 0x01, 0x02, 0x77};
static const uint8_t K6_VALS[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1,
    0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A,
    0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA,
    0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA,
    0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};

static void fill_default_mjpeg_tables(orc_decoder *d, const scan_info *scan) {
    int use_dc0 = 0, use_dc1 = 0, use_ac0 = 0, use_ac1 = 0;
    for (int i = 0; i < scan->n; i++) {
        use_dc0 |= scan->dc_table_indices[i] == 0;
        use_dc1 |= scan->dc_table_indices[i] == 1;
        use_ac0 |= scan->ac_table_indices[i] == 0;
        use_ac1 |= scan->ac_table_indices[i] == 1;
    }
    if (!d->dc_tables[0].present && use_dc0) huff_table_new(&d->dc_tables[0], K3_BITS, K34_VALS, 12, 0);
    if (!d->dc_tables[1].present && use_dc1) huff_table_new(&d->dc_tables[1], K4_BITS, K34_VALS, 12, 0);
    if (!d->ac_tables[0].present && use_ac0) huff_table_new(&d->ac_tables[0], K5_BITS, K5_VALS, 162, 1);
    if (!d->ac_tables[1].present && use_ac1) huff_table_new(&d->ac_tables[1], K6_BITS, K6_VALS, 162, 1);
}

/* ---- segment parsers (src/parser.rs) ------------------------------------------------------- */
/* validation-only use of the upsampler chooser (src/decoder.rs:375-379, src/upsampler.rs:76-105) */
static int validate_sampling(orc_decoder *d, const frame_info *f) {
    uint8_t hmax = 0, vmax = 0;
    for (int i = 0; i < f->ncomp; i++) {
        if (f->comps[i].h > hmax) hmax = f->comps[i].h;
        if (f->comps[i].v > vmax) vmax = f->comps[i].v;
    }
    for (int i = 0; i < f->ncomp; i++) {
        uint8_t h = f->comps[i].h, v = f->comps[i].v;
        int h1 = h == hmax || f->image_w == 1, v1 = v == vmax || f->image_h == 1;
        int h2 = h * 2 == hmax, v2 = v * 2 == vmax;
        if ((h1 && v1) || (h2 && v1) || (h1 && v2) || (h2 && v2)) continue;
        if (hmax % h != 0 || vmax % v != 0) FAIL(d, ORC_ERR_UNSUPPORTED, "NonIntegerSubsamplingRatio");
    }
    return ORC_OK;
}

/* src/parser.rs:161-280 */
static int parse_sof(orc_decoder *d, uint8_t marker, frame_info *f) {
    size_t length;
    TRY(read_length(d, &length));
    if (length <= 6) FAIL(d, ORC_ERR_FORMAT, "invalid length in SOF");
    int n = marker - 0xC0;
    memset(f, 0, sizeof *f);
    f->is_baseline = n == 0;
    f->is_differential = (n >= 5 && n <= 7) || (n >= 13 && n <= 15);
    f->coding_process = (n == 0 || n == 1 || n == 5 || n == 9 || n == 13) ? 0 : ((n == 2 || n == 6 || n == 10 || n == 14) ? 1 : 2);
    f->arithmetic = n >= 9;
    TRY(read_u8(d, &f->precision));
    if (f->precision == 8) {
    } else if (f->precision == 12) {
        if (f->is_baseline) FAIL(d, ORC_ERR_FORMAT, "12 bit sample precision is not allowed in baseline");
    } else if (f->coding_process != 2 || f->precision > 16) {
        FAIL(d, ORC_ERR_FORMAT, "invalid precision %u in frame header", f->precision);
    }
    uint16_t height, width;
    TRY(read_u16(d, &height));
    TRY(read_u16(d, &width));
    if (height == 0) FAIL(d, ORC_ERR_UNSUPPORTED, "DNL");
    if (width == 0) FAIL(d, ORC_ERR_FORMAT, "zero width in frame header");
    uint8_t count;
    TRY(read_u8(d, &count));
    if (count == 0) FAIL(d, ORC_ERR_FORMAT, "zero component count in frame header");
    if (f->coding_process == 1 && count > 4) FAIL(d, ORC_ERR_FORMAT, "progressive frame with more than 4 components");
    if (length != 6 + 3 * (size_t)count) FAIL(d, ORC_ERR_FORMAT, "invalid length in SOF");
    for (int i = 0; i < count; i++) {
        uint8_t id, byte, tq;
        TRY(read_u8(d, &id));
        for (int k = 0; k < i; k++)
            if (f->comps[k].identifier == id) FAIL(d, ORC_ERR_FORMAT, "duplicate frame component identifier %u", id);
        TRY(read_u8(d, &byte));
        uint8_t h = byte >> 4, v = byte & 0x0f;
        if (h == 0 || h > 4) FAIL(d, ORC_ERR_FORMAT, "invalid horizontal sampling factor %u", h);
        if (v == 0 || v > 4) FAIL(d, ORC_ERR_FORMAT, "invalid vertical sampling factor %u", v);
        TRY(read_u8(d, &tq));
        if (tq > 3 || (f->coding_process == 2 && tq != 0)) FAIL(d, ORC_ERR_FORMAT, "invalid quantization table index %u", tq);
        orc_component *c = &f->comps[i];
        c->identifier = id;
        c->h = h;
        c->v = v;
        c->tq = tq;
        c->dct_scale = 8;
    }
    f->ncomp = count;
    if (orc_update_component_sizes(width, height, f->comps, count, &f->mcu_w, &f->mcu_h)) FAIL(d, ORC_ERR_FORMAT, "invalid dimensions");
    f->image_w = f->output_w = width;
    f->image_h = f->output_h = height;
    return ORC_OK;
}

/* src/parser.rs:332-482 */
static int parse_sos(orc_decoder *d, const frame_info *f, scan_info *s) {
    size_t length;
    TRY(read_length(d, &length));
    if (length == 0) FAIL(d, ORC_ERR_FORMAT, "zero length in SOS");
    uint8_t count;
    TRY(read_u8(d, &count));
    if (count == 0 || count > 4) FAIL(d, ORC_ERR_FORMAT, "invalid component count %u in scan header", count);
    if (length != 4 + 2 * (size_t)count) FAIL(d, ORC_ERR_FORMAT, "invalid length in SOS");
    memset(s, 0, sizeof *s);
    int maxidx = 0;
    for (int i = 0; i < count; i++) {
        uint8_t id, byte;
        TRY(read_u8(d, &id));
        int ci = -1;
        for (int k = 0; k < f->ncomp; k++)
            if (f->comps[k].identifier == id) {
                ci = k;
                break;
            }
        if (ci < 0) FAIL(d, ORC_ERR_FORMAT, "scan component identifier %u does not match any of the component identifiers defined in the frame", id);
        for (int k = 0; k < i; k++)
            if (s->component_indices[k] == ci) FAIL(d, ORC_ERR_FORMAT, "duplicate scan component identifier %u", id);
        if (ci < maxidx) FAIL(d, ORC_ERR_FORMAT, "the scan component order does not follow the order in the frame header");
        TRY(read_u8(d, &byte));
        uint8_t dc = byte >> 4, ac = byte & 0x0f;
        if (dc > 3 || (f->is_baseline && dc > 1)) FAIL(d, ORC_ERR_FORMAT, "invalid dc table index %u", dc);
        if (ac > 3 || (f->is_baseline && ac > 1)) FAIL(d, ORC_ERR_FORMAT, "invalid ac table index %u", ac);
        s->component_indices[i] = ci;
        s->dc_table_indices[i] = dc;
        s->ac_table_indices[i] = ac;
        if (ci > maxidx) maxidx = ci;
    }
    s->n = count;
    uint32_t blocks_per_mcu = 0;
    for (int i = 0; i < count; i++) blocks_per_mcu += (uint32_t)f->comps[s->component_indices[i]].h * f->comps[s->component_indices[i]].v;
    if (count > 1 && blocks_per_mcu > 10) FAIL(d, ORC_ERR_FORMAT, "scan with more than one component and more than 10 blocks per MCU");
    uint8_t ss = 0, se = 0, byte = 0;
    TRY(read_u8(d, &ss));
    TRY(read_u8(d, &se));
    TRY(read_u8(d, &byte));
    uint8_t ah = byte >> 4, al = byte & 0x0f;
    if (al >= f->precision) FAIL(d, ORC_ERR_FORMAT, "invalid point transform, must be less than the frame precision");
    if (f->coding_process == 1) {
        if (se > 63 || ss > se || (ss == 0 && se != 0)) FAIL(d, ORC_ERR_FORMAT, "invalid spectral selection parameters: ss=%u, se=%u", ss, se);
        if (ss != 0 && count != 1) FAIL(d, ORC_ERR_FORMAT, "spectral selection scan with AC coefficients can't have more than one component");
        if (ah > 13 || al > 13) FAIL(d, ORC_ERR_FORMAT, "invalid successive approximation parameters: ah=%u, al=%u", ah, al);
        if (ah != 0 && ah != al + 1) FAIL(d, ORC_ERR_FORMAT, "successive approximation scan with more than one bit of improvement");
    } else if (f->coding_process == 2) {
        if (se != 0) FAIL(d, ORC_ERR_FORMAT, "spectral selection end shall be zero in lossless scan");
        if (ah != 0) FAIL(d, ORC_ERR_FORMAT, "successive approximation high shall be zero in lossless scan");
        if (ss > 7) FAIL(d, ORC_ERR_FORMAT, "invalid predictor selection value: %u", ss);
    } else {
        if (se == 0) se = 63;
        if (ss != 0 || se != 63) FAIL(d, ORC_ERR_FORMAT, "spectral selection is not allowed in non-progressive scan");
        if (ah != 0 || al != 0) FAIL(d, ORC_ERR_FORMAT, "successive approximation is not allowed in non-progressive scan");
    }
    s->ss_start = ss;
    s->ss_end = (uint8_t)(se + 1);
    s->ah = ah;
    s->al = al;
    return ORC_OK;
}

/* src/parser.rs:485-532 + de-zigzag of src/decoder.rs:485-498 */
static int parse_dqt(orc_decoder *d) {
    size_t length;
    TRY(read_length(d, &length));
    uint16_t tables[4][64];
    int got[4] = {0, 0, 0, 0};
    while (length > 0) {
        uint8_t byte;
        TRY(read_u8(d, &byte));
        size_t precision = byte >> 4, index = byte & 0x0f;
        if (precision > 1) FAIL(d, ORC_ERR_FORMAT, "invalid precision %zu in DQT", precision);
        if (index > 3) FAIL(d, ORC_ERR_FORMAT, "invalid destination identifier %zu in DQT", index);
        if (length < 65 + 64 * precision) FAIL(d, ORC_ERR_FORMAT, "invalid length in DQT");
        for (int i = 0; i < 64; i++) {
            if (precision == 0) {
                uint8_t b;
                TRY(read_u8(d, &b));
                tables[index][i] = b;
            } else {
                TRY(read_u16(d, &tables[index][i]));
            }
        }
        for (int i = 0; i < 64; i++)
            if (tables[index][i] == 0) FAIL(d, ORC_ERR_FORMAT, "quantization table contains element with a zero value");
        got[index] = 1;
        length -= 65 + 64 * precision;
    }
    for (int t = 0; t < 4; t++)
        if (got[t]) {
            for (int j = 0; j < 64; j++) d->qt[t][UNZIGZAG[j]] = tables[t][j];
            d->has_qt[t] = 1;
        }
    return ORC_OK;
}

/* src/parser.rs:536-589 + merge of src/decoder.rs:501-518 */
static int parse_dht(orc_decoder *d) {
    size_t length;
    TRY(read_length(d, &length));
    huff_table *ndc = (huff_table *)calloc(4, sizeof(huff_table));
    huff_table *nac = (huff_table *)calloc(4, sizeof(huff_table));
    int rc = ORC_OK;
#define DHT_FAIL(code, msg)                        \
    do {                                           \
        snprintf(d->err, sizeof d->err, "%s", msg); \
        rc = (code);                               \
        goto done;                                 \
    } while (0)
    while (length > 17) {
        uint8_t byte;
        if ((rc = read_u8(d, &byte))) goto done;
        uint8_t cls = byte >> 4;
        size_t index = byte & 0x0f;
        if (cls != 0 && cls != 1) DHT_FAIL(ORC_ERR_FORMAT, "invalid class in DHT");
        if (d->has_frame && d->frame.is_baseline && index > 1) DHT_FAIL(ORC_ERR_FORMAT, "a maximum of two huffman tables per class are allowed in baseline");
        if (index > 3) DHT_FAIL(ORC_ERR_FORMAT, "invalid destination identifier in DHT");
        uint8_t counts[16];
        if ((rc = read_exact(d, counts, 16))) goto done;
        size_t size = 0;
        for (int i = 0; i < 16; i++) size += counts[i];
        if (size == 0) DHT_FAIL(ORC_ERR_FORMAT, "encountered table with zero length in DHT");
        if (size > 256) DHT_FAIL(ORC_ERR_FORMAT, "encountered table with excessive length in DHT");
        if (size > length - 17) DHT_FAIL(ORC_ERR_FORMAT, "invalid length in DHT");
        uint8_t values[256];
        if ((rc = read_exact(d, values, size))) goto done;
        huff_table *dst = cls == 0 ? &ndc[index] : &nac[index];
        if (!huff_table_new(dst, counts, values, (int)size, cls == 1)) DHT_FAIL(ORC_ERR_FORMAT, "bad huffman code length");
        length -= 17 + size;
    }
    if (length != 0) DHT_FAIL(ORC_ERR_FORMAT, "invalid length in DHT");
    for (int i = 0; i < 4; i++) {
        if (ndc[i].present) d->dc_tables[i] = ndc[i];
        if (nac[i].present) d->ac_tables[i] = nac[i];
    }
done:
    free(ndc);
    free(nac);
    return rc;
#undef DHT_FAIL
}

/* src/parser.rs:613-710 + dispatch of src/decoder.rs:532-558 */
static int parse_app(orc_decoder *d, uint8_t marker) {
    size_t length, bytes_read = 0;
    TRY(read_length(d, &length));
    int n = marker - 0xE0;
    if (n == 0) {
        if (length >= 5) {
            uint8_t b[5];
            TRY(read_exact(d, b, 5));
            bytes_read = 5;
            if (!memcmp(b, "JFIF\0", 5)) d->is_jfif = 1;
            else if (!memcmp(b, "AVI1\0", 5)) d->is_mjpeg = 1;
        }
    } else if (n == 1) {
        uint8_t *buf = (uint8_t *)malloc(length ? length : 1);
        int rc = read_exact(d, buf, length);
        if (rc) {
            free(buf);
            return rc;
        }
        bytes_read = length;
        if (length >= 6 && !memcmp(buf, "Exif\0\0", 6)) {
            free(d->exif);
            d->exif = (uint8_t *)malloc(length - 6 ? length - 6 : 1);
            memcpy(d->exif, buf + 6, length - 6);
            d->exif_len = length - 6;
            d->has_exif = 1;
        } else if (length >= 29 && !memcmp(buf, "http://ns.adobe.com/xap/1.0/\0", 29)) {
            free(d->xmp);
            d->xmp = (uint8_t *)malloc(length - 29 ? length - 29 : 1);
            memcpy(d->xmp, buf + 29, length - 29);
            d->xmp_len = length - 29;
            d->has_xmp = 1;
        }
        free(buf);
    } else if (n == 2) {
        if (length > 14) {
            uint8_t b[14];
            TRY(read_exact(d, b, 14));
            bytes_read = 14;
            if (!memcmp(b, "ICC_PROFILE\0", 12)) {
                size_t dl = length - 14;
                uint8_t *data = (uint8_t *)malloc(dl ? dl : 1);
                int rc = read_exact(d, data, dl);
                if (rc) {
                    free(data);
                    return rc;
                }
                bytes_read += dl;
                d->icc = (icc_chunk *)realloc(d->icc, (d->n_icc + 1) * sizeof(icc_chunk));
                d->icc[d->n_icc].seq_no = b[12];
                d->icc[d->n_icc].num_markers = b[13];
                d->icc[d->n_icc].data = data;
                d->icc[d->n_icc].len = dl;
                d->n_icc++;
            }
        }
    } else if (n == 13) {
        if (length >= 14) {
            uint8_t b[14];
            TRY(read_exact(d, b, 14));
            bytes_read = 14;
            if (!memcmp(b, "Photoshop 3.0\0", 14)) {
                TRY(skip_bytes(d, length - 14)); /* PSIR payload is read and kept by the reference; unused */
                bytes_read = length;
            }
        }
    } else if (n == 14) {
        if (length >= 12) {
            uint8_t b[12];
            TRY(read_exact(d, b, 12));
            bytes_read = 12;
            if (!memcmp(b, "Adobe\0", 6)) {
                if (b[11] > 2) FAIL(d, ORC_ERR_FORMAT, "invalid color transform in adobe app segment");
                d->has_adobe = 1;
                d->adobe_transform = b[11];
            }
        }
    }
    return skip_bytes(d, length - bytes_read);
}

/* src/decoder.rs:766-791 */
static int read_marker(orc_decoder *d, uint8_t *m) {
    for (;;) {
        uint8_t b;
        do {
            TRY(read_u8(d, &b));
        } while (b != 0xFF);
        TRY(read_u8(d, &b));
        while (b == 0xFF) TRY(read_u8(d, &b));
        if (b != 0x00 && b != 0xFF) {
            *m = b;
            return ORC_OK;
        }
    }
}

/* ---- block decoding (src/decoder.rs:1086-1298) ---------------------------------------------- */
static int decode_block(orc_decoder *d, int16_t *coefficients, const huff_table *dc_table, const huff_table *ac_table,
                        uint8_t ss_start, uint8_t ss_end, uint8_t al, uint16_t *eob_run, int16_t *dc_predictor) {
    if (ss_start == 0) {
        uint8_t value;
        TRY(huff_decode(d, dc_table, &value));
        int16_t diff = 0;
        if (value == 0) diff = 0;
        else if (value <= 11) TRY(receive_extend(d, value, &diff));
        else FAIL(d, ORC_ERR_FORMAT, "invalid DC difference magnitude category");
        *dc_predictor = (int16_t)((uint16_t)*dc_predictor + (uint16_t)diff); /* wrapping_add */
        coefficients[0] = (int16_t)((uint16_t)*dc_predictor << al);
    }
    uint8_t index = ss_start > 1 ? ss_start : 1;
    if (index < ss_end && *eob_run > 0) {
        *eob_run -= 1;
        return ORC_OK;
    }
    while (index < ss_end) {
        int hit;
        int16_t value;
        uint8_t run;
        TRY(decode_fast_ac(d, ac_table, &hit, &value, &run));
        if (hit) {
            index = (uint8_t)(index + run);
            if (index >= ss_end) break;
            coefficients[UNZIGZAG[index]] = (int16_t)((uint16_t)value << al);
            index++;
        } else {
            uint8_t byte;
            TRY(huff_decode(d, ac_table, &byte));
            uint8_t r = byte >> 4, s = byte & 0x0f;
            if (s == 0) {
                if (r == 15) {
                    index = (uint8_t)(index + 16);
                } else {
                    *eob_run = (uint16_t)((1u << r) - 1);
                    if (r > 0) {
                        uint16_t extra;
                        TRY(get_bits(d, r, &extra));
                        *eob_run = (uint16_t)(*eob_run + extra);
                    }
                    break;
                }
            } else {
                index = (uint8_t)(index + r);
                if (index >= ss_end) break;
                int16_t v;
                TRY(receive_extend(d, s, &v));
                coefficients[UNZIGZAG[index]] = (int16_t)((uint16_t)v << al);
                index++;
            }
        }
    }
    return ORC_OK;
}

/* src/decoder.rs:1260-1298 */
static int refine_non_zeroes(orc_decoder *d, int16_t *coefficients, uint8_t start, uint8_t end, uint8_t zrl, int16_t bit,
                             uint8_t *ret) {
    uint8_t last = (uint8_t)(end - 1);
    uint8_t zero_run_length = zrl;
    for (uint8_t i = start; i < end; i++) {
        int16_t *c = &coefficients[UNZIGZAG[i]];
        if (*c == 0) {
            if (zero_run_length == 0) {
                *ret = i;
                return ORC_OK;
            }
            zero_run_length--;
        } else {
            uint16_t b;
            TRY(get_bits(d, 1, &b));
            if (b == 1 && (*c & bit) == 0) {
                int32_t v = *c > 0 ? (int32_t)*c + bit : (int32_t)*c - bit;
                if (v > 32767 || v < -32768) FAIL(d, ORC_ERR_FORMAT, "Coefficient overflow");
                *c = (int16_t)v;
            }
        }
    }
    *ret = last;
    return ORC_OK;
}

/* src/decoder.rs:1174-1258 */
static int decode_block_sa(orc_decoder *d, int16_t *coefficients, const huff_table *ac_table, uint8_t ss_start, uint8_t ss_end,
                           uint8_t al, uint16_t *eob_run) {
    int16_t bit = (int16_t)(1 << al);
    if (ss_start == 0) {
        uint16_t b;
        TRY(get_bits(d, 1, &b));
        if (b == 1) coefficients[0] |= bit;
        return ORC_OK;
    }
    if (*eob_run > 0) {
        *eob_run -= 1;
        uint8_t r;
        return refine_non_zeroes(d, coefficients, ss_start, ss_end, 64, bit, &r);
    }
    uint8_t index = ss_start;
    while (index < ss_end) {
        uint8_t byte;
        TRY(huff_decode(d, ac_table, &byte));
        uint8_t r = byte >> 4, s = byte & 0x0f;
        uint8_t zero_run_length = r;
        int16_t value = 0;
        if (s == 0) {
            if (r != 15) {
                *eob_run = (uint16_t)((1u << r) - 1);
                if (r > 0) {
                    uint16_t extra;
                    TRY(get_bits(d, r, &extra));
                    *eob_run = (uint16_t)(*eob_run + extra);
                }
                zero_run_length = 64;
            }
        } else if (s == 1) {
            uint16_t b;
            TRY(get_bits(d, 1, &b));
            value = b == 1 ? bit : (int16_t)-bit;
        } else {
            FAIL(d, ORC_ERR_FORMAT, "unexpected huffman code");
        }
        TRY(refine_non_zeroes(d, coefficients, index, ss_end, zero_run_length, bit, &index));
        if (value != 0) coefficients[UNZIGZAG[index]] = value;
        index++;
    }
    return ORC_OK;
}

/* ---- test tap ------------------------------------------------------------------------------ */
static void tap_append(orc_decoder *d, int frame_comp, const int16_t *row, size_t n) {
    if (d->no_taps) return;
    size_t need = d->tap_len[frame_comp] + n;
    if (need > d->tap_cap[frame_comp]) {
        size_t cap = d->tap_cap[frame_comp] ? d->tap_cap[frame_comp] * 2 : 4096;
        while (cap < need) cap *= 2;
        d->tap_coefs[frame_comp] = (int16_t *)realloc(d->tap_coefs[frame_comp], cap * sizeof(int16_t));
        d->tap_cap[frame_comp] = cap;
    }
    memcpy(d->tap_coefs[frame_comp] + d->tap_len[frame_comp], row, n * sizeof(int16_t));
    d->tap_len[frame_comp] = need;
}

/* ---- decode_scan (src/decoder.rs:794-1082) -------------------------------------------------- */
static int decode_scan(orc_decoder *d, const frame_info *frame, const scan_info *scan, orc_worker *worker,
                       const int finished[MAX_COMPONENTS], int *has_marker, uint8_t *out_marker,
                       uint8_t *data[MAX_COMPONENTS], size_t data_len[MAX_COMPONENTS], int *has_data) {
    orc_component components[MAX_COMPONENTS];
    int nc = scan->n;
    for (int i = 0; i < nc; i++) components[i] = frame->comps[scan->component_indices[i]];
    for (int i = 0; i < nc; i++)
        if (!d->has_qt[components[i].tq]) FAIL(d, ORC_ERR_FORMAT, "use of unset quantization table");
    if (d->is_mjpeg) fill_default_mjpeg_tables(d, scan);
    if (scan->ss_start == 0)
        for (int i = 0; i < nc; i++)
            if (!d->dc_tables[scan->dc_table_indices[i]].present) FAIL(d, ORC_ERR_FORMAT, "scan makes use of unset dc huffman table");
    if (scan->ss_end > 1)
        for (int i = 0; i < nc; i++)
            if (!d->ac_tables[scan->ac_table_indices[i]].present) FAIL(d, ORC_ERR_FORMAT, "scan makes use of unset ac huffman table");

    for (int i = 0; i < nc; i++)
        if (finished[i]) {
            if (orc_worker_start(worker, i, &components[i], d->qt[components[i].tq])) FAIL(d, ORC_ERR_INTERNAL, "%s", orc_last_error());
            d->tap_len[scan->component_indices[i]] = 0;
        }

    int is_progressive = frame->coding_process == 1;
    int is_interleaved = nc > 1;
    int16_t dummy_block[64];
    memset(dummy_block, 0, sizeof dummy_block);
    d->bits = 0;
    d->num_bits = 0;
    d->has_marker = 0; /* HuffmanDecoder::new() */
    int16_t dc_predictors[MAX_COMPONENTS] = {0, 0, 0, 0};
    uint16_t mcus_left_until_restart = d->restart_interval;
    uint8_t expected_rst_num = 0;
    uint16_t eob_run = 0;
    int16_t *mcu_row_coefficients[MAX_COMPONENTS] = {0, 0, 0, 0};
    size_t per_row[MAX_COMPONENTS] = {0, 0, 0, 0};
    int rc = ORC_OK;
    for (int i = 0; i < nc; i++) per_row[i] = (size_t)components[i].block_w * components[i].v * 64;
    if (!is_progressive)
        for (int i = 0; i < nc; i++)
            if (finished[i]) mcu_row_coefficients[i] = (int16_t *)calloc(per_row[i], sizeof(int16_t));

    uint16_t mh[MAX_COMPONENTS] = {1, 1, 1, 1}, mv[MAX_COMPONENTS] = {1, 1, 1, 1};
    uint16_t max_mcu_x, max_mcu_y;
    if (is_interleaved) {
        for (int i = 0; i < nc; i++) {
            mh[i] = components[i].h;
            mv[i] = components[i].v;
        }
        max_mcu_x = frame->mcu_w;
        max_mcu_y = frame->mcu_h;
    } else {
        max_mcu_x = components[0].block_w;
        max_mcu_y = components[0].block_h;
    }
#define SCAN_FAIL(code, ...)                          \
    do {                                              \
        snprintf(d->err, sizeof d->err, __VA_ARGS__); \
        rc = (code);                                  \
        goto done;                                    \
    } while (0)
#define SCAN_TRY(x)          \
    do {                     \
        rc = (x);            \
        if (rc) goto done;   \
    } while (0)

    for (uint32_t mcu_y = 0; mcu_y < max_mcu_y; mcu_y++) {
        if (mcu_y * 8 >= frame->image_h) break;
        for (uint32_t mcu_x = 0; mcu_x < max_mcu_x; mcu_x++) {
            if (mcu_x * 8 >= frame->image_w) break;
            if (d->restart_interval > 0) {
                if (mcus_left_until_restart == 0) {
                    int has;
                    uint8_t m;
                    SCAN_TRY(take_marker(d, &has, &m));
                    if (has && is_rst(m)) {
                        uint8_t n = (uint8_t)(m - 0xD0);
                        if (n != expected_rst_num) SCAN_FAIL(ORC_ERR_FORMAT, "found RST%u where RST%u was expected", n, expected_rst_num);
                        d->bits = 0;
                        d->num_bits = 0; /* huffman.reset() */
                        memset(dc_predictors, 0, sizeof dc_predictors);
                        eob_run = 0;
                        expected_rst_num = (uint8_t)((expected_rst_num + 1) % 8);
                        mcus_left_until_restart = d->restart_interval;
                    } else if (has) {
                        SCAN_FAIL(ORC_ERR_FORMAT, "found marker 0x%02X inside scan where RST%u was expected", m, expected_rst_num);
                    } else {
                        SCAN_FAIL(ORC_ERR_FORMAT, "no marker found where RST%u was expected", expected_rst_num);
                    }
                }
                mcus_left_until_restart--;
            }
            for (int i = 0; i < nc; i++) {
                const orc_component *component = &components[i];
                for (uint32_t v_pos = 0; v_pos < mv[i]; v_pos++)
                    for (uint32_t h_pos = 0; h_pos < mh[i]; h_pos++) {
                        int16_t *coefficients;
                        if (is_progressive) {
                            size_t block_y = (size_t)mcu_y * mv[i] + v_pos;
                            size_t block_x = (size_t)mcu_x * mh[i] + h_pos;
                            size_t off = (block_y * component->block_w + block_x) * 64;
                            int ci = scan->component_indices[i];
                            if (off + 64 > d->coefficients_len[ci]) SCAN_FAIL(ORC_ERR_INTERNAL, "panic: coefficient slice out of range");
                            coefficients = d->coefficients[ci] + off;
                        } else if (finished[i]) {
                            uint32_t cur = is_interleaved ? 0 : (mcu_y % component->v);
                            size_t block_y = (size_t)cur * mv[i] + v_pos;
                            size_t block_x = (size_t)mcu_x * mh[i] + h_pos;
                            size_t off = (block_y * component->block_w + block_x) * 64;
                            if (off + 64 > per_row[i]) SCAN_FAIL(ORC_ERR_INTERNAL, "panic: coefficient slice out of range");
                            coefficients = mcu_row_coefficients[i] + off;
                        } else {
                            coefficients = dummy_block;
                        }
                        if (scan->ah == 0)
                            SCAN_TRY(decode_block(d, coefficients, &d->dc_tables[scan->dc_table_indices[i]],
                                                  &d->ac_tables[scan->ac_table_indices[i]], scan->ss_start, scan->ss_end, scan->al,
                                                  &eob_run, &dc_predictors[i]));
                        else
                            SCAN_TRY(decode_block_sa(d, coefficients, &d->ac_tables[scan->ac_table_indices[i]], scan->ss_start,
                                                     scan->ss_end, scan->al, &eob_run));
                    }
            }
        }
        for (int i = 0; i < nc; i++) {
            const orc_component *component = &components[i];
            if (!finished[i]) continue;
            if (!is_interleaved && (mcu_y + 1) * 8 < frame->image_h && (mcu_y + 1) % component->v > 0) continue;
            const int16_t *row;
            if (is_progressive) {
                uint32_t worker_mcu_y = is_interleaved ? mcu_y : mcu_y / component->v;
                size_t off = (size_t)worker_mcu_y * per_row[i];
                int ci = scan->component_indices[i];
                if (off + per_row[i] > d->coefficients_len[ci]) SCAN_FAIL(ORC_ERR_INTERNAL, "panic: coefficient slice out of range");
                row = d->coefficients[ci] + off;
            } else {
                row = mcu_row_coefficients[i];
            }
            tap_append(d, scan->component_indices[i], row, per_row[i]);
            if (orc_worker_append_row(worker, i, row, per_row[i])) SCAN_FAIL(ORC_ERR_INTERNAL, "%s", orc_last_error());
            if (!is_progressive) memset(mcu_row_coefficients[i], 0, per_row[i] * sizeof(int16_t));
        }
    }
    {
        int has;
        uint8_t m;
        SCAN_TRY(take_marker(d, &has, &m));
        while (has && is_rst(m)) {
            char saved[256];
            memcpy(saved, d->err, sizeof saved);
            if (read_marker(d, &m)) { /* .ok(): errors become None */
                has = 0;
                memcpy(d->err, saved, sizeof saved);
            }
        }
        *has_marker = has;
        *out_marker = m;
    }
    *has_data = 0;
    for (int i = 0; i < nc; i++)
        if (finished[i]) *has_data = 1;
    if (*has_data)
        for (int i = 0; i < nc; i++)
            if (finished[i]) {
                int ci = scan->component_indices[i];
                free(data[ci]);
                data[ci] = NULL;
                if (orc_worker_get_result(worker, i, &data[ci], &data_len[ci])) SCAN_FAIL(ORC_ERR_INTERNAL, "%s", orc_last_error());
            }
done:
    for (int i = 0; i < MAX_COMPONENTS; i++) free(mcu_row_coefficients[i]);
    return rc;
#undef SCAN_FAIL
#undef SCAN_TRY
}

/* src/decoder.rs:698-764 */
static int determine_color_transform(const orc_decoder *d) {
    if (d->has_color_transform) return d->color_transform;
    const frame_info *f = &d->frame;
    if (f->ncomp == 1) return ORC_CT_GRAYSCALE;
    if (f->ncomp == 3) {
        uint8_t a = f->comps[0].identifier, b = f->comps[1].identifier, c = f->comps[2].identifier;
        if (a == 1 && b == 2 && c == 3) return ORC_CT_YCBCR;
        if (a == 1 && b == 34 && c == 35) return ORC_CT_JCS_BG_YCC;
        if (a == 82 && b == 71 && c == 66) return ORC_CT_RGB;
        if (a == 114 && b == 103 && c == 98) return ORC_CT_JCS_BG_RGB;
        if (d->is_jfif) return ORC_CT_YCBCR;
    }
    if (d->has_adobe) {
        if (d->adobe_transform == 0) {
            if (f->ncomp == 3) return ORC_CT_RGB;
            if (f->ncomp == 4) return ORC_CT_CMYK;
        } else if (d->adobe_transform == 1) {
            return ORC_CT_YCBCR;
        } else {
            return ORC_CT_YCCK;
        }
    } else if (f->ncomp == 4) {
        return ORC_CT_CMYK;
    }
    if (f->ncomp == 4) return ORC_CT_YCCK;
    if (f->ncomp == 3) return ORC_CT_YCBCR;
    return ORC_CT_UNKNOWN;
}

/* src/decoder.rs:617-696 */
static int decode_planes(orc_decoder *d, orc_worker *worker) {
    const frame_info *frame = &d->frame;
    size_t need = (size_t)frame->ncomp * frame->output_w * frame->output_h;
    if (d->buffer_limit < need) FAIL(d, ORC_ERR_FORMAT, "size of decoded image exceeds maximum allowed size");
    if (frame->coding_process == 1 && d->has_coefficients) {
        for (int i = 0; i < frame->ncomp; i++) {
            const orc_component *component = &frame->comps[i];
            if (d->coefficients_finished[i] == ~(uint64_t)0) continue;
            if (!d->has_qt[component->tq]) continue;
            if (orc_worker_start(worker, i, component, d->qt[component->tq])) FAIL(d, ORC_ERR_INTERNAL, "%s", orc_last_error());
            size_t per_row = (size_t)component->block_w * component->v * 64;
            d->tap_len[i] = 0;
            for (uint32_t mcu_y = 0; mcu_y < frame->mcu_h; mcu_y++) {
                const int16_t *row = d->coefficients[i] + (size_t)mcu_y * per_row;
                tap_append(d, i, row, per_row);
                if (orc_worker_append_row(worker, i, row, per_row)) FAIL(d, ORC_ERR_INTERNAL, "%s", orc_last_error());
            }
            free(d->planes[i]);
            d->planes[i] = NULL;
            if (orc_worker_get_result(worker, i, &d->planes[i], &d->plane_len[i])) FAIL(d, ORC_ERR_INTERNAL, "%s", orc_last_error());
        }
    }
    if (frame->coding_process == 2) FAIL(d, ORC_ERR_UNSUPPORTED, "lossless JPEG is outside the oracle's scope");
    free(d->pixels);
    d->pixels = (uint8_t *)malloc(need ? need : 1);
    d->pixels_len = 0;
    d->final_ct = determine_color_transform(d);
    for (int i = 0; i < frame->ncomp; i++) {
        free(d->tap_plane[i]);
        d->tap_plane[i] = NULL;
        d->tap_plane_len[i] = d->plane_len[i];
        if (d->planes[i] && !d->no_taps) {
            d->tap_plane[i] = (uint8_t *)malloc(d->plane_len[i] ? d->plane_len[i] : 1);
            memcpy(d->tap_plane[i], d->planes[i], d->plane_len[i]);
        }
    }
    int rc = orc_compute_image_mt(d->arith, d->nthreads, frame->comps, frame->ncomp, (const uint8_t *const *)d->planes,
                                  d->plane_len, frame->output_w, frame->output_h, d->final_ct, d->pixels, need, &d->pixels_len);
    if (rc) FAIL(d, rc, "%s", orc_last_error());
    return ORC_OK;
}

/* src/decoder.rs:297-615 */
static int decode_internal(orc_decoder *d, int stop_after_metadata) {
    if (stop_after_metadata && d->has_frame) return ORC_OK;
    if (!d->has_frame) {
        uint8_t a, b;
        TRY(read_u8(d, &a));
        if (a != 0xFF) FAIL(d, ORC_ERR_FORMAT, "first two bytes are not an SOI marker");
        TRY(read_u8(d, &b));
        if (b != 0xD8) FAIL(d, ORC_ERR_FORMAT, "first two bytes are not an SOI marker");
    }
    uint8_t previous_marker = 0xD8;
    int has_pending = 0;
    uint8_t pending = 0;
    int scans_processed = 0;
    for (int i = 0; i < MAX_COMPONENTS; i++) {
        free(d->planes[i]);
        d->planes[i] = NULL;
        d->plane_len[i] = 0;
    }
    orc_worker *worker = orc_worker_new(d->arith);
    int rc = ORC_OK;
#define DI_FAIL(code, ...)                            \
    do {                                              \
        snprintf(d->err, sizeof d->err, __VA_ARGS__); \
        rc = (code);                                  \
        goto done;                                    \
    } while (0)
#define DI_TRY(x)          \
    do {                   \
        rc = (x);          \
        if (rc) goto done; \
    } while (0)
    for (;;) {
        uint8_t marker;
        if (has_pending) {
            marker = pending;
            has_pending = 0;
        } else {
            DI_TRY(read_marker(d, &marker));
        }
        if (is_sof(marker)) {
            if (d->has_frame) DI_FAIL(ORC_ERR_UNSUPPORTED, "Hierarchical");
            frame_info *f = (frame_info *)malloc(sizeof(frame_info));
            rc = parse_sof(d, marker, f);
            if (!rc && f->is_differential) {
                snprintf(d->err, sizeof d->err, "Hierarchical");
                rc = ORC_ERR_UNSUPPORTED;
            }
            if (!rc && f->arithmetic) {
                snprintf(d->err, sizeof d->err, "ArithmeticEntropyCoding");
                rc = ORC_ERR_UNSUPPORTED;
            }
            if (!rc && f->precision != 8 && f->coding_process != 2) {
                snprintf(d->err, sizeof d->err, "SamplePrecision(%u)", f->precision);
                rc = ORC_ERR_UNSUPPORTED;
            }
            if (!rc && !(f->precision >= 2 && f->precision <= 16)) {
                snprintf(d->err, sizeof d->err, "SamplePrecision(%u)", f->precision);
                rc = ORC_ERR_UNSUPPORTED;
            }
            if (!rc && f->ncomp != 1 && f->ncomp != 3 && f->ncomp != 4) {
                snprintf(d->err, sizeof d->err, "ComponentCount(%d)", f->ncomp);
                rc = ORC_ERR_UNSUPPORTED;
            }
            if (!rc) rc = validate_sampling(d, f);
            if (rc) {
                free(f);
                goto done;
            }
            d->frame = *f;
            d->has_frame = 1;
            free(f);
            if (stop_after_metadata) goto done;
        } else if (marker == 0xDA) { /* SOS */
            if (!d->has_frame) DI_FAIL(ORC_ERR_FORMAT, "scan encountered before frame");
            frame_info *frame = &d->frame;
            scan_info scan;
            DI_TRY(parse_sos(d, frame, &scan));
            if (frame->coding_process == 1 && !d->has_coefficients) {
                for (int i = 0; i < frame->ncomp; i++) {
                    size_t n = (size_t)frame->comps[i].block_w * frame->comps[i].block_h * 64;
                    d->coefficients[i] = (int16_t *)calloc(n ? n : 1, sizeof(int16_t));
                    d->coefficients_len[i] = n;
                }
                d->has_coefficients = 1;
            }
            if (frame->coding_process == 2) DI_FAIL(ORC_ERR_UNSUPPORTED, "lossless JPEG is outside the oracle's scope");
            int finished[MAX_COMPONENTS] = {0, 0, 0, 0};
            if (scan.al == 0) {
                for (int k = 0; k < scan.n; k++) {
                    int i = scan.component_indices[k];
                    if (d->coefficients_finished[i] == ~(uint64_t)0) continue;
                    for (int j = scan.ss_start; j < scan.ss_end; j++) d->coefficients_finished[i] |= (uint64_t)1 << j;
                    if (d->coefficients_finished[i] == ~(uint64_t)0) finished[k] = 1;
                }
            }
            uint8_t *data[MAX_COMPONENTS] = {0, 0, 0, 0};
            size_t data_len[MAX_COMPONENTS] = {0, 0, 0, 0};
            int has_data = 0, hm = 0;
            uint8_t m = 0;
            rc = decode_scan(d, frame, &scan, worker, finished, &hm, &m, data, data_len, &has_data);
            if (rc) {
                for (int i = 0; i < MAX_COMPONENTS; i++) free(data[i]);
                goto done;
            }
            if (has_data) {
                for (int i = 0; i < frame->ncomp && i < MAX_COMPONENTS; i++) {
                    if (!data[i] || data_len[i] == 0) {
                        free(data[i]);
                        continue;
                    }
                    if (d->coefficients_finished[i] == ~(uint64_t)0) {
                        free(d->planes[i]);
                        d->planes[i] = data[i];
                        d->plane_len[i] = data_len[i];
                    } else {
                        free(data[i]);
                    }
                }
            }
            has_pending = hm;
            pending = m;
            scans_processed++;
        } else if (marker == 0xDB) {
            DI_TRY(parse_dqt(d));
        } else if (marker == 0xC4) {
            DI_TRY(parse_dht(d));
        } else if (marker == 0xCC) {
            DI_FAIL(ORC_ERR_UNSUPPORTED, "ArithmeticEntropyCoding");
        } else if (marker == 0xDD) { /* DRI, src/parser.rs:592-600 */
            size_t length;
            DI_TRY(read_length(d, &length));
            if (length != 2) DI_FAIL(ORC_ERR_FORMAT, "DRI with invalid length");
            DI_TRY(read_u16(d, &d->restart_interval));
        } else if (marker == 0xFE) { /* COM */
            size_t length;
            DI_TRY(read_length(d, &length));
            if (d->pos + length > d->len) {
                d->pos = d->len;
                DI_FAIL(ORC_ERR_IO, "failed to fill whole buffer");
            }
            d->pos += length;
        } else if (is_app(marker)) {
            DI_TRY(parse_app(d, marker));
        } else if (is_rst(marker)) {
            if (previous_marker != 0xDA) DI_FAIL(ORC_ERR_FORMAT, "RST found outside of entropy-coded data");
        } else if (marker == 0xDC) { /* DNL */
            if (previous_marker != 0xDA || scans_processed != 1) DI_FAIL(ORC_ERR_FORMAT, "DNL is only allowed immediately after the first scan");
            DI_FAIL(ORC_ERR_UNSUPPORTED, "DNL");
        } else if (marker == 0xDE || marker == 0xDF) {
            DI_FAIL(ORC_ERR_UNSUPPORTED, "Hierarchical");
        } else if (marker == 0xD9) {
            break;
        } else {
            DI_FAIL(ORC_ERR_FORMAT, "marker 0x%02X found where not allowed", marker);
        }
        previous_marker = marker;
    }
    if (!d->has_frame) DI_FAIL(ORC_ERR_FORMAT, "end of image encountered before frame");
    rc = decode_planes(d, worker);
done:
    orc_worker_free(worker);
    return rc;
#undef DI_FAIL
#undef DI_TRY
}

/* ---- public ------------------------------------------------------------------------------- */
orc_decoder *orc_decoder_new(const uint8_t *data, size_t len, int arith) {
    orc_decoder *d = (orc_decoder *)calloc(1, sizeof *d);
    if (!d) return NULL;
    d->data = data;
    d->len = len;
    d->arith = arith;
    d->buffer_limit = (size_t)-1;
    return d;
}
void orc_decoder_free(orc_decoder *d) {
    if (!d) return;
    for (int i = 0; i < MAX_COMPONENTS; i++) {
        free(d->coefficients[i]);
        free(d->planes[i]);
        free(d->tap_coefs[i]);
        free(d->tap_plane[i]);
    }
    for (size_t i = 0; i < d->n_icc; i++) free(d->icc[i].data);
    free(d->icc);
    free(d->exif);
    free(d->xmp);
    free(d->pixels);
    free(d->icc_joined);
    free(d);
}
void orc_decoder_set_color_transform(orc_decoder *d, int ct) {
    d->has_color_transform = 1;
    d->color_transform = ct;
}
void orc_decoder_set_max_decoding_buffer_size(orc_decoder *d, size_t max) { d->buffer_limit = max; }
void orc_decoder_set_threads(orc_decoder *d, int nthreads) { d->nthreads = nthreads; }
void orc_decoder_set_taps(orc_decoder *d, int on) { d->no_taps = !on; }
int orc_decoder_read_info(orc_decoder *d) { return decode_internal(d, 1); }
/* src/decoder.rs:171-194 */
int orc_decoder_info(const orc_decoder *d, orc_image_info *info) {
    if (!d->has_frame) return 0;
    const frame_info *f = &d->frame;
    info->width = f->output_w;
    info->height = f->output_h;
    info->pixel_format = f->ncomp == 1 ? (f->precision <= 8 ? 0 : 1) : (f->ncomp == 3 ? 2 : 3);
    info->coding_process = f->coding_process;
    return 1;
}
/* src/decoder.rs:278-290 + src/parser.rs:120-133 */
int orc_decoder_scale(orc_decoder *d, uint16_t req_w, uint16_t req_h, uint16_t *w, uint16_t *h) {
    TRY(orc_decoder_read_info(d));
    frame_info *f = &d->frame;
    int idct_size = orc_choose_idct_size(f->image_w, f->image_h, req_w, req_h);
    for (int i = 0; i < f->ncomp; i++) f->comps[i].dct_scale = (uint16_t)idct_size;
    uint16_t mw, mh;
    if (orc_update_component_sizes(f->image_w, f->image_h, f->comps, f->ncomp, &mw, &mh)) FAIL(d, ORC_ERR_FORMAT, "invalid dimensions");
    /* (w as f32 * idct_size as f32 / 8.0).ceil() as u16 */
    float fw = (float)f->image_w * (float)idct_size / 8.0f, fh = (float)f->image_h * (float)idct_size / 8.0f;
    uint16_t ow = (uint16_t)fw, oh = (uint16_t)fh;
    if ((float)ow < fw) ow++;
    if ((float)oh < fh) oh++;
    f->output_w = ow;
    f->output_h = oh;
    *w = ow;
    *h = oh;
    return ORC_OK;
}
int orc_decoder_decode(orc_decoder *d, const uint8_t **pixels, size_t *len) {
    int rc = decode_internal(d, 0);
    if (rc) return rc;
    *pixels = d->pixels;
    *len = d->pixels_len;
    return ORC_OK;
}
const char *orc_decoder_error(const orc_decoder *d) { return d->err; }
int orc_decoder_color_transform(const orc_decoder *d) { return d->final_ct; }
int orc_decoder_ncomp(const orc_decoder *d) { return d->has_frame ? d->frame.ncomp : 0; }
int orc_decoder_component(const orc_decoder *d, int i, orc_component *c, uint16_t qt[64]) {
    if (!d->has_frame || i < 0 || i >= d->frame.ncomp) return 0;
    *c = d->frame.comps[i];
    if (d->has_qt[c->tq]) memcpy(qt, d->qt[c->tq], 128);
    else memset(qt, 0, 128);
    return 1;
}
int orc_decoder_coefficients(const orc_decoder *d, int i, const int16_t **coefs, size_t *n_i16) {
    if (i < 0 || i >= MAX_COMPONENTS || !d->tap_coefs[i]) return 0;
    *coefs = d->tap_coefs[i];
    *n_i16 = d->tap_len[i];
    return 1;
}
int orc_decoder_plane(const orc_decoder *d, int i, const uint8_t **plane, size_t *len) {
    if (i < 0 || i >= MAX_COMPONENTS || !d->tap_plane[i]) return 0;
    *plane = d->tap_plane[i];
    *len = d->tap_plane_len[i];
    return 1;
}
/* src/decoder.rs:211-241 */
int orc_decoder_icc_profile(const orc_decoder *dc, const uint8_t **data, size_t *len) {
    orc_decoder *d = (orc_decoder *)dc;
    size_t num = d->n_icc;
    if (num == 0 || num >= 255) return 0;
    const icc_chunk *present[256];
    memset(present, 0, sizeof present);
    for (size_t i = 0; i < num; i++) {
        const icc_chunk *c = &d->icc[i];
        if (c->num_markers != num) return 0;
        if (c->seq_no == 0) return 0;
        if (present[c->seq_no]) return 0;
        present[c->seq_no] = c;
    }
    size_t total = 0;
    for (size_t s = 1; s <= num; s++) {
        if (!present[s]) return 0;
        total += present[s]->len;
    }
    free(d->icc_joined);
    d->icc_joined = (uint8_t *)malloc(total ? total : 1);
    size_t off = 0;
    for (size_t s = 1; s <= num; s++) {
        memcpy(d->icc_joined + off, present[s]->data, present[s]->len);
        off += present[s]->len;
    }
    d->icc_joined_len = total;
    *data = d->icc_joined;
    *len = total;
    return 1;
}
int orc_decoder_exif(const orc_decoder *d, const uint8_t **data, size_t *len) {
    if (!d->has_exif) return 0;
    *data = d->exif;
    *len = d->exif_len;
    return 1;
}
int orc_decoder_xmp(const orc_decoder *d, const uint8_t **data, size_t *len) {
    if (!d->has_xmp) return 0;
    *data = d->xmp;
    *len = d->xmp_len;
    return 1;
}
