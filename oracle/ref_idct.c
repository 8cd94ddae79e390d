/*
 * ref_idct.c -- oracle (test infrastructure, see oracle.h): dequantise + inverse DCT of one block,
 * restating /root/reference/src/idct.rs and /root/reference/src/arch/ssse3.rs.
 *
 * All scalar arithmetic is Wrapping<i32> in the reference; here it is done on uint32_t (wraps by
 * definition) and viewed as int32_t for the arithmetic right shifts.
 */
#include "oracle.h"

#include <string.h>

typedef uint32_t u32;
typedef int32_t i32;

static inline i32 sar(u32 x, int n) { return (i32)x >> n; } /* gcc: arithmetic shift */

/* src/idct.rs:568-570 */
static inline uint8_t stbi_clamp(i32 x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

/* src/idct.rs:572-574 : (x * 4096.0 + 0.5) as i32, evaluated in f32 */
static inline u32 f2f(float x) { return (u32)(i32)(x * 4096.0f + 0.5f); }

/* src/idct.rs:450-452 */
static inline u32 dequantize(int16_t c, uint16_t q) { return (u32)((i32)c * (i32)q); }

/* src/idct.rs:378-447 : kernel_x + kernel_t */
static void kernel(const u32 s[8], u32 x_scale, u32 xs[4], u32 ts[4]) {
    /* even part, src/idct.rs:378-407 */
    u32 p2 = s[2], p3 = s[6];
    u32 p1 = (p2 + p3) * f2f(0.5411961f);
    u32 t2 = p1 + p3 * f2f(-1.847759065f);
    u32 t3 = p1 + p2 * f2f(0.765366865f);
    p2 = s[0];
    p3 = s[4];
    u32 t0 = (p2 + p3) << 12;
    u32 t1 = (p2 - p3) << 12;
    xs[0] = t0 + t3 + x_scale;
    xs[3] = t0 - t3 + x_scale;
    xs[1] = t1 + t2 + x_scale;
    xs[2] = t1 - t2 + x_scale;
    /* odd part, src/idct.rs:410-439 */
    t0 = s[7];
    t1 = s[5];
    t2 = s[3];
    t3 = s[1];
    p3 = t0 + t2;
    u32 p4 = t1 + t3;
    p1 = t0 + t3;
    p2 = t1 + t2;
    u32 p5 = (p3 + p4) * f2f(1.175875602f);
    t0 *= f2f(0.298631336f);
    t1 *= f2f(2.053119869f);
    t2 *= f2f(3.072711026f);
    t3 *= f2f(1.501321110f);
    p1 = p5 + p1 * f2f(-0.899976223f);
    p2 = p5 + p2 * f2f(-2.562915447f);
    p3 = p3 * f2f(-1.961570560f);
    p4 = p4 * f2f(-0.390180644f);
    t3 += p1 + p4;
    t2 += p2 + p3;
    t1 += p2 + p4;
    t0 += p1 + p3;
    ts[0] = t0;
    ts[1] = t1;
    ts[2] = t2;
    ts[3] = t3;
}

/* src/idct.rs:260-370 */
static void idct8x8_scalar(const int16_t c[64], const uint16_t q[64], size_t stride, uint8_t *out) {
    u32 temp[64];
    for (int i = 0; i < 8; i++) {
        /* zero-AC column shortcut, src/idct.rs:279-295 (tested on the raw coefficients) */
        if (c[i + 8] == 0 && c[i + 16] == 0 && c[i + 24] == 0 && c[i + 32] == 0 && c[i + 40] == 0 &&
            c[i + 48] == 0 && c[i + 56] == 0) {
            u32 dc = dequantize(c[i], q[i]) << 2;
            for (int k = 0; k < 8; k++) temp[i + 8 * k] = dc;
        } else {
            u32 s[8], x[4], t[4];
            for (int k = 0; k < 8; k++) s[k] = dequantize(c[i + 8 * k], q[i + 8 * k]);
            kernel(s, 512, x, t);
            temp[i] = (u32)sar(x[0] + t[3], 10);
            temp[i + 56] = (u32)sar(x[0] - t[3], 10);
            temp[i + 8] = (u32)sar(x[1] + t[2], 10);
            temp[i + 48] = (u32)sar(x[1] - t[2], 10);
            temp[i + 16] = (u32)sar(x[2] + t[1], 10);
            temp[i + 40] = (u32)sar(x[2] - t[1], 10);
            temp[i + 24] = (u32)sar(x[3] + t[0], 10);
            temp[i + 32] = (u32)sar(x[3] - t[0], 10);
        }
    }
    const u32 X_SCALE = 65536u + (128u << 17); /* src/idct.rs:336 */
    for (int r = 0; r < 8; r++) {
        const u32 *ch = temp + 8 * r;
        uint8_t *o = out + (size_t)r * stride;
        if (ch[1] == 0 && ch[2] == 0 && ch[3] == 0 && ch[4] == 0 && ch[5] == 0 && ch[6] == 0 && ch[7] == 0) {
            /* src/idct.rs:344-353 */
            uint8_t dc = stbi_clamp(sar((ch[0] << 12) + X_SCALE, 17));
            for (int k = 0; k < 8; k++) o[k] = dc;
        } else {
            u32 x[4], t[4];
            kernel(ch, X_SCALE, x, t);
            o[0] = stbi_clamp(sar(x[0] + t[3], 17));
            o[7] = stbi_clamp(sar(x[0] - t[3], 17));
            o[1] = stbi_clamp(sar(x[1] + t[2], 17));
            o[6] = stbi_clamp(sar(x[1] - t[2], 17));
            o[2] = stbi_clamp(sar(x[2] + t[1], 17));
            o[5] = stbi_clamp(sar(x[2] - t[1], 17));
            o[3] = stbi_clamp(sar(x[3] + t[0], 17));
            o[4] = stbi_clamp(sar(x[3] - t[0], 17));
        }
    }
}

/* src/idct.rs:456-517 */
static void idct4x4(const int16_t c[64], const uint16_t q[64], size_t stride, uint8_t *out) {
    enum { CONST_BITS = 12, PASS1_BITS = 2, FINAL_BITS = CONST_BITS + PASS1_BITS + 3 };
    u32 temp[16];
    for (int i = 0; i < 4; i++) {
        u32 s0 = dequantize(c[i], q[i]);
        u32 s1 = dequantize(c[i + 8], q[i + 8]);
        u32 s2 = dequantize(c[i + 16], q[i + 16]);
        u32 s3 = dequantize(c[i + 24], q[i + 24]);
        u32 x0 = (s0 + s2) << PASS1_BITS;
        u32 x2 = (s0 - s2) << PASS1_BITS;
        u32 p1 = (s1 + s3) * f2f(0.541196100f);
        u32 t0 = (u32)sar(p1 + s3 * f2f(-1.847759065f) + 512u, CONST_BITS - PASS1_BITS);
        u32 t2 = (u32)sar(p1 + s1 * f2f(0.765366865f) + 512u, CONST_BITS - PASS1_BITS);
        temp[i] = x0 + t2;
        temp[i + 12] = x0 - t2;
        temp[i + 4] = x2 + t0;
        temp[i + 8] = x2 - t0;
    }
    for (int i = 0; i < 4; i++) {
        u32 s0 = temp[i * 4], s1 = temp[i * 4 + 1], s2 = temp[i * 4 + 2], s3 = temp[i * 4 + 3];
        u32 x0 = (s0 + s2) << CONST_BITS;
        u32 x2 = (s0 - s2) << CONST_BITS;
        u32 p1 = (s1 + s3) * f2f(0.541196100f);
        u32 t0 = p1 + s3 * f2f(-1.847759065f);
        u32 t2 = p1 + s1 * f2f(0.765366865f);
        x0 += (1u << (FINAL_BITS - 1)) + (128u << FINAL_BITS);
        x2 += (1u << (FINAL_BITS - 1)) + (128u << FINAL_BITS);
        uint8_t *o = out + (size_t)i * stride;
        o[0] = stbi_clamp(sar(x0 + t2, FINAL_BITS));
        o[3] = stbi_clamp(sar(x0 - t2, FINAL_BITS));
        o[1] = stbi_clamp(sar(x2 + t0, FINAL_BITS));
        o[2] = stbi_clamp(sar(x2 - t0, FINAL_BITS));
    }
}

/* src/idct.rs:519-553 */
static void idct2x2(const int16_t c[64], const uint16_t q[64], size_t stride, uint8_t *out) {
    enum { SCALE_BITS = 3 };
    u32 s00 = dequantize(c[0], q[0]), s10 = dequantize(c[8], q[8]);
    u32 x0 = s00 + s10, x2 = s00 - s10;
    u32 s01 = dequantize(c[1], q[1]), s11 = dequantize(c[9], q[9]);
    u32 x1 = s01 + s11, x3 = s01 - s11;
    x0 += (1u << (SCALE_BITS - 1)) + (128u << SCALE_BITS);
    x2 += (1u << (SCALE_BITS - 1)) + (128u << SCALE_BITS);
    out[0] = stbi_clamp(sar(x0 + x1, SCALE_BITS));
    out[1] = stbi_clamp(sar(x0 - x1, SCALE_BITS));
    out[stride] = stbi_clamp(sar(x2 + x3, SCALE_BITS));
    out[stride + 1] = stbi_clamp(sar(x2 - x3, SCALE_BITS));
}

/* src/idct.rs:555-565 : Wrapping<i32> division truncates toward zero */
static void idct1x1(const int16_t c[64], const uint16_t q[64], uint8_t *out) {
    i32 s0 = (i32)(dequantize(c[0], q[0]) + 128u * 8u) / 8;
    out[0] = stbi_clamp(s0);
}

/* ---- SSSE3 variant, portable emulation of the 16-bit lanes (src/arch/ssse3.rs) ------------- */
static inline int16_t sat16(i32 v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }
static inline int16_t adds(int16_t a, int16_t b) { return sat16((i32)a + b); }
static inline int16_t subs(int16_t a, int16_t b) { return sat16((i32)a - b); }
/* _mm_mulhrs_epi16 */
static inline int16_t mulhrs(int16_t a, int16_t b) { return (int16_t)(((((i32)a * b) >> 14) + 1) >> 1); }

/* src/arch/ssse3.rs:8-84, one lane */
static void idct8_lane(int16_t d[8]) {
    int16_t p2 = d[2], p3 = d[6];
    int16_t p1 = mulhrs(adds(p2, p3), 17734);
    int16_t t2 = subs(subs(p1, p3), mulhrs(p3, 27779));
    int16_t t3 = adds(p1, mulhrs(p2, 25079));
    p2 = d[0];
    p3 = d[4];
    int16_t t0 = adds(p2, p3);
    int16_t t1 = subs(p2, p3);
    int16_t x0 = adds(t0, t3), x3 = subs(t0, t3), x1 = adds(t1, t2), x2 = subs(t1, t2);
    t0 = d[7];
    t1 = d[5];
    t2 = d[3];
    t3 = d[1];
    p3 = adds(t0, t2);
    int16_t p4 = adds(t1, t3);
    p1 = adds(t0, t3);
    p2 = adds(t1, t2);
    int16_t p5 = adds(p3, p4);
    p5 = adds(p5, mulhrs(p5, 5763));
    t0 = mulhrs(t0, 9786);
    t1 = adds(adds(t1, t1), mulhrs(t1, 1741));
    t2 = adds(adds(t2, adds(t2, t2)), mulhrs(t2, 2383));
    t3 = adds(t3, mulhrs(t3, 16427));
    p1 = subs(p5, mulhrs(p1, 29490));
    p2 = subs(subs(subs(p5, p2), p2), mulhrs(p2, 18446));
    p3 = subs(mulhrs(p3, -31509), p3);
    p4 = mulhrs(p4, -12785);
    t3 = adds(adds(p1, p4), t3);
    t2 = adds(adds(p2, p3), t2);
    t1 = adds(adds(p2, p4), t1);
    t0 = adds(adds(p1, p3), t0);
    d[0] = adds(x0, t3);
    d[7] = subs(x0, t3);
    d[1] = adds(x1, t2);
    d[6] = subs(x1, t2);
    d[2] = adds(x2, t1);
    d[5] = subs(x2, t1);
    d[3] = adds(x3, t0);
    d[4] = subs(x3, t0);
}

/* src/arch/ssse3.rs:124-192 */
static void idct8x8_ssse3_emul(const int16_t c[64], const uint16_t q[64], size_t stride, uint8_t *out) {
    int16_t data[8][8];
    for (int r = 0; r < 8; r++)
        for (int k = 0; k < 8; k++) {
            /* _mm_mullo_epi16 then _mm_slli_epi16(.., 3): both wrap at 16 bits */
            uint16_t prod = (uint16_t)((u32)(uint16_t)c[8 * r + k] * (u32)q[8 * r + k]);
            data[r][k] = (int16_t)(uint16_t)(prod << 3);
        }
    /* idct8 works lane-wise across the eight row vectors = down the columns */
    for (int k = 0; k < 8; k++) {
        int16_t d[8];
        for (int r = 0; r < 8; r++) d[r] = data[r][k];
        idct8_lane(d);
        for (int r = 0; r < 8; r++) data[r][k] = d[r];
    }
    /* transpose8; idct8; transpose8  ==  the same 1-D transform along each row */
    for (int r = 0; r < 8; r++) idct8_lane(data[r]);
    for (int r = 0; r < 8; r++)
        for (int k = 0; k < 8; k++) {
            int16_t v = adds(data[r][k], 8224); /* OFFSET + ROUNDING_BIAS, ssse3.rs:173-177 */
            i32 s = v >> 6;
            out[(size_t)r * stride + k] = (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s)); /* packus */
        }
}

#if defined(__SSSE3__)
#include <tmmintrin.h>
static void idct8_v(__m128i d[8]) {
    __m128i p2 = d[2], p3 = d[6];
    __m128i p1 = _mm_mulhrs_epi16(_mm_adds_epi16(p2, p3), _mm_set1_epi16(17734));
    __m128i t2 = _mm_subs_epi16(_mm_subs_epi16(p1, p3), _mm_mulhrs_epi16(p3, _mm_set1_epi16(27779)));
    __m128i t3 = _mm_adds_epi16(p1, _mm_mulhrs_epi16(p2, _mm_set1_epi16(25079)));
    p2 = d[0];
    p3 = d[4];
    __m128i t0 = _mm_adds_epi16(p2, p3), t1 = _mm_subs_epi16(p2, p3);
    __m128i x0 = _mm_adds_epi16(t0, t3), x3 = _mm_subs_epi16(t0, t3);
    __m128i x1 = _mm_adds_epi16(t1, t2), x2 = _mm_subs_epi16(t1, t2);
    t0 = d[7];
    t1 = d[5];
    t2 = d[3];
    t3 = d[1];
    p3 = _mm_adds_epi16(t0, t2);
    __m128i p4 = _mm_adds_epi16(t1, t3);
    p1 = _mm_adds_epi16(t0, t3);
    p2 = _mm_adds_epi16(t1, t2);
    __m128i p5 = _mm_adds_epi16(p3, p4);
    p5 = _mm_adds_epi16(p5, _mm_mulhrs_epi16(p5, _mm_set1_epi16(5763)));
    t0 = _mm_mulhrs_epi16(t0, _mm_set1_epi16(9786));
    t1 = _mm_adds_epi16(_mm_adds_epi16(t1, t1), _mm_mulhrs_epi16(t1, _mm_set1_epi16(1741)));
    t2 = _mm_adds_epi16(_mm_adds_epi16(t2, _mm_adds_epi16(t2, t2)), _mm_mulhrs_epi16(t2, _mm_set1_epi16(2383)));
    t3 = _mm_adds_epi16(t3, _mm_mulhrs_epi16(t3, _mm_set1_epi16(16427)));
    p1 = _mm_subs_epi16(p5, _mm_mulhrs_epi16(p1, _mm_set1_epi16(29490)));
    p2 = _mm_subs_epi16(_mm_subs_epi16(_mm_subs_epi16(p5, p2), p2), _mm_mulhrs_epi16(p2, _mm_set1_epi16(18446)));
    p3 = _mm_subs_epi16(_mm_mulhrs_epi16(p3, _mm_set1_epi16(-31509)), p3);
    p4 = _mm_mulhrs_epi16(p4, _mm_set1_epi16(-12785));
    t3 = _mm_adds_epi16(_mm_adds_epi16(p1, p4), t3);
    t2 = _mm_adds_epi16(_mm_adds_epi16(p2, p3), t2);
    t1 = _mm_adds_epi16(_mm_adds_epi16(p2, p4), t1);
    t0 = _mm_adds_epi16(_mm_adds_epi16(p1, p3), t0);
    d[0] = _mm_adds_epi16(x0, t3);
    d[7] = _mm_subs_epi16(x0, t3);
    d[1] = _mm_adds_epi16(x1, t2);
    d[6] = _mm_subs_epi16(x1, t2);
    d[2] = _mm_adds_epi16(x2, t1);
    d[5] = _mm_subs_epi16(x2, t1);
    d[3] = _mm_adds_epi16(x3, t0);
    d[4] = _mm_subs_epi16(x3, t0);
}
static void transpose8_v(__m128i d[8]) {
    __m128i a0 = _mm_unpacklo_epi16(d[0], d[1]), a1 = _mm_unpacklo_epi16(d[2], d[3]);
    __m128i a2 = _mm_unpacklo_epi16(d[4], d[5]), a3 = _mm_unpacklo_epi16(d[6], d[7]);
    __m128i b0 = _mm_unpackhi_epi16(d[0], d[1]), b1 = _mm_unpackhi_epi16(d[2], d[3]);
    __m128i b2 = _mm_unpackhi_epi16(d[4], d[5]), b3 = _mm_unpackhi_epi16(d[6], d[7]);
    __m128i c0 = _mm_unpacklo_epi32(a0, a1), c1 = _mm_unpackhi_epi32(a0, a1);
    __m128i c2 = _mm_unpacklo_epi32(a2, a3), c3 = _mm_unpackhi_epi32(a2, a3);
    __m128i e0 = _mm_unpacklo_epi32(b0, b1), e1 = _mm_unpackhi_epi32(b0, b1);
    __m128i e2 = _mm_unpacklo_epi32(b2, b3), e3 = _mm_unpackhi_epi32(b2, b3);
    d[0] = _mm_unpacklo_epi64(c0, c2);
    d[1] = _mm_unpackhi_epi64(c0, c2);
    d[2] = _mm_unpacklo_epi64(c1, c3);
    d[3] = _mm_unpackhi_epi64(c1, c3);
    d[4] = _mm_unpacklo_epi64(e0, e2);
    d[5] = _mm_unpackhi_epi64(e0, e2);
    d[6] = _mm_unpacklo_epi64(e1, e3);
    d[7] = _mm_unpackhi_epi64(e1, e3);
}
int orc_idct8x8_ssse3_intrin(const int16_t c[64], const uint16_t q[64], size_t stride, uint8_t *out) {
    __m128i d[8];
    for (int i = 0; i < 8; i++)
        d[i] = _mm_slli_epi16(_mm_mullo_epi16(_mm_loadu_si128((const __m128i *)(c + 8 * i)),
                                              _mm_loadu_si128((const __m128i *)(q + 8 * i))),
                              3);
    idct8_v(d);
    transpose8_v(d);
    idct8_v(d);
    transpose8_v(d);
    for (int i = 0; i < 8; i++) {
        uint8_t buf[16];
        __m128i v = _mm_adds_epi16(d[i], _mm_set1_epi16(8224));
        _mm_storeu_si128((__m128i *)buf, _mm_packus_epi16(_mm_srai_epi16(v, 6), _mm_setzero_si128()));
        memcpy(out + (size_t)i * stride, buf, 8);
    }
    return 1;
}
#else
int orc_idct8x8_ssse3_intrin(const int16_t c[64], const uint16_t q[64], size_t stride, uint8_t *out) {
    (void)c; (void)q; (void)stride; (void)out;
    return 0;
}
#endif

/* src/idct.rs:205-257 : dispatch on scale; the SSSE3 variant only replaces the 8x8 kernel
 * (src/idct.rs:247-253, src/arch/mod.rs:37-57). */
void orc_idct_block(int arith, int scale, const int16_t c[64], const uint16_t q[64], size_t stride,
                    uint8_t *out) {
    switch (scale) {
    case 8:
        if (arith == ORC_ARITH_SSSE3_NATIVE && orc_idct8x8_ssse3_intrin(c, q, stride, out))
            break;
        if (arith == ORC_ARITH_SSSE3 || arith == ORC_ARITH_SSSE3_NATIVE)
            idct8x8_ssse3_emul(c, q, stride, out);
        else
            idct8x8_scalar(c, q, stride, out);
        break;
    case 4: idct4x4(c, q, stride, out); break;
    case 2: idct2x2(c, q, stride, out); break;
    case 1: idct1x1(c, q, out); break;
    default: break;
    }
}

/* src/idct.rs:14-28 */
int orc_choose_idct_size(uint16_t full_w, uint16_t full_h, uint16_t req_w, uint16_t req_h) {
    static const int scales[3] = {1, 2, 4};
    for (int k = 0; k < 3; k++) {
        u32 s = (u32)scales[k];
        uint16_t sw = (uint16_t)(((u32)full_w * s - 1) / 8 + 1);
        uint16_t sh = (uint16_t)(((u32)full_h * s - 1) / 8 + 1);
        if (sw >= req_w || sh >= req_h) return scales[k];
    }
    return 8;
}

/* src/parser.rs:283-310 */
static int ceil_div(u32 x, u32 y, uint16_t *r) {
    if (x == 0 || y == 0) return ORC_ERR_FORMAT;
    *r = (uint16_t)(1 + ((x - 1) / y));
    return ORC_OK;
}
int orc_update_component_sizes(uint16_t w, uint16_t h, orc_component *comps, int n, uint16_t *mcu_w,
                               uint16_t *mcu_h) {
    u32 h_max = 0, v_max = 0;
    for (int i = 0; i < n; i++) {
        if (comps[i].h > h_max) h_max = comps[i].h;
        if (comps[i].v > v_max) v_max = comps[i].v;
    }
    uint16_t mw, mh;
    if (ceil_div(w, h_max * 8, &mw) || ceil_div(h, v_max * 8, &mh)) return ORC_ERR_FORMAT;
    for (int i = 0; i < n; i++) {
        orc_component *c = &comps[i];
        if (ceil_div((u32)w * c->h * c->dct_scale, h_max * 8, &c->size_w)) return ORC_ERR_FORMAT;
        if (ceil_div((u32)h * c->v * c->dct_scale, v_max * 8, &c->size_h)) return ORC_ERR_FORMAT;
        c->block_w = (uint16_t)(mw * c->h);
        c->block_h = (uint16_t)(mh * c->v);
    }
    *mcu_w = mw;
    *mcu_h = mh;
    return ORC_OK;
}
