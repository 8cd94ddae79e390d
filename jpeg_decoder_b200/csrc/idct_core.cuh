// idct_core.cuh -- the 8x8 dequantise + inverse DCT arithmetic shared by K1 (k1_idct.cu) and the fused
// kernel (kf_fused.cu).  Reference: dequantize_and_idct_block_8x8 (src/idct.rs:241-370), kernel (378-447),
// stbi_f2f / stbi_fsh / stbi_clamp (567-578).  Arithmetic specification: SURVEY.md Appendix A.1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "ptx.cuh"

namespace b200jpg {

// ---------------------------------------------------------------------------------------------
// Scalar arithmetic, src/idct.rs.  Everything is Wrapping<i32>: unsigned ops wrap, `>>` on int is
// an arithmetic shift.
// ---------------------------------------------------------------------------------------------
// stbi_f2f(x) = (x * 4096 + 0.5) as i32 in f32, src/idct.rs:572-574; values checked in
// tests/test_oracle_kat.py against the oracle, which evaluates the f32 expression.
#define F2F_0_5411961 2217u
#define F2F_N1_847759065 ((unsigned)-7567)
#define F2F_0_765366865 3135u
#define F2F_1_175875602 4816u
#define F2F_0_298631336 1223u
#define F2F_2_053119869 8410u
#define F2F_3_072711026 12586u
#define F2F_1_501321110 6149u
#define F2F_N0_899976223 ((unsigned)-3685)
#define F2F_N2_562915447 ((unsigned)-10497)
#define F2F_N1_961570560 ((unsigned)-8034)
#define F2F_N0_390180644 ((unsigned)-1597)

__device__ __forceinline__ int sar(unsigned x, int n) { return (int)x >> n; }

// bytes (b0,b1,b2,b3) = clamp(v0..v3)
__device__ __forceinline__ unsigned pack4_sat_u8(int v0, int v1, int v2, int v3) {
    return pack_sat_u8(v1, v0, pack_sat_u8(v3, v2, 0u));
}

// One 1-D pass of the stb_image butterfly, src/idct.rs:378-447.  `s0` already carries any bias the
// caller folded in; xs is added to the even part (x_scale).
#define IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7, xs, x0, x1, x2, x3, t0, t1, t2, t3) \
    {                                                                                \
        unsigned p1_ = ((s2) + (s6)) * F2F_0_5411961;                                \
        unsigned e2_ = p1_ + (s6) * F2F_N1_847759065;                                \
        unsigned e3_ = p1_ + (s2) * F2F_0_765366865;                                 \
        unsigned e0_ = (((s0) + (s4)) << 12) + (xs);                                 \
        unsigned e1_ = (((s0) - (s4)) << 12) + (xs);                                 \
        x0 = e0_ + e3_;                                                              \
        x3 = e0_ - e3_;                                                              \
        x1 = e1_ + e2_;                                                              \
        x2 = e1_ - e2_;                                                              \
        unsigned q3_ = (s7) + (s3), q4_ = (s5) + (s1), q1_ = (s7) + (s1), q2_ = (s5) + (s3); \
        unsigned p5_ = (q3_ + q4_) * F2F_1_175875602;                                \
        q1_ = p5_ + q1_ * F2F_N0_899976223;                                          \
        q2_ = p5_ + q2_ * F2F_N2_562915447;                                          \
        q3_ = q3_ * F2F_N1_961570560;                                                \
        q4_ = q4_ * F2F_N0_390180644;                                                \
        t3 = (s1) * F2F_1_501321110 + (q1_ + q4_);                                   \
        t2 = (s3) * F2F_3_072711026 + (q2_ + q3_);                                   \
        t1 = (s5) * F2F_2_053119869 + (q2_ + q4_);                                   \
        t0 = (s7) * F2F_0_298631336 + (q1_ + q3_);                                   \
    }

// add / subtract issued on the FMA pipe: IMAD with a multiplier (+1 / -1) that only the host knows,
// read from the kernel-parameter constant bank so ptxas cannot fold it back into an IADD3
#define fma_add(a, b) ((a) * p.one + (b))
#define fma_sub(a, b) ((b) * p.minus_one + (a))
// output butterfly x +- t in one of two styles: 0 = plain add (ptxas picks IADD3, ALU pipe, half rate),
// 1 = IMAD with the opaque +-1 multiplier (FMA pipe, full rate)
#define K1_BFLY_ADD(STYLE, x, t) ((STYLE) == 1 ? fma_add((t), (x)) : ((x) + (t)))
#define K1_BFLY_SUB(STYLE, x, t) ((STYLE) == 1 ? fma_sub((x), (t)) : ((x) - (t)))

// ---------------------------------------------------------------------------------------------
// The same 1-D pass in "direct form".  All arithmetic of src/idct.rs:378-447 between two shifts is
// Wrapping<i32>, i.e. a linear map over the ring Z/2^32, so any regrouping of its sums and products is
// bit-exact.  The odd half (t0..t3 from s1,s3,s5,s7) is evaluated as a 4x4 constant matrix whose IMAD
// chains are seeded with the even half, so x + t costs the four IMADs alone (instead of 9 IMADs + 13 adds
// + the add itself) and x - t is 2x - (x + t), one IADD3; the even rotation is 2x2.  40 instructions per
// pass including the eight shifts (46 in the butterfly form), split evenly between the FMA pipe (IMAD) and
// the ALU pipe (IADD3/LEA/SHF), which both issue every other cycle on sm_100.
// The matrix entries are spelled as sums of the reference's constants so the identity is visible:
//   t3 = s1*G1 + q1' + q4',  q1' = p5 + (s7+s1)*E1,  q4' = (s5+s1)*E4,  p5 = (s1+s3+s5+s7)*D   etc.
// ---------------------------------------------------------------------------------------------
#define ID_A F2F_0_5411961
#define ID_B F2F_N1_847759065
#define ID_C F2F_0_765366865
#define ID_D F2F_1_175875602
#define ID_E1 F2F_N0_899976223
#define ID_E2 F2F_N2_562915447
#define ID_E3 F2F_N1_961570560
#define ID_E4 F2F_N0_390180644
#define ID_G1 F2F_1_501321110
#define ID_G3 F2F_3_072711026
#define ID_G5 F2F_2_053119869
#define ID_G7 F2F_0_298631336
// o0..o7 = outputs in natural order (x0+t3, x1+t2, x2+t1, x3+t0, x3-t0, x2-t1, x1-t2, x0-t3), not yet shifted
#define IDCT_1D_DIRECT(s0, s1, s2, s3, s4, s5, s6, s7, xs, o0, o1, o2, o3, o4, o5, o6, o7)                        \
    {                                                                                                              \
        const unsigned e3_ = (s2) * (unsigned)(ID_A + ID_C) + (s6) * (unsigned)(ID_A);                            \
        const unsigned e2_ = (s2) * (unsigned)(ID_A) + (s6) * (unsigned)(ID_A + ID_B);                            \
        const unsigned e0_ = (((s0) + (s4)) << 12) + (xs);                                                         \
        const unsigned e1_ = (((s0) - (s4)) << 12) + (xs);                                                         \
        const unsigned x0_ = e0_ + e3_, x3_ = e0_ - e3_, x1_ = e1_ + e2_, x2_ = e1_ - e2_;                          \
        /* x + t as one IMAD chain seeded with x; x - t = 2x - (x + t) */                                           \
        o0 = x0_ + (s1) * (unsigned)(ID_G1 + ID_D + ID_E1 + ID_E4) + (s3) * (unsigned)(ID_D) +                     \
             (s5) * (unsigned)(ID_D + ID_E4) + (s7) * (unsigned)(ID_D + ID_E1);                                    \
        o1 = x1_ + (s1) * (unsigned)(ID_D) + (s3) * (unsigned)(ID_G3 + ID_D + ID_E2 + ID_E3) +                     \
             (s5) * (unsigned)(ID_D + ID_E2) + (s7) * (unsigned)(ID_D + ID_E3);                                    \
        o2 = x2_ + (s1) * (unsigned)(ID_D + ID_E4) + (s3) * (unsigned)(ID_D + ID_E2) +                             \
             (s5) * (unsigned)(ID_G5 + ID_D + ID_E2 + ID_E4) + (s7) * (unsigned)(ID_D);                            \
        o3 = x3_ + (s1) * (unsigned)(ID_D + ID_E1) + (s3) * (unsigned)(ID_D + ID_E3) +                             \
             (s5) * (unsigned)(ID_D) + (s7) * (unsigned)(ID_G7 + ID_D + ID_E1 + ID_E3);                            \
        o7 = x0_ + x0_ - o0;                                                                                       \
        o6 = x1_ + x1_ - o1;                                                                                       \
        o5 = x2_ + x2_ - o2;                                                                                       \
        o4 = x3_ + x3_ - o3;                                                                                       \
    }

__device__ __forceinline__ unsigned sext_lo(unsigned w) { return (unsigned)(int)(short)(w & 0xffffu); }
__device__ __forceinline__ unsigned sext_hi(unsigned w) { return (unsigned)((int)w >> 16); }

// Full-precision reference form of one 8x8 block (zero-AC shortcuts included, src/idct.rs:279-295,
// 344-353).  Used by the generic kernel and as the exact slow path of the fast kernel.
static __device__ __noinline__ void idct8x8_scalar_exact(const short* __restrict__ c, const unsigned* __restrict__ q, uint8_t* dst,
                                     unsigned stride) {
    unsigned temp[64];
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        bool acz = (c[i + 8] | c[i + 16] | c[i + 24] | c[i + 32] | c[i + 40] | c[i + 48] | c[i + 56]) == 0;
        if (acz) {
            unsigned dc = ((unsigned)(int)c[i] * q[i]) << 2;
#pragma unroll
            for (int k = 0; k < 8; k++) temp[i + 8 * k] = dc;
        } else {
            unsigned s[8];
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = (unsigned)(int)c[i + 8 * k] * q[i + 8 * k];
            unsigned x0, x1, x2, x3, t0, t1, t2, t3;
            IDCT_1D(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], 512u, x0, x1, x2, x3, t0, t1, t2, t3);
            temp[i] = (unsigned)sar(x0 + t3, 10);
            temp[i + 56] = (unsigned)sar(x0 - t3, 10);
            temp[i + 8] = (unsigned)sar(x1 + t2, 10);
            temp[i + 48] = (unsigned)sar(x1 - t2, 10);
            temp[i + 16] = (unsigned)sar(x2 + t1, 10);
            temp[i + 40] = (unsigned)sar(x2 - t1, 10);
            temp[i + 24] = (unsigned)sar(x3 + t0, 10);
            temp[i + 32] = (unsigned)sar(x3 - t0, 10);
        }
    }
    const unsigned XS = 65536u + (128u << 17);
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        const unsigned* s = temp + 8 * r;
        // the row shortcut (src/idct.rs:344-353) is algebraically identical to the general form
        unsigned x0, x1, x2, x3, t0, t1, t2, t3;
        IDCT_1D(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], XS, x0, x1, x2, x3, t0, t1, t2, t3);
        uint2 o;
        o.x = pack4_sat_u8(sar(x0 + t3, 17), sar(x1 + t2, 17), sar(x2 + t1, 17), sar(x3 + t0, 17));
        o.y = pack4_sat_u8(sar(x3 - t0, 17), sar(x2 - t1, 17), sar(x1 - t2, 17), sar(x0 - t3, 17));
        *reinterpret_cast<uint2*>(dst + (size_t)r * stride) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// Register-resident block: IDP.2A dequantisation (8-bit tables) and the two direct-form passes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned dp2a_lo_su(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned dp2a_hi_su(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

#define K1_DEQ8_ROW(k, B0, B1, B2, B3)                          \
    {                                                           \
        const unsigned bias_ = (k == 0) ? 0x80000u : 0u;        \
        s[k][0] = dp2a_lo_su(raw[k].x, (B0), bias_);            \
        s[k][1] = dp2a_hi_su(raw[k].x, (B0), bias_);            \
        s[k][2] = dp2a_lo_su(raw[k].y, (B1), bias_);            \
        s[k][3] = dp2a_hi_su(raw[k].y, (B1), bias_);            \
        s[k][4] = dp2a_lo_su(raw[k].z, (B2), bias_);            \
        s[k][5] = dp2a_hi_su(raw[k].z, (B2), bias_);            \
        s[k][6] = dp2a_lo_su(raw[k].w, (B3), bias_);            \
        s[k][7] = dp2a_hi_su(raw[k].w, (B3), bias_);            \
    }

template <int SLOT>
__device__ __forceinline__ void dequant_q8_const(const uint4 (&raw)[8], unsigned (&s)[8][8], const K1QCache& qc) {
#pragma unroll
    for (int k = 0; k < 8; k++)
        K1_DEQ8_ROW(k, qc.b[SLOT][4 * k + 0], qc.b[SLOT][4 * k + 1], qc.b[SLOT][4 * k + 2], qc.b[SLOT][4 * k + 3]);
}

// 8-bit table from memory (table index >= 4 of a batch): qp4 = the table's 32 packed operand words
__device__ __forceinline__ void dequant_q8_mem(const uint4 (&raw)[8], unsigned (&s)[8][8], const uint4* __restrict__ qp4) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint4 b = __ldg(qp4 + k);
        K1_DEQ8_ROW(k, b.x, b.y, b.z, b.w);
    }
}
// 16-bit table: sign-extend + IMAD; q4 = the table as u32[64]
__device__ __forceinline__ void dequant_q16_mem(const uint4 (&raw)[8], unsigned (&s)[8][8], const uint4* __restrict__ q4) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint4 qa = __ldg(q4 + 2 * k), qb = __ldg(q4 + 2 * k + 1);
        const unsigned bias = (k == 0) ? 0x80000u : 0u;
        s[k][0] = sext_lo(raw[k].x) * qa.x + bias;
        s[k][1] = sext_hi(raw[k].x) * qa.y + bias;
        s[k][2] = sext_lo(raw[k].y) * qa.z + bias;
        s[k][3] = sext_hi(raw[k].y) * qa.w + bias;
        s[k][4] = sext_lo(raw[k].z) * qb.x + bias;
        s[k][5] = sext_hi(raw[k].z) * qb.y + bias;
        s[k][6] = sext_lo(raw[k].w) * qb.z + bias;
        s[k][7] = sext_hi(raw[k].w) * qb.w + bias;
    }
}
// Dequantises raw (the block as 8 x 16 bytes, natural order) with the table selected by qflags (DevComp::qflags):
// every s[0][i] carries the +2^19 bias that the column pass cancels (0x80000000 in its even-part constant).
// Returns non-zero when some |c*q| of the first row reaches 2^19: the caller must then take idct8x8_scalar_exact
// (only there can the reference's zero-AC column shortcut, src/idct.rs:279-295, differ from the butterfly).
__device__ __forceinline__ unsigned dequant_block(const uint4 (&raw)[8], unsigned (&s)[8][8], unsigned qflags, const K1QCache& qc,
                                                  const uint4* __restrict__ q4, const uint4* __restrict__ qp4) {
    const unsigned qslot = qflags >> 8;
    if (qflags & 1u) {
        if (qslot == 0) dequant_q8_const<0>(raw, s, qc);
        else if (qslot == 1) dequant_q8_const<1>(raw, s, qc);
        else if (qslot == 2) dequant_q8_const<2>(raw, s, qc);
        else if (qslot == 3) dequant_q8_const<3>(raw, s, qc);
        else dequant_q8_mem(raw, s, qp4);
    } else {
        dequant_q16_mem(raw, s, q4);
    }
    return (s[0][0] | s[0][1] | s[0][2] | s[0][3] | s[0][4] | s[0][5] | s[0][6] | s[0][7]) >> 20;
}
// Both passes in direct form on a dequantised (biased) block; rows[r] = the 8 output samples of row r.
__device__ __forceinline__ void idct8x8_direct(unsigned (&s)[8][8], uint2 (&rows)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        unsigned o0, o1, o2, o3, o4, o5, o6, o7;
        IDCT_1D_DIRECT(s[0][i], s[1][i], s[2][i], s[3][i], s[4][i], s[5][i], s[6][i], s[7][i], (512u + 0x80000000u), o0, o1, o2, o3,
                       o4, o5, o6, o7);
        s[0][i] = (unsigned)sar(o0, 10); s[1][i] = (unsigned)sar(o1, 10); s[2][i] = (unsigned)sar(o2, 10); s[3][i] = (unsigned)sar(o3, 10);
        s[4][i] = (unsigned)sar(o4, 10); s[5][i] = (unsigned)sar(o5, 10); s[6][i] = (unsigned)sar(o6, 10); s[7][i] = (unsigned)sar(o7, 10);
    }
    const unsigned XSD = 65536u + (128u << 17);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        unsigned o0, o1, o2, o3, o4, o5, o6, o7;
        IDCT_1D_DIRECT(s[r][0], s[r][1], s[r][2], s[r][3], s[r][4], s[r][5], s[r][6], s[r][7], XSD, o0, o1, o2, o3, o4, o5, o6, o7);
        rows[r].x = pack4_sat_u8(sar(o0, 17), sar(o1, 17), sar(o2, 17), sar(o3, 17));
        rows[r].y = pack4_sat_u8(sar(o4, 17), sar(o5, 17), sar(o6, 17), sar(o7, 17));
    }
}

// ---------------------------------------------------------------------------------------------
// Scaled IDCT (Decoder::scale, src/idct.rs:456-565) on a block held as eight 16-byte rows: only the low-frequency
// corner is read, everything stays in registers.  coefficient (r, c) = half (c & 1) of word c / 2 of raw[r].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned raw_coef(const uint4& row, int c) {
    const unsigned w = c < 2 ? row.x : (c < 4 ? row.y : (c < 6 ? row.z : row.w));
    return (c & 1) ? sext_hi(w) : sext_lo(w);
}
// src/idct.rs:456-517; q = the table as u32[64]
__device__ __forceinline__ void idct4x4_regs(const uint4 (&raw)[8], const unsigned* __restrict__ q, uint8_t* dst, unsigned stride) {
    unsigned temp[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const unsigned s0 = raw_coef(raw[0], i) * __ldg(q + i), s1 = raw_coef(raw[1], i) * __ldg(q + i + 8);
        const unsigned s2 = raw_coef(raw[2], i) * __ldg(q + i + 16), s3 = raw_coef(raw[3], i) * __ldg(q + i + 24);
        const unsigned x0 = (s0 + s2) << 2, x2 = (s0 - s2) << 2;
        const unsigned p1 = (s1 + s3) * F2F_0_5411961;
        const unsigned t0 = (unsigned)sar(p1 + s3 * F2F_N1_847759065 + 512u, 10);
        const unsigned t2 = (unsigned)sar(p1 + s1 * F2F_0_765366865 + 512u, 10);
        temp[0][i] = x0 + t2;
        temp[3][i] = x0 - t2;
        temp[1][i] = x2 + t0;
        temp[2][i] = x2 - t0;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const unsigned s0 = temp[i][0], s1 = temp[i][1], s2 = temp[i][2], s3 = temp[i][3];
        const unsigned x0 = ((s0 + s2) << 12) + (1u << 16) + (128u << 17);
        const unsigned x2 = ((s0 - s2) << 12) + (1u << 16) + (128u << 17);
        const unsigned p1 = (s1 + s3) * F2F_0_5411961;
        const unsigned t0 = p1 + s3 * F2F_N1_847759065;
        const unsigned t2 = p1 + s1 * F2F_0_765366865;
        *reinterpret_cast<unsigned*>(dst + (size_t)i * stride) = pack4_sat_u8(sar(x0 + t2, 17), sar(x2 + t0, 17), sar(x2 - t0, 17), sar(x0 - t2, 17));
    }
}
// src/idct.rs:519-553
__device__ __forceinline__ void idct2x2_regs(const uint4 (&raw)[8], const unsigned* __restrict__ q, uint8_t* dst, unsigned stride) {
    const unsigned s00 = raw_coef(raw[0], 0) * __ldg(q), s10 = raw_coef(raw[1], 0) * __ldg(q + 8);
    const unsigned s01 = raw_coef(raw[0], 1) * __ldg(q + 1), s11 = raw_coef(raw[1], 1) * __ldg(q + 9);
    const unsigned x0 = s00 + s10 + 4u + (128u << 3), x2 = s00 - s10 + 4u + (128u << 3);
    const unsigned x1 = s01 + s11, x3 = s01 - s11;
    *reinterpret_cast<unsigned short*>(dst) = (unsigned short)pack4_sat_u8(sar(x0 + x1, 3), sar(x0 - x1, 3), 0, 0);
    *reinterpret_cast<unsigned short*>(dst + stride) = (unsigned short)pack4_sat_u8(sar(x2 + x3, 3), sar(x2 - x3, 3), 0, 0);
}
// src/idct.rs:555-565 (Wrapping<i32> division truncates toward zero)
__device__ __forceinline__ void idct1x1_regs(const uint4 (&raw)[8], const unsigned* __restrict__ q, uint8_t* dst) {
    const int s0 = (int)(raw_coef(raw[0], 0) * __ldg(q) + 1024u) / 8;
    dst[0] = (uint8_t)min(max(s0, 0), 255);
}

// ---------------------------------------------------------------------------------------------
// SSSE3 arithmetic in registers (src/arch/ssse3.rs:8-84, 124-192): int16 lanes carried in int32 registers.
//   adds / subs  saturate: VIADDMNMX + VIMNMX
//   mulhrs(a, c) = (((a * c) >> 14) + 1) >> 1 == (a * c + 2^14) >> 15 (floor identities); every multiplier is a
//                  constant with |c| < 2^15, so the result always fits int16 and needs no truncation: IMAD + SHF
// ---------------------------------------------------------------------------------------------
// (spelled in PTX: left to itself the compiler recognises a saturating i16 add, narrows the whole data flow to 16-bit
// types and emulates them with compare / select / PRMT sequences -- 3x the instructions)
__device__ __forceinline__ int s3_adds(int a, int b) {
    int r;
    asm("{\n.reg .s32 t;\nadd.s32 t, %1, %2;\nmax.s32 t, t, -32768;\nmin.s32 %0, t, 32767;\n}" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ int s3_subs(int a, int b) {
    int r;
    asm("{\n.reg .s32 t;\nsub.s32 t, %1, %2;\nmax.s32 t, t, -32768;\nmin.s32 %0, t, 32767;\n}" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ int s3_mulhrs(int a, int c) { return (a * c + 16384) >> 15; }

// idct8, src/arch/ssse3.rs:8-84, on one lane (eight registers, in place)
__device__ __forceinline__ void s3_idct8(int& d0, int& d1, int& d2, int& d3, int& d4, int& d5, int& d6, int& d7) {
    int p2 = d2, p3 = d6;
    int p1 = s3_mulhrs(s3_adds(p2, p3), 17734);
    int t2 = s3_subs(s3_subs(p1, p3), s3_mulhrs(p3, 27779));
    int t3 = s3_adds(p1, s3_mulhrs(p2, 25079));
    p2 = d0;
    p3 = d4;
    int t0 = s3_adds(p2, p3), t1 = s3_subs(p2, p3);
    const int x0 = s3_adds(t0, t3), x3 = s3_subs(t0, t3), x1 = s3_adds(t1, t2), x2 = s3_subs(t1, t2);
    t0 = d7;
    t1 = d5;
    t2 = d3;
    t3 = d1;
    p3 = s3_adds(t0, t2);
    int p4 = s3_adds(t1, t3);
    p1 = s3_adds(t0, t3);
    p2 = s3_adds(t1, t2);
    int p5 = s3_adds(p3, p4);
    p5 = s3_adds(p5, s3_mulhrs(p5, 5763));
    t0 = s3_mulhrs(t0, 9786);
    t1 = s3_adds(s3_adds(t1, t1), s3_mulhrs(t1, 1741));
    t2 = s3_adds(s3_adds(t2, s3_adds(t2, t2)), s3_mulhrs(t2, 2383));
    t3 = s3_adds(t3, s3_mulhrs(t3, 16427));
    p1 = s3_subs(p5, s3_mulhrs(p1, 29490));
    p2 = s3_subs(s3_subs(s3_subs(p5, p2), p2), s3_mulhrs(p2, 18446));
    p3 = s3_subs(s3_mulhrs(p3, -31509), p3);
    p4 = s3_mulhrs(p4, -12785);
    t3 = s3_adds(s3_adds(p1, p4), t3);
    t2 = s3_adds(s3_adds(p2, p3), t2);
    t1 = s3_adds(s3_adds(p2, p4), t1);
    t0 = s3_adds(s3_adds(p1, p3), t0);
    d0 = s3_adds(x0, t3);
    d7 = s3_subs(x0, t3);
    d1 = s3_adds(x1, t2);
    d6 = s3_subs(x1, t2);
    d2 = s3_adds(x2, t1);
    d5 = s3_subs(x2, t1);
    d3 = s3_adds(x3, t0);
    d4 = s3_subs(x3, t0);
}

// dequantize_and_idct_block_8x8, src/arch/ssse3.rs:124-192, whole block in registers.  q4 = the table as u32[64].
// _mm_mullo_epi16 and _mm_slli_epi16(.., 3) both wrap at 16 bits: (c * q * 8) mod 2^16, read as int16.
__device__ __forceinline__ void idct8x8_ssse3_regs(const uint4 (&raw)[8], const uint4* __restrict__ q4, uint2 (&rows)[8]) {
    int d[8][8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint4 qa = __ldg(q4 + 2 * k), qb = __ldg(q4 + 2 * k + 1);
        const unsigned w[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
        const unsigned q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // only the low 16 bits of the product matter, so the low half needs no extraction
            d[k][2 * j] = (int)(short)(unsigned short)(w[j] * (q[2 * j] << 3));
            d[k][2 * j + 1] = (int)(short)(unsigned short)((w[j] >> 16) * (q[2 * j + 1] << 3));
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) s3_idct8(d[0][k], d[1][k], d[2][k], d[3][k], d[4][k], d[5][k], d[6][k], d[7][k]);  // down the columns (ssse3.rs:162)
#pragma unroll
    for (int r = 0; r < 8; r++) {
        s3_idct8(d[r][0], d[r][1], d[r][2], d[r][3], d[r][4], d[r][5], d[r][6], d[r][7]);  // transpose-idct8-transpose = along the rows
        // adds(.., OFFSET + ROUNDING_BIAS) >> 6, packus (ssse3.rs:173-185)
        rows[r].x = pack4_sat_u8(s3_adds(d[r][0], 8224) >> 6, s3_adds(d[r][1], 8224) >> 6, s3_adds(d[r][2], 8224) >> 6, s3_adds(d[r][3], 8224) >> 6);
        rows[r].y = pack4_sat_u8(s3_adds(d[r][4], 8224) >> 6, s3_adds(d[r][5], 8224) >> 6, s3_adds(d[r][6], 8224) >> 6, s3_adds(d[r][7], 8224) >> 6);
    }
}

}  // namespace b200jpg
