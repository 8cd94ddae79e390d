// color_core.cuh -- chroma upsampling + YCbCr -> RGB arithmetic shared by K2 (k2_color.cu) and the fused kernel
// (kf_fused.cu).  Reference: UpsamplerH2V2 (src/upsampler.rs:191-228), color_convert_line_ycbcr / ycbcr_to_rgb
// (src/decoder.rs:1406-1437, 1486-1508).  Arithmetic specification: SURVEY.md Appendix A.3 / A.4; the derivation of
// the clamped-edge triangle filter is at the top of k2_color.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace b200jpg {

// src/decoder.rs:1486-1508.  stbi_f2f(x) = (x * 2^20 + 0.5) as i32 evaluated in f32:
// 1.40200 -> 1470104, 0.34414 -> 360857, 0.71414 -> 748830, 1.77200 -> 1858077
// (checked against the oracle's f32 evaluation in tests/test_oracle_kat.py).
#define C_R_CR 1470104
#define C_G_CB 360857
#define C_G_CR 748830
#define C_B_CB 1858077
#define YCC_HALF (1 << 19)

// The same value from y << 16 and the CENTRED chroma samples cb - 128, cr - 128: literally the reference's form
// (src/decoder.rs:1490-1498: y * 2^20 + HALF, cb - 128, cr - 128).  The fast kernels get the centred samples for
// free -- the H2V2 filter subtracts 128 * 16 in its IDP.4A accumulator, the 4:4:4 path sign-extends the byte of
// (word ^ 0x80808080) inside the PRMT that extracts it -- so a pixel costs 5 IMADs instead of 7.
// `sixteen` is 16 passed through a register the compiler cannot see through, so that y16 * 16 + HALF stays one
// IMAD (FMA pipe) instead of a shift and an add on the ALU pipe, which is the busier one in these kernels.
struct YccRegs {
    int mul;
};
__device__ __forceinline__ void ycbcr_scalar_y16(int y16, int cbm, int crm, int& r, int& g, int& b, const YccRegs& k) {
    const int base = y16 * k.mul + YCC_HALF;
    r = (base + C_R_CR * crm) >> 20;
    g = (base - C_G_CB * cbm - C_G_CR * crm) >> 20;
    b = (base + C_B_CB * cbm) >> 20;
}

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

constexpr unsigned H2V2_BIAS = 8u - 128u * 16u;
struct Chroma16 {  // upsampled chroma MINUS 128 for 16 pixels of the two rows of the pair
    int odd[16];   // output row 2p-1 (near = chroma row p-1)
    int even[16];  // output row 2p   (near = chroma row p)
};

// a*: chroma row A = max(p-1,0); b*: chroma row B = min(p, in_h-1).  lo/hi = samples i0..i0+3 / i0+4..i0+7,
// L / R = clamped halo samples i0-1 / i0+8.
__device__ __forceinline__ void h2v2_16(unsigned a_lo, unsigned a_hi, unsigned aL, unsigned aR, unsigned b_lo,
                                        unsigned b_hi, unsigned bL, unsigned bR, Chroma16& o) {
    // shifted words: s0 = (L, 0, 1, 2), s1 = (3, 4, 5, 6), s2 = (7, R, -, -)
    const unsigned as0 = prmt(aL, a_lo, 0x6540), as1 = prmt(a_lo, a_hi, 0x6543), as2 = prmt(a_hi, aR, 0x0043);
    const unsigned bs0 = prmt(bL, b_lo, 0x6540), bs1 = prmt(b_lo, b_hi, 0x6543), bs2 = prmt(b_hi, bR, 0x0043);
    // P[j] = (a[i-1], a[i], b[i-1], b[i]) for i = i0 + j
    unsigned P[9];
    P[0] = prmt(as0, bs0, 0x5410);
    P[1] = prmt(a_lo, b_lo, 0x5410);
    P[2] = prmt(as0, bs0, 0x7632);
    P[3] = prmt(a_lo, b_lo, 0x7632);
    P[4] = prmt(as1, bs1, 0x5410);
    P[5] = prmt(a_hi, b_hi, 0x5410);
    P[6] = prmt(as1, bs1, 0x7632);
    P[7] = prmt(a_hi, b_hi, 0x7632);
    P[8] = prmt(as2, bs2, 0x5410);
    // weights on (a[i-1], a[i], b[i-1], b[i]); t = 3 near + far
    const unsigned W_A_EVEN = 0x03010903u;  // near = A: out[2i]   = 3 t[i] + t[i-1]
    const unsigned W_A_ODD = 0x01030309u;   // near = A: out[2i-1] = 3 t[i-1] + t[i]
    const unsigned W_B_EVEN = 0x09030301u;  // near = B: out[2i]
    const unsigned W_B_ODD = 0x03090103u;   // near = B: out[2i-1]
#pragma unroll
    for (int j = 0; j < 8; j++) {
        // (sum + 8 - 128 * 16) >> 4 == ((sum + 8) >> 4) - 128 exactly (arithmetic shift)
        o.odd[2 * j] = (int)__dp4a(P[j], W_A_EVEN, H2V2_BIAS) >> 4;
        o.odd[2 * j + 1] = (int)__dp4a(P[j + 1], W_A_ODD, H2V2_BIAS) >> 4;
        o.even[2 * j] = (int)__dp4a(P[j], W_B_EVEN, H2V2_BIAS) >> 4;
        o.even[2 * j + 1] = (int)__dp4a(P[j + 1], W_B_ODD, H2V2_BIAS) >> 4;
    }
}

// H2V1 (4:2:2, src/upsampler.rs:134-163) for 16 output pixels: out[2i] = (3 a[i] + a[i-1] + 2) >> 2,
// out[2i+1] = (3 a[i] + a[i+1] + 2) >> 2 with clamped ends -- the horizontal half of the triangle filter above, one
// IDP.4A per sample on byte pairs (a[i-1], a[i]); results CENTRED (minus 128), like h2v2_16.
// lo / hi = samples i0..i0+3 / i0+4..i0+7, L / R = the clamped halo samples i0-1 / i0+8.
__device__ __forceinline__ void h2v1_16(unsigned lo, unsigned hi, unsigned L, unsigned R, int (&o)[16]) {
    // pair j = (a[i0+j-1], a[i0+j]) sits in word w[j] at bytes pos[j], pos[j]+1: only the pairs that straddle a word
    // need a PRMT, the others are read in place by moving the weights to their byte lanes
    const unsigned s0 = prmt(L, lo, 0x0040), s1 = prmt(lo, hi, 0x0043), s2 = prmt(hi, R, 0x0043);
    const unsigned w[9] = {s0, lo, lo, lo, s1, hi, hi, hi, s2};
    constexpr unsigned pos[9] = {0, 0, 1, 2, 0, 0, 1, 2, 0};
    constexpr unsigned W_EVEN = 0x0301u, W_ODD = 0x0103u;  // out[2i] = a[i-1] + 3 a[i]; out[2i-1] = 3 a[i-1] + a[i]
    constexpr unsigned BIAS = 2u - 128u * 4u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        o[2 * j] = (int)__dp4a(w[j], W_EVEN << (8u * pos[j]), BIAS) >> 2;
        o[2 * j + 1] = (int)__dp4a(w[j + 1], W_ODD << (8u * pos[j + 1]), BIAS) >> 2;
    }
}

// H1V2 (4:4:0, src/upsampler.rs:165-189) for 16 output pixels of one row: out[x] = (3 near[x] + far[x] + 2) >> 2, the vertical
// half of the triangle filter -- byte pairs (near, far) side by side, one IDP.4A per sample; results CENTRED (minus 128).
// nv / fv = 16 samples of the near and the far chroma row.
__device__ __forceinline__ void h1v2_16(const uint4 nv, const uint4 fv, int (&o)[16]) {
    const unsigned n[4] = {nv.x, nv.y, nv.z, nv.w}, f[4] = {fv.x, fv.y, fv.z, fv.w};
    constexpr unsigned BIAS = 2u - 128u * 4u;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const unsigned p01 = prmt(n[w], f[w], 0x5140), p23 = prmt(n[w], f[w], 0x7362);  // (n0 f0 n1 f1), (n2 f2 n3 f3)
        o[4 * w + 0] = (int)__dp4a(p01, 0x00000103u, BIAS) >> 2;
        o[4 * w + 1] = (int)__dp4a(p01, 0x01030000u, BIAS) >> 2;
        o[4 * w + 2] = (int)__dp4a(p23, 0x00000103u, BIAS) >> 2;
        o[4 * w + 3] = (int)__dp4a(p23, 0x01030000u, BIAS) >> 2;
    }
}

__device__ __forceinline__ YccRegs make_ycc_regs(int3 sixteen, bool opaque) {
    YccRegs k;
    k.mul = sixteen.x;
    if (opaque) asm volatile("mov.u32 %0, %1;" : "=r"(k.mul) : "r"(sixteen.x));
    return k;
}

// Stores the first `nbytes` (<= 4*NW) bytes of ow[] at dst: 128-bit stores when dst is 16-byte aligned (always
// the case when width % 16 == 0), 32-bit stores when 4-byte aligned, byte stores otherwise / for a ragged tail.
template <int NW>
__device__ __forceinline__ void store_words(uint8_t* dst, const unsigned (&ow)[NW], unsigned nbytes) {
    const unsigned a = (unsigned)(uintptr_t)dst;
    if (nbytes == 4u * NW && (a & 15u) == 0) {
#pragma unroll
        for (int k = 0; k < NW / 4; k++)
            reinterpret_cast<uint4*>(dst)[k] = make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]);
    } else if (nbytes == 4u * NW && (a & 3u) == 0) {
#pragma unroll
        for (int k = 0; k < NW; k++) reinterpret_cast<unsigned*>(dst)[k] = ow[k];
    } else {
#pragma unroll
        for (int k = 0; k < 4 * NW; k++)
            if ((unsigned)k < nbytes) dst[k] = (uint8_t)(ow[k >> 2] >> (8 * (k & 3)));
    }
}

// The SSSE3 colour path (src/arch/ssse3.rs:208-244) from a luma byte and CENTRED chroma.  Its saturating adds can never
// saturate for 8-bit inputs -- |cb6|, |cr6| <= 8192, so |cr_140200| <= 11486, |cb_177200| <= 14517,
// |cb_034414 + cr_071414| <= 8671 and 32 <= y6 <= 16352: every sum stays inside int16 -- so plain adds are exact
// (checked exhaustively over all 2^24 inputs in tests/test_oracle_kat.py).  mulhrs(a, c) == (a * c + 2^14) >> 15.
__device__ __forceinline__ void ycbcr_ssse3_centred(int y, int cbm, int crm, int& r, int& g, int& b) {
    const int y6 = y * 64 + 32, cb6 = cbm * 64, cr6 = crm * 64;
    const int cr_140200 = ((cr6 * 13173 + 16384) >> 15) + cr6;
    const int cb_034414 = (cb6 * 11276 + 16384) >> 15;
    const int cr_071414 = (cr6 * 23401 + 16384) >> 15;
    const int cb_177200 = ((cb6 * 25297 + 16384) >> 15) + cb6;
    r = (y6 + cr_140200) >> 6;
    g = (y6 - (cb_034414 + cr_071414)) >> 6;
    b = (y6 + cb_177200) >> 6;
}

// 16 pixels: luma bytes in yv (4 words), CENTRED chroma (cb - 128, cr - 128) -> 48 output bytes at dst.
// npx = number of valid pixels of this 16-pixel group (16 except for the last group of a ragged row)
// SSSE3 = the x86 build's arithmetic: the first nss pixels of the group (0, 8 or 16: the SIMD loop covers
// (W / 8 - 1) * 8 pixels of a row, src/arch/ssse3.rs:206) take the SSSE3 formula, the rest the scalar one.
template <bool SSSE3 = false>
__device__ __forceinline__ void ycbcr_store16(const uint4 yv, const int* cb, const int* cr, uint8_t* dst, const YccRegs& sixteen,
                                              unsigned npx, unsigned nss = 0) {
    const unsigned yw[4] = {yv.x, yv.y, yv.z, yv.w};
    unsigned ow[12];
#pragma unroll
    for (int w = 0; w < 4; w++) {
        int r[4], g[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (SSSE3 && 4u * (unsigned)w < nss) {
                ycbcr_ssse3_centred((int)prmt(yw[w], 0u, 0x4440u | (unsigned)k), cb[4 * w + k], cr[4 * w + k], r[k], g[k], b[k]);
                continue;
            }
            // one PRMT puts luma byte k at bits 16..23 (y << 16); the << 4 folds into the IMADs below
            const int y16 = (int)prmt(yw[w], 0u, 0x4044u | ((unsigned)k << 8));
            ycbcr_scalar_y16(y16, cb[4 * w + k], cr[4 * w + k], r[k], g[k], b[k], sixteen);
        }
        // bytes: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        ow[3 * w + 0] = pack_sat_u8(g[0], r[0], pack_sat_u8(r[1], b[0], 0u));
        ow[3 * w + 1] = pack_sat_u8(b[1], g[1], pack_sat_u8(g[2], r[2], 0u));
        ow[3 * w + 2] = pack_sat_u8(r[3], b[2], pack_sat_u8(b[3], g[3], 0u));
    }
    store_words<12>(dst, ow, 3u * npx);
}

// selector that keeps bytes 0..m-1 of a word and replicates byte m-1 into the rest (m = 1..4)
__device__ __forceinline__ unsigned keep_sel(unsigned m) {
    return (0x3210u & ((1u << (4u * m)) - 1u)) | ((((m - 1u) * 0x1111u) << (4u * m)) & 0xffffu);
}

// bytes nvalid..7 of the 8-byte window := byte nvalid-1 (nvalid = 1..7): two PRMTs with computed selectors
__device__ __forceinline__ uint2 replicate_last_sample(uint2 v, unsigned nvalid) {
    if (nvalid <= 4u) {
        v.x = prmt(v.x, 0u, keep_sel(nvalid));
        v.y = prmt(v.x, 0u, (nvalid - 1u) * 0x1111u);
    } else {
        v.y = prmt(v.y, 0u, keep_sel(nvalid - 4u));
    }
    return v;
}

}  // namespace b200jpg
