// sbs.h -- "sparse block stream": the wire format between the host entropy decoder and the device.
//
// Dense coefficients (the reference's Worker::append_row payload, src/decoder.rs:962-983) are 128 B per
// 8x8 block and ~75-95 % zeros; shipping them over PCIe costs 3 B per 4:2:0 pixel.  The host therefore
// sends what Huffman decoding actually produced and kernel K0 (k0_expand.cu) rebuilds the dense slab in
// HBM, bit for bit, before K1 runs.
//
// One stream per image, blocks in *scan order* (the order the entropy decoder meets them):
//   bm  : u64[nb_pad]   bit k (1..63) = the AC coefficient with zig-zag index k is non-zero;
//                       bit 0 = "wide": this block's AC values are int16 (else int8)
//   dc  : i16[nb_pad]   coefficient 0 of every block
//   voff: u32[nb_pad/32 + 1]  byte offset into `vals` of every group of 32 blocks
//   vals: bytes         per block, its non-zero AC values in zig-zag order, 1 or 2 bytes each
// nb_pad = nb rounded up to 32; padding blocks are all-zero.  Scan order is either
//   SBS_PLANAR      component 0's blocks in raster order, then component 1's, ...
//   SBS_INTERLEAVED MCU by MCU, inside an MCU component by component, v then h (src/decoder.rs:978-983)
// optionally or-ed with
//   SBS_NATURAL     bit k of bm / the order of a block's values refer to NATURAL coefficient positions instead of
//                   zig-zag indices (what compacting an already dense block produces without a permutation)
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if !defined(__CUDACC__)
#include <emmintrin.h>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define B200JPG_SBS_AVX512 1
#endif
#endif

namespace b200jpg {

enum : unsigned {
    SBS_PLANAR = 0,
    SBS_INTERLEAVED = 1,
    SBS_NATURAL = 2,
    SBS_ENTROPY = 4,        // internal: see sbs_pipeline.h
    SBS_BLOCK_OFFSETS = 8,  // internal (device-made streams, entropy_dev.h): u32 value offset per block, relative to the stream start
};

struct SbsLayout {
    size_t nb = 0, nb_pad = 0, off_dc = 0, off_voff = 0, off_vals = 0;
    static SbsLayout make(size_t nblocks) {
        SbsLayout l;
        l.nb = nblocks;
        l.nb_pad = (nblocks + 31) / 32 * 32;
        l.off_dc = 8 * l.nb_pad;
        l.off_voff = 10 * l.nb_pad;
        l.off_vals = (l.off_voff + 4 * (l.nb_pad / 32 + 1) + 15) / 16 * 16;
        return l;
    }
    // upper bound of a stream's length (every AC coefficient non-zero and wide) plus store slack
    size_t worst_bytes() const { return off_vals + nb * 126 + 256; }
};

// Per-image descriptor of K0.  Plain data, shared by host and device.
struct alignas(16) K0Image {
    unsigned long long stream_off;  // byte offset of the stream inside the device stream buffer (16 B aligned)
    unsigned nb;                    // blocks in the stream
    unsigned order;                 // SBS_*
    unsigned mcu_w;                 // MCUs per MCU row (interleaved)
    unsigned bpm;                   // blocks per MCU (interleaved)
    unsigned slab_row[4];           // first row (128 B units) of each component inside the coefficient slab
    unsigned block_w[4];            // blocks per block row of each component
    unsigned first[4];              // planar: scan-order index of each component's first block
    unsigned char h[4], v[4];       // interleaved: blocks per MCU of each component
    unsigned char mcu_comp[12], mcu_hx[12], mcu_vy[12];  // interleaved: component / position of MCU block j
    unsigned pad_[3];
};

#if !defined(__CUDACC__)
// zig-zag index of natural position p (inverse of UNZIGZAG, src/decoder.rs:27-36)
static const uint8_t SBS_ZIGZAG_OF[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42,
                                          3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
                                          10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
                                          21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

// Sequential writer of one stream into caller-provided memory of at least worst_bytes().
class SbsWriter {
public:
    void begin(uint8_t* base, size_t nblocks) {
        lay_ = SbsLayout::make(nblocks);
        base_ = base;
        bm_ = (uint64_t*)base;
        dc_ = (int16_t*)(base + lay_.off_dc);
        voff_ = (uint32_t*)(base + lay_.off_voff);
        vals_ = base + lay_.off_vals;
        t_ = 0;
        vpos_ = 0;
    }
    bool active() const { return base_ != nullptr; }
    size_t blocks_written() const { return t_; }
    const SbsLayout& layout() const { return lay_; }
    void restart() { t_ = 0; vpos_ = 0; }

    // one block: AC values v[0..n) in zig-zag order, their positions in `bits` (bit k = zig-zag index k),
    // `wide` = some value does not fit int8.  v must be readable up to v[(n+15)/16*16).
    inline void put(uint64_t bits, int16_t dc, const int16_t* v, unsigned n, bool wide) {
        if ((t_ & 31) == 0) voff_[t_ >> 5] = (uint32_t)vpos_;
        bm_[t_] = (bits & ~(uint64_t)1) | (wide ? 1u : 0u);
        dc_[t_] = dc;
        t_++;
        uint8_t* o = vals_ + vpos_;
        if (!wide) {
            for (unsigned i = 0; i < n; i += 16) {
                const __m128i a = _mm_loadu_si128((const __m128i*)(v + i)), b = _mm_loadu_si128((const __m128i*)(v + i + 8));
                _mm_storeu_si128((__m128i*)(o + i), _mm_packs_epi16(a, b));
            }
            vpos_ += n;
        } else {
            for (unsigned i = 0; i < n; i += 8) _mm_storeu_si128((__m128i*)(o + 2 * i), _mm_loadu_si128((const __m128i*)(v + i)));
            vpos_ += 2 * (size_t)n;
        }
    }
    inline void put_zero() {
        if ((t_ & 31) == 0) voff_[t_ >> 5] = (uint32_t)vpos_;
        bm_[t_] = 0;
        dc_[t_] = 0;
        t_++;
    }
    // one block from dense coefficients in natural order
    inline void put_dense(const int16_t* c) {
        const __m128i z = _mm_setzero_si128();
        uint64_t nz = 0;
        for (int r = 0; r < 4; r++) {
            const __m128i a = _mm_loadu_si128((const __m128i*)(c + 16 * r)), b = _mm_loadu_si128((const __m128i*)(c + 16 * r + 8));
            const unsigned m = (unsigned)_mm_movemask_epi8(_mm_packs_epi16(_mm_cmpeq_epi16(a, z), _mm_cmpeq_epi16(b, z)));
            nz |= (uint64_t)(~m & 0xffffu) << (16 * r);
        }
        nz &= ~(uint64_t)1;
        if (!nz) {
            if ((t_ & 31) == 0) voff_[t_ >> 5] = (uint32_t)vpos_;
            bm_[t_] = 0;
            dc_[t_] = c[0];
            t_++;
            return;
        }
        alignas(16) int16_t byk[64 + 16];
        uint64_t zm = 0;
        unsigned acc = 0;
        for (uint64_t m = nz; m; m &= m - 1) {
            const unsigned p = (unsigned)__builtin_ctzll(m), k = SBS_ZIGZAG_OF[p];
            byk[k] = c[p];
            zm |= (uint64_t)1 << k;
            acc |= (unsigned)(c[p] + 128);
        }
        alignas(16) int16_t v[64 + 16];
        unsigned n = 0;
        for (uint64_t m = zm; m; m &= m - 1) v[n++] = byk[__builtin_ctzll(m)];
        put(zm, c[0], v, n, acc > 255u);
    }
    // one block from dense coefficients in natural order, for SBS_NATURAL streams: bitmap and values stay in
    // natural order, so this is one SIMD zero test and one pass over the set bits
    inline void put_dense_natural(const int16_t* c) {
        const __m128i z = _mm_setzero_si128(), k128 = _mm_set1_epi16(128);
        uint64_t nz = 0;
        __m128i hi = z;
        for (int r = 0; r < 4; r++) {
            __m128i a = _mm_loadu_si128((const __m128i*)(c + 16 * r));
            const __m128i b = _mm_loadu_si128((const __m128i*)(c + 16 * r + 8));
            const unsigned m = (unsigned)_mm_movemask_epi8(_mm_packs_epi16(_mm_cmpeq_epi16(a, z), _mm_cmpeq_epi16(b, z)));
            nz |= (uint64_t)(~m & 0xffffu) << (16 * r);
            if (r == 0) a = _mm_and_si128(a, _mm_set_epi16(-1, -1, -1, -1, -1, -1, -1, 0));  // DC travels separately
            hi = _mm_or_si128(hi, _mm_or_si128(_mm_srli_epi16(_mm_add_epi16(a, k128), 8), _mm_srli_epi16(_mm_add_epi16(b, k128), 8)));
        }
        nz &= ~(uint64_t)1;
        if ((t_ & 31) == 0) voff_[t_ >> 5] = (uint32_t)vpos_;
        const bool wide = _mm_movemask_epi8(_mm_cmpeq_epi16(hi, z)) != 0xffff;  // some (v + 128) >> 8 != 0
        bm_[t_] = nz | (wide ? 1u : 0u);
        dc_[t_] = c[0];
        t_++;
        uint8_t* o = vals_ + vpos_;
        if (!wide) {
            for (uint64_t m = nz; m; m &= m - 1) *o++ = (uint8_t)c[__builtin_ctzll(m)];
        } else {
            for (uint64_t m = nz; m; m &= m - 1) {
                const uint16_t v = (uint16_t)c[__builtin_ctzll(m)];
                *o++ = (uint8_t)v;
                *o++ = (uint8_t)(v >> 8);
            }
        }
        vpos_ = (size_t)(o - vals_);
    }
    // `count` consecutive dense blocks (natural order) -> SBS_NATURAL blocks; picks the AVX-512 VBMI2 body when the
    // CPU has it (one vpcompressb per block instead of a loop over the set bits)
    void put_dense_natural_run(const int16_t* c, size_t count) {
#ifdef B200JPG_SBS_AVX512
        static const bool vbmi2 = __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("avx512bw") &&
                                  !getenv("B200JPG_NO_AVX512");  // (the variable exists for the tests of the SSE2 body)
        if (vbmi2) {
            put_dense_natural_run_avx512(c, count);
            return;
        }
#endif
        for (size_t b = 0; b < count; b++) put_dense_natural(c + 64 * b);
    }
    // pads the tables to nb_pad and returns the stream length (multiple of 16)
    size_t finish() {
        while (t_ < lay_.nb_pad) put_zero();
        voff_[lay_.nb_pad >> 5] = (uint32_t)vpos_;
        const size_t len = (lay_.off_vals + vpos_ + 15) / 16 * 16;
        memset(vals_ + vpos_, 0, len - (lay_.off_vals + vpos_));
        return len;
    }

private:
#ifdef B200JPG_SBS_AVX512
    __attribute__((target("avx512f,avx512bw,avx512vbmi2"))) void put_dense_natural_run_avx512(const int16_t* c, size_t count) {
        for (size_t b = 0; b < count; b++, c += 64) {
            const __m512i lo = _mm512_loadu_si512(c), hi = _mm512_loadu_si512(c + 32);
            const uint64_t nz = (((uint64_t)_mm512_test_epi16_mask(hi, hi) << 32) | (uint64_t)_mm512_test_epi16_mask(lo, lo)) & ~(uint64_t)1;
            // saturating narrow + widen back: differs exactly where a value is outside int8
            const __m256i nlo = _mm512_cvtsepi16_epi8(lo), nhi = _mm512_cvtsepi16_epi8(hi);
            const uint64_t out8 = (((uint64_t)_mm512_cmpneq_epi16_mask(_mm512_cvtepi8_epi16(nhi), hi) << 32) |
                                   (uint64_t)_mm512_cmpneq_epi16_mask(_mm512_cvtepi8_epi16(nlo), lo)) & ~(uint64_t)1;
            if ((t_ & 31) == 0) voff_[t_ >> 5] = (uint32_t)vpos_;
            bm_[t_] = nz | (out8 ? 1u : 0u);
            dc_[t_] = c[0];
            t_++;
            uint8_t* o = vals_ + vpos_;
            if (!out8) {
                const __m512i bytes = _mm512_inserti64x4(_mm512_castsi256_si512(nlo), nhi, 1);
                _mm512_storeu_si512(o, _mm512_maskz_compress_epi8((__mmask64)nz, bytes));
                vpos_ += (size_t)__builtin_popcountll(nz);
            } else {
                const unsigned nlo16 = (unsigned)__builtin_popcountll(nz & 0xffffffffull);
                _mm512_storeu_si512(o, _mm512_maskz_compress_epi16((__mmask32)nz, lo));
                _mm512_storeu_si512(o + 2 * nlo16, _mm512_maskz_compress_epi16((__mmask32)(nz >> 32), hi));
                vpos_ += 2 * (size_t)__builtin_popcountll(nz);
            }
        }
    }
#endif
    SbsLayout lay_;
    uint8_t* base_ = nullptr;
    uint64_t* bm_ = nullptr;
    int16_t* dc_ = nullptr;
    uint32_t* voff_ = nullptr;
    uint8_t* vals_ = nullptr;
    size_t t_ = 0, vpos_ = 0;
};
#endif  // !__CUDACC__

}  // namespace b200jpg
