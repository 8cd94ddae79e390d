// host_decoder.cpp -- see host_decoder.h.  Follows reference src/decoder.rs:297-1298, src/parser.rs,
// src/huffman.rs, src/marker.rs (v0.3.2); error classes match the reference's Error variants.
#include "host_decoder.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

namespace b200jpg {

// src/decoder.rs:27-36
static const uint8_t UNZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

#define TRY(x)               \
    do {                     \
        int rc__ = (x);      \
        if (rc__) return rc__; \
    } while (0)

static inline bool is_sof(uint8_t m) { return m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC; }
static inline bool is_rst(uint8_t m) { return m >= 0xD0 && m <= 0xD7; }
static inline bool is_app(uint8_t m) { return m >= 0xE0 && m <= 0xEF; }

int HostDecoder::fail(int code, const char* fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    const char* prefix = code == B200JPG_ERR_FORMAT ? "invalid JPEG format: "
                         : code == B200JPG_ERR_UNSUPPORTED ? "unsupported JPEG feature: " : "";
    err_ = std::string(prefix) + buf;
    return code;
}

// ---- reader ---------------------------------------------------------------------------------
int HostDecoder::read_u8(uint8_t* b) {
    if (pos_ >= len_) return fail(B200JPG_ERR_IO, "failed to fill whole buffer");
    *b = data_[pos_++];
    return 0;
}
int HostDecoder::read_u16(uint16_t* v) {
    if (pos_ + 2 > len_) {
        pos_ = len_;
        return fail(B200JPG_ERR_IO, "failed to fill whole buffer");
    }
    *v = (uint16_t)((data_[pos_] << 8) | data_[pos_ + 1]);
    pos_ += 2;
    return 0;
}
int HostDecoder::read_exact(uint8_t* dst, size_t n) {
    if (pos_ + n > len_) {
        pos_ = len_;
        return fail(B200JPG_ERR_IO, "failed to fill whole buffer");
    }
    memcpy(dst, data_ + pos_, n);
    pos_ += n;
    return 0;
}
int HostDecoder::skip(size_t n) {
    if (pos_ + n > len_) {
        pos_ = len_;
        return fail(B200JPG_ERR_IO, "unexpected end of file");
    }
    pos_ += n;
    return 0;
}
// src/parser.rs:136-147
int HostDecoder::read_length(size_t* len) {
    uint16_t l;
    TRY(read_u16(&l));
    if (l < 2) return fail(B200JPG_ERR_FORMAT, "encountered marker with invalid length %u", l);
    *len = (size_t)l - 2;
    return 0;
}
// src/decoder.rs:766-791
int HostDecoder::read_marker(uint8_t* m) {
    for (;;) {
        uint8_t b;
        do {
            TRY(read_u8(&b));
        } while (b != 0xFF);
        TRY(read_u8(&b));
        while (b == 0xFF) TRY(read_u8(&b));
        if (b != 0x00) {
            *m = b;
            return 0;
        }
    }
}

// ---- huffman tables, src/huffman.rs:165-285 ---------------------------------------------------
static inline int16_t extend(uint16_t value, uint8_t count) {
    const uint16_t vt = (uint16_t)(1u << (count - 1));
    if (value < vt) return (int16_t)((int32_t)value - (int32_t)(1u << count) + 1);
    return (int16_t)value;
}

bool HuffTable::build(const uint8_t bits[16], const uint8_t* vals, int nvals, bool is_ac) {
    uint8_t huffsize[272];
    uint16_t huffcode[272];
    int n = 0;
    for (int i = 0; i < 16; i++) n += bits[i];
    if (n == 0 || n > 256 || n != nvals) return false;  // before anything is written: 16 counts of 255 would overrun huffsize
    n = 0;
    for (int i = 0; i < 16; i++)
        for (int k = 0; k < bits[i]; k++) huffsize[n++] = (uint8_t)(i + 1);
    uint8_t code_size = huffsize[0];
    uint32_t code = 0;
    for (int i = 0; i < n; i++) {
        while (code_size < huffsize[i]) {
            code <<= 1;
            code_size++;
        }
        if (code >= (1u << huffsize[i])) return false;  // "bad huffman code length"
        huffcode[i] = (uint16_t)code++;
    }
    present = true;
    nvalues = nvals;
    memcpy(values, vals, (size_t)nvals);
    memset(lut_value, 0, sizeof lut_value);
    memset(lut_size, 0, sizeof lut_size);
    memset(ac_value, 0, sizeof ac_value);
    memset(ac_run_size, 0, sizeof ac_run_size);
    int j = 0;
    for (int i = 0; i < 16; i++) {
        delta[i] = 0;
        maxcode[i] = -1;
        if (bits[i]) {
            delta[i] = j - (int32_t)huffcode[j];
            j += bits[i];
            maxcode[i] = huffcode[j - 1];
        }
    }
    for (int i = 0; i < n; i++) {
        const uint8_t size = huffsize[i];
        if (size > 8) continue;
        const int rem = 8 - size, start = huffcode[i] << rem;
        for (int b = 0; b < (1 << rem); b++) {
            lut_value[start + b] = vals[i];
            lut_size[start + b] = size;
        }
    }
    has_ac_lut = is_ac;
    if (is_ac)
        for (int i = 0; i < 256; i++) {
            const uint8_t value = lut_value[i], size = lut_size[i];
            const uint8_t run = value >> 4, mag = value & 0x0f;
            if (mag > 0 && size + mag <= 8) {
                const uint16_t un = (uint16_t)((((unsigned)i << size) & 0xffu) >> (8 - mag));
                ac_value[i] = extend(un, mag);
                ac_run_size[i] = (uint8_t)((run << 4) | (size + mag));
            }
        }
    return true;
}

// ---- bit reader, src/huffman.rs:20-161 --------------------------------------------------------
int HostDecoder::read_bits() {
    // Fast path (same result as the byte loop below): when the next 8 input bytes contain no 0xFF, as many
    // whole bytes as fit are appended in one go.
    if (!has_marker_ && num_bits_ <= 56 && pos_ + 8 <= len_) {
        uint64_t v;
        memcpy(&v, data_ + pos_, 8);
        // a byte is 0xFF iff its high bit is set and adding 1 to its low 7 bits reaches the high bit
        const uint64_t ff = v & 0x8080808080808080ull & ((v & 0x7f7f7f7f7f7f7f7full) + 0x0101010101010101ull);
        if (!ff) {
            const unsigned nbytes = (64u - num_bits_) >> 3;  // 1..8
            const uint64_t be = __builtin_bswap64(v);
            const uint64_t chunk = nbytes == 8 ? be : (be >> (64 - 8 * nbytes)) << (64 - 8 * nbytes);
            bits_ |= chunk >> num_bits_;
            num_bits_ = (uint8_t)(num_bits_ + 8 * nbytes);
            pos_ += nbytes;
            return 0;
        }
    }
    while (num_bits_ <= 56) {
        uint8_t byte = 0;
        if (!has_marker_) {
            TRY(read_u8(&byte));
        }
        if (byte == 0xFF) {
            uint8_t next;
            TRY(read_u8(&next));
            if (next != 0x00) {
                while (next == 0xFF) TRY(read_u8(&next));
                if (next == 0x00) return fail(B200JPG_ERR_FORMAT, "FF 00 found where marker was expected");
                has_marker_ = true;
                marker_ = next;
                continue;
            }
        }
        bits_ |= (uint64_t)byte << (56 - num_bits_);
        num_bits_ = (uint8_t)(num_bits_ + 8);
    }
    return 0;
}
int HostDecoder::get_bits(uint8_t count, uint16_t* v) {
    if (num_bits_ < count) TRY(read_bits());
    *v = count ? (uint16_t)((bits_ >> (64 - count)) & ((1u << count) - 1)) : 0;
    bits_ <<= count;
    num_bits_ = (uint8_t)(num_bits_ - count);
    return 0;
}
int HostDecoder::receive_extend(uint8_t count, int16_t* v) {
    uint16_t u;
    TRY(get_bits(count, &u));
    *v = extend(u, count);
    return 0;
}
int HostDecoder::huff_decode(const HuffTable& t, uint8_t* out) {
    if (num_bits_ < 16) TRY(read_bits());
    const unsigned idx = (unsigned)(bits_ >> 56);
    if (t.lut_size[idx] > 0) {
        bits_ <<= t.lut_size[idx];
        num_bits_ = (uint8_t)(num_bits_ - t.lut_size[idx]);
        *out = t.lut_value[idx];
        return 0;
    }
    const unsigned b16 = (unsigned)(bits_ >> 48);
    for (int i = 8; i < 16; i++) {
        const int32_t code = (int32_t)(b16 >> (15 - i));
        if (code <= t.maxcode[i]) {
            bits_ <<= (i + 1);
            num_bits_ = (uint8_t)(num_bits_ - (i + 1));
            const int32_t index = code + t.delta[i];
            if (index < 0 || index >= t.nvalues) return fail(B200JPG_ERR_INTERNAL, "huffman value index out of range (the reference panics)");
            *out = t.values[index];
            return 0;
        }
    }
    return fail(B200JPG_ERR_FORMAT, "failed to decode huffman code");
}
int HostDecoder::take_marker(bool* has, uint8_t* m) {
    TRY(read_bits());
    *has = has_marker_;
    *m = marker_;
    has_marker_ = false;
    return 0;
}

// ---- Annex K.3 tables used for MJPEG streams without DHT, src/huffman.rs:295-346 -------------------
static const uint8_t K3_BITS[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t K4_BITS[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t K_DC_VALS[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t K5_BITS[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D};
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
static const uint8_t K6_BITS[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
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

void HostDecoder::fill_default_mjpeg_tables(const ScanInfo& s) {
    bool dc0 = false, dc1 = false, ac0 = false, ac1 = false;
    for (int i = 0; i < s.n; i++) {
        dc0 |= s.dc_table[i] == 0;
        dc1 |= s.dc_table[i] == 1;
        ac0 |= s.ac_table[i] == 0;
        ac1 |= s.ac_table[i] == 1;
    }
    if (!dc_[0].present && dc0) dc_[0].build(K3_BITS, K_DC_VALS, 12, false);
    if (!dc_[1].present && dc1) dc_[1].build(K4_BITS, K_DC_VALS, 12, false);
    if (!ac_[0].present && ac0) ac_[0].build(K5_BITS, K5_VALS, 162, true);
    if (!ac_[1].present && ac1) ac_[1].build(K6_BITS, K6_VALS, 162, true);
}

// ---- segment parsers --------------------------------------------------------------------------
// src/parser.rs:161-280 and the frame checks of src/decoder.rs:338-389
int HostDecoder::parse_sof(uint8_t marker) {
    size_t length;
    TRY(read_length(&length));
    if (length <= 6) return fail(B200JPG_ERR_FORMAT, "invalid length in SOF");
    const int n = marker - 0xC0;
    FrameInfo f;
    f.is_baseline = n == 0;
    f.is_differential = (n >= 5 && n <= 7) || (n >= 13 && n <= 15);
    f.coding_process = (n == 0 || n == 1 || n == 5 || n == 9 || n == 13) ? B200JPG_CP_DCT_SEQUENTIAL
                       : (n == 2 || n == 6 || n == 10 || n == 14)        ? B200JPG_CP_DCT_PROGRESSIVE
                                                                         : B200JPG_CP_LOSSLESS;
    f.arithmetic = n >= 9;
    TRY(read_u8(&f.precision));
    if (f.precision == 8) {
    } else if (f.precision == 12) {
        if (f.is_baseline) return fail(B200JPG_ERR_FORMAT, "12 bit sample precision is not allowed in baseline");
    } else if (f.coding_process != B200JPG_CP_LOSSLESS || f.precision > 16) {
        return fail(B200JPG_ERR_FORMAT, "invalid precision %u in frame header", f.precision);
    }
    uint16_t height, width;
    TRY(read_u16(&height));
    TRY(read_u16(&width));
    if (height == 0) return fail(B200JPG_ERR_UNSUPPORTED, "DNL");
    if (width == 0) return fail(B200JPG_ERR_FORMAT, "zero width in frame header");
    uint8_t count;
    TRY(read_u8(&count));
    if (count == 0) return fail(B200JPG_ERR_FORMAT, "zero component count in frame header");
    if (f.coding_process == B200JPG_CP_DCT_PROGRESSIVE && count > 4)
        return fail(B200JPG_ERR_FORMAT, "progressive frame with more than 4 components");
    if (length != 6 + 3 * (size_t)count) return fail(B200JPG_ERR_FORMAT, "invalid length in SOF");
    for (int i = 0; i < count; i++) {
        uint8_t id, byte, tq;
        TRY(read_u8(&id));
        for (const auto& c : f.comps)
            if (c.identifier == id) return fail(B200JPG_ERR_FORMAT, "duplicate frame component identifier %u", id);
        TRY(read_u8(&byte));
        const uint8_t h = byte >> 4, v = byte & 0x0f;
        if (h == 0 || h > 4) return fail(B200JPG_ERR_FORMAT, "invalid horizontal sampling factor %u", h);
        if (v == 0 || v > 4) return fail(B200JPG_ERR_FORMAT, "invalid vertical sampling factor %u", v);
        TRY(read_u8(&tq));
        if (tq > 3 || (f.coding_process == B200JPG_CP_LOSSLESS && tq != 0))
            return fail(B200JPG_ERR_FORMAT, "invalid quantization table index %u", tq);
        b200jpg_component c;
        memset(&c, 0, sizeof c);
        c.identifier = id;
        c.h = h;
        c.v = v;
        c.tq = tq;
        c.dct_scale = 8;
        f.comps.push_back(c);
    }
    if (b200jpg_update_component_sizes(width, height, f.comps.data(), (int)f.comps.size(), &f.mcu_w, &f.mcu_h))
        return fail(B200JPG_ERR_FORMAT, "invalid dimensions");
    f.image_w = f.output_w = width;
    f.image_h = f.output_h = height;
    // src/decoder.rs:350-379
    if (f.is_differential) return fail(B200JPG_ERR_UNSUPPORTED, "Hierarchical");
    if (f.arithmetic) return fail(B200JPG_ERR_UNSUPPORTED, "ArithmeticEntropyCoding");
    if (f.precision != 8 && f.coding_process != B200JPG_CP_LOSSLESS) return fail(B200JPG_ERR_UNSUPPORTED, "SamplePrecision(%u)", f.precision);
    if (f.precision < 2 || f.precision > 16) return fail(B200JPG_ERR_UNSUPPORTED, "SamplePrecision(%u)", f.precision);
    if (count != 1 && count != 3 && count != 4) return fail(B200JPG_ERR_UNSUPPORTED, "ComponentCount(%u)", count);
    // Upsampler::new(&frame.components, image_size) only to validate the ratios (src/upsampler.rs:76-105)
    uint8_t hmax = 0, vmax = 0;
    for (const auto& c : f.comps) {
        hmax = std::max(hmax, c.h);
        vmax = std::max(vmax, c.v);
    }
    for (const auto& c : f.comps) {
        const bool h1 = c.h == hmax || width == 1, v1 = c.v == vmax || height == 1;
        const bool h2 = c.h * 2 == hmax, v2 = c.v * 2 == vmax;
        if ((h1 && v1) || (h2 && v1) || (h1 && v2) || (h2 && v2)) continue;
        if (hmax % c.h != 0 || vmax % c.v != 0) return fail(B200JPG_ERR_UNSUPPORTED, "NonIntegerSubsamplingRatio");
    }
    frame_ = f;
    has_frame_ = true;
    return 0;
}

// src/parser.rs:332-482
int HostDecoder::parse_sos(ScanInfo* s) {
    const FrameInfo& f = frame_;
    size_t length;
    TRY(read_length(&length));
    if (length == 0) return fail(B200JPG_ERR_FORMAT, "zero length in SOS");
    uint8_t count;
    TRY(read_u8(&count));
    if (count == 0 || count > 4) return fail(B200JPG_ERR_FORMAT, "invalid component count %u in scan header", count);
    if (length != 4 + 2 * (size_t)count) return fail(B200JPG_ERR_FORMAT, "invalid length in SOS");
    *s = ScanInfo();
    int maxidx = 0;
    for (int i = 0; i < count; i++) {
        uint8_t id, byte;
        TRY(read_u8(&id));
        int ci = -1;
        for (size_t k = 0; k < f.comps.size(); k++)
            if (f.comps[k].identifier == id) {
                ci = (int)k;
                break;
            }
        if (ci < 0)
            return fail(B200JPG_ERR_FORMAT, "scan component identifier %u does not match any of the component identifiers defined in the frame", id);
        for (int k = 0; k < i; k++)
            if (s->comp_index[k] == ci) return fail(B200JPG_ERR_FORMAT, "duplicate scan component identifier %u", id);
        if (ci < maxidx) return fail(B200JPG_ERR_FORMAT, "the scan component order does not follow the order in the frame header");
        TRY(read_u8(&byte));
        const uint8_t dc = byte >> 4, ac = byte & 0x0f;
        if (dc > 3 || (f.is_baseline && dc > 1)) return fail(B200JPG_ERR_FORMAT, "invalid dc table index %u", dc);
        if (ac > 3 || (f.is_baseline && ac > 1)) return fail(B200JPG_ERR_FORMAT, "invalid ac table index %u", ac);
        s->comp_index[i] = ci;
        s->dc_table[i] = dc;
        s->ac_table[i] = ac;
        maxidx = std::max(maxidx, ci);
    }
    s->n = count;
    uint32_t blocks_per_mcu = 0;
    for (int i = 0; i < count; i++) blocks_per_mcu += (uint32_t)f.comps[s->comp_index[i]].h * f.comps[s->comp_index[i]].v;
    if (count > 1 && blocks_per_mcu > 10) return fail(B200JPG_ERR_FORMAT, "scan with more than one component and more than 10 blocks per MCU");
    uint8_t ss = 0, se = 0, byte = 0;
    TRY(read_u8(&ss));
    TRY(read_u8(&se));
    TRY(read_u8(&byte));
    const uint8_t ah = byte >> 4, al = byte & 0x0f;
    if (al >= f.precision) return fail(B200JPG_ERR_FORMAT, "invalid point transform, must be less than the frame precision");
    if (f.coding_process == B200JPG_CP_DCT_PROGRESSIVE) {
        if (se > 63 || ss > se || (ss == 0 && se != 0)) return fail(B200JPG_ERR_FORMAT, "invalid spectral selection parameters: ss=%u, se=%u", ss, se);
        if (ss != 0 && count != 1) return fail(B200JPG_ERR_FORMAT, "spectral selection scan with AC coefficients can't have more than one component");
        if (ah > 13 || al > 13) return fail(B200JPG_ERR_FORMAT, "invalid successive approximation parameters: ah=%u, al=%u", ah, al);
        if (ah != 0 && ah != al + 1) return fail(B200JPG_ERR_FORMAT, "successive approximation scan with more than one bit of improvement");
    } else if (f.coding_process == B200JPG_CP_LOSSLESS) {
        if (se != 0) return fail(B200JPG_ERR_FORMAT, "spectral selection end shall be zero in lossless scan");
        if (ah != 0) return fail(B200JPG_ERR_FORMAT, "successive approximation high shall be zero in lossless scan");
        if (ss > 7) return fail(B200JPG_ERR_FORMAT, "invalid predictor selection value: %u", ss);
    } else {
        if (se == 0) se = 63;
        if (ss != 0 || se != 63) return fail(B200JPG_ERR_FORMAT, "spectral selection is not allowed in non-progressive scan");
        if (ah != 0 || al != 0) return fail(B200JPG_ERR_FORMAT, "successive approximation is not allowed in non-progressive scan");
    }
    s->ss_start = ss;
    s->ss_end = (uint8_t)(se + 1);
    s->ah = ah;
    s->al = al;
    return 0;
}

// src/parser.rs:485-532, de-zigzag src/decoder.rs:485-498
int HostDecoder::parse_dqt() {
    size_t length;
    TRY(read_length(&length));
    uint16_t tables[4][64];
    bool got[4] = {false, false, false, false};
    while (length > 0) {
        uint8_t byte;
        TRY(read_u8(&byte));
        const size_t precision = byte >> 4, index = byte & 0x0f;
        if (precision > 1) return fail(B200JPG_ERR_FORMAT, "invalid precision %zu in DQT", precision);
        if (index > 3) return fail(B200JPG_ERR_FORMAT, "invalid destination identifier %zu in DQT", index);
        if (length < 65 + 64 * precision) return fail(B200JPG_ERR_FORMAT, "invalid length in DQT");
        for (int i = 0; i < 64; i++) {
            if (precision == 0) {
                uint8_t b;
                TRY(read_u8(&b));
                tables[index][i] = b;
            } else {
                TRY(read_u16(&tables[index][i]));
            }
        }
        for (int i = 0; i < 64; i++)
            if (tables[index][i] == 0) return fail(B200JPG_ERR_FORMAT, "quantization table contains element with a zero value");
        got[index] = true;
        length -= 65 + 64 * precision;
    }
    for (int t = 0; t < 4; t++)
        if (got[t]) {
            for (int j = 0; j < 64; j++) qt_[t][UNZIGZAG[j]] = tables[t][j];
            has_qt_[t] = true;
        }
    return 0;
}

// src/parser.rs:536-589, merge src/decoder.rs:501-518
int HostDecoder::parse_dht() {
    size_t length;
    TRY(read_length(&length));
    std::vector<HuffTable> ndc(4), nac(4);
    while (length > 17) {
        uint8_t byte;
        TRY(read_u8(&byte));
        const uint8_t cls = byte >> 4;
        const size_t index = byte & 0x0f;
        if (cls != 0 && cls != 1) return fail(B200JPG_ERR_FORMAT, "invalid class %u in DHT", cls);
        if (has_frame_ && frame_.is_baseline && index > 1)
            return fail(B200JPG_ERR_FORMAT, "a maximum of two huffman tables per class are allowed in baseline");
        if (index > 3) return fail(B200JPG_ERR_FORMAT, "invalid destination identifier %zu in DHT", index);
        uint8_t counts[16];
        TRY(read_exact(counts, 16));
        size_t size = 0;
        for (int i = 0; i < 16; i++) size += counts[i];
        if (size == 0) return fail(B200JPG_ERR_FORMAT, "encountered table with zero length in DHT");
        if (size > 256) return fail(B200JPG_ERR_FORMAT, "encountered table with excessive length in DHT");
        if (size > length - 17) return fail(B200JPG_ERR_FORMAT, "invalid length in DHT");
        uint8_t values[256];
        TRY(read_exact(values, size));
        HuffTable& dst = cls == 0 ? ndc[index] : nac[index];
        if (!dst.build(counts, values, (int)size, cls == 1)) return fail(B200JPG_ERR_FORMAT, "bad huffman code length");
        length -= 17 + size;
    }
    if (length != 0) return fail(B200JPG_ERR_FORMAT, "invalid length in DHT");
    for (int i = 0; i < 4; i++) {
        if (ndc[i].present) dc_[i] = ndc[i];
        if (nac[i].present) ac_[i] = nac[i];
    }
    return 0;
}

// src/parser.rs:613-710, dispatch src/decoder.rs:532-558
int HostDecoder::parse_app(uint8_t marker) {
    size_t length, bytes_read = 0;
    TRY(read_length(&length));
    const int n = marker - 0xE0;
    if (n == 0) {
        if (length >= 5) {
            uint8_t b[5];
            TRY(read_exact(b, 5));
            bytes_read = 5;
            if (!memcmp(b, "JFIF\0", 5)) is_jfif_ = true;
            else if (!memcmp(b, "AVI1\0", 5)) is_mjpeg_ = true;
        }
    } else if (n == 1) {
        std::vector<uint8_t> buf(length);
        TRY(read_exact(buf.data(), length));
        bytes_read = length;
        if (length >= 6 && !memcmp(buf.data(), "Exif\0\0", 6)) {
            exif_.assign(buf.begin() + 6, buf.end());
            has_exif_ = true;
        } else if (length >= 29 && !memcmp(buf.data(), "http://ns.adobe.com/xap/1.0/\0", 29)) {
            xmp_.assign(buf.begin() + 29, buf.end());
            has_xmp_ = true;
        }
    } else if (n == 2) {
        if (length > 14) {
            uint8_t b[14];
            TRY(read_exact(b, 14));
            bytes_read = 14;
            if (!memcmp(b, "ICC_PROFILE\0", 12)) {
                IccChunk c;
                c.seq_no = b[12];
                c.num_markers = b[13];
                c.data.resize(length - 14);
                TRY(read_exact(c.data.data(), length - 14));
                bytes_read = length;
                icc_.push_back(std::move(c));
            }
        }
    } else if (n == 13) {
        if (length >= 14) {
            uint8_t b[14];
            TRY(read_exact(b, 14));
            bytes_read = 14;
            if (!memcmp(b, "Photoshop 3.0\0", 14)) {  // PSIR: read and dropped
                TRY(skip(length - 14));
                bytes_read = length;
            }
        }
    } else if (n == 14) {
        if (length >= 12) {
            uint8_t b[12];
            TRY(read_exact(b, 12));
            bytes_read = 12;
            if (!memcmp(b, "Adobe\0", 6)) {
                if (b[11] > 2) return fail(B200JPG_ERR_FORMAT, "invalid color transform in adobe app segment");
                has_adobe_ = true;
                adobe_ = b[11];
            }
        }
    }
    return skip(length - bytes_read);
}

// ---- block decoding, src/decoder.rs:1086-1298 --------------------------------------------------
// nz (may be null): the block's non-zero map, kept in step with every AC coefficient written (a first pass that runs twice
// over a band -- malformed files -- may overwrite a coefficient, and a value can shift out to zero: the bit follows the value)
static inline void note_coef(uint64_t* nz, unsigned index, int16_t stored) {
    if (nz) *nz = (*nz & ~((uint64_t)1 << index)) | ((uint64_t)(stored != 0) << index);
}

int HostDecoder::decode_block(int16_t* c, uint64_t* nz, const HuffTable& dc, const HuffTable& ac, const ScanInfo& s, uint16_t* eob_run,
                              int16_t* pred) {
    const uint8_t al = s.al, ss_end = s.ss_end;
    if (s.ss_start == 0) {
        uint8_t value;
        TRY(huff_decode(dc, &value));
        int16_t diff = 0;
        if (value != 0) {
            if (value > 11) return fail(B200JPG_ERR_FORMAT, "invalid DC difference magnitude category");
            TRY(receive_extend(value, &diff));
        }
        *pred = (int16_t)((uint16_t)*pred + (uint16_t)diff);  // wrapping_add, src/decoder.rs:1117
        c[0] = (int16_t)((uint16_t)*pred << al);
    }
    uint8_t index = s.ss_start > 1 ? s.ss_start : 1;
    if (index < ss_end && *eob_run > 0) {
        *eob_run -= 1;
        return 0;
    }
    while (index < ss_end) {
        // decode_fast_ac, src/huffman.rs:60-78
        if (num_bits_ < 8) TRY(read_bits());
        const unsigned idx = (unsigned)(bits_ >> 56);
        const uint8_t rs = ac.ac_run_size[idx];
        if (rs != 0) {
            bits_ <<= (rs & 0x0f);
            num_bits_ = (uint8_t)(num_bits_ - (rs & 0x0f));
            index = (uint8_t)(index + (rs >> 4));
            if (index >= ss_end) break;
            c[UNZIGZAG[index]] = (int16_t)((uint16_t)ac.ac_value[idx] << al);
            note_coef(nz, index, c[UNZIGZAG[index]]);
            index++;
            continue;
        }
        uint8_t byte;
        TRY(huff_decode(ac, &byte));
        const uint8_t r = byte >> 4, sz = byte & 0x0f;
        if (sz == 0) {
            if (r == 15) {
                index = (uint8_t)(index + 16);
            } else {
                *eob_run = (uint16_t)((1u << r) - 1);
                if (r > 0) {
                    uint16_t extra;
                    TRY(get_bits(r, &extra));
                    *eob_run = (uint16_t)(*eob_run + extra);
                }
                break;
            }
        } else {
            index = (uint8_t)(index + r);
            if (index >= ss_end) break;
            int16_t v;
            TRY(receive_extend(sz, &v));
            c[UNZIGZAG[index]] = (int16_t)((uint16_t)v << al);
            note_coef(nz, index, c[UNZIGZAG[index]]);
            index++;
        }
    }
    return 0;
}

// Where decode_block_seq puts what it decodes: the dense block of the reference (natural order,
// src/decoder.rs:1137-1172) or one block of the sparse stream (zig-zag order, sbs.h).
namespace {
struct DenseSink {
    int16_t* c;
    inline void dc(int16_t v) { c[0] = v; }
    inline void ac(unsigned k, int16_t v) { c[UNZIGZAG[k]] = v; }
    inline void done() {}
};
struct SbsSink {
    SbsWriter* w;
    uint64_t bits = 0;
    unsigned n = 0, acc = 0;
    int16_t dcv = 0;
    alignas(16) int16_t v[64 + 16];
    explicit SbsSink(SbsWriter* wr) : w(wr) {}
    inline void dc(int16_t x) { dcv = x; }
    inline void ac(unsigned k, int16_t x) {
        v[n++] = x;
        bits |= (uint64_t)1 << k;
        acc |= (unsigned)(x + 128);  // > 255 as soon as one value is outside int8
    }
    inline void done() { w->put(bits, dcv, v, n, acc > 255u); }
};
}  // namespace

// decode_block for sequential scans (ss = 0..63, al = 0), the hot loop of every baseline JPEG: identical
// decisions and refill thresholds (16 bits before a code, 8 before the fast-AC probe, `count` before
// receive_extend -- src/huffman.rs:31-96), with the bit buffer kept in locals.
template <class Sink>
int HostDecoder::decode_block_seq(Sink& sink, const HuffTable& dc, const HuffTable& ac, uint16_t* eob_run, int16_t* pred) {
    uint64_t bits = bits_;
    unsigned nb = num_bits_;
#define SEQ_REFILL()                   \
    do {                               \
        bits_ = bits;                  \
        num_bits_ = (uint8_t)nb;       \
        TRY(read_bits());              \
        bits = bits_;                  \
        nb = num_bits_;                \
    } while (0)
#define SEQ_SLOW_CODE(table, out)      \
    do {                               \
        bits_ = bits;                  \
        num_bits_ = (uint8_t)nb;       \
        TRY(huff_decode((table), &(out))); \
        bits = bits_;                  \
        nb = num_bits_;                \
    } while (0)
    {
        if (nb < 16) SEQ_REFILL();
        const unsigned idx = (unsigned)(bits >> 56);
        uint8_t value;
        const unsigned size = dc.lut_size[idx];
        if (size) {
            value = dc.lut_value[idx];
            bits <<= size;
            nb -= size;
        } else {
            SEQ_SLOW_CODE(dc, value);
        }
        int16_t diff = 0;
        if (value != 0) {
            if (value > 11) return fail(B200JPG_ERR_FORMAT, "invalid DC difference magnitude category");
            if (nb < value) SEQ_REFILL();
            const uint16_t u = (uint16_t)(bits >> (64 - value));
            bits <<= value;
            nb -= value;
            diff = extend(u, value);
        }
        *pred = (int16_t)((uint16_t)*pred + (uint16_t)diff);
        sink.dc(*pred);
    }
    if (*eob_run > 0) {
        *eob_run -= 1;
        bits_ = bits;
        num_bits_ = (uint8_t)nb;
        sink.done();
        return 0;
    }
    unsigned index = 1;
    while (index < 64) {
        if (nb < 8) SEQ_REFILL();
        const unsigned idx = (unsigned)(bits >> 56);
        const unsigned rs = ac.ac_run_size[idx];
        if (rs != 0) {  // decode_fast_ac hit: code and value bits in one probe
            bits <<= (rs & 0x0f);
            nb -= (rs & 0x0f);
            index += rs >> 4;
            if (index >= 64) break;
            sink.ac(index, ac.ac_value[idx]);
            index++;
            continue;
        }
        if (nb < 16) SEQ_REFILL();
        const unsigned idx2 = (unsigned)(bits >> 56);
        uint8_t byte;
        const unsigned size = ac.lut_size[idx2];
        if (size) {
            byte = ac.lut_value[idx2];
            bits <<= size;
            nb -= size;
        } else {
            SEQ_SLOW_CODE(ac, byte);
        }
        const unsigned r = byte >> 4, sz = byte & 0x0f;
        if (sz == 0) {
            if (r == 15) {
                index += 16;
            } else {
                *eob_run = (uint16_t)((1u << r) - 1);
                if (r > 0) {
                    if (nb < r) SEQ_REFILL();
                    *eob_run = (uint16_t)(*eob_run + (uint16_t)(bits >> (64 - r)));
                    bits <<= r;
                    nb -= r;
                }
                break;
            }
        } else {
            index += r;
            if (index >= 64) break;
            if (nb < sz) SEQ_REFILL();
            const uint16_t u = (uint16_t)(bits >> (64 - sz));
            bits <<= sz;
            nb -= sz;
            sink.ac(index, extend(u, (uint8_t)sz));
            index++;
        }
    }
    bits_ = bits;
    num_bits_ = (uint8_t)nb;
    sink.done();
    return 0;
#undef SEQ_REFILL
#undef SEQ_SLOW_CODE
}

int HostDecoder::refine_non_zeroes(int16_t* c, uint8_t start, uint8_t end, uint8_t zrl, int16_t bit, uint8_t* ret) {
    const uint8_t last = (uint8_t)(end - 1);
    uint8_t zero_run_length = zrl;
    for (uint8_t i = start; i < end; i++) {
        int16_t* co = &c[UNZIGZAG[i]];
        if (*co == 0) {
            if (zero_run_length == 0) {
                *ret = i;
                return 0;
            }
            zero_run_length--;
        } else {
            uint16_t b;
            TRY(get_bits(1, &b));
            if (b == 1 && (*co & bit) == 0) {
                const int32_t v = *co > 0 ? (int32_t)*co + bit : (int32_t)*co - bit;
                if (v > 32767 || v < -32768) return fail(B200JPG_ERR_FORMAT, "Coefficient overflow");
                *co = (int16_t)v;
            }
        }
    }
    *ret = last;
    return 0;
}

// refine_non_zeroes over the block's non-zero map: the same bits read in the same order, the same coefficients touched, the
// same index returned.  The loop above stops at the (zrl + 1)-th ZERO coefficient of [start, end) and reads one correction bit
// for every NON-ZERO one it passes; here the stop position is the (zrl + 1)-th set bit of the band's zero mask, and the
// coefficients to correct are the set bits of the map below it.  (A correction never makes a coefficient zero, so the map
// itself does not change.)
int HostDecoder::refine_non_zeroes_map(int16_t* c, uint64_t nz, uint8_t start, uint8_t end, uint8_t zrl, int16_t bit, uint8_t* ret) {
    const uint64_t below_end = end >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << end) - 1);
    const uint64_t band = start >= 64 ? 0 : below_end & ~(((uint64_t)1 << start) - 1);
    unsigned stop = end;
    bool found = false;
    if (zrl < 64) {
        uint64_t z = ~nz & band;
        if ((unsigned)__builtin_popcountll(z) > zrl) {
            for (unsigned r = zrl; r > 0; r--) z &= z - 1;
            stop = (unsigned)__builtin_ctzll(z);
            found = true;
        }
    }
    uint64_t m = nz & band & (stop >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << stop) - 1));
    while (m) {
        const unsigned i = (unsigned)__builtin_ctzll(m);
        m &= m - 1;
        if (num_bits_ < 1) TRY(read_bits());  // get_bits(1)
        const bool b = (bits_ >> 63) != 0;
        bits_ <<= 1;
        num_bits_ = (uint8_t)(num_bits_ - 1);
        int16_t* co = &c[UNZIGZAG[i]];
        if (b && (*co & bit) == 0) {
            const int32_t v = *co > 0 ? (int32_t)*co + bit : (int32_t)*co - bit;
            if (v > 32767 || v < -32768) return fail(B200JPG_ERR_FORMAT, "Coefficient overflow");
            *co = (int16_t)v;
        }
    }
    *ret = found ? (uint8_t)stop : (uint8_t)(end - 1);
    return 0;
}

int HostDecoder::decode_block_sa(int16_t* c, uint64_t* nz, const HuffTable& ac, const ScanInfo& s, uint16_t* eob_run) {
    const int16_t bit = (int16_t)(1 << s.al);
    if (s.ss_start == 0) {
        uint16_t b;
        TRY(get_bits(1, &b));
        if (b == 1) c[0] |= bit;
        return 0;
    }
    if (*eob_run > 0) {
        *eob_run -= 1;
        uint8_t r;
        return nz ? refine_non_zeroes_map(c, *nz, s.ss_start, s.ss_end, 64, bit, &r) : refine_non_zeroes(c, s.ss_start, s.ss_end, 64, bit, &r);
    }
    uint8_t index = s.ss_start;
    while (index < s.ss_end) {
        uint8_t byte;
        TRY(huff_decode(ac, &byte));
        const uint8_t r = byte >> 4, sz = byte & 0x0f;
        uint8_t zero_run_length = r;
        int16_t value = 0;
        if (sz == 0) {
            if (r != 15) {
                *eob_run = (uint16_t)((1u << r) - 1);
                if (r > 0) {
                    uint16_t extra;
                    TRY(get_bits(r, &extra));
                    *eob_run = (uint16_t)(*eob_run + extra);
                }
                zero_run_length = 64;
            }
        } else if (sz == 1) {
            uint16_t b;
            TRY(get_bits(1, &b));
            value = b == 1 ? bit : (int16_t)-bit;
        } else {
            return fail(B200JPG_ERR_FORMAT, "unexpected huffman code");
        }
        if (nz) TRY(refine_non_zeroes_map(c, *nz, index, s.ss_end, zero_run_length, bit, &index));
        else TRY(refine_non_zeroes(c, index, s.ss_end, zero_run_length, bit, &index));
        if (value != 0) {
            c[UNZIGZAG[index]] = value;
            note_coef(nz, index, value);
        }
        index++;
    }
    return 0;
}

// ---- decode_scan, src/decoder.rs:794-1082 ------------------------------------------------------
int HostDecoder::decode_scan(const ScanInfo& scan, const bool finished[4], bool* out_has_marker, uint8_t* out_marker) {
    const FrameInfo& frame = frame_;
    const int nc = scan.n;
    b200jpg_component comps[4];
    for (int i = 0; i < nc; i++) comps[i] = frame.comps[scan.comp_index[i]];
    for (int i = 0; i < nc; i++)
        if (!has_qt_[comps[i].tq]) return fail(B200JPG_ERR_FORMAT, "use of unset quantization table");
    if (is_mjpeg_) fill_default_mjpeg_tables(scan);
    if (scan.ss_start == 0)
        for (int i = 0; i < nc; i++)
            if (!dc_[scan.dc_table[i]].present) return fail(B200JPG_ERR_FORMAT, "scan makes use of unset dc huffman table");
    if (scan.ss_end > 1)
        for (int i = 0; i < nc; i++)
            if (!ac_[scan.ac_table[i]].present) return fail(B200JPG_ERR_FORMAT, "scan makes use of unset ac huffman table");

    const bool is_progressive = frame.coding_process == B200JPG_CP_DCT_PROGRESSIVE;
    const bool is_interleaved = nc > 1;
    const bool sequential = scan.ss_start == 0 && scan.ss_end == 64 && scan.al == 0 && scan.ah == 0;
    // Sparse stream straight from the Huffman loop: only when this one scan delivers every block of every
    // component, in an order K0 can map back to raster positions (interleaved MCUs, or a lone 1x1 component).
    bool direct = sbs_base_ != nullptr && sequential && !is_progressive && !sbs_direct_ && nc == (int)frame.comps.size() &&
                  (is_interleaved || (comps[0].h == 1 && comps[0].v == 1));
    for (int i = 0; i < nc; i++) direct = direct && finished[i] && scan.comp_index[i] == i;
    if (direct) sbs_.begin(sbs_base_, total_blocks());
    // where each scan component's blocks go: the progressive store, the final buffer (worker::start
    // zero-fills, src/decoder.rs:848-861, 874-880), or a dummy block
    int16_t* target[4] = {nullptr, nullptr, nullptr, nullptr};
    bool zero_per_block[4] = {false, false, false, false};
    int16_t dummy[64];
    for (int i = 0; i < nc; i++) {
        const int ci = scan.comp_index[i];
        const size_t count = (size_t)comps[i].block_w * comps[i].block_h * 64;
        if (finished[i]) memcpy(final_qt_[ci], qt_[comps[i].tq], 128);  // RowData.quantization_table
        if (is_progressive) {
            target[i] = work_[ci].data();
        } else if (finished[i] && direct) {
            have_final_[ci] = false;
        } else if (finished[i]) {
            have_final_[ci] = false;
            // An interleaved scan visits every block of the component exactly once, so each block is zeroed right
            // before it is decoded (cache-hot, one pass over the buffer instead of two); otherwise zero up front.
            if (ext_[ci]) {
                if (!is_interleaved) memset(ext_[ci], 0, count * sizeof(int16_t));
                target[i] = ext_[ci];
                zero_per_block[i] = is_interleaved;
            } else {
                final_[ci].assign(count, 0);
                target[i] = final_[ci].data();
            }
        }
    }
    bits_ = 0;
    num_bits_ = 0;
    has_marker_ = false;  // HuffmanDecoder::new()
    int16_t dc_predictors[4] = {0, 0, 0, 0};
    uint16_t mcus_left = restart_interval_;
    uint8_t expected_rst = 0;
    uint16_t eob_run = 0;
    uint32_t mh[4] = {1, 1, 1, 1}, mv[4] = {1, 1, 1, 1};
    uint32_t max_mcu_x, max_mcu_y;
    if (is_interleaved) {
        for (int i = 0; i < nc; i++) {
            mh[i] = comps[i].h;
            mv[i] = comps[i].v;
        }
        max_mcu_x = frame.mcu_w;
        max_mcu_y = frame.mcu_h;
    } else {
        max_mcu_x = comps[0].block_w;
        max_mcu_y = comps[0].block_h;
    }
    for (uint32_t mcu_y = 0; mcu_y < max_mcu_y; mcu_y++) {
        if (mcu_y * 8 >= frame.image_h) break;  // src/decoder.rs:911-913
        for (uint32_t mcu_x = 0; mcu_x < max_mcu_x; mcu_x++) {
            if (mcu_x * 8 >= frame.image_w) break;
            if (restart_interval_ > 0) {
                if (mcus_left == 0) {
                    bool has;
                    uint8_t m;
                    TRY(take_marker(&has, &m));
                    if (has && is_rst(m)) {
                        const uint8_t n = (uint8_t)(m - 0xD0);
                        if (n != expected_rst) return fail(B200JPG_ERR_FORMAT, "found RST%u where RST%u was expected", n, expected_rst);
                        bits_ = 0;
                        num_bits_ = 0;
                        memset(dc_predictors, 0, sizeof dc_predictors);
                        eob_run = 0;
                        expected_rst = (uint8_t)((expected_rst + 1) % 8);
                        mcus_left = restart_interval_;
                    } else if (has) {
                        return fail(B200JPG_ERR_FORMAT, "found marker 0x%02X inside scan where RST%u was expected", m, expected_rst);
                    } else {
                        return fail(B200JPG_ERR_FORMAT, "no marker found where RST%u was expected", expected_rst);
                    }
                }
                mcus_left--;
            }
            for (int i = 0; i < nc; i++) {
                const b200jpg_component& comp = comps[i];
                for (uint32_t v_pos = 0; v_pos < mv[i]; v_pos++)
                    for (uint32_t h_pos = 0; h_pos < mh[i]; h_pos++) {
                        if (direct) {
                            SbsSink sink(&sbs_);
                            TRY(decode_block_seq(sink, dc_[scan.dc_table[i]], ac_[scan.ac_table[i]], &eob_run, &dc_predictors[i]));
                            continue;
                        }
                        int16_t* c;
                        uint64_t* nz = nullptr;
                        if (target[i]) {
                            const size_t block_y = (size_t)mcu_y * mv[i] + v_pos, block_x = (size_t)mcu_x * mh[i] + h_pos;
                            c = target[i] + (block_y * comp.block_w + block_x) * 64;
                            if (is_progressive) nz = nz_[scan.comp_index[i]].data() + (block_y * comp.block_w + block_x);
                            if (zero_per_block[i]) memset(c, 0, 128);
                        } else {
                            c = dummy;
                            if (scan.ah == 0) memset(dummy, 0, sizeof dummy);
                        }
                        if (sequential) {
                            DenseSink sink{c};
                            TRY(decode_block_seq(sink, dc_[scan.dc_table[i]], ac_[scan.ac_table[i]], &eob_run, &dc_predictors[i]));
                            if (nz) {  // a full-band first pass inside a progressive frame: rebuild the block's map
                                uint64_t m = 0;
                                for (unsigned k = 1; k < 64; k++) m |= (uint64_t)(c[UNZIGZAG[k]] != 0) << k;
                                *nz = m;
                            }
                        } else if (scan.ah == 0)
                            TRY(decode_block(c, nz, dc_[scan.dc_table[i]], ac_[scan.ac_table[i]], scan, &eob_run, &dc_predictors[i]));
                        else
                            TRY(decode_block_sa(c, nz, ac_[scan.ac_table[i]], scan, &eob_run));
                    }
            }
        }
    }
    bool has;
    uint8_t m;
    TRY(take_marker(&has, &m));
    while (has && is_rst(m)) {
        const std::string saved = err_;
        if (read_marker(&m)) {  // .ok()
            has = false;
            err_ = saved;
        }
    }
    *out_has_marker = has;
    *out_marker = m;
    // what the worker has received by the end of this scan
    for (int i = 0; i < nc; i++)
        if (finished[i]) {
            const int ci = scan.comp_index[i];
            if (is_progressive) {
                if (ext_[ci]) memcpy(ext_[ci], work_[ci].data(), work_[ci].size() * sizeof(int16_t));
                else final_[ci] = work_[ci];
            }
            have_final_[ci] = true;
        }
    if (direct) sbs_direct_ = true;
    return 0;
}

// The conditions under which decode_scan would take its `direct` route with nothing left to go wrong before the
// first code word: every check decode_scan makes up front holds, so the host would start decoding here.
bool HostDecoder::device_scan_ok(const ScanInfo& scan) const {
    const FrameInfo& f = frame_;
    if (f.coding_process != B200JPG_CP_DCT_SEQUENTIAL || f.precision != 8 || is_mjpeg_) return false;
    if (scan.n != (int)f.comps.size() || scan.ss_start != 0 || scan.ss_end != 64 || scan.al != 0 || scan.ah != 0) return false;
    if (scan.n == 1 && (f.comps[0].h != 1 || f.comps[0].v != 1)) return false;
    unsigned bpm = 0;
    for (int i = 0; i < scan.n; i++) {
        if (scan.comp_index[i] != i || finished_mask_[i] != 0) return false;
        if (!has_qt_[f.comps[(size_t)i].tq] || !dc_[scan.dc_table[i]].present || !ac_[scan.ac_table[i]].present) return false;
        bpm += (unsigned)f.comps[(size_t)i].h * f.comps[(size_t)i].v;
    }
    if (bpm > 10 || buffer_limit_exceeded()) return false;
    return true;
}
void HostDecoder::capture_final_qtables() {
    for (size_t i = 0; i < frame_.comps.size() && i < 4; i++) memcpy(final_qt_[i], qt_[frame_.comps[i].tq], 128);
}

// Ends the sparse stream: a directly written one is padded; otherwise the dense per-component buffers the
// scans produced are compacted in planar order (the reference's worker would receive exactly these blocks).
int HostDecoder::finish_sbs() {
    sbs_len_ = 0;
    if (!sbs_base_) return 0;
    for (size_t i = 0; i < frame_.comps.size(); i++)
        if (!have_final_[i]) return 0;  // "not all components have data": reported by the caller
    if (!sbs_direct_) {
        sbs_.begin(sbs_base_, total_blocks());
        for (size_t i = 0; i < frame_.comps.size(); i++) {
            const size_t n = (size_t)frame_.comps[i].block_w * frame_.comps[i].block_h;
            const int16_t* c = coefficients((int)i);
            for (size_t b = 0; b < n; b++) sbs_.put_dense(c + 64 * b);
        }
    }
    sbs_len_ = sbs_.finish();
    return 0;
}

// ---- the marker loop, src/decoder.rs:297-615 ---------------------------------------------------
int HostDecoder::decode_internal(bool stop_after_metadata) {
    if (stop_after_metadata && has_frame_) return 0;
    if (!has_frame_) {
        uint8_t a, b;
        TRY(read_u8(&a));
        if (a != 0xFF) return fail(B200JPG_ERR_FORMAT, "first two bytes are not an SOI marker");
        TRY(read_u8(&b));
        if (b != 0xD8) return fail(B200JPG_ERR_FORMAT, "first two bytes are not an SOI marker");
    }
    uint8_t previous_marker = 0xD8;
    bool has_pending = false;
    uint8_t pending = 0;
    int scans_processed = 0;
    for (int i = 0; i < 4; i++) have_final_[i] = false;  // planes = vec![Vec::new(); n]
    for (;;) {
        uint8_t marker;
        if (has_pending) {
            marker = pending;
            has_pending = false;
        } else {
            TRY(read_marker(&marker));
        }
        if (is_sof(marker)) {
            if (has_frame_) return fail(B200JPG_ERR_UNSUPPORTED, "Hierarchical");
            TRY(parse_sof(marker));
            if (stop_after_metadata) return 0;
        } else if (marker == 0xDA) {
            if (!has_frame_) return fail(B200JPG_ERR_FORMAT, "scan encountered before frame");
            ScanInfo scan;
            TRY(parse_sos(&scan));
            const FrameInfo& frame = frame_;
            if (frame.coding_process == B200JPG_CP_DCT_PROGRESSIVE && !has_work_) {
                for (size_t i = 0; i < frame.comps.size(); i++)
                    work_[i].assign((size_t)frame.comps[i].block_w * frame.comps[i].block_h * 64, 0);
                for (size_t i = 0; i < frame.comps.size(); i++) nz_[i].assign((size_t)frame.comps[i].block_w * frame.comps[i].block_h, 0);
                has_work_ = true;
            }
            if (frame.coding_process == B200JPG_CP_LOSSLESS)
                return fail(B200JPG_ERR_UNSUPPORTED, "lossless JPEG (SOF3) bypasses the worker path and is not built (SURVEY section 2)");
            if (probe_device_ && scans_processed == 0 && device_scan_ok(scan)) {
                device_scan_.eligible = true;
                device_scan_.scan_begin = pos_;
                device_scan_.scan = scan;
                device_scan_.restart_interval = restart_interval_;
                capture_final_qtables();
                return B200JPG_INTERNAL_DEVICE_SCAN;
            }
            bool finished[4] = {false, false, false, false};
            if (scan.al == 0)
                for (int k = 0; k < scan.n; k++) {
                    const int i = scan.comp_index[k];
                    if (finished_mask_[i] == ~(uint64_t)0) continue;
                    for (int j = scan.ss_start; j < scan.ss_end; j++) finished_mask_[i] |= (uint64_t)1 << j;
                    if (finished_mask_[i] == ~(uint64_t)0) finished[k] = true;
                }
            bool hm = false;
            uint8_t m = 0;
            TRY(decode_scan(scan, finished, &hm, &m));
            has_pending = hm;
            pending = m;
            scans_processed++;
        } else if (marker == 0xDB) {
            TRY(parse_dqt());
        } else if (marker == 0xC4) {
            TRY(parse_dht());
        } else if (marker == 0xCC) {
            return fail(B200JPG_ERR_UNSUPPORTED, "ArithmeticEntropyCoding");
        } else if (marker == 0xDD) {  // src/parser.rs:592-600
            size_t length;
            TRY(read_length(&length));
            if (length != 2) return fail(B200JPG_ERR_FORMAT, "DRI with invalid length");
            TRY(read_u16(&restart_interval_));
        } else if (marker == 0xFE) {
            size_t length;
            TRY(read_length(&length));
            if (pos_ + length > len_) {
                pos_ = len_;
                return fail(B200JPG_ERR_IO, "failed to fill whole buffer");
            }
            pos_ += length;
        } else if (is_app(marker)) {
            TRY(parse_app(marker));
        } else if (is_rst(marker)) {
            if (previous_marker != 0xDA) return fail(B200JPG_ERR_FORMAT, "RST found outside of entropy-coded data");
        } else if (marker == 0xDC) {
            if (previous_marker != 0xDA || scans_processed != 1)
                return fail(B200JPG_ERR_FORMAT, "DNL is only allowed immediately after the first scan");
            return fail(B200JPG_ERR_UNSUPPORTED, "DNL");
        } else if (marker == 0xDE || marker == 0xDF) {
            return fail(B200JPG_ERR_UNSUPPORTED, "Hierarchical");
        } else if (marker == 0xD9) {
            break;
        } else {
            return fail(B200JPG_ERR_FORMAT, "marker 0x%02X found where not allowed", marker);
        }
        previous_marker = marker;
    }
    if (!has_frame_) return fail(B200JPG_ERR_FORMAT, "end of image encountered before frame");
    // decode_planes, src/decoder.rs:631-684: size limit, then render unfinished progressive components
    if (buffer_limit_exceeded()) return fail(B200JPG_ERR_FORMAT, "size of decoded image exceeds maximum allowed size");
    if (frame_.coding_process == B200JPG_CP_DCT_PROGRESSIVE && has_work_)
        for (size_t i = 0; i < frame_.comps.size(); i++) {
            if (finished_mask_[i] == ~(uint64_t)0) continue;
            if (!has_qt_[frame_.comps[i].tq]) continue;
            memcpy(final_qt_[i], qt_[frame_.comps[i].tq], 128);
            if (ext_[i]) memcpy(ext_[i], work_[i].data(), work_[i].size() * sizeof(int16_t));
            else final_[i] = work_[i];
            have_final_[i] = true;
        }
    return finish_sbs();
}

bool HostDecoder::buffer_limit_exceeded() const {
    const size_t need = frame_.comps.size() * (size_t)frame_.output_w * frame_.output_h;
    return buffer_limit_ < need;
}

// src/decoder.rs:278-290, src/parser.rs:120-133
int HostDecoder::scale(uint16_t req_w, uint16_t req_h, uint16_t* w, uint16_t* h) {
    TRY(read_info());
    const int idct = b200jpg_choose_idct_size(frame_.image_w, frame_.image_h, req_w, req_h);
    for (auto& c : frame_.comps) c.dct_scale = (uint16_t)idct;
    if (b200jpg_update_component_sizes(frame_.image_w, frame_.image_h, frame_.comps.data(), (int)frame_.comps.size(), &frame_.mcu_w, &frame_.mcu_h))
        return fail(B200JPG_ERR_FORMAT, "invalid dimensions");
    const float fw = (float)frame_.image_w * (float)idct / 8.0f, fh = (float)frame_.image_h * (float)idct / 8.0f;
    uint16_t ow = (uint16_t)fw, oh = (uint16_t)fh;  // .ceil() as u16
    if ((float)ow < fw) ow++;
    if ((float)oh < fh) oh++;
    frame_.output_w = ow;
    frame_.output_h = oh;
    *w = ow;
    *h = oh;
    return 0;
}

int HostDecoder::determine_color_transform() const {
    if (has_ct_) return ct_;
    const auto& c = frame_.comps;
    if (c.size() == 1) return B200JPG_CT_GRAYSCALE;
    if (c.size() == 3) {
        const uint8_t a = c[0].identifier, b = c[1].identifier, d = c[2].identifier;
        if (a == 1 && b == 2 && d == 3) return B200JPG_CT_YCBCR;
        if (a == 1 && b == 34 && d == 35) return B200JPG_CT_JCS_BG_YCC;
        if (a == 82 && b == 71 && d == 66) return B200JPG_CT_RGB;
        if (a == 114 && b == 103 && d == 98) return B200JPG_CT_JCS_BG_RGB;
        if (is_jfif_) return B200JPG_CT_YCBCR;
    }
    if (has_adobe_) {
        if (adobe_ == 0) {
            if (c.size() == 3) return B200JPG_CT_RGB;
            if (c.size() == 4) return B200JPG_CT_CMYK;
        } else if (adobe_ == 1) {
            return B200JPG_CT_YCBCR;
        } else {
            return B200JPG_CT_YCCK;
        }
    } else if (c.size() == 4) {
        return B200JPG_CT_CMYK;
    }
    if (c.size() == 4) return B200JPG_CT_YCCK;
    if (c.size() == 3) return B200JPG_CT_YCBCR;
    return B200JPG_CT_UNKNOWN;
}

int HostDecoder::pixel_format() const {  // src/decoder.rs:174-183
    const size_t n = frame_.comps.size();
    if (n == 1) return frame_.precision <= 8 ? B200JPG_PF_L8 : B200JPG_PF_L16;
    return n == 3 ? B200JPG_PF_RGB24 : B200JPG_PF_CMYK32;
}

bool HostDecoder::icc_profile(std::vector<uint8_t>* out) const {
    const size_t num = icc_.size();
    if (num == 0 || num >= 255) return false;
    const IccChunk* present[256] = {nullptr};
    for (const auto& c : icc_) {
        if (c.num_markers != num || c.seq_no == 0 || present[c.seq_no]) return false;
        present[c.seq_no] = &c;
    }
    out->clear();
    for (size_t s = 1; s <= num; s++) {
        if (!present[s]) return false;
        out->insert(out->end(), present[s]->data.begin(), present[s]->data.end());
    }
    return true;
}

}  // namespace b200jpg
