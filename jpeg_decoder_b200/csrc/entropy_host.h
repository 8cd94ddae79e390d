// entropy_host.h -- host side of device entropy decoding (entropy_dev.h): decides whether a scan qualifies,
// builds the decoding tables from the file's DHT segments and copies the entropy-coded bytes, with the byte
// stuffing removed (src/huffman.rs:126-160 does that while reading), into page-locked memory for the upload.
// Host threads touch every scan byte once (memchr + memcpy); Huffman decoding itself happens on the device.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "entropy_dev.h"
#include "host_decoder.h"

namespace b200jpg {

// Decoding tables of one Huffman table.  `is_ac` selects how symbols are interpreted (run/size vs DC category).
// Returns false when the table has more long-code prefixes than ENT_MAX_SUBTABLES (such files stay on the host).
inline bool ent_build_tables(const HuffTable& t, bool is_ac, EntTables* out) {
    const uint16_t invalid = (uint16_t)ent_entry(1, 0, 0);  // no code word matches: anomaly, skip one bit
    for (auto& e : out->lut) e = invalid;
    for (auto& row : out->sub)
        for (auto& e : row) e = invalid;
    unsigned nsub = 0;
    int idx = 0;
    for (int len = 1; len <= 16; len++) {
        if (t.maxcode[len - 1] < 0) continue;
        const int first = idx - t.delta[len - 1];  // first code of this length
        const int count = t.maxcode[len - 1] - first + 1;
        for (int c = 0; c < count && idx < t.nvalues; c++, idx++) {
            const unsigned sym = t.values[idx];
            unsigned s, adv;
            if (!is_ac) {  // DC: the symbol is the category = number of value bits (src/decoder.rs:1096-1110)
                s = sym & 15u;
                adv = sym <= 11 ? 1 : 0;
            } else {
                const unsigned r = sym >> 4;
                s = sym & 15u;
                if (s) adv = r + 1;
                else adv = r == 0 ? 64 : (r == 15 ? 16 : 0);  // EOB, ZRL; EOBn does not occur in sequential scans
            }
            const unsigned code = (unsigned)(first + c);
            const uint16_t entry = (uint16_t)ent_entry((unsigned)len, s, adv);
            if (len <= (int)ENT_LUT_BITS) {
                const unsigned rem = ENT_LUT_BITS - (unsigned)len;
                for (unsigned f = 0; f < (1u << rem); f++) out->lut[(code << rem) + f] = entry;
            } else {
                const unsigned prefix = code >> ((unsigned)len - ENT_LUT_BITS);
                uint16_t& l = out->lut[prefix];
                if ((l & 31u) != ENT_LINK) {
                    if (nsub == ENT_MAX_SUBTABLES) return false;
                    l = (uint16_t)(ENT_LINK | (nsub++ << 5));
                }
                const unsigned rem = 16 - (unsigned)len;
                const unsigned low = (code << rem) & ((1u << ENT_SUB_LUT_BITS) - 1u);
                for (unsigned f = 0; f < (1u << rem); f++) out->sub[l >> 5][low + f] = entry;
            }
        }
    }
    return true;
}

// Un-stuffs entropy-coded bytes from s towards *o: plain bytes are copied, FF 00 becomes FF.  Stops at the first FF that
// is followed by anything else (a marker) and returns where it is; *o has advanced by the bytes written (the FF of the
// marker not included).  Returns nullptr when the input ends without a marker or the destination would overflow.
// The bulk of a host thread's time per image is this loop (0.6 MB per 1080p file, one FF per ~256 bytes), and with a
// few CPUs per GPU it bounds the whole-file path: 32 bytes per iteration where AVX2 exists -- store the vector, and only
// when it holds an FF look at the byte behind the first one and restart right after the pair.
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) inline const uint8_t* ent_unstuff_avx2(const uint8_t* s, const uint8_t* end, uint8_t** o, uint8_t* dst_end, bool* marker) {
    uint8_t* out = *o;
    const __m256i ff = _mm256_set1_epi8((char)0xFF);
    *marker = false;
    while (s + 33 <= end && out + 96 <= dst_end) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)s);
        _mm256_storeu_si256((__m256i*)out, v);   // whatever lies behind an FF is overwritten by the next store
        const unsigned m = (unsigned)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, ff));
        if (m == 0) {
            s += 32;
            out += 32;
            continue;
        }
        const unsigned k = (unsigned)__builtin_ctz(m);
        if (s[k + 1] == 0x00) {  // stuffed byte: keep the FF, drop the 00
            s += k + 2;
            out += k + 1;
            continue;
        }
        *o = out + k;
        *marker = true;
        return s + k;
    }
    *o = out;
    return s;   // the last few bytes (or a nearly full destination) are the scalar loop's
}
// AVX-512 VBMI2 (Ice Lake / Sapphire Rapids and later): no data-dependent control flow at all.  Per 64 bytes: the bytes
// that follow an FF (mask shifted by one, carried across vectors), those of them that are 00 are dropped with one
// VPCOMPRESSB, and any that is NOT 00 -- a marker -- ends the vector loop before its vector is consumed.
__attribute__((target("avx512f,avx512bw,avx512vbmi2,popcnt"))) inline const uint8_t* ent_unstuff_vbmi2(const uint8_t* s, const uint8_t* end, uint8_t** o,
                                                                                                  uint8_t* dst_end) {
    uint8_t* out = *o;
    const __m512i ff = _mm512_set1_epi8((char)0xFF);
    unsigned long long carry = 0;  // the previous vector ended with an FF (already written)
    while (s + 64 <= end && out + 192 <= dst_end) {
        const __m512i v = _mm512_loadu_si512((const void*)s);
        const unsigned long long ffm = _mm512_cmpeq_epi8_mask(v, ff), zm = _mm512_testn_epi8_mask(v, v);
        const unsigned long long after = (ffm << 1) | carry;
        if (after & ~zm) break;  // FF followed by something else than 00: the scalar loop looks at it
        const unsigned long long drop = after & zm;
        _mm512_storeu_si512((void*)out, _mm512_maskz_compress_epi8(~drop, v));
        out += 64 - (unsigned)_mm_popcnt_u64(drop);
        carry = ffm >> 63;
        s += 64;
    }
    if (carry) {  // hand the trailing FF back: the scalar loop wants to see the pair
        s -= 1;
        out -= 1;
    }
    *o = out;
    return s;
}
#define ENT_HAVE_AVX2_PATH 1
#endif
// 0 = scalar (memchr + memcpy), 1 = AVX2, 2 = AVX-512 VBMI2; -1 = the best the CPU has.  The override exists for the tests.
inline int& ent_unstuff_level() {
    static int level = -1;
    return level;
}
inline const uint8_t* ent_unstuff(const uint8_t* s, const uint8_t* end, uint8_t** o, uint8_t* dst_end) {
#if defined(ENT_HAVE_AVX2_PATH)
    static const int best = __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("avx512bw") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
    const int level = ent_unstuff_level() < 0 ? best : (ent_unstuff_level() < best ? ent_unstuff_level() : best);
    if (level == 2) {
        s = ent_unstuff_vbmi2(s, end, o, dst_end);
    } else if (level == 1) {
        bool marker = false;
        s = ent_unstuff_avx2(s, end, o, dst_end, &marker);
        if (marker) return s;
    }
#endif
    while (s < end) {
        const uint8_t* f = (const uint8_t*)memchr(s, 0xFF, (size_t)(end - s));
        const size_t n = (size_t)((f ? f : end) - s);
        if ((size_t)(dst_end - *o) < n + 64) return nullptr;
        memcpy(*o, s, n);
        *o += n;
        if (!f || f + 1 >= end) return nullptr;
        if (f[1] != 0x00) return f;
        *(*o)++ = 0xFF;
        s = f + 2;
    }
    return nullptr;
}

inline size_t ent_payload_bound(size_t file_len) {
    return sizeof(EntHeader) + ENT_MAX_SLOTS * sizeof(EntTables) + file_len + file_len / 4 + 4096;
}
constexpr size_t ENT_MIN_INTERVAL_BYTES = 256;  // scans cut into smaller pieces than this (on average) stay on the host

// Writes the payload of a qualifying scan (HostDecoder::device_scan()) to dst.  Returns its length, or 0 when the
// entropy-coded segment turns out not to qualify: anything but stuffed bytes and the expected restart markers up to
// an EOI marker (other markers, fill bytes, a truncated file, a wrong number or order of RSTn) stays with the host
// decoder, which mirrors the reference there.
inline size_t ent_build_payload(const HostDecoder& hd, const uint8_t* file, size_t file_len, uint8_t* dst, size_t cap) {
    const DeviceScan& ds = hd.device_scan();
    if (!ds.eligible || cap < ent_payload_bound(file_len) || ds.scan_begin > file_len) return 0;
    const FrameInfo& fr = hd.frame();
    EntHeader h;
    memset(&h, 0, sizeof h);
    h.magic = ENT_MAGIC;
    // table slots: distinct (class, id) pairs in order of first use
    int slot_dc[4] = {-1, -1, -1, -1}, slot_ac[4] = {-1, -1, -1, -1};
    EntTables* tabs = (EntTables*)(dst + sizeof(EntHeader));
    unsigned nslots = 0, j = 0;
    for (int i = 0; i < ds.scan.n; i++) {
        const int d = ds.scan.dc_table[i], a = ds.scan.ac_table[i];
        if (slot_dc[d] < 0) {
            if (nslots == ENT_MAX_SLOTS) return 0;
            if (!ent_build_tables(hd.dc_table(d), false, &tabs[nslots])) return 0;
            slot_dc[d] = (int)nslots++;
        }
        if (slot_ac[a] < 0) {
            if (nslots == ENT_MAX_SLOTS) return 0;
            if (!ent_build_tables(hd.ac_table(a), true, &tabs[nslots])) return 0;
            slot_ac[a] = (int)nslots++;
        }
        const b200jpg_component& c = fr.comps[(size_t)ds.scan.comp_index[i]];
        for (unsigned q = 0; q < (unsigned)c.h * c.v; q++) {
            if (j >= 12) return 0;
            h.dcslot[j] = (uint8_t)slot_dc[d];
            h.acslot[j] = (uint8_t)slot_ac[a];
            j++;
        }
    }
    h.bpm = (uint8_t)j;
    h.nslots = nslots;
    h.tables_off = (uint32_t)sizeof(EntHeader);
    h.total_blocks = (uint32_t)hd.total_blocks();
    // intervals: ceil(MCUs / restart interval) of them, separated by RST0..7 in cyclic order (src/decoder.rs:910-931)
    const size_t total_mcus = h.total_blocks / h.bpm;
    h.restart_interval = ds.restart_interval;
    const size_t nint = ds.restart_interval ? (total_mcus + ds.restart_interval - 1) / ds.restart_interval : 1;
    if (nint == 0 || nint > ((size_t)1 << 24)) return 0;
    h.nintervals = (uint32_t)nint;
    h.intervals_off = (uint32_t)(sizeof(EntHeader) + nslots * sizeof(EntTables));
    size_t at = (h.intervals_off + nint * sizeof(EntInterval) + 15) / 16 * 16;
    if (at + 64 > cap) return 0;
    EntInterval* iv = (EntInterval*)(dst + h.intervals_off);
    h.data_off = (uint32_t)at;
    // unstuff: FF 00 -> FF; FF D0+n closes an interval; FF D9 ends the scan; anything else disqualifies
    const uint8_t* s = file + ds.scan_begin;
    const uint8_t* const end = file + file_len;
    uint8_t* o = dst + at;
    size_t cur = 0, total = 0;
    bool done = false;
    auto close_interval = [&]() -> bool {  // records interval `cur`, pads it, opens the next one
        const size_t nbytes = (size_t)(o - (dst + at));
        iv[cur].data_off = (uint32_t)at;
        iv[cur].nbytes = (uint32_t)nbytes;
        total += nbytes;
        const size_t next = (at + nbytes + 15) / 16 * 16 + 16;
        if (next + 64 > cap) return false;
        memset(o, 0, next - (at + nbytes));
        at = next;
        o = dst + at;
        cur++;
        return true;
    };
    uint8_t* const dst_end = dst + cap;
    while (s < end) {
        const uint8_t* f = ent_unstuff(s, end, &o, dst_end);   // f = an FF followed by a marker byte
        if (!f || f + 1 >= end) return 0;
        const uint8_t m = f[1];
        if (m == 0xD9) {
            if (cur + 1 != nint || !close_interval()) return 0;
            done = true;
            break;
        } else if (m >= 0xD0 && m <= 0xD7 && ds.restart_interval && cur + 1 < nint && (unsigned)(m - 0xD0) == (cur & 7u)) {
            if (!close_interval()) return 0;
            s = f + 2;
        } else {
            return 0;
        }
    }
    if (!done || total == 0 || total >= ((size_t)1 << 28) || total / nint < (nint > 1 ? ENT_MIN_INTERVAL_BYTES : 1)) return 0;
    h.scan_bytes = (uint32_t)total;
    h.payload_len = (uint32_t)at;  // `at` is already padded and 16-byte aligned
    memcpy(dst, &h, sizeof h);
    return at;
}

// Bytes of the compact stream (entropy_dev.h) the write pass produces for an image of `nblocks` blocks.
inline size_t ent_cs_bytes(size_t nblocks) {
    const size_t nb_pad = (nblocks + 31) / 32 * 32;
    return (ent_cs_header_bytes(nb_pad) + 15) / 16 * 16 + (ent_cs_values_bytes(nblocks) + 15) / 16 * 16;
}

// Kernel descriptors of one image -- one per interval -- from its payload and geometry.  coef_off[c] = byte offset of
// component c's blocks inside the coefficient slab (multiple of 128; only checked: K0 is what writes the slab),
// cs_off = where the image's compact stream goes inside the device stream buffer (ent_cs_bytes(), 16-byte aligned),
// sub0 = index of the first subsequence in the group's state arrays.  Returns the number of descriptors written (header.nintervals), 0 when payload and geometry
// disagree or `cap` is too small; *nsub_total = subsequences over all intervals.
inline unsigned ent_fill_images(const uint8_t* payload, size_t payload_len, const b200jpg_image_desc& d, const size_t coef_off[4],
                                unsigned long long payload_off, unsigned long long cs_off, unsigned sub0, EntImage* out, size_t cap,
                                unsigned* nsub_total) {
    EntHeader h;
    if (payload_len < sizeof h) return 0;
    memcpy(&h, payload, sizeof h);
    if (h.magic != ENT_MAGIC || h.payload_len != payload_len || h.nslots == 0 || h.nslots > ENT_MAX_SLOTS || h.bpm == 0 || h.bpm > 12) return 0;
    if (h.payload_len % 16 != 0 || h.tables_off % 16 != 0 || h.tables_off + h.nslots * sizeof(EntTables) > h.intervals_off) return 0;
    if (h.nintervals == 0 || h.nintervals > cap || (size_t)h.intervals_off + (size_t)h.nintervals * sizeof(EntInterval) > h.payload_len) return 0;
    EntImage base;
    memset(&base, 0, sizeof base);
    base.payload_off = payload_off;
    base.nslots = h.nslots;
    base.tables_off = h.tables_off;
    base.bpm = h.bpm;
    base.ncomp = d.ncomp;
    unsigned j = 0, nb = 0, hv[4] = {0, 0, 0, 0};
    for (int c = 0; c < d.ncomp && c < 4; c++) {
        const b200jpg_component& k = d.comps[c];
        if (k.h == 0 || k.v == 0 || coef_off[c] % 128 != 0) return 0;
        base.slab_row[c] = (unsigned)(coef_off[c] / 128);
        base.block_w[c] = k.block_w;
        base.h[c] = k.h;
        base.v[c] = k.v;
        hv[c] = (unsigned)k.h * k.v;
        nb += (unsigned)k.block_w * k.block_h;
        for (unsigned vy = 0; vy < k.v; vy++)
            for (unsigned hx = 0; hx < k.h; hx++, j++)
                if (j < 12) {
                    base.mcu_comp[j] = (unsigned char)c;
                    base.mcu_hx[j] = (unsigned char)hx;
                    base.mcu_vy[j] = (unsigned char)vy;
                }
    }
    if (d.ncomp == 1) {  // a lone component is not interleaved: one block per MCU, raster order (src/decoder.rs:895-905)
        j = 1;
        base.h[0] = base.v[0] = 1;
        hv[0] = 1;
    }
    if (j != h.bpm || nb != h.total_blocks || nb == 0) return 0;
    base.mcu_w = d.comps[0].block_w / base.h[0];
    if (base.mcu_w == 0 || cs_off % 16 != 0) return 0;
    base.nb_pad = (nb + 31u) / 32u * 32u;
    base.vals_off = (unsigned)((ent_cs_header_bytes(base.nb_pad) + 15) / 16 * 16);
    base.cs_off = cs_off;
    for (unsigned q = 12; q-- > 0;)
        if (q < h.bpm) base.comp_j0[base.mcu_comp[q]] = (unsigned char)q;  // descending: the first slot of each component wins
    bool uniform = true;
    for (unsigned q = 0; q < 12; q++) {
        base.dcslot[q] = h.dcslot[q] < h.nslots ? h.dcslot[q] : 0;
        base.acslot[q] = h.acslot[q] < h.nslots ? h.acslot[q] : 0;
        if (q < h.bpm && (base.dcslot[q] != base.dcslot[0] || base.acslot[q] != base.acslot[0])) uniform = false;
    }
    base.dec_bpm = uniform ? 1 : h.bpm;
    const unsigned total_mcus = nb / h.bpm;
    const unsigned ri = h.restart_interval ? h.restart_interval : total_mcus;
    if ((total_mcus + ri - 1) / ri != h.nintervals) return 0;
    const EntInterval* iv = (const EntInterval*)(payload + h.intervals_off);
    unsigned nsub = 0;
    for (unsigned r = 0; r < h.nintervals; r++) {
        EntImage& im = out[r];
        im = base;
        if (iv[r].data_off % 16 != 0 || (size_t)iv[r].data_off + iv[r].nbytes + 16 > h.payload_len || iv[r].nbytes >= (1u << 28)) return 0;
        const unsigned mcus = r + 1 < h.nintervals ? ri : total_mcus - r * ri;
        im.data_off = iv[r].data_off;
        im.scan_bits = iv[r].nbytes * 8u;
        im.nwords = (iv[r].nbytes + 3u) / 4u;  // the bytes after the interval are zero up to the next 16-byte boundary and beyond
        im.nsub = (im.scan_bits + ENT_SUB_BITS - 1) / ENT_SUB_BITS;
        if (im.nsub == 0) im.nsub = 1;  // an empty interval still has to deliver its blocks (from zero bits, like the reference)
        im.sub0 = sub0 + nsub;
        im.mcu0 = r * ri;
        im.tight_end = r + 1 < h.nintervals ? 1 : 0;
        im.total_blocks = mcus * h.bpm;
        for (int c = 0; c < d.ncomp && c < 4; c++) im.comp_blocks[c] = mcus * hv[c];
        nsub += im.nsub;
    }
    *nsub_total = nsub;
    return h.nintervals;
}

}  // namespace b200jpg
