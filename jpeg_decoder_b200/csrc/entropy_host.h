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

inline size_t ent_payload_bound(size_t file_len) { return sizeof(EntHeader) + ENT_MAX_SLOTS * sizeof(EntTables) + file_len + 64; }

// Writes the payload of a qualifying scan (HostDecoder::device_scan()) to dst.  Returns its length, or 0 when the
// entropy-coded segment turns out not to qualify: anything but stuffed bytes up to an EOI marker (restart or
// other markers, fill bytes, a truncated file) stays with the host decoder, which mirrors the reference there.
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
    h.data_off = (uint32_t)(sizeof(EntHeader) + nslots * sizeof(EntTables));
    h.total_blocks = (uint32_t)hd.total_blocks();
    // unstuff: FF 00 -> FF; FF D9 ends the scan; anything else disqualifies
    const uint8_t* s = file + ds.scan_begin;
    const uint8_t* const end = file + file_len;
    uint8_t* const data = dst + h.data_off;
    uint8_t* o = data;
    bool done = false;
    while (s < end) {
        const uint8_t* f = (const uint8_t*)memchr(s, 0xFF, (size_t)(end - s));
        if (!f || f + 1 >= end) return 0;
        memcpy(o, s, (size_t)(f - s));
        o += f - s;
        if (f[1] == 0x00) {
            *o++ = 0xFF;
            s = f + 2;
        } else if (f[1] == 0xD9) {
            done = true;
            break;
        } else {
            return 0;
        }
    }
    if (!done || o == data || (size_t)(o - data) >= ((size_t)1 << 28)) return 0;
    h.scan_bytes = (uint32_t)(o - data);
    size_t total = h.data_off + h.scan_bytes;
    const size_t padded = (total + 15) / 16 * 16 + 16;
    memset(dst + total, 0, padded - total);
    h.payload_len = (uint32_t)padded;
    memcpy(dst, &h, sizeof h);
    return padded;
}

// Kernel descriptor of one image from its payload header and geometry.  coef_off[c] = byte offset of component
// c's blocks inside the coefficient slab (multiple of 128), sub0 = index of its first subsequence in the group's
// state arrays.  Returns false when header and geometry disagree.
inline bool ent_fill_image(const EntHeader& h, const b200jpg_image_desc& d, const size_t coef_off[4], unsigned long long payload_off,
                           unsigned sub0, EntImage* im) {
    memset(im, 0, sizeof *im);
    if (h.magic != ENT_MAGIC || h.nslots == 0 || h.nslots > ENT_MAX_SLOTS || h.bpm == 0 || h.bpm > 12 || h.scan_bytes == 0) return false;
    if (h.data_off % 16 != 0 || h.payload_len % 16 != 0 || (size_t)h.data_off + h.scan_bytes + 16 > h.payload_len) return false;
    if (h.tables_off + h.nslots * sizeof(EntTables) > h.data_off || h.tables_off % 16 != 0) return false;
    im->payload_off = payload_off;
    im->data_off = h.data_off;
    im->scan_bits = h.scan_bytes * 8u;
    im->nwords = (h.payload_len - h.data_off) / 4u;
    im->nsub = (im->scan_bits + ENT_SUB_BITS - 1) / ENT_SUB_BITS;
    im->sub0 = sub0;
    im->nslots = h.nslots;
    im->tables_off = h.tables_off;
    im->bpm = h.bpm;
    im->ncomp = d.ncomp;
    unsigned j = 0, nb = 0;
    for (int c = 0; c < d.ncomp && c < 4; c++) {
        const b200jpg_component& k = d.comps[c];
        if (k.h == 0 || k.v == 0 || coef_off[c] % 128 != 0) return false;
        im->slab_row[c] = (unsigned)(coef_off[c] / 128);
        im->block_w[c] = k.block_w;
        im->comp_blocks[c] = (unsigned)k.block_w * k.block_h;
        im->h[c] = k.h;
        im->v[c] = k.v;
        nb += im->comp_blocks[c];
        for (unsigned vy = 0; vy < k.v; vy++)
            for (unsigned hx = 0; hx < k.h; hx++, j++)
                if (j < 12) {
                    im->mcu_comp[j] = (unsigned char)c;
                    im->mcu_hx[j] = (unsigned char)hx;
                    im->mcu_vy[j] = (unsigned char)vy;
                }
    }
    if (d.ncomp == 1) {  // a lone component is not interleaved: one block per MCU, raster order (src/decoder.rs:895-905)
        j = 1;
        im->h[0] = im->v[0] = 1;
    }
    if (j != h.bpm || nb != h.total_blocks || nb == 0) return false;
    im->total_blocks = nb;
    im->mcu_w = d.comps[0].block_w / im->h[0];
    if (im->mcu_w == 0) return false;
    bool uniform = true;
    for (unsigned q = 0; q < 12; q++) {
        im->dcslot[q] = h.dcslot[q] < h.nslots ? h.dcslot[q] : 0;
        im->acslot[q] = h.acslot[q] < h.nslots ? h.acslot[q] : 0;
        if (q < h.bpm && (im->dcslot[q] != im->dcslot[0] || im->acslot[q] != im->acslot[0])) uniform = false;
    }
    im->dec_bpm = uniform ? 1 : h.bpm;
    return true;
}

}  // namespace b200jpg
