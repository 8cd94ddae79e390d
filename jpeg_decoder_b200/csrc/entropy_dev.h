// entropy_dev.h -- Huffman entropy decoding of baseline scans ON THE DEVICE ("the step
// before the path").  The reference decodes a scan with one sequential loop per image (src/decoder.rs:794-1172,
// src/huffman.rs:20-161); the position of every code word depends on all code words before it.  What makes the
// stream parallel anyway is that Huffman codes self-synchronise: a decoder started at a wrong bit position falls
// into step with the true code-word boundaries after a few dozen symbols.  So the scan is cut into fixed
// subsequences of ENT_SUB_BITS bits, one thread each:
//
//   cold  : every thread decodes its subsequence from the guess (bit 0 of the subsequence, block start) and
//           publishes where it ended: state = (bit position p, zig-zag index k, block-in-MCU b, blocks completed)
//   sync  : thread i re-decodes subsequence i from the state thread i-1 published; repeated while any published
//           state still changes (only threads whose predecessor changed do work).  Subsequence 0 starts from the
//           true state, so a pass without changes means state[i] = f_i(state[i-1]) for every i: the chain IS the
//           sequential decode
//   prefix: exclusive sums of "blocks completed" and "AC values met" -> the block every subsequence starts in and
//           where its values go
//   write : every thread decodes once more from its predecessor's state and appends what it meets to a compact
//           per-image stream -- non-zero AC values back to back (aligned 8-byte stores of four values), one 64-bit
//           position bitmap, one DC difference and one value offset per block; it re-checks
//           state[i] = f_i(state[i-1]), so a stream that did not converge (or is malformed in any way) is detected,
//           never mis-decoded
//   dc    : DC differences -> DC values, a wrapping int16 prefix sum per component in scan order
//   expand: kernel K0 (k0_expand.cu, the one that expands the host's sparse block streams) rebuilds the dense slab
//           K1 reads from bitmaps + values with coalesced 16-byte stores, zeros included
//
// Anything the device flags goes back to the host decoder, which reproduces the reference's behaviour for
// broken streams exactly.  This header is shared by the kernels (ke_entropy.cu), the host code that prepares
// the payload (entropy_host.h) and a CPU emulation used by the tests (tests/cpp/entropy_emul.cpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ENT_HD __host__ __device__ __forceinline__
#else
#define ENT_HD inline
#endif

namespace b200jpg {

#ifndef B200JPG_ENT_SUB_BITS
#define B200JPG_ENT_SUB_BITS 1024
#endif
constexpr unsigned ENT_SUB_BITS = B200JPG_ENT_SUB_BITS;  // one thread's share of the scan (a build-time knob: profiles/r02_entropy_subbits.md)
constexpr unsigned ENT_LUT_BITS = 9;               // code words up to this length resolve with one table probe
constexpr unsigned ENT_SUB_LUT_BITS = 16 - ENT_LUT_BITS;  // longer ones with a second probe, indexed by the remaining bits
constexpr unsigned ENT_MAX_SUBTABLES = 12;         // ENT_LUT_BITS-bit prefixes that continue into longer code words, per table
constexpr unsigned ENT_MAX_SLOTS = 4;              // Huffman tables per image (baseline: 2 DC + 2 AC)
constexpr unsigned ENT_TAIL_SLACK_BITS = 2048;     // zero bits the last subsequence may read past the end of the scan
constexpr unsigned ENT_MAGIC = 0x544e4542u;        // "BENT"

// Anomalies: reasons the device result of an image is discarded and the image is decoded on the host instead.
enum : unsigned {
    ENT_BAD_SYMBOL = 1,    // no code word matches / DC category > 11 / EOBn in a sequential scan
    ENT_BAD_RUN = 2,       // a run that leaves the block (src/decoder.rs:1149-1153 ends the block silently)
    ENT_BAD_CHAIN = 4,     // state[i] != f_i(state[i-1]): the synchronisation passes did not converge
    ENT_INCOMPLETE = 8,    // the scan ended before every block was decoded
    ENT_BAD_TAIL = 32,     // whole bytes left between the last MCU of a restart interval and its RSTn marker
};

// One table entry: bits 0..4 how many bits the code word and its value bits take together (<= 16 + 11), bits 5..8 the
// number of value bits, bits 9..15 how far the zig-zag index advances (run + 1; 64 = end of block).  The decode loop
// of the counting passes then needs one mask for the bit position and one shift for the index, nothing else.
// Anomalies (no code word matches, DC category > 11, EOBn in a sequential scan: adv = 0 here) are stored as "value bits =
// 15, advance 1, the code length alone": only the write pass looks for the 15 -- it alone decodes the true code words.
// Total 31 marks a first-level entry whose prefix continues into longer code words: bits 5..15 then hold the index of
// the second-level table.
constexpr uint32_t ENT_S_ANOMALY = 15;
ENT_HD uint32_t ent_entry(unsigned len, unsigned s, unsigned adv) {
    return adv ? (len + s) | (s << 5) | (adv << 9) : len | (ENT_S_ANOMALY << 5) | (1u << 9);
}
constexpr uint32_t ENT_LINK = 31;

// Decoding tables of one Huffman table ("slot"): two levels, so that every code word resolves with at most two
// shared-memory probes whatever its length (a search over code lengths would make the whole warp wait for the one
// lane that met a rare symbol).  Built on the host, copied to shared memory by the kernels.
struct EntTables {
    uint16_t lut[1u << ENT_LUT_BITS];                          // indexed by the next ENT_LUT_BITS bits
    uint16_t sub[ENT_MAX_SUBTABLES][1u << ENT_SUB_LUT_BITS];   // indexed by the ENT_SUB_LUT_BITS bits after those
};
constexpr unsigned ENT_TABLE_U16 = (unsigned)(sizeof(EntTables) / 2);
static_assert(sizeof(EntTables) == 4096, "EntTables layout");

// What the host writes in front of the unstuffed scan bytes.  All offsets are relative to the payload start.
// A scan with restart intervals (DRI, src/decoder.rs:910-931) is a sequence of independently decodable pieces: every
// interval starts byte-aligned with fresh predictors, so each one is handed to the kernels as a scan of its own
// (a free synchronisation point); without DRI there is one interval, the whole scan.
struct EntInterval {
    uint32_t data_off;  // 16-byte aligned; followed by >= 16 zero bytes
    uint32_t nbytes;    // unstuffed entropy-coded bytes of the interval
};
struct alignas(16) EntHeader {
    uint32_t magic;
    uint32_t scan_bytes;    // unstuffed entropy-coded bytes of all intervals
    uint32_t data_off;      // first interval's bytes
    uint32_t payload_len;   // multiple of 16
    uint32_t total_blocks;  // blocks the scan must deliver
    uint32_t nslots;
    uint8_t bpm;            // blocks per MCU
    uint8_t pad_[3];
    uint8_t dcslot[12], acslot[12];  // table slot of MCU block j
    uint32_t tables_off;    // EntTables[nslots]
    uint32_t nintervals;    // >= 1
    uint32_t intervals_off; // EntInterval[nintervals]
    uint32_t restart_interval;  // MCUs per interval; 0 = no DRI
    uint32_t reserved_[3];
};
static_assert(sizeof(EntHeader) == 80, "EntHeader layout");

// Per-interval descriptor of the kernels (built by the submitter from the payload + the image geometry); "image" in
// the kernels' vocabulary, since an interval is decoded exactly like a small scan.
struct alignas(16) EntImage {
    unsigned long long payload_off;  // byte offset of the payload inside the device stream buffer
    unsigned data_off, scan_bits, nwords, nsub, sub0, total_blocks, mcu_w, nslots, tables_off;
    unsigned slab_row[4];   // first 128-byte row of each component inside the coefficient slab
    unsigned block_w[4];    // blocks per block row
    unsigned comp_blocks[4];  // blocks of each component inside this interval, 0 for absent components
    unsigned char bpm, ncomp;
    unsigned char dec_bpm;  // period of the table pattern the decoder tracks as `b`: bpm, or 1 when every block of an MCU uses the
                            // same tables (b then never influences decoding and would only delay synchronisation)
    unsigned char tight_end;  // 1: the interval is followed by an RSTn marker, which the reference only finds when no whole byte is
                              // left after the interval's last MCU (take_marker looks no further than its bit buffer,
                              // src/huffman.rs:98-124); after the LAST interval extra bytes are skipped (src/decoder.rs:961-970)
    unsigned char h[4], v[4];
    unsigned char mcu_comp[12], mcu_hx[12], mcu_vy[12], dcslot[12], acslot[12];
    unsigned mcu0;    // index of the interval's first MCU in the scan
    unsigned nb_pad;  // blocks of the whole image rounded up to 32: length of the compact stream's per-block arrays
    unsigned vals_off;              // where the image's value area starts, relative to cs_off (bytes, even)
    unsigned long long cs_off;      // compact stream of the image inside the device stream buffer: bm | dc | boff
    unsigned char comp_j0[4];       // index of each component's first block inside an MCU
    unsigned pad2_[1];
};
static_assert(sizeof(EntImage) == 192, "EntImage layout");

// The compact stream the write pass produces, per image (all blocks in scan order, index t):
//   bm  : u64[nb_pad]  bit k (1..63) = the AC coefficient with zig-zag index k is non-zero; bit 0 = values are int16
//   dc  : i16[nb_pad]  DC difference, after the dc pass the DC value
//   boff: u32[nb_pad]  byte offset (relative to the start of bm) of the block's first value
//   values: int16, zig-zag order, at boff -- every restart interval appends to a region of its own (63 values per block
//           is the worst case), so intervals need no offsets from each other
// This is sbs.h's stream with per-block instead of per-group offsets (SBS_BLOCK_OFFSETS); K0 expands both.
ENT_HD size_t ent_cs_header_bytes(size_t nb_pad) { return 14 * nb_pad; }
ENT_HD size_t ent_cs_values_bytes(size_t nblocks) { return nblocks * 126 + 64; }

// Published per subsequence.  Two states are "equal" for synchronisation purposes when p, k and b agree.
struct EntState {
    uint32_t p;   // bit position of the next code word (relative to the scan start)
    uint32_t k;   // zig-zag index the next code word writes (0 = the next code word is a DC code)
    uint32_t b;   // index of the current block inside its MCU
    uint32_t nb;  // blocks completed while decoding the subsequence
};
ENT_HD uint64_t ent_pack(const EntState& s) {
    return (uint64_t)s.p | ((uint64_t)(s.k & 127u) << 32) | ((uint64_t)(s.b & 31u) << 39) | ((uint64_t)(s.nb & 0xfffffu) << 44);
}
ENT_HD EntState ent_unpack(uint64_t v) {
    EntState s;
    s.p = (uint32_t)v;
    s.k = (uint32_t)(v >> 32) & 127u;
    s.b = (uint32_t)(v >> 39) & 31u;
    s.nb = (uint32_t)(v >> 44) & 0xfffffu;
    return s;
}
constexpr uint64_t ENT_SYNC_MASK = ((uint64_t)1 << 44) - 1;  // p, k, b

ENT_HD uint32_t ent_bswap(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0u, 0x0123u);
#else
    return __builtin_bswap32(w);
#endif
}
// Word sources of the decoder: get(i) = big-endian word i of the scan, zero beyond the end (the reference feeds zero
// bits once the data is exhausted, src/huffman.rs:126-160).  This one reads the payload where it lies; the kernels
// stage their CTA's share of the scan in shared memory first (ke_entropy.cu).
struct EntWordsGlobal {
    const uint32_t* w;
    uint32_t n;
    ENT_HD uint32_t get(uint32_t i) const { return i < n ? ent_bswap(w[i]) : 0u; }
};

#if defined(ENT_STATS)
static unsigned long long ent_stats_len[32];  // emulator only: code-length histogram
#endif

// Sinks: what happens to decoded coefficients.  kCountOnly: the decoder only reports how many values it would store.
// Counts what the write pass will append: AC values (DC differences have their own array).
struct EntCountSink {
    static constexpr bool kCountOnly = true;  // the decoder adds (value bits present && not a DC code word) itself: no branch
    uint32_t nvals = 0;
    ENT_HD void store(unsigned, int) {}
    ENT_HD bool block_done() { return false; }
};

template <class Sink>
ENT_HD void sink_count(Sink&, unsigned) {}
ENT_HD void sink_count(EntCountSink& s, unsigned n) { s.nvals += n; }

// Decodes code words starting at state `st` while they start before bit `end_bit`.  Returns the state after the
// last one (nb = blocks completed here).  The function is deterministic in (st, end_bit) whatever the bits are:
// that is all the synchronisation passes need from it.  tabs = EntTables[] viewed as uint16_t; CHECK = report
// anomalies (only the write pass decodes the true code words, only there they mean something).
template <bool CHECK, class Words, class Sink>
ENT_HD EntState ent_decode_range(const Words& words, const uint16_t* tabs, const uint8_t* dcslot, const uint8_t* acslot, unsigned bpm,
                                 EntState st, uint32_t end_bit, Sink& sink, unsigned* anomaly) {
    uint32_t p = st.p, k = st.k, b = st.b, nb = 0;
    uint32_t wi = p >> 5;
    uint32_t hi = words.get(wi), lo = words.get(wi + 1);
    uint32_t dc_off = dcslot[b] * ENT_TABLE_U16, ac_off = acslot[b] * ENT_TABLE_U16;  // change only at block boundaries
    unsigned bad = 0;
    while (p < end_bit) {
        const uint32_t sh = p & 31u;
#if defined(__CUDA_ARCH__)
        const uint32_t win = __funnelshift_l(lo, hi, sh);
#else
        const uint32_t win = sh ? (hi << sh) | (lo >> (32u - sh)) : hi;
#endif
        const uint32_t off = k == 0 ? dc_off : ac_off;
        uint32_t e = tabs[off + (win >> (32u - ENT_LUT_BITS))];
        if ((e & 31u) == ENT_LINK)
            e = tabs[off + (1u << ENT_LUT_BITS) + ((e >> 5) << ENT_SUB_LUT_BITS) + ((win >> 16) & ((1u << ENT_SUB_LUT_BITS) - 1u))];
        const uint32_t tot = e & 31u, adv = e >> 9;
        uint32_t s = (e >> 5) & 15u;
#if defined(ENT_STATS)
        ent_stats_len[tot - (s == ENT_S_ANOMALY ? 0u : s)]++;
#endif
        if (Sink::kCountOnly) {
            // what the write pass will append for this code word: an AC value that lands inside the block (a run leaving
            // the block stores nothing and is flagged there; counting it would let value offsets run past the 63 values
            // per block the stream's value region is sized for)
            sink_count(sink, (s != 0u) & (k != 0u) & (k + adv <= 64u));
        } else {
            if (s == ENT_S_ANOMALY) {
                if (CHECK) bad |= ENT_BAD_SYMBOL;
                s = 0;
            }
            if (s) {
                const uint32_t pos = k + adv - 1;
                const uint32_t u = (win << (tot - s)) >> (32u - s);
                // extend(), src/huffman.rs:98-... / Figure F.12
                const int v = u < (1u << (s - 1)) ? (int)u - (int)(1u << s) + 1 : (int)u;
                if (pos <= 63u) sink.store(pos, v);
                else if (CHECK) bad |= ENT_BAD_RUN;
            }
        }
        p += tot;
        k += adv;
        const uint32_t nwi = p >> 5;
        if (nwi != wi) {  // tot <= 27: at most one word further
            hi = lo;
            lo = words.get(nwi + 1);
            wi = nwi;
        }
        if (k >= 64u) {
            k = 0;
            nb++;
            b = b + 1 == bpm ? 0 : b + 1;
            dc_off = dcslot[b] * ENT_TABLE_U16;
            ac_off = acslot[b] * ENT_TABLE_U16;
            if (sink.block_done()) break;
        }
    }
    if (CHECK && bad) *anomaly |= bad;
    EntState r;
    r.p = p;
    r.k = k;
    r.b = b;
    r.nb = nb;
    return r;
}

ENT_HD void ent_or64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

// The write pass.  A block that starts and ends inside one subsequence is written with plain stores; the two (or
// more) threads that share a block each OR their part of the bitmap into the zero-initialised array.  Values leave
// four at a time as one aligned 8-byte store (a lane's 2-byte stores each cost a sector of their own on the way to
// L2: measured 23 M sectors per 27 images).  They are shifted into a 64-bit window from the top -- two instructions
// per value, no variable shift -- so that after four values the oldest sits in the low 16 bits (little-endian order);
// a quad is due whenever the value's address is the last of an aligned 8 bytes.  The up to three values before the
// first and after the last aligned quad of a subsequence are stored one by one: the neighbouring subsequences own
// the rest of those words.
struct EntCompactSink {
    static constexpr bool kCountOnly = false;
    unsigned long long* bm;
    int16_t* dc;
    uint32_t* boff;
    int16_t* vals;        // the interval's value region
    uint32_t vals_rel;    // its byte offset relative to bm (even; the stream itself is 16-byte aligned)
    uint32_t t;           // scan-order index of the current block inside the image
    uint32_t B, total;    // blocks of the interval delivered so far / to deliver
    uint32_t vi, vcap;    // values appended so far / capacity of the region
    uint32_t vfirst;      // vi of this subsequence's first value
    uint32_t phase;       // ((vals_rel / 2 + vi) & 3) == 3  <=>  value vi ends an aligned quad; phase = (vals_rel / 2) & 3
    uint32_t acc_lo, acc_hi;  // the last four values, the newest in the top 16 bits
    unsigned long long bits;
    bool own;             // this thread met the block's DC code word
    ENT_HD void begin(uint8_t* streams, const EntImage* im, uint32_t first_block, uint32_t first_val, bool at_block_start) {
        uint8_t* cs = streams + im->cs_off;
        bm = (unsigned long long*)cs;
        dc = (int16_t*)(cs + 8 * (size_t)im->nb_pad);
        boff = (uint32_t*)(cs + 10 * (size_t)im->nb_pad);
        const uint32_t block0 = im->mcu0 * im->bpm;
        vals_rel = im->vals_off + 126u * block0;
        vals = (int16_t*)(cs + vals_rel);
        B = first_block;
        total = im->total_blocks;
        t = block0 + B;
        vi = vfirst = first_val;
        vcap = 63u * total;
        phase = (vals_rel >> 1) & 3u;
        bits = 0;
        acc_lo = acc_hi = 0;
        own = at_block_start;
        if (own && B < total) boff[t] = vals_rel + 2u * vi;
    }
    ENT_HD bool complete() const { return B >= total; }
    // value number `slot` (0 = oldest) of the window
    ENT_HD int16_t window(unsigned slot) const { return (int16_t)(slot < 2u ? acc_lo >> (16u * slot) : acc_hi >> (16u * (slot - 2u))); }
    // only reached with B < total: begin() is not used past the end, block_done() stops there
    ENT_HD void store(unsigned pos, int v) {
        if (pos == 0) {
            dc[t] = (int16_t)v;
            return;
        }
        bits |= 1ull << pos;
        acc_lo = (acc_lo >> 16) | (acc_hi << 16);
        acc_hi = (acc_hi >> 16) | ((uint32_t)(uint16_t)v << 16);
        if (((vi + phase) & 3u) == 3u) {  // value vi closes an aligned quad
            if (vi >= vfirst + 3u) {
                if (vi < vcap) *(unsigned long long*)(vals + vi - 3) = (unsigned long long)acc_lo | ((unsigned long long)acc_hi << 32);
            } else {  // head: the quad began in the previous subsequence; mine are its last vi - vfirst + 1 values
                const unsigned n = vi - vfirst + 1u;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (unsigned j = 0; j < n; j++)
                    if (vfirst + j < vcap) vals[vfirst + j] = window(4u - n + j);
            }
        }
        vi++;
    }
    ENT_HD bool block_done() {
        if (own) bm[t] = bits | 1ull;
        else if (bits) ent_or64(&bm[t], bits);
        B++;
        t++;
        bits = 0;
        own = true;
        if (B >= total) return true;
        boff[t] = vals_rel + 2u * vi;
        return false;
    }
    // after the decode loop: the values of the last, incomplete quad (the newest m of the window); the block still in
    // progress (it belongs to the next subsequence as well)
    ENT_HD void finish() {
        unsigned m = (vi + phase) & 3u;  // values behind the last aligned quad boundary
        if (m > vi - vfirst) m = vi - vfirst;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (unsigned j = 0; j < m; j++)
            if (vi - m + j < vcap) vals[vi - m + j] = window(4u - m + j);
        if (B < total && (own || bits)) ent_or64(&bm[t], bits | (own ? 1ull : 0ull));
    }
};

// where subsequence i of an image ends (the last one ends with the scan)
ENT_HD uint32_t ent_sub_end(uint32_t i, uint32_t nsub, uint32_t scan_bits) { return i + 1 < nsub ? (i + 1) * ENT_SUB_BITS : scan_bits; }

// zig-zag index -> natural position, src/decoder.rs:27-36 (the CPU emulation's stand-in for K0 uses it)
#define ENT_UNZIGZAG_INIT                                                                                                      \
    {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, \
     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}

}  // namespace b200jpg
