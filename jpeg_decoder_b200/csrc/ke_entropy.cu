// ke_entropy.cu -- the kernels of device entropy decoding; the algorithm is described in entropy_dev.h.
//
//   ent_pass<COLD | WRITE>, ent_sync  one thread per subsequence (ENT_SUB_BITS bits of the scan), 128 per CTA; the image's
//                                  decoding tables and its 16 KB of the scan live in shared memory; the write pass
//                                  appends to the image's compact stream, which K0 (k0_expand.cu) expands
//   ent_prefix                     per image: exclusive sums of the blocks each subsequence completed and the values it met
//   ent_dc_sums/_chunks/_apply     DC differences -> DC values (wrapping int16 prefix sum per component), one thread per block
//
// grid.y = image, grid.x covers the longest scan of the launch; CTAs past the end of their image exit at once.
#include <cuda_runtime.h>

#include "entropy_dev.h"
#include "kernels.h"

namespace b200jpg {

namespace {

constexpr int ENT_COLD = 0, ENT_WRITE = 2;
#ifndef B200JPG_ENT_THREADS
#define B200JPG_ENT_THREADS 256
#endif
// subsequences per CTA (build-time knob, with ENT_SUB_BITS).  Shared memory decides how many warps an SM holds: 16.6 KB of tables per CTA
// + 136 B of scan per thread -- 64 / 128 / 256 threads: 16 / 24 / 32 warps per SM, 63.9 / 68.3 / 72.8 GP/s (profiles/r02_entropy_sync_variants.md)
constexpr unsigned ENT_THREADS = B200JPG_ENT_THREADS;

struct EntShared {
    EntImage im;
    EntTables tabs[ENT_MAX_SLOTS];
    uint8_t dcslot[12], acslot[12];
};


// The CTA's share of the scan in shared memory, TRANSPOSED: tile[n * ENT_THREADS + t] = word n of the CTA's subsequence t,
// n = 0 .. SUB_WORDS + 1 (the last code word of a subsequence may reach two words into the next one: every column
// carries copies of them).  A thread walks down its own column, so whatever offsets the lanes of a warp are at, they
// hit 32 different banks, the next word is one row further and nothing has to be bounds-checked (the only reads
// past a column are the tail slack of a scan's last subsequence in the write pass, where everything beyond is zero:
// clamped to the column's last word, which is zero there too).  Staged once with coalesced 128-bit loads (every thread
// walking its own 128-byte line through L1 thrashes it: measured 5 % hit rate), byte-swapped on the way in.
constexpr unsigned SUB_WORDS = ENT_SUB_BITS / 32;
constexpr unsigned COL_WORDS = SUB_WORDS + 2;
constexpr unsigned TILE_WORDS = ENT_THREADS * SUB_WORDS;  // words of the scan a CTA owns
constexpr size_t TILE_BYTES = (size_t)COL_WORDS * ENT_THREADS * 4;
struct EntWordsColumn {
    const uint32_t* col;  // shared: word 0 of the subsequence
    uint32_t first;       // its index in the scan
    __device__ __forceinline__ uint32_t get(uint32_t i) const { return col[min(i - first, COL_WORDS - 1u) * ENT_THREADS]; }
};
__device__ __forceinline__ void stage_put(uint32_t* tile, uint32_t w, uint32_t v) {  // w: index inside the CTA's share (+ 2)
    const uint32_t t = w / SUB_WORDS, n = w % SUB_WORDS;
    v = ent_bswap(v);
    if (t < ENT_THREADS) tile[n * ENT_THREADS + t] = v;
    if (n < 2u && t > 0u) tile[(SUB_WORDS + n) * ENT_THREADS + t - 1u] = v;
}
__device__ __forceinline__ void stage_tile(uint32_t* tile, const uint32_t* words, uint32_t nwords, uint32_t first) {
    for (unsigned q = threadIdx.x; q < TILE_WORDS / 4 + 1; q += ENT_THREADS) {  // + 1: the two words behind the share
        const uint32_t i = first + 4 * q;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (i + 3 < nwords) v = __ldg(reinterpret_cast<const uint4*>(words + i));  // payloads are 16-byte aligned and padded
        else if (i < nwords) {
            v.x = __ldg(words + i);
            v.y = i + 1 < nwords ? __ldg(words + i + 1) : 0u;
            v.z = i + 2 < nwords ? __ldg(words + i + 2) : 0u;
        }
        stage_put(tile, 4 * q, v.x);
        stage_put(tile, 4 * q + 1, v.y);
        stage_put(tile, 4 * q + 2, v.z);
        stage_put(tile, 4 * q + 3, v.w);
    }
}

__device__ __forceinline__ void load_shared(EntShared& sh, const EntImage& im, const uint8_t* payload) {
    const uint4* src = reinterpret_cast<const uint4*>(payload + im.tables_off);
    uint4* dst = reinterpret_cast<uint4*>(sh.tabs);
    const unsigned n = im.nslots * (unsigned)(sizeof(EntTables) / 16);
    for (unsigned q = threadIdx.x; q < n; q += ENT_THREADS) dst[q] = __ldg(src + q);
    if (threadIdx.x < sizeof(EntImage) / 16) reinterpret_cast<uint4*>(&sh.im)[threadIdx.x] = __ldg(reinterpret_cast<const uint4*>(&im) + threadIdx.x);
    if (threadIdx.x < 12) {
        sh.dcslot[threadIdx.x] = im.dcslot[threadIdx.x];
        sh.acslot[threadIdx.x] = im.acslot[threadIdx.x];
    }
    __syncthreads();
}

// work arrays of one launch (one entry per subsequence of every image, images back to back)
struct EntWork {
    unsigned long long* state;
    unsigned* first_block;
    unsigned* nvals;      // AC values the subsequence appends to the compact stream
    unsigned* first_val;  // exclusive prefix of nvals inside the interval
    unsigned char* ch[2];
    unsigned* counters;  // counters[r] != 0: pass r changed some state (counters[0] is set by the cold pass)
    unsigned* status;    // per image: [anomaly bits, completed]
};

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) { return __ldcg(p); }  // L2: other SMs write these
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) { __stcg(p, v); }

constexpr unsigned ENT_LOCAL_ITERS = 32;  // synchronisation rounds a CTA runs by itself before the next launch takes over

// COLD and WRITE: one decode per thread.
template <int MODE>
__global__ void __launch_bounds__(ENT_THREADS) ent_pass(const EntImage* __restrict__ imgs, uint8_t* streams, EntWork w) {
    const EntImage& im = imgs[blockIdx.y];
    if (blockIdx.x * ENT_THREADS >= im.nsub) return;
    const unsigned i = blockIdx.x * ENT_THREADS + threadIdx.x;
    const bool valid = i < im.nsub;
    const unsigned g = im.sub0 + i;
    bool active = valid;
    if (MODE == ENT_WRITE) {
        active = valid && w.first_block[g] < im.total_blocks;
        if (!__syncthreads_or(active)) return;
    }
    __shared__ EntShared sh;
    extern __shared__ uint32_t tile[];  // COL_WORDS * ENT_THREADS words (dynamic: with the tables it passes 48 KB for 256-thread CTAs)
    const uint8_t* payload = streams + im.payload_off;
    const uint32_t* gwords = reinterpret_cast<const uint32_t*>(payload + im.data_off);
    stage_tile(tile, gwords, im.nwords, blockIdx.x * TILE_WORDS);
    load_shared(sh, im, payload);  // ends with a barrier
    if (!active) return;
    const EntWordsColumn words{tile + threadIdx.x, i * SUB_WORDS};
    const uint16_t* tabs = reinterpret_cast<const uint16_t*>(sh.tabs);
    EntState st;
    st.p = i * ENT_SUB_BITS;
    st.k = st.b = st.nb = 0;
    unsigned bad = 0;
    if (MODE == ENT_COLD) {
        EntCountSink sink;
        st_state(&w.state[g], ent_pack(ent_decode_range<false>(words, tabs, sh.dcslot, sh.acslot, im.dec_bpm, st,
                                                               ent_sub_end(i, im.nsub, im.scan_bits), sink, &bad)));
        w.nvals[g] = sink.nvals;
        w.ch[0][g] = 1;  // the first synchronisation launch looks at everybody
        if (i == 0) w.counters[0] = 1;
    } else {
        if (i > 0) {
            st = ent_unpack(ld_state(&w.state[g - 1]));
            st.nb = 0;
        }
        EntCompactSink sink;
        sink.begin(streams, &sh.im, w.first_block[g], w.first_val[g], st.k == 0);
        const bool last = i + 1 == im.nsub;
        const uint32_t end = last ? im.scan_bits + ENT_TAIL_SLACK_BITS : ent_sub_end(i, im.nsub, im.scan_bits);
        const EntState e = ent_decode_range<true>(words, tabs, sh.dcslot, sh.acslot, im.dec_bpm, st, end, sink, &bad);
        sink.finish();
        if (sink.complete()) {
            w.status[2 * blockIdx.y + 1] = 1;  // every block of the interval has been delivered
            if (im.tight_end && e.p + 7 < im.scan_bits) bad |= ENT_BAD_TAIL;
        }
        else if (last) bad |= ENT_INCOMPLETE;
        // the chain check covers what the prefix sums were built from: the end state AND the number of values -- a
        // successor that re-synchronises inside its subsequence ends in the same state with a different count, which
        // would shift every later value offset of the interval (possible when the sync passes ran out of budget)
        else if (ent_pack(e) != ld_state(&w.state[g]) || sink.vi - w.first_val[g] != w.nvals[g]) bad |= ENT_BAD_CHAIN;
        if (bad) atomicOr(&w.status[2 * blockIdx.y], bad);
    }
}

// SYNC launch number `pass` (1, 2, ...).  ch[] flags carry "my state changed and my successor has not seen it yet" from
// one launch to the next (read from ch[(pass-1)&1], written to ch[pass&1]).  Inside a launch a CTA keeps iterating on
// its own 128 subsequences through shared-memory flags until they are quiet; only the hand-over to the next CTA (and
// whatever is left when ENT_LOCAL_ITERS runs out) waits for the next launch.  Launch 1 re-decodes every subsequence;
// after that a quarter of them is pending, then 5 %, 1 %, ... for four to six rounds of one subsequence walk each
// (tests/cpp/entropy_emul.cpp with ENT_EMUL_ROUNDS=1).
// Measured and not kept (profiles/r02_entropy_sync_variants.md): the rounds in a kernel of their own that stages nothing
// (2 KB of shared memory per waiting CTA instead of 35, scan and tables through L1) -- 8 % slower under load, the L1 is
// not the rounds' alone; the same with the tables staged and only the scan words through L1 -- no difference; a warp per
// pending subsequence probing 32 bit positions at once -- no faster per code word than
// the lone lane, and three times the instructions.
template <class Words>
__device__ __forceinline__ void sync_redecode(const EntImage& im, const EntWork& w, const Words& words, const uint16_t* tabs, const uint8_t* dcslot,
                                              const uint8_t* acslot, unsigned ij, unsigned long long mine, unsigned long long* v_out, bool* changed) {
    const unsigned gj = im.sub0 + ij;
    EntState st = ent_unpack(ld_state(&w.state[gj - 1]));
    st.nb = 0;
    EntCountSink sink;
    unsigned bad = 0;
    const unsigned long long v =
        ent_pack(ent_decode_range<false>(words, tabs, dcslot, acslot, im.dec_bpm, st, ent_sub_end(ij, im.nsub, im.scan_bits), sink, &bad));
    *changed = ((v ^ mine) & ENT_SYNC_MASK) != 0;
    if (v != mine) st_state(&w.state[gj], v);  // the counts may change even when (p, k, b) do not
    w.nvals[gj] = sink.nvals;
    *v_out = v;
}

__global__ void __launch_bounds__(ENT_THREADS) ent_sync(const EntImage* __restrict__ imgs, const uint8_t* streams, EntWork w, unsigned pass) {
    if (w.counters[pass - 1] == 0) return;  // converged earlier: nothing left to do
    const EntImage& im = imgs[blockIdx.y];
    if (blockIdx.x * ENT_THREADS >= im.nsub) return;
    const unsigned i = blockIdx.x * ENT_THREADS + threadIdx.x;
    const bool valid = i < im.nsub;
    const unsigned g = im.sub0 + i;
    const unsigned char* ch_in = (pass & 1u) ? w.ch[0] : w.ch[1];
    unsigned char* ch_out = (pass & 1u) ? w.ch[1] : w.ch[0];
    bool pending = valid && i > 0 && ch_in[g - 1] != 0;  // my predecessor changed and I have not re-decoded since
    if (!__syncthreads_or(pending)) {
        if (valid) ch_out[g] = 0;
        return;
    }
    // per subsequence of this CTA (index = threadIdx of its owner): last published state, flags; and the compacted work list
    __shared__ unsigned long long s_mine[ENT_THREADS];
    __shared__ unsigned char s_pend[ENT_THREADS], s_chg[ENT_THREADS], s_ever[ENT_THREADS];
    __shared__ unsigned short s_list[ENT_THREADS];
    __shared__ unsigned s_n;
    const uint8_t* payload = streams + im.payload_off;
    const uint32_t* gwords = reinterpret_cast<const uint32_t*>(payload + im.data_off);
    __shared__ EntShared sh;
    extern __shared__ uint32_t tile[];  // COL_WORDS * ENT_THREADS words (dynamic: with the tables it passes 48 KB for 256-thread CTAs)
    const unsigned t = threadIdx.x;
    s_mine[t] = valid ? ld_state(&w.state[g]) : 0ull;
    s_pend[t] = pending ? 1 : 0;
    s_chg[t] = 0;
    s_ever[t] = 0;
    stage_tile(tile, gwords, im.nwords, blockIdx.x * TILE_WORDS);
    load_shared(sh, im, payload);  // ends with a barrier
    const uint16_t* tabs = reinterpret_cast<const uint16_t*>(sh.tabs);
    bool quiet = false;
    // Every round the pending subsequences are gathered onto the first threads of the CTA: a re-decode costs the same
    // issue slots whether 3 or 32 lanes of its warp take part.
    for (unsigned iter = 0; iter < ENT_LOCAL_ITERS; iter++) {
        if (t == 0) s_n = 0;
        __syncthreads();
        if (s_pend[t]) s_list[atomicAdd(&s_n, 1u)] = (unsigned short)t;
        s_chg[t] = 0;
        __syncthreads();
        if (t < s_n) {
            const unsigned j = s_list[t];  // the subsequence this thread re-decodes (any order: rounds are order-free)
            unsigned long long v;
            bool changed;
            const unsigned ij = blockIdx.x * ENT_THREADS + j;
            sync_redecode(im, w, EntWordsColumn{tile + j, ij * SUB_WORDS}, tabs, sh.dcslot, sh.acslot, ij, s_mine[j], &v, &changed);
            s_mine[j] = v;
            s_chg[j] = changed ? 1 : 0;
            if (changed) s_ever[j] = 1;
        }
        __syncthreads();  // flags and states of this round are visible to the CTA
        const bool next = valid && t > 0 && s_chg[t - 1] != 0;  // (threads past the image's end own no state)
        s_pend[t] = next ? 1 : 0;
        if (!__syncthreads_or(next)) {
            quiet = true;  // everything inside the CTA has been consumed
            break;
        }
    }
    // for the next launch: the successor of my last thread lives in another CTA; changes of the final round are unconsumed
    const bool flag = valid && ((!quiet && s_chg[t] != 0) || (t == ENT_THREADS - 1 && s_ever[t] != 0));
    if (valid) ch_out[g] = flag ? 1 : 0;
    if (flag) w.counters[pass] = 1;
}

constexpr unsigned SCAN_THREADS = 512;

// exclusive block-wide sum of one value per thread; returns the total through *total
__device__ __forceinline__ unsigned block_exclusive(unsigned v, unsigned* total) {
    __shared__ unsigned warp_sums[SCAN_THREADS / 32];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    unsigned x = v;
#pragma unroll
    for (unsigned d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    __syncthreads();  // warp_sums may still be read from an earlier call
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    unsigned base = 0, sum = 0;
#pragma unroll
    for (unsigned k = 0; k < SCAN_THREADS / 32; k++) {
        const unsigned s = warp_sums[k];
        if (k < wid) base += s;
        sum += s;
    }
    *total = sum;
    return base + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) ent_prefix(const EntImage* __restrict__ imgs, EntWork w) {
    const EntImage& im = imgs[blockIdx.x];
    const unsigned n = im.nsub, per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const unsigned lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    unsigned blocks = 0, vals = 0;
    for (unsigned i = lo; i < hi; i++) {
        blocks += ent_unpack(w.state[im.sub0 + i]).nb;
        vals += w.nvals[im.sub0 + i];
    }
    unsigned total;
    unsigned acc_b = block_exclusive(blocks, &total), acc_v = block_exclusive(vals, &total);
    for (unsigned i = lo; i < hi; i++) {
        w.first_block[im.sub0 + i] = acc_b;
        w.first_val[im.sub0 + i] = acc_v;
        acc_b += ent_unpack(w.state[im.sub0 + i]).nb;
        acc_v += w.nvals[im.sub0 + i];
    }
}

// the DC slot (compact stream, scan order) of the q-th block of component c of the interval (MCU by MCU, v then h inside an MCU)
__device__ __forceinline__ short* dc_ptr(uint8_t* streams, const EntImage& im, unsigned c, unsigned q) {
    const unsigned hv = (unsigned)im.h[c] * im.v[c];
    const unsigned ml = q / hv, r = q - ml * hv;
    short* dc = reinterpret_cast<short*>(streams + im.cs_off + 8ull * im.nb_pad);
    return dc + (size_t)(im.mcu0 + ml) * im.bpm + im.comp_j0[c] + r;
}

// src/decoder.rs:1096-1110: dc_predictor = dc_predictor.wrapping_add(diff), per component, along the scan -- a prefix
// sum over the DC differences the write pass stored.  One thread per block, three short launches:
//   ent_dc_sums  : sum of every chunk of SCAN_THREADS consecutive blocks (in scan order) of a component
//   ent_dc_chunks: exclusive prefix of the chunk sums, one CTA per (component, image)
//   ent_dc_apply : block-wide inclusive prefix inside the chunk + the chunk's offset, written back
// grid = (chunks of the largest component of the launch, 4 components, images)
// (One CTA per (component, interval) that walks runs of consecutive blocks twice around a single block-wide scan was measured:
// 57 us whatever the number of images, against 16 + 5 + 18 us for these three -- a latency chain of 64 dependent-address loads
// per thread where these kernels have one.)
__global__ void __launch_bounds__(SCAN_THREADS) ent_dc_sums(const EntImage* __restrict__ imgs, uint8_t* streams,
                                                            unsigned* __restrict__ chunk_sums, unsigned max_chunks) {
    const EntImage& im = imgs[blockIdx.z];
    const unsigned c = blockIdx.y;
    if (c >= im.ncomp || blockIdx.x * SCAN_THREADS >= im.comp_blocks[c]) return;
    const unsigned q = blockIdx.x * SCAN_THREADS + threadIdx.x;
    const unsigned v = q < im.comp_blocks[c] ? (unsigned)(unsigned short)*dc_ptr(streams, im, c, q) : 0u;
    unsigned total;
    block_exclusive(v, &total);
    if (threadIdx.x == 0) chunk_sums[((size_t)blockIdx.z * 4 + c) * max_chunks + blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) ent_dc_chunks(const EntImage* __restrict__ imgs, unsigned* __restrict__ chunk_sums, unsigned max_chunks) {
    const EntImage& im = imgs[blockIdx.y];
    const unsigned c = blockIdx.x;
    if (c >= im.ncomp) return;
    unsigned* sums = chunk_sums + ((size_t)blockIdx.y * 4 + c) * max_chunks;
    const unsigned n = (im.comp_blocks[c] + SCAN_THREADS - 1) / SCAN_THREADS;
    unsigned carry = 0;
    for (unsigned base = 0; base < n; base += SCAN_THREADS) {  // uniform trip count: block_exclusive has barriers
        const unsigned k = base + threadIdx.x;
        const unsigned v = k < n ? sums[k] : 0u;
        unsigned total;
        const unsigned ex = block_exclusive(v, &total);
        if (k < n) sums[k] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) ent_dc_apply(const EntImage* __restrict__ imgs, uint8_t* streams,
                                                             const unsigned* __restrict__ chunk_sums, unsigned max_chunks) {
    const EntImage& im = imgs[blockIdx.z];
    const unsigned c = blockIdx.y;
    if (c >= im.ncomp || blockIdx.x * SCAN_THREADS >= im.comp_blocks[c]) return;
    const unsigned q = blockIdx.x * SCAN_THREADS + threadIdx.x;
    const bool valid = q < im.comp_blocks[c];
    short* p = dc_ptr(streams, im, c, valid ? q : 0u);
    const unsigned v = valid ? (unsigned)(unsigned short)*p : 0u;
    unsigned total;
    const unsigned ex = block_exclusive(v, &total);
    if (valid) *p = (short)(unsigned short)(chunk_sums[((size_t)blockIdx.z * 4 + c) * max_chunks + blockIdx.x] + ex + v);
}

}  // namespace

static unsigned dc_chunks(unsigned max_comp_blocks) { return (max_comp_blocks + SCAN_THREADS - 1) / SCAN_THREADS; }

size_t ent_work_bytes(unsigned total_sub, unsigned nimages, unsigned max_comp_blocks, int max_passes) {
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    return up((size_t)total_sub * 8) + 3 * up((size_t)total_sub * 4) + 2 * up(total_sub) + up((size_t)(max_passes + 2) * 4) + up((size_t)nimages * 8) +
           up((size_t)nimages * 4 * dc_chunks(max_comp_blocks) * 4);
}

cudaError_t launch_entropy(const EntImage* d_images, unsigned nimages, unsigned max_nsub, unsigned total_sub, unsigned max_comp_blocks,
                           uint8_t* d_streams, void* d_work, int max_passes, unsigned** d_status, cudaStream_t stream, uint64_t* launches) {
    if (nimages == 0 || max_nsub == 0) return cudaSuccess;
    // tables + scan tile pass 48 KB: the opt-in is an attribute of the function ON THE CURRENT DEVICE, so it is set per call
    // (microseconds per group) rather than once per process -- a process may drive several devices through several contexts
    cudaError_t attr = cudaFuncSetAttribute(ent_pass<ENT_COLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BYTES);
    if (attr == cudaSuccess) attr = cudaFuncSetAttribute(ent_pass<ENT_WRITE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BYTES);
    if (attr == cudaSuccess) attr = cudaFuncSetAttribute(ent_sync, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BYTES);
    if (attr != cudaSuccess) return attr;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    char* p = (char*)d_work;
    EntWork w;
    w.state = (unsigned long long*)p;
    p += up((size_t)total_sub * 8);
    w.first_block = (unsigned*)p;
    p += up((size_t)total_sub * 4);
    w.nvals = (unsigned*)p;
    p += up((size_t)total_sub * 4);
    w.first_val = (unsigned*)p;
    p += up((size_t)total_sub * 4);
    w.ch[0] = (unsigned char*)p;
    p += up(total_sub);
    w.ch[1] = (unsigned char*)p;
    p += up(total_sub);
    w.counters = (unsigned*)p;
    const size_t tail = up((size_t)(max_passes + 2) * 4) + up((size_t)nimages * 8);
    w.status = (unsigned*)(p + up((size_t)(max_passes + 2) * 4));
    *d_status = w.status;
    unsigned* chunk_sums = (unsigned*)(p + tail);
    const unsigned nchunks = dc_chunks(max_comp_blocks);
    cudaError_t e = cudaMemsetAsync(w.counters, 0, tail, stream);
    if (e != cudaSuccess) return e;
    const dim3 sub_grid((max_nsub + ENT_THREADS - 1) / ENT_THREADS, 1);
    // "images" are restart intervals: a group can hold more of them than grid.y / grid.z allow
    for (unsigned base = 0; base < nimages; base += 65535u) {
        const unsigned cnt = min(65535u, nimages - base);
        const EntImage* imgs = d_images + base;
        EntWork wc = w;
        wc.status = w.status + 2 * (size_t)base;
        unsigned* sums = chunk_sums + (size_t)base * 4 * nchunks;
        if (base) {
            e = cudaMemsetAsync(w.counters, 0, up((size_t)(max_passes + 2) * 4), stream);
            if (e != cudaSuccess) return e;
        }
        const dim3 grid(sub_grid.x, cnt);
        ent_pass<ENT_COLD><<<grid, ENT_THREADS, TILE_BYTES, stream>>>(imgs, d_streams, wc);
        for (int r = 1; r <= max_passes; r++) ent_sync<<<grid, ENT_THREADS, TILE_BYTES, stream>>>(imgs, d_streams, wc, (unsigned)r);
        ent_prefix<<<cnt, SCAN_THREADS, 0, stream>>>(imgs, wc);
        ent_pass<ENT_WRITE><<<grid, ENT_THREADS, TILE_BYTES, stream>>>(imgs, d_streams, wc);
        ent_dc_sums<<<dim3(nchunks, 4, cnt), SCAN_THREADS, 0, stream>>>(imgs, d_streams, sums, nchunks);
        ent_dc_chunks<<<dim3(4, cnt), SCAN_THREADS, 0, stream>>>(imgs, sums, nchunks);
        ent_dc_apply<<<dim3(nchunks, 4, cnt), SCAN_THREADS, 0, stream>>>(imgs, d_streams, sums, nchunks);
        if (launches) *launches += (uint64_t)max_passes + 6;
    }
    return cudaGetLastError();
}

}  // namespace b200jpg
