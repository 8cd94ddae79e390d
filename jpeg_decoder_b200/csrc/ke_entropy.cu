// ke_entropy.cu -- the kernels of device entropy decoding; the algorithm is described in entropy_dev.h.
//
//   ent_pass<COLD | SYNC | WRITE>  one thread per subsequence (ENT_SUB_BITS bits of the scan), 128 per CTA; the image's
//                                  decoding tables (<= 6.4 KB) live in shared memory, the scan bytes are read through L1
//                                  (every thread walks its own 128-byte line)
//   ent_prefix                     per image: exclusive sum of the blocks each subsequence completed
//   ent_dc                         per (image, component): DC differences -> DC values (wrapping int16 prefix sum)
//
// grid.y = image, grid.x covers the longest scan of the launch; CTAs past the end of their image exit at once.
#include <cuda_runtime.h>

#include "entropy_dev.h"
#include "kernels.h"

namespace b200jpg {

namespace {

constexpr int ENT_COLD = 0, ENT_SYNC = 1, ENT_WRITE = 2;
constexpr unsigned ENT_THREADS = 128;

struct EntShared {
    EntImage im;
    EntTables tabs[ENT_MAX_SLOTS];
    uint8_t dcslot[12], acslot[12];
    uint8_t unzz[64];
};

__constant__ uint8_t c_unzigzag[64] = ENT_UNZIGZAG_INIT;

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
    if (threadIdx.x < 64) sh.unzz[threadIdx.x] = c_unzigzag[threadIdx.x];
    __syncthreads();
}

// work arrays of one launch (one entry per subsequence of every image, images back to back)
struct EntWork {
    unsigned long long* state;
    unsigned* first_block;
    unsigned char* ch[2];
    unsigned* counters;  // counters[r] != 0: pass r changed some state (counters[0] is set by the cold pass)
    unsigned* status;    // per image: [anomaly bits, completed]
};

template <int MODE>
__global__ void __launch_bounds__(ENT_THREADS) ent_pass(const EntImage* __restrict__ imgs, const uint8_t* __restrict__ streams, EntWork w,
                                                        unsigned pass, short* __restrict__ coefs) {
    if (MODE == ENT_SYNC && w.counters[pass - 1] == 0) return;  // converged earlier: nothing left to do
    const EntImage& im = imgs[blockIdx.y];
    if (blockIdx.x * ENT_THREADS >= im.nsub) return;
    const unsigned i = blockIdx.x * ENT_THREADS + threadIdx.x;
    const bool valid = i < im.nsub;
    const unsigned g = im.sub0 + i;
    const unsigned char* ch_in = (pass & 1u) ? w.ch[0] : w.ch[1];
    unsigned char* ch_out = (pass & 1u) ? w.ch[1] : w.ch[0];
    bool active = valid;
    if (MODE == ENT_SYNC) {
        active = valid && i > 0 && ch_in[g - 1] != 0;
        if (!__syncthreads_or(active)) {
            if (valid) ch_out[g] = 0;
            return;
        }
    }
    if (MODE == ENT_WRITE) {
        active = valid && w.first_block[g] < im.total_blocks;
        if (!__syncthreads_or(active)) return;
    }
    __shared__ EntShared sh;
    const uint8_t* payload = streams + im.payload_off;
    load_shared(sh, im, payload);
    if (!active) {
        if (MODE == ENT_SYNC && valid) ch_out[g] = 0;
        return;
    }
    const uint32_t* words = reinterpret_cast<const uint32_t*>(payload + im.data_off);
    EntState st;
    st.p = i * ENT_SUB_BITS;
    st.k = st.b = st.nb = 0;
    if (MODE != ENT_COLD && i > 0) {
        st = ent_unpack(w.state[g - 1]);
        st.nb = 0;
    }
    unsigned bad = 0;
    if (MODE != ENT_WRITE) {
        EntNullSink sink;
        const uint64_t v = ent_pack(ent_decode_range(words, im.nwords, sh.tabs, sh.dcslot, sh.acslot, im.dec_bpm, st,
                                                     ent_sub_end(i, im.nsub, im.scan_bits), sink, &bad));
        if (MODE == ENT_COLD) {
            w.state[g] = v;
            ch_out[g] = 1;  // pass 0 writes ch[0]; pass 1 reads it
            if (i == 0) w.counters[0] = 1;
        } else {
            const bool changed = ((v ^ w.state[g]) & ENT_SYNC_MASK) != 0;
            w.state[g] = v;
            ch_out[g] = changed ? 1 : 0;
            if (changed) w.counters[pass] = 1;
        }
    } else {
        EntWriteSink sink;
        sink.begin(coefs, &sh.im, sh.unzz, w.first_block[g]);
        const bool last = i + 1 == im.nsub;
        const uint32_t end = last ? im.scan_bits + ENT_TAIL_SLACK_BITS : ent_sub_end(i, im.nsub, im.scan_bits);
        const EntState e = ent_decode_range(words, im.nwords, sh.tabs, sh.dcslot, sh.acslot, im.dec_bpm, st, end, sink, &bad);
        if (sink.B >= im.total_blocks) w.status[2 * blockIdx.y + 1] = 1;  // every block has been delivered
        else if (last) bad |= ENT_INCOMPLETE;
        else if (ent_pack(e) != w.state[g]) bad |= ENT_BAD_CHAIN;
        if (bad) atomicOr(&w.status[2 * blockIdx.y], bad);
    }
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
    unsigned sum = 0;
    for (unsigned i = lo; i < hi; i++) sum += ent_unpack(w.state[im.sub0 + i]).nb;
    unsigned total;
    unsigned acc = block_exclusive(sum, &total);
    for (unsigned i = lo; i < hi; i++) {
        w.first_block[im.sub0 + i] = acc;
        acc += ent_unpack(w.state[im.sub0 + i]).nb;
    }
}

// coefficient 0 of the q-th block of component c in scan order (MCU by MCU, v then h inside an MCU)
__device__ __forceinline__ short* dc_ptr(short* coefs, const EntImage& im, unsigned c, unsigned q) {
    const unsigned h = im.h[c], hv = h * im.v[c];
    const unsigned m = q / hv, r = q - m * hv, vy = r / h, hx = r - vy * h;
    const unsigned my = m / im.mcu_w, mx = m - my * im.mcu_w;
    return coefs + ((size_t)im.slab_row[c] + (size_t)(my * im.v[c] + vy) * im.block_w[c] + mx * h + hx) * 64;
}

// src/decoder.rs:1096-1110: dc_predictor = dc_predictor.wrapping_add(diff), per component, along the scan
__global__ void __launch_bounds__(SCAN_THREADS) ent_dc(const EntImage* __restrict__ imgs, short* __restrict__ coefs) {
    const EntImage& im = imgs[blockIdx.y];
    const unsigned c = blockIdx.x;
    if (c >= im.ncomp) return;
    const unsigned n = im.comp_blocks[c], per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const unsigned lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    unsigned sum = 0;
    for (unsigned q = lo; q < hi; q++) sum += (unsigned)(unsigned short)*dc_ptr(coefs, im, c, q);
    unsigned total;
    unsigned acc = block_exclusive(sum, &total);
    for (unsigned q = lo; q < hi; q++) {
        short* p = dc_ptr(coefs, im, c, q);
        acc += (unsigned)(unsigned short)*p;
        *p = (short)(unsigned short)acc;
    }
}

}  // namespace

size_t ent_work_bytes(unsigned total_sub, unsigned nimages, int max_passes) {
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    return up((size_t)total_sub * 8) + up((size_t)total_sub * 4) + 2 * up(total_sub) + up((size_t)(max_passes + 2) * 4) + up((size_t)nimages * 8);
}

cudaError_t launch_entropy(const EntImage* d_images, unsigned nimages, unsigned max_nsub, unsigned total_sub, const uint8_t* d_streams, void* d_work,
                           int max_passes, short* d_coefs, unsigned** d_status, cudaStream_t stream, uint64_t* launches) {
    if (nimages == 0 || max_nsub == 0) return cudaSuccess;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    char* p = (char*)d_work;
    EntWork w;
    w.state = (unsigned long long*)p;
    p += up((size_t)total_sub * 8);
    w.first_block = (unsigned*)p;
    p += up((size_t)total_sub * 4);
    w.ch[0] = (unsigned char*)p;
    p += up(total_sub);
    w.ch[1] = (unsigned char*)p;
    p += up(total_sub);
    w.counters = (unsigned*)p;
    const size_t tail = up((size_t)(max_passes + 2) * 4) + up((size_t)nimages * 8);
    w.status = (unsigned*)(p + up((size_t)(max_passes + 2) * 4));
    *d_status = w.status;
    cudaError_t e = cudaMemsetAsync(w.counters, 0, tail, stream);
    if (e != cudaSuccess) return e;
    const dim3 grid((max_nsub + ENT_THREADS - 1) / ENT_THREADS, nimages);
    ent_pass<ENT_COLD><<<grid, ENT_THREADS, 0, stream>>>(d_images, d_streams, w, 0u, nullptr);
    for (int r = 1; r <= max_passes; r++) ent_pass<ENT_SYNC><<<grid, ENT_THREADS, 0, stream>>>(d_images, d_streams, w, (unsigned)r, nullptr);
    ent_prefix<<<nimages, SCAN_THREADS, 0, stream>>>(d_images, w);
    ent_pass<ENT_WRITE><<<grid, ENT_THREADS, 0, stream>>>(d_images, d_streams, w, 0u, d_coefs);
    ent_dc<<<dim3(4, nimages), SCAN_THREADS, 0, stream>>>(d_images, d_coefs);
    if (launches) *launches += (uint64_t)max_passes + 4;
    return cudaGetLastError();
}

}  // namespace b200jpg
