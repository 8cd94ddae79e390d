// k0_expand.cu -- K0: sparse block stream -> dense coefficient slab (sm_100a).
//
// The reference hands the worker one dense Vec<i16> per MCU row (Worker::append_row,
// src/worker/mod.rs:26, filled at src/decoder.rs:962-983).  Over PCIe that is 128 B per block of mostly
// zeros, so the host ships the stream of sbs.h instead and this kernel rebuilds, in HBM, exactly the dense
// raster-order slab K1 consumes (device_types.h): same bytes as if the dense buffers had been uploaded.
//
// One warp = one group of 32 scan-order blocks.  Lane l owns block 32g+l for the bookkeeping (bitmap, byte
// count, warp scan for its value offset, destination row); the expansion itself is cooperative: for every
// block of the group lane l produces zig-zag positions l and l+32 (rank of the position among the bitmap's
// set bits = index of its value), so value bytes are read by neighbouring lanes and every one of the 64
// natural-order slots is written exactly once (no zero fill).  A 4 KB shared-memory tile per warp turns the
// scatter into 16-byte row stores: 128 contiguous bytes per block, blocks of a group are (nearly) adjacent.
//
// Roofline: HBM write of the slab (128 B / block) + stream read (~25 B / block); it runs only on the
// host-fed paths, which are PCIe / host bound by two orders of magnitude.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "sbs.h"

namespace b200jpg {

__constant__ unsigned char c_unzigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                             12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                             58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int K0_WARPS = 8;

__global__ void __launch_bounds__(K0_WARPS * 32) k0_expand(const K0Image* __restrict__ images, const uint8_t* __restrict__ streams,
                                                           short* __restrict__ slab) {
    __shared__ __align__(16) short tile[K0_WARPS][32][64];
    const K0Image& im = images[blockIdx.y];
    if (im.order & SBS_BLOCK_OFFSETS) return;  // device-made streams: k0_expand_blocks
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned g = blockIdx.x * K0_WARPS + warp;
    const unsigned nb = im.nb;
    if (g * 32u >= nb) return;
    const unsigned nb_pad = (nb + 31u) & ~31u;
    const uint8_t* s = streams + im.stream_off;
    const unsigned t = g * 32u + lane;
    const unsigned long long bm = ((const unsigned long long*)s)[t];
    const int dcv = ((const short*)(s + 8ull * nb_pad))[t];
    const bool block_offsets = (im.order & SBS_BLOCK_OFFSETS) != 0;  // device-made stream: one offset per block, relative to s
    const unsigned base = ((const unsigned*)(s + 10ull * nb_pad))[block_offsets ? t : g];
    const uint8_t* vals = block_offsets ? s : s + ((10ull * nb_pad + 4ull * (nb_pad / 32u + 1u) + 15ull) & ~15ull);

    // byte offset of this lane's values: exclusive scan of the per-block byte counts
    const unsigned bytes = (unsigned)__popcll(bm >> 1) << (unsigned)(bm & 1ull);
    unsigned incl = bytes;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += o;
    }
    const unsigned off = block_offsets ? base : base + incl - bytes;

    // destination row of this lane's block inside the dense slab
    unsigned row = 0xffffffffu;
    if (t < nb) {
        if ((im.order & SBS_INTERLEAVED) == 0) {
            const unsigned c = (t >= im.first[1]) + (t >= im.first[2]) + (t >= im.first[3]);
            row = im.slab_row[c] + (t - im.first[c]);
        } else {
            const unsigned mcu = t / im.bpm, j = t - mcu * im.bpm;
            const unsigned c = im.mcu_comp[j];
            const unsigned my = mcu / im.mcu_w, mx = mcu - my * im.mcu_w;
            row = im.slab_row[c] + (my * im.v[c] + im.mcu_vy[j]) * im.block_w[c] + mx * im.h[c] + im.mcu_hx[j];
        }
    }

    const bool natural = (im.order & SBS_NATURAL) != 0;
    const unsigned p0 = natural ? lane : c_unzigzag[lane], p1 = natural ? lane + 32u : c_unzigzag[lane + 32u];
    const unsigned long long below0 = ((1ull << lane) - 1ull) & ~1ull, below1 = ((1ull << (lane + 32u)) - 1ull) & ~1ull;
#pragma unroll 4
    for (int b = 0; b < 32; b++) {
        const unsigned long long m = __shfl_sync(0xffffffffu, bm, b);
        const unsigned o = __shfl_sync(0xffffffffu, off, b);
        const int d = __shfl_sync(0xffffffffu, dcv, b);
        const bool wide = (m & 1ull) != 0;
        int v0 = 0, v1 = 0;
        if (lane == 0) {
            v0 = d;
        } else if ((m >> lane) & 1ull) {
            const unsigned r = (unsigned)__popcll(m & below0);
            v0 = block_offsets ? (int)*(const short*)(vals + o + 2u * r)  // device-made: always int16, 2-byte aligned
                 : wide        ? (int)(short)((unsigned)vals[o + 2u * r] | ((unsigned)vals[o + 2u * r + 1u] << 8))
                               : (int)(signed char)vals[o + r];
        }
        if ((m >> (lane + 32u)) & 1ull) {
            const unsigned r = (unsigned)__popcll(m & below1);
            v1 = block_offsets ? (int)*(const short*)(vals + o + 2u * r)  // device-made: always int16, 2-byte aligned
                 : wide        ? (int)(short)((unsigned)vals[o + 2u * r] | ((unsigned)vals[o + 2u * r + 1u] << 8))
                               : (int)(signed char)vals[o + r];
        }
        tile[warp][b][p0] = (short)v0;
        tile[warp][b][p1] = (short)v1;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const unsigned b = (unsigned)i * 4u + (lane >> 3), chunk = lane & 7u;
        const unsigned r = __shfl_sync(0xffffffffu, row, (int)b);
        if (r != 0xffffffffu) *(int4*)(slab + (size_t)r * 64u + chunk * 8u) = *(const int4*)&tile[warp][b][chunk * 8u];
    }
}

// K0 for device-made streams (SBS_BLOCK_OFFSETS: every block has its own value offset, values are int16 in zig-zag order).
// One THREAD per block: it zeroes its 128-byte row of a shared-memory tile, walks the set bits of its bitmap and drops
// each value at its natural position (~15 values per block instead of 64 slots x rank computations in the warp-per-group
// kernel above: 5x fewer instructions); the warp then stores its 32 rows cooperatively, four whole 128-byte blocks per
// instruction.  Rows are swizzled in 16-byte chunks by (thread & 7), so both phases are free of bank conflicts beyond the
// unavoidable ones of scattered 2-byte stores.  (Staging the CTA's values in shared memory with coalesced 16-byte loads first was
// measured and loses: 101 -> 109 us per 27 images -- half the resident CTAs, and the scattered 2-byte stores remain.)
constexpr int K0B_THREADS = 128;
__global__ void __launch_bounds__(K0B_THREADS) k0_expand_blocks(const K0Image* __restrict__ images, const uint8_t* __restrict__ streams,
                                                                short* __restrict__ slab) {
    __shared__ __align__(16) short tile[K0B_THREADS][64];
    __shared__ unsigned char unzz[64];
    const K0Image& im = images[blockIdx.y];
    if ((im.order & SBS_BLOCK_OFFSETS) == 0) return;
    const unsigned nb = im.nb;
    if (blockIdx.x * K0B_THREADS >= nb) return;
    if (threadIdx.x < 64) unzz[threadIdx.x] = c_unzigzag[threadIdx.x];
    const unsigned lane = threadIdx.x & 31u, sw = threadIdx.x & 7u;
    const unsigned t = blockIdx.x * K0B_THREADS + threadIdx.x;
    const unsigned nb_pad = (nb + 31u) & ~31u;
    const uint8_t* s = streams + im.stream_off;
    short* mine = tile[threadIdx.x];
#pragma unroll
    for (int c = 0; c < 8; c++) *reinterpret_cast<int4*>(mine + 8 * c) = make_int4(0, 0, 0, 0);
    __syncthreads();  // unzz
    unsigned row = 0xffffffffu;
    if (t < nb) {
        unsigned long long m = ((const unsigned long long*)s)[t];
        const short* v = reinterpret_cast<const short*>(s + ((const unsigned*)(s + 10ull * nb_pad))[t]);
        mine[(0u ^ sw) << 3] = ((const short*)(s + 8ull * nb_pad))[t];  // natural position 0 = chunk 0, element 0
        m &= ~1ull;
        while (m) {
            const unsigned k = (unsigned)__ffsll((long long)m) - 1u;
            m &= m - 1ull;
            const unsigned nat = unzz[k];
            mine[(((nat >> 3) ^ sw) << 3) | (nat & 7u)] = __ldg(v++);
        }
        const unsigned mcu = t / im.bpm, j = t - mcu * im.bpm;
        if ((im.order & SBS_INTERLEAVED) == 0) {
            const unsigned c = (t >= im.first[1]) + (t >= im.first[2]) + (t >= im.first[3]);
            row = im.slab_row[c] + (t - im.first[c]);
        } else {
            const unsigned c = im.mcu_comp[j];
            const unsigned my = mcu / im.mcu_w, mx = mcu - my * im.mcu_w;
            row = im.slab_row[c] + (my * im.v[c] + im.mcu_vy[j]) * im.block_w[c] + mx * im.h[c] + im.mcu_hx[j];
        }
    }
    __syncwarp();
    const unsigned wbase = threadIdx.x & ~31u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const unsigned b = (unsigned)i * 4u + (lane >> 3), chunk = lane & 7u;
        const unsigned r = __shfl_sync(0xffffffffu, row, (int)b);
        if (r != 0xffffffffu)
            *reinterpret_cast<int4*>(slab + (size_t)r * 64u + chunk * 8u) = *reinterpret_cast<const int4*>(&tile[wbase + b][(chunk ^ (b & 7u)) << 3]);
    }
}

// zeroes the per-block arrays (bm | dc | boff = 14 bytes per block) of device-made streams before the write pass ORs into them
__global__ void __launch_bounds__(256) k0_zero_headers(const K0Image* __restrict__ images, uint8_t* __restrict__ streams) {
    const K0Image& im = images[blockIdx.y];
    if ((im.order & SBS_BLOCK_OFFSETS) == 0) return;
    const size_t n16 = (14ull * ((im.nb + 31u) & ~31u) + 15ull) / 16ull;  // nb_pad is a multiple of 32: 14 * nb_pad is a multiple of 16
    uint4* p = reinterpret_cast<uint4*>(streams + im.stream_off);
    for (size_t q = (size_t)blockIdx.x * 256u + threadIdx.x; q < n16; q += (size_t)gridDim.x * 256u) p[q] = make_uint4(0u, 0u, 0u, 0u);
}

// Upload by kernel: the streams of a group lie scattered over the host threads' page-locked rings; one launch reads them all
// through their device mappings (zero copy) and lays them out back to back in the device stream buffer.  One copy-engine
// transfer per image made ~27 operations per group compete with the pixel download for the link (the download lost 9 %,
// the uploads ran at 13-18 GB/s); a host-side gather into one staging buffer fixed that but cost the submitter thread a 0.3 B
// per pixel memcpy at 4-7 GB/s, which became the bottleneck with two GPUs per box.  This kernel costs the host nothing and
// the download 4 % (scripts/pcie_probe4.py: 29 GB/s next to a saturated download, 50 GB/s alone).
// grid = (GATHER_CTAS, items); 16-byte words, both sides 16-byte aligned.
constexpr unsigned GATHER_CTAS = 5, GATHER_THREADS = 256;
__global__ void __launch_bounds__(GATHER_THREADS) k_gather_streams(const GatherItem* __restrict__ items, uint8_t* __restrict__ dst) {
    const GatherItem it = items[blockIdx.y];
    const uint4* src = reinterpret_cast<const uint4*>(it.src);
    uint4* out = reinterpret_cast<uint4*>(dst + it.dst_off);
    for (unsigned i = blockIdx.x * GATHER_THREADS + threadIdx.x; i < it.n16; i += GATHER_CTAS * GATHER_THREADS) out[i] = src[i];
}

cudaError_t launch_gather_streams(const GatherItem* d_items, unsigned nitems, uint8_t* d_streams, cudaStream_t stream) {
    for (unsigned base = 0; base < nitems; base += 65535u) {
        const unsigned cnt = min(65535u, nitems - base);
        k_gather_streams<<<dim3(GATHER_CTAS, cnt), GATHER_THREADS, 0, stream>>>(d_items + base, d_streams);
    }
    return cudaGetLastError();
}

cudaError_t launch_k0_zero_headers(const K0Image* d_images, unsigned nimages, unsigned max_blocks, uint8_t* d_streams, cudaStream_t stream) {
    if (nimages == 0 || max_blocks == 0) return cudaSuccess;
    const unsigned n16 = (14u * ((max_blocks + 31u) & ~31u) + 15u) / 16u;
    k0_zero_headers<<<dim3(min((n16 + 255u) / 256u, 64u), nimages), 256, 0, stream>>>(d_images, d_streams);
    return cudaGetLastError();
}

cudaError_t launch_k0_expand(const K0Image* d_images, unsigned nimages, unsigned max_blocks, const uint8_t* d_streams, short* d_slab,
                             cudaStream_t stream) {
    if (nimages == 0 || max_blocks == 0) return cudaSuccess;
    const unsigned groups = (max_blocks + 31u) / 32u;
    dim3 grid((groups + K0_WARPS - 1) / K0_WARPS, nimages);
    k0_expand<<<grid, K0_WARPS * 32, 0, stream>>>(d_images, d_streams, d_slab);  // host-made streams (skips the others)
    return cudaGetLastError();
}

cudaError_t launch_k0_expand_blocks(const K0Image* d_images, unsigned nimages, unsigned max_blocks, const uint8_t* d_streams, short* d_slab,
                                    cudaStream_t stream) {
    if (nimages == 0 || max_blocks == 0) return cudaSuccess;
    k0_expand_blocks<<<dim3((max_blocks + K0B_THREADS - 1) / K0B_THREADS, nimages), K0B_THREADS, 0, stream>>>(d_images, d_streams, d_slab);
    return cudaGetLastError();
}

}  // namespace b200jpg
