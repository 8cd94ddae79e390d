// kf_fused.cu -- KF: dequantise + IDCT + chroma upsampling + YCbCr->RGB in ONE pass, coefficient slab -> pixel slab
// (sm_100a).  SURVEY section 8 row f4: the component planes never touch HBM (9.04 -> 6.02 B/px at 1080p 4:2:0,
// 15 -> 9 B/px at 4:4:4).
//
// Replaces, per image, Worker::start + append_row* + get_result (src/worker/immediate.rs:30-60) for the three
// components AND compute_image (src/decoder.rs:1300-1336 -> src/worker/mod.rs:97-128 -> src/upsampler.rs:47-63,
// 191-228 -> src/decoder.rs:1406-1437) in a single kernel.  Same arithmetic as K1 + K2 (idct_core.cuh,
// color_core.cuh), hence the same bytes.
//
// Shape.  Persistent CTAs of 8 warps, 2 per SM.  The batch is flattened into items = (image, column strip of
// <= 1920 pixels, MCU row); every CTA owns a contiguous range of items, so it walks down its columns one MCU row at a
// time.  Per item:
//   phase A  every warp takes 32-block boxes of the item (lane = one 8x8 block in registers): the coefficients arrive
//            by 2-D TMA (128B swizzle) into the warp's own 4 KB slot, the next box is requested as soon as the current
//            one sits in registers (the slot is private to the warp: no empty-barrier, no producer warp), and the
//            samples are written to plane rows staged in shared memory;
//   phase B  threads take (row pair, 16-pixel group) tasks exactly like K2: triangle filter by IDP.4A, colour by IMAD,
//            128-bit stores of interleaved RGB.
// 4:2:0 needs chroma row 8r-1 and luma row 16r-1 of the MCU row above for its first output row pair: output rows
// 16r-1 .. 16r+14 are emitted at MCU row r, and the last luma row and the last chroma rows are carried in two spare
// staged rows (alternating by the parity of r, so no extra barrier).  A CTA whose range starts in the middle of a
// column first runs phase A alone on the MCU row above.  Column strips of wide images recompute one chroma block
// to the left and right (the filter's horizontal halo).
//
// Roofline: HBM in bytes (128 B read per block + 3 B written per pixel), but -- like K1 and K2 -- bound by integer
// issue: ~800 instructions per block and ~17 per pixel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "device_types.h"
#include "kernels.h"
#include "ptx.cuh"
#include "idct_core.cuh"
#include "color_core.cuh"

namespace b200jpg {

constexpr unsigned KF_DEFAULT_WARPS = 4;
constexpr unsigned KF_BOX = 32;                // blocks per TMA box = one warp's lanes
constexpr unsigned KF_SLOT_BYTES = KF_BOX * 128;
// staged plane rows: 4:2:0 = 16 luma + 2 carry, 8 chroma + 2 carry per component; 4:4:4 = 8 rows per component
constexpr unsigned KF_YROWS_420 = 18, KF_CROWS_420 = 10, KF_ROWS_444 = 8;

size_t kf_smem_bytes(unsigned mode, unsigned ystride, unsigned cstride, unsigned KF_WARPS) {
    const size_t planes = mode == KF_MODE_420 ? (size_t)KF_YROWS_420 * ystride + 2u * KF_CROWS_420 * cstride : (size_t)3u * KF_ROWS_444 * ystride;
    return 1024 + (size_t)KF_WARPS * KF_SLOT_BYTES + planes;
}

struct KfBox {  // one 32-block box of an item, warp-uniform
    unsigned slab_row;  // TMA row coordinate of the box
    unsigned run;       // index into FColumn::run
    unsigned idx0;      // index of lane 0's block inside the run
};

__device__ __forceinline__ KfBox kf_box(const FColumn* __restrict__ col, unsigned r, unsigned b) {
    unsigned k = 0;
#pragma unroll
    for (unsigned j = 1; j < 4; j++)
        if (j < __ldg(&col->nruns) && b >= __ldg(&col->run[j].box0)) k = j;
    KfBox x;
    x.run = k;
    x.idx0 = (b - __ldg(&col->run[k].box0)) * KF_BOX;
    x.slab_row = __ldg(&col->run[k].slab_row0) + r * __ldg(&col->run[k].step) + x.idx0;
    return x;
}

__device__ __forceinline__ void sts64(unsigned addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(unsigned addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// a wait that cannot hang the GPU: a barrier that does not complete within ~1 s of polling is a bug -> trap
__device__ __forceinline__ void kf_wait(unsigned bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}

// KF_WARPS warps per CTA, 16 / KF_WARPS CTAs per SM (the register file holds 16 warps at 128 registers)
template <unsigned MODE, unsigned KF_WARPS>
__global__ void __launch_bounds__(KF_WARPS * 32, 16 / KF_WARPS)
kf_fused(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ K1QCache qc, KFParams p) {
    constexpr unsigned KF_THREADS = KF_WARPS * 32;
    extern __shared__ __align__(1024) uint8_t kf_smem[];
    __shared__ __align__(8) unsigned long long full_bar[KF_WARPS];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned smem = (smem_u32(kf_smem) + 1023u) & ~1023u;
    const unsigned slot = smem + warp * KF_SLOT_BYTES;
    const unsigned bar = smem_u32(&full_bar[warp]);
    const unsigned ystride = p.ystride, cstride = p.cstride;
    // staged planes: base address and row stride of each component
    const unsigned ybase = smem + KF_WARPS * KF_SLOT_BYTES;
    unsigned pbase[3], pstride[3];
    if (MODE == KF_MODE_420) {
        pbase[0] = ybase; pstride[0] = ystride;
        pbase[1] = ybase + KF_YROWS_420 * ystride; pstride[1] = cstride;
        pbase[2] = pbase[1] + KF_CROWS_420 * cstride; pstride[2] = cstride;
    } else {
        pbase[0] = ybase; pbase[1] = ybase + KF_ROWS_444 * ystride; pbase[2] = ybase + 2u * KF_ROWS_444 * ystride;
        pstride[0] = pstride[1] = pstride[2] = ystride;
    }

    const unsigned w_begin = p.item_base + (unsigned)(((unsigned long long)blockIdx.x * p.total_items) / gridDim.x);
    const unsigned w_end = p.item_base + (unsigned)(((unsigned long long)(blockIdx.x + 1) * p.total_items) / gridDim.x);
    if (w_end == w_begin) return;

    if (tid == 0) {
        for (unsigned w = 0; w < KF_WARPS; w++) mbar_init(smem_u32(&full_bar[w]), 1);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    __syncthreads();

    // column containing item w_begin
    unsigned lo = 0, hi = p.ncols - 1;
    while (lo < hi) {
        const unsigned mid = (lo + hi + 1) / 2;
        if (__ldg(&p.cols[mid].first_item) <= w_begin) lo = mid; else hi = mid - 1;
    }
    unsigned ci = lo;
    unsigned r = w_begin - __ldg(&p.cols[ci].first_item);
    // 4:2:0, range starting inside a column: one extra item (phase A only) on the MCU row above
    const bool warm = MODE == KF_MODE_420 && r > 0;
    if (warm) r -= 1;
    const unsigned nitems = (w_end - w_begin) + (warm ? 1u : 0u);

    const YccRegs ycc = make_ycc_regs(make_int3(p.sixteen, p.sixteen, p.sixteen), true);
    unsigned phase = 0;  // parity of this warp's next wait on its full barrier

    bool pending = false;  // warp-uniform: a box has been requested into this warp's slot and not consumed yet

    for (unsigned k = 0; k < nitems; k++) {
        const FColumn* __restrict__ col = &p.cols[ci];
        const unsigned nboxes = __ldg(&col->nboxes), nrows = __ldg(&col->nrows);
        // the item after this one (for the prefetch across the item boundary)
        unsigned nci = ci, nr = r + 1;
        if (nr == nrows) { nci = ci + 1; nr = 0; }
        const bool have_next_item = k + 1 < nitems;

        // ------------------------------------------------ phase A: dequantise + IDCT into the staged planes
        if (warp < nboxes && !pending) {  // first item, or a warp that had no box in the previous item
            if (lane == 0) {
                const KfBox x = kf_box(col, r, warp);
                mbar_expect_tx(bar, KF_SLOT_BYTES);
                tma_load_2d(slot, &tmap, 0, (int)x.slab_row, bar);
            }
            pending = true;
        }
        for (unsigned b = warp; b < nboxes; b += KF_WARPS) {
            const KfBox x = kf_box(col, r, b);
            const unsigned rcomp = __ldg(&col->run[x.run].comp), rlen = __ldg(&col->run[x.run].len), rwrap = __ldg(&col->run[x.run].wrap);
            const unsigned rdst_x = __ldg(&col->run[x.run].dst_x), rdst_row = __ldg(&col->run[x.run].dst_row);
            const uint4* cw = reinterpret_cast<const uint4*>(&p.comps[__ldg(&col->comp0) + rcomp]);
            const uint4 c0 = __ldg(cw), c1 = __ldg(cw + 1);  // DevComp: {plane_off lo, hi, stride, block_w}, {qt_index, dct_scale, nblocks, qflags}
            (void)c0;
            const unsigned qt_index = c1.x, qflags = c1.w;
            const uint4* q4 = reinterpret_cast<const uint4*>(p.qtabs + (size_t)qt_index * 64);
            const uint4* qp4 = reinterpret_cast<const uint4*>(p.qpack + (size_t)qt_index * 32);

            kf_wait(bar, phase);
            phase ^= 1u;
            const unsigned sbase = slot + lane * 128u, swz = (lane & 7u) << 4;
            uint4 raw[8];
#pragma unroll
            for (int j = 0; j < 8; j++) raw[j] = lds128(sbase + ((j * 16u) ^ swz));

            const unsigned idx = x.idx0 + lane;
            unsigned s[8][8];
            // (dequantising first also guarantees that every lane's slot reads have returned before the slot is refilled)
            const unsigned oor = dequant_block(raw, s, qflags, qc, q4, qp4);
            __syncwarp();
            {  // next box of this warp: same item, else its first one of the next item
                const bool same = b + KF_WARPS < nboxes;
                pending = same || (have_next_item && warp < __ldg(&p.cols[nci].nboxes));
                if (pending && lane == 0) {
                    const KfBox nx = same ? kf_box(col, r, b + KF_WARPS) : kf_box(&p.cols[nci], nr, warp);
                    mbar_expect_tx(bar, KF_SLOT_BYTES);
                    tma_load_2d(slot, &tmap, 0, (int)nx.slab_row, bar);
                }
            }
            if (idx >= rlen) continue;
            const unsigned brow = idx >= rwrap ? 1u : 0u;
            const unsigned bcol = idx - (brow ? rwrap : 0u);
            const unsigned dst = pbase[rcomp] + (rdst_row + 8u * brow) * pstride[rcomp] + rdst_x + 8u * bcol;
            const unsigned dstride = pstride[rcomp];
            uint2 rows[8];
            if (oor != 0) {  // |c*q| >= 2^19 in the first row: the reference form with its zero-AC column shortcut
                __align__(8) uint8_t tmp[64];
                idct8x8_scalar_exact(p.coefs + ((size_t)x.slab_row + lane) * 64, reinterpret_cast<const unsigned*>(q4), tmp, 8);
#pragma unroll
                for (int j = 0; j < 8; j++) rows[j] = *reinterpret_cast<const uint2*>(tmp + 8 * j);
            } else {
                idct8x8_direct(s, rows);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) sts64(dst + (unsigned)j * dstride, rows[j]);
        }
        __syncthreads();

        // ------------------------------------------------ phase B: upsample + colour convert + store
        const bool emit = !(warm && k == 0);
        const DevImage& img = p.images[__ldg(&col->image)];
        const unsigned W = img.width, H = img.height;
        const unsigned x0 = __ldg(&col->x0), wpx = __ldg(&col->wpx), ngroups = __ldg(&col->ngroups), gmagic = __ldg(&col->gmagic);
        uint8_t* const out = p.out + img.out_off + (size_t)x0 * 3u;
        const size_t row_bytes = (size_t)W * 3u;
        if (MODE == KF_MODE_420) {
            const unsigned par = r & 1u;
            if (emit) {
                const unsigned in_w = img.c[1].in_w, in_h = img.c[1].in_h;
                const unsigned p_first = 8u * r;
                const unsigned p_last = (r + 1 == nrows) ? H / 2u : p_first + 7u;  // inclusive
                const unsigned ntasks = (p_last - p_first + 1u) * ngroups;
                const unsigned cx_base = __ldg(&col->cx_base);
                for (unsigned t = tid; t < ntasks; t += KF_THREADS) {
                    const unsigned pi = ngroups == 1u ? t : __umulhi(t, gmagic), g = t - pi * ngroups;
                    const unsigned pr = p_first + pi;
                    // chroma rows A = max(pr-1, 0), B = min(pr, in_h-1) as staged row indices (8 + par = the carried row)
                    const unsigned rowB = min(pr, in_h - 1u) - p_first;
                    const unsigned rowA = pr == 0 ? rowB : (pi == 0 ? 8u + par : pi - 1u);
                    const unsigned gi = x0 / 2u + 8u * g;             // global index of the group's first chroma sample
                    const unsigned oM = 16u + gi - cx_base;           // its byte offset in a staged chroma row
                    const unsigned oL = gi > 0 ? oM - 1u : oM;
                    const unsigned oR = 16u + min(gi + 8u, in_w - 1u) - cx_base;
                    Chroma16 cb, cr;
#pragma unroll
                    for (int c = 1; c <= 2; c++) {
                        const unsigned a_row = pbase[c] + rowA * cstride, b_row = pbase[c] + rowB * cstride;
                        uint2 av = lds64(a_row + oM), bv = lds64(b_row + oM);
                        if (gi + 8u > in_w) {  // last group of a ragged row: block padding := last valid sample
                            av = replicate_last_sample(av, in_w - gi);
                            bv = replicate_last_sample(bv, in_w - gi);
                        }
                        h2v2_16(av.x, av.y, lds8(a_row + oL), lds8(a_row + oR), bv.x, bv.y, lds8(b_row + oL), lds8(b_row + oR), c == 1 ? cb : cr);
                    }
                    const unsigned npx = min(16u, wpx - 16u * g);
                    uint8_t* const o = out + (size_t)g * 48u + (size_t)(2u * pr) * row_bytes;
                    if (pr > 0) {  // output row 2 pr - 1: luma row 2 pi - 1 of the item, the carried one for pi == 0
                        const unsigned yrow = pi == 0 ? 16u + par : 2u * pi - 1u;
                        ycbcr_store16(lds128(ybase + yrow * ystride + 16u * g), cb.odd, cr.odd, o - row_bytes, ycc, npx);
                    }
                    if (2u * pr < H) ycbcr_store16(lds128(ybase + 2u * pi * ystride + 16u * g), cb.even, cr.even, o, ycc, npx);
                }
            }
            // carry the last luma row and the last chroma rows to the MCU row below (other parity: nobody reads it now)
            const unsigned npar = par ^ 1u;
            for (unsigned i = tid; i < ystride / 16u; i += KF_THREADS)
                sts128(ybase + (16u + npar) * ystride + 16u * i, lds128(ybase + 15u * ystride + 16u * i));
            for (unsigned i = tid; i < 2u * (cstride / 16u); i += KF_THREADS) {
                const unsigned c = i >= cstride / 16u ? 2u : 1u, off = 16u * (i - (c == 2u ? cstride / 16u : 0u));
                sts128(pbase[c] + (8u + npar) * cstride + off, lds128(pbase[c] + 7u * cstride + off));
            }
        } else {
            const unsigned nr_rows = min(8u, H - 8u * r);
            const unsigned ntasks = nr_rows * ngroups;
            for (unsigned t = tid; t < ntasks; t += KF_THREADS) {
                const unsigned j = ngroups == 1u ? t : __umulhi(t, gmagic), g = t - j * ngroups;
                const uint4 yv = lds128(pbase[0] + j * ystride + 16u * g);
                const uint4 bv = lds128(pbase[1] + j * ystride + 16u * g);
                const uint4 rv = lds128(pbase[2] + j * ystride + 16u * g);
                const unsigned bw[4] = {bv.x ^ 0x80808080u, bv.y ^ 0x80808080u, bv.z ^ 0x80808080u, bv.w ^ 0x80808080u};
                const unsigned rw[4] = {rv.x ^ 0x80808080u, rv.y ^ 0x80808080u, rv.z ^ 0x80808080u, rv.w ^ 0x80808080u};
                int cb[16], cr[16];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    cb[q] = (int)prmt(bw[q >> 2], 0u, 0x8880u + 0x1111u * (unsigned)(q & 3));
                    cr[q] = (int)prmt(rw[q >> 2], 0u, 0x8880u + 0x1111u * (unsigned)(q & 3));
                }
                ycbcr_store16(yv, cb, cr, out + (size_t)g * 48u + (size_t)(8u * r + j) * row_bytes, ycc, min(16u, wpx - 16u * g));
            }
        }
        __syncthreads();
        ci = nci;
        r = nr;
    }
}

// ---------------------------------------------------------------------------------------------
// profiling knobs (defaults measured on B200, profiles/r02_kf_sweep.md): warps per CTA and strip width
static unsigned kf_env(const char* name, unsigned dflt) {
    const char* e = getenv(name);
    return e ? (unsigned)atoi(e) : dflt;
}
unsigned kf_warps() {
    static unsigned w = 0;
    if (!w) w = kf_env("B200JPG_KF_WARPS", KF_DEFAULT_WARPS) == 4 ? 4 : 8;
    return w;
}
unsigned kf_strip_px() {
    static unsigned px = 0;
    if (!px) {
        px = kf_env("B200JPG_KF_STRIP", kf_warps() == 4 ? 960 : 1920);
        px = px < 128 ? 128 : (px > 1920 ? 1920 : px) / 16 * 16;
    }
    return px;
}

template <unsigned MODE, unsigned W>
static cudaError_t launch_kf_t(const CUtensorMap& tmap32, const K1QCache& qc, const KFParams& p, int num_sms, cudaStream_t stream) {
    const size_t smem_bytes = kf_smem_bytes(MODE, p.ystride, p.cstride, W);
    static size_t attr_set = 0;
    if (attr_set < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(kf_fused<MODE, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        attr_set = smem_bytes;
    }
    unsigned grid = (unsigned)num_sms * (16u / W);
    if (grid > p.total_items) grid = p.total_items;
    kf_fused<MODE, W><<<grid, W * 32, smem_bytes, stream>>>(tmap32, qc, p);
    return cudaGetLastError();
}

cudaError_t launch_kf(unsigned mode, const CUtensorMap& tmap32, const K1QCache& qc, const KFParams& p, int num_sms, cudaStream_t stream) {
    if (p.ncols == 0 || p.total_items == 0) return cudaSuccess;
    const bool w4 = kf_warps() == 4;
    if (mode == KF_MODE_420) return w4 ? launch_kf_t<KF_MODE_420, 4>(tmap32, qc, p, num_sms, stream) : launch_kf_t<KF_MODE_420, 8>(tmap32, qc, p, num_sms, stream);
    return w4 ? launch_kf_t<KF_MODE_444, 4>(tmap32, qc, p, num_sms, stream) : launch_kf_t<KF_MODE_444, 8>(tmap32, qc, p, num_sms, stream);
}

}  // namespace b200jpg
