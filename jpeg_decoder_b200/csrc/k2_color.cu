// k2_color.cu -- K2: chroma upsampling + colour conversion + interleave, plane slab -> pixel slab.
//
// Replaces compute_image (reference src/decoder.rs:1300-1336) -> compute_image_parallel
// (src/worker/mod.rs:97-128, src/worker/rayon.rs:193-219) -> Upsampler::upsample_and_interleave_row
// (src/upsampler.rs:47-63) with the upsamplers of src/upsampler.rs:119-250 and the colour functions of
// src/decoder.rs:1391-1508 / src/arch/ssse3.rs:196-288.  Arithmetic: SURVEY.md Appendix A.3/A.4.
//
// The reference's "fancy" upsamplers special-case the first/last sample; all of them are the
// clamped-edge form of one triangle filter:
//   H2V1: out[2i] = (3 in[i] + in[i-1] + 2) >> 2, out[2i+1] = (3 in[i] + in[i+1] + 2) >> 2 with in[-1] := in[0],
//         in[w] := in[w-1]   ((4a + 2) >> 2 == a reproduces `out[0] = in[0]`, upsampler.rs:151,161)
//   H2V2: t[i] = 3 near[i] + far[i]; out[2i] = (3 t[i] + t[i-1] + 8) >> 4, out[2i+1] = (3 t[i] + t[i+1] + 8) >> 4,
//         clamped ((4t + 8) >> 4 == (t + 2) >> 2 reproduces upsampler.rs:216,226 and the in_w == 1 case 208-213)
//   near/far rows (upsampler.rs:174-180): even row 2k -> near k, far max(k-1,0); odd row 2k+1 -> near k,
//         far min(k+1, in_h-1)   (the f32 expression evaluated exactly; tested against the oracle's f32 form)
//
// Kernels: k2_generic (every upsampler / transform / arithmetic variant, one thread per pixel),
//          k2_ycbcr420 (H2V2 chroma + YCbCr, 16 px x 2 rows per thread, DP4A filter, 128-bit I/O),
//          k2_ycbcr444 (H1V1 + YCbCr, 16 px per thread, 128-bit stores).
// Roofline: HBM.  Algorithmic bytes per image = plane bytes read once + width*height*ncomp written.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "device_types.h"
#include "kernels.h"
#include "ptx.cuh"
#include "color_core.cuh"

namespace b200jpg {


// (r, g, b) before clamping, already shifted down: no intermediate can overflow i32.
__device__ __forceinline__ void ycbcr_scalar(int y, int cb, int cr, int& r, int& g, int& b) {
    // fold the -128 of cb/cr into the constant term: one IMAD chain per channel
    const int yr = y * (1 << 20) + (YCC_HALF - 128 * C_R_CR);
    const int yg = y * (1 << 20) + (YCC_HALF + 128 * C_G_CB + 128 * C_G_CR);
    const int yb = y * (1 << 20) + (YCC_HALF - 128 * C_B_CB);
    r = (yr + C_R_CR * cr) >> 20;
    g = (yg - C_G_CB * cb - C_G_CR * cr) >> 20;
    b = (yb + C_B_CB * cb) >> 20;
}

__device__ __forceinline__ int sat16(int v) { return min(max(v, -32768), 32767); }
__device__ __forceinline__ int mulhrs16(int a, int b) { return (int)(short)((((a * b) >> 14) + 1) >> 1); }
// src/arch/ssse3.rs:208-244
__device__ __forceinline__ void ycbcr_ssse3(int y, int cb, int cr, int& r, int& g, int& b) {
    const int y6 = sat16((y << 6) + 32);
    const int cb6 = sat16((cb << 6) - 8192), cr6 = sat16((cr << 6) - 8192);
    const int cr_140200 = sat16(mulhrs16(cr6, 13173) + cr6);
    const int cb_034414 = mulhrs16(cb6, 11276);
    const int cr_071414 = mulhrs16(cr6, 23401);
    const int cb_177200 = sat16(mulhrs16(cb6, 25297) + cb6);
    r = sat16(y6 + cr_140200) >> 6;
    g = sat16(y6 - sat16(cb_034414 + cr_071414)) >> 6;
    b = sat16(y6 + cb_177200) >> 6;
}

__device__ __forceinline__ uint8_t clamp_u8(int v) { return (uint8_t)min(max(v, 0), 255); }

// One upsampled sample of one component at output position (x, y): src/upsampler.rs:119-250
__device__ __forceinline__ int upsample_at(const uint8_t* __restrict__ plane, const DevUpComp& c, unsigned x, unsigned y) {
    switch (c.kind) {
    case UP_H1V1: return plane[(size_t)y * c.stride + x];
    case UP_H2V1: {
        const uint8_t* in = plane + (size_t)y * c.stride;
        const int i = (int)(x >> 1);
        const int j = (x & 1) ? min(i + 1, (int)c.in_w - 1) : max(i - 1, 0);
        return (3 * in[i] + in[j] + 2) >> 2;
    }
    case UP_H1V2: {
        const int k = (int)(y >> 1);
        const int f = (y & 1) ? min(k + 1, (int)c.in_h - 1) : max(k - 1, 0);
        return (3 * plane[(size_t)k * c.stride + x] + plane[(size_t)f * c.stride + x] + 2) >> 2;
    }
    case UP_H2V2: {
        const int k = (int)(y >> 1);
        const int f = (y & 1) ? min(k + 1, (int)c.in_h - 1) : max(k - 1, 0);
        const uint8_t* n = plane + (size_t)k * c.stride;
        const uint8_t* fr = plane + (size_t)f * c.stride;
        const int i = (int)(x >> 1);
        const int j = (x & 1) ? min(i + 1, (int)c.in_w - 1) : max(i - 1, 0);
        const int ti = 3 * n[i] + fr[i], tj = 3 * n[j] + fr[j];
        return (3 * ti + tj + 8) >> 4;
    }
    default: return plane[(size_t)(y / c.vs) * c.stride + x / c.hs];
    }
}

// ---------------------------------------------------------------------------------------------
// Generic kernel: one thread per output pixel.
// grid.x = ceil(max_w/256) * max_h, grid.y = image
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k2_generic(K2Params p, unsigned first, unsigned xchunks) {
    const DevImage& img = p.images[first + blockIdx.y];
    if (img.path != K2_PATH_GENERIC) return;
    const unsigned y = blockIdx.x / xchunks;
    const unsigned x = (blockIdx.x % xchunks) * 256u + threadIdx.x;
    if (y >= img.height || x >= img.width) return;
    const unsigned W = img.width, n = img.ncomp;
    uint8_t* out = p.out + img.out_off;
    if (img.cc == CC_GRAY) {  // src/decoder.rs:1310-1332: compact stride -> width
        out[(size_t)y * W + x] = p.planes[img.c[0].plane_off + (size_t)y * img.c[0].stride + x];
        return;
    }
    int v[4] = {0, 0, 0, 0};
    for (unsigned k = 0; k < n; k++) v[k] = upsample_at(p.planes + img.c[k].plane_off, img.c[k], x, y);
    uint8_t* o = out + ((size_t)y * W + x) * n;
    switch (img.cc) {
    case CC_RGB:  // src/decoder.rs:1391-1404
        o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
        break;
    case CC_YCBCR: {  // src/decoder.rs:1406-1437
        int r, g, b;
        if (x < img.ssse3_pixels) ycbcr_ssse3(v[0], v[1], v[2], r, g, b);
        else ycbcr_scalar(v[0], v[1], v[2], r, g, b);
        o[0] = clamp_u8(r); o[1] = clamp_u8(g); o[2] = clamp_u8(b);
        break;
    }
    case CC_YCCK: {  // src/decoder.rs:1439-1456 (always the scalar formula)
        int r, g, b;
        ycbcr_scalar(v[0], v[1], v[2], r, g, b);
        o[0] = clamp_u8(r); o[1] = clamp_u8(g); o[2] = clamp_u8(b); o[3] = (uint8_t)(255 - v[3]);
        break;
    }
    case CC_CMYK:  // src/decoder.rs:1458-1474
        o[0] = (uint8_t)(255 - v[0]); o[1] = (uint8_t)(255 - v[1]); o[2] = (uint8_t)(255 - v[2]); o[3] = (uint8_t)(255 - v[3]);
        break;
    default:  // color_no_convert, src/decoder.rs:1476-1484: components planar inside each row
        for (unsigned k = 0; k < n; k++) out[(size_t)y * W * n + (size_t)k * W + x] = (uint8_t)v[k];
        break;
    }
}

// ---------------------------------------------------------------------------------------------
// 4:2:0 YCbCr fast path.  Thread = 16 output pixels x the output row pair (2p-1, 2p): both rows
// blend the same two chroma rows (p-1, p) with swapped weights, so every chroma byte is read once.
// Preconditions (checked by the planner): ncomp 3, comp0 H1V1, comps 1,2 H2V2, scalar arithmetic, strides and
// offsets 16-byte (luma) / 8-byte (chroma) aligned.  Any width: rows whose start is not 16-byte aligned
// (width % 16 != 0) fall back to 32-bit or byte stores inside store_words, the last group of a row is ragged.
// grid.x = ceil(G/128) * P, G = width/16, P = height/2 + 1; grid.y = image
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// The kernel: a thread walks K2_RP consecutive row pairs of its 16-pixel column group (K2_RP = 1 is the
// default: measured fastest because it keeps 8 CTAs = 32 warps per SM; 4 reuses chroma rows in registers
// but halves the occupancy -- profiles/sweep_*.jsonl).
// grid = (ceil(G/128), ceil(P/K2_RP), images), G = width/16, P = height/2 + 1
// ---------------------------------------------------------------------------------------------
int g_k2_mode = -1;

struct ChromaRow {  // one chroma row segment: samples i0..i0+7 plus the clamped halo samples
    unsigned lo, hi, L, R;
};

// in_w = samples in the row: in the last group of a ragged row the bytes at i >= in_w (block padding) are replaced by
// the last valid sample, which is exactly the reference's edge rule (out[2 in_w - 1] uses t[in_w-1] alone)
template <bool RAGGED>
__device__ __forceinline__ ChromaRow load_chroma_row(const uint8_t* row, unsigned i0, unsigned iL, unsigned iR, unsigned in_w) {
    ChromaRow c;
    uint2 v = __ldg(reinterpret_cast<const uint2*>(row + i0));
    if (RAGGED && i0 + 8u > in_w) v = replicate_last_sample(v, in_w - i0);  // last group of a ragged row only
    c.lo = v.x;
    c.hi = v.y;
    c.L = __ldg(row + iL);
    c.R = __ldg(row + iR);
    return c;
}

// RAGGED: images whose chroma width is not a multiple of 8 (the last 8-sample window of a row then needs the
// edge sample replicated); a separate instantiation so that the common case keeps 64 registers / 8 CTAs per SM.
template <unsigned K2_RP, int MINB, bool RAGGED, bool SSSE3 = false>
__global__ void __launch_bounds__(128, MINB) k2_ycbcr420(K2Params p, unsigned first) {
    const DevImage& img = p.images[first + blockIdx.z];
    if (RAGGED ? img.path != K2_PATH_420R : (img.path != K2_PATH_420 && !(img.path == K2_PATH_420T && (p.flags & K2_FLAG_LDG_TAKES_420T)))) return;
    const unsigned g = blockIdx.x * 128u + threadIdx.x;
    const unsigned W = img.width, H = img.height;
    const unsigned npairs = H / 2 + 1;
    const unsigned p0 = blockIdx.y * K2_RP;
    if (p0 >= npairs || g * 16u >= W) return;
    const unsigned p1 = min(p0 + K2_RP, npairs);
    const unsigned in_w = img.c[1].in_w, in_h = img.c[1].in_h;
    const unsigned sy = img.c[0].stride, sb = img.c[1].stride, sr = img.c[2].stride;
    const uint8_t* yplane = p.planes + img.c[0].plane_off + g * 16u;
    const uint8_t* bplane = p.planes + img.c[1].plane_off;
    const uint8_t* rplane = p.planes + img.c[2].plane_off;
    uint8_t* out = p.out + img.out_off + (size_t)g * 48u;
    const unsigned i0 = g * 8u;
    const unsigned iL = i0 > 0 ? i0 - 1 : 0, iR = min(i0 + 8u, in_w - 1);
    const YccRegs sixteen = make_ycc_regs(p.sixteen, true);
    const unsigned npx = min(16u, W - g * 16u);
    const unsigned nss = SSSE3 ? min(16u, max(img.ssse3_pixels, g * 16u) - g * 16u) : 0u;

    // chroma row A of the first pair
    const unsigned rA0 = p0 > 0 ? p0 - 1 : 0;
    ChromaRow ba = load_chroma_row<RAGGED>(bplane + (size_t)rA0 * sb, i0, iL, iR, in_w);
    ChromaRow ra = load_chroma_row<RAGGED>(rplane + (size_t)rA0 * sr, i0, iL, iR, in_w);
    for (unsigned pr = p0; pr < p1; pr++) {
        const unsigned rB = min(pr, in_h - 1);
        // all loads of the iteration up front (independent, clamped rows) so their latencies overlap
        const unsigned y_odd = pr > 0 ? 2 * pr - 1 : 0, y_even = min(2 * pr, H - 1);
        const ChromaRow bb = load_chroma_row<RAGGED>(bplane + (size_t)rB * sb, i0, iL, iR, in_w);
        const ChromaRow rb = load_chroma_row<RAGGED>(rplane + (size_t)rB * sr, i0, iL, iR, in_w);
        const uint4 yv_odd = __ldg(reinterpret_cast<const uint4*>(yplane + (size_t)y_odd * sy));
        const uint4 yv_even = __ldg(reinterpret_cast<const uint4*>(yplane + (size_t)y_even * sy));
        Chroma16 cb, cr;
        h2v2_16(ba.lo, ba.hi, ba.L, ba.R, bb.lo, bb.hi, bb.L, bb.R, cb);
        h2v2_16(ra.lo, ra.hi, ra.L, ra.R, rb.lo, rb.hi, rb.L, rb.R, cr);
        if (pr > 0) ycbcr_store16<SSSE3>(yv_odd, cb.odd, cr.odd, out + (size_t)y_odd * W * 3u, sixteen, npx, nss);      // output row 2p-1
        if (2 * pr < H) ycbcr_store16<SSSE3>(yv_even, cb.even, cr.even, out + (size_t)y_even * W * 3u, sixteen, npx, nss);  // output row 2p
        ba = bb;
        ra = rb;
    }
}

// ---------------------------------------------------------------------------------------------
// 4:4:4 YCbCr fast path: thread = 16 pixels of one row.
// grid.x = ceil(G/128) * height, grid.y = image
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// 4:2:0 YCbCr, bulk-copy fed (the default for aligned images).  Same arithmetic as k2_ycbcr420; what changes
// is how bytes reach the SM.  Persistent CTAs each own a contiguous run of row pairs of 2048-pixel strips (work list
// K2Strip[] built by the planner, so heterogeneous batches cost no empty CTAs).  Warp 4 keeps a K2T_STAGES-deep
// shared-memory ring full with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx): per row pair the two luma
// rows and ONE new chroma row per component -- the other chroma row is the previous pair's, still in the ring (its
// slot is released one iteration late).  Warps 0-3 convert out of shared memory.
// Measured on B200 (profiles/r01_k2_bulk_vs_ldg.md): 0.917 ms vs 0.923 ms per 512 images for the load/store kernel --
// both sit at ~68 % ALU-pipe and ~70 % FMA-heavy-pipe utilisation, i.e. the conversion is bound by integer issue,
// not by how the bytes arrive.
// ---------------------------------------------------------------------------------------------
constexpr unsigned K2T_STAGES = 6;
constexpr unsigned K2T_THREADS = 160;                     // warps 0-3 convert, warp 4 feeds the ring
constexpr unsigned K2T_STRIP = 2048;                      // pixels per strip = 128 threads x 16
constexpr unsigned K2T_CROW = K2T_STRIP / 2 + 32;         // chroma row buffer: 16 B pad | 1024 samples | 16 B pad
constexpr unsigned K2T_OFF_Y0 = 0, K2T_OFF_Y1 = K2T_STRIP, K2T_OFF_BB = 2 * K2T_STRIP, K2T_OFF_RB = K2T_OFF_BB + K2T_CROW,
                   K2T_OFF_BA = K2T_OFF_RB + K2T_CROW, K2T_OFF_RA = K2T_OFF_BA + K2T_CROW;
constexpr unsigned K2T_STAGE_BYTES = (K2T_OFF_RA + K2T_CROW + 127u) / 128u * 128u;

struct K2TProd {  // what the producer needs per strip; lives in shared memory, touched by one thread only
    const uint8_t* y;   // luma row 0 at the strip's first pixel
    const uint8_t* cb;  // chroma rows 0 at the first copied sample (c_lo)
    const uint8_t* cr;
    unsigned sy, sc;    // plane strides
    unsigned H, in_h;
    unsigned ybytes, cbytes, cdst;  // bytes per luma / chroma row copy, offset of the chroma copy in its row buffer
};

__global__ void __launch_bounds__(K2T_THREADS, 4) k2_ycbcr420_tma(K2Params p, const K2Strip* __restrict__ strips, unsigned nstrips,
                                                                  unsigned item_base, unsigned total_items) {
    extern __shared__ __align__(128) uint8_t k2t_smem[];
    __shared__ __align__(8) unsigned long long full_bar[K2T_STAGES];
    __shared__ __align__(8) unsigned long long empty_bar[K2T_STAGES];
    __shared__ K2TProd prod_state;
    const unsigned smem = smem_u32(k2t_smem);
    const unsigned tid = threadIdx.x, lane = tid & 31;
    const unsigned w_begin = item_base + (unsigned)(((unsigned long long)blockIdx.x * total_items) / gridDim.x);
    const unsigned w_end = item_base + (unsigned)(((unsigned long long)(blockIdx.x + 1) * total_items) / gridDim.x);
    const unsigned n = w_end - w_begin;
    if (n == 0) return;

    if (tid == 0) {
        for (unsigned st = 0; st < K2T_STAGES; st++) {
            mbar_init(smem_u32(&full_bar[st]), 1);
            mbar_init(smem_u32(&empty_bar[st]), 4);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // strip containing item w_begin: binary search on first_item
    unsigned lo = 0, hi = nstrips - 1;
    while (lo < hi) {
        const unsigned mid = (lo + hi + 1) / 2;
        if (__ldg(&strips[mid].first_item) <= w_begin) lo = mid; else hi = mid - 1;
    }
    const unsigned first_pair = w_begin - __ldg(&strips[lo].first_item);

    if (tid >= 128) {
        // ---- producer (lane 0 of warp 4): cursor + per-strip state, refreshed only when the cursor enters a new strip.
        // A consumer warp doing this on the side was measured 1.5x slower than its siblings, which then idled. ----
        if (lane != 0) return;
        unsigned p_strip = lo, p_pair = first_pair, p_npairs = 0;
        bool p_new = true;
        unsigned slot_index = 0, e_phase = 0;
        for (unsigned j = 0; j < n; j++) {
            // tile j replaces tile j - STAGES, which is released one iteration late (end of iteration j - STAGES + 1)
            if (j >= K2T_STAGES) mbar_wait(smem_u32(&empty_bar[slot_index]), e_phase);
            if (p_new) {
                const K2Strip sp = strips[p_strip];
                const DevImage& img = p.images[sp.image];
                const unsigned wpx = min(K2T_STRIP, img.width - sp.x0);
                const unsigned cs = sp.x0 / 2u;
                const unsigned c_lo = cs >= 16u ? cs - 16u : 0u;
                const unsigned c_hi = min(img.c[1].stride, cs + K2T_STRIP / 2u + 16u);
                prod_state.y = p.planes + img.c[0].plane_off + sp.x0;
                prod_state.cb = p.planes + img.c[1].plane_off + c_lo;
                prod_state.cr = p.planes + img.c[2].plane_off + c_lo;
                prod_state.sy = img.c[0].stride;
                prod_state.sc = img.c[1].stride;
                prod_state.H = img.height;
                prod_state.in_h = img.c[1].in_h;
                prod_state.ybytes = (wpx + 15u) & ~15u;
                prod_state.cbytes = c_hi - c_lo;
                prod_state.cdst = c_lo + 16u - cs;
                p_npairs = sp.npairs;
                p_new = false;
            }
            const unsigned pr = p_pair, ybytes = prod_state.ybytes, cbytes = prod_state.cbytes, sc = prod_state.sc;
            const bool need_a = pr > 0 && j == 0;  // first tile of this CTA: chroma row p-1 is not in the ring
            const bool has_even = 2u * pr < prod_state.H;
            const unsigned slot = smem + slot_index * K2T_STAGE_BYTES;
            const unsigned bar = smem_u32(&full_bar[slot_index]);
            mbar_expect_tx(bar, (need_a ? 4u : 2u) * cbytes + (pr > 0 ? ybytes : 0u) + (has_even ? ybytes : 0u));
            const uint8_t* yp = prod_state.y + (size_t)(2u * pr) * prod_state.sy;
            if (pr > 0) bulk_load_1d(slot + K2T_OFF_Y0, yp - prod_state.sy, ybytes, bar);
            if (has_even) bulk_load_1d(slot + K2T_OFF_Y1, yp, ybytes, bar);
            const size_t rb = (size_t)min(pr, prod_state.in_h - 1u) * sc;
            const unsigned cdst = slot + prod_state.cdst;
            bulk_load_1d(cdst + K2T_OFF_BB, prod_state.cb + rb, cbytes, bar);
            bulk_load_1d(cdst + K2T_OFF_RB, prod_state.cr + rb, cbytes, bar);
            if (need_a) {
                const size_t ra = (size_t)(pr - 1u) * sc;
                bulk_load_1d(cdst + K2T_OFF_BA, prod_state.cb + ra, cbytes, bar);
                bulk_load_1d(cdst + K2T_OFF_RA, prod_state.cr + ra, cbytes, bar);
            }
            if (++p_pair == p_npairs) {
                p_pair = 0;
                p_strip++;
                p_new = true;
            }
            if (++slot_index == K2T_STAGES) {
                slot_index = 0;
                if (j >= K2T_STAGES) e_phase ^= 1u;
            }
        }
        return;
    }

    // ---- consumers (warps 0-3): cursor + per-strip constants of this thread, refreshed only on a strip change ----
    unsigned c_strip = lo, c_pair = first_pair, c_npairs = 0, c_H = 0, c_row_bytes = 0, c_npx = 0, c_oL = 0, c_oR = 0;
    uint8_t* c_out = nullptr;
    bool c_new = true;
    const unsigned oM = 16u + tid * 8u;
    const YccRegs ycc = make_ycc_regs(p.sixteen, true);
    unsigned stage = 0, phase = 0;  // slot / parity of the tile being consumed
    for (unsigned it = 0; it < n; it++) {
        if (c_new) {
            const K2Strip sp = strips[c_strip];
            const DevImage& img = p.images[sp.image];
            const unsigned W = img.width, wpx = min(K2T_STRIP, W - sp.x0);
            const unsigned cs = sp.x0 / 2u, gi = cs + tid * 8u;  // global index of this thread's first chroma sample
            c_H = img.height;
            c_row_bytes = W * 3u;
            c_npx = tid * 16u < wpx ? min(16u, wpx - tid * 16u) : 0u;
            c_oL = gi > 0 ? oM - 1u : oM;                                   // clamped halo samples, as buffer offsets
            c_oR = 16u + (min(gi + 8u, img.c[1].in_w - 1u) - cs);
            c_out = p.out + img.out_off + ((size_t)sp.x0 + tid * 16u) * 3u;
            c_npairs = sp.npairs;
            c_new = false;
        }
        const unsigned pr = c_pair;
        const unsigned slot = smem + stage * K2T_STAGE_BYTES;
        // chroma row A (= max(p-1,0)): row B of this slot when p == 0, this slot's A buffers for the CTA's first tile,
        // otherwise row B of the previous pair, which still sits in the previous slot
        const unsigned prev_stage = stage == 0 ? K2T_STAGES - 1u : stage - 1u;
        const unsigned prev = smem + prev_stage * K2T_STAGE_BYTES;
        const unsigned a_b = pr == 0 ? slot + K2T_OFF_BB : (it == 0 ? slot + K2T_OFF_BA : prev + K2T_OFF_BB);
        const unsigned a_r = a_b + K2T_CROW;
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        if (c_npx) {
            Chroma16 cb, cr;
            {
                const uint2 av = lds64(a_b + oM), bv = lds64(slot + K2T_OFF_BB + oM);
                h2v2_16(av.x, av.y, lds8(a_b + c_oL), lds8(a_b + c_oR), bv.x, bv.y, lds8(slot + K2T_OFF_BB + c_oL), lds8(slot + K2T_OFF_BB + c_oR), cb);
            }
            {
                const uint2 av = lds64(a_r + oM), bv = lds64(slot + K2T_OFF_RB + oM);
                h2v2_16(av.x, av.y, lds8(a_r + c_oL), lds8(a_r + c_oR), bv.x, bv.y, lds8(slot + K2T_OFF_RB + c_oL), lds8(slot + K2T_OFF_RB + c_oR), cr);
            }
            uint8_t* row = c_out + (size_t)(2u * pr) * c_row_bytes;
            if (pr > 0) ycbcr_store16(lds128(slot + K2T_OFF_Y0 + tid * 16u), cb.odd, cr.odd, row - c_row_bytes, ycc, c_npx);
            if (2u * pr < c_H) ycbcr_store16(lds128(slot + K2T_OFF_Y1 + tid * 16u), cb.even, cr.even, row, ycc, c_npx);
        }
        // release the PREVIOUS slot: its chroma rows were this iteration's row A
        __syncwarp();
        if (lane == 0 && it > 0) mbar_arrive(smem_u32(&empty_bar[prev_stage]));
        if (++stage == K2T_STAGES) { stage = 0; phase ^= 1u; }
        if (++c_pair == c_npairs) {
            c_pair = 0;
            c_strip++;
            c_new = true;
        }
    }
}

// 16 bytes of a plane row whose stride is a multiple of 8 (not necessarily 16): two aligned 8-byte loads, the
// second only when it still lies inside the row
__device__ __forceinline__ uint4 load_row16(const uint8_t* row, unsigned g, unsigned stride) {
    const uint2* src = reinterpret_cast<const uint2*>(row + g * 16u);
    const uint2 lo = __ldg(src);
    uint2 hi = make_uint2(0u, 0u);
    if (g * 16u + 8u < stride) hi = __ldg(src + 1);
    return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

template <bool SSSE3>
__global__ void __launch_bounds__(128) k2_ycbcr444(K2Params p, unsigned first, unsigned gchunks) {
    const DevImage& img = p.images[first + blockIdx.y];
    if (img.path != K2_PATH_444) return;
    const unsigned y = blockIdx.x / gchunks;
    const unsigned g = (blockIdx.x % gchunks) * 128u + threadIdx.x;
    const unsigned W = img.width;
    if (y >= img.height || g * 16u >= W) return;
    const uint4 yv = load_row16(p.planes + img.c[0].plane_off + (size_t)y * img.c[0].stride, g, img.c[0].stride);
    const uint4 bv = load_row16(p.planes + img.c[1].plane_off + (size_t)y * img.c[1].stride, g, img.c[1].stride);
    const uint4 rv = load_row16(p.planes + img.c[2].plane_off + (size_t)y * img.c[2].stride, g, img.c[2].stride);
    // byte ^ 0x80 read as a signed byte is byte - 128: one LOP3 per word, the sign extension rides in the PRMT
    const unsigned bw[4] = {bv.x ^ 0x80808080u, bv.y ^ 0x80808080u, bv.z ^ 0x80808080u, bv.w ^ 0x80808080u};
    const unsigned rw[4] = {rv.x ^ 0x80808080u, rv.y ^ 0x80808080u, rv.z ^ 0x80808080u, rv.w ^ 0x80808080u};
    int cb[16], cr[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        cb[k] = (int)prmt(bw[k >> 2], 0u, 0x8880u + 0x1111u * (unsigned)(k & 3));
        cr[k] = (int)prmt(rw[k >> 2], 0u, 0x8880u + 0x1111u * (unsigned)(k & 3));
    }
    ycbcr_store16<SSSE3>(yv, cb, cr, p.out + img.out_off + ((size_t)y * W + g * 16u) * 3u, make_ycc_regs(p.sixteen, true), min(16u, W - g * 16u),
                         SSSE3 ? min(16u, max(img.ssse3_pixels, g * 16u) - g * 16u) : 0u);
}

// ---------------------------------------------------------------------------------------------
// 4:2:2 YCbCr (H2V1 chroma, src/upsampler.rs:134-163; the layout of MJPEG / camera files): thread = 16 pixels of one row,
// 8 chroma samples + the two clamped halo samples per component.  grid.x = ceil(G/128) * height, grid.y = image
// ---------------------------------------------------------------------------------------------
template <bool SSSE3>
__global__ void __launch_bounds__(128) k2_ycbcr422(K2Params p, unsigned first, unsigned gchunks) {
    const DevImage& img = p.images[first + blockIdx.y];
    if (img.path != K2_PATH_422) return;
    const unsigned y = blockIdx.x / gchunks;
    const unsigned g = (blockIdx.x % gchunks) * 128u + threadIdx.x;
    const unsigned W = img.width;
    if (y >= img.height || g * 16u >= W) return;
    const uint4 yv = load_row16(p.planes + img.c[0].plane_off + (size_t)y * img.c[0].stride, g, img.c[0].stride);
    const unsigned in_w = img.c[1].in_w, i0 = g * 8u;
    const unsigned iL = i0 > 0 ? i0 - 1u : 0u, iR = min(i0 + 8u, in_w - 1u);
    int cb[16], cr[16];
#pragma unroll
    for (int c = 1; c <= 2; c++) {
        const uint8_t* row = p.planes + img.c[c].plane_off + (size_t)y * img.c[c].stride;
        uint2 v = __ldg(reinterpret_cast<const uint2*>(row + i0));
        if (i0 + 8u > in_w) v = replicate_last_sample(v, in_w - i0);  // block padding := last valid sample (the edge rule)
        h2v1_16(v.x, v.y, __ldg(row + iL), __ldg(row + iR), c == 1 ? cb : cr);
    }
    ycbcr_store16<SSSE3>(yv, cb, cr, p.out + img.out_off + ((size_t)y * W + g * 16u) * 3u, make_ycc_regs(p.sixteen, true), min(16u, W - g * 16u),
                         SSSE3 ? min(16u, max(img.ssse3_pixels, g * 16u) - g * 16u) : 0u);
}

// ---------------------------------------------------------------------------------------------
// 4:4:0 YCbCr (H1V2 chroma, src/upsampler.rs:165-189: what a losslessly rotated 4:2:2 file becomes): thread = 16 pixels of one
// row, 16 samples of the near and of the far chroma row per component (far = the row above for even output rows, below for
// odd ones, clamped: the f32 expression of upsampler.rs:174-180 as integer min / max).  grid.x = ceil(G/128) * height, grid.y = image
// ---------------------------------------------------------------------------------------------
template <bool SSSE3>
__global__ void __launch_bounds__(128) k2_ycbcr440(K2Params p, unsigned first, unsigned gchunks) {
    const DevImage& img = p.images[first + blockIdx.y];
    if (img.path != K2_PATH_440) return;
    const unsigned y = blockIdx.x / gchunks;
    const unsigned g = (blockIdx.x % gchunks) * 128u + threadIdx.x;
    const unsigned W = img.width;
    if (y >= img.height || g * 16u >= W) return;
    const uint4 yv = load_row16(p.planes + img.c[0].plane_off + (size_t)y * img.c[0].stride, g, img.c[0].stride);
    const unsigned k = y >> 1;
    int cb[16], cr[16];
#pragma unroll
    for (int c = 1; c <= 2; c++) {
        const unsigned f = (y & 1u) ? min(k + 1u, img.c[c].in_h - 1u) : (k > 0u ? k - 1u : 0u);
        const uint8_t* plane = p.planes + img.c[c].plane_off;
        const uint4 nv = load_row16(plane + (size_t)k * img.c[c].stride, g, img.c[c].stride);
        const uint4 fv = load_row16(plane + (size_t)f * img.c[c].stride, g, img.c[c].stride);
        h1v2_16(nv, fv, c == 1 ? cb : cr);
    }
    ycbcr_store16<SSSE3>(yv, cb, cr, p.out + img.out_off + ((size_t)y * W + g * 16u) * 3u, make_ycc_regs(p.sixteen, true), min(16u, W - g * 16u),
                         SSSE3 ? min(16u, max(img.ssse3_pixels, g * 16u) - g * 16u) : 0u);
}

// ---------------------------------------------------------------------------------------------
// Every component at full resolution and nothing to compute but bytes: RGB (src/decoder.rs:1391-1404), CMYK (1458-1474),
// YCCK (1439-1456), ColorTransform::None (1476-1484).  Thread = 16 pixels of one row: 16-byte loads per component,
// interleave with PRMT, 16-byte stores.  grid.x = ceil(G/128) * height, grid.y = image
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k2_bytes(K2Params p, unsigned first, unsigned gchunks) {
    const DevImage& img = p.images[first + blockIdx.y];
    if (img.path != K2_PATH_BYTES) return;
    const unsigned y = blockIdx.x / gchunks;
    const unsigned g = (blockIdx.x % gchunks) * 128u + threadIdx.x;
    const unsigned W = img.width, n = img.ncomp;
    if (y >= img.height || g * 16u >= W) return;
    const unsigned npx = min(16u, W - g * 16u);
    unsigned v[4][4];
#pragma unroll
    for (unsigned k = 0; k < 4; k++) {
        uint4 t = make_uint4(0u, 0u, 0u, 0u);
        if (k < n) t = load_row16(p.planes + img.c[k].plane_off + (size_t)y * img.c[k].stride, g, img.c[k].stride);
        v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w;
    }
    uint8_t* const row = p.out + img.out_off + (size_t)y * W * n;
    if (img.cc == CC_NOCONVERT) {  // the components side by side inside every row
        for (unsigned k = 0; k < n; k++) store_words<4>(row + (size_t)k * W + g * 16u, v[k], npx);
        return;
    }
    if (img.cc == CC_RGB) {
        unsigned ow[12];
#pragma unroll
        for (int w = 0; w < 4; w++) {  // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
            const unsigned r = v[0][w], gg = v[1][w], b = v[2][w];
            ow[3 * w + 0] = prmt(prmt(r, gg, 0x0140), b, 0x2410);
            ow[3 * w + 1] = prmt(prmt(r, gg, 0x0625), b, 0x2150);
            ow[3 * w + 2] = prmt(prmt(r, gg, 0x0073), b, 0x7106);
        }
        store_words<12>(row + (size_t)g * 48u, ow, 3u * npx);
        return;
    }
    unsigned ow[16];
    if (img.cc == CC_CMYK) {  // 255 - x on every byte, then C M Y K per pixel
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const unsigned c = ~v[0][w], m = ~v[1][w], yy = ~v[2][w], k = ~v[3][w];
            const unsigned cm01 = prmt(c, m, 0x5140), cm23 = prmt(c, m, 0x7362), yk01 = prmt(yy, k, 0x5140), yk23 = prmt(yy, k, 0x7362);
            ow[4 * w + 0] = prmt(cm01, yk01, 0x5410);
            ow[4 * w + 1] = prmt(cm01, yk01, 0x7632);
            ow[4 * w + 2] = prmt(cm23, yk23, 0x5410);
            ow[4 * w + 3] = prmt(cm23, yk23, 0x7632);
        }
    } else {  // CC_YCCK: YCbCr -> RGB (always the scalar formula, src/decoder.rs:1447), K inverted
        const YccRegs ycc = make_ycc_regs(p.sixteen, true);
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const unsigned bw = v[1][w] ^ 0x80808080u, rw = v[2][w] ^ 0x80808080u, kw = ~v[3][w];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int y16 = (int)prmt(v[0][w], 0u, 0x4044u | ((unsigned)q << 8));
                const int cbm = (int)prmt(bw, 0u, 0x8880u + 0x1111u * (unsigned)q), crm = (int)prmt(rw, 0u, 0x8880u + 0x1111u * (unsigned)q);
                const int kk = (int)prmt(kw, 0u, 0x4440u | (unsigned)q);
                int r, gg, b;
                ycbcr_scalar_y16(y16, cbm, crm, r, gg, b, ycc);
                ow[4 * w + q] = pack_sat_u8(gg, r, pack_sat_u8(kk, b, 0u));
            }
        }
    }
    store_words<16>(row + (size_t)g * 64u, ow, 4u * npx);
}

// ---------------------------------------------------------------------------------------------
// 1-component images: compact the plane from stride block_w*dct_scale to `width` (src/decoder.rs:1310-1332).
// Thread = 16 bytes of one row.  grid = (ceil(ceil(W/16)/128), height, images)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k2_gray_crop(K2Params p, unsigned first) {
    const DevImage& img = p.images[first + blockIdx.z];
    if (img.path != K2_PATH_GRAY) return;
    const unsigned g = blockIdx.x * 128u + threadIdx.x, y = blockIdx.y;
    const unsigned W = img.width;
    if (y >= img.height || g * 16u >= W) return;
    const uint4 v = load_row16(p.planes + img.c[0].plane_off + (size_t)y * img.c[0].stride, g, img.c[0].stride);
    const unsigned ow[4] = {v.x, v.y, v.z, v.w};
    store_words<4>(p.out + img.out_off + (size_t)y * W + g * 16u, ow, min(16u, W - g * 16u));
}

// ---------------------------------------------------------------------------------------------
cudaError_t launch_k2_generic(const K2Params& p, unsigned first, unsigned count, unsigned max_w, unsigned max_h,
                              cudaStream_t stream) {
    if (count == 0 || max_w == 0 || max_h == 0) return cudaSuccess;
    const unsigned xchunks = (max_w + 255u) / 256u;
    dim3 grid(xchunks * max_h, count);
    k2_generic<<<grid, 256, 0, stream>>>(p, first, xchunks);
    return cudaGetLastError();
}
int k2_mode() {  // experiment knob (profiling only): 0 = default, 1 = load/store kernels only, 4 = 4 row pairs per thread
    if (g_k2_mode < 0) {
        const char* e = getenv("B200JPG_K2_MODE");
        g_k2_mode = e ? atoi(e) : 0;
    }
    return g_k2_mode;
}
cudaError_t launch_k2_420(const K2Params& p, unsigned first, unsigned count, unsigned max_w, unsigned max_h, bool ragged,
                          cudaStream_t stream) {
    if (count == 0 || max_w == 0 || max_h == 0) return cudaSuccess;
    const unsigned npairs = max_h / 2u + 1u;
    const unsigned rp = k2_mode() == 4 ? 4u : 1u;
    dim3 grid(((max_w + 15u) / 16u + 127u) / 128u, (npairs + rp - 1u) / rp, count);
    if (p.flags & K2_FLAG_SSSE3) {  // the x86 build's colour arithmetic on the first (W / 8 - 1) * 8 pixels of every row
        if (ragged) k2_ycbcr420<1, 6, true, true><<<dim3(grid.x, npairs, count), 128, 0, stream>>>(p, first);
        else k2_ycbcr420<1, 6, false, true><<<dim3(grid.x, npairs, count), 128, 0, stream>>>(p, first);
        return cudaGetLastError();
    }
    if (ragged) k2_ycbcr420<1, 7, true><<<dim3(grid.x, npairs, count), 128, 0, stream>>>(p, first);
    else if (rp == 4) k2_ycbcr420<4, 5, false><<<grid, 128, 0, stream>>>(p, first);
    else k2_ycbcr420<1, 8, false><<<grid, 128, 0, stream>>>(p, first);
    return cudaGetLastError();
}
cudaError_t launch_k2_420_tma(const K2Params& p, const K2Strip* strips, unsigned nstrips, unsigned item_base, unsigned total_items,
                              int num_sms, cudaStream_t stream) {
    if (nstrips == 0 || total_items == 0) return cudaSuccess;
    const size_t smem_bytes = (size_t)K2T_STAGES * K2T_STAGE_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k2_ycbcr420_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    unsigned grid = (unsigned)num_sms * 4u;
    if (grid > total_items) grid = total_items;
    k2_ycbcr420_tma<<<grid, K2T_THREADS, smem_bytes, stream>>>(p, strips, nstrips, item_base, total_items);
    return cudaGetLastError();
}
// ---------------------------------------------------------------------------------------------
// K3: interleaved RGB8 -> the layout of an on-GPU consumer (b200jpg_batch_format_device).  Thread = 16 pixels of one
// row: three 16-byte loads, then 3 x 16-byte stores (planar u8) or 12 / 3 x 4 float4 stores.  HBM-bound by design.
// ---------------------------------------------------------------------------------------------
struct K3Params {
    const DevImage* images;
    const uint8_t* src;
    uint8_t* dst;
    float scale[3], bias[3];
    int format;
};
__global__ void __launch_bounds__(128) k3_format(K3Params p, unsigned first, unsigned gchunks) {
    const DevImage& img = p.images[first + blockIdx.y];
    if (img.ncomp != 3 || img.width == 0) return;
    const unsigned y = blockIdx.x / gchunks;
    const unsigned g = (blockIdx.x % gchunks) * 128u + threadIdx.x;
    const unsigned W = img.width, H = img.height;
    if (y >= H || g * 16u >= W) return;
    const unsigned npx = min(16u, W - g * 16u);
    const uint8_t* src = p.src + img.out_off + ((size_t)y * W + g * 16u) * 3u;
    unsigned char px[48];
    if (npx == 16u && ((uintptr_t)src & 15u) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) reinterpret_cast<uint4*>(px)[k] = __ldg(reinterpret_cast<const uint4*>(src) + k);
    } else {
        for (unsigned k = 0; k < 48u; k++) px[k] = k < 3u * npx ? src[k] : 0;
    }
    const size_t plane = (size_t)W * H;
    if (p.format == 1) {  // CHW uint8
        uint8_t* base = p.dst + img.out_off + (size_t)y * W + g * 16u;
        for (int c = 0; c < 3; c++) {
            unsigned ow[4];
#pragma unroll
            for (int w = 0; w < 4; w++)
                ow[w] = px[12 * w + c] | ((unsigned)px[12 * w + 3 + c] << 8) | ((unsigned)px[12 * w + 6 + c] << 16) | ((unsigned)px[12 * w + 9 + c] << 24);
            store_words<4>(base + (size_t)c * plane, ow, npx);
        }
        return;
    }
    float* fdst = reinterpret_cast<float*>(p.dst + 4u * img.out_off);
    if (p.format == 2) {  // HWC float32
        float* o = fdst + ((size_t)y * W + g * 16u) * 3u;
        for (unsigned k = 0; k < 3u * npx; k++) o[k] = fmaf((float)px[k], p.scale[k % 3u], p.bias[k % 3u]);
    } else {              // CHW float32
        for (int c = 0; c < 3; c++) {
            float* o = fdst + (size_t)c * plane + (size_t)y * W + g * 16u;
            for (unsigned k = 0; k < npx; k++) o[k] = fmaf((float)px[3 * k + c], p.scale[c], p.bias[c]);
        }
    }
}
cudaError_t launch_k3_format(const DevImage* images, unsigned first, unsigned count, unsigned max_w, unsigned max_h, const void* src, void* dst,
                             int format, const float scale[3], const float bias[3], cudaStream_t stream) {
    if (count == 0 || max_w == 0 || max_h == 0) return cudaSuccess;
    K3Params p;
    p.images = images;
    p.src = (const uint8_t*)src;
    p.dst = (uint8_t*)dst;
    for (int c = 0; c < 3; c++) {
        p.scale[c] = scale ? scale[c] : 1.0f;
        p.bias[c] = bias ? bias[c] : 0.0f;
    }
    p.format = format;
    const unsigned gchunks = ((max_w + 15u) / 16u + 127u) / 128u;
    k3_format<<<dim3(gchunks * max_h, count), 128, 0, stream>>>(p, first, gchunks);
    return cudaGetLastError();
}

// kernels whose grid is (ceil(G/128) * max_h, images): path selects k2_ycbcr444 / k2_ycbcr422 / k2_ycbcr440 / k2_bytes
cudaError_t launch_k2_rows16(unsigned path, const K2Params& p, unsigned first, unsigned count, unsigned max_w, unsigned max_h, cudaStream_t stream) {
    if (count == 0 || max_w == 0 || max_h == 0) return cudaSuccess;
    const unsigned gchunks = ((max_w + 15u) / 16u + 127u) / 128u;
    dim3 grid(gchunks * max_h, count);
    const bool s3 = (p.flags & K2_FLAG_SSSE3) != 0;
    if (path == K2_PATH_422) s3 ? k2_ycbcr422<true><<<grid, 128, 0, stream>>>(p, first, gchunks) : k2_ycbcr422<false><<<grid, 128, 0, stream>>>(p, first, gchunks);
    else if (path == K2_PATH_440) s3 ? k2_ycbcr440<true><<<grid, 128, 0, stream>>>(p, first, gchunks) : k2_ycbcr440<false><<<grid, 128, 0, stream>>>(p, first, gchunks);
    else if (path == K2_PATH_BYTES) k2_bytes<<<grid, 128, 0, stream>>>(p, first, gchunks);
    else s3 ? k2_ycbcr444<true><<<grid, 128, 0, stream>>>(p, first, gchunks) : k2_ycbcr444<false><<<grid, 128, 0, stream>>>(p, first, gchunks);
    return cudaGetLastError();
}
cudaError_t launch_k2_gray(const K2Params& p, unsigned first, unsigned count, unsigned max_w, unsigned max_h, cudaStream_t stream) {
    if (count == 0 || max_w == 0 || max_h == 0) return cudaSuccess;
    dim3 grid(((max_w + 15u) / 16u + 127u) / 128u, max_h, count);
    k2_gray_crop<<<grid, 128, 0, stream>>>(p, first);
    return cudaGetLastError();
}

}  // namespace b200jpg
