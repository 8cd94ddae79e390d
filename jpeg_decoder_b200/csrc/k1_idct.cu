// k1_idct.cu -- K1: dequantise + inverse DCT, coefficient slab -> plane slab (sm_100a).
//
// Replaces dequantize_and_idct_block (reference src/idct.rs:205-239) as driven by
// ImmediateWorker::append_row_immediate (src/worker/immediate.rs:39-60) /
// append_row_locked (src/worker/rayon.rs:71-112), and the SSSE3 variant
// src/arch/ssse3.rs:124-192.  Arithmetic specification: SURVEY.md Appendix A.1/A.2.
//
// Two kernels:
//   k1_idct8_tma      the hot one: scalar arithmetic, 8x8.  Persistent 4-warp CTAs streaming 128-block
//                     tiles (16 KB) through a 3-stage shared-memory ring with 2-D TMA (128B swizzle) +
//                     mbarriers, one thread = one block in registers, IDP.2A dequantisation with the
//                     tables in the constant bank, coalesced 8-byte row stores.
//   k1_idct_generic   every other case (scaled IDCT 4x4/2x2/1x1, SSSE3 arithmetic), plain loads.
//
// Roofline: HBM.  Algorithmic bytes per block = 128 read + dct_scale^2 written (192 at scale 8).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "device_types.h"
#include "kernels.h"
#include "ptx.cuh"
#include "idct_core.cuh"

#define K1_DEFAULT_MODE 8

namespace b200jpg {


// src/idct.rs:456-517
__device__ void idct4x4_scalar(const short* __restrict__ c, const unsigned* __restrict__ q, uint8_t* dst, unsigned stride) {
    unsigned temp[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        unsigned s0 = (unsigned)(int)c[i] * q[i], s1 = (unsigned)(int)c[i + 8] * q[i + 8];
        unsigned s2 = (unsigned)(int)c[i + 16] * q[i + 16], s3 = (unsigned)(int)c[i + 24] * q[i + 24];
        unsigned x0 = (s0 + s2) << 2, x2 = (s0 - s2) << 2;
        unsigned p1 = (s1 + s3) * F2F_0_5411961;
        unsigned t0 = (unsigned)sar(p1 + s3 * F2F_N1_847759065 + 512u, 10);
        unsigned t2 = (unsigned)sar(p1 + s1 * F2F_0_765366865 + 512u, 10);
        temp[i] = x0 + t2;
        temp[i + 12] = x0 - t2;
        temp[i + 4] = x2 + t0;
        temp[i + 8] = x2 - t0;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        unsigned s0 = temp[i * 4], s1 = temp[i * 4 + 1], s2 = temp[i * 4 + 2], s3 = temp[i * 4 + 3];
        unsigned x0 = ((s0 + s2) << 12) + (1u << 16) + (128u << 17);
        unsigned x2 = ((s0 - s2) << 12) + (1u << 16) + (128u << 17);
        unsigned p1 = (s1 + s3) * F2F_0_5411961;
        unsigned t0 = p1 + s3 * F2F_N1_847759065;
        unsigned t2 = p1 + s1 * F2F_0_765366865;
        *reinterpret_cast<unsigned*>(dst + (size_t)i * stride) =
            pack4_sat_u8(sar(x0 + t2, 17), sar(x2 + t0, 17), sar(x2 - t0, 17), sar(x0 - t2, 17));
    }
}

// src/idct.rs:519-553
__device__ void idct2x2_scalar(const short* __restrict__ c, const unsigned* __restrict__ q, uint8_t* dst, unsigned stride) {
    unsigned s00 = (unsigned)(int)c[0] * q[0], s10 = (unsigned)(int)c[8] * q[8];
    unsigned s01 = (unsigned)(int)c[1] * q[1], s11 = (unsigned)(int)c[9] * q[9];
    unsigned x0 = s00 + s10 + 4u + (128u << 3), x2 = s00 - s10 + 4u + (128u << 3);
    unsigned x1 = s01 + s11, x3 = s01 - s11;
    unsigned r0 = pack4_sat_u8(sar(x0 + x1, 3), sar(x0 - x1, 3), 0, 0);
    unsigned r1 = pack4_sat_u8(sar(x2 + x3, 3), sar(x2 - x3, 3), 0, 0);
    *reinterpret_cast<unsigned short*>(dst) = (unsigned short)r0;
    *reinterpret_cast<unsigned short*>(dst + stride) = (unsigned short)r1;
}

// src/idct.rs:555-565 (Wrapping<i32> division truncates toward zero)
__device__ void idct1x1_scalar(const short* __restrict__ c, const unsigned* __restrict__ q, uint8_t* dst) {
    int s0 = (int)((unsigned)(int)c[0] * q[0] + 1024u) / 8;
    dst[0] = (uint8_t)min(max(s0, 0), 255);
}

// ---------------------------------------------------------------------------------------------
// SSSE3 arithmetic, src/arch/ssse3.rs:8-84, 124-192: int16 lanes, saturating add/sub, mulhrs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int sat16(int v) { return min(max(v, -32768), 32767); }
__device__ __forceinline__ int adds16(int a, int b) { return sat16(a + b); }
__device__ __forceinline__ int subs16(int a, int b) { return sat16(a - b); }
__device__ __forceinline__ int mulhrs16(int a, int b) { return (int)(short)((((a * b) >> 14) + 1) >> 1); }

__device__ void idct8_ssse3_lane(int* d, int st) {
    int p2 = d[2 * st], p3 = d[6 * st];
    int p1 = mulhrs16(adds16(p2, p3), 17734);
    int t2 = subs16(subs16(p1, p3), mulhrs16(p3, 27779));
    int t3 = adds16(p1, mulhrs16(p2, 25079));
    p2 = d[0];
    p3 = d[4 * st];
    int t0 = adds16(p2, p3), t1 = subs16(p2, p3);
    int x0 = adds16(t0, t3), x3 = subs16(t0, t3), x1 = adds16(t1, t2), x2 = subs16(t1, t2);
    t0 = d[7 * st];
    t1 = d[5 * st];
    t2 = d[3 * st];
    t3 = d[1 * st];
    p3 = adds16(t0, t2);
    int p4 = adds16(t1, t3);
    p1 = adds16(t0, t3);
    p2 = adds16(t1, t2);
    int p5 = adds16(p3, p4);
    p5 = adds16(p5, mulhrs16(p5, 5763));
    t0 = mulhrs16(t0, 9786);
    t1 = adds16(adds16(t1, t1), mulhrs16(t1, 1741));
    t2 = adds16(adds16(t2, adds16(t2, t2)), mulhrs16(t2, 2383));
    t3 = adds16(t3, mulhrs16(t3, 16427));
    p1 = subs16(p5, mulhrs16(p1, 29490));
    p2 = subs16(subs16(subs16(p5, p2), p2), mulhrs16(p2, 18446));
    p3 = subs16(mulhrs16(p3, -31509), p3);
    p4 = mulhrs16(p4, -12785);
    t3 = adds16(adds16(p1, p4), t3);
    t2 = adds16(adds16(p2, p3), t2);
    t1 = adds16(adds16(p2, p4), t1);
    t0 = adds16(adds16(p1, p3), t0);
    d[0] = adds16(x0, t3);
    d[7 * st] = subs16(x0, t3);
    d[1 * st] = adds16(x1, t2);
    d[6 * st] = subs16(x1, t2);
    d[2 * st] = adds16(x2, t1);
    d[5 * st] = subs16(x2, t1);
    d[3 * st] = adds16(x3, t0);
    d[4 * st] = subs16(x3, t0);
}

__device__ void idct8x8_ssse3(const short* __restrict__ c, const unsigned* __restrict__ q, uint8_t* dst, unsigned stride) {
    int data[64];
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        // _mm_mullo_epi16 then _mm_slli_epi16(.., 3): both wrap at 16 bits (ssse3.rs:151-158)
        unsigned prod = ((unsigned)(unsigned short)c[i] * (q[i] & 0xffffu)) & 0xffffu;
        data[i] = (int)(short)((prod << 3) & 0xffffu);
    }
#pragma unroll 1
    for (int k = 0; k < 8; k++) idct8_ssse3_lane(data + k, 8);  // down the columns (ssse3.rs:162)
#pragma unroll 1
    for (int r = 0; r < 8; r++) idct8_ssse3_lane(data + 8 * r, 1);  // transpose-idct8-transpose = along rows
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        const int* d = data + 8 * r;
        uint2 o;  // adds(.., 8224) >> 6, packus (ssse3.rs:173-185)
        o.x = pack4_sat_u8(adds16(d[0], 8224) >> 6, adds16(d[1], 8224) >> 6, adds16(d[2], 8224) >> 6, adds16(d[3], 8224) >> 6);
        o.y = pack4_sat_u8(adds16(d[4], 8224) >> 6, adds16(d[5], 8224) >> 6, adds16(d[6], 8224) >> 6, adds16(d[7], 8224) >> 6);
        *reinterpret_cast<uint2*>(dst + (size_t)r * stride) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// Generic kernel: one CTA per tile, one thread per block, coefficients read straight from HBM.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K1_TILE) k1_idct_generic(K1Params p, int arith) {
    const DevTile tile = p.tiles[blockIdx.x];
    const unsigned t = threadIdx.x;
    if (t >= tile.nvalid) return;
    const DevComp comp = p.comps[tile.comp];
    unsigned bx = (tile.bxy & 0xffffu) + t, by = tile.bxy >> 16;
    by += bx / comp.block_w;
    bx %= comp.block_w;
    const unsigned* q = p.qtabs + (size_t)comp.qt_index * 64;
    // stage the block into local memory with 16-byte loads
    __align__(16) short c[64];
    const uint4* src = reinterpret_cast<const uint4*>(p.coefs + ((size_t)tile.slab_row + t) * 64);
#pragma unroll
    for (int k = 0; k < 8; k++) reinterpret_cast<uint4*>(c)[k] = __ldg(src + k);
    const unsigned sc = comp.dct_scale;
    uint8_t* dst = p.planes + comp.plane_off + (size_t)by * sc * comp.stride + (size_t)bx * sc;
    // dispatch: src/idct.rs:212-238 ; the SSSE3 variant only replaces the 8x8 kernel (src/idct.rs:247-253)
    if (sc == 8) {
        if (arith == 1) idct8x8_ssse3(c, q, dst, comp.stride);
        else idct8x8_scalar_exact(c, q, dst, comp.stride);
    } else if (sc == 4) {
        idct4x4_scalar(c, q, dst, comp.stride);
    } else if (sc == 2) {
        idct2x2_scalar(c, q, dst, comp.stride);
    } else {
        idct1x1_scalar(c, q, dst);
    }
}

// ---------------------------------------------------------------------------------------------
// The hot kernel.
// ---------------------------------------------------------------------------------------------
constexpr int K1_STAGE_BYTES = K1_TILE * 128;

// Persistent CTAs of 4 warps (4 CTAs = 16 warps per SM), each owning a contiguous range of 128-block
// tiles.  Thread 0 keeps a 3-stage shared-memory ring full with 2-D TMA loads (128B swizzle, so that the
// "thread t reads block t" pattern is bank-conflict free); the slot it refills at the top of an
// iteration was drained one whole tile earlier.  Each thread then does a whole 8x8 block in registers:
//  * 8-bit quantisation tables (every baseline JPEG) are dequantised with IDP.2A straight from the packed
//    int16 pairs: s = dp2a(w, {q_even, 0, 0, q_odd}) -- no sign-extension instructions -- and the packed
//    table words of the first four tables of a batch live in the kernel-parameter constant bank, so they
//    are instruction operands rather than loads; 16-bit tables take the load + IMAD path;
//  * column pass, row pass, clamp and pack without transposes or shared-memory temporaries;
//  * 8-byte row stores: 32 lanes x 8 B = 256 contiguous bytes per plane row.
// ---------------------------------------------------------------------------------------------
constexpr int K1_STAGES = 3;


// MODE = column-pass style + 4 * row-pass style of the output butterflies (K1_BFLY_ADD): 0 plain, 1 IMAD.
// Measured (sweep_r01*.jsonl): ALU-pipe instructions issue at half rate on sm_100, IMAD at full rate, and the
// kernel is bound by the ALU pipe in style 0 and by issue slots in style 1; 5 (= both IMAD) is the fastest.  All variants are bit-identical; they only move work
// between the ALU and FMA pipes (the kernel is integer-issue bound, not HBM bound).
// ARITH: 0 = scalar (src/idct.rs), 1 = SSSE3 int16 lanes (src/arch/ssse3.rs) for the 8x8 blocks.
// SCALED: components may have dct_scale 4 / 2 / 1 (Decoder::scale); those blocks take the reduced transforms, which
// exist in scalar arithmetic only (src/idct.rs:247-253).  A separate instantiation: the hot one stays lean.
template <int MODE, int ARITH = 0, bool SCALED = false>
__global__ void __launch_bounds__(K1_TILE, 4)
k1_idct8_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ K1QCache qc, K1Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) unsigned long long full_bar[K1_STAGES];
    __shared__ __align__(8) unsigned long long empty_bar[K1_STAGES];

    const unsigned tid = threadIdx.x, lane = tid & 31;
    const unsigned t_begin = (unsigned)(((unsigned long long)blockIdx.x * p.ntiles) / gridDim.x);
    const unsigned t_end = (unsigned)(((unsigned long long)(blockIdx.x + 1) * p.ntiles) / gridDim.x);
    const unsigned n = t_end - t_begin;

    unsigned row_ahead = 0;  // thread 0: slab row of the tile the next refill will fetch
    if (tid == 0) {
        for (int st = 0; st < K1_STAGES; st++) {
            mbar_init(smem_u32(&full_bar[st]), 1);
            mbar_init(smem_u32(&empty_bar[st]), K1_TILE / 32);
        }
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
        for (unsigned it = 0; it < (unsigned)K1_STAGES && it < n; it++) {
            const unsigned bar = smem_u32(&full_bar[it]);
            mbar_expect_tx(bar, K1_STAGE_BYTES);
            tma_load_2d(smem + it * K1_STAGE_BYTES, &tmap, 0, (int)__ldg(&p.tiles[t_begin + it].slab_row), bar);
        }
        if (n > (unsigned)K1_STAGES) row_ahead = __ldg(&p.tiles[t_begin + K1_STAGES].slab_row);
    }
    __syncthreads();

    const unsigned row_off = tid * 128u, swz = (tid & 7u) << 4;
    unsigned cur_comp = 0xffffffffu;
    DevComp comp;
    comp.plane_off = 0; comp.stride = 0; comp.block_w = 1; comp.qt_index = 0; comp.dct_scale = 8; comp.nblocks = 0; comp.qflags = 0;
    const uint4* q4 = nullptr;
    const uint4* qp4 = nullptr;

    for (unsigned it = 0; it < n; it++) {
        const unsigned stage = it % K1_STAGES, round = it / K1_STAGES;
        // ---- refill: the slot drained during the previous iteration receives tile it-1+STAGES ----
        if (tid == 0 && it > 0 && it - 1 + K1_STAGES < n) {
            const unsigned ps = (it - 1) % K1_STAGES, pr = (it - 1) / K1_STAGES;
            mbar_wait(smem_u32(&empty_bar[ps]), pr & 1);
            const unsigned bar = smem_u32(&full_bar[ps]);
            mbar_expect_tx(bar, K1_STAGE_BYTES);
            tma_load_2d(smem + ps * K1_STAGE_BYTES, &tmap, 0, (int)row_ahead, bar);
            if (it + K1_STAGES < n) row_ahead = __ldg(&p.tiles[t_begin + it + K1_STAGES].slab_row);
        }
        const DevTile tile = p.tiles[t_begin + it];
        if (tile.comp != cur_comp) {  // warp-uniform
            cur_comp = tile.comp;
            comp = p.comps[cur_comp];
            q4 = reinterpret_cast<const uint4*>(p.qtabs + (size_t)comp.qt_index * 64);
            qp4 = reinterpret_cast<const uint4*>(p.qpack + (size_t)comp.qt_index * 32);
        }
        mbar_wait(smem_u32(&full_bar[stage]), round & 1);
        const unsigned sbase = smem + stage * K1_STAGE_BYTES + row_off;
        uint4 raw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) raw[k] = lds128(sbase + ((k * 16u) ^ swz));
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));

        if (tid >= tile.nvalid) continue;
        unsigned bx = (tile.bxy & 0xffffu) + tid, by = tile.bxy >> 16;
        if (comp.block_w >= (unsigned)K1_TILE) {
            if (bx >= comp.block_w) { bx -= comp.block_w; by += 1; }
        } else {
            by += bx / comp.block_w;
            bx %= comp.block_w;
        }
        const unsigned sc = SCALED ? comp.dct_scale : 8u;
        uint8_t* dst = p.planes + comp.plane_off + (size_t)by * sc * comp.stride + (size_t)bx * sc;
        if (SCALED && sc != 8u) {  // warp-uniform: a tile belongs to one component
            const unsigned* q = reinterpret_cast<const unsigned*>(q4);
            if (sc == 4u) idct4x4_regs(raw, q, dst, comp.stride);
            else if (sc == 2u) idct2x2_regs(raw, q, dst, comp.stride);
            else idct1x1_regs(raw, q, dst);
            continue;
        }

        if constexpr (ARITH == 1) {
            uint2 rows[8];
            idct8x8_ssse3_regs(raw, q4, rows);
#pragma unroll
            for (int r = 0; r < 8; r++) *reinterpret_cast<uint2*>(dst + (size_t)r * comp.stride) = rows[r];
            continue;
        }
        unsigned s[8][8];
        const unsigned oor = dequant_block(raw, s, comp.qflags, qc, q4, qp4);
        if (oor != 0) {
            idct8x8_scalar_exact(p.coefs + ((size_t)tile.slab_row + tid) * 64, reinterpret_cast<const unsigned*>(q4), dst, comp.stride);
            continue;
        }
        if constexpr ((MODE & 8) != 0) {
            // direct form (idct_core.cuh): fewest instructions, IMAD and ALU work in balance
            uint2 rows[8];
            idct8x8_direct(s, rows);
#pragma unroll
            for (int r = 0; r < 8; r++) *reinterpret_cast<uint2*>(dst + (size_t)r * comp.stride) = rows[r];
        } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned x0, x1, x2, x3, t0, t1, t2, t3;
            constexpr int CS = MODE & 3;  // output butterfly style of the column pass
            IDCT_1D(s[0][i], s[1][i], s[2][i], s[3][i], s[4][i], s[5][i], s[6][i], s[7][i], (512u + 0x80000000u),
                    x0, x1, x2, x3, t0, t1, t2, t3);
            s[0][i] = (unsigned)sar(K1_BFLY_ADD(CS, x0, t3), 10);
            s[7][i] = (unsigned)sar(K1_BFLY_SUB(CS, x0, t3), 10);
            s[1][i] = (unsigned)sar(K1_BFLY_ADD(CS, x1, t2), 10);
            s[6][i] = (unsigned)sar(K1_BFLY_SUB(CS, x1, t2), 10);
            s[2][i] = (unsigned)sar(K1_BFLY_ADD(CS, x2, t1), 10);
            s[5][i] = (unsigned)sar(K1_BFLY_SUB(CS, x2, t1), 10);
            s[3][i] = (unsigned)sar(K1_BFLY_ADD(CS, x3, t0), 10);
            s[4][i] = (unsigned)sar(K1_BFLY_SUB(CS, x3, t0), 10);
        }
        const unsigned XS = 65536u + (128u << 17);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            unsigned x0, x1, x2, x3, t0, t1, t2, t3;
            constexpr int RS = (MODE >> 2) & 3;  // output butterfly style of the row pass
            uint2 o;
            IDCT_1D(s[r][0], s[r][1], s[r][2], s[r][3], s[r][4], s[r][5], s[r][6], s[r][7], XS, x0, x1, x2, x3, t0, t1,
                    t2, t3);
            o.x = pack4_sat_u8(sar(K1_BFLY_ADD(RS, x0, t3), 17), sar(K1_BFLY_ADD(RS, x1, t2), 17), sar(K1_BFLY_ADD(RS, x2, t1), 17),
                               sar(K1_BFLY_ADD(RS, x3, t0), 17));
            o.y = pack4_sat_u8(sar(K1_BFLY_SUB(RS, x3, t0), 17), sar(K1_BFLY_SUB(RS, x2, t1), 17), sar(K1_BFLY_SUB(RS, x1, t2), 17),
                               sar(K1_BFLY_SUB(RS, x0, t3), 17));
            *reinterpret_cast<uint2*>(dst + (size_t)r * comp.stride) = o;
        }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_k1_generic(const K1Params& p, int arith, cudaStream_t stream) {
    if (p.ntiles == 0) return cudaSuccess;
    k1_idct_generic<<<p.ntiles, K1_TILE, 0, stream>>>(p, arith);
    return cudaGetLastError();
}

// experiment knob (profiling only): B200JPG_K1_MODE=0..7 selects the pipe-balance variant
int g_k1_mode = -1;
static int k1_mode() {
    if (g_k1_mode < 0) {
        const char* e = getenv("B200JPG_K1_MODE");
        g_k1_mode = e ? atoi(e) & 15 : K1_DEFAULT_MODE;
    }
    return g_k1_mode;
}

size_t k1_tma_smem_bytes() { return (size_t)K1_STAGES * K1_STAGE_BYTES + 1024; }

cudaError_t launch_k1_tma(const CUtensorMap& tmap, const K1QCache& qc, const K1Params& p, int arith, bool scaled, int num_sms, cudaStream_t stream) {
    if (p.ntiles == 0) return cudaSuccess;
    if (arith == 1 || scaled) {  // the less travelled variants: SSSE3 arithmetic and / or scaled components
        static bool attr_set3 = false;
        if (!attr_set3) {
            const void* fns[3] = {(const void*)k1_idct8_tma<8, 1, false>, (const void*)k1_idct8_tma<8, 1, true>, (const void*)k1_idct8_tma<8, 0, true>};
            for (const void* f : fns) {
                cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_tma_smem_bytes());
                if (e != cudaSuccess) return e;
            }
            attr_set3 = true;
        }
        unsigned grid3 = (unsigned)num_sms * 4u;
        if (grid3 > p.ntiles) grid3 = p.ntiles;
        if (arith == 1 && scaled) k1_idct8_tma<8, 1, true><<<grid3, K1_TILE, k1_tma_smem_bytes(), stream>>>(tmap, qc, p);
        else if (arith == 1) k1_idct8_tma<8, 1, false><<<grid3, K1_TILE, k1_tma_smem_bytes(), stream>>>(tmap, qc, p);
        else k1_idct8_tma<8, 0, true><<<grid3, K1_TILE, k1_tma_smem_bytes(), stream>>>(tmap, qc, p);
        return cudaGetLastError();
    }
    static bool attr_set = false;
    if (!attr_set) {
        const void* fns[3] = {(const void*)k1_idct8_tma<0>, (const void*)k1_idct8_tma<5>, (const void*)k1_idct8_tma<8>};
        for (const void* f : fns) {
            cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_tma_smem_bytes());
            if (e != cudaSuccess) return e;
        }
        attr_set = true;
    }
    unsigned grid = (unsigned)num_sms * 4u;
    if (grid > p.ntiles) grid = p.ntiles;
    switch (k1_mode() & 15) {
#define K1_CASE(M) case M: k1_idct8_tma<M><<<grid, K1_TILE, k1_tma_smem_bytes(), stream>>>(tmap, qc, p); break;
        K1_CASE(0)
        K1_CASE(5)
        default: k1_idct8_tma<8><<<grid, K1_TILE, k1_tma_smem_bytes(), stream>>>(tmap, qc, p); break;
    }
    return cudaGetLastError();
}

}  // namespace b200jpg
