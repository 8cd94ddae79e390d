// kernels.h -- launch interface between the host planner (pipeline.cu) and the kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"
#include "sbs.h"

namespace b200jpg {

struct K1Params {
    const DevTile* tiles;
    const DevComp* comps;
    const unsigned* qtabs;  // u32[64] per table, natural order
    const unsigned* qpack;  // u32[32] per table: {q[2j], 0, 0, q[2j+1]} bytes, valid for 8-bit tables
    const short* coefs;     // coefficient slab base
    uint8_t* planes;        // plane slab base
    unsigned ntiles;
    unsigned one, minus_one;  // 1 and 0xffffffff (see fma_add in k1_idct.cu)
};

struct K2Params {
    const DevImage* images;
    const uint8_t* planes;
    uint8_t* out;
    unsigned nimages;
    unsigned flags;  // K2_FLAG_*
    int3 sixteen;  // (16, 16, 16), see ycbcr_scalar_y16 in k2_color.cu
};

// packed 8-bit tables of up to four components, passed by value = constant bank operands
struct K1QCache {
    unsigned b[4][32];
};

// K0: sparse block streams -> dense coefficient slab; max_blocks = largest K0Image::nb of the launch
cudaError_t launch_k0_expand(const K0Image* d_images, unsigned nimages, unsigned max_blocks, const uint8_t* d_streams, short* d_slab,
                             cudaStream_t stream);
// device entropy decoding (entropy_dev.h, ke_entropy.cu): cold + max_passes sync + prefix + write + dc on `stream`.
// d_work: ent_work_bytes() bytes; *d_status: per image [anomaly bits, completed] inside d_work, valid after the launch.
struct EntImage;
// d_streams holds the payloads and receives the compact streams (EntImage::cs_off) K0 then expands.
size_t ent_work_bytes(unsigned total_sub, unsigned nimages, unsigned max_comp_blocks, int max_passes);
cudaError_t launch_entropy(const EntImage* d_images, unsigned nimages, unsigned max_nsub, unsigned total_sub, unsigned max_comp_blocks,
                           uint8_t* d_streams, void* d_work, int max_passes, unsigned** d_status, cudaStream_t stream, uint64_t* launches);
// device-made streams (SBS_BLOCK_OFFSETS) of the same descriptor list: one thread per block
cudaError_t launch_k0_expand_blocks(const K0Image* d_images, unsigned nimages, unsigned max_blocks, const uint8_t* d_streams, short* d_slab,
                                    cudaStream_t stream);
cudaError_t launch_k0_zero_headers(const K0Image* d_images, unsigned nimages, unsigned max_blocks, uint8_t* d_streams, cudaStream_t stream);
// one entry per stream a group uploads by kernel (k0_expand.cu: k_gather_streams)
struct GatherItem {
    const void* src;             // page-locked host memory, read through its device mapping (unified addressing)
    unsigned long long dst_off;  // byte offset inside the device stream buffer
    unsigned n16, pad_;          // length in 16-byte words
};
cudaError_t launch_gather_streams(const GatherItem* d_items, unsigned nitems, uint8_t* d_streams, cudaStream_t stream);
cudaError_t launch_k1_generic(const K1Params& p, int arith, cudaStream_t stream);
size_t k1_tma_smem_bytes();
cudaError_t launch_k1_tma(const CUtensorMap& tmap, const K1QCache& qc, const K1Params& p, int arith, bool scaled, int num_sms, cudaStream_t stream);

// K2: max_w/max_h = largest output size in [first, first+count); the grid covers that and images
// smaller than it exit early.  `path` selects the kernel; images whose DevImage::path differs are
// skipped by it, so a heterogeneous batch is covered by one launch per path in use.
cudaError_t launch_k2_generic(const K2Params& p, unsigned first, unsigned count, unsigned max_w, unsigned max_h,
                              cudaStream_t stream);
cudaError_t launch_k2_420(const K2Params& p, unsigned first, unsigned count, unsigned max_w, unsigned max_h, bool ragged,
                          cudaStream_t stream);

cudaError_t launch_k2_rows16(unsigned path, const K2Params& p, unsigned first, unsigned count, unsigned max_w, unsigned max_h,
                             cudaStream_t stream);
cudaError_t launch_k3_format(const DevImage* images, unsigned first, unsigned count, unsigned max_w, unsigned max_h, const void* src, void* dst,
                             int format, const float scale[3], const float bias[3], cudaStream_t stream);
cudaError_t launch_k2_420_tma(const K2Params& p, const K2Strip* strips, unsigned nstrips, unsigned item_base, unsigned total_items,
                              int num_sms, cudaStream_t stream);
int k2_mode();
constexpr unsigned K2_FLAG_SSSE3 = 2u;           // K2Params::flags: SSSE3 colour arithmetic on DevImage::ssse3_pixels pixels per row
constexpr unsigned K2_FLAG_LDG_TAKES_420T = 1u;  // K2Params::flags: the load/store 4:2:0 kernel also takes the bulk-copy path's images
cudaError_t launch_k2_gray(const K2Params& p, unsigned first, unsigned count, unsigned max_w, unsigned max_h, cudaStream_t stream);

// KF: fused dequant + IDCT + upsample + colour (kf_fused.cu); tmap32 = the coefficient slab as [rows][64 x u16] with a
// 32-row box and 128B swizzle.  `mode` = KF_MODE_*; cols[0..ncols) are the columns of that mode inside the launch range.
struct KFParams {
    const FColumn* cols;
    const DevComp* comps;
    const unsigned* qtabs;
    const unsigned* qpack;
    const short* coefs;
    const DevImage* images;
    uint8_t* out;
    unsigned ncols, item_base, total_items;
    unsigned ystride, cstride;  // bytes per staged luma / chroma row (batch-wide maxima, multiples of 16)
    int sixteen;
};
size_t kf_smem_bytes(unsigned mode, unsigned ystride, unsigned cstride, unsigned warps);
unsigned kf_warps();     // warps per CTA of the fused kernel (4 or 8)
unsigned kf_strip_px();  // widest column strip the planner may cut (pixels, multiple of 16)
cudaError_t launch_kf(unsigned mode, const CUtensorMap& tmap32, const K1QCache& qc, const KFParams& p, int num_sms, cudaStream_t stream);

extern int g_k1_mode, g_k2_mode;  // profiling knobs, see b200jpg_debug_set_kernel_modes

}  // namespace b200jpg
