// device_types.h -- plain structs shared by the host planner and the CUDA kernels.
//
// HBM layout (DESIGN.md "Data layout"):
//   coefficient slab : for every (image, component) a dense run of block_w*block_h blocks, 64 int16
//                      each (128 B), natural coefficient order, blocks in raster order -- exactly
//                      what the reference pushes through Worker::append_row (src/decoder.rs:962-983).
//                      Component runs are 1024 B aligned.
//   plane slab       : one u8 plane per (image, component), stride = block_w*dct_scale, rows =
//                      block_h*dct_scale (src/worker/immediate.rs:30-37), 256 B aligned.
//   pixel slab       : per image width*height*ncomp interleaved bytes (src/worker/mod.rs:107-110),
//                      256 B aligned.
#pragma once
#include <stdint.h>

namespace b200jpg {

constexpr int K1_TILE = 128;  // blocks per K1 tile (one thread per block)

// One per (image, component).  32 B, read with two 16-byte loads.
struct __align__(16) DevComp {
    unsigned long long plane_off;  // byte offset of the plane inside the plane slab
    unsigned stride;               // block_w * dct_scale
    unsigned block_w;              // blocks per block row
    unsigned qt_index;             // index into the u32[64] quantisation tables
    unsigned dct_scale;            // 1, 2, 4, 8
    unsigned nblocks;              // block_w * block_h
    unsigned qflags;               // bit0: all 64 entries <= 255 (IDP.2A path); bits 8..15: constant-bank slot, 0xff = none
};

// One per K1 tile: up to K1_TILE consecutive blocks (raster order) of one component.  16 B.
struct __align__(16) DevTile {
    unsigned comp;      // index into DevComp[]
    unsigned slab_row;  // index of the tile's first block inside the coefficient slab (128 B units)
    unsigned bxy;       // bx0 | (by0 << 16): block coordinates of the first block
    unsigned nvalid;    // blocks of this tile that belong to the component (1..K1_TILE)
};

enum : unsigned { UP_H1V1 = 0, UP_H2V1 = 1, UP_H1V2 = 2, UP_H2V2 = 3, UP_GENERIC = 4 };
enum : unsigned { CC_NOCONVERT = 0, CC_RGB = 1, CC_YCBCR = 2, CC_CMYK = 3, CC_YCCK = 4, CC_GRAY = 5 };
enum : unsigned { K2_PATH_GENERIC = 0, K2_PATH_420 = 1, K2_PATH_444 = 2, K2_PATH_GRAY = 3, K2_PATH_420R = 4, K2_PATH_420T = 5, K2_PATH_422 = 6,
                  K2_PATH_BYTES = 7, K2_PATH_440 = 8, K2_NPATHS = 9 };

struct DevUpComp {
    unsigned long long plane_off;
    unsigned stride;  // row_stride, src/upsampler.rs:35
    unsigned in_w;    // component.size.width
    unsigned in_h;    // component.size.height
    unsigned kind;    // UP_*
    unsigned hs, vs;  // generic scaling factors
};

// One per image.  Read through the constant/L1 path by every K2 thread.
struct __align__(16) DevImage {
    unsigned long long out_off;  // byte offset of the image inside the pixel slab
    unsigned width, height;      // output size
    unsigned ncomp;
    unsigned cc;                 // CC_*
    unsigned ssse3_pixels;       // pixels per row converted with the SSSE3 formula (0 in scalar mode)
    unsigned path;               // K2_PATH_* the planner chose for this image
    DevUpComp c[4];
};

// One per (image on the bulk-copy 4:2:0 path, 2048-pixel strip): the flattened work list of k2_ycbcr420_tma.
struct __align__(16) K2Strip {
    unsigned image;       // index into DevImage[]
    unsigned x0;          // first pixel of the strip
    unsigned first_item;  // index of the strip's first row pair in the flattened (strip, row pair) sequence
    unsigned npairs;      // height / 2 + 1
};

// ---- fused kernel KF (kf_fused.cu): dense coefficients -> pixels, planes staged in shared memory only ----------
// Work list: one FColumn per (image, column strip of <= 1920 pixels); an item = one MCU row of a column.  Per MCU row a
// column's blocks form up to four contiguous runs of slab rows (luma block rows, Cb, Cr), each fetched as 32-block boxes.
enum : unsigned { KF_MODE_444 = 0, KF_MODE_420 = 1, KF_NMODES = 2 };

struct __align__(16) FRun {
    unsigned slab_row0;  // slab row (128 B units) of the run's first block in MCU row 0 of the image
    unsigned step;       // slab rows from one MCU row to the next (v * block_w of the component)
    unsigned len;        // blocks of the run per MCU row (halo blocks included)
    unsigned wrap;       // blocks per block row inside the run (== len unless two luma block rows were merged)
    unsigned comp;       // component 0..2
    unsigned dst_x;      // byte offset inside a staged plane row where block column 0 of the run lands
    unsigned dst_row;    // staged row of the run's first block row (0 or 8)
    unsigned box0;       // index of the run's first 32-block box within the item
};

struct __align__(16) FColumn {
    unsigned image;       // index into DevImage[]
    unsigned comp0;       // index of the image's first DevComp (three consecutive entries)
    unsigned first_item;  // index of the column's first MCU row in the flattened item sequence of its mode
    unsigned nrows;       // MCU rows
    unsigned x0;          // first output pixel of the strip (multiple of the MCU width)
    unsigned wpx;         // output pixels of the strip
    unsigned cx_base;     // 4:2:0: global index of the chroma sample staged at byte 16 of a chroma row
    unsigned nruns;
    unsigned nboxes;      // boxes per item
    unsigned ngroups;     // 16-pixel groups per output row of the strip
    unsigned gmagic;      // ceil(2^32 / ngroups): task / ngroups == __umulhi(task, gmagic) for every task index used
    unsigned pad;
    FRun run[4];
};

}  // namespace b200jpg
