// batch_internal.h -- the batch plan object and the planner/launcher entry points shared by pipeline.cu
// (dense host pipeline, worker API) and sbs_pipeline.cu (sparse-stream host pipeline).  Not part of the C ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/b200jpg.h"
#include "context.h"
#include "device_types.h"
#include "kernels.h"

using namespace b200jpg;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ImageLayout {
    size_t coef_off[4] = {0, 0, 0, 0};
    size_t plane_off[4] = {0, 0, 0, 0};
    size_t coef_bytes[4] = {0, 0, 0, 0};
    size_t out_off = 0, out_len = 0;
    unsigned tile_first = 0, tile_count = 0;
    int status = B200JPG_OK;
};

struct b200jpg_batch {
    b200jpg_ctx* ctx = nullptr;
    size_t n = 0;
    std::vector<ImageLayout> layout;
    std::vector<DevComp> comps;
    std::vector<DevTile> tiles;
    std::vector<DevImage> images;
    std::vector<unsigned> qtabs;  // 64 per table
    std::vector<unsigned> qpack;  // 32 per table (8-bit tables: {q[2j], 0, 0, q[2j+1]})
    std::vector<unsigned char> qt_is8;
    K1QCache qcache;
    b200jpg_batch_info info{};
    bool all_scale8 = true;
    bool k1_tma_aligned = true;
    unsigned path_max_w[K2_NPATHS] = {}, path_max_h[K2_NPATHS] = {};
    bool path_used[K2_NPATHS] = {};
    std::vector<K2Strip> strips;         // work list of the bulk-copy 4:2:0 kernel (images on K2_PATH_420T)
    std::vector<unsigned> strip_first;   // per image: index of its first strip (n + 1 entries)
    unsigned strip_items = 0;
    K2Strip* d_strips = nullptr;
    // fused kernel KF (kf_fused.cu): per mode the column work list of the images it can take, in image order
    std::vector<FColumn> fcols[KF_NMODES];
    std::vector<unsigned> fcol_first[KF_NMODES];  // per image: index of its first column (n + 1 entries)
    std::vector<signed char> fmode;               // per image: KF_MODE_* or -1 (needs K1 + K2)
    unsigned fitems[KF_NMODES] = {0, 0};
    unsigned f_ystride[KF_NMODES] = {0, 0}, f_cstride[KF_NMODES] = {0, 0};
    FColumn* d_fcols[KF_NMODES] = {nullptr, nullptr};
    CUtensorMap tmap32;  // same slab, 32-row box
    // device copies of the tables
    DevComp* d_comps = nullptr;
    DevTile* d_tiles = nullptr;
    DevImage* d_images = nullptr;
    unsigned* d_qtabs = nullptr;
    unsigned* d_qpack = nullptr;
    // tensor map cache (one slab pointer at a time)
    const void* tmap_base = nullptr;
    CUtensorMap tmap;
    // internal slabs for the host pipeline
    void* d_coefs = nullptr;
    void* d_planes = nullptr;
    void* d_out = nullptr;
    bool planes_absolute = false;  // plane_off holds absolute device addresses (worker path)
    bool outs_absolute = false;    // out_off holds absolute device addresses (whole-file engine, pixels staying on the device)
    bool slabs_borrowed = false;   // d_coefs/d_planes/d_out belong to the context's scratch cache
    bool tables_borrowed = false;  // the d_* tables live in a caller's TableArena
    size_t table_bytes = 0;        // bytes of that arena in use
    // b200jpg_batch_run_host, host_compact = AUTO: whether compaction pays depends on what bounds the host side
    // (PCIe: it does; host DRAM bandwidth: it does not -- the CPU then reads what the DMA engine would have read), so
    // a batch that is run repeatedly times its second dense run and its second compacted run and keeps the faster.
    int compact_decision = -1;     // -1 undecided, 0 dense upload, 1 host compaction
    int compact_trials = 0;
    double compact_ms[2] = {0, 0};
};

// Caller-provided home of a plan's device tables: `bytes` of device memory at `d` mirrored by page-locked host
// memory at `h`.  The planner fills `h`, enqueues ONE copy on the upload stream and never cudaMalloc/cudaFree's
// (cudaFree synchronises the whole device, which would serialise a multi-stream pipeline).
struct TableArena {
    char* d = nullptr;
    char* h = nullptr;
    size_t bytes = 0;
};

struct PlanOverrides {
    const unsigned long long (*plane_addr)[4] = nullptr;  // per image absolute device addresses of the planes
    const unsigned long long* out_addr = nullptr;         // per image absolute device address of the pixels (no pixel slab)
    const TableArena* arena = nullptr;                    // nullptr: cudaMalloc each table
    cudaStream_t upload_stream = nullptr;                 // nullptr: the context's main stream
};


int b200jpg_fail(b200jpg_ctx* ctx, int code, const std::string& msg);
int b200jpg_cuda_fail(b200jpg_ctx* ctx, cudaError_t e, const char* what);
#define CU_TRY(ctx, call)                                                  \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) return b200jpg_cuda_fail((ctx), e_, #call); \
    } while (0)

// Validates and lays out n images, uploads the tables (PlanOverrides says where to and on which stream).
int batch_create_impl(b200jpg_ctx* ctx, const b200jpg_image_desc* imgs, size_t n, int* statuses, const PlanOverrides& ov,
                      b200jpg_batch** out);
void batch_release_device(b200jpg_batch* b);
// K1 over tiles [tile_first, tile_first+tile_count), K2 over images [img_first, img_first+img_count)
int batch_launch(b200jpg_batch* b, const void* d_coefs, void* d_planes, void* d_out, int stages, unsigned tile_first,
                 unsigned tile_count, unsigned img_first, unsigned img_count, cudaStream_t stream);
