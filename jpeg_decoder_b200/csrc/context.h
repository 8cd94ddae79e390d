// context.h -- the context object shared by pipeline.cu and files_api.cpp (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);

struct b200jpg_ctx {
    int device = 0;
    int arith = B200JPG_ARITH_SCALAR;
    int k1_kernel = B200JPG_KERNEL_AUTO;
    int k2_kernel = B200JPG_KERNEL_AUTO;
    int host_compact = B200JPG_COMPACT_AUTO;
    int host_threads = 0;
    int entropy = B200JPG_ENTROPY_AUTO;
    int fuse = B200JPG_FUSE_AUTO;
    cudaStream_t stream = nullptr;   // main stream (caller's or ours)
    cudaStream_t stream2 = nullptr;  // second stream of the host pipeline
    bool own_stream = false;
    int num_sms = 148;
    std::atomic<uint64_t> launches{0};  // bumped from the stream-engine threads as well as from API calls
    uint64_t device_scans = 0, device_scan_retries = 0;  // b200jpg_decode_files: scans Huffman-decoded on the GPU / sent back to the host
    PFN_tensorMapEncodeTiled encode = nullptr;
    std::string err;  // guarded by err_mu (b200jpg_fail is called from paths that do not hold mu)
    std::mutex err_mu;
    // grow-only caches so that repeated batches / file chunks do not pay cudaMalloc / cudaHostAlloc each time
    struct Buf {
        void* p = nullptr;
        size_t cap = 0;
    };
    std::mutex mu;
    Buf scratch[3];  // device: coefficient, plane and pixel slabs of the host pipeline
    bool scratch_busy = false;
    // the sparse-stream pipeline and the whole-file engine built on it (sbs_pipeline.h, files_api.cpp), created on
    // first use and destroyed with the context
    void* sbs_pipeline = nullptr;
    void (*sbs_pipeline_free)(void*) = nullptr;
    void* files_engine = nullptr;
    void (*files_engine_free)(void*) = nullptr;
};


// Where Huffman decoding of qualifying scans happens (B200JPG_ENTROPY_*); the environment variable exists for tests and
// benchmarks that want to compare both routes with one binary.
inline bool b200jpg_device_entropy_enabled(const b200jpg_ctx* ctx) {
    if (const char* e = getenv("B200JPG_ENTROPY")) {
        if (!strcmp(e, "host")) return false;
        if (!strcmp(e, "device")) return true;
    }
    return ctx && ctx->entropy != B200JPG_ENTROPY_HOST;
}
