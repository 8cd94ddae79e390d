// probe_zc.cu -- experiment only (scripts/pcie_probe4.py): upload by a kernel that reads page-locked host memory through its
// device-side mapping (zero copy) instead of a copy-engine transfer.  Built beside the scripts: nvcc -shared -o scripts/libprobe_zc.so.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void zc_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" int probe_zc_copy(const void* host_pinned, void* dev, size_t bytes, int blocks, void* stream) {
    zc_copy<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)host_pinned, (uint4*)dev, bytes / 16);
    return (int)cudaGetLastError();
}
