// Launch interface of kernels.cu (host side).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "device_types.h"

namespace fgpu {

constexpr int kScanThreads = 256;
constexpr int kRowsPerThread = 8;
constexpr int kTileRows = kScanThreads * kRowsPerThread;  // 2048: fixed when run directories are built

size_t scan_smem_bytes(int key_words, int n_numbufs);
cudaError_t launch_table_init(const QueryDesc& q, cudaStream_t st);
cudaError_t launch_scan(const QueryDesc* d_q, const QueryDesc& q, int sm_count, cudaStream_t st);
cudaError_t launch_finalize(const FinalizeDesc& f, cudaStream_t st);
cudaError_t launch_merge(const QueryDesc& q, const void* partial, cudaStream_t st);
cudaError_t launch_decode(const ChunkDesc& c, int32_t* out_i32, long long* out_i64, uint8_t* out_valid, int sm_count,
                          cudaStream_t st);

}  // namespace fgpu
