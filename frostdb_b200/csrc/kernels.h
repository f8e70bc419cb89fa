// Launch interface of kernels.cu (host side).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "device_types.h"

namespace fgpu {

constexpr int kScanThreads = 256;
constexpr int kRowsPerThread = 4;
constexpr int kTileRows = kScanThreads * kRowsPerThread;  // 1024 rows per CTA iteration
constexpr int kVecThreads = 128;  // scan kernel: 4 independent warps per CTA
constexpr int kMaxVecRows = 1024;  // rows per warp vector (multiple of kIndexRows)
constexpr int kIndexRows = 128;  // chunk index granularity (one warp): fixed when run directories are built

cudaError_t launch_table_init(const QueryDesc& q, cudaStream_t st, bool zero_counters = true);
cudaError_t launch_scan(const QueryDesc* d_q, const QueryDesc& q, int sm_count, cudaStream_t st);
// sorted-run scan (runs_scan.cu): nl range leaves, nk run-length keys, na Sum(int64) aggregates
constexpr int kRunsThreads = 128;
cudaError_t launch_runs(const RunsDesc& d, int nl, int nk, int na, int sm_count, cudaStream_t st);
cudaError_t runs_blocks_per_sm(const RunsDesc& d, int nl, int nk, int na, int* per_sm);  // resident CTAs per SM for this shape
// sorted-run scan, CTA-cooperative with TMA-staged column tiles (runs_tma.cu): same descriptors as launch_runs
uint32_t runs_tma_tile_rows(uint32_t base, uint32_t nc);  // rows per tile of a row group that stages nc columns
void runs_tma_plan(RunsDesc& d, int nq, uint32_t base_tile, uint32_t col_region, int force_stages, int force_warps, int force_chunk);  // tile rows, ring depth, stage layout
size_t runs_tma_smem_bytes(const RunsDesc& d);
int runs_tma_ctas_per_sm(const RunsDesc& d);
cudaError_t launch_runs_tma(const RunsDesc& d, int nl, int nk, int na, int sm_count, cudaStream_t st);
// filter-only plans over PLAIN columns (take_rows.cu): count per span, scan, ordered write
cudaError_t launch_take(const TakeDesc& d, int sm_count, cudaStream_t st);
int take_resident_warps(int sm_count);
// tile aggregate (tile_agg.cu): TMA-staged column tiles, CTA-private shared-memory table
cudaError_t launch_tile_agg(const TileAggDesc& d, int sm_count, cudaStream_t st);
size_t tile_agg_smem_bytes(const TileAggDesc& d);
// dictionary column chunks -> flat code arrays (kernels.cu)
cudaError_t launch_flatten(const FlatJob* d_jobs, uint32_t n_jobs, uint32_t total_blocks, int sm_count, cudaStream_t st);
// partial-table exchange over peer-mapped mailboxes (comm.cu)
cudaError_t launch_comm_push(const CommPush& p, int sm_count, cudaStream_t st);
cudaError_t launch_comm_wait(const CommWait& w, cudaStream_t st);
cudaError_t launch_merge_dense(const QueryDesc& q, const CommMerge& m, cudaStream_t st);
cudaError_t launch_rows(const QueryDesc* d_q, const QueryDesc& q, int sm_count, cudaStream_t st);
cudaError_t launch_finalize(const FinalizeDesc& f, cudaStream_t st);
cudaError_t launch_merge(const QueryDesc& q, const void* partial, cudaStream_t st);
cudaError_t launch_finalize_dense(const DenseOut& f, cudaStream_t st);
cudaError_t launch_gather_pages(const void* stage, void* image, const PageCopy* table, uint32_t n, cudaStream_t st);
cudaError_t launch_make_seeds(void* image, uint64_t jobs_off, uint32_t n_jobs, uint32_t max_chunks, cudaStream_t st);
cudaError_t launch_decode(const ChunkDesc& c, int32_t* out_i32, long long* out_i64, uint8_t* out_valid, int sm_count,
                          cudaStream_t st);

}  // namespace fgpu
