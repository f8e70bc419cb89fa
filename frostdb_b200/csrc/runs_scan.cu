// k_runs: scan + range filter + group-by Sum/Count for row groups of SORTED parts (sm_100a).
//
// Takes over, for the dominant plan shape, what TableScan -> PredicateFilter -> HashAggregate do per
// record in the reference (physicalplan/filter.go:276-323, physicalplan/aggregate.go:386-560) when
//   * the filter is a conjunction of <= 2 int64 range leaves on PLAIN non-null columns (or absent),
//   * every group key is a dictionary-string column that is run-length only in the row group (what
//     compaction leaves behind: rows sorted by the key columns, table.go:1296-1346),
//   * every stored aggregate is Sum(int64) over a PLAIN non-null column (Count needs no input).
//
// Work decomposition: a row group is cut into spans of span_blocks * block_rows consecutive rows; warp w
// of the grid takes spans w, w + W, ...  Inside a span the warp streams blocks of block_rows rows of the
// staged columns through a private cp.async ring (16-byte LDGSTS, coalesced 512 bytes per instruction)
// and keeps, for every key column, a WARP-UNIFORM cursor into the run directory.  The distance from the
// current step to the nearest run end tells how many whole 32-row steps belong to one group; those steps
// run in a loop of two shared-memory loads, a range test and two adds per lane, with no vote, shuffle or
// atomic.  A step that straddles one run end is split by lane index.  Partial sums live in registers per
// lane and are reduced + added to the dense aggregate table once per group change.
//
// Algorithmic bytes per row: 8 per distinct staged column (the run directories are O(groups)).
#include <cuda_runtime.h>

#include <unordered_map>

#include "device_types.h"
#include "kernels.h"
#include "agg_ops.cuh"

namespace fgpu {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t kNoSlot = 0xffffffffu;
constexpr int kWarps = kRunsThreads / 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ unsigned long long lds64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}

// GEN = false: every stored aggregate is Sum(int64) (plain 64-bit adds, the headline shape).
// GEN = true: Sum / Min / Max over int64 or float64 through the shared reducers (agg_ops.cuh).
template <int NL, int NK, int NA, bool GEN>
__global__ void __launch_bounds__(kRunsThreads, 5) k_runs(const __grid_constant__ RunsDesc d) {
  extern __shared__ __align__(128) uint8_t dyn[];
  constexpr int NC = (NL + NA) > 0 ? (NL + NA) : 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * kWarps + warp, GW = gridDim.x * kWarps;
  const uint32_t BR = d.block_rows, D = d.n_ring, n_cols = d.n_cols;
  const uint32_t col_bytes = BR * 8u, slot_bytes = n_cols * col_bytes;
  const uint32_t span_rows = d.span_blocks * BR;
  const uint32_t ring_s = smem_u32(dyn) + uint32_t(warp) * D * slot_bytes;

  // ---- prefetch iterator: walks the warp's spans block by block, D - 1 blocks ahead of the consumer ----
  uint32_t i_span = gw, i_rg = 0, i_row = 0, i_end = 0;
  const uint8_t* i_src[NC];  // this lane's next 16-byte piece of every staged column (null: not staged here)
  auto issue_open = [&]() {  // i_span < n_spans: position on the first block of the span
    while (i_span >= __ldg(d.rg_first_span + i_rg + 1)) i_rg++;
    const RunsRg* R = d.rgs + i_rg;
    const uint32_t n_rows = __ldg(&R->n_rows);
    i_row = (i_span - __ldg(d.rg_first_span + i_rg)) * span_rows;
    i_end = min(n_rows, i_row + span_rows);
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const uint8_t* col = (uint32_t(c) < n_cols) ? reinterpret_cast<const uint8_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&R->col[c]))) : nullptr;
      i_src[c] = col ? col + size_t(i_row) * 8u + uint32_t(lane) * 16u : nullptr;
    }
  };
  auto issue_block = [&](uint32_t rs) {
    if (i_span < d.n_spans) {
      const uint32_t n = i_end - i_row;  // rows left in the span
      const uint32_t dst = ring_s + rs * slot_bytes + uint32_t(lane) * 16u;
#pragma unroll
      for (int c = 0; c < NC; c++) {
        if (i_src[c] != nullptr) {
          const uint8_t* src = i_src[c];
          const uint32_t dc = dst + uint32_t(c) * col_bytes;
          if (n >= BR) {
            if (BR == 256u) {
              cp_async16(dc, src);
              cp_async16(dc + 512u, src + 512);
              cp_async16(dc + 1024u, src + 1024);
              cp_async16(dc + 1536u, src + 1536);
            } else {
              const uint32_t iters = BR >> 6;
#pragma unroll 2
              for (uint32_t j = 0; j < iters; j++) cp_async16(dc + j * 512u, src + j * 512u);
            }
          } else {
            for (uint32_t o = uint32_t(lane) * 16u; o < n * 8u; o += 512u) cp_async16(dc + o - uint32_t(lane) * 16u, src + o - uint32_t(lane) * 16u);
          }
          i_src[c] = src + size_t(BR) * 8u;
        }
      }
      i_row += BR;
      if (i_row >= i_end) {
        i_span += GW;
        if (i_span < d.n_spans) issue_open();
      }
    }
    cp_async_commit();  // exactly one group per call, empty ones included
  };
  if (i_span < d.n_spans) issue_open();
  for (uint32_t k = 0; k + 1 < D; k++) issue_block(k);
  uint32_t rs = 0, rs_ahead = D - 1;

  // ---- running group (warp-uniform slot, per-lane partials) ----
  uint32_t cs = kNoSlot, cnt = 0, sel = 0;
  unsigned long long part[NA > 0 ? NA : 1];
  uint32_t afunc[NA > 0 ? NA : 1];
  bool aflt[NA > 0 ? NA : 1];
#pragma unroll
  for (int a = 0; a < NA; a++) {
    afunc[a] = GEN ? (d.agg_func[a] & 0xffu) : 1u;
    aflt[a] = GEN && (d.agg_func[a] >> 8) != 0;
    part[a] = GEN ? (unsigned long long)agg_identity(uint8_t(afunc[a]), aflt[a]) : 0ull;
  }
  // folds value bits v into the partial of aggregate q when act
  auto fold = [&](int q, bool act, unsigned long long v) {
    if constexpr (GEN) {
      if (act) part[q] = (unsigned long long)agg_combine(uint8_t(afunc[q]), aflt[q], (long long)part[q], (long long)v);
    } else {
      part[q] += act ? v : 0ull;
    }
  };
  auto flush = [&]() {
    const uint32_t tt = __reduce_add_sync(FULL, cnt);
    if (tt == 0) return;  // a group that received no row leaves the table untouched
    if (lane == 0) atomicAdd(d.t_rows + cs, (unsigned long long)tt);
    sel += tt;
    cnt = 0;
#pragma unroll
    for (int a = 0; a < NA; a++) {
      unsigned long long v = part[a];
      if constexpr (GEN) {
#pragma unroll
        for (int o = 16; o; o >>= 1)
          v = (unsigned long long)agg_combine(uint8_t(afunc[a]), aflt[a], (long long)v, (long long)__shfl_xor_sync(FULL, v, o));
        if (lane == 0) apply_agg(uint8_t(afunc[a]), aflt[a], d.t_agg[a] + cs, (long long)v);
        part[a] = (unsigned long long)agg_identity(uint8_t(afunc[a]), aflt[a]);
      } else {
        // 64-bit warp sum from three 32-bit REDUX sums over 22-bit pieces (32 x 2^22 fits): no shuffle chain
        const uint32_t s0 = __reduce_add_sync(FULL, uint32_t(v) & 0x3fffffu);
        const uint32_t s1 = __reduce_add_sync(FULL, uint32_t(v >> 22) & 0x3fffffu);
        const uint32_t s2 = __reduce_add_sync(FULL, uint32_t(v >> 44));
        v = (unsigned long long)s0 + ((unsigned long long)s1 << 22) + ((unsigned long long)s2 << 44);
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(d.t_agg[a] + cs), v);
        part[a] = 0;
      }
    }
  };

  uint32_t p_rg = 0;
  for (uint32_t p_span = gw; p_span < d.n_spans; p_span += GW) {
    while (p_span >= __ldg(d.rg_first_span + p_rg + 1)) p_rg++;
    const RunsRg* R = d.rgs + p_rg;
    const uint32_t row0 = (p_span - __ldg(d.rg_first_span + p_rg)) * span_rows;
    const uint32_t span_end = min(__ldg(&R->n_rows), row0 + span_rows);
    long long lo[NL > 0 ? NL : 1], hi[NL > 0 ? NL : 1];
#pragma unroll
    for (int l = 0; l < NL; l++) {
      lo[l] = __ldg(&R->lo[l]);
      hi[l] = __ldg(&R->hi[l]);
    }
    const bool all_pass = NL == 0 || __ldg(&R->all_pass) != 0;  // warp-uniform: statistics decided the filter
    // key cursors: the run that holds the span's first row (seed of its 128-row chunk)
    const Run* runs[NK > 0 ? NK : 1];
    uint32_t kk[NK > 0 ? NK : 1], kend[NK > 0 ? NK : 1], kadd[NK > 0 ? NK : 1];
#pragma unroll
    for (int k = 0; k < NK; k++) {
      runs[k] = reinterpret_cast<const Run*>(__ldg(reinterpret_cast<const unsigned long long*>(&R->runs[k])));
      const Seed* sd = reinterpret_cast<const Seed*>(__ldg(reinterpret_cast<const unsigned long long*>(&R->seeds[k]))) + row0 / uint32_t(kIndexRows);
      const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(sd));        // k, start, end, off
      const uint32_t sv = __ldg(reinterpret_cast<const uint32_t*>(sd) + 4);  // val
      kk[k] = s4.x;
      kend[k] = s4.z;
      kadd[k] = (sv + 1u) * d.stride[k];
    }
    // dictionary leaves: one more cursor per leaf, its run says whether the rows of the run pass
    const uint32_t n_pred = d.n_pred;
    const Run* pruns[kRunsPreds];
    const uint8_t* plut[kRunsPreds];
    uint32_t pk[kRunsPreds], pend[kRunsPreds];
    bool ppass[kRunsPreds];
#pragma unroll
    for (int i = 0; i < kRunsPreds; i++) {
      pruns[i] = nullptr;
      plut[i] = nullptr;
      pk[i] = 0;
      pend[i] = 0xffffffffu;
      ppass[i] = true;
      if (uint32_t(i) < n_pred) {
        pruns[i] = reinterpret_cast<const Run*>(__ldg(reinterpret_cast<const unsigned long long*>(&R->pred_runs[i])));
        if (pruns[i] != nullptr) {
          plut[i] = reinterpret_cast<const uint8_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&R->pred_lut[i])));
          const Seed* sd = reinterpret_cast<const Seed*>(__ldg(reinterpret_cast<const unsigned long long*>(&R->pred_seeds[i]))) + row0 / uint32_t(kIndexRows);
          const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(sd));
          const uint32_t sv = __ldg(reinterpret_cast<const uint32_t*>(sd) + 4);
          pk[i] = s4.x;
          pend[i] = s4.z;
          ppass[i] = (sv == 0xffffffffu) ? (d.pred_null[i] != 0) : (__ldg(plut[i] + sv) != 0);
        }
      }
    }
    auto advance_preds = [&](uint32_t row) {
#pragma unroll
      for (int i = 0; i < kRunsPreds; i++) {
        if (row >= pend[i]) {  // (pend == 0xffffffff: no such leaf here)
          uint32_t k = pk[i], e;
          do {
            k++;
            e = __ldg(&pruns[i][k + 1].start);
          } while (row >= e);
          pk[i] = k;
          pend[i] = e;
          const uint32_t v = __ldg(&pruns[i][k].val);
          ppass[i] = (v == 0xffffffffu) ? (d.pred_null[i] != 0) : (__ldg(plut[i] + v) != 0);
        }
      }
    };
    // moves the cursors onto the runs that hold `row` (same row in every lane: the loads broadcast)
    auto advance = [&](uint32_t row) {
#pragma unroll
      for (int k = 0; k < NK; k++) {
        if (row >= kend[k]) {
          uint32_t i = kk[k], e;
          do {
            i++;
            e = __ldg(&runs[k][i + 1].start);
          } while (row >= e);
          kk[k] = i;
          kend[k] = e;
          kadd[k] = (__ldg(&runs[k][i].val) + 1u) * d.stride[k];
        }
      }
    };
    for (uint32_t r0 = row0; r0 < span_end; r0 += BR) {
      issue_block(rs_ahead);
      rs_ahead = (rs_ahead + 1 == D) ? 0 : rs_ahead + 1;
      // the block's copies are the oldest pending group of every lane
      if (D == 2) cp_async_wait<1>();
      else if (D == 3) cp_async_wait<2>();
      else if (D == 4) cp_async_wait<3>();
      else cp_async_wait<0>();
      __syncwarp();
      const uint32_t base = ring_s + rs * slot_bytes + uint32_t(lane) * 8u;
      uint32_t lcol[NL > 0 ? NL : 1], acol[NA > 0 ? NA : 1];
#pragma unroll
      for (int l = 0; l < NL; l++) lcol[l] = base + d.leaf_col[l] * col_bytes;
#pragma unroll
      for (int a = 0; a < NA; a++) acol[a] = base + d.agg_col[a] * col_bytes;
      auto passes = [&](int s) -> bool {
        bool act = true;
#pragma unroll
        for (int l = 0; l < NL; l++) {
          const long long x = (long long)lds64(lcol[l] + uint32_t(s) * 256u);
          act = act && x >= lo[l] && x <= hi[l];
        }
        return act;
      };
      // The block is consumed segment by segment: a segment is the stretch of rows up to the nearest run
      // end of any key column (or the block end), i.e. rows of ONE group.  Its whole 32-row steps run
      // unmasked; the partial steps at its two ends are masked by row index.  A step that holds several
      // run ends is simply visited once per segment.
      const uint32_t rend = min(span_end, r0 + BR);
      const uint32_t idx0 = uint32_t(lane);
      uint32_t row = r0;
      while (row < rend) {
        if (n_pred) {  // rows of a run that fails a dictionary leaf are skipped as a whole
          advance_preds(row);
          if (!(ppass[0] && ppass[1])) {
            row = min(rend, (!ppass[0] && !ppass[1]) ? max(pend[0], pend[1]) : (!ppass[0] ? pend[0] : pend[1]));
            continue;
          }
        }
        advance(row);
        uint32_t seg_end = min(rend, min(pend[0], pend[1])), us = 0;
#pragma unroll
        for (int k = 0; k < NK; k++) {
          seg_end = min(seg_end, kend[k]);
          us += kadd[k];
        }
        if (us != cs) {
          if (cs != kNoSlot) flush();
          cs = us;
        }
        const uint32_t a = row - r0, b = seg_end - r0;  // block-relative rows [a, b)
        const uint32_t sa = (a + 31u) >> 5, sb = b >> 5;  // whole steps [sa, sb)
        auto masked_step = [&](uint32_t s) {
          const bool in = (s * 32u + idx0 - a) < (b - a);
          const bool act = in && (all_pass || passes(int(s)));
          cnt += act ? 1u : 0u;
#pragma unroll
          for (int q = 0; q < NA; q++) fold(q, act, lds64(acol[q] + s * 256u));
        };
        if (sa > sb) {
          masked_step(a >> 5);  // the segment lies inside one step
        } else {
          if (a & 31u) masked_step(a >> 5);
          if (all_pass) {
            cnt += sb - sa;
#pragma unroll 4
            for (uint32_t s = sa; s < sb; s++) {
#pragma unroll
              for (int q = 0; q < NA; q++) fold(q, true, lds64(acol[q] + s * 256u));
            }
          } else {
#pragma unroll 2
            for (uint32_t s = sa; s < sb; s++) {
              const bool act = passes(int(s));
              cnt += act ? 1u : 0u;
#pragma unroll
              for (int q = 0; q < NA; q++) fold(q, act, lds64(acol[q] + s * 256u));
            }
          }
          if (b & 31u) masked_step(sb);
        }
        row = seg_end;
      }
      __syncwarp();  // every lane is done with ring slot rs before it is refilled
      rs = (rs + 1 == D) ? 0 : rs + 1;
    }
  }
  if (cs != kNoSlot) flush();
  if (lane == 0 && sel) atomicAdd(d.counters, (unsigned long long)sel);  // rows that passed the filter
}

template <int NL, int NK>
cudaError_t launch_na(const RunsDesc& d, int na, bool gen, dim3 grid, size_t smem, cudaStream_t st, bool query_only, int* per_sm) {
  auto go = [&](auto kern) -> cudaError_t {
    if (query_only) {  // per kernel instance: attribute and occupancy are looked up once per shared-memory size
      struct Cache { size_t cfg_smem = 0, occ_smem = ~size_t(0); int occ = 0; };
      static std::unordered_map<const void*, Cache> cache;  // (every instance has the same function type: key by address)
      Cache& c = cache[reinterpret_cast<const void*>(kern)];
      if (smem > c.cfg_smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        c.cfg_smem = smem;
      }
      if (c.occ_smem != smem) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.occ, kern, kRunsThreads, smem);
        if (e != cudaSuccess) return e;
        c.occ_smem = smem;
      }
      *per_sm = c.occ;
      return cudaSuccess;
    }
    kern<<<grid, kRunsThreads, smem, st>>>(d);
    return cudaGetLastError();
  };
  if (gen && na > 0) return na == 1 ? go(k_runs<NL, NK, 1, true>) : (na == 2 ? go(k_runs<NL, NK, 2, true>) : go(k_runs<NL, NK, 3, true>));
  switch (na) {
    case 0: return go(k_runs<NL, NK, 0, false>);
    case 1: return go(k_runs<NL, NK, 1, false>);
    case 2: return go(k_runs<NL, NK, 2, false>);
    default: return go(k_runs<NL, NK, 3, false>);
  }
}
template <int NL>
cudaError_t launch_nk(const RunsDesc& d, int nk, int na, bool gen, dim3 grid, size_t smem, cudaStream_t st, bool query_only, int* per_sm) {
  switch (nk) {
    case 0: return launch_na<NL, 0>(d, na, gen, grid, smem, st, query_only, per_sm);
    case 1: return launch_na<NL, 1>(d, na, gen, grid, smem, st, query_only, per_sm);
    case 2: return launch_na<NL, 2>(d, na, gen, grid, smem, st, query_only, per_sm);
    default: return launch_na<NL, 3>(d, na, gen, grid, smem, st, query_only, per_sm);
  }
}
cudaError_t launch_nl(const RunsDesc& d, int nl, int nk, int na, dim3 grid, size_t smem, cudaStream_t st, bool query_only, int* per_sm) {
  bool gen = false;  // any reducer other than Sum(int64)
  for (int a = 0; a < na; a++) gen = gen || d.agg_func[a] != 1u;
  switch (nl) {
    case 0: return launch_nk<0>(d, nk, na, gen, grid, smem, st, query_only, per_sm);
    case 1: return launch_nk<1>(d, nk, na, gen, grid, smem, st, query_only, per_sm);
    default: return launch_nk<2>(d, nk, na, gen, grid, smem, st, query_only, per_sm);
  }
}

}  // namespace

size_t runs_smem_bytes(const RunsDesc& d) { return size_t(kWarps) * d.n_ring * d.n_cols * d.block_rows * 8; }

cudaError_t runs_blocks_per_sm(const RunsDesc& d, int nl, int nk, int na, int* per_sm) {
  return launch_nl(d, nl, nk, na, dim3(1), runs_smem_bytes(d), nullptr, true, per_sm);
}

cudaError_t launch_runs(const RunsDesc& d, int nl, int nk, int na, int sm_count, cudaStream_t st) {
  if (d.n_spans == 0) return cudaSuccess;
  const size_t smem = runs_smem_bytes(d);
  int per_sm = 0;
  cudaError_t e = launch_nl(d, nl, nk, na, dim3(1), smem, st, true, &per_sm);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = uint32_t(sm_count) * uint32_t(per_sm);
  const uint32_t need = (d.n_spans + kWarps - 1) / kWarps;
  if (grid > need) grid = need;
  return launch_nl(d, nl, nk, na, dim3(grid), smem, st, false, nullptr);
}

}  // namespace fgpu
