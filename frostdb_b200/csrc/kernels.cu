// sm_100a kernels of libfrostgpu.
//
//   k_scan      K1+K2+K4/K5 fused: per 2048-row tile, decode the projected column chunks from their
//               stored Parquet encoding (PLAIN / RLE-bit-packed hybrid / RLE_DICTIONARY) in shared
//               memory, evaluate the predicate tree into a per-row selection, and fold the selected
//               rows into the aggregate table with warp-aggregated atomics.  Replaces
//               ParquetConverter.Convert (pqarrow/arrow.go:264-373) + PredicateFilter.Callback
//               (query/physicalplan/filter.go:255-323) + HashAggregate.Callback
//               (query/physicalplan/aggregate.go:263-490) + Distinction.Callback (distinct.go:70-170).
//   k_finalize  K6: aggregate table -> compacted result columns (finishAggregate, aggregate.go:543-633).
//   k_merge     K6: folds gathered partial tables of other ranks into the local table
//               (Synchronizer + final HashAggregate, synchronize.go:16-53, physicalplan.go:438-471).
//   k_decode    K1 standalone: one column chunk -> dense Arrow-style buffers.
//
// This is HBM-bound integer/indexing work: no tensor cores.  Loads of PLAIN columns are coalesced
// 8-byte-per-lane streams, hybrid streams are read through 4-byte aligned windows, atomics are
// aggregated per warp with __match_any_sync before touching the table.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.h"
#include "kernels.h"

namespace fgpu {

namespace {

constexpr int NT = kScanThreads;      // threads per CTA
constexpr int RPT = kRowsPerThread;   // rows per thread
constexpr int TILE = NT * RPT;        // rows per tile
constexpr int NWARP = NT / 32;
static_assert(TILE == kTileRows, "tile size mismatch");
static_assert(RPT == 8, "blocked decode assumes 8 rows per thread");

struct SmemLayout {
  uint16_t* ridx;      // [TILE] run index per value slot (scratch of the hybrid expander)
  uint32_t* tmp;       // [TILE] decoded values of the column being processed
  uint32_t* leafbits;  // [TILE] one bit per predicate leaf
  uint8_t* vbyte;      // [NT] validity of 8 consecutive rows
  uint16_t* vrank;     // [NT] number of valid rows before each group of 8
  uint32_t* wscr;      // [64] cross-warp scratch
  unsigned long long* keyw;  // [key_words][TILE] packed group key (dense mode: word 0 = slot index)
  long long* numbuf;   // [n_numbufs][TILE] staged numeric columns
  uint32_t* numnull;   // [n_numbufs][TILE/32] null bits of staged numeric columns
};

__device__ __forceinline__ SmemLayout carve(uint8_t* base, int key_words, int n_numbufs) {
  SmemLayout s;
  size_t off = 0;
  s.keyw = reinterpret_cast<unsigned long long*>(base + off);
  off += size_t(key_words) * TILE * 8;
  s.numbuf = reinterpret_cast<long long*>(base + off);
  off += size_t(n_numbufs) * TILE * 8;
  s.tmp = reinterpret_cast<uint32_t*>(base + off);
  off += TILE * 4;
  s.leafbits = reinterpret_cast<uint32_t*>(base + off);
  off += TILE * 4;
  s.ridx = reinterpret_cast<uint16_t*>(base + off);
  off += TILE * 2;
  s.numnull = reinterpret_cast<uint32_t*>(base + off);
  off += size_t(n_numbufs) * (TILE / 32) * 4;
  s.wscr = reinterpret_cast<uint32_t*>(base + off);
  off += 64 * 4;
  s.vrank = reinterpret_cast<uint16_t*>(base + off);
  off += NT * 2;
  s.vbyte = base + off;
  return s;
}

// ---- block-wide helpers ---------------------------------------------------------------------------

// In-place inclusive max-scan over ridx[0..TILE).  Markers are increasing run indices, so the
// running maximum at a position is the run that covers it.
__device__ __forceinline__ void block_max_scan_u16(uint16_t* ridx, uint32_t* wscr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint4 v = reinterpret_cast<uint4*>(ridx)[tid];
  uint32_t a[8] = {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16,
                   v.z & 0xffffu, v.z >> 16, v.w & 0xffffu, v.w >> 16};
#pragma unroll
  for (int i = 1; i < 8; i++) a[i] = max(a[i], a[i - 1]);
  uint32_t incl = a[7];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl = max(incl, o);
  }
  uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0;
  if (lane == 31) wscr[warp] = incl;
  __syncthreads();
  uint32_t base = excl;
  for (int w = 0; w < warp; w++) base = max(base, wscr[w]);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = max(a[i], base);
  v.x = a[0] | (a[1] << 16);
  v.y = a[2] | (a[3] << 16);
  v.z = a[4] | (a[5] << 16);
  v.w = a[6] | (a[7] << 16);
  reinterpret_cast<uint4*>(ridx)[tid] = v;
  __syncthreads();
}

// Exclusive sum-scan of one value per thread; returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t x, uint32_t* wscr, uint32_t* total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t incl = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) wscr[32 + warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
  for (int w = 0; w < NWARP; w++) {
    uint32_t t = wscr[32 + w];
    if (w < warp) base += t;
    tot += t;
  }
  *total = tot;
  __syncthreads();
  return base + incl - x;
}

__device__ __forceinline__ uint32_t extract_bits(const uint8_t* stream, uint32_t off, uint64_t bit, uint32_t w) {
  uint64_t byte = uint64_t(off) + (bit >> 3);
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(stream) + (byte >> 2);
  uint32_t sh = uint32_t(byte & 3) * 8 + uint32_t(bit & 7);
  uint32_t lo = __ldg(wp), hi = __ldg(wp + 1);
  uint32_t v = __funnelshift_r(lo, hi, sh);
  uint32_t mask = (w >= 32) ? 0xffffffffu : ((1u << w) - 1u);
  return v & mask;
}

// Expands `nv` values of a hybrid stream starting at value ordinal v0 into out[0..nv).
// first_run is the directory index of the run that holds ordinal v0.  Blocked mapping: thread t
// produces slots 8t..8t+7.  Ends with a barrier: out[] is visible to the whole CTA on return.
__device__ __forceinline__ void decode_hybrid(const uint8_t* __restrict__ stream, const Run* __restrict__ runs,
                                              uint32_t first_run, uint32_t n_runs, uint32_t v0, uint32_t nv,
                                              uint32_t* out, uint16_t* ridx, uint32_t* wscr) {
  const int tid = threadIdx.x;
  reinterpret_cast<uint4*>(ridx)[tid] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const uint32_t vend = v0 + nv;
  for (uint32_t k = first_run + 1 + tid; k < n_runs; k += NT) {
    uint32_t st = __ldg(&runs[k].start);
    if (st >= vend) break;
    ridx[st - v0] = uint16_t(k - first_run);
  }
  __syncthreads();
  block_max_scan_u16(ridx, wscr);
  const uint32_t s0 = uint32_t(tid) * RPT;
  if (s0 < nv) {
    uint4 rv = reinterpret_cast<const uint4*>(ridx)[tid];
    uint32_t rr[8] = {rv.x & 0xffffu, rv.x >> 16, rv.y & 0xffffu, rv.y >> 16,
                      rv.z & 0xffffu, rv.z >> 16, rv.w & 0xffffu, rv.w >> 16};
    uint32_t cur = 0xffffffffu;
    uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < RPT; j++) {
      uint32_t s = s0 + j;
      if (s < nv) {
        if (rr[j] != cur) {
          cur = rr[j];
          r = __ldg(reinterpret_cast<const uint4*>(runs + first_run + cur));
        }
        uint32_t val;
        if ((r.w & 1u) == 0) {
          val = r.z;
        } else {
          uint32_t w = (r.w >> 8) & 0xffu;
          val = extract_bits(stream, r.y, uint64_t(v0 + s - r.x) * w, w);
        }
        out[s] = val;
      }
    }
  }
  __syncthreads();
}

// Definition levels of the tile -> vbyte/vrank; returns the number of valid rows.
__device__ __forceinline__ uint32_t decode_validity(const ChunkDesc& c, uint32_t tile_in_rg, uint32_t r0, uint32_t n,
                                                    const SmemLayout& sm) {
  decode_hybrid(c.def, c.def_runs, __ldg(&c.tile_defrun[tile_in_rg]), c.n_defruns, r0, n, sm.tmp, sm.ridx, sm.wscr);
  const int tid = threadIdx.x;
  uint32_t byte = 0;
#pragma unroll
  for (int j = 0; j < RPT; j++) {
    uint32_t s = uint32_t(tid) * RPT + j;
    if (s < n && sm.tmp[s] != 0) byte |= 1u << j;
  }
  uint32_t total;
  uint32_t excl = block_excl_sum(__popc(byte), sm.wscr, &total);
  sm.vbyte[tid] = uint8_t(byte);
  sm.vrank[tid] = uint16_t(excl);
  __syncthreads();
  return total;
}

__device__ __forceinline__ bool row_valid(const SmemLayout& sm, uint32_t i, uint32_t* pos) {
  uint32_t byte = sm.vbyte[i >> 3];
  uint32_t bit = i & 7u;
  *pos = uint32_t(sm.vrank[i >> 3]) + __popc(byte & ((1u << bit) - 1u));
  return (byte >> bit) & 1u;
}

// ---- predicate / expression evaluation ---------------------------------------------------------------

__device__ __forceinline__ bool cmp_i64(uint8_t op, long long a, long long b) {
  switch (op) {
    case 1: return a == b;
    case 2: return a != b;
    case 3: return a < b;
    case 4: return a <= b;
    case 5: return a > b;
    case 6: return a >= b;
    default: return false;
  }
}
__device__ __forceinline__ bool cmp_f64(uint8_t op, double a, double b) {
  switch (op) {
    case 1: return a == b;
    case 2: return a != b;
    case 3: return a < b;
    case 4: return a <= b;
    case 5: return a > b;
    case 6: return a >= b;
    default: return false;
  }
}

__device__ __forceinline__ bool eval_filter(const QueryDesc& q, uint32_t bits) {
  // postfix program over leaf bits; stack kept in a 32-bit word
  uint32_t stack = 0;
  int sp = 0;
  for (int i = 0; i < q.n_filter_prog; i++) {
    uint8_t op = q.filter_prog[i];
    if (op < 0x80) {
      stack = (stack & ~(1u << sp)) | (((bits >> op) & 1u) << sp);
      sp++;
    } else {
      uint32_t b = (stack >> (sp - 1)) & 1u, a = (stack >> (sp - 2)) & 1u;
      uint32_t r = (op == 0x80) ? (a & b) : (a | b);
      sp -= 2;
      stack = (stack & ~(1u << sp)) | (r << sp);
      sp++;
    }
  }
  return stack & 1u;
}

struct NumVal {
  long long bits;
  bool null;
};

// Value of numeric slot `slot` for tile row i (raw 0 for NULL, as builder.AppendValue leaves it:
// pqarrow/builder/utils.go:54-58 + optbuilders.go:337-340).
__device__ __forceinline__ NumVal load_num(const QueryDesc& q, const ChunkDesc* __restrict__ chunks, const SmemLayout& sm,
                                           int slot, uint32_t r0, uint32_t i) {
  NumVal v;
  int nb = q.slot_numbuf[slot];
  const ChunkDesc& c = chunks[slot];
  if (c.kind == CK_ABSENT) {
    v.bits = 0;
    v.null = true;
  } else if (nb < 0 || (c.kind == CK_PLAIN64 && !c.has_nulls)) {
    v.bits = __ldg(reinterpret_cast<const long long*>(c.values) + r0 + i);
    v.null = false;
  } else {
    v.bits = sm.numbuf[size_t(nb) * TILE + i];
    v.null = (sm.numnull[nb * (TILE / 32) + (i >> 5)] >> (i & 31)) & 1u;
  }
  return v;
}

__device__ __forceinline__ long long eval_prog(const QueryDesc& q, const AggDesc& a, const ChunkDesc* __restrict__ chunks,
                                               const SmemLayout& sm, uint32_t r0, uint32_t i) {
  // Arithmetic ignores validity and Div by zero yields NULL, i.e. a raw 0 in the aggregated
  // array (query/physicalplan/project.go:169-395, :216-218).
  long long st[8];
  int sp = 0;
  for (int p = a.prog_off; p < a.prog_off + a.prog_len; p++) {
    const ProgOp& o = q.prog[p];
    if (o.op == PO_LOAD) {
      st[sp++] = load_num(q, chunks, sm, o.slot, r0, i).bits;
    } else if (o.op == PO_CONST) {
      st[sp++] = o.imm;
    } else {
      long long rb = st[--sp], lb = st[--sp], res;
      if (a.is_float) {
        double l = __longlong_as_double(lb), r = __longlong_as_double(rb), x;
        switch (o.op) {
          case PO_ADD: x = l + r; break;
          case PO_SUB: x = l - r; break;
          case PO_MUL: x = l * r; break;
          default: x = (r == 0.0) ? 0.0 : l / r; break;
        }
        res = __double_as_longlong(x);
      } else {
        unsigned long long l = (unsigned long long)lb, r = (unsigned long long)rb;
        switch (o.op) {
          case PO_ADD: res = (long long)(l + r); break;
          case PO_SUB: res = (long long)(l - r); break;
          case PO_MUL: res = (long long)(l * r); break;
          default:
            if (rb == 0) res = 0;
            else if (rb == -1) res = (long long)(0ull - l);  // avoids INT64_MIN / -1 trap semantics; Go wraps
            else res = lb / rb;
            break;
        }
      }
      st[sp++] = res;
    }
  }
  return st[0];
}

// ---- aggregate table -----------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// Exact-key open addressing: a slot is claimed by CAS on its tag (0 -> 1), the key words are
// written, then the tag is published as fingerprint|2.  Unlike the reference, which keys groups
// only by a 64-bit hash (aggregate.go:411), equal tags are confirmed against the stored key.
__device__ __forceinline__ uint32_t hash_find_or_insert(const QueryDesc& q, const unsigned long long* kw, bool* overflow) {
  const int W = q.key_words;
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (int w = 0; w < W; w++) h = mix64(h ^ kw[w]) + 0x9e3779b97f4a7c15ull * (w + 1);
  const uint32_t fp = uint32_t(h >> 32) | 2u;
  const uint32_t mask = q.table_slots - 1;
  uint32_t s = uint32_t(h) & mask;
  volatile uint32_t* tag = q.t_tag;
  volatile unsigned long long* keys = q.t_keys;
  for (uint32_t probes = 0; probes <= mask; probes++) {
    uint32_t t = tag[s];
    if (t == 0) {
      uint32_t old = atomicCAS(q.t_tag + s, 0u, 1u);
      if (old == 0) {
        for (int w = 0; w < W; w++) keys[size_t(s) * W + w] = kw[w];
        __threadfence();
        tag[s] = fp;
        return s;
      }
      t = old;
    }
    while (t == 1u) t = tag[s];
    if (t == fp) {
      bool eq = true;
      for (int w = 0; w < W; w++) eq &= (keys[size_t(s) * W + w] == kw[w]);
      if (eq) return s;
    }
    s = (s + 1) & mask;
  }
  *overflow = true;
  return 0;
}

template <typename T, typename Op>
__device__ __forceinline__ T peer_reduce(T v, unsigned peers, int lane, Op op) {
  if (peers == 0xffffffffu) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
  }
  const int cnt = __popc(peers);
  if (cnt == 1) return v;
  const int rank = __popc(peers & ((1u << lane) - 1u));
  for (int stride = 1; stride < cnt; stride <<= 1) {
    int srcrank = rank + stride;
    int src = (srcrank < cnt) ? int(__fns(peers, 0, srcrank + 1)) : lane;
    T o = __shfl_sync(peers, v, src);
    if (srcrank < cnt && (rank & (2 * stride - 1)) == 0) v = op(v, o);
  }
  return v;  // valid in the lowest lane of the peer group
}

__device__ __forceinline__ void atomic_min_f64(long long* addr, double v) {
  // Go's `if v < minV` (aggregate.go:847-857): NaN never replaces, ties keep the stored value.
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_f64(long long* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

__device__ __forceinline__ void apply_agg(uint8_t func, bool is_float, long long* cell, long long bits) {
  if (func == 1 /*sum*/) {
    if (is_float) atomicAdd(reinterpret_cast<double*>(cell), __longlong_as_double(bits));
    else atomicAdd(reinterpret_cast<unsigned long long*>(cell), (unsigned long long)bits);
  } else if (func == 2 /*min*/) {
    if (is_float) atomic_min_f64(cell, __longlong_as_double(bits));
    else atomicMin(cell, bits);
  } else if (func == 3 /*max*/) {
    if (is_float) atomic_max_f64(cell, __longlong_as_double(bits));
    else atomicMax(cell, bits);
  }
}

__device__ __forceinline__ long long reduce_agg(uint8_t func, bool is_float, long long bits, unsigned peers, int lane) {
  if (func == 1) {
    if (is_float)
      return __double_as_longlong(peer_reduce(__longlong_as_double(bits), peers, lane, [](double a, double b) { return a + b; }));
    return (long long)peer_reduce((unsigned long long)bits, peers, lane,
                                  [](unsigned long long a, unsigned long long b) { return a + b; });
  } else if (func == 2) {
    if (is_float)
      return __double_as_longlong(peer_reduce(__longlong_as_double(bits), peers, lane, [](double a, double b) { return (b < a) ? b : a; }));
    return peer_reduce(bits, peers, lane, [](long long a, long long b) { return (b < a) ? b : a; });
  } else {
    if (is_float)
      return __double_as_longlong(peer_reduce(__longlong_as_double(bits), peers, lane, [](double a, double b) { return (b > a) ? b : a; }));
    return peer_reduce(bits, peers, lane, [](long long a, long long b) { return (b > a) ? b : a; });
  }
}

}  // namespace

// ======================================================================================================
// k_scan
// ======================================================================================================
__global__ void __launch_bounds__(NT) k_scan(const QueryDesc* __restrict__ qp) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const QueryDesc& q = *qp;
  const SmemLayout sm = carve(smem_raw, q.key_words, int(q.n_numbufs));
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned long long selected_local = 0;
  bool overflow = false;

  for (uint32_t tile = blockIdx.x; tile < q.n_tiles; tile += gridDim.x) {
    // locate the row group of this tile (rg_first_tile is ascending, n_rg + 1 entries)
    int lo = 0, hi = q.n_rg;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (__ldg(&q.rg_first_tile[mid]) <= tile) lo = mid; else hi = mid;
    }
    const int rg = lo;
    const uint32_t tile_in_rg = tile - __ldg(&q.rg_first_tile[rg]);
    const uint32_t r0 = tile_in_rg * TILE;
    const uint32_t rg_rows = __ldg(&q.rg_rows[rg]);
    const uint32_t n = min(uint32_t(TILE), rg_rows - r0);
    const ChunkDesc* __restrict__ chunks = q.chunks + size_t(rg) * q.n_slots;
    const LeafRt* __restrict__ lrt = q.leaf_rt + size_t(rg) * q.n_leaves;

    // reset per-row state (strided mapping: row i = j*NT + tid is owned by one thread throughout)
#pragma unroll
    for (int j = 0; j < RPT; j++) {
      uint32_t i = j * NT + tid;
      sm.leafbits[i] = 0;
      for (int w = 0; w < q.key_words; w++) sm.keyw[size_t(w) * TILE + i] = 0;
    }
    // constant leaves (missing-column rules) contribute their bit up front
    {
      uint32_t cbits = 0;
      for (int l = 0; l < q.n_leaves; l++)
        if (lrt[l].mode == LM_ALL) cbits |= 1u << l;
      if (cbits) {
#pragma unroll
        for (int j = 0; j < RPT; j++) sm.leafbits[j * NT + tid] = cbits;
      }
    }
    __syncthreads();

    // ---- per column: decode, then feed keys / leaves / staging ------------------------------------
    for (int slot = 0; slot < q.n_slots; slot++) {
      const ChunkDesc& c = chunks[slot];
      const uint8_t stype = q.slot_type[slot];
      if (stype == ST_DICT) {
        uint32_t nv = n, v0 = r0;
        const bool absent = (c.kind == CK_ABSENT);
        const bool nulls = !absent && c.has_nulls;
        if (!absent) {
          if (nulls) {
            nv = decode_validity(c, tile_in_rg, r0, n, sm);
            v0 = __ldg(&c.tile_val0[tile_in_rg]);
          }
          decode_hybrid(c.values, c.runs, __ldg(&c.tile_run[tile_in_rg]), c.n_runs, v0, nv, sm.tmp, sm.ridx, sm.wscr);
        }
#pragma unroll
        for (int j = 0; j < RPT; j++) {
          uint32_t i = j * NT + tid;
          if (i >= n) continue;
          uint32_t idx = kNullIdx;
          if (!absent) {
            if (nulls) {
              uint32_t pos;
              if (row_valid(sm, i, &pos)) idx = sm.tmp[pos];
            } else {
              idx = sm.tmp[i];
            }
          }
          // group keys on this column
          for (int k = 0; k < q.n_keys; k++) {
            const KeyDesc& kd = q.keys[k];
            if (kd.slot != slot) continue;
            uint64_t code = (idx == kNullIdx) ? 0ull : uint64_t(__ldg(&c.lut[idx])) + 1ull;
            if (q.table_mode == TM_DENSE) sm.keyw[i] += code * kd.dense_stride;
            else sm.keyw[size_t(kd.word) * TILE + i] |= code << kd.shift;
          }
          // predicate leaves on this column
          if (q.slot_used_by_leaf[slot]) {
            uint32_t bits = 0;
            for (int l = 0; l < q.n_leaves; l++) {
              if (q.leaves[l].slot != slot || lrt[l].mode != LM_EVAL) continue;
              uint32_t r = (idx == kNullIdx) ? lrt[l].null_result : __ldg(&lrt[l].lut[idx]);
              bits |= (r & 1u) << l;
            }
            sm.leafbits[i] |= bits;
          }
        }
        __syncthreads();  // tmp / vbyte are reused by the next column
      } else {
        // numeric column: stage when nullable or dictionary-encoded, else it is read in place
        const int nb = q.slot_numbuf[slot];
        const bool staged = (nb >= 0) && c.kind != CK_ABSENT && !(c.kind == CK_PLAIN64 && !c.has_nulls);
        if (staged) {
          uint32_t nv = n, v0 = r0;
          if (c.has_nulls) {
            nv = decode_validity(c, tile_in_rg, r0, n, sm);
            v0 = __ldg(&c.tile_val0[tile_in_rg]);
          }
          if (c.kind == CK_DICT64)
            decode_hybrid(c.values, c.runs, __ldg(&c.tile_run[tile_in_rg]), c.n_runs, v0, nv, sm.tmp, sm.ridx, sm.wscr);
          long long* nbuf = sm.numbuf + size_t(nb) * TILE;
          uint32_t* nnull = sm.numnull + nb * (TILE / 32);
#pragma unroll
          for (int j = 0; j < RPT; j++) {
            uint32_t i = j * NT + tid;
            bool valid = i < n;
            uint32_t pos = i;
            if (valid && c.has_nulls) valid = row_valid(sm, i, &pos);
            long long val = 0;
            if (valid) {
              if (c.kind == CK_DICT64) val = __ldg(&c.dict64[sm.tmp[pos]]);
              else val = __ldg(reinterpret_cast<const long long*>(c.values) + v0 + pos);
            }
            nbuf[i] = val;
            unsigned nullmask = __ballot_sync(0xffffffffu, !valid);
            if (lane == 0) nnull[i >> 5] = nullmask;
          }
          __syncthreads();
        }
        // leaves on this numeric column and int64 group keys
        const bool has_key = [&] {
          for (int k = 0; k < q.n_keys; k++)
            if (q.keys[k].slot == slot) return true;
          return false;
        }();
        if (q.slot_used_by_leaf[slot] || has_key) {
#pragma unroll
          for (int j = 0; j < RPT; j++) {
            uint32_t i = j * NT + tid;
            if (i >= n) continue;
            NumVal v = load_num(q, chunks, sm, slot, r0, i);
            if (q.slot_used_by_leaf[slot]) {
              uint32_t bits = 0;
              for (int l = 0; l < q.n_leaves; l++) {
                const LeafDesc& ld = q.leaves[l];
                if (ld.slot != slot || lrt[l].mode != LM_EVAL) continue;
                bool r = false;
                if (!v.null) {  // NULL compares to NULL: not selected (binaryscalarexpr.go:143-150)
                  if (ld.cmp_float) {
                    double x = (stype == ST_F64) ? __longlong_as_double(v.bits) : double(v.bits);
                    r = cmp_f64(ld.op, x, ld.lit_f);
                  } else {
                    r = cmp_i64(ld.op, v.bits, ld.lit_i);
                  }
                }
                bits |= uint32_t(r) << l;
              }
              sm.leafbits[i] |= bits;
            }
            if (has_key) {
              for (int k = 0; k < q.n_keys; k++) {
                const KeyDesc& kd = q.keys[k];
                if (kd.slot != slot) continue;
                // NULL and 0 hash alike in the reference (dynparquet/hashed.go:254-262)
                sm.keyw[size_t(kd.word) * TILE + i] = v.null ? 0ull : (unsigned long long)v.bits;
              }
            }
          }
        }
      }
    }
    __syncthreads();

    // ---- selection + aggregation (strided mapping, whole warps stay converged for the ballots) ----
#pragma unroll 1
    for (int j = 0; j < RPT; j++) {
      uint32_t i = j * NT + tid;
      bool active = i < n;
      if (active && q.n_filter_prog > 0) active = eval_filter(q, sm.leafbits[i]);
      unsigned amask = __ballot_sync(0xffffffffu, active);
      if (amask == 0) continue;
      if (lane == 0) selected_local += __popc(amask);
      if (!active) continue;
      uint32_t slot_idx;
      if (q.table_mode == TM_DENSE) {
        slot_idx = uint32_t(sm.keyw[i]);
      } else {
        unsigned long long kw[kMaxKeyWords];
        for (int w = 0; w < q.key_words; w++) kw[w] = sm.keyw[size_t(w) * TILE + i];
        slot_idx = hash_find_or_insert(q, kw, &overflow);
      }
      unsigned peers = __match_any_sync(amask, slot_idx);
      int leader = __ffs(peers) - 1;
      if (lane == leader) atomicAdd(q.t_rows + slot_idx, (unsigned long long)__popc(peers));
      for (int a = 0; a < q.n_aggs; a++) {
        const AggDesc& ad = q.aggs[a];
        if (ad.func == 4 /*count*/) continue;  // = rows of the group (aggregate.go:937-950)
        long long bits = eval_prog(q, ad, chunks, sm, r0, i);
        bits = reduce_agg(ad.func, ad.is_float, bits, peers, lane);
        if (lane == leader) apply_agg(ad.func, ad.is_float, q.t_agg[a] + slot_idx, bits);
      }
    }
    __syncthreads();
  }
  if (lane == 0 && selected_local) atomicAdd(q.counters + 0, selected_local);
  if (overflow) atomicExch(q.counters + 1, 1ull);
}

// ======================================================================================================
// k_table_init: rows = 0, sum = 0, min = +max, max = -max (int64) / +-inf (double), tags = 0
// ======================================================================================================
__global__ void k_table_init(QueryDesc q) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t s = size_t(blockIdx.x) * blockDim.x + threadIdx.x; s < q.table_slots; s += stride) {
    q.t_rows[s] = 0;
    if (q.table_mode == TM_HASH) q.t_tag[s] = 0;
    for (int a = 0; a < q.n_aggs; a++) {
      const AggDesc& ad = q.aggs[a];
      long long init = 0;
      if (ad.func == 2) init = ad.is_float ? 0x7ff0000000000000ll : 0x7fffffffffffffffll;
      if (ad.func == 3) init = ad.is_float ? (long long)0xfff0000000000000ull : (long long)0x8000000000000000ull;
      if (ad.func != 4) q.t_agg[a][s] = init;
    }
  }
}

// ======================================================================================================
// k_finalize
// ======================================================================================================
__global__ void k_finalize(FinalizeDesc f) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t s = size_t(blockIdx.x) * blockDim.x + threadIdx.x; s < f.table_slots; s += stride) {
    unsigned long long rows = f.t_rows[s];
    if (rows == 0) continue;
    unsigned int o = atomicAdd(f.out_count, 1u);
    if (o >= f.max_out) continue;
    f.out_rows[o] = rows;
    for (int k = 0; k < f.n_keys; k++) {
      const KeyDesc& kd = f.keys[k];
      long long code;
      if (f.table_mode == TM_DENSE) {
        code = (long long)((s / kd.dense_stride) % f.dense_radix[k]);
      } else {
        unsigned long long w = f.t_keys[s * f.key_words + kd.word];
        if (kd.is_int64) code = (long long)w;
        else code = (long long)((w >> kd.shift) & ((kd.bits >= 64) ? ~0ull : ((1ull << kd.bits) - 1ull)));
      }
      f.out_keys[size_t(k) * f.max_out + o] = code;
    }
    for (int a = 0; a < f.n_aggs; a++)
      f.out_aggs[size_t(a) * f.max_out + o] = f.t_agg[a] ? f.t_agg[a][s] : (long long)rows;
  }
}

// ======================================================================================================
// k_merge: fold one remote partial table (same QueryDesc shape) into the local one.
// Partial layout (position independent): [rows u64 x S][agg_0 i64 x S]...[tags u32 x S][keys u64 x S*W]
// ======================================================================================================
__global__ void k_merge(QueryDesc q, const uint8_t* __restrict__ partial) {
  const size_t S = q.table_slots;
  const unsigned long long* p_rows = reinterpret_cast<const unsigned long long*>(partial);
  const long long* p_agg = reinterpret_cast<const long long*>(partial + S * 8);
  int n_stored = 0;
  int agg_pos[kMaxAggs];
  for (int a = 0; a < q.n_aggs; a++) agg_pos[a] = (q.aggs[a].func == 4) ? -1 : n_stored++;
  const uint32_t* p_tag = reinterpret_cast<const uint32_t*>(partial + S * 8 * (1 + n_stored));
  const unsigned long long* p_keys =
      reinterpret_cast<const unsigned long long*>(partial + S * 8 * (1 + n_stored) + ((S * 4 + 7) & ~size_t(7)));
  bool overflow = false;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t s = size_t(blockIdx.x) * blockDim.x + threadIdx.x; s < S; s += stride) {
    unsigned long long rows = p_rows[s];
    if (rows == 0) continue;
    uint32_t dst;
    if (q.table_mode == TM_DENSE) {
      dst = uint32_t(s);
    } else {
      unsigned long long kw[kMaxKeyWords];
      for (int w = 0; w < q.key_words; w++) kw[w] = p_keys[s * q.key_words + w];
      (void)p_tag;
      dst = hash_find_or_insert(q, kw, &overflow);
      if (overflow) break;
    }
    atomicAdd(q.t_rows + dst, rows);
    for (int a = 0; a < q.n_aggs; a++) {
      if (agg_pos[a] < 0) continue;
      apply_agg(q.aggs[a].func, q.aggs[a].is_float, q.t_agg[a] + dst, p_agg[size_t(agg_pos[a]) * S + s]);
    }
  }
  if (overflow) atomicExch(q.counters + 1, 1ull);
}

// ======================================================================================================
// k_decode: one column chunk -> dense buffers (K1 standalone).
//   DICT_STR : out_i32[row] = global dictionary id, -1 for NULL
//   numeric  : out_i64[row] = value (0 for NULL), out_valid[row] = 0/1
// ======================================================================================================
__global__ void __launch_bounds__(NT) k_decode(ChunkDesc c, uint32_t n_tiles, int32_t* __restrict__ out_i32,
                                               long long* __restrict__ out_i64, uint8_t* __restrict__ out_valid) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const SmemLayout sm = carve(smem_raw, 0, 0);
  const int tid = threadIdx.x;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r0 = tile * TILE;
    const uint32_t n = min(uint32_t(TILE), c.n_rows - r0);
    uint32_t nv = n, v0 = r0;
    if (c.has_nulls) {
      nv = decode_validity(c, tile, r0, n, sm);
      v0 = __ldg(&c.tile_val0[tile]);
    }
    if (c.kind == CK_DICT_STR || c.kind == CK_DICT64)
      decode_hybrid(c.values, c.runs, __ldg(&c.tile_run[tile]), c.n_runs, v0, nv, sm.tmp, sm.ridx, sm.wscr);
#pragma unroll
    for (int j = 0; j < RPT; j++) {
      uint32_t i = j * NT + tid;
      if (i >= n) continue;
      bool valid = true;
      uint32_t pos = i;
      if (c.has_nulls) valid = row_valid(sm, i, &pos);
      if (c.kind == CK_DICT_STR) {
        out_i32[r0 + i] = valid ? int32_t(__ldg(&c.lut[sm.tmp[pos]])) : -1;
      } else {
        long long v = 0;
        if (valid) {
          if (c.kind == CK_DICT64) v = __ldg(&c.dict64[sm.tmp[pos]]);
          else v = __ldg(reinterpret_cast<const long long*>(c.values) + v0 + pos);
        }
        out_i64[r0 + i] = v;
        out_valid[r0 + i] = valid ? 1 : 0;
      }
    }
    __syncthreads();
  }
}

// ======================================================================================================
// host launchers
// ======================================================================================================
size_t scan_smem_bytes(int key_words, int n_numbufs) {
  return size_t(key_words) * TILE * 8 + size_t(n_numbufs) * TILE * 8 + TILE * 4 + TILE * 4 + TILE * 2 +
         size_t(n_numbufs) * (TILE / 32) * 4 + 64 * 4 + NT * 2 + NT + 16;
}

cudaError_t launch_table_init(const QueryDesc& q, cudaStream_t st) {
  int blocks = int((size_t(q.table_slots) + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_table_init<<<blocks, 256, 0, st>>>(q);
  return cudaGetLastError();
}

cudaError_t launch_scan(const QueryDesc* d_q, const QueryDesc& q, int sm_count, cudaStream_t st) {
  if (q.n_tiles == 0) return cudaSuccess;
  size_t smem = scan_smem_bytes(q.key_words, int(q.n_numbufs));
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan, NT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = uint32_t(sm_count) * uint32_t(per_sm);
  if (grid > q.n_tiles) grid = q.n_tiles;
  k_scan<<<grid, NT, smem, st>>>(d_q);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeDesc& f, cudaStream_t st) {
  int blocks = int((size_t(f.table_slots) + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_finalize<<<blocks, 256, 0, st>>>(f);
  return cudaGetLastError();
}

cudaError_t launch_merge(const QueryDesc& q, const void* partial, cudaStream_t st) {
  int blocks = int((size_t(q.table_slots) + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_merge<<<blocks, 256, 0, st>>>(q, static_cast<const uint8_t*>(partial));
  return cudaGetLastError();
}

cudaError_t launch_decode(const ChunkDesc& c, int32_t* out_i32, long long* out_i64, uint8_t* out_valid, int sm_count,
                          cudaStream_t st) {
  uint32_t n_tiles = (c.n_rows + TILE - 1) / TILE;
  if (n_tiles == 0) return cudaSuccess;
  size_t smem = scan_smem_bytes(0, 0);
  uint32_t grid = uint32_t(sm_count) * 4;
  if (grid > n_tiles) grid = n_tiles;
  k_decode<<<grid, NT, smem, st>>>(c, n_tiles, out_i32, out_i64, out_valid);
  return cudaGetLastError();
}

}  // namespace fgpu
