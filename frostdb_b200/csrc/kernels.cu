// sm_100a kernels of libfrostgpu.
//
//   k_scan<KW>   K1+K2+K4/K5 fused.  Every consumer warp streams one 128-row chunk of a row group at a time:
//                it walks the stored Parquet encodings of the projected column chunks directly in HBM
//                (PLAIN int64/double, RLE/bit-packed hybrid definition levels and dictionary indices)
//                with per-lane run cursors seeded from a per-chunk seed, evaluates the predicate
//                leaves into per-row bits, and folds the selected rows into the aggregate table.
//                Rows whose group does not change inside a warp (the common case for parts sorted in
//                compaction order) are accumulated in registers and flushed with ONE warp-reduced
//                atomic per aggregate when the group changes; mixed groups fall back to
//                __match_any_sync peer reduction.  Replaces ParquetConverter.Convert
//                (pqarrow/arrow.go:264-373), PredicateFilter.Callback (filter.go:255-323),
//                HashAggregate.Callback (aggregate.go:263-490), Distinction.Callback (distinct.go:70-170).
//   k_rows       K1+K2+K3: same decode + predicate, then order-preserving stream compaction
//                (warp ballot + decoupled look-back prefix over tiles) of the projected columns:
//                PredicateFilter.filter() (filter.go:276-323) + Projection.
//   k_finalize   K6: aggregate table -> compacted result columns (finishAggregate, aggregate.go:543-633).
//   k_merge      K6: folds gathered partial tables into the local one (Synchronizer + final
//                HashAggregate, synchronize.go:16-53, physicalplan.go:438-471).
//   k_decode     K1 standalone: one column chunk -> dense buffers.
//
// HBM-bound integer / indexing work: no tensor cores.  In k_scan a dedicated producer warp keeps a
// shared-memory ring of tiles full with cp.async.bulk (TMA bulk copies completing on mbarriers): the
// PLAIN column slices of the tile and the 8 cursor seeds of every hybrid stream.  The 8 consumer warps
// therefore never wait on an HBM round trip in the common case; what they still read from global
// memory (run directory entries past the seed, bit-packed payload, LUTs) is small and L2 resident.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.h"
#include "kernels.h"

namespace fgpu {

namespace {

constexpr int NT = kScanThreads;     // threads per CTA
constexpr int NWARP = NT / 32;       // warps per CTA = chunks per tile
constexpr int STEPS = kIndexRows / 32;  // 32-row steps per chunk
constexpr int TILE = kTileRows;
static_assert(NWARP * kIndexRows == TILE, "tile = one chunk per warp");
static_assert(STEPS == 4, "unrolled for 4 steps");

constexpr uint32_t kNoSlot = 0xffffffffu;
constexpr uint32_t FULL = 0xffffffffu;

// ---- mbarrier / bulk-copy primitives (sm_90+ PTX) -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA, non-tensor form); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- hybrid stream cursor ------------------------------------------------------------------------------
struct HybCur {
  const Run* runs;
  const uint8_t* stream;
  const uint32_t* lut;  // string dictionaries: bit-packed index -> global id (RLE values are global ids already)
  uint32_t k, start, end, val, meta, off;
};

__device__ __forceinline__ void hc_load(HybCur& c) {
  uint4 r = __ldg(reinterpret_cast<const uint4*>(c.runs + c.k));
  c.start = r.x;
  c.off = r.y;
  c.val = r.z;
  c.meta = r.w;
  c.end = __ldg(&c.runs[c.k + 1].start);
}
// `seed` may live in shared memory (staged by the producer) or in global memory.
__device__ __forceinline__ uint32_t hc_init(HybCur& c, const Run* runs, const uint8_t* stream, const uint32_t* lut, const Seed* seed) {
  const uint4 a = *reinterpret_cast<const uint4*>(seed);
  const uint4 b = *(reinterpret_cast<const uint4*>(seed) + 1);
  c.runs = runs;
  c.stream = stream;
  c.lut = lut;
  c.k = a.x;
  c.start = a.y;
  c.end = a.z;
  c.off = a.w;
  c.val = b.x;
  c.meta = b.y;
  return b.z;  // val0
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
// Value at ordinal `ord` (ordinals are requested in increasing order per lane).
__device__ __forceinline__ uint32_t hc_get(HybCur& c, uint32_t ord) {
  while (ord >= c.end) {
    c.k++;
    hc_load(c);
  }
  if ((c.meta & 1u) == 0) return c.val;
  uint32_t w = (c.meta >> 8) & 0xffu;
  uint32_t v = extract_bits(c.stream, c.off, uint64_t(ord - c.start) * w, w);
  return c.lut ? __ldg(c.lut + v) : v;
}

// Seeds of the chunk's streams: staged copies in shared memory when the producer provided them.
struct ChunkSeeds {
  const Seed* val;  // value stream seed of this chunk (shared or global), nullptr when the column has none
  const Seed* def;  // definition-level stream seed
};
__device__ __forceinline__ ChunkSeeds global_seeds(const ChunkDesc& c, uint32_t chunk) {
  ChunkSeeds s;
  s.val = (c.kind == CK_DICT_STR || c.kind == CK_DICT64) ? c.seeds + chunk : nullptr;
  s.def = c.has_nulls ? c.def_seeds + chunk : nullptr;
  return s;
}

// Decodes the chunk's STEPS x 32 rows of a dictionary-encoded string column: gid[j] = GLOBAL dictionary
// id of row (c0 + 32 j + lane), kNullIdx for NULL / out of range / absent column.
__device__ __forceinline__ void decode_dict_chunk(const ChunkDesc& c, const ChunkSeeds& sd, uint32_t c0, uint32_t n_rows, int lane,
                                                  uint32_t (&gid)[STEPS]) {
  if (c.kind == CK_ABSENT) {
#pragma unroll
    for (int j = 0; j < STEPS; j++) gid[j] = kNullIdx;
    return;
  }
  const uint32_t lt = (1u << lane) - 1u;
  HybCur vc, dc;
  uint32_t vbase = c0;
  const bool nulls = c.has_nulls;
  if (nulls) vbase = hc_init(dc, c.def_runs, c.def, nullptr, sd.def);
  hc_init(vc, c.runs, c.values, c.lut, sd.val);
#pragma unroll
  for (int j = 0; j < STEPS; j++) {
    uint32_t r = c0 + j * 32 + lane;
    bool valid = r < n_rows;
    uint32_t ord = r;
    if (nulls) {
      if (valid) valid = hc_get(dc, r) != 0;
      unsigned m = __ballot_sync(FULL, valid);
      ord = vbase + __popc(m & lt);
      vbase += __popc(m);
    }
    gid[j] = valid ? hc_get(vc, ord) : kNullIdx;
  }
}

// Decodes the chunk's rows of a numeric column: bits[j] raw 8 bytes (0 for NULL, as
// builder.AppendValue leaves NULL slots: pqarrow/builder/utils.go:54-58, optbuilders.go:337-340),
// nullmask bit j set when the row is NULL / absent / out of range.  `staged` is the tile's slice of a
// PLAIN column in shared memory (indexed by row - tile_r0) or nullptr.
__device__ __forceinline__ void decode_num_chunk(const ChunkDesc& c, const ChunkSeeds& sd, const long long* staged, uint32_t tile_r0,
                                                 uint32_t c0, uint32_t n_rows, int lane, long long (&bits)[STEPS], uint32_t& nullmask) {
  nullmask = 0;
  if (c.kind == CK_ABSENT) {
#pragma unroll
    for (int j = 0; j < STEPS; j++) bits[j] = 0;
    nullmask = (1u << STEPS) - 1u;
    return;
  }
  const long long* vals = reinterpret_cast<const long long*>(c.values);
  if (c.kind == CK_PLAIN64 && !c.has_nulls) {
#pragma unroll
    for (int j = 0; j < STEPS; j++) {
      uint32_t r = c0 + j * 32 + lane;
      bool inb = r < n_rows;
      long long v = 0;
      if (inb) v = staged ? staged[r - tile_r0] : __ldg(vals + r);
      bits[j] = v;
      if (!inb) nullmask |= 1u << j;
    }
    return;
  }
  const uint32_t lt = (1u << lane) - 1u;
  HybCur vc, dc;
  uint32_t vbase = c0;
  const bool nulls = c.has_nulls;
  if (nulls) vbase = hc_init(dc, c.def_runs, c.def, nullptr, sd.def);
  const bool dict = c.kind == CK_DICT64;
  if (dict) hc_init(vc, c.runs, c.values, nullptr, sd.val);
#pragma unroll
  for (int j = 0; j < STEPS; j++) {
    uint32_t r = c0 + j * 32 + lane;
    bool valid = r < n_rows;
    uint32_t ord = r;
    if (nulls) {
      if (valid) valid = hc_get(dc, r) != 0;
      unsigned m = __ballot_sync(FULL, valid);
      ord = vbase + __popc(m & lt);
      vbase += __popc(m);
    }
    long long v = 0;
    if (valid) v = dict ? __ldg(&c.dict64[hc_get(vc, ord)]) : __ldg(vals + ord);
    bits[j] = v;
    if (!valid) nullmask |= 1u << j;
  }
}

// What the consumer knows about the tile it is working on.
struct TileCtx {
  const QueryDesc* q;
  const ChunkDesc* chunks;  // [n_slots] of the row group
  const LeafRt* lrt;        // [n_leaves] of the row group
  const uint8_t* stage;     // shared-memory stage of this tile (nullptr: nothing staged, read HBM)
  uint32_t tile_r0;         // first row of the tile inside the row group
  uint32_t chunk;           // this warp's chunk index inside the row group
  uint32_t chunk_in_tile;
  uint32_t c0;              // first row of the chunk
  uint32_t n_rows;          // rows of the row group
};
__device__ __forceinline__ size_t stage_plain_bytes() { return size_t(TILE) * 8; }
__device__ __forceinline__ size_t stage_seed_bytes() { return size_t(NWARP) * sizeof(Seed); }
__device__ __forceinline__ const long long* staged_plain(const TileCtx& t, int slot) {
  if (!t.stage) return nullptr;
  int p = t.q->slot_plain_stage[slot];
  if (p < 0) return nullptr;
  const ChunkDesc& c = t.chunks[slot];
  if (c.kind != CK_PLAIN64 || c.has_nulls) return nullptr;
  return reinterpret_cast<const long long*>(t.stage + size_t(p) * stage_plain_bytes());
}
__device__ __forceinline__ ChunkSeeds seeds_for(const TileCtx& t, int slot) {
  const ChunkDesc& c = t.chunks[slot];
  ChunkSeeds s = global_seeds(c, t.chunk);
  if (t.stage) {
    const uint8_t* base = t.stage + size_t(t.q->n_stage_plain) * stage_plain_bytes();
    int sv = t.q->slot_seed_stage[slot][0], sdf = t.q->slot_seed_stage[slot][1];
    if (s.val && sv >= 0) s.val = reinterpret_cast<const Seed*>(base + size_t(sv) * stage_seed_bytes()) + t.chunk_in_tile;
    if (s.def && sdf >= 0) s.def = reinterpret_cast<const Seed*>(base + size_t(sdf) * stage_seed_bytes()) + t.chunk_in_tile;
  }
  return s;
}
__device__ __forceinline__ void tile_dict(const TileCtx& t, int slot, int lane, uint32_t (&gid)[STEPS]) {
  decode_dict_chunk(t.chunks[slot], seeds_for(t, slot), t.c0, t.n_rows, lane, gid);
}
__device__ __forceinline__ void tile_num(const TileCtx& t, int slot, int lane, long long (&bits)[STEPS], uint32_t& nullmask) {
  decode_num_chunk(t.chunks[slot], seeds_for(t, slot), staged_plain(t, slot), t.tile_r0, t.c0, t.n_rows, lane, bits, nullmask);
}

// ---- predicate --------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cmp_i64(uint8_t op, long long a, long long b) {
  switch (op) {
    case 1: return a == b;
    case 2: return a != b;
    case 3: return a < b;
    case 4: return a <= b;
    case 5: return a > b;
    default: return a >= b;
  }
}
__device__ __forceinline__ bool cmp_f64(uint8_t op, double a, double b) {
  switch (op) {
    case 1: return a == b;
    case 2: return a != b;
    case 3: return a < b;
    case 4: return a <= b;
    case 5: return a > b;
    default: return a >= b;
  }
}

__device__ __forceinline__ bool eval_filter(const QueryDesc& q, uint32_t bits) {
  if (q.filter_kind == FK_AND) return (bits & q.filter_mask) == q.filter_mask;
  if (q.filter_kind == FK_OR) return (bits & q.filter_mask) != 0;
  uint32_t stack = 0;  // postfix program over leaf bits; stack kept in a 32-bit word
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

// Evaluates every predicate leaf of the chunk: leafbits[j] bit l = leaf l selects row j.
// NULL semantics: binaryscalarexpr.go:143-150 (numeric NULL never selected), :165-172,:205-212
// (dictionary == NULL / != NULL), missing columns :47-73 (LM_ALL / LM_NONE precomputed on the host).
__device__ __forceinline__ void eval_leaves(const TileCtx& t, int lane, uint32_t (&leafbits)[STEPS]) {
  const QueryDesc& q = *t.q;
  const LeafRt* __restrict__ lrt = t.lrt;
  uint32_t cbits = 0;
  for (int l = 0; l < q.n_leaves; l++)
    if (lrt[l].mode == LM_ALL) cbits |= 1u << l;
#pragma unroll
  for (int j = 0; j < STEPS; j++) leafbits[j] = cbits;
  for (int slot = 0; slot < q.n_slots; slot++) {
    if (!q.slot_used_by_leaf[slot]) continue;
    if (q.slot_type[slot] == ST_DICT) {
      uint32_t gid[STEPS];
      tile_dict(t, slot, lane, gid);
      for (int l = 0; l < q.n_leaves; l++) {
        if (q.leaves[l].slot != slot || lrt[l].mode != LM_EVAL) continue;
        const uint8_t* lut = lrt[l].lut;
        const uint32_t nullres = lrt[l].null_result;
#pragma unroll
        for (int j = 0; j < STEPS; j++) {
          uint32_t r = (gid[j] == kNullIdx) ? nullres : uint32_t(__ldg(lut + gid[j]));
          leafbits[j] |= (r & 1u) << l;
        }
      }
    } else {
      long long bits[STEPS];
      uint32_t nullmask;
      tile_num(t, slot, lane, bits, nullmask);
      const bool f64col = q.slot_type[slot] == ST_F64;
      for (int l = 0; l < q.n_leaves; l++) {
        const LeafDesc& ld = q.leaves[l];
        if (ld.slot != slot || lrt[l].mode != LM_EVAL) continue;
#pragma unroll
        for (int j = 0; j < STEPS; j++) {
          bool r = false;
          if (!((nullmask >> j) & 1u)) {
            if (ld.cmp_float) {
              double x = f64col ? __longlong_as_double(bits[j]) : double(bits[j]);
              r = cmp_f64(ld.op, x, ld.lit_f);
            } else {
              r = cmp_i64(ld.op, bits[j], ld.lit_i);
            }
          }
          leafbits[j] |= uint32_t(r) << l;
        }
      }
    }
  }
}

// Selected rows of the chunk: bit j = this lane's row of step j passes the predicate.
__device__ __forceinline__ uint32_t chunk_selection(const TileCtx& t, int lane) {
  const QueryDesc& q = *t.q;
  uint32_t actbits = 0;
  if (q.n_filter_prog > 0) {
    uint32_t leafbits[STEPS];
    eval_leaves(t, lane, leafbits);
#pragma unroll
    for (int j = 0; j < STEPS; j++)
      if (t.c0 + j * 32 + lane < t.n_rows && eval_filter(q, leafbits[j])) actbits |= 1u << j;
  } else {
#pragma unroll
    for (int j = 0; j < STEPS; j++)
      if (t.c0 + j * 32 + lane < t.n_rows) actbits |= 1u << j;
  }
  return actbits;
}

// ---- aggregate expressions ----------------------------------------------------------------------------------
// Arithmetic ignores validity and Div by zero yields NULL, i.e. a raw 0 in the aggregated array
// (query/physicalplan/project.go:169-395, :216-218).
__device__ __forceinline__ long long apply_arith(uint8_t op, bool is_float, long long lb, long long rb) {
  if (is_float) {
    double l = __longlong_as_double(lb), r = __longlong_as_double(rb), x;
    switch (op) {
      case PO_ADD: x = l + r; break;
      case PO_SUB: x = l - r; break;
      case PO_MUL: x = l * r; break;
      default: x = (r == 0.0) ? 0.0 : l / r; break;
    }
    return __double_as_longlong(x);
  }
  unsigned long long l = (unsigned long long)lb, r = (unsigned long long)rb;
  switch (op) {
    case PO_ADD: return (long long)(l + r);
    case PO_SUB: return (long long)(l - r);
    case PO_MUL: return (long long)(l * r);
    default:
      if (rb == 0) return 0;
      if (rb == -1) return (long long)(0ull - l);  // Go wraps INT64_MIN / -1
      return lb / rb;
  }
}

// Values of aggregate `a` for the chunk's rows.  Evaluated column-at-a-time over a small operand stack
// of 8-row vectors (depth <= 3; deeper expressions are rejected by the host).
__device__ __forceinline__ void eval_agg_values(const TileCtx& t, const AggDesc& a, int lane, long long (&out)[STEPS]) {
  const QueryDesc& q = *t.q;
  long long s1[STEPS], s2[STEPS];
  int sp = 0;
  for (int p = a.prog_off; p < a.prog_off + a.prog_len; p++) {
    const ProgOp& o = q.prog[p];
    if (o.op == PO_LOAD || o.op == PO_CONST) {
      long long v[STEPS];
      if (o.op == PO_LOAD) {
        uint32_t nm;
        tile_num(t, o.slot, lane, v, nm);
      } else {
#pragma unroll
        for (int j = 0; j < STEPS; j++) v[j] = o.imm;
      }
      if (sp == 0) {
#pragma unroll
        for (int j = 0; j < STEPS; j++) out[j] = v[j];
      } else if (sp == 1) {
#pragma unroll
        for (int j = 0; j < STEPS; j++) s1[j] = v[j];
      } else {
#pragma unroll
        for (int j = 0; j < STEPS; j++) s2[j] = v[j];
      }
      sp++;
    } else {
      if (sp == 2) {
#pragma unroll
        for (int j = 0; j < STEPS; j++) out[j] = apply_arith(o.op, a.is_float, out[j], s1[j]);
      } else {
#pragma unroll
        for (int j = 0; j < STEPS; j++) s1[j] = apply_arith(o.op, a.is_float, s1[j], s2[j]);
      }
      sp--;
    }
  }
}

// ---- aggregate table ----------------------------------------------------------------------------------------
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
template <int KW>
__device__ __forceinline__ uint32_t hash_find_or_insert(const QueryDesc& q, const unsigned long long (&kw)[KW], bool* overflow) {
  const int W = q.key_words;
  uint64_t h = 0x9e3779b97f4a7c15ull;
#pragma unroll
  for (int w = 0; w < KW; w++)
    if (w < W) h = mix64(h ^ kw[w]) + 0x9e3779b97f4a7c15ull * (w + 1);
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
#pragma unroll
        for (int w = 0; w < KW; w++)
          if (w < W) keys[size_t(s) * W + w] = kw[w];
        __threadfence();
        tag[s] = fp;
        return s;
      }
      t = old;
    }
    while (t == 1u) t = tag[s];
    if (t == fp) {
      bool eq = true;
#pragma unroll
      for (int w = 0; w < KW; w++)
        if (w < W) eq &= (keys[size_t(s) * W + w] == kw[w]);
      if (eq) return s;
    }
    s = (s + 1) & mask;
  }
  *overflow = true;
  return 0;
}

template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v = op(v, __shfl_xor_sync(FULL, v, d));
  return v;
}

template <typename T, typename Op>
__device__ __forceinline__ T peer_reduce(T v, unsigned peers, int lane, Op op) {
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

__device__ __forceinline__ long long agg_identity(uint8_t func, bool is_float) {
  if (func == 2) return is_float ? 0x7ff0000000000000ll : 0x7fffffffffffffffll;
  if (func == 3) return is_float ? (long long)0xfff0000000000000ull : (long long)0x8000000000000000ull;
  return 0;  // sum: +0 (int) / +0.0 (double)
}

__device__ __forceinline__ long long agg_combine(uint8_t func, bool is_float, long long a, long long b) {
  if (func == 1) {
    if (is_float) return __double_as_longlong(__longlong_as_double(a) + __longlong_as_double(b));
    return (long long)((unsigned long long)a + (unsigned long long)b);
  }
  if (func == 2) {
    if (is_float) return (__longlong_as_double(b) < __longlong_as_double(a)) ? b : a;
    return (b < a) ? b : a;
  }
  if (is_float) return (__longlong_as_double(b) > __longlong_as_double(a)) ? b : a;
  return (b > a) ? b : a;
}

// Warp-reduces the per-lane partial of one aggregate and lets lane 0 apply it to the table cell.
__device__ __noinline__ void flush_agg(uint8_t func, bool is_float, long long* cell, long long acc, int lane) {
  acc = warp_reduce(acc, [=](long long a, long long b) { return agg_combine(func, is_float, a, b); });
  if (lane == 0) apply_agg(func, is_float, cell, acc);
}

// One step whose active lanes belong to different groups: peer reduction per distinct group.
__device__ __noinline__ void mixed_agg(uint8_t func, bool is_float, long long* column, uint32_t slot, bool active, long long bits,
                                       int lane) {
  unsigned amask = __ballot_sync(FULL, active);
  if (!active) return;
  unsigned peers = __match_any_sync(amask, slot);
  long long r = peer_reduce(bits, peers, lane, [=](long long a, long long b) { return agg_combine(func, is_float, a, b); });
  if (lane == __ffs(peers) - 1) apply_agg(func, is_float, column + slot, r);
}
__device__ __noinline__ void mixed_rows(unsigned long long* rows, uint32_t slot, bool active, int lane) {
  unsigned amask = __ballot_sync(FULL, active);
  if (!active) return;
  unsigned peers = __match_any_sync(amask, slot);
  if (lane == __ffs(peers) - 1) atomicAdd(rows + slot, (unsigned long long)__popc(peers));
}

// Locates the row group of a tile (rg_first_tile ascending, n_rg + 1 entries).
__device__ __forceinline__ int find_rg(const uint32_t* __restrict__ first_tile, int n_rg, uint32_t tile) {
  int lo = 0, hi = n_rg;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(&first_tile[mid]) <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

}  // namespace

// ======================================================================================================
// k_scan
// ======================================================================================================
// Shared-memory ring: n_stages stages, each = n_stage_plain x (TILE x 8 B) + n_stage_seeds x (8 seeds x 32 B).
// Warps 0..7 consume, warp 8 produces.  full[s]: 1 arrival (producer, with expect_tx); empty[s]: 8 arrivals.
constexpr int kMaxStages = 8;

__device__ __forceinline__ size_t stage_bytes(const QueryDesc& q) {
  return size_t(q.n_stage_plain) * stage_plain_bytes() + size_t(q.n_stage_seeds) * stage_seed_bytes();
}

template <int KW>
__global__ void __launch_bounds__(NT + 32) k_scan(const QueryDesc* __restrict__ qp) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ QueryDesc sq;
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(qp);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sq);
    for (uint32_t i = threadIdx.x; i < sizeof(QueryDesc) / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const QueryDesc& q = sq;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = q.n_stages;
  const size_t sbytes = stage_bytes(q);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NWARP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NWARP) {
    // ================= producer warp: keeps the ring full =================
    if (lane == 0 && S > 0) {
      uint32_t it = 0;
      for (uint32_t tile = blockIdx.x; tile < q.n_tiles; tile += gridDim.x, it++) {
        const int st = int(it % uint32_t(S));
        const uint32_t ph = (it / uint32_t(S)) & 1u;
        mbar_wait(&empty_bar[st], ph ^ 1u);
        const int rg = find_rg(q.rg_first_tile, q.n_rg, tile);
        const uint32_t tile_in_rg = tile - __ldg(&q.rg_first_tile[rg]);
        const uint32_t n_rows = __ldg(&q.rg_rows[rg]);
        const uint32_t r0 = tile_in_rg * TILE;
        const uint32_t n = min(uint32_t(TILE), n_rows - r0);
        const uint32_t n_chunks = (n + kIndexRows - 1) / kIndexRows;
        const ChunkDesc* __restrict__ chunks = q.chunks + size_t(rg) * q.n_slots;
        uint8_t* stage = ring + size_t(st) * sbytes;
        // pass 1: bytes of this stage
        uint32_t bytes = 0;
        const uint32_t plain_sz = (n * 8u + 15u) & ~15u;
        const uint32_t seed_sz = n_chunks * uint32_t(sizeof(Seed));
        for (int p = 0; p < q.n_stage_plain; p++) {
          const ChunkDesc& c = chunks[q.stage_plain_slot[p]];
          if (c.kind == CK_PLAIN64 && !c.has_nulls) bytes += plain_sz;
        }
        for (int t = 0; t < q.n_stage_seeds; t++) {
          const ChunkDesc& c = chunks[q.stage_seed_slot[t]];
          const bool present = q.stage_seed_is_def[t] ? (c.kind != CK_ABSENT && c.has_nulls) : (c.kind == CK_DICT_STR || c.kind == CK_DICT64);
          if (present) bytes += seed_sz;
        }
        mbar_expect_tx(&full_bar[st], bytes);
        // pass 2: issue the copies
        for (int p = 0; p < q.n_stage_plain; p++) {
          const ChunkDesc& c = chunks[q.stage_plain_slot[p]];
          if (c.kind == CK_PLAIN64 && !c.has_nulls)
            bulk_g2s(stage + size_t(p) * stage_plain_bytes(), c.values + size_t(r0) * 8, plain_sz, &full_bar[st]);
        }
        uint8_t* seed_base = stage + size_t(q.n_stage_plain) * stage_plain_bytes();
        for (int t = 0; t < q.n_stage_seeds; t++) {
          const ChunkDesc& c = chunks[q.stage_seed_slot[t]];
          const bool is_def = q.stage_seed_is_def[t];
          const bool present = is_def ? (c.kind != CK_ABSENT && c.has_nulls) : (c.kind == CK_DICT_STR || c.kind == CK_DICT64);
          if (present)
            bulk_g2s(seed_base + size_t(t) * stage_seed_bytes(), (is_def ? c.def_seeds : c.seeds) + size_t(tile_in_rg) * NWARP, seed_sz,
                     &full_bar[st]);
        }
      }
    }
    return;
  }

  // ================= consumer warps =================
  const bool dense = q.table_mode == TM_DENSE;
  // warp-uniform running group + per-lane partial aggregates carried across chunks
  uint32_t cur_slot = kNoSlot;
  uint32_t cur_cnt = 0;
  long long acc[kMaxAggs];
#pragma unroll 1
  for (int a = 0; a < q.n_aggs; a++) acc[a] = agg_identity(q.aggs[a].func, q.aggs[a].is_float);
  unsigned long long selected_local = 0;
  bool overflow = false;
  // hash mode: last key -> slot of this lane
  unsigned long long last_kw[KW];
  uint32_t last_slot = kNoSlot;
#pragma unroll
  for (int w = 0; w < KW; w++) last_kw[w] = ~0ull;

  uint32_t it = 0;
  for (uint32_t tile = blockIdx.x; tile < q.n_tiles; tile += gridDim.x, it++) {
    const int st = S > 0 ? int(it % uint32_t(S)) : 0;
    const uint32_t ph = S > 0 ? (it / uint32_t(S)) & 1u : 0;
    const int rg = find_rg(q.rg_first_tile, q.n_rg, tile);
    const uint32_t tile_in_rg = tile - __ldg(&q.rg_first_tile[rg]);
    TileCtx t;
    t.q = &q;
    t.n_rows = __ldg(&q.rg_rows[rg]);
    t.tile_r0 = tile_in_rg * TILE;
    t.chunk_in_tile = uint32_t(warp);
    t.chunk = tile_in_rg * NWARP + warp;  // 128-row chunk index inside the row group
    t.c0 = t.chunk * kIndexRows;
    t.chunks = q.chunks + size_t(rg) * q.n_slots;
    t.lrt = q.leaf_rt + size_t(rg) * q.n_leaves;
    t.stage = S > 0 ? ring + size_t(st) * sbytes : nullptr;
    if (S > 0) mbar_wait(&full_bar[st], ph);

    if (t.c0 < t.n_rows) {
      // ---- predicate -> active mask per step ----------------------------------------------------
      const uint32_t actbits = chunk_selection(t, lane);
      if (__ballot_sync(FULL, actbits != 0) != 0) {
        // ---- group key per row -------------------------------------------------------------------
        unsigned long long kw[KW][STEPS];
#pragma unroll
        for (int w = 0; w < KW; w++)
#pragma unroll
          for (int j = 0; j < STEPS; j++) kw[w][j] = 0;
        for (int k = 0; k < q.n_keys; k++) {
          const KeyDesc& kd = q.keys[k];
          if (kd.is_int64) {
            long long v[STEPS];
            uint32_t nm;
            tile_num(t, kd.slot, lane, v, nm);
            // NULL and 0 hash alike in the reference (dynparquet/hashed.go:254-262): NULL slots hold 0
#pragma unroll
            for (int w = 0; w < KW; w++)
              if (w == kd.word)
#pragma unroll
                for (int j = 0; j < STEPS; j++) kw[w][j] = (unsigned long long)v[j];
          } else {
            uint32_t gid[STEPS];
            tile_dict(t, kd.slot, lane, gid);
#pragma unroll
            for (int j = 0; j < STEPS; j++) {
              unsigned long long code = (gid[j] == kNullIdx) ? 0ull : (unsigned long long)gid[j] + 1ull;
              if (dense) {
                kw[0][j] += code * kd.dense_stride;
              } else {
#pragma unroll
                for (int w = 0; w < KW; w++)
                  if (w == kd.word) kw[w][j] |= code << kd.shift;
              }
            }
          }
        }
        // ---- table slot per row --------------------------------------------------------------------
        uint32_t slot[STEPS];
#pragma unroll
        for (int j = 0; j < STEPS; j++) {
          slot[j] = kNoSlot;
          if (!((actbits >> j) & 1u)) continue;
          if (dense) {
            slot[j] = uint32_t(kw[0][j]);
          } else {
            bool same = last_slot != kNoSlot;
#pragma unroll
            for (int w = 0; w < KW; w++) same &= (kw[w][j] == last_kw[w]);
            if (!same) {
              unsigned long long key[KW];
#pragma unroll
              for (int w = 0; w < KW; w++) key[w] = kw[w][j];
              last_slot = hash_find_or_insert<KW>(q, key, &overflow);
#pragma unroll
              for (int w = 0; w < KW; w++) last_kw[w] = key[w];
            }
            slot[j] = last_slot;
          }
        }
        // ---- step classification: empty, one group for the whole warp (uslot), or mixed -------------
        uint32_t uslot[STEPS];   // warp-uniform: group of the step, kNoSlot when mixed or empty
        uint32_t mixedbits = 0;  // warp-uniform: bit j set when the step's active lanes span several groups
#pragma unroll
        for (int j = 0; j < STEPS; j++) {
          bool active = (actbits >> j) & 1u;
          unsigned amask = __ballot_sync(FULL, active);
          uslot[j] = kNoSlot;
          if (amask == 0) continue;
          if (lane == 0) selected_local += __popc(amask);
          uint32_t s = __shfl_sync(FULL, slot[j], __ffs(amask) - 1);
          bool uni = __all_sync(FULL, !active || slot[j] == s);
          if (uni) uslot[j] = s;
          else mixedbits |= 1u << j;
        }
        // ---- rows-per-group counter (also every Count aggregate, aggregate.go:937-950) ---------------
        {
          uint32_t cs = cur_slot;
#pragma unroll
          for (int j = 0; j < STEPS; j++) {
            bool active = (actbits >> j) & 1u;
            if ((mixedbits >> j) & 1u) {
              if (cs != kNoSlot) {
                uint32_t tt = __reduce_add_sync(FULL, cur_cnt);
                if (lane == 0 && tt) atomicAdd(q.t_rows + cs, (unsigned long long)tt);
                cur_cnt = 0;
                cs = kNoSlot;
              }
              mixed_rows(q.t_rows, slot[j], active, lane);
            } else if (uslot[j] != kNoSlot) {
              if (uslot[j] != cs) {
                if (cs != kNoSlot) {
                  uint32_t tt = __reduce_add_sync(FULL, cur_cnt);
                  if (lane == 0 && tt) atomicAdd(q.t_rows + cs, (unsigned long long)tt);
                  cur_cnt = 0;
                }
                cs = uslot[j];
              }
              cur_cnt += active ? 1u : 0u;
            }
          }
        }
        // ---- aggregates: per-lane partials while the warp stays in one group --------------------------
        // (acc[] is indexed dynamically on purpose: it is touched once per chunk, the per-row work runs on `part`)
#pragma unroll 1
        for (int a = 0; a < q.n_aggs; a++) {
          const AggDesc& ad = q.aggs[a];
          if (ad.func == 4 /*count*/) continue;
          long long v[STEPS];
          eval_agg_values(t, ad, lane, v);
          uint32_t cs = cur_slot;
          long long part = acc[a];
          const long long ident = agg_identity(ad.func, ad.is_float);
#pragma unroll
          for (int j = 0; j < STEPS; j++) {
            bool active = (actbits >> j) & 1u;
            if ((mixedbits >> j) & 1u) {
              if (cs != kNoSlot) {
                flush_agg(ad.func, ad.is_float, q.t_agg[a] + cs, part, lane);
                part = ident;
                cs = kNoSlot;
              }
              mixed_agg(ad.func, ad.is_float, q.t_agg[a], slot[j], active, v[j], lane);
            } else if (uslot[j] != kNoSlot) {
              if (uslot[j] != cs) {
                if (cs != kNoSlot) {
                  flush_agg(ad.func, ad.is_float, q.t_agg[a] + cs, part, lane);
                  part = ident;
                }
                cs = uslot[j];
              }
              if (active) part = agg_combine(ad.func, ad.is_float, part, v[j]);
            }
          }
          acc[a] = part;
        }
        // the running group after this chunk (identical for every aggregate by construction)
#pragma unroll
        for (int j = 0; j < STEPS; j++) {
          if ((mixedbits >> j) & 1u) cur_slot = kNoSlot;
          else if (uslot[j] != kNoSlot) cur_slot = uslot[j];
        }
      }
    }
    // this warp is done with the stage
    if (S > 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
  }
  // ---- final flush ---------------------------------------------------------------------------------
  if (cur_slot != kNoSlot) {
    uint32_t tt = __reduce_add_sync(FULL, cur_cnt);
    if (lane == 0 && tt) atomicAdd(q.t_rows + cur_slot, (unsigned long long)tt);
#pragma unroll 1
    for (int a = 0; a < q.n_aggs; a++) {
      if (q.aggs[a].func == 4) continue;
      flush_agg(q.aggs[a].func, q.aggs[a].is_float, q.t_agg[a] + cur_slot, acc[a], lane);
    }
  }
  if (lane == 0 && selected_local) atomicAdd(q.counters + 0, selected_local);
  if (overflow) atomicExch(q.counters + 1, 1ull);
}

// ======================================================================================================
// k_rows: predicate + order-preserving compaction of the projected columns
// ======================================================================================================
__global__ void __launch_bounds__(NT) k_rows(const QueryDesc* __restrict__ qp) {
  __shared__ QueryDesc sq;
  __shared__ uint32_t warp_cnt[NWARP];
  __shared__ unsigned long long tile_base_s;
  __shared__ uint32_t tile_ticket;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(qp);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sq);
    for (uint32_t i = threadIdx.x; i < sizeof(QueryDesc) / 4; i += NT) dst[i] = src[i];
  }
  __syncthreads();
  const QueryDesc& q = sq;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  volatile unsigned long long* state = q.tile_state;  // [n_tiles]: bits 62..63 flag (1 aggregate, 2 prefix), low bits value

  for (;;) {
    // tiles are handed out in order so that every predecessor of a tile is running or finished
    if (threadIdx.x == 0) tile_ticket = atomicAdd(reinterpret_cast<unsigned int*>(q.counters + 2), 1u);
    __syncthreads();
    const uint32_t tile = tile_ticket;
    if (tile >= q.n_tiles) break;
    const int rg = find_rg(q.rg_first_tile, q.n_rg, tile);
    const uint32_t tile_in_rg = tile - __ldg(&q.rg_first_tile[rg]);
    TileCtx t;
    t.q = &q;
    t.n_rows = __ldg(&q.rg_rows[rg]);
    t.tile_r0 = tile_in_rg * TILE;
    t.chunk_in_tile = uint32_t(warp);
    t.chunk = tile_in_rg * NWARP + warp;
    t.c0 = t.chunk * kIndexRows;
    t.chunks = q.chunks + size_t(rg) * q.n_slots;
    t.lrt = q.leaf_rt + size_t(rg) * q.n_leaves;
    t.stage = nullptr;

    uint32_t actbits = 0;
    if (t.c0 < t.n_rows) actbits = chunk_selection(t, lane);
    // rank of every selected row inside the warp's chunk, in row order
    uint32_t rank[STEPS];
    uint32_t wtotal = 0;
#pragma unroll
    for (int j = 0; j < STEPS; j++) {
      unsigned m = __ballot_sync(FULL, (actbits >> j) & 1u);
      rank[j] = wtotal + __popc(m & lt);
      wtotal += __popc(m);
    }
    if (lane == 0) warp_cnt[warp] = wtotal;
    __syncthreads();
    uint32_t wbase = 0, ttotal = 0;
    for (int w = 0; w < NWARP; w++) {
      if (w < warp) wbase += warp_cnt[w];
      ttotal += warp_cnt[w];
    }
    // decoupled look-back over the preceding tiles (one warp, 32 predecessors per probe)
    if (warp == 0) {
      if (lane == 0) state[tile] = (1ull << 62) | ttotal;
      __threadfence();
      unsigned long long excl = 0;
      int64_t look = int64_t(tile) - 1;
      while (look >= 0) {
        int64_t tt = look - lane;
        unsigned long long s = (tt >= 0) ? state[tt] : (2ull << 62);
        while (__any_sync(FULL, (s >> 62) == 0)) s = (tt >= 0) ? state[tt] : (2ull << 62);  // predecessors publish soon
        unsigned pmask = __ballot_sync(FULL, (s >> 62) == 2);
        int first_prefix = pmask ? __ffs(pmask) - 1 : 32;
        unsigned long long contrib = (lane <= first_prefix) ? (s & ((1ull << 62) - 1)) : 0;
        contrib = warp_reduce(contrib, [](unsigned long long a, unsigned long long b) { return a + b; });
        excl += contrib;
        if (pmask) break;
        look -= 32;
      }
      if (lane == 0) {
        __threadfence();
        state[tile] = (2ull << 62) | (excl + ttotal);
        tile_base_s = excl;
      }
    }
    __syncthreads();
    const unsigned long long base = tile_base_s + wbase;
    // ---- write the projected columns of the selected rows ---------------------------------------
    if (t.c0 < t.n_rows && wtotal > 0) {
      for (int o = 0; o < q.n_out; o++) {
        const int slot = q.out_slot[o];
        if (q.slot_type[slot] == ST_DICT) {
          uint32_t gid[STEPS];
          tile_dict(t, slot, lane, gid);
          int32_t* out = reinterpret_cast<int32_t*>(q.out_data[o]);
#pragma unroll
          for (int j = 0; j < STEPS; j++)
            if ((actbits >> j) & 1u) out[base + rank[j]] = (gid[j] == kNullIdx) ? -1 : int32_t(gid[j]);
        } else {
          long long v[STEPS];
          uint32_t nm;
          tile_num(t, slot, lane, v, nm);
          long long* out = reinterpret_cast<long long*>(q.out_data[o]);
          uint8_t* valid = q.out_valid[o];
#pragma unroll
          for (int j = 0; j < STEPS; j++)
            if ((actbits >> j) & 1u) {
              out[base + rank[j]] = v[j];
              valid[base + rank[j]] = ((nm >> j) & 1u) ? 0 : 1;
            }
        }
      }
    }
    if (threadIdx.x == 0 && tile + 1 == q.n_tiles) q.counters[0] = tile_base_s + ttotal;  // total selected rows
    __syncthreads();
  }
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
      if (ad.func != 4) q.t_agg[a][s] = agg_identity(ad.func, ad.is_float);
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
// Partial layout (position independent): [rows u64 x S][stored agg i64 x S]...[tags u32 x S (8-aligned)][keys u64 x S*W]
// ======================================================================================================
__global__ void k_merge(QueryDesc q, const uint8_t* __restrict__ partial) {
  const size_t S = q.table_slots;
  const unsigned long long* p_rows = reinterpret_cast<const unsigned long long*>(partial);
  const long long* p_agg = reinterpret_cast<const long long*>(partial + S * 8);
  int n_stored = 0;
  int agg_pos[kMaxAggs];
  for (int a = 0; a < q.n_aggs; a++) agg_pos[a] = (q.aggs[a].func == 4) ? -1 : n_stored++;
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
      for (int w = 0; w < kMaxKeyWords; w++) kw[w] = (w < q.key_words) ? p_keys[s * q.key_words + w] : 0;
      dst = hash_find_or_insert<kMaxKeyWords>(q, kw, &overflow);
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
__global__ void __launch_bounds__(NT) k_decode(ChunkDesc c, uint32_t n_chunks, int32_t* __restrict__ out_i32,
                                               long long* __restrict__ out_i64, uint8_t* __restrict__ out_valid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t chunk = blockIdx.x * NWARP + warp; chunk < n_chunks; chunk += gridDim.x * NWARP) {
    const uint32_t c0 = chunk * kIndexRows;
    const ChunkSeeds sd = global_seeds(c, chunk);
    if (c.kind == CK_DICT_STR) {
      uint32_t gid[STEPS];
      decode_dict_chunk(c, sd, c0, c.n_rows, lane, gid);
#pragma unroll
      for (int j = 0; j < STEPS; j++) {
        uint32_t r = c0 + j * 32 + lane;
        if (r < c.n_rows) out_i32[r] = (gid[j] == kNullIdx) ? -1 : int32_t(gid[j]);
      }
    } else {
      long long v[STEPS];
      uint32_t nm;
      decode_num_chunk(c, sd, nullptr, 0, c0, c.n_rows, lane, v, nm);
#pragma unroll
      for (int j = 0; j < STEPS; j++) {
        uint32_t r = c0 + j * 32 + lane;
        if (r < c.n_rows) {
          out_i64[r] = v[j];
          out_valid[r] = ((nm >> j) & 1u) ? 0 : 1;
        }
      }
    }
  }
}

// ======================================================================================================
// host launchers
// ======================================================================================================
namespace {
template <int KW>
cudaError_t launch_scan_kw(const QueryDesc* d_q, const QueryDesc& q, int sm_count, cudaStream_t st) {
  const size_t smem = size_t(q.n_stages) * (size_t(q.n_stage_plain) * TILE * 8 + size_t(q.n_stage_seeds) * NWARP * sizeof(Seed));
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_scan<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan<KW>, NT + 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = uint32_t(sm_count) * uint32_t(per_sm);
  if (grid > q.n_tiles) grid = q.n_tiles;
  k_scan<KW><<<grid, NT + 32, smem, st>>>(d_q);
  return cudaGetLastError();
}
int grid_for(size_t n) {
  int blocks = int((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  return blocks < 1 ? 1 : blocks;
}
}  // namespace

cudaError_t launch_table_init(const QueryDesc& q, cudaStream_t st) {
  k_table_init<<<grid_for(q.table_slots), 256, 0, st>>>(q);
  return cudaGetLastError();
}

cudaError_t launch_scan(const QueryDesc* d_q, const QueryDesc& q, int sm_count, cudaStream_t st) {
  if (q.n_tiles == 0) return cudaSuccess;
  if (q.table_mode == TM_DENSE || q.key_words <= 1) return launch_scan_kw<1>(d_q, q, sm_count, st);
  if (q.key_words == 2) return launch_scan_kw<2>(d_q, q, sm_count, st);
  return launch_scan_kw<kMaxKeyWords>(d_q, q, sm_count, st);
}

cudaError_t launch_rows(const QueryDesc* d_q, const QueryDesc& q, int sm_count, cudaStream_t st) {
  if (q.n_tiles == 0) return cudaSuccess;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rows, NT, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = uint32_t(sm_count) * uint32_t(per_sm);
  if (grid > q.n_tiles) grid = q.n_tiles;
  k_rows<<<grid, NT, 0, st>>>(d_q);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeDesc& f, cudaStream_t st) {
  k_finalize<<<grid_for(f.table_slots), 256, 0, st>>>(f);
  return cudaGetLastError();
}

cudaError_t launch_merge(const QueryDesc& q, const void* partial, cudaStream_t st) {
  k_merge<<<grid_for(q.table_slots), 256, 0, st>>>(q, static_cast<const uint8_t*>(partial));
  return cudaGetLastError();
}

cudaError_t launch_decode(const ChunkDesc& c, int32_t* out_i32, long long* out_i64, uint8_t* out_valid, int sm_count,
                          cudaStream_t st) {
  uint32_t n_chunks = (c.n_rows + kIndexRows - 1) / kIndexRows;
  if (n_chunks == 0) return cudaSuccess;
  uint32_t grid = (n_chunks + NWARP - 1) / NWARP;
  if (grid > uint32_t(sm_count) * 8) grid = uint32_t(sm_count) * 8;
  k_decode<<<grid, NT, 0, st>>>(c, n_chunks, out_i32, out_i64, out_valid);
  return cudaGetLastError();
}

}  // namespace fgpu
