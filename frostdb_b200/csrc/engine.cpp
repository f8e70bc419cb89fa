// libfrostgpu host engine: context, part registry uploads, plan compilation, execution, result
// export, and the extern "C" surface of include/frostgpu.h.
//
// There is deliberately no CPU execution path in this file: without a CUDA device fgpu_init fails
// with FGPU_ERR_NO_DEVICE and nothing else can run.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/frostgpu.h"
#include "arrow_export.h"
#include "device_types.h"
#include "kernels.h"
#include "part_store.h"

using namespace fgpu;

namespace {

thread_local std::string g_err;

int32_t fail(int32_t code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      cudaGetLastError();                                                                      \
      return fail(_e == cudaErrorMemoryAllocation ? FGPU_ERR_OOM : FGPU_ERR_CUDA,              \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                         \
    }                                                                                          \
  } while (0)

// Stream-ordered device allocation (cudaMallocAsync on the engine stream, pool kept warm): a query's
// scratch and a streamed part's columns cost microseconds to allocate instead of a cudaMalloc each.
struct DevBuf {
  void* p = nullptr;
  size_t n = 0;
  cudaStream_t st = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { reset(); }
  void reset() {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t bytes, cudaStream_t stream) {
    reset();
    st = stream;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(&p, bytes, stream);
    if (e == cudaSuccess) n = bytes;
    else p = nullptr;
    return e;
  }
};

struct ExprNode {
  int32_t kind = 0, op = 0, left = -1, right = -1;
  std::string name;
  int32_t lit_type = FGPU_SCALAR_NULL;
  int64_t lit_i = 0;
  double lit_f = 0;
  std::string lit_bytes;
  fgpu_match_fn match = nullptr;
  void* match_user = nullptr;
};

const char* op_string(int op) {  // logicalplan.Op.String(), expr.go:35-72
  switch (op) {
    case FGPU_OP_EQ: return "==";
    case FGPU_OP_NOT_EQ: return "!=";
    case FGPU_OP_LT: return "<";
    case FGPU_OP_LT_EQ: return "<=";
    case FGPU_OP_GT: return ">";
    case FGPU_OP_GT_EQ: return ">=";
    case FGPU_OP_REGEX_MATCH: return "=~";
    case FGPU_OP_REGEX_NOT_MATCH: return "!~";
    case FGPU_OP_AND: return "&&";
    case FGPU_OP_OR: return "||";
    case FGPU_OP_ADD: return "+";
    case FGPU_OP_SUB: return "-";
    case FGPU_OP_MUL: return "*";
    case FGPU_OP_DIV: return "/";
    case FGPU_OP_CONTAINS: return "contains";
    case FGPU_OP_NOT_CONTAINS: return "not contains";
    default: return "?";
  }
}

const char* agg_string(int f) {  // logicalplan.AggFunc.String(), expr.go:731-750
  switch (f) {
    case FGPU_AGG_SUM: return "sum";
    case FGPU_AGG_MIN: return "min";
    case FGPU_AGG_MAX: return "max";
    case FGPU_AGG_COUNT: return "count";
    case FGPU_AGG_AVG: return "avg";
    default: return "?";
  }
}

}  // namespace

// Page-locked host blocks, recycled: result images of cached plans and the columns of large results (which the
// Arrow consumer holds until it releases the array, possibly after the context is gone: the pool is shared-owned).
// Freeing page-locked memory synchronises the device, so blocks are only released when the pool itself dies.
struct PinnedPool {
  std::mutex mu;
  std::multimap<size_t, uint8_t*> free_blocks;
  size_t free_bytes = 0;
  ~PinnedPool() {
    for (auto& kv : free_blocks) cudaFreeHost(kv.second);
  }
  uint8_t* take(size_t bytes, size_t* cap) {
    size_t want = 4096;
    while (want < bytes) want <<= 1;
    if (bytes > (size_t(1) << 24)) want = (bytes + (size_t(1) << 22) - 1) & ~((size_t(1) << 22) - 1);  // large blocks: 4 MiB granules
    {
      std::lock_guard<std::mutex> lk(mu);
      auto it = free_blocks.lower_bound(want);
      if (it != free_blocks.end() && it->first <= want * 2) {
        uint8_t* p = it->second;
        *cap = it->first;
        free_bytes -= it->first;
        free_blocks.erase(it);
        return p;
      }
    }
    uint8_t* p = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&p), want, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    *cap = want;
    return p;
  }
  void give(uint8_t* p, size_t cap) {
    std::lock_guard<std::mutex> lk(mu);
    free_blocks.emplace(cap, p);
    free_bytes += cap;
  }
};

// Mailbox communicator of one rank (see comm.cu).
struct CommHandle {  // what travels between the ranks (FGPU_COMM_HANDLE_BYTES)
  uint32_t magic;
  int32_t device;
  uint64_t pid, ptr, total, slot_bytes;
  int32_t n, rank;
  cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(CommHandle) <= FGPU_COMM_HANDLE_BYTES, "handle fits its blob");
constexpr uint32_t kCommMagic = 0x46474d42u;  // "FGMB"
constexpr size_t kCommDataOff = 4096;        // flags live in front of the slots

struct Comm {
  int rank = -1, n = 0;
  uint64_t slot_bytes = 0, total = 0, seq = 0;
  uint8_t* mailbox = nullptr;  // [flags: 2 sets x kMaxRanks x {seq, bytes}] ... [2 sets x n slots]
  bool open = false;
  uint8_t* peer[kMaxRanks] = {nullptr};
  bool peer_ipc[kMaxRanks] = {false};
  size_t flag_off(uint64_t set, int r) const { return (size_t(set) * kMaxRanks + size_t(r)) * 16; }
  size_t data_off(uint64_t set, int r) const { return kCommDataOff + (size_t(set) * size_t(n) + size_t(r)) * slot_bytes; }
};

struct fgpu_ctx {
  Comm comm;
  int device = 0;
  int sm_count = 148;
  int tile_rows = kTileRows;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::mutex mu;
  std::map<std::string, Table> tables;
  std::vector<ColumnImage*> pending_uploads;  // staging to release after the next stream sync
  uint64_t epoch_counter = 0;                 // source of Table::epoch stamps
  std::map<std::string, std::shared_ptr<struct QueryPlan>> plans;  // by plan signature (bounded, see fgpu_query_prepare)
  // page-locked arena for column metadata on its way to the device: a pageable source would make every
  // cudaMemcpyAsync wait for the transfers queued before it.  Reset by release_staging() after a sync.
  uint8_t* arena = nullptr;
  size_t arena_cap = 0, arena_used = 0;
  void* arena_take(size_t n) {
    if (!arena) {
      size_t want = size_t(64) << 20;
      if (cudaHostAlloc(reinterpret_cast<void**>(&arena), want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        arena = nullptr;
        return nullptr;
      }
      arena_cap = want;
    }
    size_t off = (arena_used + 63) & ~size_t(63);
    if (off + n > arena_cap) return nullptr;  // caller falls back to the pageable source
    arena_used = off + n;
    return arena + off;
  }
  std::shared_ptr<PinnedPool> pinned = std::make_shared<PinnedPool>();
  uint8_t* pinned_take(size_t bytes, size_t* cap) { return pinned->take(bytes, cap); }
  void pinned_give(uint8_t* p, size_t cap) { pinned->give(p, cap); }
  // page-locked scratch for the per-query descriptor upload and the counters read-back (guarded by mu)
  uint8_t* scratch = nullptr;
  size_t scratch_bytes = 0;
  cudaError_t ensure_scratch(size_t bytes) {
    if (bytes <= scratch_bytes) return cudaSuccess;
    if (scratch) cudaFreeHost(scratch);
    scratch = nullptr;
    scratch_bytes = 0;
    size_t want = std::max<size_t>(bytes * 2, 1 << 20);
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&scratch), want, cudaHostAllocDefault);
    if (e == cudaSuccess) scratch_bytes = want;
    return e;
  }
};

// FROSTGPU_PROFILE=1: host-side phase times of every Execute on stderr (development aid).
extern "C" char** environ;
// The FROSTGPU_* development switches as one string.  Called per Execute: the scan of the environment is skipped
// while the environment block looks unchanged (same array, same entry pointers: setenv / putenv replace an entry).
const std::string& env_switches() {
  static const char* const kNames[] = {"FROSTGPU_NO_TILE", "FROSTGPU_NO_RUNS", "FROSTGPU_NO_PRUNE", "FROSTGPU_NO_FAST", "FROSTGPU_NO_FUSE",
                                       "FROSTGPU_NO_TAKE", "FROSTGPU_TA_TILE", "FROSTGPU_TA_STAGES", "FROSTGPU_TA_GLOBAL", "FROSTGPU_TA_CHUNK", "FROSTGPU_TA_CARRY",
                                       "FROSTGPU_VL", "FROSTGPU_RING", "FROSTGPU_RUNS_BR", "FROSTGPU_RUNS_RING", "FROSTGPU_RUNS_SPAN",
                                       "FROSTGPU_RUNS_V1", "FROSTGPU_RT_TILE", "FROSTGPU_RT_STAGES", "FROSTGPU_RT_SPAN", "FROSTGPU_RT_WARPS",
                                       "FROSTGPU_NO_EXEC_CACHE", "FROSTGPU_NO_PLAN_CACHE", "FROSTGPU_PROFILE"};
  static std::mutex mu;
  static std::string cached;
  static uint64_t cached_sig = 0;
  uint64_t sig = 0x9e3779b97f4a7c15ull ^ uint64_t(reinterpret_cast<uintptr_t>(environ));
  for (char** e = environ; e && *e; e++) sig = (sig ^ uint64_t(reinterpret_cast<uintptr_t>(*e))) * 0x100000001b3ull;
  std::lock_guard<std::mutex> lk(mu);
  if (sig != cached_sig || cached_sig == 0) {
    cached.clear();
    for (const char* n : kNames) {
      const char* v = getenv(n);
      if (v) { cached += n; cached += '='; cached += v; cached += ';'; }
    }
    cached_sig = sig;
  }
  return cached;
}
struct PhaseClock {
  bool on = env_switches().find("FROSTGPU_PROFILE") != std::string::npos;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  std::string line;
  void mark(const char* what) {
    if (!on) return;
    auto t = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof buf, " %s=%.0fus", what, std::chrono::duration<double, std::micro>(t - t0).count());
    line += buf;
    t0 = t;
  }
  void flush(const char* tag) {
    if (on && !line.empty()) fprintf(stderr, "[frostgpu %s]%s\n", tag, line.c_str());
    line.clear();
  }
};

// A prepared plan.  Prepared queries with the same plan text share ONE of these through the context's plan table,
// so that the shim's prepare-per-Execute still reuses the compiled plan and its device state.
struct QueryPlan {
  // the compiled plan of the last Execute, reused while the table (its parts, its dictionaries) and the read
  // transaction are the same: a prepared query re-executed on an unchanged table skips plan compilation
  mutable std::shared_ptr<void> plan_cache;  // Compiled
  mutable uint64_t cache_epoch = 0, cache_tx = 0;
  mutable std::string cache_env;  // the FROSTGPU_* development switches the plan was compiled under
  fgpu_ctx* ctx = nullptr;
  std::string table;
  int32_t kind = 0;
  std::vector<ExprNode> exprs;
  int32_t filter = -1;
  std::vector<int32_t> group_by;
  std::vector<fgpu_agg> aggs;

  std::string expr_name(int i) const {  // Expr.Name(), expr.go:181,326,568,623
    const ExprNode& e = exprs[size_t(i)];
    switch (e.kind) {
      case FGPU_EXPR_COLUMN:
      case FGPU_EXPR_DYNCOLUMN: return e.name;
      case FGPU_EXPR_LITERAL: {
        if (e.lit_type == FGPU_SCALAR_INT64) return std::to_string(e.lit_i);
        if (e.lit_type == FGPU_SCALAR_FLOAT64) {
          char buf[64];
          snprintf(buf, sizeof buf, "%g", e.lit_f);
          return buf;
        }
        if (e.lit_type == FGPU_SCALAR_STRING) return e.lit_bytes;
        return "null";
      }
      case FGPU_EXPR_BINARY: return expr_name(e.left) + " " + op_string(e.op) + " " + expr_name(e.right);
    }
    return "?";
  }
};

struct fgpu_query {
  fgpu_ctx* ctx = nullptr;
  std::shared_ptr<QueryPlan> plan;
};

struct KeyOut {
  std::string name;
  bool is_int64 = false;
  bool is_float = false;
  const GlobalDict* dict = nullptr;
  std::vector<std::string> dict_snapshot;  // values by id at execute time
};

struct fgpu_result {
  fgpu_ctx* ctx = nullptr;
  fgpu_stats stats{};
  // compiled state kept for partial/merge
  QueryDesc qd{};
  FinalizeDesc fd{};
  DevBuf table, aux, cnt;
  size_t table_bytes = 0;
  unsigned int* cnt_ptr = nullptr;  // group counter inside `aux` when the scan counted the result rows
  bool groups_known = false;  // the scan's own stream already counted the result rows
  unsigned int n_groups = 0;
  std::vector<KeyOut> keys;
  std::vector<std::string> agg_names;
  std::vector<uint8_t> agg_is_float;
  // finished records
  std::vector<std::vector<OwnedColumn>> records;
  std::vector<int64_t> record_rows;
  size_t next = 0;
  bool finalized = false;
  bool rows_plan = false;
  bool rows_no_nulls = false;  // rows plan produced by the take kernels: no projected value is NULL
  // collective Execute between its two halves: the partial table is pushed, the merge is still to come
  bool pending = false;
  void* pending_cached = nullptr;              // ExecCache: the plan-owned device state holds the table (else `table`)
  std::shared_ptr<void> plan_keep;             // keeps that plan alive
  uint64_t pending_seq = 0, pending_bytes = 0;
};

namespace {

// Uploads one column of one part if it is not resident yet: the host-assembled meta region with one
// copy, the PLAIN value regions straight from the source file (one async copy per contiguous extent),
// all on the engine stream so that kernels queued behind need no extra synchronisation.
int32_t ensure_resident(fgpu_ctx* ctx, Table* table, Part* part, const std::string& column, uint64_t* h2d_bytes,
                        double* build_us = nullptr, double* upload_us = nullptr) {
  auto t0 = std::chrono::steady_clock::now();
  build_column(kIndexRows, table, part, column, /*device_seeds=*/true);
  auto t1 = std::chrono::steady_clock::now();
  if (build_us) *build_us += std::chrono::duration<double, std::micro>(t1 - t0).count();
  struct Tail {
    double* out; std::chrono::steady_clock::time_point t;
    ~Tail() { if (out) *out += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t).count(); }
  } tail{upload_us, t1};
  ColumnImage& img = part->images[column];
  if (!img.error.empty()) return FGPU_OK;  // surfaces as an error only if a query projects the column
  if (img.resident) return FGPU_OK;
  void* dev = nullptr;
  CUDA_TRY(cudaMallocAsync(&dev, img.dev_bytes, ctx->stream));
  const void* meta_src = img.meta.data();
  if (void* pin = ctx->arena_take(img.meta.size())) {
    std::memcpy(pin, img.meta.data(), img.meta.size());
    meta_src = pin;
  }
  cudaError_t e = cudaMemcpyAsync(dev, meta_src, img.meta.size(), cudaMemcpyHostToDevice, ctx->stream);
  // PLAIN pages: neighbouring page payloads (separated by a page header) travel as ONE transfer of the
  // raw bytes into a staging buffer and are gathered into the dense value array on the device; a payload
  // without neighbours goes straight to its place.
  uint64_t copied = img.meta.size();
  {
    struct Span { const uint8_t* src; uint64_t len, stage_off; size_t first, n; };
    std::vector<Span> spans;
    constexpr uint64_t kGap = 4096;
    for (size_t i = 0; i < img.extents.size(); i++) {
      const Extent& x = img.extents[i];
      if (!spans.empty() && x.src >= spans.back().src + spans.back().len && uint64_t(x.src - (spans.back().src + spans.back().len)) <= kGap) {
        spans.back().len = uint64_t(x.src - spans.back().src) + x.len;
        spans.back().n++;
      } else {
        spans.push_back({x.src, x.len, 0, i, 1});
      }
    }
    uint64_t stage_bytes = 0;
    std::vector<PageCopy> copies;
    for (Span& sp : spans) {
      if (sp.n == 1) continue;
      sp.stage_off = stage_bytes;
      stage_bytes += ((sp.len + 15) & ~uint64_t(15)) + 16;
      for (size_t i = sp.first; i < sp.first + sp.n; i++)
        copies.push_back({sp.stage_off + uint64_t(img.extents[i].src - sp.src), img.extents[i].dst_off, img.extents[i].len});
    }
    DevBuf stage;
    const uint64_t table_off = stage_bytes;
    if (!copies.empty() && e == cudaSuccess) {
      e = stage.alloc(stage_bytes + copies.size() * sizeof(PageCopy), ctx->stream);
      if (e == cudaSuccess) {
        const void* tsrc = copies.data();
        if (void* pin = ctx->arena_take(copies.size() * sizeof(PageCopy))) {
          std::memcpy(pin, copies.data(), copies.size() * sizeof(PageCopy));
          tsrc = pin;
        }
        e = cudaMemcpyAsync(static_cast<uint8_t*>(stage.p) + table_off, tsrc, copies.size() * sizeof(PageCopy), cudaMemcpyHostToDevice, ctx->stream);
      }
    }
    for (const Span& sp : spans) {
      if (e != cudaSuccess) break;
      copied += sp.len;
      if (sp.n == 1) e = cudaMemcpyAsync(static_cast<uint8_t*>(dev) + img.extents[sp.first].dst_off, sp.src, sp.len, cudaMemcpyHostToDevice, ctx->stream);
      else e = cudaMemcpyAsync(static_cast<uint8_t*>(stage.p) + sp.stage_off, sp.src, sp.len, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess && !copies.empty())
      e = launch_gather_pages(stage.p, dev, reinterpret_cast<const PageCopy*>(static_cast<uint8_t*>(stage.p) + table_off), uint32_t(copies.size()), ctx->stream);
    // `stage` is released in stream order when it goes out of scope
  }
  if (e == cudaSuccess) e = launch_make_seeds(dev, img.seed_jobs_off, img.n_seed_jobs, img.max_seed_chunks, ctx->stream);
  if (e != cudaSuccess) {
    cudaFreeAsync(dev, ctx->stream);
    cudaGetLastError();
    return fail(FGPU_ERR_CUDA, std::string("column upload: ") + cudaGetErrorString(e));
  }
  img.dev = dev;
  img.resident = true;
  patch_column_pointers(part, column, static_cast<const uint8_t*>(dev));
  if (h2d_bytes) *h2d_bytes += copied;
  ctx->pending_uploads.push_back(&img);
  return FGPU_OK;
}

int bit_width_u32(uint64_t max_value) {
  int w = 0;
  while (w < 64 && (max_value >> w) != 0) w++;
  return w;
}

// Host side of several columns of one part in parallel (page walk, run directories, dictionary interning): the
// columns are independent — every column has its own dictionary, image and chunk records, all created here
// before the workers start so that no container is resized concurrently.
void prebuild_columns(Table* table, Part* part, const std::vector<std::string>& columns) {
  std::vector<const std::string*> todo;
  for (const std::string& c : columns) {
    ColumnImage& img = part->images[c];  // creates the entry
    if (img.built) continue;
    for (const SchemaLeaf& l : part->pf.leaves)
      if (l.name == c && l.phys == PT_BYTE_ARRAY) (void)table->dicts[c];
    todo.push_back(&c);
  }
  if (todo.size() < 2) return;  // a single column is built by its caller
  unsigned hw = std::thread::hardware_concurrency();
  const size_t n_threads = std::min<size_t>(todo.size(), std::min<size_t>(hw ? hw : 4, 32));  // (a part has ~20 columns: one wave)
  std::atomic<size_t> next{0};
  std::vector<std::thread> workers;
  for (size_t t = 0; t < n_threads; t++)
    workers.emplace_back([&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= todo.size()) break;
        build_column(kIndexRows, table, part, *todo[i], /*device_seeds=*/true);
      }
    });
  for (std::thread& w : workers) w.join();
}

// Flat code arrays of one dictionary column of a part (every row group, one allocation), derived on the device
// from the resident hybrid image the first time the tile-aggregate kernel needs the column (k_flatten).
int32_t ensure_flat(fgpu_ctx* ctx, Part* part, const std::string& column) {
  ColumnImage& img = part->images[column];
  if (img.flat_dev) return FGPU_OK;
  if (!img.resident) return fail(FGPU_ERR_INVALID, "flat codes requested for a column that is not resident: " + column);
  std::vector<FlatJob> jobs;
  std::vector<ChunkHost*> chs;
  std::vector<uint64_t> offs;
  uint64_t bytes = 0;
  uint32_t blocks = 0;
  for (RowGroupHost& rg : part->rgs) {
    auto it = rg.cols.find(column);
    if (it == rg.cols.end()) continue;
    ChunkHost& ch = it->second;
    if (!ch.error.empty() || ch.desc.kind != CK_DICT_STR) continue;
    uint32_t max_gid = 0;
    for (uint32_t g : ch.lut_host) max_gid = std::max(max_gid, g);
    FlatJob j{};
    j.chunk = ch.desc;
    j.bias = ch.desc.has_nulls ? 0u : 1u;
    const int bits = bit_width_u32(uint64_t(max_gid) + 1u - j.bias);
    j.w = bits <= 8 ? 8u : (bits <= 16 ? 16u : 32u);  // byte-aligned codes: one shared-memory load per row and column
    j.first_block = blocks;
    const uint32_t nb = (rg.n_rows + uint32_t(kIndexRows) - 1) / uint32_t(kIndexRows);
    blocks += nb;
    offs.push_back(bytes);
    bytes += (uint64_t(nb) * (kIndexRows / 8) * j.w + 64 + 127) & ~uint64_t(127);
    jobs.push_back(j);
    chs.push_back(&ch);
  }
  if (jobs.empty()) return FGPU_OK;
  void* dev = nullptr;
  CUDA_TRY(cudaMallocAsync(&dev, bytes + jobs.size() * sizeof(FlatJob), ctx->stream));
  for (size_t i = 0; i < jobs.size(); i++) jobs[i].out = static_cast<uint8_t*>(dev) + offs[i];
  const void* src = jobs.data();
  if (void* pin = ctx->arena_take(jobs.size() * sizeof(FlatJob))) {
    std::memcpy(pin, jobs.data(), jobs.size() * sizeof(FlatJob));
    src = pin;
  }
  FlatJob* d_jobs = reinterpret_cast<FlatJob*>(static_cast<uint8_t*>(dev) + bytes);
  cudaError_t e = cudaMemcpyAsync(d_jobs, src, jobs.size() * sizeof(FlatJob), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = launch_flatten(d_jobs, uint32_t(jobs.size()), blocks, ctx->sm_count, ctx->stream);
  if (e != cudaSuccess) {
    cudaFreeAsync(dev, ctx->stream);
    cudaGetLastError();
    return fail(FGPU_ERR_CUDA, std::string("flat codes: ") + cudaGetErrorString(e));
  }
  img.flat_dev = dev;
  for (size_t i = 0; i < jobs.size(); i++) {
    chs[i]->flat = jobs[i].out;
    chs[i]->flat_w = uint8_t(jobs[i].w);
    chs[i]->flat_bias = uint8_t(jobs[i].bias);
  }
  return FGPU_OK;
}

// After the stream has been synchronised the host staging of uploaded columns can go.
void release_staging(fgpu_ctx* ctx) {
  ctx->arena_used = 0;
  for (ColumnImage* img : ctx->pending_uploads) {
    std::vector<uint8_t>().swap(img->meta);
    std::vector<Extent>().swap(img->extents);
  }
  ctx->pending_uploads.clear();
}

void free_part(fgpu_ctx* ctx, Part* p) {
  for (auto& kv : p->images) {
    if (kv.second.dev) cudaFreeAsync(kv.second.dev, ctx->stream);
    if (kv.second.flat_dev) cudaFreeAsync(kv.second.flat_dev, ctx->stream);
  }
}

struct VisibleRG {
  Part* part;
  RowGroupHost* rg;
  uint64_t skip_slots = 0;               // bit s: slot s is not read in this row group (its leaf is decided by statistics)
  uint8_t leaf_mode[kMaxLeaves] = {0};   // LeafMode per leaf decided from the chunk statistics (LM_EVAL: undecided)
};


// One predicate leaf on the host.
struct LeafHost {
  std::string column;
  int slot = -1;  // -1: column absent from every visible row group
  int op = 0;
  const ExprNode* lit = nullptr;
  const ExprNode* bin = nullptr;
  size_t lut_off = size_t(-1);  // dictionary leaves: offset of the per-global-id result bytes in the LUT blob
  // numeric leaves, canonical form (see LeafDesc)
  bool numeric = false, never = false, neg = false, cmp_float = false, null_literal = false;
  int64_t lo_i = 0, hi_i = 0;
  double lo_f = 0, hi_f = 0;
  uint8_t missing_mode = LM_ALL;  // behaviour on row groups that lack the column
};

bool bytes_contains(const std::string& hay, const std::string& needle) {
  if (needle.empty()) return true;
  return hay.find(needle) != std::string::npos;
}

// Missing-column rules, binaryscalarexpr.go:47-73 and regexpfilter.go:23-33.  Returns LM_ALL / LM_NONE.
uint8_t missing_column_mode(const LeafHost& l) {
  const ExprNode& lit = *l.lit;
  switch (l.op) {
    case FGPU_OP_EQ:
      if (lit.lit_type == FGPU_SCALAR_STRING && !lit.lit_bytes.empty()) return LM_NONE;
      return LM_ALL;
    case FGPU_OP_NOT_EQ:
      if (lit.lit_type == FGPU_SCALAR_NULL) return LM_NONE;
      return LM_ALL;
    case FGPU_OP_LT: case FGPU_OP_LT_EQ: case FGPU_OP_GT: case FGPU_OP_GT_EQ:
      return LM_NONE;
    case FGPU_OP_REGEX_MATCH:
    case FGPU_OP_REGEX_NOT_MATCH: {
      bool empty_match = l.bin->match ? l.bin->match(l.bin->match_user, reinterpret_cast<const uint8_t*>(""), 0) != 0 : false;
      bool not_match = l.op == FGPU_OP_REGEX_NOT_MATCH;
      return ((not_match && !empty_match) || (!not_match && empty_match)) ? LM_ALL : LM_NONE;
    }
    default:
      return LM_ALL;  // contains / not contains fall through to "all rows" (:71)
  }
}

// Result of a dictionary leaf for one dictionary value (non-NULL row).
bool dict_leaf_value(const LeafHost& l, const std::string& v) {
  const ExprNode& lit = *l.lit;
  switch (l.op) {
    case FGPU_OP_EQ: return lit.lit_type != FGPU_SCALAR_NULL && v == lit.lit_bytes;
    case FGPU_OP_NOT_EQ: return lit.lit_type == FGPU_SCALAR_NULL ? true : v != lit.lit_bytes;
    case FGPU_OP_CONTAINS:
      return lit.lit_type == FGPU_SCALAR_NULL ? true : bytes_contains(v, lit.lit_bytes);
    case FGPU_OP_NOT_CONTAINS:
      return lit.lit_type == FGPU_SCALAR_NULL ? true : !bytes_contains(v, lit.lit_bytes);
    case FGPU_OP_REGEX_MATCH:
      return l.bin->match(l.bin->match_user, reinterpret_cast<const uint8_t*>(v.data()), v.size()) != 0;
    case FGPU_OP_REGEX_NOT_MATCH:
      return l.bin->match(l.bin->match_user, reinterpret_cast<const uint8_t*>(v.data()), v.size()) == 0;
  }
  return false;
}

// "column op literal" on an int64 column as an inclusive range, optionally negated (see LeafDesc).
void canonical_int_range(int op, int64_t v, int64_t* lo, int64_t* hi, bool* neg) {
  const int64_t kMin = std::numeric_limits<int64_t>::min(), kMax = std::numeric_limits<int64_t>::max();
  *lo = kMax; *hi = kMin;  // empty
  *neg = false;
  switch (op) {
    case FGPU_OP_EQ: *lo = *hi = v; break;
    case FGPU_OP_NOT_EQ: *lo = *hi = v; *neg = true; break;
    case FGPU_OP_LT: if (v != kMin) { *lo = kMin; *hi = v - 1; } break;
    case FGPU_OP_LT_EQ: *lo = kMin; *hi = v; break;
    case FGPU_OP_GT: if (v != kMax) { *lo = v + 1; *hi = kMax; } break;
    default: *lo = v; *hi = kMax; break;
  }
}

// What the chunk statistics [mn, mx] (bounds of the non-null values) decide about an int64 range leaf for a
// whole row group: LM_NONE no row can pass, LM_ALL every row passes (needs a chunk without NULLs), LM_EVAL
// undecided.  The row-group filter of the reference (expr/binaryscalarexpr.go:84-190) answers the first question.
uint8_t stats_leaf_mode(int64_t lo, int64_t hi, bool neg, int64_t mn, int64_t mx, bool no_nulls) {
  const bool disjoint = mx < lo || mn > hi;
  const bool inside = mn >= lo && mx <= hi;
  if (!neg) return disjoint ? LM_NONE : ((inside && no_nulls) ? LM_ALL : LM_EVAL);
  return inside ? LM_NONE : ((disjoint && no_nulls) ? LM_ALL : LM_EVAL);
}

// The decision compile() takes for an int64 range leaf on one chunk, also behind the host-only test hook
// fgpu_rowgroup_leaf_mode: an all-NULL chunk has no usable bounds (nothing is decided from statistics).
uint8_t chunk_leaf_mode(int64_t lo, int64_t hi, bool neg, bool has_bounds, int64_t mn, int64_t mx, int64_t null_count, int64_t num_values) {
  if (!has_bounds || (null_count >= 0 && null_count == num_values)) return LM_EVAL;
  return stats_leaf_mode(lo, hi, neg, mn, mx, null_count == 0);
}

// Device and page-locked state of a compiled plan that every Execute of a prepared query reuses while the plan is
// valid (same table epoch, same read transaction): the descriptor block is uploaded once, the aggregate table and
// its counters are one persistent allocation, and the table comes back to page-locked memory with ONE copy and ONE
// stream synchronisation per Execute; dense tables of this size are compacted into the result record on the host.
struct ExecCache {
  DevBuf aux, table, out;         // descriptors; [aggregate table | counters 128 B]; compacted result image
  DenseOut dout{};
  size_t out_bytes = 0;
  std::vector<std::unique_ptr<OwnedColumn>> dict_template;  // per key: the whole dictionary, ready to copy (small dictionaries)
  fgpu_ctx* ctx = nullptr;
  uint8_t* pinned = nullptr;      // host image of the compacted result (from the context's page-locked pool)
  size_t pinned_cap = 0;
  size_t table_bytes = 0;
  bool ready = false;
  QueryDesc qd{};
  const QueryDesc* qdesc_dev = nullptr;
  bool has_runs = false, has_ta = false, runs_v1 = false;
  RunsDesc rd{};
  int runs_nl = 0, runs_nk = 0, runs_na = 0;
  TileAggDesc td{};
  fgpu_stats stats{};             // everything about the scan that does not change between Executes
  std::vector<KeyOut> keys;       // with the dictionary snapshots
  std::vector<uint32_t> dense_radix;
  std::vector<std::string> agg_names;
  std::vector<uint8_t> agg_is_float;
  std::shared_ptr<PinnedPool> pool;
  ~ExecCache() {
    if (pinned && pool) pool->give(pinned, pinned_cap);
  }
};

struct Compiled {
  std::unique_ptr<ExecCache> exec;
  std::vector<VisibleRG> rgs;
  std::vector<std::string> slot_names;
  std::vector<uint8_t> slot_types;
  std::vector<uint8_t> slot_needed;  // 0: only named by Count(), whose result does not depend on the values
  std::vector<LeafHost> leaves;
  std::vector<uint8_t> filter_prog;
  std::vector<KeyOut> keys;
  std::vector<int> key_slots;
  QueryDesc qd{};
  std::vector<uint32_t> dense_radix;
  uint64_t total_rows = 0;
  uint64_t group_bound = 0;
  uint64_t h2d_bytes = 0;  // column uploads this query triggered
  uint32_t pruned_row_groups = 0;
  uint64_t touched_rows = ~uint64_t(0);  // rows of the row groups that survive pruning (~0: nothing was pruned)
  bool runs_shape = false;  // dense keys, conjunction of numeric leaves, plain aggregate inputs (any number of them)
};

int32_t compile_filter(const QueryPlan& q, int node, Compiled* c, std::map<std::string, int>& slot_of) {
  const ExprNode& e = q.exprs[size_t(node)];
  if (e.kind != FGPU_EXPR_BINARY) return fail(FGPU_ERR_UNSUPPORTED, "unsupported boolean expression");
  if (e.op == FGPU_OP_AND || e.op == FGPU_OP_OR) {
    int32_t rc = compile_filter(q, e.left, c, slot_of);
    if (rc) return rc;
    rc = compile_filter(q, e.right, c, slot_of);
    if (rc) return rc;
    c->filter_prog.push_back(e.op == FGPU_OP_AND ? 0x80 : 0x81);
    return FGPU_OK;
  }
  switch (e.op) {
    case FGPU_OP_EQ: case FGPU_OP_NOT_EQ: case FGPU_OP_LT: case FGPU_OP_LT_EQ: case FGPU_OP_GT: case FGPU_OP_GT_EQ:
    case FGPU_OP_REGEX_MATCH: case FGPU_OP_REGEX_NOT_MATCH: case FGPU_OP_CONTAINS: case FGPU_OP_NOT_CONTAINS:
      break;
    default:
      return fail(FGPU_ERR_UNSUPPORTED, std::string("unsupported filter operator ") + op_string(e.op));
  }
  const ExprNode& l = q.exprs[size_t(e.left)];
  const ExprNode& r = q.exprs[size_t(e.right)];
  if (l.kind != FGPU_EXPR_COLUMN) return fail(FGPU_ERR_INVALID, "left side of binary expression must be a column");
  if (r.kind != FGPU_EXPR_LITERAL) return fail(FGPU_ERR_INVALID, "right side of binary expression must be a literal");
  if ((e.op == FGPU_OP_REGEX_MATCH || e.op == FGPU_OP_REGEX_NOT_MATCH) && !e.match)
    return fail(FGPU_ERR_INVALID, "regex leaf without a matcher");
  if (c->leaves.size() >= size_t(kMaxLeaves)) return fail(FGPU_ERR_UNSUPPORTED, "too many predicate leaves");
  LeafHost lh;
  lh.column = l.name;
  lh.op = e.op;
  lh.lit = &r;
  lh.bin = &e;
  auto it = slot_of.find(l.name);
  lh.slot = (it == slot_of.end()) ? -1 : it->second;
  c->filter_prog.push_back(uint8_t(c->leaves.size()));
  c->leaves.push_back(std::move(lh));
  return FGPU_OK;
}

int32_t compile_agg_expr(const QueryPlan& q, int node, std::map<std::string, int>& slot_of, const Compiled& c,
                         std::vector<ProgOp>* prog, bool* any_float, bool* any_int_col) {
  const ExprNode& e = q.exprs[size_t(node)];
  if (e.kind == FGPU_EXPR_COLUMN) {
    auto it = slot_of.find(e.name);
    if (it == slot_of.end()) return fail(FGPU_ERR_NOT_FOUND, "aggregate field(s) not found [\"" + e.name + "\"], aggregations are not possible without it");
    uint8_t t = c.slot_types[size_t(it->second)];
    if (t == ST_DICT) return fail(FGPU_ERR_UNSUPPORTED, "aggregation over a non-numeric column: " + e.name);
    if (t == ST_F64) *any_float = true; else *any_int_col = true;
    ProgOp op{};
    op.op = PO_LOAD;
    op.slot = uint8_t(it->second);
    prog->push_back(op);
    return FGPU_OK;
  }
  if (e.kind == FGPU_EXPR_LITERAL) {
    ProgOp op{};
    op.op = PO_CONST;
    if (e.lit_type == FGPU_SCALAR_INT64) op.imm = e.lit_i;
    else if (e.lit_type == FGPU_SCALAR_FLOAT64) { std::memcpy(&op.imm, &e.lit_f, 8); op.slot = 1; *any_float = true; }
    else return fail(FGPU_ERR_UNSUPPORTED, "non-numeric literal in arithmetic expression");
    prog->push_back(op);
    return FGPU_OK;
  }
  if (e.kind == FGPU_EXPR_BINARY && e.op >= FGPU_OP_ADD && e.op <= FGPU_OP_DIV) {
    int32_t rc = compile_agg_expr(q, e.left, slot_of, c, prog, any_float, any_int_col);
    if (rc) return rc;
    rc = compile_agg_expr(q, e.right, slot_of, c, prog, any_float, any_int_col);
    if (rc) return rc;
    ProgOp op{};
    op.op = uint8_t(PO_ADD + (e.op - FGPU_OP_ADD));
    prog->push_back(op);
    return FGPU_OK;
  }
  return fail(FGPU_ERR_UNSUPPORTED, "unsupported aggregate expression");
}

void collect_columns(const QueryPlan& q, int node, std::vector<std::string>* out) {
  if (node < 0) return;
  const ExprNode& e = q.exprs[size_t(node)];
  if (e.kind == FGPU_EXPR_COLUMN) out->push_back(e.name);
  if (e.kind == FGPU_EXPR_BINARY) {
    collect_columns(q, e.left, out);
    collect_columns(q, e.right, out);
  }
}

// The row-group filter LSM.Scan applies to the row groups of Parquet parts before they reach the plan
// (index/lsm.go:437; expr/filter.go:208-268 builds And / Or / BinaryScalarExpr, everything else is AlwaysTrue;
// expr/binaryscalarexpr.go:42-190 answers "may this chunk hold a matching value" from the null count and the
// bounds).  Its rules for a column that is MISSING from the row group (:47-73) differ from the physical plan's
// (physicalplan/binaryscalarexpr.go:47-73): `missing == 5`, `missing != 5`, `missing != ""` and every ordering
// comparison drop the row group here, so its rows never reach the plan.
int stat_compare(const ChunkHost& ch, bool use_max, const ExprNode& lit, bool* ok) {
  *ok = false;
  if (ch.phys == PT_INT64 && lit.lit_type == FGPU_SCALAR_INT64 && ch.has_minmax) {
    const int64_t v = use_max ? ch.max_bits : ch.min_bits;
    *ok = true;
    return v < lit.lit_i ? -1 : (v > lit.lit_i ? 1 : 0);
  }
  if (ch.phys == PT_DOUBLE && lit.lit_type == FGPU_SCALAR_FLOAT64 && ch.has_minmax) {
    double v;
    std::memcpy(&v, use_max ? &ch.max_bits : &ch.min_bits, 8);
    *ok = true;
    return v < lit.lit_f ? -1 : (v > lit.lit_f ? 1 : 0);
  }
  if (ch.phys == PT_BYTE_ARRAY && lit.lit_type == FGPU_SCALAR_STRING && ch.has_minmax_str) {
    const std::string& v = use_max ? ch.max_str : ch.min_str;
    *ok = true;
    const int c = v.compare(lit.lit_bytes);
    return c < 0 ? -1 : (c > 0 ? 1 : 0);
  }
  return 0;
}

// Split-block bloom filter of the chunk against a non-null literal (expr/binaryscalarexpr.go:104-118: parquet-go hashes
// the PLAIN encoding of the value with XXH64): false = the value is definitely not in the chunk.  Chunks without a
// filter, and literals whose type does not match the column's, answer true.
bool bloom_may_contain(const ChunkHost& ch, const ExprNode& lit) {
  if (ch.bloom.empty()) return true;
  uint64_t h;
  if (ch.phys == PT_INT64 && lit.lit_type == FGPU_SCALAR_INT64) {
    uint8_t b[8];
    std::memcpy(b, &lit.lit_i, 8);
    h = xxhash64(b, 8, 0);
  } else if (ch.phys == PT_DOUBLE && lit.lit_type == FGPU_SCALAR_FLOAT64) {
    uint8_t b[8];
    std::memcpy(b, &lit.lit_f, 8);
    h = xxhash64(b, 8, 0);
  } else if (ch.phys == PT_BYTE_ARRAY && lit.lit_type == FGPU_SCALAR_STRING) {
    h = xxhash64(reinterpret_cast<const uint8_t*>(lit.lit_bytes.data()), lit.lit_bytes.size(), 0);
  } else {
    return true;
  }
  return sbbf_check(ch.bloom.data(), uint32_t(ch.bloom.size()), h);
}

bool rg_may_match(const QueryPlan& q, int node, const RowGroupHost& rg) {
  if (node < 0) return true;
  const ExprNode& e = q.exprs[size_t(node)];
  if (e.kind != FGPU_EXPR_BINARY) return true;
  if (e.op == FGPU_OP_AND) return rg_may_match(q, e.left, rg) && rg_may_match(q, e.right, rg);
  if (e.op == FGPU_OP_OR) return rg_may_match(q, e.left, rg) || rg_may_match(q, e.right, rg);
  if (e.op < FGPU_OP_EQ || e.op > FGPU_OP_GT_EQ) return true;  // AlwaysTrueFilter
  const ExprNode& l = q.exprs[size_t(e.left)];
  const ExprNode& lit = q.exprs[size_t(e.right)];
  if (l.kind != FGPU_EXPR_COLUMN || lit.kind != FGPU_EXPR_LITERAL) return true;
  auto it = rg.cols.find(l.name);
  if (it == rg.cols.end()) {  // :47-73
    if (lit.lit_type == FGPU_SCALAR_NULL) {
      if (e.op == FGPU_OP_EQ) return true;
      if (e.op == FGPU_OP_NOT_EQ) return false;
    }
    if (lit.lit_type == FGPU_SCALAR_STRING) {
      if (e.op == FGPU_OP_EQ && lit.lit_bytes.empty()) return true;
      if (e.op == FGPU_OP_NOT_EQ && !lit.lit_bytes.empty()) return true;
    }
    return false;
  }
  const ChunkHost& ch = it->second;
  const int64_t nulls = ch.null_count;  // -1: not recorded
  const bool full_of_nulls = nulls >= 0 && uint64_t(nulls) == rg.n_rows;
  bool ok;
  if (e.op == FGPU_OP_EQ) {
    if (lit.lit_type == FGPU_SCALAR_NULL) return nulls != 0;
    if (full_of_nulls) return false;
    // the reference asks the bloom filter when the chunk has one and the bounds otherwise (:104-118); both are true
    // negatives, so both are asked here
    if (!bloom_may_contain(ch, lit)) return false;
    const int cmax = stat_compare(ch, true, lit, &ok);
    if (!ok) return true;
    const int cmin = stat_compare(ch, false, lit, &ok);
    if (!ok) return true;
    return cmax >= 0 && cmin <= 0;
  }
  if (lit.lit_type == FGPU_SCALAR_NULL) return true;
  if (full_of_nulls) return false;
  switch (e.op) {
    case FGPU_OP_LT_EQ: { const int c = stat_compare(ch, false, lit, &ok); return !ok || c <= 0; }
    case FGPU_OP_LT: { const int c = stat_compare(ch, false, lit, &ok); return !ok || c < 0; }
    case FGPU_OP_GT: { const int c = stat_compare(ch, true, lit, &ok); return !ok || c > 0; }
    case FGPU_OP_GT_EQ: { const int c = stat_compare(ch, true, lit, &ok); return !ok || c >= 0; }
    default: return true;  // != : left to the execution engine
  }
}

// Resolves the plan against the visible row groups and fills everything of QueryDesc that does
// not need device memory.
int32_t compile(fgpu_ctx* ctx, const QueryPlan& q, uint64_t tx, Compiled* c) {
  Table& table = ctx->tables[q.table];  // a table without parts scans nothing (empty LSM)
  for (auto& p : table.parts) {
    if (p->tx > tx) continue;  // index/lsm.go:416
    for (auto& rg : p->rgs) {
      c->rgs.push_back({p.get(), &rg});
      c->total_rows += rg.n_rows;
    }
  }
  if (c->rgs.empty()) {  // nothing to scan: no records, exactly like an Iterator over an empty LSM
    c->qd.table_mode = TM_DENSE;
    c->qd.key_words = 1;
    c->qd.table_slots = 1;
    c->qd.tile_rows = ctx->tile_rows;
    return FGPU_OK;
  }
  // every column name present in a visible row group, in name order, with its type
  std::map<std::string, uint8_t> present;
  for (const VisibleRG& v : c->rgs) {
    for (auto& kv : v.rg->cols) {
      uint8_t t = kv.second.phys == PT_BYTE_ARRAY ? ST_DICT : (kv.second.phys == PT_DOUBLE ? ST_F64 : ST_I64);
      if (kv.second.phys != PT_BYTE_ARRAY && kv.second.phys != PT_DOUBLE && kv.second.phys != PT_INT64) t = 0xff;
      auto it = present.find(kv.first);
      if (it == present.end()) present.emplace(kv.first, t);
      else if (it->second != t) return fail(FGPU_ERR_UNSUPPORTED, "column " + kv.first + " changes type between parts");
    }
  }
  // string columns whose dictionaries were preloaded (identically on every rank of a multi-GPU run) count
  // as present even when no local row group holds them: the key set, and with it the shape of the partial
  // aggregate table, must not depend on which parts a rank happens to own
  for (auto& kv : table.dicts)
    if (kv.second.preloaded && !present.count(kv.first)) present.emplace(kv.first, uint8_t(ST_DICT));
  std::map<std::string, int> slot_of;
  auto want = [&](const std::string& name) -> int32_t {
    if (slot_of.count(name)) return FGPU_OK;
    auto it = present.find(name);
    if (it == present.end()) return FGPU_OK;  // absent everywhere: handled by the callers
    if (it->second == 0xff) return fail(FGPU_ERR_UNSUPPORTED, "column " + name + " has a type the GPU engine does not read");
    if (c->slot_names.size() >= size_t(kMaxSlots)) return fail(FGPU_ERR_UNSUPPORTED, "query touches too many columns");
    slot_of[name] = int(c->slot_names.size());
    c->slot_names.push_back(name);
    c->slot_types.push_back(it->second);
    return FGPU_OK;
  };
  // group-by / distinct columns; a computed key (arithmetic over int64 columns, the sqlparse pre-projection
  // `(timestamp / 1000) * 1000 as bucket`) is named after its expression and carries the expression's node
  std::vector<std::string> key_names;
  std::map<std::string, int32_t> key_expr;  // computed keys: name -> expression node
  for (int32_t g : q.group_by) {
    const ExprNode& e = q.exprs[size_t(g)];
    if (e.kind == FGPU_EXPR_COLUMN) {
      if (present.count(e.name)) key_names.push_back(e.name);
    } else if (e.kind == FGPU_EXPR_DYNCOLUMN) {
      std::string prefix = e.name + ".";  // DynamicColumn.MatchColumn, expr.go:564
      for (auto& kv : present)
        if (kv.first.compare(0, prefix.size(), prefix) == 0) key_names.push_back(kv.first);
    } else if (e.kind == FGPU_EXPR_BINARY && e.op >= FGPU_OP_ADD && e.op <= FGPU_OP_DIV && q.kind == FGPU_PLAN_AGGREGATE) {
      const std::string name = q.expr_name(g);
      if (present.count(name)) return fail(FGPU_ERR_UNSUPPORTED, "computed group key shadows a column: " + name);
      key_names.push_back(name);
      key_expr[name] = g;
    } else {
      return fail(FGPU_ERR_UNSUPPORTED, "this group-by expression is not supported on the GPU path");
    }
  }
  {
    std::vector<std::string> uniq;
    for (auto& n : key_names)
      if (std::find(uniq.begin(), uniq.end(), n) == uniq.end()) uniq.push_back(n);
    key_names.swap(uniq);
  }
  if (key_names.size() > size_t(kMaxKeys)) return fail(FGPU_ERR_UNSUPPORTED, "too many group-by columns");
  for (auto& n : key_names) {
    if (key_expr.count(n)) continue;
    int32_t rc = want(n);
    if (rc) return rc;
  }
  // filter + aggregate + computed-key input columns
  std::vector<std::string> cols;
  for (auto& kv : key_expr) collect_columns(q, kv.second, &cols);
  collect_columns(q, q.filter, &cols);
  for (const fgpu_agg& a : q.aggs) collect_columns(q, a.expr, &cols);
  for (auto& n : cols) {
    int32_t rc = want(n);
    if (rc) return rc;
  }
  QueryDesc& qd = c->qd;
  qd.n_slots = int32_t(c->slot_names.size());
  for (int s = 0; s < qd.n_slots; s++) {
    qd.slot_type[s] = c->slot_types[size_t(s)];
  }
  // Count(col) counts rows, NULLs included (aggregate.go:929-950): its input column is never read,
  // so it is neither uploaded nor staged unless something else in the query needs it.
  {
    std::vector<std::string> need = key_names;
    for (auto& kv : key_expr) collect_columns(q, kv.second, &need);
    collect_columns(q, q.filter, &need);
    for (const fgpu_agg& a : q.aggs)
      if (a.func != FGPU_AGG_COUNT) collect_columns(q, a.expr, &need);
    c->slot_needed.assign(c->slot_names.size(), q.kind == FGPU_PLAN_AGGREGATE ? 0 : 1);
    for (auto& n : need) {
      auto it = slot_of.find(n);
      if (it != slot_of.end()) c->slot_needed[size_t(it->second)] = 1;
    }
  }
  // filter
  if (q.filter >= 0) {
    int32_t rc = compile_filter(q, q.filter, c, slot_of);
    if (rc) return rc;
    if (c->filter_prog.size() > size_t(kMaxFilterProg)) return fail(FGPU_ERR_UNSUPPORTED, "filter expression too large");
    // stack depth check (the kernel keeps the stack in 32 bits)
    int depth = 0, maxd = 0;
    for (uint8_t op : c->filter_prog) {
      if (op < 0x80) depth++; else depth--;
      maxd = std::max(maxd, depth);
    }
    if (maxd > 30) return fail(FGPU_ERR_UNSUPPORTED, "filter expression too deep");
    qd.n_filter_prog = int32_t(c->filter_prog.size());
    std::memcpy(qd.filter_prog, c->filter_prog.data(), c->filter_prog.size());
    // pure conjunctions / disjunctions of leaves reduce to one mask test per row
    bool all_and = true, all_or = true;
    uint32_t mask = 0;
    for (uint8_t op : c->filter_prog) {
      if (op < 0x80) mask |= 1u << op;
      else if (op == 0x80) all_or = false;
      else all_and = false;
    }
    qd.filter_mask = mask;
    qd.filter_kind = all_and ? FK_AND : (all_or ? FK_OR : FK_PROGRAM);
  }
  for (size_t l = 0; l < c->leaves.size(); l++) {
    LeafHost& lh = c->leaves[l];
    lh.missing_mode = missing_column_mode(lh);
    if (lh.slot < 0) continue;
    uint8_t st = c->slot_types[size_t(lh.slot)];
    const ExprNode& lit = *lh.lit;
    if (st == ST_DICT) {
      switch (lh.op) {
        case FGPU_OP_EQ: case FGPU_OP_NOT_EQ: case FGPU_OP_CONTAINS: case FGPU_OP_NOT_CONTAINS:
          if (lit.lit_type != FGPU_SCALAR_NULL && lit.lit_type != FGPU_SCALAR_STRING)
            return fail(FGPU_ERR_UNSUPPORTED, "dictionary column compared with a non-string literal");
          break;
        case FGPU_OP_REGEX_MATCH: case FGPU_OP_REGEX_NOT_MATCH:
          break;
        default:  // binaryscalarexpr.go:100-108
          return fail(FGPU_ERR_INVALID, std::string("unsupported operator: ") + op_string(lh.op));
      }
    } else {
      if (lh.op < FGPU_OP_EQ || lh.op > FGPU_OP_GT_EQ)
        return fail(FGPU_ERR_UNSUPPORTED, std::string("operator ") + op_string(lh.op) + " on a numeric column");
      if (lit.lit_type == FGPU_SCALAR_STRING) return fail(FGPU_ERR_UNSUPPORTED, "numeric column compared with a string literal");
      lh.numeric = true;
      lh.null_literal = lit.lit_type == FGPU_SCALAR_NULL;  // arrow compute against NULL: nothing selected (:143-150)
      lh.cmp_float = (st == ST_F64) || lit.lit_type == FGPU_SCALAR_FLOAT64;
      const double kInf = std::numeric_limits<double>::infinity();
      if (lh.cmp_float) {
        const double v = lit.lit_type == FGPU_SCALAR_FLOAT64 ? lit.lit_f : double(lit.lit_i);
        lh.lo_f = kInf; lh.hi_f = -kInf;  // empty
        if (!std::isnan(v)) {
          switch (lh.op) {
            case FGPU_OP_EQ: lh.lo_f = lh.hi_f = v; break;
            case FGPU_OP_NOT_EQ: lh.lo_f = lh.hi_f = v; lh.neg = true; break;
            case FGPU_OP_LT: lh.lo_f = -kInf; lh.hi_f = std::nextafter(v, -kInf); break;
            case FGPU_OP_LT_EQ: lh.lo_f = -kInf; lh.hi_f = v; break;
            case FGPU_OP_GT: lh.lo_f = std::nextafter(v, kInf); lh.hi_f = kInf; break;
            default: lh.lo_f = v; lh.hi_f = kInf; break;
          }
          if (lh.op == FGPU_OP_LT && v == -kInf) { lh.lo_f = kInf; lh.hi_f = -kInf; }
          if (lh.op == FGPU_OP_GT && v == kInf) { lh.lo_f = kInf; lh.hi_f = -kInf; }
        } else if (lh.op == FGPU_OP_NOT_EQ) {
          lh.neg = true;  // x != NaN holds for every non-NULL x
        }
      } else {
        bool neg = false;
        canonical_int_range(lh.op, lit.lit_i, &lh.lo_i, &lh.hi_i, &neg);
        lh.neg = neg;
      }
    }
  }
  // Conjunction: comparisons on the same numeric column intersect into one range leaf
  // (timestamp >= a AND timestamp < b  ->  timestamp in [a, b-1]); a missing column keeps the
  // strictest of the merged leaves' missing-column behaviours.
  if (qd.filter_kind == FK_AND && !getenv("FROSTGPU_NO_FUSE")) {
    std::vector<LeafHost> fused;
    for (LeafHost& lh : c->leaves) {
      bool merged = false;
      if (lh.numeric && !lh.neg && !lh.null_literal && lh.slot >= 0) {
        for (LeafHost& f : fused) {
          if (!f.numeric || f.neg || f.null_literal || f.slot != lh.slot || f.cmp_float != lh.cmp_float) continue;
          if (f.cmp_float) { f.lo_f = std::max(f.lo_f, lh.lo_f); f.hi_f = std::min(f.hi_f, lh.hi_f); }
          else { f.lo_i = std::max(f.lo_i, lh.lo_i); f.hi_i = std::min(f.hi_i, lh.hi_i); }
          if (lh.missing_mode == LM_NONE) f.missing_mode = LM_NONE;
          merged = true;
          break;
        }
      }
      if (!merged) fused.push_back(lh);
    }
    if (fused.size() != c->leaves.size()) {
      c->leaves.swap(fused);
      c->filter_prog.clear();
      for (size_t l = 0; l < c->leaves.size(); l++) {
        c->filter_prog.push_back(uint8_t(l));
        if (l > 0) c->filter_prog.push_back(0x80);
      }
      qd.n_filter_prog = int32_t(c->filter_prog.size());
      std::memset(qd.filter_prog, 0, sizeof qd.filter_prog);
      std::memcpy(qd.filter_prog, c->filter_prog.data(), c->filter_prog.size());
      qd.filter_mask = (c->leaves.size() >= 32) ? 0xffffffffu : ((1u << c->leaves.size()) - 1u);
    }
  }
  qd.n_leaves = int32_t(c->leaves.size());
  for (int l = 0; l < qd.n_leaves; l++) {
    const LeafHost& lh = c->leaves[size_t(l)];
    LeafDesc& ld = qd.leaves[l];
    ld = LeafDesc{};
    ld.slot = uint8_t(lh.slot < 0 ? 0xff : lh.slot);
    ld.op = uint8_t(lh.op);
    if (lh.slot < 0) continue;
    qd.slot_used_by_leaf[lh.slot] = 1;
    ld.cmp_float = lh.cmp_float;
    ld.neg = lh.neg;
    ld.lo_i = lh.lo_i; ld.hi_i = lh.hi_i;
    ld.lo_f = lh.lo_f; ld.hi_f = lh.hi_f;
  }
  // ---- row-group pruning by chunk statistics ----------------------------------------------------------
  // LSM.Scan asks the filter's TrueNegativeFilter whether a row group may hold matching rows before it
  // hands the row group to the plan (index/lsm.go:401-454; expr/binaryscalarexpr.go:41-190 compares the
  // literal with the chunk's min/max).  Same decision here, from the footer statistics, before anything
  // is uploaded: an int64 range leaf whose chunk bounds miss the range selects nothing in that row group
  // (under a conjunction the row group is dropped: never uploaded, never scanned); bounds inside the
  // range of a chunk without NULLs select every row, and the leaf's column is then not read there.
  {
    const bool conj = qd.n_filter_prog > 0 && qd.filter_kind == FK_AND;
    const bool may_skip = q.kind != FGPU_PLAN_FILTER;
    std::vector<uint8_t> base(c->slot_names.size(), may_skip ? 0 : 1);  // needed regardless of the filter
    if (may_skip) {
      std::vector<std::string> need = key_names;
      for (auto& kv : key_expr) collect_columns(q, kv.second, &need);
      for (const fgpu_agg& a : q.aggs)
        if (a.func != FGPU_AGG_COUNT) collect_columns(q, a.expr, &need);
      for (auto& n : need) {
        auto it = slot_of.find(n);
        if (it != slot_of.end()) base[size_t(it->second)] = 1;
      }
    }
    std::vector<VisibleRG> kept;
    const bool prune_on = !getenv("FROSTGPU_NO_PRUNE");
    // AggFuncPushDown (logicalplan/optimize.go:166-193) + MaxAgg (expr/filter.go:156-207): a global Max(column)
    // directly above the scan, without filter, turns into a row-group filter that memoises the largest chunk
    // maximum seen so far, in scan order, and only lets row groups through that exceed it.
    bool max_agg = prune_on && q.kind == FGPU_PLAN_AGGREGATE && q.group_by.empty() && q.aggs.size() == 1 && q.aggs[0].func == FGPU_AGG_MAX &&
                   q.filter < 0 && q.exprs[size_t(q.aggs[0].expr)].kind == FGPU_EXPR_COLUMN;
    const std::string max_col = max_agg ? q.exprs[size_t(q.aggs[0].expr)].name : std::string();
    bool have_max = false;
    int64_t max_i = 0;
    double max_f = 0;
    for (VisibleRG& v : c->rgs) {
      if (max_agg && !v.part->arrow) {
        auto it = v.rg->cols.find(max_col);
        if (it == v.rg->cols.end()) continue;  // no such field in this row group: nothing to contribute
        const ChunkHost& ch = it->second;
        const bool all_null = ch.null_count >= 0 && uint64_t(ch.null_count) == v.rg->n_rows;
        if (all_null) continue;
        if (ch.has_minmax && (ch.phys == PT_INT64 || ch.phys == PT_DOUBLE)) {
          bool greater;
          if (ch.phys == PT_INT64) {
            greater = !have_max || ch.max_bits > max_i;
            if (greater) max_i = ch.max_bits;
          } else {
            double mx;
            std::memcpy(&mx, &ch.max_bits, 8);
            greater = !have_max || mx > max_f;
            if (greater) max_f = mx;
          }
          if (!greater) continue;
          have_max = true;
        }
      }
      // Parquet parts: the reference's own row-group filter first (L0 Arrow records are never filtered, lsm.go:420-427)
      if (prune_on && !v.part->arrow && q.filter >= 0 && !rg_may_match(q, q.filter, *v.rg)) continue;
      bool drop = false;
      for (size_t l = 0; l < c->leaves.size() && prune_on; l++) {
        const LeafHost& lh = c->leaves[l];
        uint8_t mode = LM_EVAL;
        auto it = lh.slot < 0 ? v.rg->cols.end() : v.rg->cols.find(lh.column);
        if (it == v.rg->cols.end()) {
          mode = lh.missing_mode;  // the physical plan's rule for the rows of a row group that got this far (physicalplan/binaryscalarexpr.go:47-73)
        } else if (lh.numeric && !lh.cmp_float && !lh.null_literal && c->slot_types[size_t(lh.slot)] == ST_I64 && it->second.has_minmax) {
          mode = chunk_leaf_mode(lh.lo_i, lh.hi_i, lh.neg, true, it->second.min_bits, it->second.max_bits, it->second.null_count, int64_t(v.rg->n_rows));
        }
        else if (it != v.rg->cols.end() && c->slot_types[size_t(lh.slot)] == ST_DICT && lh.op == FGPU_OP_EQ &&
                 lh.lit->lit_type == FGPU_SCALAR_STRING && it->second.has_minmax_str) {
          // string equality against the chunk's bounding strings (binaryscalarexpr.go:98-104; bounds may be
          // truncated by the writer, they still bound): outside them no value can be equal
          const std::string& lit = lh.lit->lit_bytes;
          if (lit < it->second.min_str || lit > it->second.max_str) mode = LM_NONE;
        }
        // equality with a value the chunk's bloom filter rules out: the leaf is false on every row of the row group
        // (also under a disjunction, where the row group itself stays)
        if (mode == LM_EVAL && it != v.rg->cols.end() && lh.op == FGPU_OP_EQ && !lh.null_literal && !v.part->arrow &&
            !bloom_may_contain(it->second, *lh.lit))
          mode = LM_NONE;
        v.leaf_mode[l] = mode;
        if (conj && mode == LM_NONE) drop = true;
      }
      if (drop) continue;
      for (size_t s = 0; s < c->slot_names.size(); s++) {
        bool needed = base[s] != 0;
        for (size_t l = 0; l < c->leaves.size() && !needed; l++)
          if (c->leaves[l].slot == int(s) && v.leaf_mode[l] == LM_EVAL) needed = true;
        if (!needed || !c->slot_needed[s]) v.skip_slots |= 1ull << s;
      }
      kept.push_back(v);
    }
    c->pruned_row_groups = uint32_t(c->rgs.size() - kept.size());
    c->rgs.swap(kept);
    c->touched_rows = 0;
    for (const VisibleRG& v : c->rgs) c->touched_rows += v.rg->n_rows;
    // Every row group ruled out: nothing is scanned, but the plan keeps its full shape (keys, aggregates, table
    // slots): a rank whose time range misses the filter must still produce a partial table its peers can merge.
  }
  // Lazy residency: the columns this query projects are built and uploaded now (parts put with
  // FGPU_PUT_BORROW_PINNED upload nothing until a query needs it; optimize.go:36-73 physical projection).
  {
    PhaseClock pc;
    double build_us = 0, upload_us = 0;
    // numeric columns first: their bytes are the bulk of the transfer, which then runs while the host
    // assembles the dictionary columns' run directories
    for (int pass = 0; pass < 2; pass++) {
    Part* last = nullptr;
    for (const VisibleRG& v : c->rgs) {
      if (v.part == last) continue;
      last = v.part;
      for (size_t si = 0; si < c->slot_names.size(); si++) {
        const std::string& name = c->slot_names[si];
        if ((c->slot_types[si] == ST_DICT) != (pass == 1)) continue;
        bool read = false;
        for (const VisibleRG& w : c->rgs)
          if (w.part == v.part && !((w.skip_slots >> si) & 1)) { read = true; break; }
        if (!read) continue;
        if (std::find(v.part->columns.begin(), v.part->columns.end(), name) == v.part->columns.end()) continue;
        int32_t rc = ensure_resident(ctx, &table, v.part, name, &c->h2d_bytes, pc.on ? &build_us : nullptr, pc.on ? &upload_us : nullptr);
        if (rc) return rc;
        const ColumnImage& img = v.part->images[name];
        if (!img.error.empty()) return fail(FGPU_ERR_UNSUPPORTED, "column " + name + ": " + img.error);
      }
    }
    }
    if (pc.on && c->h2d_bytes) {
      char buf[128];
      snprintf(buf, sizeof buf, " build=%.0fus enqueue=%.0fus bytes=%llu", build_us, upload_us, (unsigned long long)c->h2d_bytes);
      pc.line = buf;
      pc.flush("residency");
    }
  }
  // shared-memory ring of the scan kernel: stage every numeric slot's PLAIN slice (up to kMaxStagePlain)
  // and the chunk seeds of every hybrid stream (up to kMaxStageSeeds); the rest is read from HBM directly
  for (int s = 0; s < kMaxSlots; s++) {
    qd.slot_plain_stage[s] = -1;
    qd.slot_seed_stage[s][0] = qd.slot_seed_stage[s][1] = -1;
  }
  for (int s = 0; s < qd.n_slots; s++) {
    bool any_plain = false, any_vals = false, any_def = false;
    if (!c->slot_needed[size_t(s)]) continue;
    for (const VisibleRG& v : c->rgs) {
      auto it = v.rg->cols.find(c->slot_names[size_t(s)]);
      if (it == v.rg->cols.end() || ((v.skip_slots >> s) & 1)) continue;
      const ChunkDesc& d = it->second.desc;
      if (d.kind == CK_PLAIN64 && !d.has_nulls) any_plain = true;
      if (d.kind == CK_DICT_STR || d.kind == CK_DICT64) any_vals = true;
      if (d.has_nulls) any_def = true;
    }
    if (any_plain && qd.n_stage_plain < kMaxStagePlain) {
      qd.slot_plain_stage[s] = int8_t(qd.n_stage_plain);
      qd.stage_plain_slot[qd.n_stage_plain++] = uint8_t(s);
    }
    if (any_vals && qd.n_stage_seeds < kMaxStageSeeds) {
      qd.slot_seed_stage[s][0] = int8_t(qd.n_stage_seeds);
      qd.stage_seed_slot[qd.n_stage_seeds] = uint8_t(s);
      qd.stage_seed_is_def[qd.n_stage_seeds++] = 0;
    }
    if (any_def && qd.n_stage_seeds < kMaxStageSeeds) {
      qd.slot_seed_stage[s][1] = int8_t(qd.n_stage_seeds);
      qd.stage_seed_slot[qd.n_stage_seeds] = uint8_t(s);
      qd.stage_seed_is_def[qd.n_stage_seeds++] = 1;
    }
  }
  // aggregates
  if (q.aggs.size() > size_t(kMaxAggs)) return fail(FGPU_ERR_UNSUPPORTED, "too many aggregates");
  qd.n_aggs = int32_t(q.aggs.size());
  std::vector<ProgOp> prog;
  for (size_t a = 0; a < q.aggs.size(); a++) {
    AggDesc& ad = qd.aggs[a];
    int f = q.aggs[a].func;
    if (f != FGPU_AGG_SUM && f != FGPU_AGG_MIN && f != FGPU_AGG_MAX && f != FGPU_AGG_COUNT)
      return fail(FGPU_ERR_UNSUPPORTED, std::string("aggregate function not supported on the GPU path: ") + agg_string(f));
    ad.func = uint8_t(f);
    bool any_float = false, any_int_col = false;
    size_t off = prog.size();
    int32_t rc = compile_agg_expr(q, q.aggs[a].expr, slot_of, *c, &prog, &any_float, &any_int_col);
    if (rc) return rc;
    if (any_float && any_int_col) return fail(FGPU_ERR_UNSUPPORTED, "arithmetic mixes int64 and float64 columns");
    if (any_float) {  // integer literals take part as doubles
      for (size_t p = off; p < prog.size(); p++)
        if (prog[p].op == PO_CONST && prog[p].slot == 0) {
          double d = double(prog[p].imm);
          std::memcpy(&prog[p].imm, &d, 8);
        }
    }
    ad.is_float = any_float;
    ad.prog_off = uint8_t(off);
    ad.prog_len = uint8_t(prog.size() - off);
    // the aggregated column must exist in every record (aggregate.go:367-380)
    std::vector<std::string> acols;
    collect_columns(q, q.aggs[a].expr, &acols);
    for (const VisibleRG& v : c->rgs)
      for (auto& n : acols)
        if (!v.rg->cols.count(n))
          return fail(FGPU_ERR_NOT_FOUND, "aggregate field(s) not found [\"" + n + "\"], aggregations are not possible without it");
  }
  if (prog.size() > size_t(kMaxProg)) return fail(FGPU_ERR_UNSUPPORTED, "aggregate expressions too large");
  for (size_t p = 0; p < prog.size(); p++) qd.prog[p] = prog[p];
  // the kernel evaluates aggregate expressions on an operand stack of three shared-memory vectors
  for (size_t a = 0; a < q.aggs.size(); a++) {
    int depth = 0, maxd = 0;
    for (int p = qd.aggs[a].prog_off; p < qd.aggs[a].prog_off + qd.aggs[a].prog_len; p++) {
      depth += (prog[size_t(p)].op == PO_LOAD || prog[size_t(p)].op == PO_CONST) ? 1 : -1;
      maxd = std::max(maxd, depth);
    }
    if (maxd > 3) return fail(FGPU_ERR_UNSUPPORTED, "aggregate expression nests too deeply for the GPU path");
  }
  qd.tile_rows = ctx->tile_rows;
  qd.n_rg = int32_t(c->rgs.size());
  if (q.kind == FGPU_PLAN_FILTER) {
    // Filter -> Projection(columns): the resolved columns are the output, no table
    if (key_names.size() > size_t(kMaxOut)) return fail(FGPU_ERR_UNSUPPORTED, "too many projected columns");
    qd.n_out = int32_t(key_names.size());
    for (size_t k = 0; k < key_names.size(); k++) {
      int slot = slot_of.at(key_names[k]);
      qd.out_slot[k] = uint8_t(slot);
      KeyOut ko;
      ko.name = key_names[k];
      ko.is_int64 = c->slot_types[size_t(slot)] != ST_DICT;
      ko.is_float = c->slot_types[size_t(slot)] == ST_F64;
      if (!ko.is_int64) ko.dict = &table.dicts[key_names[k]];  // (a column whose every row group was pruned was never built: empty dictionary)
      c->keys.push_back(std::move(ko));
      c->key_slots.push_back(slot);
    }
    return FGPU_OK;
  }
  // keys and table shape
  qd.n_keys = int32_t(key_names.size());
  bool all_dict = true;
  long double product = 1;
  c->dense_radix.assign(key_names.size(), 0);
  for (size_t k = 0; k < key_names.size(); k++) {
    const auto kx = key_expr.find(key_names[k]);
    if (kx != key_expr.end()) {  // computed int64 key: its program sits behind the aggregates' in qd.prog
      bool any_float = false, any_int_col = false;
      const size_t off = prog.size();
      int32_t rc = compile_agg_expr(q, kx->second, slot_of, *c, &prog, &any_float, &any_int_col);
      if (rc) return rc;
      if (any_float) return fail(FGPU_ERR_UNSUPPORTED, "float64 group-by expressions are not supported");
      if (prog.size() > size_t(kMaxProg) || prog.size() - off > 255) return fail(FGPU_ERR_UNSUPPORTED, "group-by expression too large");
      int depth = 0, maxd = 0;
      for (size_t p = off; p < prog.size(); p++) {
        depth += (prog[p].op == PO_LOAD || prog[p].op == PO_CONST) ? 1 : -1;
        maxd = std::max(maxd, depth);
      }
      if (maxd > 3) return fail(FGPU_ERR_UNSUPPORTED, "group-by expression nests too deeply for the GPU path");
      for (size_t p = off; p < prog.size(); p++) qd.prog[p] = prog[p];
      std::vector<std::string> kcols;
      collect_columns(q, kx->second, &kcols);
      for (const VisibleRG& v : c->rgs)
        for (auto& n : kcols)
          if (!v.rg->cols.count(n)) return fail(FGPU_ERR_NOT_FOUND, "group-by expression column not found: " + n);
      KeyOut ko;
      ko.name = key_names[k];
      ko.is_int64 = true;
      all_dict = false;
      product *= 1e18L;
      c->keys.push_back(std::move(ko));
      c->key_slots.push_back(-1);
      qd.keys[k].slot = 0xff;
      qd.keys[k].is_int64 = 1;
      qd.keys[k].prog_off = uint8_t(off);
      qd.keys[k].prog_len = uint8_t(prog.size() - off);
      continue;
    }
    int slot = slot_of.at(key_names[k]);
    KeyOut ko;
    ko.name = key_names[k];
    ko.is_int64 = c->slot_types[size_t(slot)] != ST_DICT;
    if (c->slot_types[size_t(slot)] == ST_F64) return fail(FGPU_ERR_UNSUPPORTED, "float64 group-by columns are not supported");
    if (ko.is_int64) {
      all_dict = false;
      product *= 1e18L;
    } else {
      ko.dict = &table.dicts[key_names[k]];
      uint32_t card = ko.dict->cardinality();
      c->dense_radix[k] = card + 1;
      product *= (long double)(card + 1);
    }
    c->keys.push_back(std::move(ko));
    c->key_slots.push_back(slot);
    qd.keys[k].slot = uint8_t(slot);
    qd.keys[k].is_int64 = c->keys.back().is_int64;
  }
  long double bound = std::min<long double>(product, (long double)std::max<uint64_t>(c->total_rows, 1));
  c->group_bound = uint64_t(bound);
  const uint64_t kDenseMax = 1ull << 22;
  if (all_dict && product <= (long double)kDenseMax) {
    qd.table_mode = TM_DENSE;
    qd.key_words = 1;
    uint64_t stride = 1;
    for (size_t k = key_names.size(); k-- > 0;) {
      qd.keys[k].dense_stride = uint32_t(stride);
      stride *= c->dense_radix[k];
    }
    qd.table_slots = uint32_t(std::max<uint64_t>(stride, 1));
  } else {
    qd.table_mode = TM_HASH;
    int word = 0, used = 0;
    for (size_t k = 0; k < key_names.size(); k++) {
      if (c->keys[k].is_int64) continue;
      int bits = std::max(1, bit_width_u32(c->dense_radix[k] - 1));
      if (used + bits > 64) { word++; used = 0; }
      qd.keys[k].word = uint8_t(word);
      qd.keys[k].shift = uint8_t(used);
      qd.keys[k].bits = uint32_t(bits);
      used += bits;
    }
    int words = (used > 0 || word > 0) ? word + 1 : 0;
    for (size_t k = 0; k < key_names.size(); k++) {
      if (!c->keys[k].is_int64) continue;
      qd.keys[k].word = uint8_t(words++);
      qd.keys[k].shift = 0;
      qd.keys[k].bits = 64;
    }
    if (words > kMaxKeyWords) return fail(FGPU_ERR_UNSUPPORTED, "group key wider than the packed-key limit");
    qd.key_words = std::max(words, 1);
    uint64_t cap = 1024;
    while (cap < 2 * c->group_bound && cap < (1ull << 28)) cap <<= 1;
    qd.table_slots = uint32_t(cap);
  }
  // ---- scan kernel: vector length, ring depth and the per-warp shared-memory layout ----------------
  {
    auto envi = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
    int vl = envi("FROSTGPU_VL", 256);
    if (vl != 128 && vl != 256 && vl != 512 && vl != 1024) vl = 512;
    int ring = envi("FROSTGPU_RING", 3);
    if (ring < 2) ring = 2;
    if (ring > 4) ring = 4;
    bool any_expr = false;
    for (int k = 0; k < qd.n_keys; k++)
      if (qd.keys[k].prog_len) any_expr = true;
    for (int a = 0; a < qd.n_aggs; a++)
      if (qd.aggs[a].func != FGPU_AGG_COUNT && !(qd.aggs[a].prog_len == 1 && qd.prog[qd.aggs[a].prog_off].op == PO_LOAD)) any_expr = true;
    auto r128 = [](size_t x) { return (x + 127) & ~size_t(127); };
    for (;;) {
      size_t off = r128(size_t(ring) * 8);
      size_t slot = r128(size_t(qd.n_stage_plain) * vl * 8 + size_t(qd.n_stage_seeds) * sizeof(Seed));
      if (slot == 0) slot = 128;
      qd.slot_bytes = uint32_t(slot);
      qd.wr_ring = uint32_t(off); off += size_t(ring) * slot;
      qd.wr_act = uint32_t(off); off = r128(off + size_t(vl / 32) * 4);
      qd.wr_leaf = uint32_t(off); off = r128(off + ((qd.n_filter_prog > 0 && qd.filter_kind != FK_AND) ? size_t(vl) * 4 : 0));
      qd.wr_slot = uint32_t(off); off = r128(off + size_t(vl) * 4);
      qd.wr_keyw = uint32_t(off); off = r128(off + (qd.table_mode == TM_HASH ? size_t(qd.key_words) * vl * 8 : 0));
      qd.wr_tmp1 = uint32_t(off); off = r128(off + (any_expr ? size_t(vl) * 8 : 0));
      qd.wr_tmp2 = uint32_t(off); off = r128(off + (any_expr ? size_t(vl) * 16 : 0));
      qd.wr_acc = uint32_t(off); off = r128(off + size_t(qd.n_aggs) * 32 * 8 + 16 + 32 * 4 + 32 * 4 + (qd.table_mode == TM_HASH ? size_t(qd.key_words) * 32 * 8 : 0));
      qd.wr_cdesc = uint32_t(off); off = r128(off + size_t(std::max(qd.n_slots, 1)) * sizeof(ChunkDesc));
      qd.wr_clrt = uint32_t(off); off = r128(off + size_t(std::max(qd.n_leaves, 1)) * sizeof(LeafRt));
      qd.wr_fplan = uint32_t(off); off = r128(off + 512);
      qd.wr_iplan = uint32_t(off); off = r128(off + 8 + 16 * size_t(kMaxStagePlain + kMaxStageSeeds));
      qd.wr_bytes = uint32_t(off);
      if (size_t(qd.wr_bytes) * (kVecThreads / 32) <= 200 * 1024 || vl == 128) break;
      vl /= 2;  // many staged columns / wide keys: shorter vectors
    }
    qd.vl = vl;
    qd.n_ring = ring;
    // fused fast path: conjunction of numeric leaves, dense dictionary keys, simple aggregates
    bool fast = qd.table_mode == TM_DENSE && (qd.n_filter_prog == 0 || qd.filter_kind == FK_AND) && qd.n_leaves <= 2 && qd.n_keys <= 3 &&
                !getenv("FROSTGPU_NO_FAST");
    for (int l = 0; l < qd.n_leaves && fast; l++) {
      const LeafHost& lh = c->leaves[size_t(l)];
      if (lh.slot >= 0 && c->slot_types[size_t(lh.slot)] == ST_DICT) fast = false;
      if (lh.slot >= 0 && lh.null_literal) fast = false;
    }
    int stored = 0;
    for (int a = 0; a < qd.n_aggs && fast; a++) {
      if (qd.aggs[a].func == FGPU_AGG_COUNT) continue;
      stored++;
      if (!(qd.aggs[a].prog_len == 1 && qd.prog[qd.aggs[a].prog_off].op == PO_LOAD)) fast = false;
    }
    {  // the sorted-run kernel: up to kRunsAggs reducers, and dictionary leaves evaluated once per run
      bool rs = qd.table_mode == TM_DENSE && (qd.n_filter_prog == 0 || qd.filter_kind == FK_AND) && qd.n_keys <= kRunsKeys &&
                qd.n_leaves <= kRunsLeaves + kRunsPreds && !getenv("FROSTGPU_NO_FAST");
      for (int l = 0; l < qd.n_leaves && rs; l++)
        if (c->leaves[size_t(l)].slot >= 0 && c->leaves[size_t(l)].numeric && c->leaves[size_t(l)].null_literal) rs = false;
      for (int a = 0; a < qd.n_aggs && rs; a++)
        if (qd.aggs[a].func != FGPU_AGG_COUNT && !(qd.aggs[a].prog_len == 1 && qd.prog[qd.aggs[a].prog_off].op == PO_LOAD)) rs = false;
      c->runs_shape = rs;
    }
    if (stored > 2) fast = false;
    qd.fast_ok = fast ? 1 : 0;
  }
  return FGPU_OK;
}

size_t table_layout(const QueryDesc& qd, size_t* off_aggs, size_t* off_tags, size_t* off_keys) {
  size_t S = qd.table_slots;
  int n_stored = 0;
  for (int a = 0; a < qd.n_aggs; a++)
    if (qd.aggs[a].func != FGPU_AGG_COUNT) n_stored++;
  *off_aggs = S * 8;
  *off_tags = S * 8 * size_t(1 + n_stored);
  size_t tags = (qd.table_mode == TM_HASH) ? ((S * 4 + 7) & ~size_t(7)) : 0;
  *off_keys = *off_tags + tags;
  size_t keys = (qd.table_mode == TM_HASH) ? S * 8 * size_t(qd.key_words) : 0;
  return *off_keys + keys;
}

void bind_table(QueryDesc* qd, uint8_t* base) {
  size_t off_aggs, off_tags, off_keys;
  table_layout(*qd, &off_aggs, &off_tags, &off_keys);
  size_t S = qd->table_slots;
  qd->t_rows = reinterpret_cast<unsigned long long*>(base);
  int pos = 0;
  for (int a = 0; a < kMaxAggs; a++) qd->t_agg[a] = nullptr;
  for (int a = 0; a < qd->n_aggs; a++) {
    if (qd->aggs[a].func == FGPU_AGG_COUNT) continue;
    qd->t_agg[a] = reinterpret_cast<long long*>(base + off_aggs + size_t(pos) * S * 8);
    pos++;
  }
  qd->t_tag = qd->table_mode == TM_HASH ? reinterpret_cast<uint32_t*>(base + off_tags) : nullptr;
  qd->t_keys = qd->table_mode == TM_HASH ? reinterpret_cast<unsigned long long*>(base + off_keys) : nullptr;
}

// Runs init + scan.  On success the result owns the device table.
// The development / test switches that change how a plan is compiled or launched: a cached plan is only reused
// under the switches it was built with.
inline bool env_has(const std::string& env, const char* name) { return env.find(name) != std::string::npos; }

int32_t finalize_dense_host(fgpu_ctx* ctx, fgpu_result* res, ExecCache& x);
void bind_table(QueryDesc* qd, uint8_t* base);

// First half of the exchange, enqueued behind the scan: this rank's partial table goes into slot [rank] of every
// rank's mailbox, then the flags are raised (comm.cu).
int32_t comm_push(fgpu_ctx* ctx, const void* table, size_t table_bytes, uint64_t* out_seq, uint64_t* out_bytes) {
  Comm& cm = ctx->comm;
  if (!cm.open) return fail(FGPU_ERR_INVALID, "collective Execute without an open communicator (fgpu_comm_open)");
  const size_t bytes = (table_bytes + 15) & ~size_t(15);
  if (bytes > cm.slot_bytes) return fail(FGPU_ERR_UNSUPPORTED, "partial table (" + std::to_string(bytes) + " bytes) larger than the exchange slot");
  const uint64_t seq = ++cm.seq, set = seq & 1u;
  CommPush p{};
  p.done = reinterpret_cast<unsigned int*>(cm.mailbox + 2048);  // (inside the zeroed header region, behind the flags)
  p.src = static_cast<const uint8_t*>(table);
  p.bytes = bytes;
  p.seq = seq;
  p.n = cm.n;
  for (int r = 0; r < cm.n; r++) {
    p.dst[r] = cm.peer[r] + cm.data_off(set, cm.rank);
    p.flag[r] = reinterpret_cast<unsigned long long*>(cm.peer[r] + cm.flag_off(set, cm.rank));
  }
  CUDA_TRY(launch_comm_push(p, ctx->sm_count, ctx->stream));
  *out_seq = seq;
  *out_bytes = bytes;
  return FGPU_OK;
}

// Second half: wait for every rank's flag; dense tables are then folded from the n mailbox slots into `table`
// itself (it holds the FINAL table afterwards).  Hash tables: the caller merges the slots with k_merge.
int32_t comm_wait_merge(fgpu_ctx* ctx, const QueryDesc& qd, void* table, uint64_t seq, uint64_t bytes) {
  Comm& cm = ctx->comm;
  const uint64_t set = seq & 1u;
  CommWait w{};
  w.flags = reinterpret_cast<const unsigned long long*>(cm.mailbox + cm.flag_off(set, 0));
  w.seq = seq;
  w.bytes = bytes;
  const char* to = getenv("FROSTGPU_COMM_TIMEOUT_MS");
  w.timeout_ns = uint64_t(to ? std::max(1, atoi(to)) : 10000) * 1000000ull;
  w.counters = qd.counters;
  w.n = cm.n;
  CUDA_TRY(launch_comm_wait(w, ctx->stream));
  if (qd.table_mode == TM_DENSE) {
    CommMerge m{};
    m.n = cm.n;
    for (int r = 0; r < cm.n; r++) m.src[r] = cm.mailbox + cm.data_off(set, r);
    QueryDesc q2 = qd;
    bind_table(&q2, static_cast<uint8_t*>(table));
    CUDA_TRY(launch_merge_dense(q2, m, ctx->stream));
  }
  return FGPU_OK;
}

int32_t comm_check(unsigned long long code) {
  if (code == 1) return fail(FGPU_ERR_CUDA, "collective Execute: a peer did not deliver its partial table within the timeout");
  if (code == 2) return fail(FGPU_ERR_UNSUPPORTED, "collective Execute: partial table shapes differ between ranks (dictionaries not preloaded identically?)");
  return FGPU_OK;
}

// Re-issues the launches of a cached plan (see ExecCache): table init, scans, one copy back, one synchronisation.
// Tail of a cached Execute: compact the (final) dense table into result columns on the device and bring them back.
// The result image is written by the kernel straight into a page-locked block of the context's pool (host memory the
// device addresses directly): no copy operation, and the block later travels to the Arrow consumer as it is.
int32_t cached_tail(fgpu_ctx* ctx, ExecCache& x) {
  cudaStream_t s = ctx->stream;
  if (!x.pinned) {
    x.pinned = ctx->pinned_take(x.out_bytes, &x.pinned_cap);
    if (!x.pinned) return fail(FGPU_ERR_OOM, "page-locked memory for the result image");
  }
  // k_finalize_dense writes the block over PCIe itself: 22 us of kernel time for 400 KB against 4 + 9 for kernel + copy
  // engine, but the Execute is 8 us shorter end to end without the copy operation behind the kernel (measured)
  x.dout.out = x.pinned;
  CUDA_TRY(launch_finalize_dense(x.dout, s));
  return FGPU_OK;
}

// Collective tail of a cached Execute: ONE launch waits for the peers' flags, folds the n partial tables of this rank's
// mailbox and compacts the result into the page-locked block (synchronize.go:16-53 + the final aggregate of
// physicalplan.go:438-471).
int32_t cached_tail_merged(fgpu_ctx* ctx, ExecCache& x, uint64_t seq, uint64_t bytes) {
  Comm& cm = ctx->comm;
  const uint64_t set = seq & 1u;
  DenseOut& f = x.dout;
  f.n_src = cm.n;
  for (int r = 0; r < cm.n; r++) f.src[r] = cm.mailbox + cm.data_off(set, r);
  f.flags = reinterpret_cast<const unsigned long long*>(cm.mailbox + cm.flag_off(set, 0));
  f.seq = seq;
  f.bytes = bytes;
  const char* to = getenv("FROSTGPU_COMM_TIMEOUT_MS");
  f.timeout_ns = uint64_t(to ? std::max(1, atoi(to)) : 10000) * 1000000ull;
  f.err = x.qd.counters + 3;
  const int32_t rc = cached_tail(ctx, x);
  f.n_src = 0;
  return rc;
}

// collective: 0 no exchange, 1 push + wait + merge in one go, 2 push only (fgpu_query_execute_collective_begin)
int32_t run_cached(fgpu_ctx* ctx, ExecCache& x, fgpu_result* res, int collective) {
  PhaseClock pc;
  struct Tail { PhaseClock& pc; ~Tail() { pc.flush("cached"); } } tail{pc};
  cudaStream_t s = ctx->stream;
  res->stats = x.stats;
  fgpu_stats& st = res->stats;
  CUDA_TRY(cudaEventRecord(ctx->ev[0], s));
  CUDA_TRY(launch_table_init(x.qd, s));
  CUDA_TRY(cudaEventRecord(ctx->ev[1], s));
  if (x.has_runs) {
    if (x.runs_v1) CUDA_TRY(launch_runs(x.rd, x.runs_nl, x.runs_nk, x.runs_na, ctx->sm_count, s));
    else CUDA_TRY(launch_runs_tma(x.rd, x.runs_nl, x.runs_nk, x.runs_na, ctx->sm_count, s));
  }
  if (x.has_ta) CUDA_TRY(launch_tile_agg(x.td, ctx->sm_count, s));
  CUDA_TRY(launch_scan(x.qdesc_dev, x.qd, ctx->sm_count, s));
  CUDA_TRY(cudaEventRecord(ctx->ev[2], s));
  if (collective) {
    uint64_t seq = 0, bytes = 0;
    int32_t rc = comm_push(ctx, x.table.p, x.table_bytes, &seq, &bytes);
    if (rc) return rc;
    st.kernel_launches++;  // k_comm_push
    if (collective == 2) {  // the merge half follows in fgpu_query_execute_collective_end
      res->pending = true;
      res->pending_cached = &x;
      res->pending_seq = seq;
      res->pending_bytes = bytes;
      return FGPU_OK;
    }
    if (x.qd.table_mode == TM_DENSE) {
      if (int32_t rcm = cached_tail_merged(ctx, x, seq, bytes)) return rcm;
    } else {
      rc = comm_wait_merge(ctx, x.qd, x.table.p, seq, bytes);
      if (rc) return rc;
      if (int32_t rct = cached_tail(ctx, x)) return rct;
    }
  } else if (int32_t rc = cached_tail(ctx, x)) {
    return rc;
  }
  CUDA_TRY(cudaEventRecord(ctx->ev[3], s));
  pc.mark("launch");
  CUDA_TRY(cudaStreamSynchronize(s));
  pc.mark("sync");
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
  st.scan_kernel_ms = ms;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
  st.total_device_ms = ms;
  const unsigned long long* counters = reinterpret_cast<const unsigned long long*>(x.pinned + 32);
  st.rows_selected = counters[0];
  st.d2h_bytes += x.out_bytes;
  if (int32_t rc = comm_check(counters[3])) return rc;
  const int32_t rc = finalize_dense_host(ctx, res, x);
  pc.mark("export");
  return rc;
}

// collective: 0 no exchange; 1 / 2: the first run of a plan pushes its partial table and leaves the result pending
// (the merge half is collective_end); a cached plan does what run_cached does for the mode.
int32_t run_scan(fgpu_ctx* ctx, const QueryPlan& q, uint64_t tx, fgpu_result* res, bool count_groups = false, bool cacheable = false,
                 int collective = 0) {
  PhaseClock pc;
  Table& plan_table = ctx->tables[q.table];
  if (plan_table.epoch == 0) plan_table.epoch = ++ctx->epoch_counter;
  const std::string env_now = env_switches();
  if (!q.plan_cache || q.cache_epoch != plan_table.epoch || q.cache_tx != tx || q.cache_env != env_now || env_has(env_now, "FROSTGPU_NO_PLAN_CACHE")) {
    std::shared_ptr<void> fresh(new Compiled(), [](void* p) { delete static_cast<Compiled*>(p); });
    int32_t rc0 = compile(ctx, q, tx, static_cast<Compiled*>(fresh.get()));
    q.plan_cache.reset();
    if (rc0) return rc0;
    q.plan_cache = std::move(fresh);
    q.cache_epoch = plan_table.epoch;
    q.cache_tx = tx;
    q.cache_env = env_now;
  } else {
    static_cast<Compiled*>(q.plan_cache.get())->h2d_bytes = 0;  // nothing is uploaded by a cached plan
  }
  Compiled& c = *static_cast<Compiled*>(q.plan_cache.get());
  if (cacheable && c.exec && c.exec->ready && !env_has(env_now, "FROSTGPU_NO_EXEC_CACHE")) return run_cached(ctx, *c.exec, res, collective);
  c.exec.reset();
  for (LeafHost& lh : c.leaves) lh.lut_off = size_t(-1);  // per-Execute state of the plan
  int32_t rc = FGPU_OK;
  (void)rc;
  pc.mark("compile");
  QueryDesc qd = c.qd;  // a copy: the plan in `c` may be reused by the next Execute
  const int n_rg = qd.n_rg, n_slots = qd.n_slots, n_leaves = qd.n_leaves;
  fgpu_stats& st = res->stats;
  st.rows_scanned = c.total_rows;
  st.rows_touched = c.touched_rows == ~uint64_t(0) ? c.total_rows : c.touched_rows;
  st.row_groups = uint32_t(n_rg);
  st.row_groups_pruned = c.pruned_row_groups;

  // ---- per row group tables: chunk descriptors, leaf runtime, leaf LUTs, tile prefix -----------
  std::vector<ChunkDesc> chunks(size_t(n_rg) * std::max(n_slots, 1));
  std::vector<LeafRt> lrt(size_t(n_rg) * std::max(n_leaves, 1));
  std::vector<uint32_t> first_tile(size_t(n_rg) + 1), rg_rows(std::max(n_rg, 1));
  std::vector<uint8_t> lutbytes;
  std::vector<std::pair<size_t, size_t>> lut_fix;  // (index into lrt, offset into lutbytes)
  uint32_t tiles = 0;
  const uint32_t tile_len = q.kind == FGPU_PLAN_FILTER ? uint32_t(kTileRows) : uint32_t(qd.vl);  // rows plan: CTA tiles; scan: warp vectors
  // ---- sorted-run scan: query-level eligibility (see runs_scan.cu) --------------------------------
  // One launch over the qualifying row groups; a row group whose statistics already decided every
  // leaf is flagged all_pass (its leaf columns are neither uploaded nor staged).
  struct RunsBatch {
    RunsDesc rd{};
    int nl = 0;
    int col_slot[kRunsCols] = {0};
    std::vector<RunsRg> rgs;
    std::vector<uint32_t> rows, first_span;
    size_t o_rg = 0, o_span = 0;
  };
  RunsBatch batches[1];
  int runs_nl = 0, runs_nk = 0, runs_na = 0, runs_np = 0;
  int runs_leaf_index[kRunsLeaves] = {0}, runs_pred_leaf[kRunsPreds] = {0};  // query leaf behind every range / dictionary leaf
  int runs_leaf_slot[kRunsLeaves] = {0}, runs_agg_slot[kRunsAggs] = {0}, runs_agg_index[kRunsAggs] = {0};
  uint32_t runs_agg_func[kRunsAggs] = {0};
  bool runs_q = q.kind != FGPU_PLAN_FILTER && c.runs_shape && !getenv("FROSTGPU_NO_RUNS");
  const bool runs_v1 = getenv("FROSTGPU_RUNS_V1") != nullptr;  // the warp-private cp.async kernel (runs_scan.cu), kept for A/B measurements
  if (runs_q) {
    for (int l = 0; l < n_leaves && runs_q; l++) {
      const LeafDesc& ld = qd.leaves[l];
      if (ld.slot != 0xff && qd.slot_type[ld.slot] == ST_DICT) {  // dictionary leaf: evaluated once per run
        if (runs_np >= kRunsPreds) runs_q = false;
        else runs_pred_leaf[runs_np++] = l;
      } else if (ld.slot == 0xff) {  // column absent from every row group: ALL / NONE everywhere, nothing to evaluate
        if (runs_np >= kRunsPreds) runs_q = false;
        else runs_pred_leaf[runs_np++] = l;
      } else if (ld.cmp_float || ld.neg || qd.slot_type[ld.slot] != ST_I64 || runs_nl >= kRunsLeaves) {
        runs_q = false;
      } else {
        runs_leaf_index[runs_nl] = l;
        runs_leaf_slot[runs_nl++] = ld.slot;
      }
    }
    for (int a = 0; a < qd.n_aggs && runs_q; a++) {
      if (qd.aggs[a].func == FGPU_AGG_COUNT) continue;
      if ((qd.aggs[a].func != FGPU_AGG_SUM && qd.aggs[a].func != FGPU_AGG_MIN && qd.aggs[a].func != FGPU_AGG_MAX) || runs_na >= kRunsAggs) { runs_q = false; break; }
      runs_agg_slot[runs_na] = qd.prog[qd.aggs[a].prog_off].slot;
      runs_agg_func[runs_na] = uint32_t(qd.aggs[a].func) | (qd.aggs[a].is_float ? 0x100u : 0u);
      runs_agg_index[runs_na++] = a;
    }
    if (qd.n_keys > kRunsKeys) runs_q = false;
    runs_nk = qd.n_keys;
  }
  if (runs_q) {  // distinct staged columns of each batch
    for (int bi = 0; bi < 1; bi++) {
      RunsBatch& B = batches[bi];
      auto col_of = [&](int slot) {
        for (uint32_t i = 0; i < B.rd.n_cols; i++)
          if (B.col_slot[i] == slot) return i;
        B.col_slot[B.rd.n_cols] = slot;
        return B.rd.n_cols++;
      };
      B.nl = runs_nl;
      for (int l = 0; l < B.nl; l++) B.rd.leaf_col[l] = col_of(runs_leaf_slot[l]);
      for (int a = 0; a < runs_na; a++) {
        B.rd.agg_col[a] = col_of(runs_agg_slot[a]);
        B.rd.agg_func[a] = runs_agg_func[a];
      }
      for (int k = 0; k < runs_nk; k++) B.rd.stride[k] = qd.keys[k].dense_stride;
      B.rd.n_pred = uint32_t(runs_np);
    }
  }
  // ---- tile aggregate: query-level eligibility (see tile_agg.cu) -------------------------------------
  // Row groups the sorted-run kernel does not take (bit-packed / short-run / nullable keys, L0 records) go to
  // the tile-aggregate kernel when the plan is a conjunction of range + dictionary leaves over a dense table
  // with plain aggregate inputs; what is left after that stays with the general scan kernel.
  struct TaBatch {
    TileAggDesc td{};
    std::vector<TileAggRg> rgs;
    std::vector<uint32_t> rows, first_tile;
    size_t o_rg = 0, o_tile = 0;
    int plain_slot[kTaPlain] = {0}, code_slot[kTaCodes] = {0};
    uint32_t code_wmax[kTaCodes] = {0};
    int leaf_q[kTaLeaves] = {0}, pred_q[kTaPreds] = {0}, agg_index[kTaAggs] = {0};
    bool agg_bounded = true;          // every int64 Sum input of every row group has statistics inside [0, 2^32)
    uint64_t agg_max = 0;             // the largest of those chunk maxima
  } TA;
  bool ta_q = q.kind != FGPU_PLAN_FILTER && qd.table_mode == TM_DENSE && (qd.n_filter_prog == 0 || qd.filter_kind == FK_AND) &&
              qd.n_keys <= kTaKeys && !getenv("FROSTGPU_NO_TILE");
  {
    TileAggDesc& td = TA.td;
    auto plain_of = [&](int slot) -> int {
      for (uint32_t i = 0; i < td.n_plain; i++)
        if (TA.plain_slot[i] == slot) return int(i);
      if (td.n_plain >= uint32_t(kTaPlain)) return -1;
      TA.plain_slot[td.n_plain] = slot;
      return int(td.n_plain++);
    };
    auto code_of = [&](int slot) -> int {
      for (uint32_t i = 0; i < td.n_codes; i++)
        if (TA.code_slot[i] == slot) return int(i);
      if (td.n_codes >= uint32_t(kTaCodes)) return -1;
      TA.code_slot[td.n_codes] = slot;
      return int(td.n_codes++);
    };
    for (int l = 0; l < n_leaves && ta_q; l++) {
      const LeafDesc& ld = qd.leaves[l];
      if (ld.slot == 0xff) continue;  // column absent from every row group: ALL / NONE per row group, nothing to evaluate
      if (qd.slot_type[ld.slot] == ST_DICT) {
        const int cc = code_of(ld.slot);
        if (td.np >= uint32_t(kTaPreds) || cc < 0) { ta_q = false; break; }
        TA.pred_q[td.np] = l;
        td.pred_code[td.np++] = uint32_t(cc);
      } else {
        const int pp = plain_of(ld.slot);
        if (c.leaves[size_t(l)].null_literal || td.nl >= uint32_t(kTaLeaves) || pp < 0) { ta_q = false; break; }
        TA.leaf_q[td.nl] = l;
        td.leaf_plain[td.nl] = uint32_t(pp);
        td.leaf_flags[td.nl++] = (ld.cmp_float ? 1u : 0u) | (qd.slot_type[ld.slot] == ST_F64 ? 2u : 0u) | (ld.neg ? 4u : 0u);
      }
    }
    for (int k = 0; k < qd.n_keys && ta_q; k++) {
      const int cc = code_of(qd.keys[k].slot);
      if (qd.keys[k].is_int64 || cc < 0) { ta_q = false; break; }
      td.key_code[td.nk] = uint32_t(cc);
      td.key_stride[td.nk++] = qd.keys[k].dense_stride;
    }
    for (int a = 0; a < qd.n_aggs && ta_q; a++) {
      const AggDesc& ad = qd.aggs[a];
      if (ad.func == FGPU_AGG_COUNT) continue;
      const bool simple = ad.prog_len == 1 && qd.prog[ad.prog_off].op == PO_LOAD;
      const int pp = simple ? plain_of(qd.prog[ad.prog_off].slot) : -1;
      if ((ad.func != FGPU_AGG_SUM && ad.func != FGPU_AGG_MIN && ad.func != FGPU_AGG_MAX) || td.na >= uint32_t(kTaAggs) || pp < 0) { ta_q = false; break; }
      TA.agg_index[td.na] = a;
      td.agg_plain[td.na] = uint32_t(pp);
      td.cell64[td.na] = (ad.func == FGPU_AGG_SUM && !ad.is_float) ? 0u : 1u;
      td.agg_func[td.na++] = uint32_t(ad.func) | (ad.is_float ? 0x100u : 0u);
    }
  }
  uint32_t ta_empty_rgs = 0;
  int gi = 0;  // row groups that stay with the general scan kernel
  for (int g0 = 0; g0 < n_rg; g0++) {
    RowGroupHost& rg = *c.rgs[size_t(g0)].rg;
    const int g = gi;
    for (int s = 0; s < n_slots; s++) {
      ChunkDesc d{};
      d.kind = CK_ABSENT;
      d.n_rows = rg.n_rows;
      auto it = ((c.rgs[size_t(g0)].skip_slots >> s) & 1) ? rg.cols.end() : rg.cols.find(c.slot_names[size_t(s)]);
      if (it != rg.cols.end()) {
        ChunkHost& ch = it->second;
        if (!ch.error.empty())
          return fail(FGPU_ERR_UNSUPPORTED, "column " + c.slot_names[size_t(s)] + ": " + ch.error);
        d = ch.desc;
        st.algorithmic_bytes += ch.stored_bytes;
        st.metadata_bytes += ch.meta_bytes;
      }
      chunks[size_t(g) * n_slots + s] = d;
    }
    for (int l = 0; l < n_leaves; l++) {
      LeafHost& lh = c.leaves[size_t(l)];
      LeafRt rt{};
      auto it = lh.slot < 0 ? rg.cols.end() : rg.cols.find(lh.column);
      if (it == rg.cols.end()) {
        rt.mode = lh.missing_mode;
      } else if (c.slot_types[size_t(lh.slot)] == ST_DICT) {
        rt.mode = LM_EVAL;
        rt.null_result = (lh.op == FGPU_OP_EQ && lh.lit->lit_type == FGPU_SCALAR_NULL) ? 1 : 0;
        if (lh.lut_off == size_t(-1)) {
          // one result byte per GLOBAL dictionary id, evaluated once per distinct dictionary entry
          const GlobalDict& gd = ctx->tables[q.table].dicts[lh.column];
          lh.lut_off = lutbytes.size();
          lutbytes.resize(lh.lut_off + gd.values.size() + 1);
          for (size_t g = 0; g < gd.values.size(); g++) lutbytes[lh.lut_off + g] = dict_leaf_value(lh, gd.values[g]) ? 1 : 0;
        }
        lut_fix.emplace_back(size_t(g) * n_leaves + l, lh.lut_off);
      } else {
        rt.mode = lh.null_literal ? LM_NONE : LM_EVAL;
      }
      if (c.rgs[size_t(g0)].leaf_mode[l] != LM_EVAL) rt.mode = c.rgs[size_t(g0)].leaf_mode[l];
      lrt[size_t(g) * n_leaves + l] = rt;
    }
    // ---- does this row group qualify for the sorted-run kernel? ----
    bool runs_ok = runs_q && rg.n_rows > 0;
    const int bi = 0;
    bool all_pass = false;
    if (runs_ok) {
      bool all_all = runs_nl > 0;
      for (int l = 0; l < runs_nl; l++) {
        const uint8_t m = lrt[size_t(g) * n_leaves + runs_leaf_index[l]].mode;
        if (m != LM_ALL) all_all = false;
        if (m == LM_NONE) runs_ok = false;  // (only with pruning switched off)
      }
      all_pass = all_all;
    }
    RunsRg rr{};
    if (runs_ok) {
      RunsBatch& B = batches[bi];
      rr.n_rows = rg.n_rows;
      for (int l = 0; l < B.nl; l++) {
        const int ql = runs_leaf_index[l];
        const bool all = lrt[size_t(g) * n_leaves + ql].mode == LM_ALL;
        rr.lo[l] = all ? std::numeric_limits<int64_t>::min() : qd.leaves[ql].lo_i;
        rr.hi[l] = all ? std::numeric_limits<int64_t>::max() : qd.leaves[ql].hi_i;
      }
      for (int i = 0; i < runs_np && runs_ok; i++) {  // dictionary leaves: one result per run
        const int ql = runs_pred_leaf[i];
        const LeafRt& rt = lrt[size_t(g) * n_leaves + ql];
        rr.pred_runs[i] = nullptr;
        if (rt.mode == LM_ALL) continue;                      // decided: passes everywhere in this row group
        if (rt.mode == LM_NONE) { runs_ok = false; break; }   // (only with pruning switched off)
        const ChunkDesc& d = chunks[size_t(g) * n_slots + qd.leaves[ql].slot];
        // (short runs — a leaf on a column far down the sort order — are the tile-aggregate kernel's business: its
        // per-row result byte costs one shared load where a cursor would move every few rows)
        if (d.kind != CK_DICT_STR || d.row_runs == nullptr || uint64_t(d.n_row_runs) * 32 > uint64_t(rg.n_rows) + 1024) { runs_ok = false; break; }
        rr.pred_runs[i] = d.row_runs;
        rr.pred_seeds[i] = d.row_seeds;
        rr.pred_n_runs[i] = d.n_row_runs;
        rr.pred_lut[i] = reinterpret_cast<const uint8_t*>(uintptr_t(c.leaves[size_t(ql)].lut_off));  // offset, rebased below
        B.rd.pred_null[i] = rt.null_result;
      }
      rr.all_pass = all_pass ? 1 : 0;
      const uint8_t* any_col = nullptr;
      for (uint32_t i = 0; i < B.rd.n_cols && runs_ok; i++) {
        const ChunkDesc& d = chunks[size_t(g) * n_slots + B.col_slot[i]];
        rr.col[i] = nullptr;
        if ((c.rgs[size_t(g0)].skip_slots >> B.col_slot[i]) & 1) continue;  // a decided leaf's own column
        if (d.kind != CK_PLAIN64 || d.has_nulls) runs_ok = false;
        rr.col[i] = d.values;
        any_col = d.values;
      }
      // all_pass: the leaf columns stay null (not staged).  Otherwise a decided leaf still runs its
      // (widened) range test: any staged column stands in for the one that was not uploaded.
      for (uint32_t i = 0; i < B.rd.n_cols && runs_ok && !all_pass; i++) {
        if (rr.col[i]) continue;
        if (!any_col) runs_ok = false;
        rr.col[i] = any_col;
      }
      if (all_pass)  // aggregate inputs must be there
        for (int a = 0; a < runs_na && runs_ok; a++)
          if (!rr.col[B.rd.agg_col[a]]) runs_ok = false;
      for (int k = 0; k < runs_nk && runs_ok; k++) {
        const ChunkDesc& d = chunks[size_t(g) * n_slots + qd.keys[k].slot];
        // run-length only, and runs long enough that a 32-row step rarely holds two run ends
        // (row_runs: the directory in row space; it exists when the column chunk is run-length only, NULLs included)
        if (d.kind != CK_DICT_STR || d.row_runs == nullptr || uint64_t(d.n_row_runs) * 32 > uint64_t(rg.n_rows) + 1024) runs_ok = false;
        rr.runs[k] = d.row_runs;
        rr.seeds[k] = d.row_seeds;
        rr.n_runs[k] = d.n_row_runs;
      }
    }
    if (runs_ok) {
      batches[bi].rgs.push_back(rr);
      batches[bi].rows.push_back(rg.n_rows);
      lut_fix.erase(std::remove_if(lut_fix.begin(), lut_fix.end(), [&](const std::pair<size_t, size_t>& f) { return f.first >= size_t(g) * n_leaves; }), lut_fix.end());
      continue;  // slot g of the general tables is reused by the next row group
    }
    // ---- does this row group go to the tile-aggregate kernel? ----
    if (ta_q && rg.n_rows > 0) {
      const TileAggDesc& td = TA.td;
      bool ok = true, none = false;
      for (int l = 0; l < n_leaves; l++)
        if (lrt[size_t(g) * n_leaves + l].mode == LM_NONE) none = true;  // conjunction: no row of this row group passes
      TileAggRg tr{};
      tr.n_rows = rg.n_rows;
      bool plain_needed[kTaPlain] = {false};
      for (uint32_t i = 0; i < td.nl && ok; i++) {
        const int ql = TA.leaf_q[i];
        const LeafDesc& ld = qd.leaves[ql];
        if (lrt[size_t(g) * n_leaves + ql].mode == LM_ALL) { tr.leaf_skip[i] = 1; continue; }
        plain_needed[td.leaf_plain[i]] = true;
        if (ld.cmp_float) { std::memcpy(&tr.lo[i], &ld.lo_f, 8); std::memcpy(&tr.hi[i], &ld.hi_f, 8); }
        else { tr.lo[i] = ld.lo_i; tr.hi[i] = ld.hi_i; }
      }
      for (uint32_t i = 0; i < td.na; i++) plain_needed[td.agg_plain[i]] = true;
      for (uint32_t i = 0; i < td.n_plain && ok && !none; i++) {
        const ChunkDesc& d = chunks[size_t(g) * n_slots + TA.plain_slot[i]];
        tr.plain[i] = nullptr;
        if (!plain_needed[i]) continue;
        if (d.kind != CK_PLAIN64 || d.has_nulls) { ok = false; break; }
        tr.plain[i] = d.values;
      }
      bool code_needed[kTaCodes] = {false};
      for (uint32_t i = 0; i < td.np && ok; i++) {
        const int ql = TA.pred_q[i];
        const LeafRt& rt = lrt[size_t(g) * n_leaves + ql];
        tr.pred_lut[i] = nullptr;
        if (rt.mode != LM_EVAL) continue;  // decided: passes (a NONE row group is dropped as a whole)
        code_needed[td.pred_code[i]] = true;
        tr.pred_lut[i] = reinterpret_cast<const uint8_t*>(uintptr_t(c.leaves[size_t(ql)].lut_off) + 1);  // offset + 1, rebased below
        TA.td.pred_null[i] = rt.null_result;
      }
      for (uint32_t i = 0; i < td.nk; i++) code_needed[td.key_code[i]] = true;
      for (uint32_t i = 0; i < td.n_codes && ok && !none; i++) {
        const ChunkDesc& d = chunks[size_t(g) * n_slots + TA.code_slot[i]];
        tr.codes[i] = nullptr;
        if (!code_needed[i] || d.kind == CK_ABSENT) continue;  // absent column: NULL for every row
        if (d.kind != CK_DICT_STR) { ok = false; break; }
        const std::string& name = c.slot_names[size_t(TA.code_slot[i])];
        int32_t frc = ensure_flat(ctx, c.rgs[size_t(g0)].part, name);
        if (frc) return frc;
        const ChunkHost& ch = rg.cols.at(name);
        if (!ch.flat) { ok = false; break; }
        tr.codes[i] = ch.flat;
        tr.code_w[i] = ch.flat_w;
        tr.code_bias[i] = ch.flat_bias;
        TA.code_wmax[i] = std::max<uint32_t>(TA.code_wmax[i], ch.flat_w);
      }
      if (ok) {
        lut_fix.erase(std::remove_if(lut_fix.begin(), lut_fix.end(), [&](const std::pair<size_t, size_t>& f) { return f.first >= size_t(g) * n_leaves; }), lut_fix.end());
        if (none) { ta_empty_rgs++; continue; }
        for (uint32_t i = 0; i < td.na; i++) {  // bounds of the int64 Sum inputs (see TileAggDesc::sums_fit32)
          if (td.cell64[i]) continue;
          const ChunkHost& ch = rg.cols.at(c.slot_names[size_t(TA.plain_slot[td.agg_plain[i]])]);
          if (!ch.has_minmax || ch.phys != PT_INT64 || ch.min_bits < 0 || ch.max_bits > 0xffffffffll) TA.agg_bounded = false;
          else TA.agg_max = std::max<uint64_t>(TA.agg_max, uint64_t(ch.max_bits));
        }
        TA.rgs.push_back(tr);
        TA.rows.push_back(rg.n_rows);
        continue;  // slot g of the general tables is reused by the next row group
      }
    }
    first_tile[size_t(g)] = tiles;
    rg_rows[size_t(g)] = rg.n_rows;
    tiles += (rg.n_rows + tile_len - 1) / tile_len;
    gi++;
  }
  first_tile[size_t(gi)] = tiles;
  qd.n_rg = gi;
  qd.n_tiles = tiles;
  // spans of the sorted-run kernel: sized so that every warp of the grid gets several
  for (RunsBatch& B : batches) {
    if (B.rgs.empty()) continue;
    RunsDesc& rd = B.rd;
    auto envi = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
    rd.n_rg = uint32_t(B.rgs.size());
    uint64_t total = 0;
    for (uint32_t r : B.rows) total += r;
    int br = 0;
    uint64_t span_blocks = 0;
    if (!runs_v1) {
      // CTA-cooperative kernel (runs_tma.cu): tiles of br rows, spans of span_blocks tiles dealt to the CTAs in turn
      // CTA-cooperative kernel (runs_tma.cu): every row group is cut into tiles of about 32 KB of column data; the
      // launch's tile table {row group, first row} is dealt to the CTAs in chunks of consecutive entries
      uint32_t base_tile = 4096;
      { const int f = envi("FROSTGPU_RT_TILE", 0); if (f == 1024 || f == 2048 || f == 4096 || f == 8192) base_tile = uint32_t(f); }
      uint32_t tiles_rt = 0, gidx = 0, col_region = 0;
      for (RunsRg& r : B.rgs) {
        uint32_t nc = 0;
        for (uint32_t i = 0; i < rd.n_cols; i++) {
          r.col_pos[i] = uint8_t(nc);
          if (r.col[i]) nc++;
        }
        r.tile_rows = runs_tma_tile_rows(base_tile, nc);
        col_region = std::max(col_region, nc * r.tile_rows * 8u);
        for (uint32_t r0 = 0; r0 < r.n_rows; r0 += r.tile_rows) {
          B.first_span.push_back(gidx);
          B.first_span.push_back(r0);
          tiles_rt++;
        }
        gidx++;
      }
      runs_tma_plan(rd, runs_nk + runs_np, base_tile, col_region, envi("FROSTGPU_RT_STAGES", 0), envi("FROSTGPU_RT_WARPS", 0), envi("FROSTGPU_RT_SPAN", 0));
      rd.n_spans = tiles_rt;
      (void)br; (void)span_blocks;
      continue;
    } else {
      br = envi("FROSTGPU_RUNS_BR", 256);
      if (br != 128 && br != 256 && br != 512) br = 256;
      int ring = envi("FROSTGPU_RUNS_RING", 2);
      ring = std::min(4, std::max(2, ring));
      if (rd.n_cols == 0) { br = 128; ring = 2; }  // nothing is staged (count per group): the ring is idle
      rd.block_rows = uint32_t(br);
      rd.n_ring = uint32_t(ring);
      int per_sm = 1;
      CUDA_TRY(runs_blocks_per_sm(rd, B.nl, runs_nk, runs_na, &per_sm));
      const uint64_t warps = uint64_t(ctx->sm_count) * uint64_t(std::max(per_sm, 1)) * (kRunsThreads / 32);
      span_blocks = total / (warps * 8 * uint64_t(br));
      span_blocks = std::min<uint64_t>(64, std::max<uint64_t>(4, span_blocks));
      if (const char* e = getenv("FROSTGPU_RUNS_SPAN")) span_blocks = uint64_t(std::max(1, atoi(e)));
    }
    rd.span_blocks = uint32_t(span_blocks);
    const uint64_t span_rows = span_blocks * uint64_t(br);
    uint32_t spans = 0;
    for (uint32_t r : B.rows) {
      B.first_span.push_back(spans);
      spans += uint32_t((uint64_t(r) + span_rows - 1) / span_rows);
    }
    B.first_span.push_back(spans);
    rd.n_spans = spans;
  }

  // ---- tile aggregate: tile size, ring depth and the shared-memory table -------------------------------
  if (!TA.rgs.empty()) {
    TileAggDesc& td = TA.td;
    auto envi = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
    td.n_rg = uint32_t(TA.rgs.size());
    td.table_slots = qd.table_slots;
    // replicas of small tables: same-address CAS on 64-bit cells serialises, 32 lanes on 17 slots
    uint32_t rl = 0;
    while (rl < 5 && (uint64_t(qd.table_slots) << (rl + 1)) <= 2048) rl++;
    td.rep_log2 = rl;
    const uint64_t cells = uint64_t(qd.table_slots) << rl;
    uint64_t tb = (cells * 4 + 7) & ~uint64_t(7);
    for (uint32_t a = 0; a < td.na; a++) {
      td.cell_off[a] = uint32_t(tb);
      tb += (cells * (td.cell64[a] ? 8 : 4) + 7) & ~uint64_t(7);
    }
    const uint64_t kSmemMax = 227 * 1024;
    // tile rows: multiples of the rows the consumer threads take per turn (kTaConsumerWarps * 32 lanes * 4 rows)
    const int cand[10][2] = {{6144, 3}, {3072, 4}, {3072, 3}, {6144, 2}, {3072, 2}, {1536, 4}, {1536, 3}, {1536, 2}, {768, 3}, {768, 2}};
    auto slot_bytes_for = [&](uint32_t T) {
      uint64_t off = 0;
      for (uint32_t i = 0; i < td.n_plain; i++) { td.plain_off[i] = uint32_t(off); off += uint64_t(T) * 8; }
      for (uint32_t i = 0; i < td.n_codes; i++) {
        td.code_off[i] = uint32_t(off);
        off += (uint64_t(T) * std::max<uint32_t>(TA.code_wmax[i], 8) / 8 + 127) & ~uint64_t(127);
      }
      return std::max<uint64_t>(off, 128);
    };
    bool fits = false;
    const int force_t = envi("FROSTGPU_TA_TILE", 0), force_s = envi("FROSTGPU_TA_STAGES", 0);
    for (int smem_table = tb < kSmemMax && !getenv("FROSTGPU_TA_GLOBAL") ? 1 : 0; smem_table >= 0 && !fits; smem_table--) {
      for (const auto& cs : cand) {
        if ((force_t && cs[0] != force_t) || (force_s && cs[1] != force_s)) continue;
        const uint64_t sb = slot_bytes_for(uint32_t(cs[0]));
        const uint64_t total = 128 + uint64_t(cs[1]) * (256 + sb) + (smem_table ? tb : 0);
        if (total > kSmemMax) continue;
        td.tile_rows = uint32_t(cs[0]);
        td.n_stages = uint32_t(cs[1]);
        td.slot_bytes = uint32_t(sb);
        td.smem_table = uint32_t(smem_table);
        fits = true;
        break;
      }
    }
    if (!fits) return fail(FGPU_ERR_UNSUPPORTED, "tile aggregate: the projected columns do not fit the shared-memory ring");
    td.table_bytes = uint32_t(td.smem_table ? tb : 0);
    uint32_t t = 0;
    for (uint32_t r : TA.rows) {
      TA.first_tile.push_back(t);
      t += (r + td.tile_rows - 1) / td.tile_rows;
    }
    TA.first_tile.push_back(t);
    td.n_tiles = t;
    td.chunk_tiles = std::max<uint32_t>(1, t / (uint32_t(ctx->sm_count) * 4));  // ~4 turns per CTA: a CTA stays inside one row group for a whole chunk
    td.keys8 = 1;
    for (uint32_t i = 0; i < td.nk; i++)
      if (TA.code_wmax[td.key_code[i]] > 8) td.keys8 = 0;
    {  // can a Sum leave 32 bits inside one CTA?  rows per CTA <= its share of the tiles (+ one chunk), all in one slot
      const uint64_t grid = uint64_t(std::max(1, ctx->sm_count));
      const uint64_t rows_per_cta = (uint64_t(t) / grid + 2 * uint64_t(td.chunk_tiles) + 1) * td.tile_rows;
      td.sums_fit32 = (TA.agg_bounded && !getenv("FROSTGPU_TA_CARRY") && TA.agg_max * rows_per_cta < (1ull << 32)) ? 1u : 0u;
    }
    if (const char* e = getenv("FROSTGPU_TA_CHUNK")) td.chunk_tiles = uint32_t(std::max(1, atoi(e)));
  }

  // ---- filter-only plans over PLAIN columns: ordered take (take_rows.cu) instead of k_rows ------------
  TakeDesc td{};
  std::vector<TakeRg> take_rgs;
  std::vector<uint32_t> take_first_span;
  bool take_q = q.kind == FGPU_PLAN_FILTER && qd.n_out >= 1 && qd.n_out <= kTakeOut && n_leaves <= kTakeLeaves + kTakePreds &&
                (qd.n_filter_prog == 0 || qd.filter_kind == FK_AND) && gi > 0 && !getenv("FROSTGPU_NO_TAKE");
  // range leaves (PLAIN int64 columns) and dictionary leaves (flat codes + result byte per dictionary id) of the plan
  int take_range[kTakeLeaves] = {0}, take_pred[kTakePreds] = {0}, n_tr = 0, n_tp = 0;
  for (int l = 0; l < n_leaves && take_q; l++) {
    const LeafDesc& ld = qd.leaves[l];
    if (ld.slot == 0xff) { take_q = false; break; }
    if (qd.slot_type[ld.slot] == ST_DICT) {
      if (n_tp >= kTakePreds || gi != n_rg) take_q = false;
      else take_pred[n_tp++] = l;
    } else if (ld.cmp_float || ld.neg || qd.slot_type[ld.slot] != ST_I64 || c.leaves[size_t(l)].null_literal || n_tr >= kTakeLeaves) {
      take_q = false;
    } else {
      take_range[n_tr++] = l;
    }
  }
  for (int o = 0; o < qd.n_out && take_q; o++)
    if (qd.slot_type[qd.out_slot[o]] == ST_DICT) take_q = false;
  for (int g = 0; g < gi && take_q; g++) {
    TakeRg tr{};
    tr.n_rows = rg_rows[size_t(g)];
    bool all = n_leaves > 0;
    for (int l = 0; l < n_leaves && take_q; l++) {
      const uint8_t m = lrt[size_t(g) * n_leaves + l].mode;
      if (m == LM_NONE) take_q = false;  // (only with pruning switched off)
      if (m != LM_ALL) all = false;
    }
    tr.all_pass = all ? 1 : 0;
    for (int i = 0; i < n_tr && take_q; i++) {
      const int l = take_range[i];
      const bool decided = lrt[size_t(g) * n_leaves + l].mode == LM_ALL;
      tr.lo[i] = decided ? std::numeric_limits<int64_t>::min() : qd.leaves[l].lo_i;
      tr.hi[i] = decided ? std::numeric_limits<int64_t>::max() : qd.leaves[l].hi_i;
      tr.leaf_col[i] = nullptr;  // decided (or nothing to evaluate at all): the kernels skip the leaf
      if (all || decided) continue;
      const ChunkDesc& d = chunks[size_t(g) * n_slots + qd.leaves[l].slot];
      if (d.kind != CK_PLAIN64 || d.has_nulls) take_q = false;
      tr.leaf_col[i] = d.values;
    }
    for (int i = 0; i < n_tp && take_q; i++) {
      const int l = take_pred[i];
      const LeafRt& rt = lrt[size_t(g) * n_leaves + l];
      tr.pred_codes[i] = nullptr;
      if (all || rt.mode == LM_ALL) continue;
      const ChunkDesc& d = chunks[size_t(g) * n_slots + qd.leaves[l].slot];
      if (d.kind != CK_DICT_STR) { take_q = false; break; }
      const std::string& name = c.slot_names[size_t(qd.leaves[l].slot)];
      int32_t frc = ensure_flat(ctx, c.rgs[size_t(g)].part, name);  // (rows plans: every row group is a general one, g == its index in c.rgs)
      if (frc) return frc;
      const ChunkHost& ch = c.rgs[size_t(g)].rg->cols.at(name);
      if (!ch.flat) { take_q = false; break; }
      tr.pred_codes[i] = ch.flat;
      tr.pred_w[i] = ch.flat_w;
      tr.pred_bias[i] = ch.flat_bias;
      tr.pred_null[i] = rt.null_result;
      tr.pred_lut[i] = reinterpret_cast<const uint8_t*>(uintptr_t(c.leaves[size_t(l)].lut_off) + 1);  // offset + 1, rebased below
    }
    for (int o = 0; o < qd.n_out && take_q; o++) {
      const ChunkDesc& d = chunks[size_t(g) * n_slots + qd.out_slot[o]];
      if (d.kind != CK_PLAIN64 || d.has_nulls) take_q = false;
      tr.out_col[o] = d.values;
    }
    take_rgs.push_back(tr);
  }
  if (take_q) {
    const uint64_t warps = uint64_t(take_resident_warps(ctx->sm_count));
    uint64_t total = 0;
    for (const TakeRg& r : take_rgs) total += r.n_rows;
    uint64_t span_blocks = std::min<uint64_t>(64, std::max<uint64_t>(4, total / (warps * 4 * 256)));
    td.span_blocks = uint32_t(span_blocks);
    const uint64_t span_rows = span_blocks * 256;
    uint32_t spans = 0;
    for (const TakeRg& r : take_rgs) {
      take_first_span.push_back(spans);
      spans += uint32_t((uint64_t(r.n_rows) + span_rows - 1) / span_rows);
    }
    take_first_span.push_back(spans);
    td.n_spans = spans;
    td.n_rg = uint32_t(take_rgs.size());
    td.nl = uint32_t(n_tr);
    td.np = uint32_t(n_tp);
    td.n_out = uint32_t(qd.n_out);
  }

  // ---- device memory -----------------------------------------------------------------------
  const bool rows_plan = q.kind == FGPU_PLAN_FILTER;
  // small dense tables of a re-executable query: persistent device state owned by the plan (ExecCache)
  const bool cache_mode = cacheable && !rows_plan && qd.table_mode == TM_DENSE && qd.table_slots <= 65536 && !getenv("FROSTGPU_NO_EXEC_CACHE");
  if (cache_mode) c.exec.reset(new ExecCache());
  DevBuf& tbuf = cache_mode ? c.exec->table : res->table;
  DevBuf& abuf = cache_mode ? c.exec->aux : res->aux;
  size_t tbytes = 0;
  if (!rows_plan) {
    size_t off_aggs, off_tags, off_keys;
    tbytes = table_layout(qd, &off_aggs, &off_tags, &off_keys);
    CUDA_TRY(tbuf.alloc(tbytes + 128, ctx->stream));  // (+ counters in cache mode; slack for the 16-byte exchange copies)
    res->table_bytes = tbytes;
    bind_table(&qd, static_cast<uint8_t*>(tbuf.p));
  }
  // rows plan: one output array per projected column (worst case every row is selected) + look-back state
  std::vector<size_t> out_off(size_t(qd.n_out)), valid_off(size_t(qd.n_out));
  size_t out_bytes = 0, state_off = 0;
  if (rows_plan) {
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    for (int o = 0; o < qd.n_out; o++) {
      bool dict = qd.slot_type[qd.out_slot[o]] == ST_DICT;
      out_off[size_t(o)] = out_bytes;
      out_bytes = al(out_bytes + size_t(c.total_rows) * (dict ? 4 : 8));
      valid_off[size_t(o)] = out_bytes;
      if (!dict) out_bytes = al(out_bytes + size_t(c.total_rows));
    }
    state_off = out_bytes;
    out_bytes = al(out_bytes + size_t(tiles) * 8);
    CUDA_TRY(tbuf.alloc(out_bytes, ctx->stream));
    uint8_t* ob = static_cast<uint8_t*>(tbuf.p);
    for (int o = 0; o < qd.n_out; o++) {
      qd.out_data[o] = ob + out_off[size_t(o)];
      qd.out_valid[o] = ob + valid_off[size_t(o)];
    }
    qd.tile_state = reinterpret_cast<unsigned long long*>(ob + state_off);
  }

  auto align16 = [](size_t x) { return (x + 15) & ~size_t(15); };
  size_t o_chunks = 0;
  size_t o_lrt = align16(o_chunks + chunks.size() * sizeof(ChunkDesc));
  size_t o_first = align16(o_lrt + lrt.size() * sizeof(LeafRt));
  size_t o_rows = align16(o_first + first_tile.size() * 4);
  size_t o_lut = align16(o_rows + rg_rows.size() * 4);
  size_t o_cnt = align16(o_lut + lutbytes.size());
  for (RunsBatch& B : batches) {
    B.o_rg = o_cnt;
    B.o_span = align16(B.o_rg + B.rgs.size() * sizeof(RunsRg));
    o_cnt = align16(B.o_span + B.first_span.size() * 4);
  }
  if (!TA.rgs.empty()) {
    TA.o_rg = o_cnt;
    TA.o_tile = align16(TA.o_rg + TA.rgs.size() * sizeof(TileAggRg));
    o_cnt = align16(TA.o_tile + TA.first_tile.size() * 4);
  }
  size_t o_take_rg = 0, o_take_span = 0, o_take_count = 0;
  if (take_q) {
    o_take_rg = o_cnt;
    o_take_span = align16(o_take_rg + take_rgs.size() * sizeof(TakeRg));
    o_take_count = align16(o_take_span + take_first_span.size() * 4);
    o_cnt = align16(o_take_count + (size_t(td.n_spans) + 1) * 8);
  }
  // device block: [tables | counters 64 B | group counter 16 B | QueryDesc]; the host image of it lives in
  // page-locked scratch so that one asynchronous copy uploads everything
  const size_t o_qd = align16(o_cnt + 64 + 16);
  const size_t aux_bytes = o_qd + sizeof(QueryDesc);
  const size_t o_ret = align16(aux_bytes);  // scratch only: counters + group count read back
  CUDA_TRY(abuf.alloc(aux_bytes, ctx->stream));
  uint8_t* aux = static_cast<uint8_t*>(abuf.p);
  for (auto& fx : lut_fix) lrt[fx.first].lut = aux + o_lut + fx.second;
  CUDA_TRY(ctx->ensure_scratch(o_ret + 128));
  struct { uint8_t* p; uint8_t* data() { return p; } } hostaux{ctx->scratch};
  std::memset(hostaux.data(), 0, o_ret + 128);
  std::memcpy(hostaux.data() + o_chunks, chunks.data(), chunks.size() * sizeof(ChunkDesc));
  std::memcpy(hostaux.data() + o_lrt, lrt.data(), lrt.size() * sizeof(LeafRt));
  std::memcpy(hostaux.data() + o_first, first_tile.data(), first_tile.size() * 4);
  std::memcpy(hostaux.data() + o_rows, rg_rows.data(), rg_rows.size() * 4);
  if (!lutbytes.empty()) std::memcpy(hostaux.data() + o_lut, lutbytes.data(), lutbytes.size());
  for (RunsBatch& B : batches) {
    if (B.rgs.empty()) continue;
    for (RunsRg& r : B.rgs)
      for (int i = 0; i < runs_np; i++)
        if (r.pred_runs[i]) r.pred_lut[i] = aux + o_lut + size_t(uintptr_t(r.pred_lut[i]));
    std::memcpy(hostaux.data() + B.o_rg, B.rgs.data(), B.rgs.size() * sizeof(RunsRg));
    std::memcpy(hostaux.data() + B.o_span, B.first_span.data(), B.first_span.size() * 4);
  }
  if (!TA.rgs.empty()) {
    for (TileAggRg& r : TA.rgs)
      for (int i = 0; i < kTaPreds; i++)
        if (r.pred_lut[i]) r.pred_lut[i] = aux + o_lut + (size_t(uintptr_t(r.pred_lut[i])) - 1);
    std::memcpy(hostaux.data() + TA.o_rg, TA.rgs.data(), TA.rgs.size() * sizeof(TileAggRg));
    std::memcpy(hostaux.data() + TA.o_tile, TA.first_tile.data(), TA.first_tile.size() * 4);
  }
  if (take_q) {
    for (TakeRg& r : take_rgs)
      for (int i = 0; i < kTakePreds; i++)
        if (r.pred_lut[i]) r.pred_lut[i] = aux + o_lut + (size_t(uintptr_t(r.pred_lut[i])) - 1);
    std::memcpy(hostaux.data() + o_take_rg, take_rgs.data(), take_rgs.size() * sizeof(TakeRg));
    std::memcpy(hostaux.data() + o_take_span, take_first_span.data(), take_first_span.size() * 4);
  }
  qd.chunks = reinterpret_cast<const ChunkDesc*>(aux + o_chunks);
  qd.leaf_rt = reinterpret_cast<const LeafRt*>(aux + o_lrt);
  qd.rg_first_tile = reinterpret_cast<const uint32_t*>(aux + o_first);
  qd.rg_rows = reinterpret_cast<const uint32_t*>(aux + o_rows);
  qd.counters = cache_mode ? reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(tbuf.p) + tbytes)  // behind the table: one copy brings both back
                           : reinterpret_cast<unsigned long long*>(aux + o_cnt);
  const QueryDesc* qdesc_dev = reinterpret_cast<const QueryDesc*>(aux + o_qd);

  FinalizeDesc& fd = res->fd;
  fd = FinalizeDesc{};
  fd.table_mode = qd.table_mode;
  fd.key_words = qd.key_words;
  fd.n_keys = qd.n_keys;
  fd.n_aggs = qd.n_aggs;
  fd.table_slots = qd.table_slots;
  for (int k = 0; k < qd.n_keys; k++) {
    fd.keys[k] = qd.keys[k];
    fd.dense_radix[k] = c.dense_radix[size_t(k)] ? c.dense_radix[size_t(k)] : 1;
  }
  fd.t_rows = qd.t_rows;
  for (int a = 0; a < kMaxAggs; a++) fd.t_agg[a] = qd.t_agg[a];
  fd.t_tag = qd.t_tag;
  fd.t_keys = qd.t_keys;
  cudaStream_t s = ctx->stream;
  pc.mark("describe");
  CUDA_TRY(cudaEventRecord(ctx->ev[0], s));
  std::memcpy(hostaux.data() + o_qd, &qd, sizeof(QueryDesc));
  CUDA_TRY(cudaMemcpyAsync(aux, hostaux.data(), aux_bytes, cudaMemcpyHostToDevice, s));
  if (rows_plan) {
    CUDA_TRY(cudaMemsetAsync(qd.tile_state, 0, size_t(tiles) * 8, s));
    CUDA_TRY(cudaEventRecord(ctx->ev[1], s));
    if (take_q) {
      td.rgs = reinterpret_cast<const TakeRg*>(aux + o_take_rg);
      td.rg_first_span = reinterpret_cast<const uint32_t*>(aux + o_take_span);
      td.span_count = reinterpret_cast<unsigned long long*>(aux + o_take_count);
      td.total = qd.counters;
      for (int o = 0; o < qd.n_out; o++) td.out_data[o] = static_cast<long long*>(qd.out_data[o]);
      CUDA_TRY(launch_take(td, ctx->sm_count, s));
      st.kernel_launches += 2;
      res->rows_no_nulls = true;
    } else {
      CUDA_TRY(launch_rows(qdesc_dev, qd, ctx->sm_count, s));
    }
  } else {
    CUDA_TRY(launch_table_init(qd, s));
    CUDA_TRY(cudaEventRecord(ctx->ev[1], s));
    for (RunsBatch& B : batches) {
      if (B.rgs.empty()) continue;
      RunsDesc& rd = B.rd;
      rd.rgs = reinterpret_cast<const RunsRg*>(aux + B.o_rg);
      rd.rg_first_span = reinterpret_cast<const uint32_t*>(aux + B.o_span);
      rd.t_rows = qd.t_rows;
      for (int a = 0; a < runs_na; a++) rd.t_agg[a] = qd.t_agg[runs_agg_index[a]];
      rd.counters = qd.counters;
      if (runs_v1) CUDA_TRY(launch_runs(rd, B.nl, runs_nk, runs_na, ctx->sm_count, s));
      else CUDA_TRY(launch_runs_tma(rd, B.nl, runs_nk, runs_na, ctx->sm_count, s));
      st.kernel_launches++;
      st.row_groups_runs += uint32_t(B.rgs.size());
    }
    if (!TA.rgs.empty()) {
      TileAggDesc& td = TA.td;
      td.rgs = reinterpret_cast<const TileAggRg*>(aux + TA.o_rg);
      td.rg_first_tile = reinterpret_cast<const uint32_t*>(aux + TA.o_tile);
      td.t_rows = qd.t_rows;
      for (uint32_t a = 0; a < td.na; a++) td.t_agg[a] = qd.t_agg[TA.agg_index[a]];
      td.counters = qd.counters;
      CUDA_TRY(launch_tile_agg(td, ctx->sm_count, s));
      st.kernel_launches++;
      st.row_groups_tiles += uint32_t(TA.rgs.size());
    }
    CUDA_TRY(launch_scan(qdesc_dev, qd, ctx->sm_count, s));
  }
  CUDA_TRY(cudaEventRecord(ctx->ev[2], s));
  st.kernel_launches += (rows_plan ? 0 : 1) + (tiles ? 1 : 0);
  st.h2d_bytes += aux_bytes + sizeof(QueryDesc);
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(hostaux.data() + o_ret);
  res->cnt_ptr = reinterpret_cast<unsigned int*>(qd.counters + 8);  // 16 bytes behind the counters, zeroed by the upload / the table init
  if (collective) {
    if (rows_plan) return fail(FGPU_ERR_INVALID, "rows plans have no partial table: execute them per rank");
    int32_t rcx = comm_push(ctx, tbuf.p, tbytes, &res->pending_seq, &res->pending_bytes);
    if (rcx) return rcx;
    res->pending = true;
    res->plan_keep = q.plan_cache;
    count_groups = false;  // the table is not final yet
  }
  if (cache_mode) {
    // ---- the plan keeps this device state; the table comes back with one copy and is compacted on the host ----
    ExecCache& x = *c.exec;
    x.table_bytes = tbytes;
    x.ctx = ctx;
    {
      DenseOut& f = x.dout;
      f.table_slots = qd.table_slots;
      f.max_out = qd.table_slots;
      f.n_keys = uint32_t(qd.n_keys);
      f.n_aggs = uint32_t(qd.n_aggs);
      for (int k = 0; k < qd.n_keys; k++) {
        f.stride[k] = qd.keys[k].dense_stride;
        f.radix[k] = c.dense_radix[size_t(k)] ? c.dense_radix[size_t(k)] : 1;
      }
      f.t_rows = qd.t_rows;
      for (int a = 0; a < kMaxAggs; a++) f.t_agg[a] = qd.t_agg[a];
      f.counters = qd.counters;
      {  // what the collective tail needs to fold partial tables itself (layout of table_layout())
        int pos = 0;
        for (int a = 0; a < qd.n_aggs; a++) {
          f.agg_func[a] = qd.aggs[a].func;
          f.agg_is_float[a] = qd.aggs[a].is_float;
          f.agg_pos[a] = qd.aggs[a].func == FGPU_AGG_COUNT ? int8_t(-1) : int8_t(pos++);
        }
      }
      x.out_bytes = 256 + size_t(qd.n_keys) * ((size_t(f.max_out) * 4 + 7) & ~size_t(7)) + size_t(qd.n_aggs) * f.max_out * 8;
      CUDA_TRY(x.out.alloc(256, s));  // working header (k_finalize_dense leaves it zero)
      CUDA_TRY(cudaMemsetAsync(x.out.p, 0, 256, s));
      f.hdr = static_cast<uint32_t*>(x.out.p);
      f.out = nullptr;  // a page-locked block per Execute (cached_tail)
    }
    x.pool = ctx->pinned;
    x.pinned = ctx->pinned_take(x.out_bytes, &x.pinned_cap);
    if (!x.pinned) return fail(FGPU_ERR_OOM, "page-locked memory for the result image");
    if (!res->pending) {
      if (int32_t rct = cached_tail(ctx, x)) return rct;
    } else {
      CUDA_TRY(cudaMemcpyAsync(x.pinned + 32, qd.counters, 32, cudaMemcpyDeviceToHost, s));  // (rows selected by this rank's scan)
    }
    CUDA_TRY(cudaEventRecord(ctx->ev[3], s));
    pc.mark("launch");
    CUDA_TRY(cudaStreamSynchronize(s));  // hostaux stays valid until here
    pc.mark("sync");
    release_staging(ctx);
    st.h2d_bytes += c.h2d_bytes;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
    st.scan_kernel_ms = ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
    st.total_device_ms = ms;
    const unsigned long long* cnts = reinterpret_cast<const unsigned long long*>(x.pinned + 32);
    st.rows_selected = cnts[0];
    x.qd = qd;
    x.qdesc_dev = qdesc_dev;
    x.has_runs = !batches[0].rgs.empty();
    x.runs_v1 = runs_v1;
    x.rd = batches[0].rd;
    x.runs_nl = batches[0].nl;
    x.runs_nk = runs_nk;
    x.runs_na = runs_na;
    x.has_ta = !TA.rgs.empty();
    x.td = TA.td;
    x.keys = c.keys;
    for (KeyOut& k : x.keys) {
      if (k.dict) {
        const uint32_t card = k.dict->cardinality();
        k.dict_snapshot.reserve(card);
        for (uint32_t i = 0; i < card; i++) k.dict_snapshot.push_back(k.dict->value(i));
        k.dict = nullptr;
      }
    }
    x.dense_radix = c.dense_radix;
    for (const KeyOut& k : x.keys) {  // small dictionaries are exported whole: the host then only copies index columns
      std::unique_ptr<OwnedColumn> d;
      if (k.dict_snapshot.size() <= 4096) {
        d = std::make_unique<OwnedColumn>();
        d->format = "z";
        d->length = int64_t(k.dict_snapshot.size());
        d->offsets.push_back(0);
        for (const std::string& v : k.dict_snapshot) {
          d->data.insert(d->data.end(), v.begin(), v.end());
          d->offsets.push_back(int32_t(d->data.size()));
        }
      }
      x.dict_template.push_back(std::move(d));
    }
    for (size_t a = 0; a < q.aggs.size(); a++) {
      x.agg_names.push_back(std::string(agg_string(q.aggs[a].func)) + "(" + q.expr_name(q.aggs[a].expr) + ")");
      x.agg_is_float.push_back(qd.aggs[a].func != FGPU_AGG_COUNT && qd.aggs[a].is_float);
    }
    st.kernel_launches++;  // k_finalize_dense (cached_tail)
    x.stats = st;  // what a cached Execute reports: no uploads, no compile
    x.stats.h2d_bytes = 0;
    x.stats.d2h_bytes = 0;
    x.stats.rows_selected = 0;
    x.ready = true;
    st.d2h_bytes += x.out_bytes;
    res->qd = qd;
    pc.mark("keep");
    pc.flush("scan");
    if (res->pending) {
      res->pending_cached = &x;
      return FGPU_OK;
    }
    return finalize_dense_host(ctx, res, x);
  }
  if (count_groups && !rows_plan) {  // count the result rows behind the scan: one host round trip less
    fd.out_count = res->cnt_ptr;
    fd.max_out = 0;
    fd.out_keys = nullptr;
    fd.out_aggs = nullptr;
    fd.out_rows = nullptr;
    CUDA_TRY(launch_finalize(fd, s));
    st.kernel_launches++;
    res->groups_known = true;
  }
  CUDA_TRY(cudaMemcpyAsync(counters, qd.counters, 64 + 16, cudaMemcpyDeviceToHost, s));
  pc.mark("launch");
  CUDA_TRY(cudaStreamSynchronize(s));  // hostaux / counters stay valid until here
  pc.mark("sync");
  release_staging(ctx);
  st.h2d_bytes += c.h2d_bytes;
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
  st.scan_kernel_ms = ms;
  st.rows_selected = counters[0];
  if (res->groups_known) res->n_groups = static_cast<unsigned int>(counters[8] & 0xffffffffull);
  st.d2h_bytes += 64;
  if (counters[1]) return fail(FGPU_ERR_UNSUPPORTED, "aggregate hash table overflow (more groups than the sized capacity)");
  if (int32_t rcc = comm_check(counters[3])) return rcc;

  // keep what finalize / merge need
  res->qd = qd;
  res->keys = c.keys;  // (the plan may be reused by the next Execute)
  for (KeyOut& k : res->keys) {
    if (k.dict) {
      uint32_t card = k.dict->cardinality();
      k.dict_snapshot.reserve(card);
      for (uint32_t i = 0; i < card; i++) k.dict_snapshot.push_back(k.dict->value(i));
      k.dict = nullptr;
    }
  }
  res->rows_plan = rows_plan;
  for (size_t a = 0; a < q.aggs.size(); a++) {
    res->agg_names.push_back(std::string(agg_string(q.aggs[a].func)) + "(" + q.expr_name(q.aggs[a].expr) + ")");
    // Count yields int64; Sum/Min/Max keep the input type (aggregate.go:734-950)
    res->agg_is_float.push_back(qd.aggs[a].func != FGPU_AGG_COUNT && qd.aggs[a].is_float);
  }
  pc.mark("keep");
  pc.flush("scan");
  return FGPU_OK;
}

// Compacted result image in page-locked memory -> one Arrow record (finishAggregate, aggregate.go:543-633): the
// device already produced dictionary indices and aggregate values per result row (k_finalize_dense); the host copies.
int32_t finalize_dense_host(fgpu_ctx* ctx, fgpu_result* res, ExecCache& x) {
  (void)ctx;
  PhaseClock pc;
  const QueryDesc& qd = x.qd;
  const uint32_t* hdr = reinterpret_cast<const uint32_t*>(x.pinned);
  const size_t G = std::min<size_t>(hdr[0], x.dout.max_out);
  res->stats.groups = G;
  const int nk = qd.n_keys, na = qd.n_aggs;
  const size_t key_bytes = (size_t(x.dout.max_out) * 4 + 7) & ~size_t(7);
  std::vector<OwnedColumn> cols;
  cols.reserve(size_t(nk + na));
  // the columns of the record are slices of the page-locked block the kernel wrote; it goes back to the pool when the
  // consumer has released every array
  std::shared_ptr<void> keep;
  if (G > 0) {
    std::shared_ptr<PinnedPool> pool = x.pool;
    uint8_t* blk = x.pinned;
    const size_t cap = x.pinned_cap;
    keep = std::shared_ptr<void>(blk, [pool, blk, cap](void*) { pool->give(blk, cap); });
  }
  for (int k = 0; k < nk; k++) {
    const KeyOut& ko = x.keys[size_t(k)];
    uint32_t* idx = reinterpret_cast<uint32_t*>(x.pinned + 256 + size_t(k) * key_bytes);
    OwnedColumn col;
    col.name = ko.name;
    col.length = int64_t(G);
    col.format = "I";  // dictionary<uint32, binary>
    if (G > 0) {
      col.ext = reinterpret_cast<const uint8_t*>(idx);
      col.ext_keep = keep;
    }
    if (hdr[16 + k]) {  // NULL keys: validity bitmap, index 0 in the NULL slots
      col.validity.assign((G + 7) / 8, 0);
      for (size_t i = 0; i < G; i++) {
        if (idx[i] == 0xffffffffu) {
          idx[i] = 0;
          col.null_count++;
        } else {
          set_bit(col.validity, int64_t(i));
        }
      }
    }
    auto dict = std::make_unique<OwnedColumn>();
    if (const OwnedColumn* t = x.dict_template[size_t(k)].get()) {
      dict->format = t->format;
      dict->length = t->length;
      dict->data = t->data;
      dict->offsets = t->offsets;
    } else {  // large dictionary: only the values this result uses
      std::vector<uint32_t> remap(ko.dict_snapshot.size(), 0xffffffffu), used;
      const std::vector<uint8_t>& val = col.validity;
      auto is_valid = [&](size_t i) { return val.empty() || ((val[i >> 3] >> (i & 7)) & 1); };
      for (size_t i = 0; i < G; i++)
        if (is_valid(i)) remap[idx[i]] = 0;
      for (size_t g = 0; g < remap.size(); g++)
        if (remap[g] == 0) {
          remap[g] = uint32_t(used.size());
          used.push_back(uint32_t(g));
        }
      for (size_t i = 0; i < G; i++)
        if (is_valid(i)) idx[i] = remap[idx[i]];
      dict->format = "z";
      dict->length = int64_t(used.size());
      dict->offsets.push_back(0);
      for (uint32_t gid : used) {
        const std::string& v = ko.dict_snapshot[gid];
        dict->data.insert(dict->data.end(), v.begin(), v.end());
        dict->offsets.push_back(int32_t(dict->data.size()));
      }
    }
    col.dictionary = std::move(dict);
    cols.push_back(std::move(col));
  }
  const uint8_t* aggs = x.pinned + 256 + size_t(nk) * key_bytes;
  for (int a = 0; a < na; a++) {
    OwnedColumn col;
    col.name = x.agg_names[size_t(a)];
    col.format = x.agg_is_float[size_t(a)] ? "g" : "l";
    col.length = int64_t(G);
    if (G > 0) {
      col.ext = aggs + size_t(a) * x.dout.max_out * 8;
      col.ext_keep = keep;
    }
    cols.push_back(std::move(col));
  }
  res->stats.algorithmic_bytes += (size_t(nk) + size_t(na)) * G * 8;
  if (G > 0) {  // finishAggregate skips empty aggregates (aggregate.go:547-549)
    res->records.push_back(std::move(cols));
    res->record_rows.push_back(int64_t(G));
    x.pinned = nullptr;  // the block now belongs to the record; the next Execute takes another one
    x.pinned_cap = 0;
  }
  res->finalized = true;
  pc.mark("arrow");
  pc.flush("finalize-host");
  return FGPU_OK;
}

// Table -> host columns -> one Arrow record.
// Rows plan: selected rows of the projected columns -> one Arrow record (filter.go:276-323 + Projection).
int32_t finalize_rows(fgpu_ctx* ctx, fgpu_result* res) {
  cudaStream_t s = ctx->stream;
  const QueryDesc& qd = res->qd;
  const size_t R = size_t(res->stats.rows_selected);
  res->stats.groups = R;
  // Every projected column comes back into a page-locked block of the context's pool with ONE copy each and ONE
  // synchronisation for all of them; the block itself becomes the Arrow buffer (it returns to the pool when the
  // consumer releases the array), so a large selection costs its PCIe transfer and nothing else.
  struct Block { uint8_t* p = nullptr; size_t cap = 0; };
  std::shared_ptr<PinnedPool> pool = ctx->pinned;
  auto keep_of = [pool](Block b) { return std::shared_ptr<void>(b.p, [pool, b](void*) { pool->give(b.p, b.cap); }); };
  std::vector<Block> data(size_t(qd.n_out)), valid(size_t(qd.n_out));
  for (int o = 0; o < qd.n_out && R; o++) {
    const bool dict = !res->keys[size_t(o)].is_int64;
    const size_t bytes = R * (dict ? 4 : 8);
    data[size_t(o)].p = pool->take(bytes, &data[size_t(o)].cap);
    if (!data[size_t(o)].p) return fail(FGPU_ERR_OOM, "page-locked memory for the result rows");
    CUDA_TRY(cudaMemcpyAsync(data[size_t(o)].p, qd.out_data[o], bytes, cudaMemcpyDeviceToHost, s));
    res->stats.d2h_bytes += bytes;
    if (!dict && !res->rows_no_nulls) {
      valid[size_t(o)].p = pool->take(R, &valid[size_t(o)].cap);
      if (!valid[size_t(o)].p) return fail(FGPU_ERR_OOM, "page-locked memory for the result rows");
      CUDA_TRY(cudaMemcpyAsync(valid[size_t(o)].p, qd.out_valid[o], R, cudaMemcpyDeviceToHost, s));
      res->stats.d2h_bytes += R;
    }
  }
  CUDA_TRY(cudaEventRecord(ctx->ev[3], s));
  CUDA_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]);
  res->stats.total_device_ms = ms;
  std::vector<OwnedColumn> cols;
  for (int o = 0; o < qd.n_out; o++) {
    const KeyOut& ko = res->keys[size_t(o)];
    OwnedColumn col;
    col.name = ko.name;
    col.length = int64_t(R);
    if (R) {
      col.ext = data[size_t(o)].p;
      col.ext_keep = keep_of(data[size_t(o)]);
    }
    if (!ko.is_int64) {
      // dictionary ids (-1 = NULL): indices into a dictionary of the values this result uses
      col.format = "I";
      int32_t* ids = reinterpret_cast<int32_t*>(data[size_t(o)].p);
      uint32_t* idx = reinterpret_cast<uint32_t*>(data[size_t(o)].p);
      std::vector<int32_t> remap(ko.dict_snapshot.size(), -1);
      bool any_null = false;
      for (size_t i = 0; i < R; i++) {
        if (ids[i] >= 0) remap[size_t(ids[i])] = 0;
        else any_null = true;
      }
      auto dict = std::make_unique<OwnedColumn>();
      dict->format = "z";
      dict->offsets.push_back(0);
      int32_t next = 0;
      for (size_t g = 0; g < remap.size(); g++) {
        if (remap[g] < 0) continue;
        remap[g] = next++;
        const std::string& v = ko.dict_snapshot[g];
        dict->data.insert(dict->data.end(), v.begin(), v.end());
        dict->offsets.push_back(int32_t(dict->data.size()));
      }
      dict->length = next;
      if (any_null) col.validity.assign((R + 7) / 8, 0);
      for (size_t i = 0; i < R; i++) {
        if (ids[i] < 0) { idx[i] = 0; col.null_count++; }
        else { idx[i] = uint32_t(remap[size_t(ids[i])]); if (any_null) set_bit(col.validity, int64_t(i)); }
      }
      col.dictionary = std::move(dict);
    } else {
      col.format = ko.is_float ? "g" : "l";
      if (valid[size_t(o)].p) {
        const uint8_t* v = valid[size_t(o)].p;
        size_t nulls = 0;
        for (size_t i = 0; i < R; i++) nulls += v[i] ? 0 : 1;
        if (nulls) {
          col.validity.assign((R + 7) / 8, 0);
          for (size_t i = 0; i < R; i++)
            if (v[i]) set_bit(col.validity, int64_t(i));
          col.null_count = int64_t(nulls);
        }
        pool->give(valid[size_t(o)].p, valid[size_t(o)].cap);
      }
    }
    res->stats.algorithmic_bytes += R * (ko.is_int64 ? 8 : 4);
    cols.push_back(std::move(col));
  }
  if (R > 0) {  // an empty selection emits no record (filter.go:264-266)
    res->records.push_back(std::move(cols));
    res->record_rows.push_back(int64_t(R));
  }
  res->finalized = true;
  res->table.reset();
  res->aux.reset();
  return FGPU_OK;
}

int32_t finalize_result(fgpu_ctx* ctx, fgpu_result* res) {
  if (res->rows_plan) return finalize_rows(ctx, res);
  PhaseClock pc;
  FinalizeDesc& fd = res->fd;
  cudaStream_t s = ctx->stream;
  // Upper bound of result rows: every slot could be occupied; count first to size the output.
  // (k_finalize with max_out == 0 only counts.)
  DevBuf& cnt = res->cnt;
  unsigned int n_groups = res->n_groups;
  if (!res->groups_known) {
    CUDA_TRY(cnt.alloc(16, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(cnt.p, 0, 16, s));
    res->cnt_ptr = static_cast<unsigned int*>(cnt.p);
    fd.out_count = static_cast<unsigned int*>(cnt.p);
    fd.max_out = 0;
    fd.out_keys = nullptr;
    fd.out_aggs = nullptr;
    fd.out_rows = nullptr;
    CUDA_TRY(launch_finalize(fd, s));
    CUDA_TRY(cudaMemcpyAsync(&n_groups, cnt.p, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    res->stats.kernel_launches += 1;
  }
  res->stats.groups = n_groups;
  const size_t G = n_groups;
  const int nk = fd.n_keys, na = fd.n_aggs;
  // every result column lands in its own page-locked block of the context's pool; int64 keys and aggregates hand
  // that block to the Arrow consumer as it is (returned to the pool on release), dictionary keys are re-indexed
  struct Block { uint8_t* p = nullptr; size_t cap = 0; };
  std::shared_ptr<PinnedPool> pool = ctx->pinned;
  auto keep_of = [pool](Block b) { return std::shared_ptr<void>(b.p, [pool, b](void*) { pool->give(b.p, b.cap); }); };
  std::vector<Block> blocks(size_t(nk + na));
  if (G > 0) {
    DevBuf out;
    size_t bytes = (size_t(nk) + size_t(na) + 1) * G * 8;
    CUDA_TRY(out.alloc(bytes, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(res->cnt_ptr, 0, 16, s));
    fd.out_count = res->cnt_ptr;
    fd.max_out = n_groups;
    fd.out_keys = static_cast<long long*>(out.p);
    fd.out_aggs = fd.out_keys + size_t(nk) * G;
    fd.out_rows = reinterpret_cast<unsigned long long*>(fd.out_aggs + size_t(na) * G);
    CUDA_TRY(launch_finalize(fd, s));
    res->stats.kernel_launches += 1;
    for (int c = 0; c < nk + na; c++) {
      blocks[size_t(c)].p = pool->take(G * 8, &blocks[size_t(c)].cap);
      if (!blocks[size_t(c)].p) return fail(FGPU_ERR_OOM, "page-locked memory for the result columns");
      CUDA_TRY(cudaMemcpyAsync(blocks[size_t(c)].p, fd.out_keys + size_t(c) * G, G * 8, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    res->stats.d2h_bytes += (size_t(nk) + size_t(na)) * G * 8;
  }
  CUDA_TRY(cudaEventRecord(ctx->ev[3], s));
  CUDA_TRY(cudaEventSynchronize(ctx->ev[3]));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]);
  res->stats.total_device_ms = ms;
  pc.mark("finalize_device");

  // ---- build Arrow columns (aggregate.go:551-625: group columns first, then aggregates) --------
  std::vector<OwnedColumn> cols;
  for (int k = 0; k < nk; k++) {
    const KeyOut& ko = res->keys[size_t(k)];
    OwnedColumn col;
    col.name = ko.name;
    col.length = int64_t(G);
    const long long* codes = reinterpret_cast<const long long*>(blocks[size_t(k)].p);
    if (ko.is_int64) {
      col.format = "l";
      if (G) {
        col.ext = blocks[size_t(k)].p;
        col.ext_keep = keep_of(blocks[size_t(k)]);
      }
    } else {
      // dictionary<uint32, binary> holding only the values this result uses
      col.format = "I";
      col.validity.assign((G + 7) / 8, 0);
      col.data.resize(G * 4);
      uint32_t* idx = reinterpret_cast<uint32_t*>(col.data.data());
      // used dictionary ids in increasing order -> dense indices (O(groups + cardinality))
      std::vector<uint32_t> remap(ko.dict_snapshot.size(), 0xffffffffu), used;
      for (size_t i = 0; i < G; i++)
        if (codes[i] != 0) remap[size_t(codes[i] - 1)] = 0;
      for (size_t g = 0; g < remap.size(); g++)
        if (remap[g] == 0) {
          remap[g] = uint32_t(used.size());
          used.push_back(uint32_t(g));
        }
      for (size_t i = 0; i < G; i++) {
        if (codes[i] == 0) {
          idx[i] = 0;
          col.null_count++;
        } else {
          idx[i] = remap[size_t(codes[i] - 1)];
          set_bit(col.validity, int64_t(i));
        }
      }
      if (col.null_count == 0) col.validity.clear();
      auto dict = std::make_unique<OwnedColumn>();
      dict->format = "z";
      dict->length = int64_t(used.size());
      dict->offsets.push_back(0);
      for (uint32_t gid : used) {
        const std::string& v = ko.dict_snapshot[gid];
        dict->data.insert(dict->data.end(), v.begin(), v.end());
        dict->offsets.push_back(int32_t(dict->data.size()));
      }
      col.dictionary = std::move(dict);
      if (G) pool->give(blocks[size_t(k)].p, blocks[size_t(k)].cap);  // the codes are re-indexed into col.data
    }
    cols.push_back(std::move(col));
  }
  for (int a = 0; a < na; a++) {
    OwnedColumn col;
    col.name = res->agg_names[size_t(a)];
    col.format = res->agg_is_float[size_t(a)] ? "g" : "l";
    col.length = int64_t(G);
    if (G) {
      col.ext = blocks[size_t(nk + a)].p;
      col.ext_keep = keep_of(blocks[size_t(nk + a)]);
    }
    cols.push_back(std::move(col));
  }
  res->stats.algorithmic_bytes += (size_t(nk) + size_t(na)) * G * 8;
  if (G > 0) {  // finishAggregate skips empty aggregates (aggregate.go:547-549)
    res->records.push_back(std::move(cols));
    res->record_rows.push_back(int64_t(G));
  }
  res->finalized = true;
  res->table.reset();
  res->aux.reset();
  res->cnt.reset();
  pc.mark("arrow");
  pc.flush("finalize");
  return FGPU_OK;
}


// Second half of a collective Execute: wait for the peers, merge, bring the final result to the host.
int32_t collective_end(fgpu_ctx* ctx, fgpu_result* res) {
  if (!res->pending) return fail(FGPU_ERR_INVALID, "no collective Execute pending on this result");
  res->pending = false;
  cudaStream_t s = ctx->stream;
  if (ExecCache* x = static_cast<ExecCache*>(res->pending_cached)) {
    if (x->qd.table_mode == TM_DENSE) {
      if (int32_t rcm = cached_tail_merged(ctx, *x, res->pending_seq, res->pending_bytes)) return rcm;
    } else {
      int32_t rc = comm_wait_merge(ctx, x->qd, x->table.p, res->pending_seq, res->pending_bytes);
      if (rc) return rc;
      if (int32_t rct = cached_tail(ctx, *x)) return rct;
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    const unsigned long long* counters = reinterpret_cast<const unsigned long long*>(x->pinned + 32);
    res->stats.d2h_bytes += x->out_bytes;
    if (int32_t rcc = comm_check(counters[3])) return rcc;
    return finalize_dense_host(ctx, res, *x);
  }
  QueryDesc& qd = res->qd;
  int32_t rc = comm_wait_merge(ctx, qd, res->table.p, res->pending_seq, res->pending_bytes);
  if (rc) return rc;
  FinalizeDesc& fd = res->fd;
  if (qd.table_mode != TM_DENSE) {  // hash tables: fold every rank's slot into a fresh table
    DevBuf merged;
    CUDA_TRY(merged.alloc(res->table_bytes + 128, s));
    QueryDesc q2 = qd;
    bind_table(&q2, static_cast<uint8_t*>(merged.p));
    CUDA_TRY(launch_table_init(q2, s, /*zero_counters=*/false));
    for (int r = 0; r < ctx->comm.n; r++) CUDA_TRY(launch_merge(q2, ctx->comm.mailbox + ctx->comm.data_off(res->pending_seq & 1u, r), s));
    std::swap(res->table.p, merged.p);
    std::swap(res->table.n, merged.n);
    qd = q2;
    fd.t_rows = qd.t_rows;
    for (int a = 0; a < kMaxAggs; a++) fd.t_agg[a] = qd.t_agg[a];
    fd.t_tag = qd.t_tag;
    fd.t_keys = qd.t_keys;
  }
  if (res->cnt_ptr) {  // count the groups of the final table behind the merge
    CUDA_TRY(cudaMemsetAsync(res->cnt_ptr, 0, 16, s));
    fd.out_count = res->cnt_ptr;
    fd.max_out = 0;
    fd.out_keys = nullptr;
    fd.out_aggs = nullptr;
    fd.out_rows = nullptr;
    CUDA_TRY(launch_finalize(fd, s));
    res->stats.kernel_launches++;
  }
  CUDA_TRY(ctx->ensure_scratch(128));
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(ctx->scratch);
  CUDA_TRY(cudaMemcpyAsync(counters, qd.counters, 64 + 16, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (counters[1]) return fail(FGPU_ERR_UNSUPPORTED, "aggregate hash table overflow while merging partials");
  if (int32_t rcc = comm_check(counters[3])) return rcc;
  if (res->cnt_ptr) {
    res->n_groups = static_cast<unsigned int>(counters[8] & 0xffffffffull);
    res->groups_known = true;
  }
  return finalize_result(ctx, res);
}

}  // namespace

// =====================================================================================================
// extern "C"
// =====================================================================================================
// No C++ exception crosses the C ABI: every entry point that returns a status runs inside this guard.
#define API_TRY try {
#define API_CATCH                                                                      \
  }                                                                                    \
  catch (const std::bad_alloc&) { return fail(FGPU_ERR_OOM, "out of host memory"); }   \
  catch (const std::exception& ex) { return fail(FGPU_ERR_INVALID, std::string("internal error: ") + ex.what()); } \
  catch (...) { return fail(FGPU_ERR_INVALID, "internal error"); }

extern "C" {

const char* fgpu_last_error(void) { return g_err.c_str(); }
int32_t fgpu_abi_version(void) { return FGPU_ABI_VERSION; }

int32_t fgpu_init(const fgpu_config* cfg, fgpu_ctx** out) {
  API_TRY
  if (!cfg || !out) return fail(FGPU_ERR_INVALID, "null argument");
  if (cfg->abi_version != FGPU_ABI_VERSION) return fail(FGPU_ERR_INVALID, "ABI version mismatch");
  if (cfg->tile_rows != 0 && cfg->tile_rows != kTileRows) return fail(FGPU_ERR_INVALID, "tile_rows must be 0 (default)");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(FGPU_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  if (cfg->device < 0 || cfg->device >= n) return fail(FGPU_ERR_INVALID, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(cfg->device));
  auto ctx = std::make_unique<fgpu_ctx>();
  ctx->device = cfg->device;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  ctx->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  {
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, cfg->device));
    uint64_t keep = ~0ull;  // freed blocks stay in the pool: allocation is a pointer bump afterwards
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  for (auto& ev : ctx->ev) CUDA_TRY(cudaEventCreate(&ev));
  *out = ctx.release();
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_shutdown(fgpu_ctx* ctx) {
  API_TRY
  if (!ctx) return FGPU_OK;
  cudaSetDevice(ctx->device);
  if (ctx->comm.mailbox) {
    for (int r = 0; r < ctx->comm.n; r++)
      if (ctx->comm.peer_ipc[r] && ctx->comm.peer[r]) cudaIpcCloseMemHandle(ctx->comm.peer[r]);
    cudaFree(ctx->comm.mailbox);
    ctx->comm = Comm{};
  }
  for (auto& kv : ctx->plans) kv.second->plan_cache.reset();  // device state of compiled plans (queries still held by the caller stay valid handles)
  ctx->plans.clear();
  for (auto& t : ctx->tables)
    for (auto& p : t.second.parts) free_part(ctx, p.get());
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->scratch) cudaFreeHost(ctx->scratch);
  if (ctx->arena) cudaFreeHost(ctx->arena);
  delete ctx;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_part_put_parquet(fgpu_ctx* ctx, const char* table, uint64_t part_id, uint64_t tx, const uint8_t* file,
                              uint64_t len, int32_t flags) {
  API_TRY
  if (!ctx || !table || !file) return fail(FGPU_ERR_INVALID, "null argument");
  if (flags != FGPU_PUT_DEFAULT && flags != FGPU_PUT_BORROW_PINNED) return fail(FGPU_ERR_INVALID, "unknown put flags");
  std::lock_guard<std::mutex> lk(ctx->mu);
  PhaseClock pc;
  struct PutTail { PhaseClock& pc; ~PutTail() { pc.mark("put"); pc.flush("put"); } } put_tail{pc};
  CUDA_TRY(cudaSetDevice(ctx->device));
  Table& t = ctx->tables[table];
  for (auto& p : t.parts)
    if (p->id == part_id) return fail(FGPU_ERR_INVALID, "part id already registered");
  auto part = std::make_unique<Part>();
  part->id = part_id;
  part->tx = tx;
  part->borrowed = flags == FGPU_PUT_BORROW_PINNED;
  std::string err;
  if (!open_part(file, len, part.get(), &err)) return fail(FGPU_ERR_PARQUET, err);
  if (!part->borrowed) {
    // The caller may reuse the buffer on return: every column goes to the device now (host images built in
    // parallel, one worker per column).
    prebuild_columns(&t, part.get(), part->columns);
    for (const std::string& col : part->columns) {
      int32_t rc = ensure_resident(ctx, &t, part.get(), col, nullptr);
      if (rc) {
        free_part(ctx, part.get());
        return rc;
      }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    release_staging(ctx);
    if (e != cudaSuccess) {
      free_part(ctx, part.get());
      return fail(FGPU_ERR_CUDA, std::string("part upload: ") + cudaGetErrorString(e));
    }
    part->file = nullptr;
  }
  t.parts.push_back(std::move(part));
  t.epoch = ++ctx->epoch_counter;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_part_put_arrow(fgpu_ctx* ctx, const char* table, uint64_t part_id, uint64_t tx, struct ArrowSchema* schema,
                            struct ArrowArray* array) {
  API_TRY
  // the library owns both structs from here on, whatever happens
  struct Release {
    ArrowSchema* s; ArrowArray* a;
    ~Release() {
      if (s && s->release) s->release(s);
      if (a && a->release) a->release(a);
    }
  } rel{schema, array};
  if (!ctx || !table || !schema || !array) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  Table& t = ctx->tables[table];
  for (auto& p : t.parts)
    if (p->id == part_id) return fail(FGPU_ERR_INVALID, "part id already registered");
  auto part = std::make_unique<Part>();
  part->id = part_id;
  part->tx = tx;
  std::string err;
  part->arrow = true;
  if (!build_arrow_part(&t, part.get(), schema, array, &err)) return fail(FGPU_ERR_UNSUPPORTED, "Arrow part: " + err);
  // the record's buffers go away on return: every column image is uploaded now
  for (const std::string& col : part->columns) {
    int32_t rc = ensure_resident(ctx, &t, part.get(), col, nullptr);
    if (rc) {
      free_part(ctx, part.get());
      return rc;
    }
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  release_staging(ctx);
  if (e != cudaSuccess) {
    free_part(ctx, part.get());
    return fail(FGPU_ERR_CUDA, std::string("part upload: ") + cudaGetErrorString(e));
  }
  t.parts.push_back(std::move(part));
  t.epoch = ++ctx->epoch_counter;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_part_drop(fgpu_ctx* ctx, const char* table, uint64_t part_id) {
  API_TRY
  if (!ctx || !table) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->tables.find(table);
  if (it == ctx->tables.end()) return fail(FGPU_ERR_NOT_FOUND, std::string("table not found: ") + table);
  auto& parts = it->second.parts;
  for (size_t i = 0; i < parts.size(); i++) {
    if (parts[i]->id == part_id) {
      cudaSetDevice(ctx->device);
      cudaStreamSynchronize(ctx->stream);
      release_staging(ctx);
      free_part(ctx, parts[i].get());
      parts.erase(parts.begin() + long(i));
      it->second.epoch = ++ctx->epoch_counter;
      return FGPU_OK;
    }
  }
  return fail(FGPU_ERR_NOT_FOUND, "part not found");
  API_CATCH
}

int32_t fgpu_table_drop(fgpu_ctx* ctx, const char* table) {
  API_TRY
  if (!ctx || !table) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->tables.find(table);
  if (it == ctx->tables.end()) return FGPU_OK;
  PhaseClock pc;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  release_staging(ctx);
  for (auto& p : it->second.parts) free_part(ctx, p.get());
  ctx->tables.erase(it);
  pc.mark("drop");
  pc.flush("drop");
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_query_prepare(fgpu_ctx* ctx, const fgpu_plan* plan, fgpu_query** out) {
  API_TRY
  if (!ctx || !plan || !out || !plan->table) return fail(FGPU_ERR_INVALID, "null argument");
  if (plan->kind != FGPU_PLAN_AGGREGATE && plan->kind != FGPU_PLAN_DISTINCT && plan->kind != FGPU_PLAN_FILTER)
    return fail(FGPU_ERR_INVALID, "unknown plan kind");
  auto q = std::make_shared<QueryPlan>();
  q->ctx = ctx;
  q->table = plan->table;
  q->kind = plan->kind;
  q->filter = plan->filter;
  // the plan's text: two prepared queries with the same signature share one QueryPlan (and its compiled state)
  std::string sig;
  auto put = [&](const void* p, size_t n) { sig.append(static_cast<const char*>(p), n); };
  auto put_str = [&](const std::string& v) { const uint64_t n = v.size(); put(&n, 8); sig += v; };
  put_str(q->table);
  put(&plan->kind, 4);
  put(&plan->filter, 4);
  auto check = [&](int32_t i) { return i >= 0 && i < plan->n_exprs; };
  for (int32_t i = 0; i < plan->n_exprs; i++) {
    const fgpu_expr& e = plan->exprs[i];
    ExprNode n;
    n.kind = e.kind;
    n.op = e.op;
    n.left = e.left;
    n.right = e.right;
    put(&e.kind, 4); put(&e.op, 4); put(&e.left, 4); put(&e.right, 4);
    if (e.kind == FGPU_EXPR_COLUMN || e.kind == FGPU_EXPR_DYNCOLUMN) {
      if (!e.name) return fail(FGPU_ERR_INVALID, "column expression without a name");
      n.name = e.name;
      put_str(n.name);
    } else if (e.kind == FGPU_EXPR_LITERAL) {
      n.lit_type = e.literal.type;
      n.lit_i = e.literal.i64;
      n.lit_f = e.literal.f64;
      if (e.literal.type == FGPU_SCALAR_STRING) {
        if (e.literal.len && !e.literal.bytes) return fail(FGPU_ERR_INVALID, "string literal without bytes");
        n.lit_bytes.assign(reinterpret_cast<const char*>(e.literal.bytes), size_t(e.literal.len));
      }
      put(&n.lit_type, 4); put(&n.lit_i, 8); put(&n.lit_f, 8);
      put_str(n.lit_bytes);
    } else if (e.kind == FGPU_EXPR_BINARY) {
      if (!check(e.left) || !check(e.right) || e.left >= i || e.right >= i)
        return fail(FGPU_ERR_INVALID, "binary expression children must precede their parent");
      n.match = e.match;
      n.match_user = e.match_user;
      const void* fn = reinterpret_cast<const void*>(e.match);
      put(&fn, sizeof fn); put(&e.match_user, sizeof e.match_user);  // a different matcher is a different plan
    } else {
      return fail(FGPU_ERR_INVALID, "unknown expression kind");
    }
    q->exprs.push_back(std::move(n));
  }
  if (plan->filter >= 0 && !check(plan->filter)) return fail(FGPU_ERR_INVALID, "filter index out of range");
  for (int32_t i = 0; i < plan->n_group_by; i++) {
    if (!check(plan->group_by[i])) return fail(FGPU_ERR_INVALID, "group-by index out of range");
    q->group_by.push_back(plan->group_by[i]);
    put(&plan->group_by[i], 4);
  }
  if (plan->kind != FGPU_PLAN_AGGREGATE && plan->n_aggs != 0) return fail(FGPU_ERR_INVALID, "DISTINCT / FILTER plan with aggregates");
  if (plan->kind == FGPU_PLAN_AGGREGATE && plan->n_aggs == 0) return fail(FGPU_ERR_INVALID, "AGGREGATE plan without aggregates");
  sig += "|";
  for (int32_t i = 0; i < plan->n_aggs; i++) {
    if (!check(plan->aggs[i].expr)) return fail(FGPU_ERR_INVALID, "aggregate expression index out of range");
    q->aggs.push_back(plan->aggs[i]);
    put(&plan->aggs[i].func, 4); put(&plan->aggs[i].expr, 4);
  }
  auto handle = std::make_unique<fgpu_query>();
  handle->ctx = ctx;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->plans.find(sig);
    if (it != ctx->plans.end()) {
      handle->plan = it->second;
    } else {
      if (ctx->plans.size() >= 256) {  // bounded: forget the plans nobody holds any more, then everything
        CUDA_TRY(cudaSetDevice(ctx->device));
        for (auto p = ctx->plans.begin(); p != ctx->plans.end();) p = p->second.use_count() == 1 ? ctx->plans.erase(p) : std::next(p);
        if (ctx->plans.size() >= 256) ctx->plans.clear();
      }
      ctx->plans.emplace(std::move(sig), q);
      handle->plan = std::move(q);
    }
  }
  *out = handle.release();
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_query_free(fgpu_query* q) {
  API_TRY
  if (q && q->ctx) {  // the last reference to a plan may own device state
    std::lock_guard<std::mutex> lk(q->ctx->mu);
    cudaSetDevice(q->ctx->device);
    q->plan.reset();
  }
  delete q;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_query_execute(fgpu_ctx* ctx, fgpu_query* q, uint64_t tx_watermark, fgpu_result** out) {
  API_TRY
  if (!ctx || !q || !out) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  auto res = std::make_unique<fgpu_result>();
  res->ctx = ctx;
  int32_t rc = run_scan(ctx, *q->plan, tx_watermark, res.get(), /*count_groups=*/true, /*cacheable=*/true);
  if (rc) return rc;
  if (!res->finalized) rc = finalize_result(ctx, res.get());
  if (rc) return rc;
  *out = res.release();
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_query_execute_partial(fgpu_ctx* ctx, fgpu_query* q, uint64_t tx_watermark, fgpu_result** out, void** dev_ptr,
                                   uint64_t* nbytes) {
  API_TRY
  if (!ctx || !q || !out || !dev_ptr || !nbytes) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  auto res = std::make_unique<fgpu_result>();
  res->ctx = ctx;
  if (q->plan->kind == FGPU_PLAN_FILTER) return fail(FGPU_ERR_INVALID, "rows plans have no partial table: execute them per rank");
  int32_t rc = run_scan(ctx, *q->plan, tx_watermark, res.get());
  if (rc) return rc;
  *dev_ptr = res->table.p;
  *nbytes = res->table_bytes;
  *out = res.release();
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_result_merge_partials(fgpu_ctx* ctx, fgpu_result* r, const void* gathered, uint64_t nbytes, int32_t n) {
  API_TRY
  if (!ctx || !r) return fail(FGPU_ERR_INVALID, "null argument");
  if (r->finalized) return fail(FGPU_ERR_INVALID, "result already finalised");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (n > 0) {
    if (!gathered) return fail(FGPU_ERR_INVALID, "null gathered buffer");
    if (nbytes != r->table_bytes) return fail(FGPU_ERR_INVALID, "partial table size differs between ranks (dictionaries not preloaded identically?)");
    // Start from an empty table and fold every rank's partial in, this rank's own included.
    DevBuf merged;
    CUDA_TRY(merged.alloc(r->table_bytes, ctx->stream));
    QueryDesc qd = r->qd;
    bind_table(&qd, static_cast<uint8_t*>(merged.p));
    CUDA_TRY(launch_table_init(qd, ctx->stream));
    for (int32_t i = 0; i < n; i++) {
      CUDA_TRY(launch_merge(qd, static_cast<const uint8_t*>(gathered) + size_t(i) * nbytes, ctx->stream));
      r->stats.kernel_launches++;
    }
    // the result now lives in the merged table; its groups are counted right behind the merges so that
    // finalisation needs one host round trip less
    FinalizeDesc& fd = r->fd;
    fd.t_rows = qd.t_rows;
    for (int a = 0; a < kMaxAggs; a++) fd.t_agg[a] = qd.t_agg[a];
    fd.t_tag = qd.t_tag;
    fd.t_keys = qd.t_keys;
    if (r->cnt_ptr) {
      CUDA_TRY(cudaMemsetAsync(r->cnt_ptr, 0, 16, ctx->stream));
      fd.out_count = r->cnt_ptr;
      fd.max_out = 0;
      fd.out_keys = nullptr;
      fd.out_aggs = nullptr;
      fd.out_rows = nullptr;
      CUDA_TRY(launch_finalize(fd, ctx->stream));
      r->stats.kernel_launches++;
    }
    CUDA_TRY(ctx->ensure_scratch(128));
    unsigned long long* counters = reinterpret_cast<unsigned long long*>(ctx->scratch);
    CUDA_TRY(cudaMemcpyAsync(counters, qd.counters, 64 + 16, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (counters[1]) return fail(FGPU_ERR_UNSUPPORTED, "aggregate hash table overflow while merging partials");
    if (r->cnt_ptr) {
      r->n_groups = static_cast<unsigned int>(counters[8] & 0xffffffffull);
      r->groups_known = true;
    }
    std::swap(r->table.p, merged.p);
    std::swap(r->table.n, merged.n);
    r->qd = qd;
  }
  return finalize_result(ctx, r);
  API_CATCH
}

int32_t fgpu_comm_export(fgpu_ctx* ctx, int32_t rank, int32_t n_ranks, uint64_t slot_bytes, uint8_t* handle) {
  API_TRY
  if (!ctx || !handle) return fail(FGPU_ERR_INVALID, "null argument");
  if (n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks) return fail(FGPU_ERR_INVALID, "rank / n_ranks out of range");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  Comm& cm = ctx->comm;
  if (cm.mailbox) return fail(FGPU_ERR_INVALID, "communicator already exported: fgpu_comm_close first");
  cm.rank = rank;
  cm.n = n_ranks;
  cm.slot_bytes = (std::max<uint64_t>(slot_bytes, 4096) + 255) & ~uint64_t(255);
  cm.total = kCommDataOff + 2 * uint64_t(n_ranks) * cm.slot_bytes;
  cm.seq = 0;
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, cm.total));  // (cudaMalloc, not the stream-ordered pool: the allocation is exported through CUDA IPC)
  cm.mailbox = static_cast<uint8_t*>(p);
  CUDA_TRY(cudaMemset(cm.mailbox, 0, kCommDataOff));
  CommHandle h{};
  h.magic = kCommMagic;
  h.device = ctx->device;
  h.pid = uint64_t(getpid());
  h.ptr = uint64_t(reinterpret_cast<uintptr_t>(cm.mailbox));
  h.total = cm.total;
  h.slot_bytes = cm.slot_bytes;
  h.n = n_ranks;
  h.rank = rank;
  CUDA_TRY(cudaIpcGetMemHandle(&h.ipc, cm.mailbox));
  std::memset(handle, 0, FGPU_COMM_HANDLE_BYTES);
  std::memcpy(handle, &h, sizeof h);
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_comm_open(fgpu_ctx* ctx, const uint8_t* all_handles) {
  API_TRY
  if (!ctx || !all_handles) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  Comm& cm = ctx->comm;
  if (!cm.mailbox) return fail(FGPU_ERR_INVALID, "fgpu_comm_export first");
  if (cm.open) return fail(FGPU_ERR_INVALID, "communicator already open");
  for (int r = 0; r < cm.n; r++) {
    CommHandle h;
    std::memcpy(&h, all_handles + size_t(r) * FGPU_COMM_HANDLE_BYTES, sizeof h);
    if (h.magic != kCommMagic || h.rank != r || h.n != cm.n || h.slot_bytes != cm.slot_bytes)
      return fail(FGPU_ERR_INVALID, "handle " + std::to_string(r) + " does not belong to this communicator (rank order, n_ranks and slot_bytes must agree)");
    if (r == cm.rank) {
      cm.peer[r] = cm.mailbox;
      cm.peer_ipc[r] = false;
    } else if (h.pid == uint64_t(getpid())) {  // another context of this process: plain peer access
      if (h.device != ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          cudaGetLastError();
          return fail(FGPU_ERR_CUDA, std::string("peer access to device ") + std::to_string(h.device) + ": " + cudaGetErrorString(e));
        }
        cudaGetLastError();
      }
      cm.peer[r] = reinterpret_cast<uint8_t*>(uintptr_t(h.ptr));
      cm.peer_ipc[r] = false;
    } else {
      void* p = nullptr;
      CUDA_TRY(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
      cm.peer[r] = static_cast<uint8_t*>(p);
      cm.peer_ipc[r] = true;
    }
  }
  cm.open = true;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_comm_close(fgpu_ctx* ctx) {
  API_TRY
  if (!ctx) return FGPU_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  Comm& cm = ctx->comm;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < cm.n; r++)
    if (cm.peer_ipc[r] && cm.peer[r]) cudaIpcCloseMemHandle(cm.peer[r]);
  if (cm.mailbox) cudaFree(cm.mailbox);
  cudaGetLastError();
  cm = Comm{};
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_query_execute_collective(fgpu_ctx* ctx, fgpu_query* q, uint64_t tx_watermark, fgpu_result** out) {
  API_TRY
  if (!ctx || !q || !out) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  auto res = std::make_unique<fgpu_result>();
  res->ctx = ctx;
  // a cached plan: scan -> push -> wait -> merge -> one copy back, one synchronisation; the first run of a plan
  // pushes, synchronises once more for its own bookkeeping and then runs the merge half
  int32_t rc = run_scan(ctx, *q->plan, tx_watermark, res.get(), /*count_groups=*/false, /*cacheable=*/true, /*collective=*/1);
  if (rc) return rc;
  if (res->pending) rc = collective_end(ctx, res.get());
  if (rc) return rc;
  *out = res.release();
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_query_execute_collective_begin(fgpu_ctx* ctx, fgpu_query* q, uint64_t tx_watermark, fgpu_result** out) {
  API_TRY
  if (!ctx || !q || !out) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  auto res = std::make_unique<fgpu_result>();
  res->ctx = ctx;
  int32_t rc = run_scan(ctx, *q->plan, tx_watermark, res.get(), /*count_groups=*/false, /*cacheable=*/true, /*collective=*/2);
  if (rc) return rc;
  res->plan_keep = q->plan->plan_cache;
  *out = res.release();
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_query_execute_collective_end(fgpu_ctx* ctx, fgpu_result* r) {
  API_TRY
  if (!ctx || !r) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  return collective_end(ctx, r);
  API_CATCH
}

int32_t fgpu_result_partial_is_additive(const fgpu_result* r, int32_t* out) {
  API_TRY
  if (!r || !out) return fail(FGPU_ERR_INVALID, "null argument");
  bool add = !r->rows_plan && r->qd.table_mode == TM_DENSE;
  for (int a = 0; a < r->qd.n_aggs && add; a++)
    if (r->qd.aggs[a].func != FGPU_AGG_COUNT && !(r->qd.aggs[a].func == FGPU_AGG_SUM && !r->qd.aggs[a].is_float)) add = false;
  *out = add ? 1 : 0;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_rowgroup_leaf_mode(int32_t op, int64_t literal, int32_t has_bounds, int64_t min_value, int64_t max_value, int64_t null_count,
                                int64_t num_values, int32_t* out_mode) {
  API_TRY
  if (!out_mode) return fail(FGPU_ERR_INVALID, "null argument");
  if (op < FGPU_OP_EQ || op > FGPU_OP_GT_EQ) return fail(FGPU_ERR_INVALID, "not a comparison operator");
  int64_t lo, hi;
  bool neg;
  canonical_int_range(op, literal, &lo, &hi, &neg);
  *out_mode = chunk_leaf_mode(lo, hi, neg, has_bounds != 0, min_value, max_value, null_count, num_values);
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_xxhash64(const uint8_t* data, uint64_t len, uint64_t seed, uint64_t* out) {
  API_TRY
  if ((!data && len) || !out) return fail(FGPU_ERR_INVALID, "null argument");
  *out = xxhash64(data, size_t(len), seed);
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_bloom_check(const uint8_t* bitset, uint64_t nbytes, uint64_t hash, int32_t* out_may_contain) {
  API_TRY
  if (!bitset || !out_may_contain) return fail(FGPU_ERR_INVALID, "null argument");
  if (nbytes < 32 || nbytes % 32 != 0 || nbytes > 0xffffffffull) return fail(FGPU_ERR_INVALID, "a split-block filter is a multiple of 32 bytes");
  *out_may_contain = sbbf_check(bitset, uint32_t(nbytes), hash) ? 1 : 0;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_bloom_insert(uint8_t* bitset, uint64_t nbytes, uint64_t hash) {
  API_TRY
  if (!bitset) return fail(FGPU_ERR_INVALID, "null argument");
  if (nbytes < 32 || nbytes % 32 != 0 || nbytes > 0xffffffffull) return fail(FGPU_ERR_INVALID, "a split-block filter is a multiple of 32 bytes");
  sbbf_insert(bitset, uint32_t(nbytes), hash);
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_parquet_rowgroup_may_match_eq(const uint8_t* file, uint64_t len, int32_t row_group, const char* column, int32_t lit_type,
                                           int64_t lit_i64, double lit_f64, const uint8_t* lit_bytes, uint64_t lit_len, int32_t* out) {
  API_TRY
  if (!file || !column || !out) return fail(FGPU_ERR_INVALID, "null argument");
  Part part;
  std::string err;
  if (!open_part(file, len, &part, &err)) return fail(FGPU_ERR_PARQUET, err);
  if (row_group < 0 || size_t(row_group) >= part.rgs.size()) return fail(FGPU_ERR_INVALID, "no such row group");
  QueryPlan q;
  ExprNode col, lit, eq;
  col.kind = FGPU_EXPR_COLUMN;
  col.name = column;
  lit.kind = FGPU_EXPR_LITERAL;
  lit.lit_type = lit_type;
  lit.lit_i = lit_i64;
  lit.lit_f = lit_f64;
  if (lit_bytes && lit_len) lit.lit_bytes.assign(reinterpret_cast<const char*>(lit_bytes), size_t(lit_len));
  eq.kind = FGPU_EXPR_BINARY;
  eq.op = FGPU_OP_EQ;
  eq.left = 0;
  eq.right = 1;
  q.exprs = {col, lit, eq};
  *out = rg_may_match(q, 2, part.rgs[size_t(row_group)]) ? 1 : 0;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_result_next(fgpu_result* r, struct ArrowSchema* out_schema, struct ArrowArray* out_array) {
  API_TRY
  if (!r || !out_schema || !out_array) return fail(FGPU_ERR_INVALID, "null argument");
  if (!r->finalized) return fail(FGPU_ERR_INVALID, r->pending ? "collective Execute pending: call fgpu_query_execute_collective_end first"
                                                               : "partial result: call fgpu_result_merge_partials first");
  if (r->next >= r->records.size()) return fail(FGPU_ERR_END, "no more records");
  export_record(std::move(r->records[r->next]), r->record_rows[r->next], out_schema, out_array);
  r->next++;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_result_stats(const fgpu_result* r, fgpu_stats* out) {
  API_TRY
  if (!r || !out) return fail(FGPU_ERR_INVALID, "null argument");
  *out = r->stats;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_result_free(fgpu_result* r) {
  API_TRY
  if (r) {
    if (r->ctx) cudaSetDevice(r->ctx->device);
    delete r;
  }
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_host_alloc(uint64_t bytes, void** out) {
  API_TRY
  if (!out) return fail(FGPU_ERR_INVALID, "null argument");
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? FGPU_ERR_OOM : FGPU_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  }
  *out = p;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_host_free(void* p) {
  API_TRY
  if (p) cudaFreeHost(p);
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_dict_export(fgpu_ctx* ctx, const char* table, const char* column, uint8_t* buf, uint64_t cap,
                         uint64_t* out_len, uint32_t* out_count) {
  API_TRY
  if (!ctx || !table || !column || !out_len || !out_count) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->tables.find(table);
  *out_len = 0;
  *out_count = 0;
  if (it == ctx->tables.end()) return FGPU_OK;
  auto dit = it->second.dicts.find(column);
  if (dit == it->second.dicts.end()) return FGPU_OK;
  const GlobalDict& d = dit->second;
  uint64_t need = 0;
  for (auto& v : d.values) need += 4 + v.size();
  *out_len = need;
  *out_count = uint32_t(d.values.size());
  if (!buf) return FGPU_OK;
  if (cap < need) return fail(FGPU_ERR_INVALID, "buffer too small");
  uint8_t* p = buf;
  for (auto& v : d.values) {
    uint32_t l = uint32_t(v.size());
    std::memcpy(p, &l, 4);
    p += 4;
    std::memcpy(p, v.data(), v.size());
    p += v.size();
  }
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_dict_preload(fgpu_ctx* ctx, const char* table, const char* column, const uint8_t* blob, uint64_t len, uint32_t count) {
  API_TRY
  if (!ctx || !table || !column || (!blob && len)) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  GlobalDict& d = ctx->tables[table].dicts[column];
  d.preloaded = true;
  const uint8_t* p = blob;
  const uint8_t* end = blob + len;
  for (uint32_t i = 0; i < count; i++) {
    if (end - p < 4) return fail(FGPU_ERR_INVALID, "dictionary blob truncated");
    uint32_t l;
    std::memcpy(&l, p, 4);
    p += 4;
    if (l > uint64_t(end - p)) return fail(FGPU_ERR_INVALID, "dictionary blob truncated");
    d.intern(reinterpret_cast<const char*>(p), l);
    p += l;
  }
  ctx->tables[table].epoch = ++ctx->epoch_counter;
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_parquet_dict_values(const uint8_t* file, uint64_t len, const char* column, uint8_t* buf, uint64_t cap, uint64_t* out_len,
                                 uint32_t* out_count) {
  API_TRY
  if (!file || !column || !out_len || !out_count) return fail(FGPU_ERR_INVALID, "null argument");
  ParsedFile pf;
  std::string err;
  if (!parse_parquet(file, len, &pf, &err)) return fail(FGPU_ERR_PARQUET, err);
  GlobalDict d;  // first-seen order over the file's row groups
  for (size_t c = 0; c < pf.leaves.size(); c++) {
    if (pf.leaves[c].name != column) continue;
    for (const RowGroupMeta& rg : pf.row_groups) {
      const ChunkMeta& cm = rg.chunks[c];
      if (!cm.error.empty()) return fail(FGPU_ERR_PARQUET, std::string(column) + ": " + cm.error);
      const uint8_t* p = cm.dict;
      const uint8_t* end = cm.dict + cm.dict_len;
      for (uint32_t i = 0; i < cm.dict_num_values; i++) {
        if (end - p < 4) return fail(FGPU_ERR_PARQUET, "dictionary page truncated");
        uint32_t l;
        std::memcpy(&l, p, 4);
        p += 4;
        if (l > uint64_t(end - p)) return fail(FGPU_ERR_PARQUET, "dictionary entry overruns page");
        d.intern(reinterpret_cast<const char*>(p), l);
        p += l;
      }
    }
  }
  uint64_t need = 0;
  for (auto& v : d.values) need += 4 + v.size();
  *out_len = need;
  *out_count = uint32_t(d.values.size());
  if (!buf) return FGPU_OK;
  if (cap < need) return fail(FGPU_ERR_INVALID, "buffer too small");
  uint8_t* o = buf;
  for (auto& v : d.values) {
    uint32_t l = uint32_t(v.size());
    std::memcpy(o, &l, 4);
    o += 4;
    std::memcpy(o, v.data(), v.size());
    o += v.size();
  }
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_part_decode_column(fgpu_ctx* ctx, const char* table, uint64_t part_id, const char* column,
                                struct ArrowSchema* out_schema, struct ArrowArray* out_array) {
  API_TRY
  if (!ctx || !table || !column || !out_schema || !out_array) return fail(FGPU_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  auto it = ctx->tables.find(table);
  if (it == ctx->tables.end()) return fail(FGPU_ERR_NOT_FOUND, std::string("table not found: ") + table);
  Part* part = nullptr;
  for (auto& p : it->second.parts)
    if (p->id == part_id) part = p.get();
  if (!part) return fail(FGPU_ERR_NOT_FOUND, "part not found");
  {
    int32_t rc = ensure_resident(ctx, &it->second, part, column, nullptr);
    if (rc) return rc;
    const ColumnImage& img = part->images[column];
    if (!img.error.empty() && img.error != "column not in part") return fail(FGPU_ERR_UNSUPPORTED, std::string(column) + ": " + img.error);
  }
  uint64_t total = 0;
  int phys = -1;
  for (auto& rg : part->rgs) {
    auto c = rg.cols.find(column);
    if (c == rg.cols.end()) return fail(FGPU_ERR_NOT_FOUND, std::string("column not found: ") + column);
    if (!c->second.error.empty()) return fail(FGPU_ERR_UNSUPPORTED, std::string(column) + ": " + c->second.error);
    phys = c->second.phys;
    total += rg.n_rows;
  }
  const bool is_str = phys == PT_BYTE_ARRAY;
  DevBuf d_vals, d_valid;
  CUDA_TRY(d_vals.alloc(total * (is_str ? 4 : 8), ctx->stream));
  if (!is_str) CUDA_TRY(d_valid.alloc(total, ctx->stream));
  uint64_t row = 0;
  for (auto& rg : part->rgs) {
    const ChunkDesc& cd = rg.cols.at(column).desc;
    CUDA_TRY(launch_decode(cd, is_str ? static_cast<int32_t*>(d_vals.p) + row : nullptr,
                           is_str ? nullptr : static_cast<long long*>(d_vals.p) + row,
                           is_str ? nullptr : static_cast<uint8_t*>(d_valid.p) + row, ctx->sm_count, ctx->stream));
    row += rg.n_rows;
  }
  OwnedColumn col;
  col.name = column;
  col.length = int64_t(total);
  if (is_str) {
    std::vector<int32_t> ids(total);
    CUDA_TRY(cudaMemcpyAsync(ids.data(), d_vals.p, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const GlobalDict& gd = it->second.dicts.at(column);
    col.format = "I";
    col.validity.assign((total + 7) / 8, 0);
    col.data.resize(total * 4);
    uint32_t* idx = reinterpret_cast<uint32_t*>(col.data.data());
    for (uint64_t i = 0; i < total; i++) {
      if (ids[i] < 0) { idx[i] = 0; col.null_count++; }
      else { idx[i] = uint32_t(ids[i]); set_bit(col.validity, int64_t(i)); }
    }
    if (col.null_count == 0) col.validity.clear();
    auto dict = std::make_unique<OwnedColumn>();
    dict->format = "z";
    dict->length = int64_t(gd.cardinality());
    dict->offsets.push_back(0);
    for (uint32_t g = 0; g < gd.cardinality(); g++) {
      const std::string& v = gd.value(g);
      dict->data.insert(dict->data.end(), v.begin(), v.end());
      dict->offsets.push_back(int32_t(dict->data.size()));
    }
    col.dictionary = std::move(dict);
  } else {
    col.format = phys == PT_DOUBLE ? "g" : "l";
    col.data.resize(total * 8);
    std::vector<uint8_t> valid(total);
    CUDA_TRY(cudaMemcpyAsync(col.data.data(), d_vals.p, total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(valid.data(), d_valid.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    col.validity.assign((total + 7) / 8, 0);
    for (uint64_t i = 0; i < total; i++) {
      if (valid[i]) set_bit(col.validity, int64_t(i));
      else col.null_count++;
    }
    if (col.null_count == 0) col.validity.clear();
  }
  release_staging(ctx);
  export_column(std::move(col), out_schema, out_array);
  return FGPU_OK;
  API_CATCH
}

int32_t fgpu_parquet_describe(const uint8_t* file, uint64_t len, int32_t tile_rows, char* buf, uint64_t cap,
                              uint64_t* out_len) {
  API_TRY
  if (!file || !out_len) return fail(FGPU_ERR_INVALID, "null argument");
  if (tile_rows == 0) tile_rows = kTileRows;
  if (tile_rows != kTileRows) return fail(FGPU_ERR_INVALID, "tile_rows must be 0 (default)");
  std::string err;
  std::string js = describe_part_json(file, len, kIndexRows, &err);
  if (js.empty()) return fail(FGPU_ERR_PARQUET, err);
  *out_len = js.size();
  if (!buf) return FGPU_OK;
  if (cap < js.size()) return fail(FGPU_ERR_INVALID, "buffer too small");
  std::memcpy(buf, js.data(), js.size());
  return FGPU_OK;
  API_CATCH
}

}  // extern "C"
