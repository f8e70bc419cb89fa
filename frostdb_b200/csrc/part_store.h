// Device-resident part registry: what Table.Iterator / LSM.Scan enumerate in the reference
// (table.go:740-868, index/lsm.go:401-454) lives here as a list of parts whose column chunks are
// already laid out in HBM for the scan kernel.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "device_types.h"
#include "parquet_meta.h"
#include "../../include/frostgpu.h"

namespace fgpu {

// Table-wide dictionary of one string column: chunk-local dictionary indices are translated to
// these ids at upload, so group keys compare as small integers across parts and row groups.
struct GlobalDict {
  std::vector<std::string> values;
  std::unordered_map<std::string, uint32_t> index;
  bool preloaded = false;  // fgpu_dict_preload named this column: it exists on some rank of a multi-GPU run
  uint32_t intern(const char* p, size_t n) {
    std::string s(p, n);
    auto it = index.find(s);
    if (it != index.end()) return it->second;
    uint32_t id = uint32_t(values.size());
    values.push_back(s);
    index.emplace(std::move(s), id);
    return id;
  }
  uint32_t cardinality() const { return uint32_t(values.size()); }
  const std::string& value(uint32_t id) const { return values[id]; }
};

struct ChunkHost {
  ChunkDesc desc{};            // device pointers (valid once the part is uploaded)
  int32_t phys = -1;           // Parquet physical type
  uint64_t stored_bytes = 0;   // footer total_compressed_size (algorithmic bytes)
  uint64_t meta_bytes = 0;     // directories / indexes / LUTs
  std::string error;           // non-empty: unreadable; an error only if a query projects it
  bool has_minmax = false;     // footer statistics of an INT64 / DOUBLE chunk (known before the column is built)
  int64_t min_bits = 0, max_bits = 0;
  int64_t null_count = -1;     // -1: not recorded
  bool has_minmax_str = false; // BYTE_ARRAY chunk: bounding strings of the non-null values
  std::string min_str, max_str;
  std::vector<uint8_t> bloom;      // split-block bloom filter of the chunk (copied out of the file at put; empty: none)
  std::vector<uint32_t> lut_host;  // CK_DICT_STR: chunk dictionary index -> *local* global id
  // section offsets inside the part image (patched into desc after upload)
  int64_t off_values = -1, off_runs = -1, off_seeds = -1, off_def = -1, off_def_runs = -1, off_def_seeds = -1, off_lut = -1,
          off_dict64 = -1;
  int64_t dev_seeds = -1, dev_def_seeds = -1;  // seeds derived on the device: offsets inside the image's device-only region
  int64_t off_row_runs = -1, off_row_seeds = -1, dev_row_seeds = -1;  // row-space directory of a nullable run-length column
  // flat code array of a CK_DICT_STR chunk (device pointer; derived on first use by k_flatten, see tile_agg.cu)
  const uint8_t* flat = nullptr;
  uint8_t flat_w = 0, flat_bias = 0;
};

struct RowGroupHost {
  uint32_t n_rows = 0;
  std::map<std::string, ChunkHost> cols;
};

// A byte range of the source Parquet file that is copied to the device as is (PLAIN value regions):
// no host-side copy is made, the H2D copy reads straight from the (ideally pinned) source buffer.
struct Extent {
  const uint8_t* src;
  uint64_t len;
  uint64_t dst_off;  // offset inside the column image
};

// Device residency of one column of one part (all its row groups in one allocation).
struct ColumnImage {
  bool built = false;      // host side done (ChunkHosts filled, dictionary interned)
  bool resident = false;   // bytes are on the device
  void* dev = nullptr;
  uint64_t dev_bytes = 0;
  std::vector<uint8_t> meta;     // host staging of everything that is not an extent (kept until uploaded)
  std::vector<Extent> extents;
  uint64_t seed_jobs_off = 0;    // SeedJob[n_seed_jobs] inside the meta region (device-derived seeds)
  uint32_t n_seed_jobs = 0, max_seed_chunks = 0;
  std::string error;             // whole-column error (a chunk the engine cannot read)
  void* flat_dev = nullptr;      // flat code arrays of every row group of this column (one allocation) + the job table
};

struct Part {
  uint64_t id = 0, tx = 0;
  const uint8_t* file = nullptr;    // source bytes: borrowed (caller keeps them alive) or `owned`
  uint64_t file_bytes = 0;
  std::vector<uint8_t> owned;
  bool borrowed = false;
  bool arrow = false;               // L0 Arrow record: LSM.Scan hands it to the plan unfiltered (index/lsm.go:420-427)
  ParsedFile pf;
  std::vector<RowGroupHost> rgs;
  std::vector<std::string> columns;  // schema order
  std::map<std::string, ColumnImage> images;
};

struct Table {
  uint64_t epoch = 0;  // changes whenever the set of parts or a dictionary does (0: not stamped yet); plans cache against it
  std::vector<std::unique_ptr<Part>> parts;  // insertion order
  std::map<std::string, GlobalDict> dicts;   // by column name
};

// Parses footer + page headers and prepares the row-group skeleton (no column is built yet).
bool open_part(const uint8_t* file, uint64_t len, Part* part, std::string* err);

// Builds the host side of one column of a part (no CUDA calls): run directories, seeds, dictionary
// LUTs into `image.meta`, PLAIN value regions as extents.  Interns dictionary entries into `table`.
// ChunkDesc pointers are offsets until patch_column_pointers() is called.
// device_seeds: leave the cursor seeds to k_make_seeds (ColumnImage::seed_jobs_*) instead of building them here.
void build_column(int index_rows, Table* table, Part* part, const std::string& column, bool device_seeds = false);

// L0 Arrow-record part (parts/arrow.go:14-55): one row group holding the record's rows.  Every column image
// is built here (the record is released by the caller afterwards): int64 / float64 arrays become PLAIN
// chunks (non-null values packed, the validity bitmap IS the bit-packed definition-level stream),
// dictionary<int, binary|utf8> and plain binary/utf8 arrays become dictionary chunks whose index stream is
// one 32-bit "bit-packed" run.  Returns false and sets err for unsupported layouts.
bool build_arrow_part(Table* table, Part* part, const ::ArrowSchema* schema, const ::ArrowArray* array, std::string* err);

// Patches device pointers into the column's ChunkDescs once the image lives at `dev_base`.
void patch_column_pointers(Part* part, const std::string& column, const uint8_t* dev_base);

// JSON description for fgpu_parquet_describe (host-only tests).
std::string describe_part_json(const uint8_t* file, uint64_t len, int tile_rows, std::string* err);

}  // namespace fgpu
