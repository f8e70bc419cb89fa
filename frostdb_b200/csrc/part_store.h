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

namespace fgpu {

// Table-wide dictionary of one string column: chunk-local dictionary indices are translated to
// these ids at upload, so group keys compare as small integers across parts and row groups.
struct GlobalDict {
  std::vector<std::string> values;
  std::unordered_map<std::string, uint32_t> index;
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
  std::vector<uint32_t> lut_host;  // CK_DICT_STR: chunk dictionary index -> *local* global id
  // section offsets inside the part image (patched into desc after upload)
  int64_t off_values = -1, off_runs = -1, off_seeds = -1, off_def = -1, off_def_runs = -1, off_def_seeds = -1, off_lut = -1,
          off_dict64 = -1;
};

struct RowGroupHost {
  uint32_t n_rows = 0;
  std::map<std::string, ChunkHost> cols;
};

struct Part {
  uint64_t id = 0, tx = 0;
  std::vector<RowGroupHost> rgs;
  std::vector<std::string> columns;  // schema order
  void* dev = nullptr;               // the part's single device allocation
  uint64_t dev_bytes = 0;
  uint64_t file_bytes = 0;
  std::vector<uint8_t> image;        // host image (kept only until upload)
};

struct Table {
  std::vector<std::unique_ptr<Part>> parts;  // insertion order
  std::map<std::string, GlobalDict> dicts;   // by column name
};

// Builds the host image of a part (no CUDA calls): sections for every readable column chunk,
// run directories, tile indexes, dictionary LUTs.  Interns dictionary entries into `table`.
bool build_part_image(const uint8_t* file, uint64_t len, int tile_rows, Table* table, Part* part, std::string* err);

// Patches device pointers into every ChunkDesc once the image lives at `dev_base`.
void patch_part_pointers(Part* part, const uint8_t* dev_base);

// JSON description for fgpu_parquet_describe (host-only tests).
std::string describe_part_json(const uint8_t* file, uint64_t len, int tile_rows, std::string* err);

}  // namespace fgpu
