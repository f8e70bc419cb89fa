// Host-side Parquet reader for libfrostgpu: footer (Thrift compact), page walk, hybrid-run walk.
// Produces what the device needs to decode pages without touching Thrift: per column chunk the
// page payload extents, and per hybrid stream (dictionary indices, definition levels) a run
// directory.  Takes over the host half of pqarrow.ParquetConverter / parquet-go's page readers
// (reference: pqarrow/arrow.go:711-823 writeColumnToArray page loop; format facts SURVEY.md App. C).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fgpu {

enum PhysType : int32_t {
  PT_BOOLEAN = 0, PT_INT32 = 1, PT_INT64 = 2, PT_INT96 = 3, PT_FLOAT = 4, PT_DOUBLE = 5,
  PT_BYTE_ARRAY = 6, PT_FIXED_LEN_BYTE_ARRAY = 7
};
enum PqEncoding : int32_t {
  ENC_PLAIN = 0, ENC_PLAIN_DICTIONARY = 2, ENC_RLE = 3, ENC_BIT_PACKED = 4,
  ENC_DELTA_BINARY_PACKED = 5, ENC_DELTA_LENGTH_BYTE_ARRAY = 6, ENC_DELTA_BYTE_ARRAY = 7,
  ENC_RLE_DICTIONARY = 8, ENC_BYTE_STREAM_SPLIT = 9
};
enum PageType : int32_t { PG_DATA = 0, PG_INDEX = 1, PG_DICTIONARY = 2, PG_DATA_V2 = 3 };

struct SchemaLeaf {
  std::string name;  // dotted path, e.g. "labels.job"
  int32_t phys = -1;
  int32_t max_def = 0;
  int32_t max_rep = 0;
  bool is_unsigned = false;  // converted type UINT_64 / logical INT(64,false)
};

struct PageInfo {
  int32_t type = 0;        // PG_DATA or PG_DATA_V2
  int32_t encoding = 0;    // of the values
  uint32_t num_values = 0; // rows in the page (flat columns), nulls included
  int64_t num_nulls = -1;  // -1: unknown (v1 page)
  const uint8_t* def = nullptr;  // RLE/bit-packed hybrid of definition levels (no length prefix)
  uint32_t def_len = 0;
  const uint8_t* values = nullptr;  // encoded values region
  uint32_t values_len = 0;
};

struct ChunkMeta {
  int32_t leaf = -1;  // index into ParsedFile::leaves
  int32_t codec = 0;
  int64_t num_values = 0;
  int64_t total_compressed_size = 0;  // footer figure: the algorithmic bytes of this chunk
  int64_t null_count = -1;            // from chunk statistics when present
  bool has_minmax = false;            // INT64 / DOUBLE chunks: footer statistics carry both bounds
  int64_t min_bits = 0, max_bits = 0; // raw 8-byte bounds of the non-null values
  bool has_minmax_str = false;        // BYTE_ARRAY chunks: both (possibly truncated, still bounding) strings recorded
  std::string min_str, max_str;
  // split-block bloom filter of the chunk (parquet-go writes one per sorting column, dynparquet/schema.go:1111-1157):
  // the bitset (aliases `file`), null when the chunk has none or it uses an algorithm / hash / compression other than
  // BLOCK / XXHASH / UNCOMPRESSED
  const uint8_t* bloom = nullptr;
  uint32_t bloom_bytes = 0;
  const uint8_t* dict = nullptr;      // dictionary page payload (PLAIN)
  uint32_t dict_len = 0;
  uint32_t dict_num_values = 0;
  std::vector<PageInfo> pages;
  std::string error;  // non-empty: chunk is unreadable by this engine (reason); only an error if projected
  // page walk state: footers are parsed eagerly, page headers only for the chunks a query touches
  int64_t data_page_offset = -1, dict_page_offset = -1;
  bool pages_walked = false;
};

struct RowGroupMeta {
  int64_t num_rows = 0;
  std::vector<ChunkMeta> chunks;  // one per leaf, in leaf order
};

struct ParsedFile {
  std::vector<SchemaLeaf> leaves;
  std::vector<RowGroupMeta> row_groups;
  std::vector<std::pair<std::string, std::string>> kv;
  int64_t num_rows = 0;
  std::string created_by;
};

// Parses footer and walks every page header.  Pointers in the result alias `file`.
// Returns false and sets err on malformed input.
// With walk_pages == false only the footer is read; walk_chunk_pages() then fills a chunk on demand.
bool parse_parquet(const uint8_t* file, uint64_t len, ParsedFile* out, std::string* err, bool walk_pages = true);

// Walks the page headers of one column chunk (idempotent): fills pages / dict, or sets error.
void walk_chunk_pages(const uint8_t* file, uint64_t len, const SchemaLeaf& leaf, int64_t rg_rows, ChunkMeta* cm);

// XXH64 (the hash of Parquet bloom filters, seed 0 over the PLAIN encoding of the value: the 8 little-endian bytes of an
// INT64 / DOUBLE, the bytes of a BYTE_ARRAY without their length prefix).
uint64_t xxhash64(const uint8_t* data, size_t len, uint64_t seed);
// Split-block bloom filter check (Parquet format, BloomFilter.md): false = the value is definitely not in the chunk.
bool sbbf_check(const uint8_t* bitset, uint32_t bytes, uint64_t hash);
// Sets the bits of `hash` (the writer's half; used by the tests to build filters).
void sbbf_insert(uint8_t* bitset, uint32_t bytes, uint64_t hash);

// One run of an RLE/bit-packed hybrid stream (Parquet "RLE" encoding).
struct HostRun {
  uint32_t start;  // ordinal of the run's first value within the chunk's value sequence
  uint32_t off;    // byte offset of the payload inside the chunk's concatenated stream section
  uint32_t val;    // RLE: the repeated value
  uint32_t meta;   // bit0: 1 = bit-packed, 0 = RLE; bits 8..15: bit width
};

// Walks one hybrid stream holding `count` values of width `w` (count clamps padded tails).
// Appends runs (start offsets relative to `start0`, payload offsets relative to `off0` + position
// in `data`).  Zero-length runs are never emitted.  Returns false on overrun/malformed input.
bool walk_hybrid(const uint8_t* data, uint32_t len, int w, uint32_t count, uint32_t start0,
                 uint32_t off0, std::vector<HostRun>* runs, std::string* err);

// Reads value i of a hybrid stream described by runs (host-side reference used by describe/tests).
uint32_t hybrid_value_at(const uint8_t* stream, const std::vector<HostRun>& runs, uint32_t ordinal, bool* was_rle = nullptr);

}  // namespace fgpu
