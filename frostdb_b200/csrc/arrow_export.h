// Builds Arrow C Data Interface structs that own their buffers (released through the standard
// release callbacks).  Result records cross the C-ABI this way; on the Go side arrow-go's cdata
// package imports them, in the tests pyarrow's RecordBatch._import_from_c does.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/frostgpu.h"

namespace fgpu {

struct OwnedColumn {
  std::string name;
  std::string format;  // "l" int64, "g" float64, "I" uint32 (dictionary indices), "z" binary, "b" bool
  int64_t length = 0;
  int64_t null_count = 0;
  std::vector<uint8_t> validity;  // bitmap, empty = no nulls
  std::vector<uint8_t> data;      // fixed width values, or binary value bytes
  // fixed width values living in memory the column does not own as a vector (page-locked result blocks that go
  // back to their pool when the consumer releases the array): `ext` wins over `data` when set
  const uint8_t* ext = nullptr;
  std::shared_ptr<void> ext_keep;
  std::vector<int32_t> offsets;   // binary: length + 1 offsets
  std::unique_ptr<OwnedColumn> dictionary;  // for dictionary-encoded columns: the values
};

// Exports a struct array ("+s") with the given children.  Takes ownership of cols.
void export_record(std::vector<OwnedColumn>&& cols, int64_t length, ArrowSchema* out_schema, ArrowArray* out_array);
// Exports a single column.
void export_column(OwnedColumn&& col, ArrowSchema* out_schema, ArrowArray* out_array);

inline void set_bit(std::vector<uint8_t>& bm, int64_t i) { bm[size_t(i >> 3)] |= uint8_t(1u << (i & 7)); }

}  // namespace fgpu
