#include "part_store.h"

#include <algorithm>
#include <cstring>
#include <sstream>
#include <unordered_map>

namespace fgpu {
namespace {

constexpr size_t kAlign = 128;
constexpr size_t kTailPad = 16;

// Column image layout: [meta region: every section the host assembles][extent region: PLAIN value
// bytes copied verbatim from the source file].  Section offsets of the extent region are recorded
// relative to its start and rebased when the meta size is final.
struct ImageWriter {
  std::vector<uint8_t>& buf;
  std::vector<Extent>& extents;
  uint64_t ext_size = 0;
  uint64_t dev_size = 0;            // device-only region behind the extents (seeds derived on the device)
  bool device_seeds = false;
  std::vector<SeedJob> jobs;
  ImageWriter(std::vector<uint8_t>& b, std::vector<Extent>& e) : buf(b), extents(e) {}
  uint64_t reserve_dev(uint64_t n) {
    dev_size = (dev_size + kAlign - 1) / kAlign * kAlign;
    uint64_t off = dev_size;
    dev_size += n + kTailPad;
    return off;
  }
  size_t begin() {
    size_t n = (buf.size() + kAlign - 1) / kAlign * kAlign;
    buf.resize(n, 0);
    return n;
  }
  void append(const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    buf.insert(buf.end(), b, b + n);
  }
  void end() { buf.resize(buf.size() + kTailPad, 0); }
  int64_t section(const void* p, size_t n) {
    size_t off = begin();
    append(p, n);
    end();
    return int64_t(off);
  }
  // reserves an aligned hole in the extent region; returns its offset inside that region
  uint64_t begin_extent() {
    ext_size = (ext_size + kAlign - 1) / kAlign * kAlign;
    return ext_size;
  }
  void add_extent(const uint8_t* src, uint64_t len) {
    if (len == 0) return;
    if (!extents.empty() && extents.back().src + extents.back().len == src && extents.back().dst_off + extents.back().len == ext_size)
      extents.back().len += len;  // contiguous in the file too: one copy
    else
      extents.push_back({src, len, ext_size});
    ext_size += len;
  }
  void end_extent() { ext_size += kTailPad; }
};

inline void or_bits(std::vector<uint64_t>& bm, uint64_t pos, uint64_t bits8) {
  size_t w = size_t(pos >> 6);
  unsigned sh = unsigned(pos & 63);
  bm[w] |= bits8 << sh;
  if (sh > 56 && w + 1 < bm.size()) bm[w + 1] |= bits8 >> (64 - sh);
}

inline void set_range(std::vector<uint64_t>& bm, uint64_t pos, uint64_t len) {
  uint64_t end = pos + len;
  while (pos < end) {
    size_t w = size_t(pos >> 6);
    unsigned sh = unsigned(pos & 63);
    uint64_t take = std::min<uint64_t>(64 - sh, end - pos);
    uint64_t mask = (take == 64) ? ~0ull : (((1ull << take) - 1ull) << sh);
    bm[w] |= mask;
    pos += take;
  }
}

inline uint64_t popcount_range(const std::vector<uint64_t>& bm, uint64_t pos, uint64_t end) {
  uint64_t n = 0;
  while (pos < end) {
    size_t w = size_t(pos >> 6);
    unsigned sh = unsigned(pos & 63);
    uint64_t take = std::min<uint64_t>(64 - sh, end - pos);
    uint64_t mask = (take == 64) ? ~0ull : (((1ull << take) - 1ull) << sh);
    n += uint64_t(__builtin_popcountll(bm[w] & mask));
    pos += take;
  }
  return n;
}

bool is_dict_encoding(int32_t e) { return e == ENC_RLE_DICTIONARY || e == ENC_PLAIN_DICTIONARY; }

// Builds every section of one column chunk into the image.
void build_chunk(const ChunkMeta& cm, const SchemaLeaf& leaf, uint32_t n_rows, int tile_rows, GlobalDict* dict,
                 ImageWriter& w, ChunkHost* out) {
  out->phys = leaf.phys;
  out->stored_bytes = uint64_t(cm.total_compressed_size);
  out->has_minmax = cm.has_minmax;
  out->min_bits = cm.min_bits;
  out->max_bits = cm.max_bits;
  out->null_count = (cm.null_count < 0 && leaf.max_def == 0) ? 0 : cm.null_count;
  out->has_minmax_str = cm.has_minmax_str;
  out->min_str = cm.min_str;
  out->max_str = cm.max_str;
  out->desc.n_rows = n_rows;
  if (!cm.error.empty()) { out->error = cm.error; return; }
  if (leaf.phys != PT_INT64 && leaf.phys != PT_DOUBLE && leaf.phys != PT_BYTE_ARRAY) {
    out->error = "unsupported physical type " + std::to_string(leaf.phys);
    return;
  }
  if (leaf.is_unsigned) { out->error = "uint64 columns are not supported"; return; }
  if (leaf.max_def > 1) { out->error = "nested optional columns are not supported"; return; }
  bool any_dict = false, any_plain = false;
  for (const PageInfo& pg : cm.pages) {
    if (is_dict_encoding(pg.encoding)) any_dict = true;
    else if (pg.encoding == ENC_PLAIN) any_plain = true;
    else { out->error = "unsupported page encoding " + std::to_string(pg.encoding); return; }
  }
  if (any_dict && any_plain) { out->error = "column chunk mixes dictionary and PLAIN pages"; return; }
  if (leaf.phys == PT_BYTE_ARRAY && any_plain) { out->error = "PLAIN byte-array pages are not supported"; return; }
  const uint32_t T = uint32_t(tile_rows);
  const uint32_t n_tiles = (n_rows + T - 1) / T;

  // ---- definition levels --------------------------------------------------------------------
  std::string err;
  std::vector<uint8_t> defstream;
  std::vector<HostRun> defruns;
  std::vector<uint64_t> valid;  // bitset, only when max_def == 1
  std::vector<uint32_t> page_nn(cm.pages.size());
  uint32_t n_values = n_rows;
  bool all_present = false;  // optional column whose definition levels are one run of 1s per page: no bitset needed
  if (leaf.max_def == 1) {
    uint32_t row = 0;
    for (size_t p = 0; p < cm.pages.size(); p++) {
      const PageInfo& pg = cm.pages[p];
      size_t first = defruns.size();
      // a merged RLE run may extend the previous page's last run: remember where this page starts
      uint32_t page_start = row;
      if (!walk_hybrid(pg.def, pg.def_len, 1, pg.num_values, row, uint32_t(defstream.size()), &defruns, &err)) {
        out->error = "definition levels: " + err;
        return;
      }
      defstream.insert(defstream.end(), pg.def, pg.def + pg.def_len);
      (void)first;
      row += pg.num_values;
      (void)page_start;
    }
    if (row != n_rows) { out->error = "definition levels do not cover the row group"; return; }
    all_present = defruns.size() == 1 && (defruns[0].meta & 1u) == 0 && defruns[0].val == 1;
    for (size_t p = 0; p < cm.pages.size() && all_present; p++)
      if (cm.pages[p].num_nulls > 0) {
        out->error = "page num_nulls disagrees with its definition levels";
        return;
      }
  }
  if (leaf.max_def == 1 && !all_present) {
    valid.assign((size_t(n_rows) + 63) / 64 + 1, 0);
    uint32_t row = 0;
    for (size_t k = 0; k < defruns.size(); k++) {
      const HostRun& r = defruns[k];
      uint32_t end = (k + 1 < defruns.size()) ? defruns[k + 1].start : n_rows;
      uint32_t len = end - r.start;
      if ((r.meta & 1u) == 0) {
        if (r.val) set_range(valid, r.start, len);
      } else {
        for (uint32_t done = 0; done < len; done += 8) {
          uint64_t b = defstream[r.off + done / 8];
          if (len - done < 8) b &= (1ull << (len - done)) - 1ull;
          or_bits(valid, uint64_t(r.start) + done, b);
        }
      }
    }
    row = 0;
    uint64_t total_valid = 0;
    for (size_t p = 0; p < cm.pages.size(); p++) {
      const PageInfo& pg = cm.pages[p];
      uint64_t nn = popcount_range(valid, row, uint64_t(row) + pg.num_values);
      if (pg.num_nulls >= 0 && uint64_t(pg.num_values) - uint64_t(pg.num_nulls) != nn) {
        out->error = "page num_nulls disagrees with its definition levels";
        return;
      }
      page_nn[p] = uint32_t(nn);
      total_valid += nn;
      row += pg.num_values;
    }
    n_values = uint32_t(total_valid);
  } else {
    for (size_t p = 0; p < cm.pages.size(); p++) page_nn[p] = cm.pages[p].num_values;
  }
  const bool has_nulls = (n_values != n_rows);

  // ---- values -------------------------------------------------------------------------------
  std::vector<uint8_t> vstream;
  std::vector<HostRun> vruns;
  if (any_dict || (!any_plain && leaf.phys == PT_BYTE_ARRAY)) {
    uint32_t vo = 0;
    for (size_t p = 0; p < cm.pages.size(); p++) {
      const PageInfo& pg = cm.pages[p];
      if (page_nn[p] == 0) continue;
      if (pg.values_len < 1) { out->error = "dictionary-index page without bit width"; return; }
      int bw = pg.values[0];
      if (bw > 32) { out->error = "dictionary index bit width > 32"; return; }
      if (!walk_hybrid(pg.values + 1, pg.values_len - 1, bw, page_nn[p], vo, uint32_t(vstream.size()), &vruns, &err)) {
        out->error = "dictionary indices: " + err;
        return;
      }
      vstream.insert(vstream.end(), pg.values + 1, pg.values + pg.values_len);
      vo += page_nn[p];
    }
    out->desc.kind = (leaf.phys == PT_BYTE_ARRAY) ? CK_DICT_STR : CK_DICT64;
  } else {
    for (size_t p = 0; p < cm.pages.size(); p++) {
      const PageInfo& pg = cm.pages[p];
      uint64_t need = uint64_t(page_nn[p]) * 8;
      if (need > pg.values_len) { out->error = "PLAIN page shorter than its value count"; return; }
    }
    out->desc.kind = CK_PLAIN64;
  }

  // ---- dictionary ---------------------------------------------------------------------------
  std::vector<int64_t> dict64;
  uint32_t dict_size = 0;
  if (out->desc.kind == CK_DICT_STR) {
    const uint8_t* p = cm.dict;
    const uint8_t* end = cm.dict + cm.dict_len;
    out->lut_host.reserve(cm.dict_num_values);
    for (uint32_t i = 0; i < cm.dict_num_values; i++) {
      if (end - p < 4) { out->error = "dictionary page truncated"; return; }
      uint32_t l;
      std::memcpy(&l, p, 4);
      p += 4;
      if (l > uint64_t(end - p)) { out->error = "dictionary entry overruns page"; return; }
      out->lut_host.push_back(dict->intern(reinterpret_cast<const char*>(p), l));
      p += l;
    }
    dict_size = cm.dict_num_values;
  } else if (out->desc.kind == CK_DICT64) {
    if (uint64_t(cm.dict_num_values) * 8 > cm.dict_len) { out->error = "numeric dictionary page truncated"; return; }
    dict64.resize(cm.dict_num_values);
    std::memcpy(dict64.data(), cm.dict, size_t(cm.dict_num_values) * 8);
    dict_size = cm.dict_num_values;
  }
  if (out->desc.kind != CK_PLAIN64) {
    for (const HostRun& r : vruns)
      if ((r.meta & 1u) == 0 && r.val >= dict_size) { out->error = "dictionary index out of range"; return; }
    if (n_values > 0 && dict_size == 0) { out->error = "dictionary-encoded values without a dictionary page"; return; }
  }

  // ---- RLE run values of string dictionaries become global ids (no LUT hop for run-length data) ----
  if (out->desc.kind == CK_DICT_STR)
    for (HostRun& r : vruns)
      if ((r.meta & 1u) == 0) r.val = out->lut_host[r.val];

  // ---- chunk seeds ------------------------------------------------------------------------------
  const uint32_t n_chunks = n_tiles;  // `tile_rows` is the seed granularity (kIndexRows)
  std::vector<uint32_t> chunk_val0((has_nulls || !w.device_seeds) ? n_chunks : 0);
  if (!has_nulls && w.device_seeds) {
    // k_make_seeds derives chunk * 128 itself
  } else if (has_nulls) {
    uint64_t acc = 0;
    for (uint32_t t = 0; t < n_chunks; t++) {
      chunk_val0[t] = uint32_t(acc);
      uint64_t b = uint64_t(t) * T, e = std::min<uint64_t>(uint64_t(n_rows), b + T);
      acc += popcount_range(valid, b, e);
    }
  } else {
    for (uint32_t t = 0; t < n_chunks; t++) chunk_val0[t] = t * T;
  }
  auto make_seeds = [&](const std::vector<HostRun>& runs, uint32_t total, bool is_def) {
    std::vector<Seed> seeds(n_chunks);
    size_t k = 0;
    for (uint32_t t = 0; t < n_chunks; t++) {
      uint32_t first = is_def ? t * T : chunk_val0[t];
      while (k + 1 < runs.size() && runs[k + 1].start <= first) k++;
      Seed sd{};
      sd.val0 = chunk_val0[t];
      if (runs.empty() || first >= total) {  // nothing left to decode from this chunk on
        sd.k = uint32_t(runs.size());
        sd.start = total;
        sd.end = 0xffffffffu;
      } else {
        const HostRun& r = runs[k];
        sd.k = uint32_t(k);
        sd.start = r.start;
        sd.end = (k + 1 < runs.size()) ? runs[k + 1].start : total;
        sd.off = r.off;
        sd.val = r.val;
        sd.meta = r.meta;
      }
      seeds[t] = sd;
    }
    return seeds;
  };

  // ---- write sections ----------------------------------------------------------------------------
  const size_t before = w.buf.size();
  out->desc.has_nulls = has_nulls ? 1 : 0;
  out->desc.n_values = n_values;
  out->desc.dict_size = dict_size;
  uint64_t plain_bytes = 0;
  if (out->desc.kind == CK_PLAIN64) {
    // value regions go to the device straight from the file: negative offsets mark "extent region"
    const uint64_t eoff = w.begin_extent();
    for (size_t p = 0; p < cm.pages.size(); p++) {
      w.add_extent(cm.pages[p].values, uint64_t(page_nn[p]) * 8);
      plain_bytes += uint64_t(page_nn[p]) * 8;
    }
    w.end_extent();
    out->off_values = -int64_t(eoff) - 1;
  } else {
    out->off_values = w.section(vstream.data(), vstream.size());
  }
  uint64_t val0_off = ~0ull;
  if (w.device_seeds && has_nulls) val0_off = uint64_t(w.section(chunk_val0.data(), chunk_val0.size() * 4));
  auto seed_job = [&](int64_t runs_off, uint32_t n_runs, uint32_t total, bool is_def, bool use_val0 = true) {
    SeedJob j{};
    j.runs_off = uint64_t(runs_off);
    j.seeds_off = w.reserve_dev(uint64_t(n_chunks) * sizeof(Seed));
    j.val0_off = use_val0 ? val0_off : ~0ull;
    j.n_runs = n_runs;
    j.total = total;
    j.n_chunks = n_chunks;
    j.is_def = is_def ? 1 : 0;
    w.jobs.push_back(j);
    return int64_t(j.seeds_off);
  };
  // ---- row-space directory of a nullable dictionary column whose two streams are run-length only ----
  // (what a sorted part leaves: NULLs first, then one run per value).  The sorted-run kernel walks it like
  // the directory of a column without NULLs; 0xffffffff stands for NULL (contributes 0 to the dense slot).
  std::vector<HostRun> row_runs;
  if (out->desc.kind == CK_DICT_STR && has_nulls) {
    bool rle_only = true;
    for (const HostRun& r : vruns) rle_only = rle_only && (r.meta & 1u) == 0;
    for (const HostRun& r : defruns) rle_only = rle_only && (r.meta & 1u) == 0;
    if (rle_only) {
      auto push = [&](uint32_t start, uint32_t val) {
        if (!row_runs.empty() && row_runs.back().val == val) return;
        row_runs.push_back(HostRun{start, 0, val, 0});
      };
      size_t vk = 0;
      uint32_t vord = 0;  // ordinal of the next non-null value
      for (size_t k = 0; k < defruns.size(); k++) {
        const uint32_t s0 = defruns[k].start, e0 = (k + 1 < defruns.size()) ? defruns[k + 1].start : n_rows;
        if (defruns[k].val == 0) { push(s0, 0xffffffffu); continue; }
        uint32_t row = s0;
        while (row < e0) {
          while (vk + 1 < vruns.size() && vruns[vk + 1].start <= vord) vk++;
          const uint32_t vend = (vk + 1 < vruns.size()) ? vruns[vk + 1].start : n_values;
          const uint32_t take = std::min(e0 - row, vend - vord);
          push(row, vruns[vk].val);
          row += take;
          vord += take;
        }
      }
    }
  }
  if (out->desc.kind != CK_PLAIN64) {
    out->desc.n_runs = uint32_t(vruns.size());
    uint32_t bp = 0;
    for (const HostRun& r : vruns) bp += (r.meta & 1u);
    out->desc.n_bp_runs = bp;
    std::vector<Seed> seeds;
    if (!w.device_seeds) seeds = make_seeds(vruns, n_values, false);
    HostRun sentinel{n_values, 0, 0, 0};
    vruns.push_back(sentinel);
    out->off_runs = w.section(vruns.data(), vruns.size() * sizeof(HostRun));
    if (w.device_seeds) out->dev_seeds = seed_job(out->off_runs, out->desc.n_runs, n_values, false);
    else out->off_seeds = w.section(seeds.data(), seeds.size() * sizeof(Seed));
  }
  if (has_nulls) {
    out->desc.n_defruns = uint32_t(defruns.size());
    std::vector<Seed> seeds;
    if (!w.device_seeds) seeds = make_seeds(defruns, n_rows, true);
    HostRun sentinel{n_rows, 0, 0, 0};
    defruns.push_back(sentinel);
    out->off_def = w.section(defstream.data(), defstream.size());
    out->off_def_runs = w.section(defruns.data(), defruns.size() * sizeof(HostRun));
    if (w.device_seeds) out->dev_def_seeds = seed_job(out->off_def_runs, out->desc.n_defruns, n_rows, true);
    else out->off_def_seeds = w.section(seeds.data(), seeds.size() * sizeof(Seed));
  }
  if (!row_runs.empty()) {
    out->desc.n_row_runs = uint32_t(row_runs.size());
    std::vector<Seed> seeds;
    if (!w.device_seeds) seeds = make_seeds(row_runs, n_rows, true);
    row_runs.push_back(HostRun{n_rows, 0, 0, 0});
    out->off_row_runs = w.section(row_runs.data(), row_runs.size() * sizeof(HostRun));
    if (w.device_seeds) out->dev_row_seeds = seed_job(out->off_row_runs, out->desc.n_row_runs, n_rows, true, /*use_val0=*/false);
    else out->off_row_seeds = w.section(seeds.data(), seeds.size() * sizeof(Seed));
  }
  // Bit-packed indices are not range-checked row by row (neither here nor in the kernels): the tables they index are
  // padded to 2^w entries instead (w = widest bit-packed run of the chunk), so that any w-bit pattern of a corrupt
  // chunk reads a valid entry (dictionary entry 0) instead of memory behind the table.
  uint32_t max_w = 0;
  for (const HostRun& r : vruns)
    if (r.meta & 1u) max_w = std::max<uint32_t>(max_w, (r.meta >> 8) & 0xffu);
  const size_t padded = max_w > 0 && max_w <= 24 ? (size_t(1) << max_w) : 0;
  if (max_w > 24 && out->desc.kind != CK_PLAIN64 && (uint64_t(1) << std::min<uint32_t>(max_w, 40)) > uint64_t(dict_size) * 2 + 2) {
    out->error = "bit-packed dictionary indices wider than the dictionary";
    return;
  }
  if (out->desc.kind == CK_DICT_STR) {
    if (padded > out->lut_host.size()) {
      std::vector<uint32_t> lut = out->lut_host;
      lut.resize(padded, lut.empty() ? 0u : lut[0]);
      out->off_lut = w.section(lut.data(), lut.size() * 4);
    } else {
      out->off_lut = w.section(out->lut_host.data(), out->lut_host.size() * 4);
    }
  }
  if (out->desc.kind == CK_DICT64) {
    if (padded > dict64.size()) dict64.resize(padded, dict64.empty() ? 0 : dict64[0]);
    out->off_dict64 = w.section(dict64.data(), dict64.size() * 8);
  }
  (void)plain_bytes;
  const size_t payload = vstream.size() + (has_nulls ? defstream.size() : 0);
  out->meta_bytes = (w.buf.size() - before) - std::min(w.buf.size() - before, payload);
  if (out->dev_seeds >= 0) out->meta_bytes += uint64_t(n_chunks) * sizeof(Seed);
  if (out->dev_def_seeds >= 0) out->meta_bytes += uint64_t(n_chunks) * sizeof(Seed);
}

// Closes a column image: seed jobs table, final offsets of the extent and device-only regions.
void finish_image(ImageWriter& w, ColumnImage& img, Part* part, const std::string& column) {
  if (!w.jobs.empty()) {
    img.seed_jobs_off = uint64_t(w.section(w.jobs.data(), w.jobs.size() * sizeof(SeedJob)));
    img.n_seed_jobs = uint32_t(w.jobs.size());
    for (const SeedJob& j : w.jobs) img.max_seed_chunks = std::max(img.max_seed_chunks, j.n_chunks);
  }
  if (img.meta.empty()) img.meta.resize(kAlign, 0);
  // rebase extent offsets behind the (128-aligned) meta region, the device-only region behind the extents
  const uint64_t meta_size = (img.meta.size() + kAlign - 1) / kAlign * kAlign;
  const uint64_t dev_base = (meta_size + w.ext_size + kAlign - 1) / kAlign * kAlign;
  for (Extent& e : img.extents) e.dst_off += meta_size;
  for (RowGroupHost& h : part->rgs) {
    ChunkHost& ch = h.cols[column];
    if (ch.off_values < -0) ch.off_values = int64_t(meta_size) + (-ch.off_values - 1);
    if (ch.dev_seeds >= 0) ch.off_seeds = int64_t(dev_base) + ch.dev_seeds;
    if (ch.dev_def_seeds >= 0) ch.off_def_seeds = int64_t(dev_base) + ch.dev_def_seeds;
    if (ch.dev_row_seeds >= 0) ch.off_row_seeds = int64_t(dev_base) + ch.dev_row_seeds;
  }
  for (uint32_t i = 0; i < img.n_seed_jobs; i++) {
    SeedJob* j = reinterpret_cast<SeedJob*>(img.meta.data() + img.seed_jobs_off) + i;
    j->seeds_off += dev_base;
  }
  img.dev_bytes = dev_base + w.dev_size + kAlign;
}


}  // namespace

bool open_part(const uint8_t* file, uint64_t len, Part* part, std::string* err) {
  if (!parse_parquet(file, len, &part->pf, err, /*walk_pages=*/false)) return false;  // page headers: on demand, per column
  part->file = file;
  part->file_bytes = len;
  part->columns.clear();
  for (const SchemaLeaf& l : part->pf.leaves) part->columns.push_back(l.name);
  for (const RowGroupMeta& rg : part->pf.row_groups) {
    if (rg.num_rows == 0) continue;
    if (rg.num_rows > 0x7fffffffll) { *err = "row group with more than 2^31 rows"; return false; }
    RowGroupHost h;
    h.n_rows = uint32_t(rg.num_rows);
    for (size_t c = 0; c < rg.chunks.size(); c++) {
      ChunkHost ch;  // skeleton: type and footer size are known before the column is built
      ch.phys = part->pf.leaves[c].phys;
      ch.stored_bytes = uint64_t(rg.chunks[c].total_compressed_size);
      ch.has_minmax = rg.chunks[c].has_minmax;
      ch.min_bits = rg.chunks[c].min_bits;
      ch.max_bits = rg.chunks[c].max_bits;
      ch.null_count = rg.chunks[c].null_count;
      if (ch.null_count < 0 && part->pf.leaves[c].max_def == 0) ch.null_count = 0;  // required column
      ch.has_minmax_str = rg.chunks[c].has_minmax_str;
      ch.min_str = rg.chunks[c].min_str;
      ch.max_str = rg.chunks[c].max_str;
      // (parquet-go sizes the filter by the chunk's values, 10 bits each: anything beyond a few MB is left alone)
      if (rg.chunks[c].bloom && rg.chunks[c].bloom_bytes <= (8u << 20))
        ch.bloom.assign(rg.chunks[c].bloom, rg.chunks[c].bloom + rg.chunks[c].bloom_bytes);
      ch.desc.n_rows = h.n_rows;
      h.cols.emplace(part->pf.leaves[c].name, std::move(ch));
    }
    part->rgs.push_back(std::move(h));
  }
  return true;
}

void build_column(int index_rows, Table* table, Part* part, const std::string& column, bool device_seeds) {
  ColumnImage& img = part->images[column];
  if (img.built) return;
  img.built = true;
  size_t leaf = part->pf.leaves.size();
  for (size_t c = 0; c < part->pf.leaves.size(); c++)
    if (part->pf.leaves[c].name == column) leaf = c;
  if (leaf == part->pf.leaves.size()) { img.error = "column not in part"; return; }
  const SchemaLeaf& sl = part->pf.leaves[leaf];
  GlobalDict* dict = (sl.phys == PT_BYTE_ARRAY) ? &table->dicts[column] : nullptr;
  ImageWriter w(img.meta, img.extents);
  w.device_seeds = device_seeds;
  size_t g = 0;
  for (RowGroupMeta& rg : part->pf.row_groups) {
    if (rg.num_rows == 0) continue;
    RowGroupHost& h = part->rgs[g++];
    ChunkHost ch;
    walk_chunk_pages(part->file, part->file_bytes, sl, rg.num_rows, &rg.chunks[leaf]);  // first touch of this chunk
    build_chunk(rg.chunks[leaf], sl, h.n_rows, index_rows, dict, w, &ch);
    if (!ch.error.empty() && img.error.empty()) img.error = ch.error;
    ch.bloom = std::move(h.cols[column].bloom);  // (footer-level fact of the skeleton the built chunk replaces)
    h.cols[column] = std::move(ch);
  }
  finish_image(w, img, part, column);
}

namespace {

inline bool bit_at(const uint8_t* bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

// Integer view of an Arrow index buffer element.
inline int64_t arrow_index(const void* buf, char fmt, int64_t i) {
  switch (fmt) {
    case 'c': return static_cast<const int8_t*>(buf)[i];
    case 'C': return static_cast<const uint8_t*>(buf)[i];
    case 's': return static_cast<const int16_t*>(buf)[i];
    case 'S': return static_cast<const uint16_t*>(buf)[i];
    case 'i': return static_cast<const int32_t*>(buf)[i];
    case 'I': return static_cast<const uint32_t*>(buf)[i];
    case 'l': return static_cast<const int64_t*>(buf)[i];
    default: return int64_t(static_cast<const uint64_t*>(buf)[i]);
  }
}

// One Arrow column -> one column image (one chunk).
bool build_arrow_column(Table* table, Part* part, const std::string& name, const ::ArrowSchema* cs, const ::ArrowArray* ca,
                        uint32_t n_rows, std::string* err) {
  const std::string fmt = cs->format ? cs->format : "";
  const bool is_dict = cs->dictionary != nullptr;
  const bool is_i64 = !is_dict && fmt == "l";
  const bool is_f64 = !is_dict && fmt == "g";
  const bool is_str = !is_dict && (fmt == "u" || fmt == "z");
  if (!is_dict && !is_i64 && !is_f64 && !is_str) { *err = "column " + name + ": unsupported Arrow type '" + fmt + "'"; return false; }
  if (is_dict && (fmt.size() != 1 || std::string("cCsSiIlL").find(fmt[0]) == std::string::npos)) {
    *err = "column " + name + ": unsupported dictionary index type '" + fmt + "'";
    return false;
  }
  if (uint64_t(ca->length) != n_rows) { *err = "column " + name + ": length differs from the record"; return false; }
  const int64_t off = ca->offset;
  const uint8_t* validity = (ca->n_buffers > 0 && ca->null_count != 0) ? static_cast<const uint8_t*>(ca->buffers[0]) : nullptr;
  auto valid_at = [&](int64_t i) { return !validity || bit_at(validity, off + i); };

  ColumnImage& img = part->images[name];
  img.built = true;
  ImageWriter w(img.meta, img.extents);
  w.device_seeds = true;
  ChunkHost ch;
  ch.desc.n_rows = n_rows;
  const uint32_t T = 128;  // seed granularity (kIndexRows)
  const uint32_t n_chunks = (n_rows + T - 1) / T;

  // ---- validity -> definition levels (the Arrow bitmap is the bit-packed level stream, LSB first) ----
  uint32_t n_values = 0;
  std::vector<uint8_t> defbits((size_t(n_rows) + 7) / 8 + 8, 0);
  std::vector<uint32_t> val0(n_chunks);
  for (uint32_t i = 0; i < n_rows; i++) {
    if (i % T == 0) val0[i / T] = n_values;
    if (valid_at(i)) {
      defbits[i >> 3] |= uint8_t(1u << (i & 7));
      n_values++;
    }
  }
  const bool has_nulls = n_values != n_rows;
  ch.null_count = int64_t(n_rows) - int64_t(n_values);
  ch.desc.has_nulls = has_nulls ? 1 : 0;
  ch.desc.n_values = n_values;

  uint64_t val0_off = ~0ull;
  if (has_nulls) val0_off = uint64_t(w.section(val0.data(), val0.size() * 4));
  auto seed_job = [&](int64_t runs_off, uint32_t n_runs, uint32_t total, bool is_def) {
    SeedJob j{};
    j.runs_off = uint64_t(runs_off);
    j.seeds_off = w.reserve_dev(uint64_t(n_chunks) * sizeof(Seed));
    j.val0_off = val0_off;
    j.n_runs = n_runs;
    j.total = total;
    j.n_chunks = n_chunks;
    j.is_def = is_def ? 1 : 0;
    w.jobs.push_back(j);
    return int64_t(j.seeds_off);
  };

  if (is_i64 || is_f64) {
    ch.phys = is_i64 ? PT_INT64 : PT_DOUBLE;
    ch.desc.kind = CK_PLAIN64;
    const int64_t* src = static_cast<const int64_t*>(ca->buffers[1]) + off;
    std::vector<int64_t> packed;
    const int64_t* vals = src;
    if (has_nulls) {
      packed.reserve(n_values);
      for (uint32_t i = 0; i < n_rows; i++)
        if (valid_at(i)) packed.push_back(src[i]);
      vals = packed.data();
    }
    if (is_i64 && n_values > 0) {  // bounds for row-group pruning, as a Parquet writer would record them
      int64_t mn = vals[0], mx = vals[0];
      for (uint32_t i = 1; i < n_values; i++) { mn = std::min(mn, vals[i]); mx = std::max(mx, vals[i]); }
      ch.has_minmax = true;
      ch.min_bits = mn;
      ch.max_bits = mx;
    }
    ch.off_values = w.section(vals, size_t(n_values) * 8);
    ch.stored_bytes = uint64_t(n_values) * 8 + (has_nulls ? (n_rows + 7) / 8 : 0);
  } else {
    ch.phys = PT_BYTE_ARRAY;
    ch.desc.kind = CK_DICT_STR;
    GlobalDict* gd = &table->dicts[name];
    std::vector<uint32_t> idx;  // chunk-local index of every non-null row
    idx.reserve(n_values);
    if (is_dict) {
      const ::ArrowArray* da = ca->dictionary;
      const std::string dfmt = cs->dictionary->format ? cs->dictionary->format : "";
      if (dfmt != "u" && dfmt != "z") { *err = "column " + name + ": dictionary values must be binary/utf8, got '" + dfmt + "'"; return false; }
      const int32_t* doff = static_cast<const int32_t*>(da->buffers[1]) + da->offset;
      const char* dbytes = static_cast<const char*>(da->buffers[2]);
      ch.lut_host.reserve(size_t(da->length));
      for (int64_t i = 0; i < da->length; i++) ch.lut_host.push_back(gd->intern(dbytes ? dbytes + doff[i] : "", size_t(doff[i + 1] - doff[i])));
      for (uint32_t i = 0; i < n_rows; i++) {
        if (!valid_at(i)) continue;
        int64_t v = arrow_index(ca->buffers[1], fmt[0], off + i);
        if (v < 0 || v >= da->length) { *err = "column " + name + ": dictionary index out of range"; return false; }
        idx.push_back(uint32_t(v));
      }
    } else {
      const int32_t* soff = static_cast<const int32_t*>(ca->buffers[1]) + off;
      const char* sbytes = static_cast<const char*>(ca->buffers[2]);
      std::unordered_map<std::string, uint32_t> local;
      for (uint32_t i = 0; i < n_rows; i++) {
        if (!valid_at(i)) continue;
        std::string v(sbytes ? sbytes + soff[i] : "", size_t(soff[i + 1] - soff[i]));
        auto it = local.find(v);
        if (it == local.end()) {
          it = local.emplace(v, uint32_t(ch.lut_host.size())).first;
          ch.lut_host.push_back(gd->intern(v.data(), v.size()));
        }
        idx.push_back(it->second);
      }
    }
    ch.desc.dict_size = uint32_t(ch.lut_host.size());
    // index stream: raw 32-bit indices = one bit-packed run of width 32
    std::vector<HostRun> vruns;
    if (n_values > 0) vruns.push_back(HostRun{0, 0, 0, 1u | (32u << 8)});
    ch.desc.n_runs = uint32_t(vruns.size());
    ch.desc.n_bp_runs = uint32_t(vruns.size());
    vruns.push_back(HostRun{n_values, 0, 0, 0});
    ch.off_values = w.section(idx.data(), idx.size() * 4);
    ch.off_runs = w.section(vruns.data(), vruns.size() * sizeof(HostRun));
    ch.dev_seeds = seed_job(ch.off_runs, ch.desc.n_runs, n_values, false);
    ch.off_lut = w.section(ch.lut_host.data(), ch.lut_host.size() * 4);
    ch.stored_bytes = uint64_t(n_values) * 4 + (has_nulls ? (n_rows + 7) / 8 : 0);
  }
  if (has_nulls) {
    std::vector<HostRun> druns;
    druns.push_back(HostRun{0, 0, 0, 1u | (1u << 8)});
    ch.desc.n_defruns = 1;
    druns.push_back(HostRun{n_rows, 0, 0, 0});
    ch.off_def = w.section(defbits.data(), defbits.size());
    ch.off_def_runs = w.section(druns.data(), druns.size() * sizeof(HostRun));
    ch.dev_def_seeds = seed_job(ch.off_def_runs, 1, n_rows, true);
  }
  ch.meta_bytes = uint64_t(n_chunks) * sizeof(Seed) * ((ch.dev_seeds >= 0) + (ch.dev_def_seeds >= 0));
  part->rgs[0].cols[name] = std::move(ch);
  finish_image(w, img, part, name);
  return true;
}

}  // namespace

bool build_arrow_part(Table* table, Part* part, const ::ArrowSchema* schema, const ::ArrowArray* array, std::string* err) {
  if (!schema || !array || !schema->format || std::string(schema->format) != "+s") { *err = "Arrow part must be a struct array (record batch)"; return false; }
  if (schema->n_children != array->n_children) { *err = "Arrow schema / array children differ"; return false; }
  if (array->length < 0 || array->length > 0x7fffffffll) { *err = "record with more than 2^31 rows"; return false; }
  if (array->offset != 0 || array->null_count > 0) { *err = "sliced or nullable record structs are not supported"; return false; }
  const uint32_t n_rows = uint32_t(array->length);
  part->columns.clear();
  part->rgs.clear();
  if (n_rows == 0) return true;  // an empty record scans nothing
  RowGroupHost h;
  h.n_rows = n_rows;
  part->rgs.push_back(std::move(h));
  for (int64_t c = 0; c < schema->n_children; c++) {
    const ::ArrowSchema* cs = schema->children[c];
    const std::string name = cs->name ? cs->name : "";
    if (name.empty()) { *err = "unnamed column in Arrow part"; return false; }
    part->columns.push_back(name);
    if (!build_arrow_column(table, part, name, cs, array->children[c], n_rows, err)) return false;
  }
  return true;
}

void patch_column_pointers(Part* part, const std::string& column, const uint8_t* base) {
  for (RowGroupHost& rg : part->rgs) {
    auto it = rg.cols.find(column);
    if (it == rg.cols.end()) continue;
    ChunkHost& c = it->second;
    if (!c.error.empty()) continue;
    auto at = [&](int64_t off) -> const uint8_t* { return off < 0 ? nullptr : base + off; };
    c.desc.values = at(c.off_values);
    c.desc.runs = reinterpret_cast<const Run*>(at(c.off_runs));
    c.desc.seeds = reinterpret_cast<const Seed*>(at(c.off_seeds));
    c.desc.def = at(c.off_def);
    c.desc.def_runs = reinterpret_cast<const Run*>(at(c.off_def_runs));
    c.desc.def_seeds = reinterpret_cast<const Seed*>(at(c.off_def_seeds));
    c.desc.lut = reinterpret_cast<const uint32_t*>(at(c.off_lut));
    c.desc.dict64 = reinterpret_cast<const int64_t*>(at(c.off_dict64));
    if (c.off_row_runs >= 0) {
      c.desc.row_runs = reinterpret_cast<const Run*>(at(c.off_row_runs));
      c.desc.row_seeds = reinterpret_cast<const Seed*>(at(c.off_row_seeds));
    } else if (c.desc.kind == CK_DICT_STR && !c.desc.has_nulls && c.desc.n_bp_runs == 0) {
      c.desc.row_runs = c.desc.runs;  // no NULLs: value ordinals are rows
      c.desc.row_seeds = c.desc.seeds;
      c.desc.n_row_runs = c.desc.n_runs;
    } else {
      c.desc.row_runs = nullptr;
      c.desc.row_seeds = nullptr;
    }
  }
}

// ---- host-only description (tests without a GPU) ---------------------------------------------------
namespace {
void json_escape(std::ostringstream& o, const std::string& s) {
  o << '"';
  for (unsigned char ch : s) {
    if (ch == '"' || ch == '\\') o << '\\' << ch;
    else if (ch < 0x20 || ch >= 0x7f) {
      char buf[8];
      snprintf(buf, sizeof buf, "\\u%04x", ch);
      o << buf;
    } else o << ch;
  }
  o << '"';
}
}  // namespace

std::string describe_part_json(const uint8_t* file, uint64_t len, int tile_rows, std::string* err) {
  Table table;
  Part part;
  if (!open_part(file, len, &part, err)) return "";
  std::vector<std::vector<uint8_t>> host_images;  // one flat host copy per column (meta + extents)
  uint64_t image_bytes = 0;
  for (const std::string& name : part.columns) {
    build_column(tile_rows, &table, &part, name);
    ColumnImage& img = part.images[name];
    std::vector<uint8_t> flat(size_t(img.dev_bytes), 0);
    std::memcpy(flat.data(), img.meta.data(), img.meta.size());
    for (const Extent& e : img.extents) std::memcpy(flat.data() + e.dst_off, e.src, size_t(e.len));
    host_images.push_back(std::move(flat));
    patch_column_pointers(&part, name, host_images.back().data());  // pointers into the host image
    image_bytes += img.dev_bytes;
  }
  const uint32_t T = uint32_t(tile_rows);
  std::ostringstream o;
  o << "{\"image_bytes\":" << image_bytes << ",\"file_bytes\":" << len << ",\"row_groups\":[";
  for (size_t g = 0; g < part.rgs.size(); g++) {
    const RowGroupHost& rg = part.rgs[g];
    if (g) o << ',';
    o << "{\"n_rows\":" << rg.n_rows << ",\"columns\":{";
    bool firstc = true;
    for (const std::string& name : part.columns) {
      const ChunkHost& c = rg.cols.at(name);
      if (!firstc) o << ',';
      firstc = false;
      json_escape(o, name);
      o << ":{";
      if (!c.error.empty()) {
        o << "\"error\":";
        json_escape(o, c.error);
        o << '}';
        continue;
      }
      const ChunkDesc& d = c.desc;
      o << "\"kind\":" << int(d.kind) << ",\"has_nulls\":" << int(d.has_nulls) << ",\"n_values\":" << d.n_values
        << ",\"n_runs\":" << d.n_runs << ",\"n_defruns\":" << d.n_defruns << ",\"dict_size\":" << d.dict_size
        << ",\"stored_bytes\":" << c.stored_bytes << ",\"meta_bytes\":" << c.meta_bytes << ",\"null_count\":" << c.null_count;
      if (c.has_minmax && c.phys == PT_INT64) o << ",\"min\":" << c.min_bits << ",\"max\":" << c.max_bits;
      // Decode through the same directories / tile indexes the kernel uses (small inputs only).
      if (rg.n_rows <= 65536) {
        std::vector<HostRun> vruns, druns;
        if (d.runs) vruns.assign(reinterpret_cast<const HostRun*>(d.runs), reinterpret_cast<const HostRun*>(d.runs) + d.n_runs + 1);
        if (d.def_runs) druns.assign(reinterpret_cast<const HostRun*>(d.def_runs), reinterpret_cast<const HostRun*>(d.def_runs) + d.n_defruns + 1);
        const GlobalDict* gd = nullptr;
        if (d.kind == CK_DICT_STR) gd = &table.dicts.at(name);
        o << ",\"decoded\":[";
        uint32_t vord = 0;
        bool tiles_ok = true;
        for (uint32_t r = 0; r < rg.n_rows; r++) {
          if (r) o << ',';
          if (r % T == 0) {
            auto seed_ok = [&](const Seed& sd, const std::vector<HostRun>& runs, uint32_t first, uint32_t total) {
              if (first >= total) return true;
              if (sd.k >= runs.size() - 1) return false;
              const HostRun& rr = runs[sd.k];
              return rr.start <= first && runs[sd.k + 1].start > first && sd.start == rr.start && sd.end == runs[sd.k + 1].start &&
                     sd.off == rr.off && sd.val == rr.val && sd.meta == rr.meta;
            };
            if (d.kind != CK_PLAIN64) {
              const Seed& sd = d.seeds[r / T];
              if (sd.val0 != vord || !seed_ok(sd, vruns, vord, d.n_values)) tiles_ok = false;
            }
            if (d.has_nulls) {
              const Seed& sd = d.def_seeds[r / T];
              if (sd.val0 != vord || !seed_ok(sd, druns, r, rg.n_rows)) tiles_ok = false;
            }
          }
          bool valid = true;
          if (d.has_nulls) valid = hybrid_value_at(d.def, druns, r) != 0;
          if (!valid) { o << "null"; continue; }
          if (d.kind == CK_PLAIN64) {
            int64_t v;
            std::memcpy(&v, d.values + size_t(vord) * 8, 8);
            if (c.phys == PT_DOUBLE) {
              double f;
              std::memcpy(&f, &v, 8);
              char buf[40];
              snprintf(buf, sizeof buf, "%.17g", f);
              o << buf;
            } else o << v;
          } else {
            bool was_rle = false;
            uint32_t idx = hybrid_value_at(d.values, vruns, vord, &was_rle);
            if (d.kind == CK_DICT_STR) json_escape(o, gd->values[was_rle ? idx : d.lut[idx]]);
            else if (c.phys == PT_DOUBLE) {
              double f;
              std::memcpy(&f, &d.dict64[idx], 8);
              char buf[40];
              snprintf(buf, sizeof buf, "%.17g", f);
              o << buf;
            } else o << d.dict64[idx];
          }
          vord++;
        }
        o << "],\"tile_index_ok\":" << (tiles_ok ? "true" : "false");
        // the row-space directory (sorted-run kernel) must say the same as the two streams it was merged from
        if (d.row_runs) {
          std::vector<HostRun> rr(reinterpret_cast<const HostRun*>(d.row_runs), reinterpret_cast<const HostRun*>(d.row_runs) + d.n_row_runs + 1);
          bool ok = rr.back().start == rg.n_rows;
          uint32_t vo = 0;
          for (uint32_t r = 0; r < rg.n_rows && ok; r++) {
            const uint32_t got = hybrid_value_at(nullptr, rr, r);
            const bool valid = !d.has_nulls || hybrid_value_at(d.def, druns, r) != 0;
            if (!valid) { ok = got == 0xffffffffu; continue; }
            bool was_rle = false;
            uint32_t idx = hybrid_value_at(d.values, vruns, vo++, &was_rle);
            ok = got == (was_rle ? idx : d.lut[idx]);
            if (r % T == 0) {
              const Seed& sd = d.row_seeds[r / T];
              ok = ok && sd.k < rr.size() - 1 && rr[sd.k].start <= r && rr[sd.k + 1].start > r && sd.end == rr[sd.k + 1].start && sd.val == rr[sd.k].val;
            }
          }
          o << ",\"n_row_runs\":" << d.n_row_runs << ",\"row_runs_ok\":" << (ok ? "true" : "false");
        }
      }
      o << '}';
    }
    o << "}}";
  }
  o << "]}";
  return o.str();
}

}  // namespace fgpu
