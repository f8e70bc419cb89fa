/*
 * frost_oracle.c — CPU restatement of FrostDB's scan -> filter -> hash-aggregate / distinct path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under frostdb_b200/ links, loads or calls this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: the reference is Go and cannot be compiled here (no Go toolchain), so this port
 * is pinned against the reference's own golden vectors transcribed in tests/golden/ (logictest
 * exec files, aggregate_test.go, expr/binaryscalarexpr_test.go) and cross-checked against pyarrow for
 * the Parquet decoding, which in the reference lives in the un-vendored third-party module
 * github.com/parquet-go/parquet-go v0.24.0 (go.mod) and is restated here from the public Parquet
 * format specification.  Semantics with no reference test behind them are marked UNPINNED.
 *
 * The code follows the reference's algorithm shape on purpose (it is also the CPU baseline):
 *   scan       -> visible parts, row groups ruled out by the filter's  index/lsm.go:401-454
 *                 TrueNegativeFilter over the chunk statistics        query/expr/{filter,binaryscalarexpr}.go
 *                 (pinned by TestBinaryScalarOperation, expr/binaryscalarexpr_test.go:56-212 ->
 *                 tests/golden/rowgroup_filter_cases.py; it never changes a result); `==` asks the
 *                 chunk's split-block bloom filter when it has one, binaryscalarexpr.go:104-118
 *                 (parquet-go's filter restated from the Parquet format's BloomFilter.md: XXH64 of the
 *                 PLAIN-encoded value, 8 salted bits in one 32-byte block; pinned by the xxHash known
 *                 answers and a plain-Python second implementation, tests/bloom_file.py)
 *   row group  -> decode projected columns to Arrow-like arrays    pqarrow/arrow.go:264-373,711-823
 *                 dictionary columns: one memo-table insert per row pqarrow/writer/writer.go:381-405
 *   filter     -> leaf bitmaps, AND/OR, compaction of all columns   query/physicalplan/filter.go:167-323
 *                                                                   binaryscalarexpr.go:41-311
 *   aggregate  -> per row hash of the key values, map lookup,        aggregate.go:263-490
 *                 append the value to the group's buffer            dynparquet/hashed.go:86-272
 *   finish     -> reduce every group's buffer                        aggregate.go:527-633,734-971
 *   workers    -> one chain per thread over a shared row-group queue, table.go:760-865
 *                 partial results merged by carried hash            physicalplan.go:438-471, synchronize.go
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/frostgpu.h" /* plan descriptor structs only (the interface being tested) */

/* ------------------------------------------------------------------------------------------------ */
/* small utilities                                                                                   */
/* ------------------------------------------------------------------------------------------------ */
static void* xmalloc(size_t n) {
  void* p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}
static void* xcalloc(size_t n, size_t m) {
  void* p = calloc(n ? n : 1, m ? m : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}
static void* xrealloc(void* q, size_t n) {
  void* p = realloc(q, n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}

/* 64-bit byte-string hash standing in for metro.Hash64(value, 0) (hashed.go:207): hash values are
 * never observable in results (random maphash seed per query, physicalplan.go:447), only equality. */
static uint64_t hash_bytes(const uint8_t* p, size_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xD6E8FEB86659FD93ull);
  while (n >= 8) {
    uint64_t k;
    memcpy(&k, p, 8);
    k *= 0xC2B2AE3D27D4EB4Full;
    k = (k << 31) | (k >> 33);
    k *= 0x9E3779B185EBCA87ull;
    h ^= k;
    h = ((h << 27) | (h >> 37)) * 5 + 0x52DCE729;
    p += 8;
    n -= 8;
  }
  uint64_t t = 0;
  for (size_t i = 0; i < n; i++) t |= (uint64_t)p[i] << (8 * i);
  h ^= t * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 33;
  h *= 0xFF51AFD7ED558CCDull;
  h ^= h >> 33;
  h *= 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 33;
  if (h == 0) h = 1; /* 0 means NULL in the group hash (aggregate.go:401-403) */
  return h;
}
/* hashCombine, aggregate.go:245-247 */
static inline uint64_t hash_combine(uint64_t lhs, uint64_t rhs) {
  return lhs ^ (rhs + 0x9e3779b9ull + (lhs << 6) + (lhs >> 2));
}

/* ------------------------------------------------------------------------------------------------ */
/* Thrift compact protocol (Parquet metadata)                                                        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { const uint8_t* p; const uint8_t* end; int err; } trd;

static uint8_t t_byte(trd* r) { if (r->p >= r->end) { r->err = 1; return 0; } return *r->p++; }
static uint64_t t_uvar(trd* r) {
  uint64_t v = 0; int sh = 0;
  for (;;) { uint8_t b = t_byte(r); if (r->err) return 0; v |= (uint64_t)(b & 0x7f) << sh; if (!(b & 0x80)) return v; sh += 7; if (sh > 63) { r->err = 1; return 0; } }
}
static int64_t t_zz(trd* r) { uint64_t u = t_uvar(r); return (int64_t)(u >> 1) ^ -(int64_t)(u & 1); }
static void t_skipn(trd* r, uint64_t n) { if (n > (uint64_t)(r->end - r->p)) { r->err = 1; return; } r->p += n; }
static int t_field(trd* r, int* last, int* id, int* type) {
  uint8_t h = t_byte(r);
  if (r->err || h == 0) return 0;
  int delta = h >> 4; *type = h & 15;
  if (delta == 0) *id = (int)t_zz(r); else *id = *last + delta;
  *last = *id;
  return 1;
}
static void t_list(trd* r, int* et, uint32_t* n) { uint8_t h = t_byte(r); *et = h & 15; uint32_t s = h >> 4; if (s == 15) s = (uint32_t)t_uvar(r); *n = s; }
static void t_skip(trd* r, int type, int depth) {
  if (r->err || depth > 32) { r->err = 1; return; }
  switch (type) {
    case 1: case 2: return;
    case 3: t_byte(r); return;
    case 4: case 5: case 6: t_uvar(r); return;
    case 7: t_skipn(r, 8); return;
    case 8: t_skipn(r, t_uvar(r)); return;
    case 9: case 10: { int et; uint32_t n; t_list(r, &et, &n); for (uint32_t i = 0; i < n && !r->err; i++) { if (et == 1 || et == 2) t_byte(r); else t_skip(r, et, depth + 1); } return; }
    case 11: { uint64_t n = t_uvar(r); if (!n) return; uint8_t kv = t_byte(r); for (uint64_t i = 0; i < n && !r->err; i++) { t_skip(r, kv >> 4, depth + 1); t_skip(r, kv & 15, depth + 1); } return; }
    case 12: { int last = 0, id, t; while (t_field(r, &last, &id, &t)) t_skip(r, t, depth + 1); return; }
    default: r->err = 1;
  }
}
static char* t_string(trd* r) {
  uint64_t n = t_uvar(r);
  if (r->err || n > (uint64_t)(r->end - r->p)) { r->err = 1; return NULL; }
  char* s = xmalloc(n + 1);
  memcpy(s, r->p, n); s[n] = 0; r->p += n;
  return s;
}

/* ------------------------------------------------------------------------------------------------ */
/* Parquet file model                                                                                */
/* ------------------------------------------------------------------------------------------------ */
enum { PQ_INT64 = 2, PQ_DOUBLE = 5, PQ_BYTE_ARRAY = 6 };

typedef struct {
  char* name;   /* dotted path */
  int phys;
  int optional; /* max definition level (0/1) */
  int repeated;
} o_leaf;

typedef struct {
  int phys, codec;
  int64_t num_values, total_compressed, data_off, dict_off;
  /* chunk statistics (the bounds and null count the reference reads through the ColumnIndex) */
  int64_t null_count;                  /* -1: not recorded */
  const uint8_t *smin, *smax;          /* PLAIN-encoded bounds inside the footer, NULL: not recorded */
  uint32_t smin_len, smax_len;
  /* split-block bloom filter (parquet-go writes one per sorting column, dynparquet/schema.go:1111-1157) */
  int64_t bloom_off;                   /* ColumnMetaData.bloom_filter_offset, -1: none */
  const uint8_t* bloom;                /* the bitset inside the file, NULL: none / not BLOCK + XXHASH + UNCOMPRESSED */
  uint32_t bloom_bytes;
} o_chunk;

typedef struct { int64_t num_rows; o_chunk* chunks; } o_rg;

typedef struct {
  const uint8_t* file; uint64_t len; uint64_t tx;
  int n_leaves; o_leaf* leaves;
  int n_rgs; o_rg* rgs;
} o_part;

struct oracle_table { int n_parts, cap; o_part** parts; char err[512]; };
typedef struct oracle_table oracle_table;

static int parse_footer(o_part* part, char* err) {
  const uint8_t* f = part->file; uint64_t len = part->len;
  if (len < 12 || memcmp(f, "PAR1", 4) || memcmp(f + len - 4, "PAR1", 4)) { snprintf(err, 512, "not a parquet file"); return -1; }
  uint32_t flen; memcpy(&flen, f + len - 8, 4);
  if ((uint64_t)flen + 12 > len) { snprintf(err, 512, "bad footer length"); return -1; }
  trd r = { f + len - 8 - flen, f + len - 8, 0 };
  /* schema elements are collected raw, then flattened */
  typedef struct { int type, rep, nchild; char* name; } se;
  se* schema = NULL; int n_schema = 0;
  int last = 0, id, t;
  while (t_field(&r, &last, &id, &t)) {
    if (id == 2) {
      int et; uint32_t n; t_list(&r, &et, &n);
      schema = xcalloc(n, sizeof(se)); n_schema = (int)n;
      for (uint32_t i = 0; i < n; i++) {
        se* e = &schema[i]; e->type = -1;
        int l2 = 0, id2, t2;
        while (t_field(&r, &l2, &id2, &t2)) {
          if (id2 == 1) e->type = (int)t_zz(&r);
          else if (id2 == 3) e->rep = (int)t_zz(&r);
          else if (id2 == 4) e->name = t_string(&r);
          else if (id2 == 5) e->nchild = (int)t_zz(&r);
          else t_skip(&r, t2, 0);
        }
      }
    } else if (id == 4) {
      int et; uint32_t n; t_list(&r, &et, &n);
      part->rgs = xcalloc(n, sizeof(o_rg)); part->n_rgs = (int)n;
      for (uint32_t g = 0; g < n; g++) {
        o_rg* rg = &part->rgs[g];
        int l2 = 0, id2, t2;
        while (t_field(&r, &l2, &id2, &t2)) {
          if (id2 == 1) {
            int et2; uint32_t nc; t_list(&r, &et2, &nc);
            rg->chunks = xcalloc(nc, sizeof(o_chunk));
            for (uint32_t c = 0; c < nc; c++) {
              o_chunk* ch = &rg->chunks[c]; ch->data_off = -1; ch->dict_off = -1; ch->null_count = -1; ch->bloom_off = -1;
              int l3 = 0, id3, t3;
              while (t_field(&r, &l3, &id3, &t3)) {
                if (id3 == 3) {
                  int l4 = 0, id4, t4;
                  while (t_field(&r, &l4, &id4, &t4)) {
                    if (id4 == 1) ch->phys = (int)t_zz(&r);
                    else if (id4 == 4) ch->codec = (int)t_zz(&r);
                    else if (id4 == 5) ch->num_values = t_zz(&r);
                    else if (id4 == 7) ch->total_compressed = t_zz(&r);
                    else if (id4 == 9) ch->data_off = t_zz(&r);
                    else if (id4 == 11) ch->dict_off = t_zz(&r);
                    else if (id4 == 14) ch->bloom_off = t_zz(&r);
                    else if (id4 == 12) { /* Statistics: 5/6 max_value/min_value, 1/2 the deprecated pair, 3 null_count */
                      const uint8_t *omin = NULL, *omax = NULL; uint32_t ominl = 0, omaxl = 0;
                      int l5 = 0, id5, t5;
                      while (t_field(&r, &l5, &id5, &t5)) {
                        if (id5 == 3) ch->null_count = t_zz(&r);
                        else if ((id5 == 1 || id5 == 2 || id5 == 5 || id5 == 6) && t5 == 8) {
                          uint64_t bl = t_uvar(&r); const uint8_t* bp = r.p; t_skipn(&r, bl);
                          if (id5 == 5) { ch->smax = bp; ch->smax_len = (uint32_t)bl; }
                          else if (id5 == 6) { ch->smin = bp; ch->smin_len = (uint32_t)bl; }
                          else if (id5 == 1) { omax = bp; omaxl = (uint32_t)bl; }
                          else { omin = bp; ominl = (uint32_t)bl; }
                        } else t_skip(&r, t5, 0);
                      }
                      if (!ch->smin || !ch->smax) { ch->smin = omin; ch->smin_len = ominl; ch->smax = omax; ch->smax_len = omaxl; }
                    }
                    else t_skip(&r, t4, 0);
                  }
                } else t_skip(&r, t3, 0);
              }
            }
          } else if (id2 == 3) rg->num_rows = t_zz(&r);
          else t_skip(&r, t2, 0);
        }
      }
    } else t_skip(&r, t, 0);
  }
  if (r.err || n_schema == 0) { snprintf(err, 512, "malformed footer"); return -1; }
  /* flatten: depth-first, dotted names */
  part->leaves = xcalloc((size_t)n_schema, sizeof(o_leaf));
  struct { int remaining; char* prefix; int def, rep; } stack[32];
  int sp = 0;
  stack[0].remaining = schema[0].nchild; stack[0].prefix = NULL; stack[0].def = 0; stack[0].rep = 0;
  for (int i = 1; i < n_schema; i++) {
    while (sp >= 0 && stack[sp].remaining == 0) sp--;
    if (sp < 0) { snprintf(err, 512, "schema tree malformed"); return -1; }
    stack[sp].remaining--;
    se* e = &schema[i];
    int def = stack[sp].def + (e->rep != 0), rep = stack[sp].rep + (e->rep == 2);
    size_t pl = stack[sp].prefix ? strlen(stack[sp].prefix) + 1 : 0;
    char* path = xmalloc(pl + strlen(e->name) + 1);
    if (pl) { strcpy(path, stack[sp].prefix); strcat(path, "."); strcat(path, e->name); } else strcpy(path, e->name);
    if (e->nchild > 0) {
      if (sp + 1 >= 32) { snprintf(err, 512, "schema too deep"); return -1; }
      sp++; stack[sp].remaining = e->nchild; stack[sp].prefix = path; stack[sp].def = def; stack[sp].rep = rep;
    } else {
      o_leaf* l = &part->leaves[part->n_leaves++];
      l->name = path; l->phys = e->type; l->optional = def; l->repeated = rep;
    }
  }
  for (int i = 0; i < n_schema; i++) free(schema[i].name);
  free(schema);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Column decode: one row group column -> Arrow-like array                                           */
/* ------------------------------------------------------------------------------------------------ */
enum { C_ABSENT = 0, C_I64 = 1, C_F64 = 2, C_DICT = 3 };

typedef struct {
  int type;
  int64_t n;
  uint8_t* valid;   /* one byte per row; NULL = all valid */
  int64_t* i64;     /* C_I64 / C_F64 (raw bits); NULL slots hold 0 (optbuilders.go:337-340) */
  uint32_t* idx;    /* C_DICT: index into dict, undefined for NULL rows */
  /* record-level dictionary built by memoisation, like array.BinaryDictionaryBuilder */
  uint32_t n_dict, cap_dict;
  const uint8_t** dval; uint32_t* dlen;
  uint32_t* memo; uint32_t memo_cap;
} o_col;

static void col_free(o_col* c) {
  free(c->valid); free(c->i64); free(c->idx); free(c->dval); free(c->dlen); free(c->memo);
  memset(c, 0, sizeof *c);
}

/* BinaryDictionaryBuilder.Append: memo-table lookup/insert of the value bytes (one per row). */
static uint32_t dict_append(o_col* c, const uint8_t* v, uint32_t len) {
  if ((c->n_dict + 1) * 2 > c->memo_cap) {
    uint32_t nc = c->memo_cap ? c->memo_cap * 2 : 64;
    uint32_t* nm = xmalloc(nc * sizeof(uint32_t));
    memset(nm, 0xff, nc * sizeof(uint32_t));
    for (uint32_t i = 0; i < c->n_dict; i++) {
      uint32_t s = (uint32_t)hash_bytes(c->dval[i], c->dlen[i]) & (nc - 1);
      while (nm[s] != 0xffffffffu) s = (s + 1) & (nc - 1);
      nm[s] = i;
    }
    free(c->memo); c->memo = nm; c->memo_cap = nc;
  }
  uint32_t s = (uint32_t)hash_bytes(v, len) & (c->memo_cap - 1);
  for (;;) {
    uint32_t e = c->memo[s];
    if (e == 0xffffffffu) break;
    if (c->dlen[e] == len && memcmp(c->dval[e], v, len) == 0) return e;
    s = (s + 1) & (c->memo_cap - 1);
  }
  if (c->n_dict == c->cap_dict) {
    c->cap_dict = c->cap_dict ? c->cap_dict * 2 : 32;
    c->dval = xrealloc(c->dval, c->cap_dict * sizeof(*c->dval));
    c->dlen = xrealloc(c->dlen, c->cap_dict * sizeof(*c->dlen));
  }
  c->dval[c->n_dict] = v; c->dlen[c->n_dict] = len;
  c->memo[s] = c->n_dict;
  return c->n_dict++;
}

/* Sequential RLE / bit-packed hybrid reader (Parquet "RLE" encoding). */
typedef struct { const uint8_t* p; const uint8_t* end; int w; uint32_t rle_left, bp_left; uint32_t rle_val; uint64_t bitbuf; int bitcnt; } hyb;
static void hyb_init(hyb* h, const uint8_t* p, const uint8_t* end, int w) { memset(h, 0, sizeof *h); h->p = p; h->end = end; h->w = w; }
static int hyb_next(hyb* h, uint32_t* out) {
  for (;;) {
    if (h->rle_left) { h->rle_left--; *out = h->rle_val; return 0; }
    if (h->bp_left) {
      while (h->bitcnt < h->w) { uint64_t b = (h->p < h->end) ? *h->p : 0; h->p++; h->bitbuf |= b << h->bitcnt; h->bitcnt += 8; }
      *out = (uint32_t)(h->bitbuf & ((h->w >= 32) ? 0xffffffffull : ((1ull << h->w) - 1)));
      h->bitbuf >>= h->w; h->bitcnt -= h->w; h->bp_left--;
      if (h->bp_left == 0) { h->bitbuf = 0; h->bitcnt = 0; }
      return 0;
    }
    uint64_t hd = 0; int sh = 0;
    for (;;) { if (h->p >= h->end) return -1; uint8_t b = *h->p++; hd |= (uint64_t)(b & 0x7f) << sh; if (!(b & 0x80)) break; sh += 7; }
    if (hd & 1) { h->bp_left = (uint32_t)((hd >> 1) * 8); h->bitbuf = 0; h->bitcnt = 0; if (h->w == 0) { /* zero-width: all zeros */ } }
    else {
      h->rle_left = (uint32_t)(hd >> 1);
      uint32_t v = 0; int nb = (h->w + 7) / 8;
      for (int i = 0; i < nb; i++) { if (h->p >= h->end) return -1; v |= (uint32_t)(*h->p++) << (8 * i); }
      h->rle_val = v;
    }
    if (!h->rle_left && !h->bp_left) return -1;
  }
}

typedef struct { int type, uncompressed, compressed, num_values, encoding, def_enc, num_nulls, def_len, rep_len; } pghdr;

static int read_page_header(trd* r, pghdr* h) {
  memset(h, 0, sizeof *h); h->type = -1; h->num_nulls = -1; h->def_enc = 3;
  int last = 0, id, t;
  while (t_field(r, &last, &id, &t)) {
    if (id == 1) h->type = (int)t_zz(r);
    else if (id == 2) h->uncompressed = (int)t_zz(r);
    else if (id == 3) h->compressed = (int)t_zz(r);
    else if (id == 5) { int l2 = 0, id2, t2; while (t_field(r, &l2, &id2, &t2)) { if (id2 == 1) h->num_values = (int)t_zz(r); else if (id2 == 2) h->encoding = (int)t_zz(r); else if (id2 == 3) h->def_enc = (int)t_zz(r); else t_skip(r, t2, 0); } }
    else if (id == 7) { int l2 = 0, id2, t2; while (t_field(r, &l2, &id2, &t2)) { if (id2 == 1) h->num_values = (int)t_zz(r); else if (id2 == 2) h->encoding = (int)t_zz(r); else t_skip(r, t2, 0); } }
    else if (id == 8) { int l2 = 0, id2, t2; while (t_field(r, &l2, &id2, &t2)) { if (id2 == 1) h->num_values = (int)t_zz(r); else if (id2 == 2) h->num_nulls = (int)t_zz(r); else if (id2 == 4) h->encoding = (int)t_zz(r); else if (id2 == 5) h->def_len = (int)t_zz(r); else if (id2 == 6) h->rep_len = (int)t_zz(r); else t_skip(r, t2, 0); } }
    else t_skip(r, t, 0);
  }
  return r->err ? -1 : 0;
}

/* writeColumnToArray (pqarrow/arrow.go:711-823): page loop, values appended through the builder. */
static int decode_chunk(const o_part* part, const o_leaf* leaf, const o_chunk* ch, int64_t rg_rows, o_col* out, char* err) {
  memset(out, 0, sizeof *out);
  if (ch->codec != 0) { snprintf(err, 512, "column %s: compressed chunks unsupported", leaf->name); return -1; }
  if (leaf->repeated || leaf->optional > 1) { snprintf(err, 512, "column %s: nested columns unsupported", leaf->name); return -1; }
  if (leaf->phys != PQ_INT64 && leaf->phys != PQ_DOUBLE && leaf->phys != PQ_BYTE_ARRAY) { snprintf(err, 512, "column %s: unsupported physical type %d", leaf->name, leaf->phys); return -1; }
  out->type = leaf->phys == PQ_BYTE_ARRAY ? C_DICT : (leaf->phys == PQ_DOUBLE ? C_F64 : C_I64);
  out->n = rg_rows;
  if (leaf->optional) out->valid = xmalloc((size_t)rg_rows);
  if (out->type == C_DICT) out->idx = xmalloc((size_t)rg_rows * 4); else out->i64 = xcalloc((size_t)rg_rows, 8);
  int64_t start = ch->data_off;
  if (ch->dict_off > 0 && ch->dict_off < start) start = ch->dict_off;
  if (start < 4 || (uint64_t)start >= part->len) { snprintf(err, 512, "column %s: bad chunk offset", leaf->name); return -1; }
  const uint8_t* p = part->file + start;
  const uint8_t* end = part->file + part->len;
  /* page dictionary (PLAIN): entry pointers */
  const uint8_t** pd_val = NULL; uint32_t* pd_len = NULL; uint32_t pd_n = 0;
  const uint8_t* pd_num = NULL;
  int64_t row = 0; int rc = 0;
  while (row < rg_rows) {
    trd r = { p, end, 0 };
    pghdr h;
    if (read_page_header(&r, &h) || h.compressed < 0 || (uint64_t)h.compressed > (uint64_t)(end - r.p)) { snprintf(err, 512, "column %s: bad page header", leaf->name); rc = -1; break; }
    const uint8_t* pay = r.p; const uint8_t* pend = pay + h.compressed;
    if (h.type == 2) { /* dictionary page */
      if (leaf->phys == PQ_BYTE_ARRAY) {
        pd_val = xmalloc((size_t)h.num_values * sizeof(*pd_val)); pd_len = xmalloc((size_t)h.num_values * 4); pd_n = (uint32_t)h.num_values;
        const uint8_t* q = pay;
        for (uint32_t i = 0; i < pd_n; i++) { uint32_t l; if (pend - q < 4) { rc = -1; break; } memcpy(&l, q, 4); q += 4; if (l > (uint64_t)(pend - q)) { rc = -1; break; } pd_val[i] = q; pd_len[i] = l; q += l; }
        if (rc) { snprintf(err, 512, "column %s: bad dictionary page", leaf->name); break; }
      } else { pd_num = pay; pd_n = (uint32_t)h.num_values; }
    } else if (h.type == 0 || h.type == 3) {
      const uint8_t* q = pay; const uint8_t* defp = NULL; uint32_t deflen = 0;
      if (h.type == 0) {
        if (leaf->optional) { uint32_t dl; memcpy(&dl, q, 4); q += 4; defp = q; deflen = dl; q += dl; }
      } else { defp = pay; deflen = (uint32_t)h.def_len; q = pay + h.def_len + h.rep_len; }
      int nvals = h.num_values;
      if (row + nvals > rg_rows) { snprintf(err, 512, "column %s: pages overrun row group", leaf->name); rc = -1; break; }
      hyb dh; if (leaf->optional) hyb_init(&dh, defp, defp + deflen, 1);
      int is_dict = (h.encoding == 8 || h.encoding == 2);
      if (!is_dict && h.encoding != 0) { snprintf(err, 512, "column %s: unsupported encoding %d", leaf->name, h.encoding); rc = -1; break; }
      if (leaf->phys == PQ_BYTE_ARRAY && !is_dict) { snprintf(err, 512, "column %s: PLAIN strings unsupported", leaf->name); rc = -1; break; }
      hyb vh; int vh_init = 0;
      for (int i = 0; i < nvals; i++) {
        uint32_t d = 1;
        if (leaf->optional) { if (hyb_next(&dh, &d)) { rc = -1; break; } out->valid[row + i] = (uint8_t)d; }
        if (!d) { if (out->idx) out->idx[row + i] = 0; continue; } /* AppendNull */
        if (is_dict) {
          if (!vh_init) { if (q >= pend) { rc = -1; break; } int w = *q; hyb_init(&vh, q + 1, pend, w); vh_init = 1; }
          uint32_t ix; if (hyb_next(&vh, &ix) || ix >= pd_n) { rc = -1; break; }
          if (out->type == C_DICT) out->idx[row + i] = dict_append(out, pd_val[ix], pd_len[ix]);
          else memcpy(&out->i64[row + i], pd_num + (size_t)ix * 8, 8);
        } else {
          if (pend - q < 8) { rc = -1; break; }
          memcpy(&out->i64[row + i], q, 8); q += 8;
        }
      }
      if (rc) { snprintf(err, 512, "column %s: page decode failed", leaf->name); break; }
      row += nvals;
    }
    p = pend;
  }
  free(pd_val); free(pd_len);
  if (rc) col_free(out);
  return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* Query model                                                                                       */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { const char* name; o_col col; int present; } rec_col;
typedef struct { int64_t n; int n_cols; rec_col* cols; int qidx; } record; /* one Arrow record (one row group) */

static rec_col* rec_find(record* r, const char* name) {
  for (int i = 0; i < r->n_cols; i++) if (r->cols[i].present && strcmp(r->cols[i].name, name) == 0) return &r->cols[i];
  return NULL;
}

/* ---- filter (filter.go:167-215, binaryscalarexpr.go:41-311, regexpfilter.go:17-165) -------------- */
static int str_contains(const uint8_t* h, uint32_t hl, const uint8_t* n, uint64_t nl) {
  if (nl == 0) return 1;
  if (nl > hl) return 0;
  return memmem(h, hl, n, nl) != NULL;
}

static int eval_leaf(const fgpu_plan* plan, const fgpu_expr* e, record* rec, uint8_t* bm, char* err) {
  const fgpu_expr* l = &plan->exprs[e->left];
  const fgpu_expr* rt = &plan->exprs[e->right];
  if (l->kind != FGPU_EXPR_COLUMN || rt->kind != FGPU_EXPR_LITERAL) { snprintf(err, 512, "left side of binary expression must be a column"); return -1; }
  const fgpu_scalar* lit = &rt->literal;
  int64_t n = rec->n;
  rec_col* rc = rec_find(rec, l->name);
  int op = e->op;
  if (!rc) { /* missing column: binaryscalarexpr.go:47-73, regexpfilter.go:23-33 */
    int all = 1;
    if (op == FGPU_OP_EQ) { if (lit->type == FGPU_SCALAR_STRING && lit->len != 0) all = 0; }
    else if (op == FGPU_OP_NOT_EQ) { if (lit->type == FGPU_SCALAR_NULL) all = 0; }
    else if (op == FGPU_OP_LT || op == FGPU_OP_LT_EQ || op == FGPU_OP_GT || op == FGPU_OP_GT_EQ) all = 0;
    else if (op == FGPU_OP_REGEX_MATCH || op == FGPU_OP_REGEX_NOT_MATCH) {
      int em = e->match(e->match_user, (const uint8_t*)"", 0) != 0;
      all = (op == FGPU_OP_REGEX_NOT_MATCH) ? !em : em;
    }
    memset(bm, all, (size_t)n);
    return 0;
  }
  o_col* c = &rc->col;
  if (c->type == C_DICT) {
    if (op == FGPU_OP_LT || op == FGPU_OP_LT_EQ || op == FGPU_OP_GT || op == FGPU_OP_GT_EQ) { snprintf(err, 512, "unsupported operator"); return -1; }
    int is_null_lit = lit->type == FGPU_SCALAR_NULL;
    for (int64_t i = 0; i < n; i++) {
      int isnull = c->valid && !c->valid[i];
      int res;
      if (op == FGPU_OP_EQ && is_null_lit) res = isnull;                                         /* :205-212 */
      else if ((op == FGPU_OP_NOT_EQ || op == FGPU_OP_CONTAINS || op == FGPU_OP_NOT_CONTAINS) && is_null_lit) res = !isnull; /* :165-172, :282-289 */
      else if (isnull) res = 0;
      else {
        const uint8_t* v = c->dval[c->idx[i]]; uint32_t vl = c->dlen[c->idx[i]];
        switch (op) {
          case FGPU_OP_EQ: res = (vl == lit->len && memcmp(v, lit->bytes, vl) == 0); break;
          case FGPU_OP_NOT_EQ: res = !(vl == lit->len && memcmp(v, lit->bytes, vl) == 0); break;
          case FGPU_OP_CONTAINS: res = str_contains(v, vl, lit->bytes, lit->len); break;
          case FGPU_OP_NOT_CONTAINS: res = !str_contains(v, vl, lit->bytes, lit->len); break;
          case FGPU_OP_REGEX_MATCH: res = e->match(e->match_user, v, vl) != 0; break;
          case FGPU_OP_REGEX_NOT_MATCH: res = e->match(e->match_user, v, vl) == 0; break;
          default: snprintf(err, 512, "unsupported operator on dictionary column"); return -1;
        }
      }
      bm[i] = (uint8_t)res;
    }
    return 0;
  }
  /* numeric: arrow compute comparison, NULL -> not selected (:143-150) */
  if (op < FGPU_OP_EQ || op > FGPU_OP_GT_EQ) { snprintf(err, 512, "unsupported operator on numeric column"); return -1; }
  if (lit->type == FGPU_SCALAR_NULL) { memset(bm, 0, (size_t)n); return 0; }
  if (lit->type == FGPU_SCALAR_STRING) { snprintf(err, 512, "numeric column compared with string"); return -1; }
  int as_float = c->type == C_F64 || lit->type == FGPU_SCALAR_FLOAT64;
  double lf = lit->type == FGPU_SCALAR_FLOAT64 ? lit->f64 : (double)lit->i64;
  for (int64_t i = 0; i < n; i++) {
    if (c->valid && !c->valid[i]) { bm[i] = 0; continue; }
    int res;
    if (as_float) {
      double x; if (c->type == C_F64) memcpy(&x, &c->i64[i], 8); else x = (double)c->i64[i];
      switch (op) { case FGPU_OP_EQ: res = x == lf; break; case FGPU_OP_NOT_EQ: res = x != lf; break; case FGPU_OP_LT: res = x < lf; break; case FGPU_OP_LT_EQ: res = x <= lf; break; case FGPU_OP_GT: res = x > lf; break; default: res = x >= lf; }
    } else {
      int64_t x = c->i64[i], y = lit->i64;
      switch (op) { case FGPU_OP_EQ: res = x == y; break; case FGPU_OP_NOT_EQ: res = x != y; break; case FGPU_OP_LT: res = x < y; break; case FGPU_OP_LT_EQ: res = x <= y; break; case FGPU_OP_GT: res = x > y; break; default: res = x >= y; }
    }
    bm[i] = (uint8_t)res;
  }
  return 0;
}

static int eval_bool(const fgpu_plan* plan, int node, record* rec, uint8_t* bm, char* err) {
  const fgpu_expr* e = &plan->exprs[node];
  if (e->kind != FGPU_EXPR_BINARY) { snprintf(err, 512, "unsupported boolean expression"); return -1; }
  if (e->op == FGPU_OP_AND || e->op == FGPU_OP_OR) {
    if (eval_bool(plan, e->left, rec, bm, err)) return -1;
    if (e->op == FGPU_OP_AND) { /* short-circuit on empty left (filter.go:178) */
      int any = 0; for (int64_t i = 0; i < rec->n; i++) any |= bm[i];
      if (!any) return 0;
    }
    uint8_t* rb = xmalloc((size_t)rec->n);
    if (eval_bool(plan, e->right, rec, rb, err)) { free(rb); return -1; }
    if (e->op == FGPU_OP_AND) for (int64_t i = 0; i < rec->n; i++) bm[i] &= rb[i];
    else for (int64_t i = 0; i < rec->n; i++) bm[i] |= rb[i];
    free(rb);
    return 0;
  }
  return eval_leaf(plan, e, rec, bm, err);
}

/* filter(): keep selected rows of every column (filter.go:276-323; the range-slice + Concatenate there
 * is an order-preserving compaction). */
static void compact_record(record* rec, const uint8_t* bm) {
  int64_t n = rec->n, m = 0;
  for (int64_t i = 0; i < n; i++) m += bm[i];
  for (int ci = 0; ci < rec->n_cols; ci++) {
    if (!rec->cols[ci].present) continue;
    o_col* c = &rec->cols[ci].col;
    int64_t o = 0;
    for (int64_t i = 0; i < n; i++) {
      if (!bm[i]) continue;
      if (c->valid) c->valid[o] = c->valid[i];
      if (c->i64) c->i64[o] = c->i64[i];
      if (c->idx) c->idx[o] = c->idx[i];
      o++;
    }
    c->n = m;
  }
  rec->n = m;
}

/* ---- aggregate -------------------------------------------------------------------------------------- */
typedef struct { int64_t* v; int64_t n, cap; } vbuf; /* a group's buffered values (raw bits) */

typedef struct {
  uint64_t hash;
  int n_keys_cap;
  /* first-seen key values, indexed by global key-column id; -1 length = NULL / not seen */
  const uint8_t** kval; int64_t* klen; int64_t* kint; uint8_t* kset;
  vbuf* bufs;     /* per aggregate (partial stage) */
  int64_t* acc;   /* per aggregate (final stage), raw bits */
  int64_t rows;
} group;

typedef struct {
  /* key columns discovered so far (first-seen order, aggregate.go:499-504) */
  int n_keycols, cap_keycols; char** keycol_name; uint8_t* keycol_is_int; uint64_t* keycol_namehash;
  /* open-addressing map hash -> group index (map[uint64]hashtuple, aggregate.go:411) */
  uint32_t* slots; uint32_t n_slots;
  group* groups; uint32_t n_groups, cap_groups;
  int n_aggs;
} hashagg;

static void ha_init(hashagg* a, int n_aggs) { memset(a, 0, sizeof *a); a->n_aggs = n_aggs; a->n_slots = 1024; a->slots = xmalloc(a->n_slots * 4); memset(a->slots, 0xff, a->n_slots * 4); }

static int ha_keycol(hashagg* a, const char* name, int is_int) {
  for (int i = 0; i < a->n_keycols; i++) if (strcmp(a->keycol_name[i], name) == 0) return i;
  if (a->n_keycols == a->cap_keycols) {
    a->cap_keycols = a->cap_keycols ? a->cap_keycols * 2 : 8;
    a->keycol_name = xrealloc(a->keycol_name, (size_t)a->cap_keycols * sizeof(char*));
    a->keycol_is_int = xrealloc(a->keycol_is_int, (size_t)a->cap_keycols);
    a->keycol_namehash = xrealloc(a->keycol_namehash, (size_t)a->cap_keycols * 8);
  }
  a->keycol_name[a->n_keycols] = strdup(name);
  a->keycol_is_int[a->n_keycols] = (uint8_t)is_int;
  a->keycol_namehash[a->n_keycols] = hash_bytes((const uint8_t*)name, strlen(name)); /* scalar.Hash(seed, field name) */
  return a->n_keycols++;
}

static group* ha_lookup(hashagg* a, uint64_t h, int* is_new) {
  if ((a->n_groups + 1) * 2 > a->n_slots) {
    uint32_t ns = a->n_slots * 2;
    uint32_t* nsl = xmalloc(ns * 4); memset(nsl, 0xff, ns * 4);
    for (uint32_t g = 0; g < a->n_groups; g++) { uint32_t s = (uint32_t)(a->groups[g].hash * 0x9E3779B97F4A7C15ull >> 32) & (ns - 1); while (nsl[s] != 0xffffffffu) s = (s + 1) & (ns - 1); nsl[s] = g; }
    free(a->slots); a->slots = nsl; a->n_slots = ns;
  }
  uint32_t s = (uint32_t)(h * 0x9E3779B97F4A7C15ull >> 32) & (a->n_slots - 1);
  for (;;) {
    uint32_t g = a->slots[s];
    if (g == 0xffffffffu) break;
    if (a->groups[g].hash == h) { *is_new = 0; return &a->groups[g]; }
    s = (s + 1) & (a->n_slots - 1);
  }
  if (a->n_groups == a->cap_groups) { a->cap_groups = a->cap_groups ? a->cap_groups * 2 : 256; a->groups = xrealloc(a->groups, (size_t)a->cap_groups * sizeof(group)); }
  group* g = &a->groups[a->n_groups];
  memset(g, 0, sizeof *g);
  g->hash = h;
  a->slots[s] = a->n_groups++;
  *is_new = 1;
  return g;
}

static void group_set_key(group* g, int kc, int total, const uint8_t* v, int64_t len, int64_t iv, int isnull) {
  if (g->n_keys_cap < total) {
    int nc = total + 8;
    g->kval = xrealloc(g->kval, (size_t)nc * sizeof(*g->kval)); g->klen = xrealloc(g->klen, (size_t)nc * 8);
    g->kint = xrealloc(g->kint, (size_t)nc * 8); g->kset = xrealloc(g->kset, (size_t)nc);
    for (int i = g->n_keys_cap; i < nc; i++) { g->kval[i] = NULL; g->klen[i] = -1; g->kint[i] = 0; g->kset[i] = 0; }
    g->n_keys_cap = nc;
  }
  g->kset[kc] = 1; g->kval[kc] = v; g->klen[kc] = isnull ? -1 : len; g->kint[kc] = iv;
}

static inline void vbuf_push(vbuf* b, int64_t v) { if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 16; b->v = xrealloc(b->v, (size_t)b->cap * 8); } b->v[b->n++] = v; }

/* Arithmetic projections (project.go:169-395): validity ignored, Div by zero -> NULL (raw 0). */
static int expr_is_float(const fgpu_plan* plan, int node, record* rec, int* has_int_col) {
  const fgpu_expr* e = &plan->exprs[node];
  if (e->kind == FGPU_EXPR_COLUMN) { rec_col* c = rec_find(rec, e->name); if (c && c->col.type == C_F64) return 1; if (c) *has_int_col = 1; return 0; }
  if (e->kind == FGPU_EXPR_LITERAL) return e->literal.type == FGPU_SCALAR_FLOAT64;
  if (e->kind == FGPU_EXPR_BINARY) return expr_is_float(plan, e->left, rec, has_int_col) | expr_is_float(plan, e->right, rec, has_int_col);
  return 0;
}
static int eval_arith(const fgpu_plan* plan, int node, record* rec, int as_float, int64_t* out, char* err) {
  const fgpu_expr* e = &plan->exprs[node];
  int64_t n = rec->n;
  if (e->kind == FGPU_EXPR_COLUMN) {
    rec_col* c = rec_find(rec, e->name);
    if (!c) { snprintf(err, 512, "aggregate field(s) not found [\"%s\"], aggregations are not possible without it", e->name); return -1; }
    if (c->col.type == C_DICT) { snprintf(err, 512, "aggregation over non-numeric column %s", e->name); return -1; }
    memcpy(out, c->col.i64, (size_t)n * 8);
    return 0;
  }
  if (e->kind == FGPU_EXPR_LITERAL) {
    int64_t bits;
    if (as_float) { double d = e->literal.type == FGPU_SCALAR_FLOAT64 ? e->literal.f64 : (double)e->literal.i64; memcpy(&bits, &d, 8); }
    else bits = e->literal.i64;
    for (int64_t i = 0; i < n; i++) out[i] = bits;
    return 0;
  }
  if (e->kind == FGPU_EXPR_BINARY && e->op >= FGPU_OP_ADD && e->op <= FGPU_OP_DIV) {
    int64_t* r = xmalloc((size_t)n * 8);
    if (eval_arith(plan, e->left, rec, as_float, out, err) || eval_arith(plan, e->right, rec, as_float, r, err)) { free(r); return -1; }
    for (int64_t i = 0; i < n; i++) {
      if (as_float) {
        double a, b, x; memcpy(&a, &out[i], 8); memcpy(&b, &r[i], 8);
        switch (e->op) { case FGPU_OP_ADD: x = a + b; break; case FGPU_OP_SUB: x = a - b; break; case FGPU_OP_MUL: x = a * b; break; default: x = (b == 0.0) ? 0.0 : a / b; }
        memcpy(&out[i], &x, 8);
      } else {
        uint64_t a = (uint64_t)out[i], b = (uint64_t)r[i]; int64_t x;
        switch (e->op) { case FGPU_OP_ADD: x = (int64_t)(a + b); break; case FGPU_OP_SUB: x = (int64_t)(a - b); break; case FGPU_OP_MUL: x = (int64_t)(a * b); break;
          default: x = (r[i] == 0) ? 0 : (r[i] == -1 ? (int64_t)(0 - a) : out[i] / r[i]); }
        out[i] = x;
      }
    }
    free(r);
    return 0;
  }
  snprintf(err, 512, "unsupported aggregate expression");
  return -1;
}

typedef struct { const char* name; o_col* col; int kc; uint64_t* hashes; } keyref;

/* Expr.Name() of an arithmetic expression (expr.go:181,326,623): "left op right", literals printed plainly. */
static void expr_name(const fgpu_plan* plan, int node, char* out, size_t cap) {
  const fgpu_expr* e = &plan->exprs[node];
  size_t n = strlen(out);
  if (e->kind == FGPU_EXPR_COLUMN || e->kind == FGPU_EXPR_DYNCOLUMN) { snprintf(out + n, cap - n, "%s", e->name); return; }
  if (e->kind == FGPU_EXPR_LITERAL) {
    if (e->literal.type == FGPU_SCALAR_INT64) snprintf(out + n, cap - n, "%lld", (long long)e->literal.i64);
    else if (e->literal.type == FGPU_SCALAR_FLOAT64) snprintf(out + n, cap - n, "%g", e->literal.f64);
    else snprintf(out + n, cap - n, "null");
    return;
  }
  static const char* const ops[] = {"?", "==", "!=", "<", "<=", ">", ">=", "=~", "!~", "&&", "||", "+", "-", "*", "/", "contains", "not contains"};
  expr_name(plan, e->left, out, cap);
  n = strlen(out);
  snprintf(out + n, cap - n, " %s ", (e->op >= 0 && e->op <= 16) ? ops[e->op] : "?");
  expr_name(plan, e->right, out, cap);
}

/* HashAggregate.Callback (aggregate.go:263-490) for one record. */
static int ha_callback(hashagg* a, const fgpu_plan* plan, record* rec, uint8_t* agg_is_float, char* err) {
  int64_t n = rec->n;
  if (n == 0) return 0;
  /* match group-by columns by name against the record's fields, in field order (:286-304) */
  keyref* keys = xcalloc((size_t)rec->n_cols + 1, sizeof(keyref)); int nk = 0;
  for (int ci = 0; ci < rec->n_cols; ci++) {
    if (!rec->cols[ci].present) continue;
    const char* fname = rec->cols[ci].name;
    for (int g = 0; g < plan->n_group_by; g++) {
      const fgpu_expr* ge = &plan->exprs[plan->group_by[g]];
      int match = 0;
      if (ge->kind == FGPU_EXPR_COLUMN) match = strcmp(ge->name, fname) == 0;
      else if (ge->kind == FGPU_EXPR_DYNCOLUMN) { size_t pl = strlen(ge->name); match = strncmp(ge->name, fname, pl) == 0 && fname[pl] == '.'; }
      else continue; /* computed key: the pre-projection's column, appended behind the physical ones below */
      if (match) {
        int dup = 0; for (int k = 0; k < nk; k++) if (keys[k].col == &rec->cols[ci].col) dup = 1;
        if (dup) continue;
        o_col* c = &rec->cols[ci].col;
        if (c->type == C_F64) { free(keys); snprintf(err, 512, "float64 group-by unsupported"); return -1; }
        keys[nk].name = fname; keys[nk].col = c; keys[nk].kc = ha_keycol(a, fname, c->type != C_DICT);
        nk++;
      }
    }
  }
  /* computed group keys: the sqlparse pre-projection evaluates `(timestamp / 1000) * 1000 as bucket` into an int64
     column of the record (binaryExprProjection, project.go:58-167; Div by zero -> NULL, which hashes like 0);
     the aggregate then matches it by name like any other column. */
  o_col* computed = xcalloc((size_t)plan->n_group_by + 1, sizeof(o_col)); char** cnames = xcalloc((size_t)plan->n_group_by + 1, sizeof(char*)); int n_comp = 0;
  keys = xrealloc(keys, (size_t)(rec->n_cols + plan->n_group_by + 1) * sizeof(keyref));
  for (int g = 0; g < plan->n_group_by; g++) {
    const fgpu_expr* ge = &plan->exprs[plan->group_by[g]];
    if (ge->kind != FGPU_EXPR_BINARY) continue;
    int has_int = 0;
    if (expr_is_float(plan, plan->group_by[g], rec, &has_int)) { snprintf(err, 512, "float64 group-by unsupported"); for (int i = 0; i < n_comp; i++) { free(computed[i].i64); free(cnames[i]); } free(computed); free(cnames); free(keys); return -1; }
    o_col* c = &computed[n_comp];
    c->type = C_I64; c->i64 = xmalloc((size_t)n * 8); c->valid = NULL;
    if (eval_arith(plan, plan->group_by[g], rec, 0, c->i64, err)) { for (int i = 0; i <= n_comp; i++) free(computed[i].i64); for (int i = 0; i < n_comp; i++) free(cnames[i]); free(computed); free(cnames); free(keys); return -1; }
    char* nm = xcalloc(1, 512); expr_name(plan, plan->group_by[g], nm, 512);
    cnames[n_comp] = nm;
    keys[nk].name = nm; keys[nk].col = c; keys[nk].kc = ha_keycol(a, nm, 1); keys[nk].hashes = NULL;
    nk++; n_comp++;
  }
  /* per-column hashes: HashArray (hashed.go:86-105) — per ROW, as the reference does */
  for (int k = 0; k < nk; k++) {
    o_col* c = keys[k].col;
    uint64_t* h = xmalloc((size_t)n * 8);
    if (c->type == C_DICT) { for (int64_t i = 0; i < n; i++) h[i] = (c->valid && !c->valid[i]) ? 0 : hash_bytes(c->dval[c->idx[i]], c->dlen[c->idx[i]]); }
    else { for (int64_t i = 0; i < n; i++) h[i] = (c->valid && !c->valid[i]) ? 0 : (uint64_t)c->i64[i]; }  /* hashInt64Array :254-262 */
    keys[k].hashes = h;
  }
  /* aggregate inputs */
  int na = plan->n_aggs;
  int64_t** vals = xcalloc((size_t)na + 1, sizeof(int64_t*));
  int rc = 0;
  for (int j = 0; j < na && !rc; j++) {
    int has_int = 0;
    int fl = expr_is_float(plan, plan->aggs[j].expr, rec, &has_int);
    if (fl && has_int) { snprintf(err, 512, "arithmetic mixes int64 and float64 columns"); rc = -1; break; }
    agg_is_float[j] = (uint8_t)(fl && plan->aggs[j].func != FGPU_AGG_COUNT);
    vals[j] = xmalloc((size_t)n * 8);
    rc = eval_arith(plan, plan->aggs[j].expr, rec, fl, vals[j], err);
  }
  if (!rc) {
    for (int64_t i = 0; i < n; i++) {
      uint64_t hash = 0;
      for (int k = 0; k < nk; k++) {
        if (keys[k].hashes[i] == 0) continue;  /* NULL (and int64 0) contribute nothing (:400-403) */
        hash = hash_combine(hash, hash_combine(a->keycol_namehash[keys[k].kc], keys[k].hashes[i]));
      }
      int is_new;
      group* g = ha_lookup(a, hash, &is_new);
      if (is_new) {
        g->bufs = xcalloc((size_t)na + 1, sizeof(vbuf));
        for (int k = 0; k < nk; k++) {  /* updateGroupByCols (:492-525): first-seen row's values */
          o_col* c = keys[k].col;
          int isnull = c->valid && !c->valid[i];
          if (c->type == C_DICT) group_set_key(g, keys[k].kc, a->n_keycols, isnull ? NULL : c->dval[c->idx[i]], isnull ? 0 : c->dlen[c->idx[i]], 0, isnull);
          else group_set_key(g, keys[k].kc, a->n_keycols, NULL, 0, isnull ? 0 : c->i64[i], isnull);
        }
      }
      g->rows++;
      for (int j = 0; j < na; j++) vbuf_push(&g->bufs[j], vals[j][i]);  /* builder.AppendValue (:471-486) */
    }
  }
  for (int k = 0; k < nk; k++) free(keys[k].hashes);
  for (int j = 0; j < na; j++) free(vals[j]);
  for (int i = 0; i < n_comp; i++) { free(computed[i].i64); free(cnames[i]); }
  free(computed); free(cnames);
  free(vals); free(keys);
  return rc;
}

/* finishAggregate + reducers (aggregate.go:543-633, 734-950) on one partial chain */
static void ha_finish(hashagg* a, const fgpu_plan* plan, const uint8_t* agg_is_float) {
  for (uint32_t gi = 0; gi < a->n_groups; gi++) {
    group* g = &a->groups[gi];
    g->acc = xcalloc((size_t)plan->n_aggs + 1, 8);
    for (int j = 0; j < plan->n_aggs; j++) {
      vbuf* b = &g->bufs[j];
      int f = plan->aggs[j].func;
      if (f == FGPU_AGG_COUNT) { g->acc[j] = b->n; }
      else if (f == FGPU_AGG_SUM) {
        if (agg_is_float[j]) { double s = 0; for (int64_t i = 0; i < b->n; i++) { double x; memcpy(&x, &b->v[i], 8); s += x; } memcpy(&g->acc[j], &s, 8); }
        else { uint64_t s = 0; for (int64_t i = 0; i < b->n; i++) s += (uint64_t)b->v[i]; g->acc[j] = (int64_t)s; }
      } else if (f == FGPU_AGG_MIN || f == FGPU_AGG_MAX) {
        if (agg_is_float[j]) { double m; memcpy(&m, &b->v[0], 8); for (int64_t i = 0; i < b->n; i++) { double x; memcpy(&x, &b->v[i], 8); if (f == FGPU_AGG_MIN ? x < m : x > m) m = x; } memcpy(&g->acc[j], &m, 8); }
        else { int64_t m = b->v[0]; for (int64_t i = 0; i < b->n; i++) { int64_t x = b->v[i]; if (f == FGPU_AGG_MIN ? x < m : x > m) m = x; } g->acc[j] = m; }
      }
      free(b->v); b->v = NULL;
    }
    free(g->bufs); g->bufs = NULL;
  }
}

/* Final stage: merge partial groups by carried hash (aggregate.go:385-396; runAggregation :955-971:
 * Count becomes Sum of the partial counts, the others are idempotent). */
static void ha_merge(hashagg* fin, hashagg* part, const fgpu_plan* plan, const uint8_t* agg_is_float) {
  int* remap = xmalloc((size_t)(part->n_keycols + 1) * sizeof(int));
  for (int k = 0; k < part->n_keycols; k++) remap[k] = ha_keycol(fin, part->keycol_name[k], part->keycol_is_int[k]);
  for (uint32_t gi = 0; gi < part->n_groups; gi++) {
    group* pg = &part->groups[gi];
    int is_new;
    group* g = ha_lookup(fin, pg->hash, &is_new);
    if (is_new) {
      g->acc = xcalloc((size_t)plan->n_aggs + 1, 8);
      memcpy(g->acc, pg->acc, (size_t)plan->n_aggs * 8);
      g->rows = pg->rows;
      for (int k = 0; k < part->n_keycols && k < pg->n_keys_cap; k++)
        if (pg->kset[k]) group_set_key(g, remap[k], fin->n_keycols, pg->kval[k], pg->klen[k], pg->kint[k], pg->klen[k] < 0 && !part->keycol_is_int[k] ? 1 : (part->keycol_is_int[k] ? pg->klen[k] < 0 : 0));
      continue;
    }
    g->rows += pg->rows;
    for (int j = 0; j < plan->n_aggs; j++) {
      int f = plan->aggs[j].func;
      if (f == FGPU_AGG_COUNT) g->acc[j] += pg->acc[j];
      else if (f == FGPU_AGG_SUM) {
        if (agg_is_float[j]) { double x, y; memcpy(&x, &g->acc[j], 8); memcpy(&y, &pg->acc[j], 8); x += y; memcpy(&g->acc[j], &x, 8); }
        else g->acc[j] = (int64_t)((uint64_t)g->acc[j] + (uint64_t)pg->acc[j]);
      } else if (agg_is_float[j]) { double x, y; memcpy(&x, &g->acc[j], 8); memcpy(&y, &pg->acc[j], 8); if (f == FGPU_AGG_MIN ? y < x : y > x) memcpy(&g->acc[j], &pg->acc[j], 8); }
      else { if (f == FGPU_AGG_MIN ? pg->acc[j] < g->acc[j] : pg->acc[j] > g->acc[j]) g->acc[j] = pg->acc[j]; }
    }
  }
  free(remap);
}

static void ha_free(hashagg* a) {
  for (uint32_t g = 0; g < a->n_groups; g++) { group* gr = &a->groups[g]; free(gr->kval); free(gr->klen); free(gr->kint); free(gr->kset); free(gr->acc); if (gr->bufs) { for (int j = 0; j < a->n_aggs; j++) free(gr->bufs[j].v); free(gr->bufs); } }
  for (int k = 0; k < a->n_keycols; k++) free(a->keycol_name[k]);
  free(a->keycol_name); free(a->keycol_is_int); free(a->keycol_namehash); free(a->slots); free(a->groups);
  memset(a, 0, sizeof *a);
}

/* ------------------------------------------------------------------------------------------------ */
/* Execution                                                                                         */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { o_part* part; int rg; } rg_ref;

typedef struct {
  oracle_table* table; const fgpu_plan* plan;
  rg_ref* queue; int n_queue; volatile int next;
  char** proj; int n_proj; /* physical projection: exact names and "prefix." dynamic matchers */
  uint8_t* proj_dyn;
  int64_t max_rows;        /* stop handing out row groups once this many rows are taken (bounded samples) */
  volatile int64_t rows_taken;
} qctx;

typedef struct { qctx* q; hashagg agg; uint8_t agg_is_float[16]; int rc; char err[512]; int64_t rows_scanned, rows_selected; record* kept; int n_kept, cap_kept; } worker;

static int proj_wants(qctx* q, const char* name) {
  for (int i = 0; i < q->n_proj; i++) {
    if (q->proj_dyn[i]) { size_t pl = strlen(q->proj[i]); if (strncmp(q->proj[i], name, pl) == 0 && name[pl] == '.') return 1; }
    else if (strcmp(q->proj[i], name) == 0) return 1;
  }
  return 0;
}

static void collect_proj(const fgpu_plan* plan, int node, qctx* q) {
  if (node < 0) return;
  const fgpu_expr* e = &plan->exprs[node];
  if (e->kind == FGPU_EXPR_COLUMN || e->kind == FGPU_EXPR_DYNCOLUMN) {
    q->proj = xrealloc(q->proj, (size_t)(q->n_proj + 1) * sizeof(char*));
    q->proj_dyn = xrealloc(q->proj_dyn, (size_t)q->n_proj + 1);
    q->proj[q->n_proj] = (char*)e->name; q->proj_dyn[q->n_proj] = e->kind == FGPU_EXPR_DYNCOLUMN; q->n_proj++;
  } else if (e->kind == FGPU_EXPR_BINARY) { collect_proj(plan, e->left, q); collect_proj(plan, e->right, q); }
}

static void record_free(record* rec) { for (int i = 0; i < rec->n_cols; i++) if (rec->cols[i].present) col_free(&rec->cols[i].col); free(rec->cols); rec->cols = NULL; }

/* ---- row-group filter: LSM.Scan asks the plan's TrueNegativeFilter before a row group reaches the plan
   (index/lsm.go:401-454).  BooleanExpr (expr/filter.go:251-268) turns And/Or into AndExpr/OrExpr
   (:208-250), comparisons into BinaryScalarExpr, everything else into AlwaysTrueFilter (:129-131);
   BinaryScalarExpr.Eval / BinaryScalarOperation (expr/binaryscalarexpr.go:41-190) answer "may this
   column chunk hold a matching value" from the null count and the min/max of the chunk.  The
   reference reads them through the ColumnIndex of parquet-go files; here they come from the chunk
   statistics of the footer (the min/max over the pages).  Without recorded bounds the answer is
   "maybe", as for a NULL bound (:141-146). ---- */
static int stat_cmp(int phys, const uint8_t* a, uint32_t al, const fgpu_scalar* lit, int* ok) {
  /* compare(v1, v2) switches on the chunk value's kind (:296-311); only like-typed literals are decided here */
  *ok = 0;
  if (phys == PQ_INT64 && lit->type == FGPU_SCALAR_INT64 && al == 8) { int64_t v; memcpy(&v, a, 8); *ok = 1; return v < lit->i64 ? -1 : (v > lit->i64 ? 1 : 0); }
  if (phys == PQ_DOUBLE && lit->type == FGPU_SCALAR_FLOAT64 && al == 8) { double v; memcpy(&v, a, 8); *ok = 1; return v < lit->f64 ? -1 : (v > lit->f64 ? 1 : 0); }
  if (phys == PQ_BYTE_ARRAY && lit->type == FGPU_SCALAR_STRING) {
    uint64_t m = al < lit->len ? al : lit->len; int c = m ? memcmp(a, lit->bytes, (size_t)m) : 0; *ok = 1;
    return c ? (c < 0 ? -1 : 1) : (al < lit->len ? -1 : (al > lit->len ? 1 : 0));
  }
  return 0;
}

/* ---- bloom filters: what parquet-go's ColumnChunk.BloomFilter().Check(value) computes (Parquet BloomFilter.md) ---- */
static inline uint64_t xx_rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
#define XXP1 11400714785074694791ull
#define XXP2 14029467366897019727ull
#define XXP3 1609587929392839161ull
#define XXP4 9650029242287828579ull
#define XXP5 2870177450012600261ull
static inline uint64_t xx_round(uint64_t acc, uint64_t in) { return xx_rotl(acc + in * XXP2, 31) * XXP1; }
static inline uint64_t xx_merge(uint64_t acc, uint64_t v) { return (acc ^ xx_round(0, v)) * XXP1 + XXP4; }
static uint64_t xxh64(const uint8_t* p, size_t len) { /* seed 0 */
  const uint8_t* end = p + len; uint64_t h, k; uint32_t k4;
  if (len >= 32) {
    uint64_t v1 = XXP1 + XXP2, v2 = XXP2, v3 = 0, v4 = 0ull - XXP1;
    do {
      memcpy(&k, p, 8); v1 = xx_round(v1, k); memcpy(&k, p + 8, 8); v2 = xx_round(v2, k);
      memcpy(&k, p + 16, 8); v3 = xx_round(v3, k); memcpy(&k, p + 24, 8); v4 = xx_round(v4, k);
      p += 32;
    } while (p + 32 <= end);
    h = xx_rotl(v1, 1) + xx_rotl(v2, 7) + xx_rotl(v3, 12) + xx_rotl(v4, 18);
    h = xx_merge(h, v1); h = xx_merge(h, v2); h = xx_merge(h, v3); h = xx_merge(h, v4);
  } else h = XXP5;
  h += (uint64_t)len;
  while (p + 8 <= end) { memcpy(&k, p, 8); h ^= xx_round(0, k); h = xx_rotl(h, 27) * XXP1 + XXP4; p += 8; }
  if (p + 4 <= end) { memcpy(&k4, p, 4); h ^= (uint64_t)k4 * XXP1; h = xx_rotl(h, 23) * XXP2 + XXP3; p += 4; }
  while (p < end) { h ^= (uint64_t)*p * XXP5; h = xx_rotl(h, 11) * XXP1; p++; }
  h ^= h >> 33; h *= XXP2; h ^= h >> 29; h *= XXP3; h ^= h >> 32;
  return h;
}
static int sbbf_check(const uint8_t* bits, uint32_t bytes, uint64_t h) {
  static const uint32_t salt[8] = {0x47b6137bu, 0x44974d91u, 0x8824ad5bu, 0xa2b7289du, 0x705495c7u, 0x2df1424bu, 0x9efc4947u, 0x5c6bfb31u};
  uint64_t nb = bytes / 32; if (!nb) return 1;
  const uint8_t* b = bits + (((h >> 32) * nb) >> 32) * 32;
  uint32_t key = (uint32_t)h, w;
  for (int i = 0; i < 8; i++) { memcpy(&w, b + 4 * i, 4); if (!(w & (1u << ((key * salt[i]) >> 27)))) return 0; }
  return 1;
}
/* BloomFilterHeader {1 numBytes, 2 algorithm {1 BLOCK}, 3 hash {1 XXHASH}, 4 compression {1 UNCOMPRESSED}} + bitset */
static void resolve_bloom(const o_part* part, o_chunk* ch) {
  if (ch->bloom_off < 0 || (uint64_t)ch->bloom_off >= part->len) { ch->bloom_off = -1; return; }
  trd r = {part->file + ch->bloom_off, part->file + part->len, 0};
  int64_t nbytes = -1; int ok = 1, last = 0, id, t;
  while (t_field(&r, &last, &id, &t)) {
    if (id == 1) nbytes = t_zz(&r);
    else if ((id == 2 || id == 3 || id == 4) && t == 12) { int one = 0, l2 = 0, id2, t2; while (t_field(&r, &l2, &id2, &t2)) { one = id2 == 1; t_skip(&r, t2, 0); } ok = ok && one; }
    else t_skip(&r, t, 0);
  }
  ch->bloom_off = -1; /* resolved */
  if (r.err || !ok || nbytes < 32 || nbytes % 32 || (uint64_t)(r.p - part->file) + (uint64_t)nbytes > part->len) return;
  ch->bloom = r.p; ch->bloom_bytes = (uint32_t)nbytes;
}
/* 1: the filter may hold the literal (or cannot be asked), 0: it definitely does not */
static int bloom_check_lit(const o_chunk* ch, const fgpu_scalar* lit) {
  if (ch->phys == PQ_INT64 && lit->type == FGPU_SCALAR_INT64) { uint8_t b[8]; memcpy(b, &lit->i64, 8); return sbbf_check(ch->bloom, ch->bloom_bytes, xxh64(b, 8)); }
  if (ch->phys == PQ_DOUBLE && lit->type == FGPU_SCALAR_FLOAT64) { uint8_t b[8]; memcpy(b, &lit->f64, 8); return sbbf_check(ch->bloom, ch->bloom_bytes, xxh64(b, 8)); }
  if (ch->phys == PQ_BYTE_ARRAY && lit->type == FGPU_SCALAR_STRING) return sbbf_check(ch->bloom, ch->bloom_bytes, xxh64((const uint8_t*)lit->bytes, (size_t)lit->len));
  return 1;
}

/* BinaryScalarOperation (expr/binaryscalarexpr.go:84-190) over one chunk's statistics.  nulls < 0: unknown. */
static int chunk_may_match(const o_chunk* ch, int64_t nulls, int64_t num_rows, int op, const fgpu_scalar* lit) {
  int full_of_nulls = nulls >= 0 && nulls == num_rows;
  int ok;
  if (op == FGPU_OP_EQ) {
    if (lit->type == FGPU_SCALAR_NULL) return nulls != 0; /* unknown count: maybe */
    if (full_of_nulls) return 0;
    if (ch->bloom) return bloom_check_lit(ch, lit); /* :104-118: with a bloom filter the bounds are not consulted */
    if (!ch->smin || !ch->smax) return 1;
    int cmax = stat_cmp(ch->phys, ch->smax, ch->smax_len, lit, &ok); if (!ok) return 1;
    int cmin = stat_cmp(ch->phys, ch->smin, ch->smin_len, lit, &ok); if (!ok) return 1;
    return cmax >= 0 && cmin <= 0; /* compare(right, Max) <= 0 && compare(right, Min) >= 0 */
  }
  if (lit->type == FGPU_SCALAR_NULL) return 1;
  if (full_of_nulls) return 0;
  switch (op) {
    case FGPU_OP_LT_EQ: if (!ch->smin) return 1; { int c = stat_cmp(ch->phys, ch->smin, ch->smin_len, lit, &ok); return !ok || c <= 0; }
    case FGPU_OP_LT:    if (!ch->smin) return 1; { int c = stat_cmp(ch->phys, ch->smin, ch->smin_len, lit, &ok); return !ok || c < 0; }
    case FGPU_OP_GT:    if (!ch->smax) return 1; { int c = stat_cmp(ch->phys, ch->smax, ch->smax_len, lit, &ok); return !ok || c > 0; }
    case FGPU_OP_GT_EQ: if (!ch->smax) return 1; { int c = stat_cmp(ch->phys, ch->smax, ch->smax_len, lit, &ok); return !ok || c >= 0; }
    default: return 1; /* != : delegated to the execution engine */
  }
}

static int rg_may_match(const fgpu_plan* plan, int node, const o_part* part, const o_rg* rg) {
  if (node < 0) return 1;
  const fgpu_expr* e = &plan->exprs[node];
  if (e->kind != FGPU_EXPR_BINARY) return 1;
  if (e->op == FGPU_OP_AND) return rg_may_match(plan, e->left, part, rg) && rg_may_match(plan, e->right, part, rg);
  if (e->op == FGPU_OP_OR) return rg_may_match(plan, e->left, part, rg) || rg_may_match(plan, e->right, part, rg);
  if (e->op < FGPU_OP_EQ || e->op > FGPU_OP_GT_EQ) return 1; /* AlwaysTrueFilter */
  const fgpu_expr* l = &plan->exprs[e->left];
  const fgpu_expr* rt = &plan->exprs[e->right];
  if (l->kind != FGPU_EXPR_COLUMN || rt->kind != FGPU_EXPR_LITERAL) return 1;
  const fgpu_scalar* lit = &rt->literal;
  int ci = -1;
  for (int c = 0; c < part->n_leaves; c++) if (strcmp(part->leaves[c].name, l->name) == 0) { ci = c; break; }
  if (ci < 0) { /* column not in this row group (:47-73) */
    if (lit->type == FGPU_SCALAR_NULL) { if (e->op == FGPU_OP_EQ) return 1; if (e->op == FGPU_OP_NOT_EQ) return 0; }
    if (lit->type == FGPU_SCALAR_STRING) { if (e->op == FGPU_OP_EQ && lit->len == 0) return 1; if (e->op == FGPU_OP_NOT_EQ && lit->len != 0) return 1; }
    return 0;
  }
  o_chunk* ch = &rg->chunks[ci];
  if (ch->bloom_off >= 0) resolve_bloom(part, ch); /* (idempotent; the result only depends on the file) */
  int64_t nulls = ch->null_count >= 0 ? ch->null_count : (part->leaves[ci].optional ? -1 : 0);
  return chunk_may_match(ch, nulls, rg->num_rows, e->op, lit);
}

/* Test hook: the row-group filter's answer for "column == int64 literal" on one row group of a Parquet file. */
int oracle_parquet_rowgroup_may_match_eq_i64(const uint8_t* file, uint64_t len, int row_group, const char* column, int lit_is_null, int64_t lit);

/* Test hook: BinaryScalarOperation on an int64 chunk described by its statistics alone, so that the cases of
   the reference's TestBinaryScalarOperation (expr/binaryscalarexpr_test.go:56-212) can be replayed. */
int oracle_chunk_may_match_i64(int has_minmax, int64_t mn, int64_t mx, int64_t null_count, int64_t num_values, int op, int lit_is_null, int64_t lit) {
  o_chunk ch; memset(&ch, 0, sizeof ch);
  ch.phys = PQ_INT64; ch.null_count = null_count;
  if (has_minmax) { ch.smin = (const uint8_t*)&mn; ch.smin_len = 8; ch.smax = (const uint8_t*)&mx; ch.smax_len = 8; }
  fgpu_scalar s; memset(&s, 0, sizeof s);
  s.type = lit_is_null ? FGPU_SCALAR_NULL : FGPU_SCALAR_INT64; s.i64 = lit;
  return chunk_may_match(&ch, null_count, num_values, op, &s);
}

static void* worker_main(void* arg) {
  worker* w = arg; qctx* q = w->q;
  for (;;) {
    int i = __sync_fetch_and_add(&q->next, 1);
    if (i >= q->n_queue) break;
    o_part* part = q->queue[i].part; o_rg* rg = &part->rgs[q->queue[i].rg];
    if (q->max_rows > 0) { int64_t before = __sync_fetch_and_add(&q->rows_taken, rg->num_rows); if (before >= q->max_rows) break; }
    /* Convert: decode the physically projected columns of this row group (optimize.go:36-73) */
    record rec; rec.qidx = i; rec.n = rg->num_rows; rec.n_cols = part->n_leaves; rec.cols = xcalloc((size_t)part->n_leaves + 1, sizeof(rec_col));
    for (int c = 0; c < part->n_leaves && !w->rc; c++) {
      rec.cols[c].name = part->leaves[c].name;
      if (!proj_wants(q, part->leaves[c].name)) continue;
      if (decode_chunk(part, &part->leaves[c], &rg->chunks[c], rg->num_rows, &rec.cols[c].col, w->err)) { w->rc = -1; break; }
      rec.cols[c].present = 1;
    }
    if (w->rc) { record_free(&rec); break; }
    w->rows_scanned += rec.n;
    if (q->plan->filter >= 0) {
      uint8_t* bm = xmalloc((size_t)rec.n + 1);
      if (eval_bool(q->plan, q->plan->filter, &rec, bm, w->err)) { w->rc = -1; free(bm); record_free(&rec); break; }
      compact_record(&rec, bm);
      free(bm);
    }
    w->rows_selected += rec.n;
    if (q->plan->kind != FGPU_PLAN_FILTER && rec.n > 0 && ha_callback(&w->agg, q->plan, &rec, w->agg_is_float, w->err)) { w->rc = -1; record_free(&rec); break; }
    /* group key strings point into the record's dictionaries (which point into the file): keep the
       per-record dictionary arrays alive until the result is built */
    if (w->n_kept == w->cap_kept) { w->cap_kept = w->cap_kept ? w->cap_kept * 2 : 16; w->kept = xrealloc(w->kept, (size_t)w->cap_kept * sizeof(record)); }
    w->kept[w->n_kept++] = rec;
  }
  if (!w->rc) ha_finish(&w->agg, q->plan, w->agg_is_float);
  return NULL;
}

/* ---- public API ---------------------------------------------------------------------------------------- */
typedef struct oracle_result {
  int64_t n_groups; int n_keys, n_aggs;
  char** key_names; uint8_t* key_is_int;
  const uint8_t** key_str; int64_t* key_len; /* [n_keys][n_groups]; len -1 = NULL */
  int64_t* key_int;
  int64_t* aggs;                             /* [n_aggs][n_groups] raw bits */
  uint8_t* agg_is_float;
  int64_t rows_scanned, rows_selected;
  /* owned storage */
  worker* workers; int n_workers; hashagg fin;
} oracle_result;

oracle_table* oracle_table_new(void) { return xcalloc(1, sizeof(oracle_table)); }

/* The buffer is borrowed: the caller keeps it alive until oracle_table_free. */
int oracle_table_add_parquet(oracle_table* t, const uint8_t* file, uint64_t len, uint64_t tx) {
  o_part* p = xcalloc(1, sizeof(o_part));
  p->file = file; p->len = len; p->tx = tx;
  if (parse_footer(p, t->err)) { free(p); return -1; }
  if (t->n_parts == t->cap) { t->cap = t->cap ? t->cap * 2 : 8; t->parts = xrealloc(t->parts, (size_t)t->cap * sizeof(o_part*)); }
  t->parts[t->n_parts++] = p;
  return 0;
}
const char* oracle_table_error(oracle_table* t) { return t->err; }

void oracle_table_free(oracle_table* t);
/* Test hook (declared above): -1 on a malformed file / unknown row group. */
int oracle_parquet_rowgroup_may_match_eq_i64(const uint8_t* file, uint64_t len, int row_group, const char* column, int lit_is_null, int64_t lit) {
  oracle_table* t = oracle_table_new();
  if (!t) return -1;
  int rc = -1;
  if (oracle_table_add_parquet(t, file, len, 1) == 0 && row_group >= 0 && row_group < t->parts[0]->n_rgs) {
    fgpu_expr ex[3]; memset(ex, 0, sizeof ex);
    ex[0].kind = FGPU_EXPR_COLUMN; ex[0].name = column; ex[0].left = ex[0].right = -1;
    ex[1].kind = FGPU_EXPR_LITERAL; ex[1].left = ex[1].right = -1;
    ex[1].literal.type = lit_is_null ? FGPU_SCALAR_NULL : FGPU_SCALAR_INT64; ex[1].literal.i64 = lit;
    ex[2].kind = FGPU_EXPR_BINARY; ex[2].op = FGPU_OP_EQ; ex[2].left = 0; ex[2].right = 1;
    fgpu_plan plan; memset(&plan, 0, sizeof plan);
    plan.n_exprs = 3; plan.exprs = ex; plan.filter = 2;
    rc = rg_may_match(&plan, 2, t->parts[0], &t->parts[0]->rgs[row_group]);
  }
  oracle_table_free(t);
  return rc;
}
void oracle_table_free(oracle_table* t) {
  if (!t) return;
  for (int i = 0; i < t->n_parts; i++) { o_part* p = t->parts[i]; for (int l = 0; l < p->n_leaves; l++) free(p->leaves[l].name); free(p->leaves); for (int g = 0; g < p->n_rgs; g++) free(p->rgs[g].chunks); free(p->rgs); free(p); }
  free(t->parts); free(t);
}

void oracle_result_free(oracle_result* r) {
  if (!r) return;
  for (int i = 0; i < r->n_workers; i++) { worker* w = &r->workers[i]; for (int k = 0; k < w->n_kept; k++) record_free(&w->kept[k]); free(w->kept); ha_free(&w->agg); }
  free(r->workers); ha_free(&r->fin);
  for (int k = 0; k < r->n_keys; k++) free(r->key_names[k]);
  free(r->key_names); free(r->key_is_int); free(r->key_str); free(r->key_len); free(r->key_int); free(r->aggs); free(r->agg_is_float);
  free(r);
}

/* Executes the plan over every part with tx <= watermark using n_threads worker chains.
 * max_rows > 0 bounds the scan to (about) that many rows (cpu_baseline samples). */
int oracle_execute(oracle_table* t, const fgpu_plan* plan, uint64_t tx_watermark, int n_threads, int64_t max_rows, oracle_result** out) {
  if (plan->n_aggs > 15) { snprintf(t->err, 512, "too many aggregates"); return -1; }
  qctx q; memset(&q, 0, sizeof q);
  q.table = t; q.plan = plan; q.max_rows = max_rows;
  for (int p = 0; p < t->n_parts; p++) if (t->parts[p]->tx <= tx_watermark) q.n_queue += t->parts[p]->n_rgs;
  q.queue = xcalloc((size_t)q.n_queue + 1, sizeof(rg_ref));
  const int no_prune = getenv("FROST_ORACLE_NO_PRUNE") != NULL; /* tests: the filter never changes a result */
  int qi = 0;
  int64_t pruned_rows = 0; /* row groups the filter rules out still count as scanned rows of the table */
  for (int p = 0; p < t->n_parts; p++) if (t->parts[p]->tx <= tx_watermark) for (int g = 0; g < t->parts[p]->n_rgs; g++) if (t->parts[p]->rgs[g].num_rows > 0) {
    if (!no_prune && !rg_may_match(plan, plan->filter, t->parts[p], &t->parts[p]->rgs[g])) { pruned_rows += t->parts[p]->rgs[g].num_rows; continue; }
    q.queue[qi].part = t->parts[p]; q.queue[qi].rg = g; qi++;
  }
  q.n_queue = qi;
  collect_proj(plan, plan->filter, &q);
  for (int g = 0; g < plan->n_group_by; g++) collect_proj(plan, plan->group_by[g], &q);
  for (int a = 0; a < plan->n_aggs; a++) collect_proj(plan, plan->aggs[a].expr, &q);
  if (n_threads < 1) n_threads = 1;
  oracle_result* r = xcalloc(1, sizeof *r);
  r->workers = xcalloc((size_t)n_threads, sizeof(worker)); r->n_workers = n_threads;
  pthread_t* th = xcalloc((size_t)n_threads, sizeof(pthread_t));
  for (int i = 0; i < n_threads; i++) { r->workers[i].q = &q; ha_init(&r->workers[i].agg, plan->n_aggs); }
  if (n_threads == 1) worker_main(&r->workers[0]);
  else { for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, worker_main, &r->workers[i]); for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL); }
  free(th);
  int rc = 0;
  for (int i = 0; i < n_threads; i++) if (r->workers[i].rc) { rc = -1; snprintf(t->err, 512, "%s", r->workers[i].err); }
  /* Synchronizer + final HashAggregate */
  ha_init(&r->fin, plan->n_aggs);
  uint8_t aif[16] = {0};
  for (int i = 0; i < n_threads; i++) for (int j = 0; j < plan->n_aggs; j++) aif[j] |= r->workers[i].agg_is_float[j];
  if (!rc) for (int i = 0; i < n_threads; i++) { ha_merge(&r->fin, &r->workers[i].agg, plan, aif); r->rows_scanned += r->workers[i].rows_scanned; r->rows_selected += r->workers[i].rows_selected; }
  free(q.queue); free(q.proj); free(q.proj_dyn);
  if (rc) { oracle_result_free(r); return -1; }
  r->rows_scanned += pruned_rows;
  if (plan->kind == FGPU_PLAN_FILTER) {
    /* Filter -> Projection(columns): the compacted rows themselves, in scan order. */
    int total_recs = 0; for (int i = 0; i < n_threads; i++) total_recs += r->workers[i].n_kept;
    record** recs = xcalloc((size_t)total_recs + 1, sizeof(record*)); int nr = 0;
    for (int i = 0; i < n_threads; i++) for (int k = 0; k < r->workers[i].n_kept; k++) recs[nr++] = &r->workers[i].kept[k];
    for (int a = 1; a < nr; a++) { record* x = recs[a]; int b = a - 1; while (b >= 0 && recs[b]->qidx > x->qidx) { recs[b + 1] = recs[b]; b--; } recs[b + 1] = x; }
    hashagg* f = &r->fin; int64_t rows = 0;
    for (int a = 0; a < nr; a++) {
      if (recs[a]->n == 0) continue;
      rows += recs[a]->n;
      for (int ci = 0; ci < recs[a]->n_cols; ci++) {
        if (!recs[a]->cols[ci].present) continue;
        const char* fname = recs[a]->cols[ci].name;
        for (int g = 0; g < plan->n_group_by; g++) {
          const fgpu_expr* ge = &plan->exprs[plan->group_by[g]];
          int match = 0;
          if (ge->kind == FGPU_EXPR_COLUMN) match = strcmp(ge->name, fname) == 0;
          else if (ge->kind == FGPU_EXPR_DYNCOLUMN) { size_t pl = strlen(ge->name); match = strncmp(ge->name, fname, pl) == 0 && fname[pl] == '.'; }
          if (match) { int t = recs[a]->cols[ci].col.type; ha_keycol(f, fname, t == C_DICT ? 0 : (t == C_F64 ? 2 : 1)); }
        }
      }
    }
    r->n_groups = rows; r->n_keys = f->n_keycols; r->n_aggs = 0;
    r->key_names = xcalloc((size_t)r->n_keys + 1, sizeof(char*)); r->key_is_int = xcalloc((size_t)r->n_keys + 1, 1);
    size_t kg = (size_t)r->n_keys * (size_t)rows;
    r->key_str = xcalloc(kg + 1, sizeof(*r->key_str)); r->key_len = xcalloc(kg + 1, 8); r->key_int = xcalloc(kg + 1, 8);
    r->aggs = xcalloc(1, 8); r->agg_is_float = xcalloc(1, 1);
    for (int k = 0; k < r->n_keys; k++) { r->key_names[k] = strdup(f->keycol_name[k]); r->key_is_int[k] = f->keycol_is_int[k]; }
    int64_t base = 0;
    for (int a = 0; a < nr; a++) {
      record* rec = recs[a];
      for (int k = 0; k < r->n_keys; k++) {
        rec_col* c = rec_find(rec, r->key_names[k]);
        for (int64_t i = 0; i < rec->n; i++) {
          size_t o = (size_t)k * (size_t)rows + (size_t)(base + i);
          if (!c || (c->col.valid && !c->col.valid[i])) { r->key_len[o] = -1; continue; }
          if (c->col.type == C_DICT) { r->key_str[o] = c->col.dval[c->col.idx[i]]; r->key_len[o] = c->col.dlen[c->col.idx[i]]; }
          else { r->key_int[o] = c->col.i64[i]; r->key_len[o] = 0; }
        }
      }
      base += rec->n;
    }
    free(recs);
    *out = r;
    return 0;
  }
  hashagg* f = &r->fin;
  r->n_groups = f->n_groups; r->n_keys = f->n_keycols; r->n_aggs = plan->n_aggs;
  r->key_names = xcalloc((size_t)r->n_keys + 1, sizeof(char*)); r->key_is_int = xcalloc((size_t)r->n_keys + 1, 1);
  size_t kg = (size_t)r->n_keys * (size_t)r->n_groups;
  r->key_str = xcalloc(kg + 1, sizeof(*r->key_str)); r->key_len = xcalloc(kg + 1, 8); r->key_int = xcalloc(kg + 1, 8);
  r->aggs = xcalloc((size_t)r->n_aggs * (size_t)r->n_groups + 1, 8); r->agg_is_float = xcalloc((size_t)r->n_aggs + 1, 1);
  memcpy(r->agg_is_float, aif, (size_t)r->n_aggs);
  for (int k = 0; k < r->n_keys; k++) { r->key_names[k] = strdup(f->keycol_name[k]); r->key_is_int[k] = f->keycol_is_int[k]; }
  for (int64_t g = 0; g < r->n_groups; g++) {
    group* gr = &f->groups[g];
    for (int k = 0; k < r->n_keys; k++) {
      size_t o = (size_t)k * (size_t)r->n_groups + (size_t)g;
      if (k < gr->n_keys_cap && gr->kset[k]) { r->key_str[o] = gr->kval[k]; r->key_len[o] = gr->klen[k]; r->key_int[o] = gr->kint[k]; }
      else { r->key_str[o] = NULL; r->key_len[o] = -1; } /* NULL back-fill (aggregate.go:568-575) */
    }
    for (int j = 0; j < r->n_aggs; j++) r->aggs[(size_t)j * (size_t)r->n_groups + (size_t)g] = gr->acc[j];
  }
  *out = r;
  return 0;
}

/* flat accessors for ctypes */
int64_t oracle_result_groups(oracle_result* r) { return r->n_groups; }
int oracle_result_n_keys(oracle_result* r) { return r->n_keys; }
int oracle_result_n_aggs(oracle_result* r) { return r->n_aggs; }
const char* oracle_result_key_name(oracle_result* r, int k) { return r->key_names[k]; }
int oracle_result_key_is_int(oracle_result* r, int k) { return r->key_is_int[k]; }
int oracle_result_agg_is_float(oracle_result* r, int j) { return r->agg_is_float[j]; }
const uint8_t* const* oracle_result_key_str(oracle_result* r, int k) { return r->key_str + (size_t)k * (size_t)r->n_groups; }
const int64_t* oracle_result_key_len(oracle_result* r, int k) { return r->key_len + (size_t)k * (size_t)r->n_groups; }
const int64_t* oracle_result_key_int(oracle_result* r, int k) { return r->key_int + (size_t)k * (size_t)r->n_groups; }
const int64_t* oracle_result_agg(oracle_result* r, int j) { return r->aggs + (size_t)j * (size_t)r->n_groups; }
int64_t oracle_result_rows_scanned(oracle_result* r) { return r->rows_scanned; }
int64_t oracle_result_rows_selected(oracle_result* r) { return r->rows_selected; }

/* Decode-only entry for the K1 parity tests: column `name` of row group `rg` of part `p`. */
int oracle_decode_column(oracle_table* t, int p, int rg, const char* name, int64_t* n_rows, int* type, uint8_t** valid, int64_t** i64,
                         uint32_t** idx, uint32_t* n_dict, const uint8_t*** dval, uint32_t** dlen) {
  if (p < 0 || p >= t->n_parts || rg < 0 || rg >= t->parts[p]->n_rgs) { snprintf(t->err, 512, "bad part/row group"); return -1; }
  o_part* part = t->parts[p];
  for (int c = 0; c < part->n_leaves; c++) {
    if (strcmp(part->leaves[c].name, name)) continue;
    o_col col;
    if (decode_chunk(part, &part->leaves[c], &part->rgs[rg].chunks[c], part->rgs[rg].num_rows, &col, t->err)) return -1;
    *n_rows = col.n; *type = col.type; *valid = col.valid; *i64 = col.i64; *idx = col.idx; *n_dict = col.n_dict; *dval = col.dval; *dlen = col.dlen;
    free(col.memo);
    return 0;
  }
  snprintf(t->err, 512, "column not found: %s", name);
  return -2;
}
void oracle_free(void* p) { free(p); }
