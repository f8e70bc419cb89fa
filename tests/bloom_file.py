"""A minimal Parquet writer for the bloom-filter tests: ONE row group, required INT64 columns, PLAIN data page v1, and a
split-block bloom filter per column (BloomFilterHeader + bitset, ColumnMetaData.bloom_filter_offset / _length) — the layout
parquet-go gives FrostDB's sorting columns (dynparquet/schema.go:1111-1157).  pyarrow 24 cannot write bloom filters from
Python, so the footer is assembled here with a small Thrift compact-protocol encoder; pyarrow reads the files back.
Also: XXH64 and the split-block filter in plain Python (Parquet format, BloomFilter.md), independent of the library."""
from __future__ import annotations

import struct
from typing import Dict, List

MASK64 = (1 << 64) - 1
P1, P2, P3, P4, P5 = 11400714785074694791, 14029467366897019727, 1609587929392839161, 9650029242287828579, 2870177450012600261
SALT = [0x47b6137b, 0x44974d91, 0x8824ad5b, 0xa2b7289d, 0x705495c7, 0x2df1424b, 0x9efc4947, 0x5c6bfb31]


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & MASK64


def _round(acc, inp):
    return (_rotl((acc + inp * P2) & MASK64, 31) * P1) & MASK64


def _merge(acc, v):
    return ((acc ^ _round(0, v)) * P1 + P4) & MASK64


def xxh64(data: bytes, seed: int = 0) -> int:
    n, p = len(data), 0
    if n >= 32:
        v1, v2, v3, v4 = (seed + P1 + P2) & MASK64, (seed + P2) & MASK64, seed, (seed - P1) & MASK64
        while p + 32 <= n:
            a, b, c, d = struct.unpack_from("<4Q", data, p)
            v1, v2, v3, v4 = _round(v1, a), _round(v2, b), _round(v3, c), _round(v4, d)
            p += 32
        h = (_rotl(v1, 1) + _rotl(v2, 7) + _rotl(v3, 12) + _rotl(v4, 18)) & MASK64
        for v in (v1, v2, v3, v4):
            h = _merge(h, v)
    else:
        h = (seed + P5) & MASK64
    h = (h + n) & MASK64
    while p + 8 <= n:
        h ^= _round(0, struct.unpack_from("<Q", data, p)[0])
        h = (_rotl(h, 27) * P1 + P4) & MASK64
        p += 8
    if p + 4 <= n:
        h ^= (struct.unpack_from("<I", data, p)[0] * P1) & MASK64
        h = (_rotl(h, 23) * P2 + P3) & MASK64
        p += 4
    while p < n:
        h ^= (data[p] * P5) & MASK64
        h = (_rotl(h, 11) * P1) & MASK64
        p += 1
    h ^= h >> 33
    h = (h * P2) & MASK64
    h ^= h >> 29
    h = (h * P3) & MASK64
    h ^= h >> 32
    return h


def sbbf_mask(hash64: int, n_bytes: int):
    """(block index, the 8 mask words) of a hash in a filter of n_bytes."""
    block = ((hash64 >> 32) * (n_bytes // 32)) >> 32
    key = hash64 & 0xffffffff
    return block, [1 << (((key * s) & 0xffffffff) >> 27) for s in SALT]


def sbbf_insert(bitset: bytearray, hash64: int) -> None:
    block, mask = sbbf_mask(hash64, len(bitset))
    for i, m in enumerate(mask):
        off = block * 32 + 4 * i
        struct.pack_into("<I", bitset, off, struct.unpack_from("<I", bitset, off)[0] | m)


def sbbf_check(bitset: bytes, hash64: int) -> bool:
    block, mask = sbbf_mask(hash64, len(bitset))
    return all(struct.unpack_from("<I", bitset, block * 32 + 4 * i)[0] & m for i, m in enumerate(mask))


# ---- Thrift compact protocol (just what the footer needs) ---------------------------------------------------------
def _uvar(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7f
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _zz(n: int) -> bytes:
    return _uvar((n << 1) ^ (n >> 63))


class Struct:
    """Fields in ascending id order; types: 5 i32, 6 i64, 8 binary, 9 list, 12 struct."""

    def __init__(self):
        self.buf, self.last = bytearray(), 0

    def _hdr(self, fid: int, t: int):
        d = fid - self.last
        if 0 < d <= 15:
            self.buf.append((d << 4) | t)
        else:
            self.buf.append(t)
            self.buf += _zz(fid)
        self.last = fid

    def i32(self, fid, v):
        self._hdr(fid, 5); self.buf += _zz(v); return self

    def i64(self, fid, v):
        self._hdr(fid, 6); self.buf += _zz(v); return self

    def binary(self, fid, v: bytes):
        self._hdr(fid, 8); self.buf += _uvar(len(v)) + v; return self

    def struct(self, fid, s: "Struct"):
        self._hdr(fid, 12); self.buf += s.done(); return self

    def list(self, fid, etype: int, items: List[bytes]):
        self._hdr(fid, 9)
        n = len(items)
        self.buf += bytes([(n << 4) | etype]) if n < 15 else bytes([0xf0 | etype]) + _uvar(n)
        for it in items:
            self.buf += it
        return self

    def done(self) -> bytes:
        return bytes(self.buf) + b"\x00"


def write_int64_file(columns: Dict[str, List[int]], bits_per_value: int = 10) -> bytes:
    """One row group, every column a required INT64 with a bloom filter over its values."""
    n_rows = len(next(iter(columns.values())))
    out = bytearray(b"PAR1")
    chunks = []
    for name, vals in columns.items():
        assert len(vals) == n_rows
        payload = struct.pack(f"<{n_rows}q", *vals)
        dph = Struct().i32(1, n_rows).i32(2, 0).i32(3, 3).i32(4, 3)                  # num_values, PLAIN, RLE, RLE
        ph = Struct().i32(1, 0).i32(2, len(payload)).i32(3, len(payload)).struct(5, dph).done()  # DATA_PAGE
        page_off = len(out)
        out += ph + payload
        n_bytes = max(32, ((n_rows * bits_per_value // 8) + 31) // 32 * 32)
        bitset = bytearray(n_bytes)
        for v in vals:
            sbbf_insert(bitset, xxh64(struct.pack("<q", v)))
        empty = Struct()
        one = lambda: Struct().struct(1, Struct())                                    # union with field 1 = {}
        bh = Struct().i32(1, n_bytes).struct(2, one()).struct(3, one()).struct(4, one()).done()
        bloom_off = len(out)
        out += bh + bitset
        stats = Struct().i64(3, 0).binary(5, struct.pack("<q", max(vals))).binary(6, struct.pack("<q", min(vals)))
        md = (Struct().i32(1, 2).list(2, 5, [_zz(0), _zz(3)]).list(3, 8, [_uvar(len(name)) + name.encode()]).i32(4, 0)
              .i64(5, n_rows).i64(6, len(ph) + len(payload)).i64(7, len(ph) + len(payload)).i64(9, page_off)
              .struct(12, stats).i64(14, bloom_off).i32(15, len(bh) + n_bytes))
        chunks.append((Struct().i64(2, page_off).struct(3, md).done(), len(ph) + len(payload), bytes(bitset)))
    schema = [Struct().binary(4, b"schema").i32(5, len(columns)).done()]
    for name in columns:
        schema.append(Struct().i32(1, 2).i32(3, 0).binary(4, name.encode()).done())   # INT64, REQUIRED
    rg = Struct().list(1, 12, [c[0] for c in chunks]).i64(2, sum(c[1] for c in chunks)).i64(3, n_rows).done()
    footer = Struct().i32(1, 1).list(2, 12, schema).i64(3, n_rows).list(4, 12, [rg]).binary(6, b"frostgpu tests (bloom_file.py)").done()
    out += footer + struct.pack("<I", len(footer)) + b"PAR1"
    return bytes(out)
