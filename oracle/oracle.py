"""ctypes wrapper of the C oracle (oracle/frost_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
frostdb_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libfrost_oracle.so")
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE])
    lib = C.CDLL(LIB_PATH)
    lib.oracle_table_new.restype = C.c_void_p
    lib.oracle_table_add_parquet.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
    lib.oracle_table_error.argtypes = [C.c_void_p]
    lib.oracle_table_error.restype = C.c_char_p
    lib.oracle_table_free.argtypes = [C.c_void_p]
    lib.oracle_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
    lib.oracle_result_free.argtypes = [C.c_void_p]
    for name, res in [("groups", C.c_int64), ("n_keys", C.c_int), ("n_aggs", C.c_int), ("rows_scanned", C.c_int64),
                      ("rows_selected", C.c_int64)]:
        f = getattr(lib, "oracle_result_" + name)
        f.argtypes = [C.c_void_p]
        f.restype = res
    for name, res in [("key_name", C.c_char_p), ("key_is_int", C.c_int), ("agg_is_float", C.c_int),
                      ("key_str", C.POINTER(C.c_void_p)), ("key_len", C.POINTER(C.c_int64)),
                      ("key_int", C.POINTER(C.c_int64)), ("agg", C.POINTER(C.c_int64))]:
        f = getattr(lib, "oracle_result_" + name)
        f.argtypes = [C.c_void_p, C.c_int]
        f.restype = res
    lib.oracle_decode_column.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.oracle_free.argtypes = [C.c_void_p]
    lib.oracle_chunk_may_match_i64.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64]
    lib.oracle_chunk_may_match_i64.restype = C.c_int
    lib.oracle_parquet_rowgroup_may_match_eq_i64.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_char_p, C.c_int, C.c_int64]
    lib.oracle_parquet_rowgroup_may_match_eq_i64.restype = C.c_int
    _lib = lib
    return lib


class OracleError(RuntimeError):
    pass


class OracleTable:
    """The parts of one table, as Parquet buffers (kept alive here; the C side borrows them)."""

    def __init__(self):
        self.lib = load()
        self.h = C.c_void_p(self.lib.oracle_table_new())
        self._bufs = []
        self.watermark = 0

    def add_parquet(self, buf: bytes, tx: Optional[int] = None) -> None:
        if tx is None:
            tx = self.watermark + 1
        keep = (C.c_char * len(buf)).from_buffer_copy(buf)
        self._bufs.append(keep)
        if self.lib.oracle_table_add_parquet(self.h, C.addressof(keep), len(buf), tx) != 0:
            raise OracleError(self.lib.oracle_table_error(self.h).decode())
        self.watermark = max(self.watermark, tx)

    def add_pinned(self, address: int, length: int, keep, tx: Optional[int] = None) -> None:
        """Borrow an existing buffer (e.g. a numpy array) without copying."""
        if tx is None:
            tx = self.watermark + 1
        self._bufs.append(keep)
        if self.lib.oracle_table_add_parquet(self.h, address, length, tx) != 0:
            raise OracleError(self.lib.oracle_table_error(self.h).decode())
        self.watermark = max(self.watermark, tx)

    def close(self):
        if self.h:
            self.lib.oracle_table_free(self.h)
            self.h = None
            self._bufs = []

    def execute(self, plan_struct, *, tx: Optional[int] = None, threads: int = 1, max_rows: int = 0) -> "OracleResult":
        """plan_struct: a frostdb_b200._lib.Plan (the same POD plan the product receives)."""
        res = C.c_void_p()
        rc = self.lib.oracle_execute(self.h, C.byref(plan_struct), self.watermark if tx is None else tx, threads, max_rows,
                                     C.byref(res))
        if rc != 0:
            raise OracleError(self.lib.oracle_table_error(self.h).decode())
        return OracleResult(self.lib, res)

    def decode_column(self, part: int, rg: int, name: str) -> list:
        lib = self.lib
        n, typ, nd = C.c_int64(), C.c_int(), C.c_uint32()
        valid, i64, idx, dval, dlen = (C.c_void_p() for _ in range(5))
        rc = lib.oracle_decode_column(self.h, part, rg, name.encode(), C.byref(n), C.byref(typ), C.byref(valid), C.byref(i64),
                                      C.byref(idx), C.byref(nd), C.byref(dval), C.byref(dlen))
        if rc != 0:
            raise OracleError(lib.oracle_table_error(self.h).decode())
        rows = n.value
        v = np.ctypeslib.as_array(C.cast(valid, C.POINTER(C.c_uint8)), (rows,)).copy() if valid.value else np.ones(rows, np.uint8)
        out: list
        if typ.value == 3:
            ix = np.ctypeslib.as_array(C.cast(idx, C.POINTER(C.c_uint32)), (rows,)).copy() if rows else np.zeros(0, np.uint32)
            ptrs = C.cast(dval, C.POINTER(C.c_void_p))
            lens = C.cast(dlen, C.POINTER(C.c_uint32))
            d = [C.string_at(ptrs[i], lens[i]).decode() for i in range(nd.value)]
            out = [d[ix[i]] if v[i] else None for i in range(rows)]
        else:
            raw = np.ctypeslib.as_array(C.cast(i64, C.POINTER(C.c_int64)), (rows,)).copy() if rows else np.zeros(0, np.int64)
            vals = raw.view(np.float64) if typ.value == 2 else raw
            out = [vals[i].item() if v[i] else None for i in range(rows)]
        for p in (valid, i64, idx, dval, dlen):
            if p.value:
                lib.oracle_free(p)
        return out


class OracleResult:
    def __init__(self, lib, h):
        self.lib, self.h = lib, h
        self.n_groups = lib.oracle_result_groups(h)
        self.rows_scanned = lib.oracle_result_rows_scanned(h)
        self.rows_selected = lib.oracle_result_rows_selected(h)

    def to_batch(self, agg_names: List[str]) -> pa.RecordBatch:
        """Result as one record: group columns (first-seen order) then aggregates, named like
        aggregate.go:615-618 names them."""
        lib, h, G = self.lib, self.h, self.n_groups
        names, arrays = [], []
        for k in range(lib.oracle_result_n_keys(h)):
            names.append(lib.oracle_result_key_name(h, k).decode())
            lens = np.ctypeslib.as_array(lib.oracle_result_key_len(h, k), (G,)) if G else np.zeros(0, np.int64)
            kind = lib.oracle_result_key_is_int(h, k)
            if kind:
                ints = np.ctypeslib.as_array(lib.oracle_result_key_int(h, k), (G,)).copy() if G else np.zeros(0, np.int64)
                vals = ints.view(np.float64) if kind == 2 else ints
                arrays.append(pa.array(vals, mask=(lens < 0) if G and (lens < 0).any() else None))
            else:
                ptrs = lib.oracle_result_key_str(h, k)
                arrays.append(pa.array([None if lens[g] < 0 else C.string_at(ptrs[g], int(lens[g])) for g in range(G)],
                                       type=pa.binary()))
        for j in range(lib.oracle_result_n_aggs(h)):
            raw = np.ctypeslib.as_array(lib.oracle_result_agg(h, j), (G,)).copy() if G else np.zeros(0, np.int64)
            names.append(agg_names[j])
            arrays.append(pa.array(raw.view(np.float64) if lib.oracle_result_agg_is_float(h, j) else raw))
        return pa.RecordBatch.from_arrays(arrays, names=names)

    def close(self):
        if self.h:
            self.lib.oracle_result_free(self.h)
            self.h = None


def chunk_may_match_i64(has_minmax: bool, mn: int, mx: int, null_count: int, num_values: int, op: int, literal) -> bool:
    """BinaryScalarOperation of the row-group filter on an int64 chunk given by its statistics (literal None = NULL)."""
    return bool(load().oracle_chunk_may_match_i64(int(has_minmax), mn, mx, null_count, num_values, op, int(literal is None), literal or 0))


def parquet_rowgroup_may_match_eq_i64(buf: bytes, row_group: int, column: str, literal) -> bool:
    """The oracle's row-group filter for `column == literal` (None: NULL) on one row group of a Parquet file: null count, then
    the chunk's bloom filter when it has one, else its bounds (expr/binaryscalarexpr.go:84-128)."""
    rc = load().oracle_parquet_rowgroup_may_match_eq_i64(buf, len(buf), row_group, column.encode(), int(literal is None), literal or 0)
    if rc < 0:
        raise ValueError("malformed Parquet file or unknown row group")
    return bool(rc)
