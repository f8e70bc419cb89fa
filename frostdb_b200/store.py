"""GPUEngine (owner of the fgpu_ctx) and a minimal ColumnStore/DB/Table mirror.

Only what the scan path needs from `frostdb.ColumnStore / DB / Table` (db.go, table.go) is here:
creating a table with a dynparquet schema, inserting rows (each insert becomes one immutable part
with the next transaction id, as Table.InsertRecord + compaction would leave it), the read
watermark, and `TableProvider`.  WAL, snapshots, LSM compaction, object storage are out of scope
(SURVEY.md §8): in a FrostDB process the Go database keeps doing all of that and hands finished
parts to `fgpu_part_put_parquet`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterator, List, Optional

import pyarrow as pa

from . import _lib
from . import dynparquet as dp


class GPUEngine:
    """One per process per GPU: wraps fgpu_init / fgpu_shutdown."""

    def __init__(self, device: int = 0):
        lib = _lib.load()
        cfg = _lib.Config(abi_version=_lib.ABI_VERSION, device=device, tile_rows=0, flags=0, staging_bytes=0)
        h = C.c_void_p()
        _lib.check(lib.fgpu_init(C.byref(cfg), C.byref(h)))
        self.handle = h
        self.device = device
        self._watermarks: Dict[str, int] = {}
        self._next_part: Dict[str, int] = {}

    def close(self) -> None:
        if self.handle:
            _lib.load().fgpu_shutdown(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- parts -------------------------------------------------------------------------------
    def put_parquet(self, table: str, buf, tx: Optional[int] = None, part_id: Optional[int] = None, borrow: bool = False) -> int:
        lib = _lib.load()
        pid = self._next_part.get(table, 0) if part_id is None else part_id
        self._next_part[table] = max(self._next_part.get(table, 0), pid + 1)
        if tx is None:
            tx = self._watermarks.get(table, 0) + 1
        if isinstance(buf, (bytes, bytearray, memoryview)):
            src = (C.c_char * len(buf)).from_buffer_copy(buf)
            addr, n = C.addressof(src), len(buf)
        else:  # numpy uint8 array: no copy
            addr, n = buf.ctypes.data, buf.nbytes
        if borrow and isinstance(buf, (bytes, bytearray, memoryview)):
            raise ValueError("borrowed parts must live in caller-owned (pinned) memory: pass a numpy array")
        _lib.check(lib.fgpu_part_put_parquet(self.handle, table.encode(), pid, tx, addr, n,
                                             _lib.PUT_BORROW_PINNED if borrow else _lib.PUT_DEFAULT))
        self._watermarks[table] = max(self._watermarks.get(table, 0), tx)
        return pid

    def put_arrow(self, table: str, record: pa.RecordBatch, tx: Optional[int] = None, part_id: Optional[int] = None) -> int:
        """An L0 part: the Arrow record itself (parts/arrow.go:14-55), handed over through the C Data Interface."""
        lib = _lib.load()
        pid = self._next_part.get(table, 0) if part_id is None else part_id
        self._next_part[table] = max(self._next_part.get(table, 0), pid + 1)
        if tx is None:
            tx = self._watermarks.get(table, 0) + 1
        schema, array = _lib.ArrowSchema(), _lib.ArrowArray()
        record._export_to_c(C.addressof(array), C.addressof(schema))
        _lib.check(lib.fgpu_part_put_arrow(self.handle, table.encode(), pid, tx, C.addressof(schema), C.addressof(array)))
        self._watermarks[table] = max(self._watermarks.get(table, 0), tx)
        return pid

    def drop_part(self, table: str, part_id: int) -> None:
        _lib.check(_lib.load().fgpu_part_drop(self.handle, table.encode(), part_id))

    def drop_table(self, table: str) -> None:
        _lib.check(_lib.load().fgpu_table_drop(self.handle, table.encode()))
        self._watermarks.pop(table, None)
        self._next_part.pop(table, None)

    def table_watermark(self, table: str) -> int:
        return self._watermarks.get(table, 0)

    # ---- results -------------------------------------------------------------------------------
    def drain(self, res: C.c_void_p) -> Iterator[pa.RecordBatch]:
        lib = _lib.load()
        while True:
            schema, array = _lib.ArrowSchema(), _lib.ArrowArray()
            rc = lib.fgpu_result_next(res, C.addressof(schema), C.addressof(array))
            if rc == _lib.FGPU_ERR_END:
                return
            _lib.check(rc)
            yield pa.RecordBatch._import_from_c(C.addressof(array), C.addressof(schema))

    def stats(self, res: C.c_void_p) -> dict:
        st = _lib.Stats()
        _lib.check(_lib.load().fgpu_result_stats(res, C.byref(st)))
        return st.asdict()

    def decode_column(self, table: str, part_id: int, column: str) -> pa.Array:
        """K1 standalone: the column of one part decoded on the GPU (fgpu_part_decode_column)."""
        schema, array = _lib.ArrowSchema(), _lib.ArrowArray()
        _lib.check(_lib.load().fgpu_part_decode_column(self.handle, table.encode(), part_id, column.encode(),
                                                       C.addressof(schema), C.addressof(array)))
        return pa.Array._import_from_c(C.addressof(array), C.addressof(schema))

    # ---- multi-GPU exchange inside the library (fgpu_comm_*) -------------------------------------------
    def comm_open(self, rank: int, n_ranks: int, exchange, slot_bytes: int = 64 << 20) -> None:
        """One process per GPU: exports this rank's mailbox, lets `exchange(handle_bytes) -> [handle_bytes] * n`
        move the handles (any transport), opens the peers."""
        lib = _lib.load()
        h = (C.c_uint8 * _lib.COMM_HANDLE_BYTES)()
        _lib.check(lib.fgpu_comm_export(self.handle, rank, n_ranks, slot_bytes, h))
        handles = exchange(bytes(h))
        assert len(handles) == n_ranks
        _lib.check(lib.fgpu_comm_open(self.handle, b"".join(handles)))

    def comm_close(self) -> None:
        _lib.check(_lib.load().fgpu_comm_close(self.handle))

    def execute_collective(self, query, tx: int) -> C.c_void_p:
        res = C.c_void_p()
        _lib.check(_lib.load().fgpu_query_execute_collective(self.handle, query, tx, C.byref(res)))
        return res

    def execute_collective_begin(self, query, tx: int) -> C.c_void_p:
        res = C.c_void_p()
        _lib.check(_lib.load().fgpu_query_execute_collective_begin(self.handle, query, tx, C.byref(res)))
        return res

    def execute_collective_end(self, res: C.c_void_p) -> None:
        _lib.check(_lib.load().fgpu_query_execute_collective_end(self.handle, res))

    # ---- cross-rank dictionaries ------------------------------------------------------------------
    def dict_export(self, table: str, column: str) -> List[bytes]:
        lib = _lib.load()
        n, cnt = C.c_uint64(0), C.c_uint32(0)
        _lib.check(lib.fgpu_dict_export(self.handle, table.encode(), column.encode(), None, 0, C.byref(n), C.byref(cnt)))
        buf = C.create_string_buffer(max(n.value, 1))
        _lib.check(lib.fgpu_dict_export(self.handle, table.encode(), column.encode(), C.addressof(buf), n.value, C.byref(n), C.byref(cnt)))
        out, raw, p = [], buf.raw, 0
        for _ in range(cnt.value):
            l = int.from_bytes(raw[p:p + 4], "little")
            out.append(raw[p + 4:p + 4 + l])
            p += 4 + l
        return out

    def dict_preload(self, table: str, column: str, values: List[bytes]) -> None:
        """Interns `values` in order (call with the same list on every rank before putting parts)."""
        blob = b"".join(len(v).to_bytes(4, "little") + v for v in values)
        src = (C.c_char * max(len(blob), 1)).from_buffer_copy(blob or b"\0")
        _lib.check(_lib.load().fgpu_dict_preload(self.handle, table.encode(), column.encode(), C.addressof(src), len(blob), len(values)))


def comm_setup(engines: "List[GPUEngine]", slot_bytes: int = 64 << 20, exchange=None) -> None:
    """Opens the mailbox communicator (fgpu_comm_*) over `engines`.

    In-process ranks: pass every rank's engine.  One process per GPU: pass `[engine]` plus `exchange`, a
    function `(rank_handle: bytes) -> List[bytes]` that returns all ranks' handles in rank order (e.g. built on
    torch.distributed.all_gather_object); rank and world size then come from the returned list."""
    lib = _lib.load()
    if exchange is None:
        handles = []
        for r, e in enumerate(engines):
            h = (C.c_uint8 * _lib.COMM_HANDLE_BYTES)()
            _lib.check(lib.fgpu_comm_export(e.handle, r, len(engines), slot_bytes, h))
            handles.append(bytes(h))
        blob = b"".join(handles)
        for e in engines:
            _lib.check(lib.fgpu_comm_open(e.handle, blob))
        return
    raise ValueError("multi-process setup: use GPUEngine.comm_open(rank, n, exchange)")


class Table:
    """The slice of frostdb.Table the scan path sees."""

    def __init__(self, db: "DB", name: str, schema: dp.Schema):
        self.db, self.name, self.schema = db, name, schema

    def Insert(self, columns: Dict[str, object], **write_opts) -> int:
        """One insert -> one part (the state compaction leaves behind, table.go:1267-1355)."""
        buf = dp.write_part(self.schema, columns, **write_opts)
        return self.InsertParquet(buf)

    def InsertParquet(self, buf: bytes) -> int:
        return self.db.engine.put_parquet(self.name, buf)

    def InsertRecord(self, record: pa.RecordBatch) -> int:
        """Table.InsertRecord (table.go:505-560): the record becomes an L0 part as it is (no sort, no Parquet)."""
        return self.db.engine.put_arrow(self.name, record)


class DB:
    def __init__(self, store: "ColumnStore", name: str):
        self.store, self.name = store, name
        self.engine = store.engine
        self.tables: Dict[str, Table] = {}

    def Table(self, name: str, schema: dp.Schema) -> Table:  # db.go Table(name, config)
        if name not in self.tables:
            self.tables[name] = Table(self, name, schema)
        return self.tables[name]

    def TableProvider(self) -> "DBTableProvider":
        return DBTableProvider(self)


class DBTableProvider:
    def __init__(self, db: DB):
        self.db = db

    def GetTable(self, name: str) -> Table:
        return self.db.tables[name]

    def gpu_engine(self) -> GPUEngine:
        return self.db.engine


class ColumnStore:
    def __init__(self, device: int = 0):
        self.engine = GPUEngine(device)
        self.dbs: Dict[str, DB] = {}

    def DB(self, ctx, name: str) -> DB:
        if name not in self.dbs:
            self.dbs[name] = DB(self, name)
        return self.dbs[name]

    def Close(self) -> None:
        self.engine.close()


class PinnedBuffer:
    """Page-locked host memory from fgpu_host_alloc, viewed as a numpy uint8 array (`.array`)."""

    def __init__(self, nbytes: int):
        import numpy as np
        self._lib = _lib.load()
        p = C.c_void_p()
        _lib.check(self._lib.fgpu_host_alloc(nbytes, C.byref(p)))
        self.ptr, self.nbytes = p, nbytes
        self.array = np.frombuffer((C.c_uint8 * nbytes).from_address(p.value), dtype=np.uint8)

    def close(self):
        if self.ptr:
            self.array = None
            self._lib.fgpu_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
