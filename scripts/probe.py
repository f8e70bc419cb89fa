"""Kernel-time probe: which part of the scan costs what (runs on the GPU box)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_data as bd
from frostdb_b200 import _lib, logicalplan as lp
from frostdb_b200.physicalplan import GPUScan
from frostdb_b200.store import GPUEngine

rows = int(os.environ.get("PROBE_ROWS", 32 * 1024 * 1024))
paths = bd.generate_parts(100_000_000, 16)[: rows // bd.PART_ROWS]
bufs = [np.fromfile(p, dtype=np.uint8) for p in paths]
eng = GPUEngine(0)
for b in bufs:
    eng.put_parquet("t", b)
lib = _lib.load()
lo, hi = bd.T0 + rows // 4, bd.T0 + 3 * rows // 4
F = lp.And(lp.Col("timestamp").GtEq(lp.Literal(lo)), lp.Col("timestamp").Lt(lp.Literal(hi)))
F1 = lp.Col("timestamp").GtEq(lp.Literal(lo))
cases = {
    "count, no filter, no keys": (None, [lp.Count(lp.Col("value"))], []),
    "sum, no filter, no keys": (None, [lp.Sum(lp.Col("value"))], []),
    "count, filter(1 leaf)": (F1, [lp.Count(lp.Col("value"))], []),
    "count, filter(2 leaves)": (F, [lp.Count(lp.Col("value"))], []),
    "count, key l00": (None, [lp.Count(lp.Col("value"))], [lp.Col("labels.l00")]),
    "count, keys l00,l01": (None, [lp.Count(lp.Col("value"))], [lp.Col("labels.l00"), lp.Col("labels.l01")]),
    "sum, keys l00,l01": (None, [lp.Sum(lp.Col("value"))], [lp.Col("labels.l00"), lp.Col("labels.l01")]),
    "headline": (F, [lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.Col("labels.l00"), lp.Col("labels.l01")]),
    "sum, key l05 (random, nulls)": (None, [lp.Sum(lp.Col("value"))], [lp.Col("labels.l05")]),
}
for name, (f, aggs, groups) in cases.items():
    scan = GPUScan(eng, "t", f, _lib.PLAN_AGGREGATE, groups, aggs)
    q, keep = scan.prepare()
    ms = []
    for i in range(5):
        res = C.c_void_p()
        _lib.check(lib.fgpu_query_execute(eng.handle, q, eng.table_watermark("t"), C.byref(res)))
        st = eng.stats(res)
        lib.fgpu_result_free(res)
        ms.append(st["scan_kernel_ms"])
    print(f"{name:34s} kernel_ms min {min(ms[1:]):8.3f}  rows/s {rows / min(ms[1:]) * 1e3:.3e}  GB/s {st['algorithmic_bytes'] / min(ms[1:]) / 1e6:8.1f} groups {st['groups']}", flush=True)
    lib.fgpu_query_free(q)
eng.close()
