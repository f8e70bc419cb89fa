"""Kernel-time probe of the paths next to the headline: filter-only compaction (cfg 5), dictionary
predicates, unsorted parts, hash-table mode (runs on the GPU box)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_data as bd
from frostdb_b200 import _lib, logicalplan as lp
from frostdb_b200.physicalplan import GPUScan
from frostdb_b200.store import GPUEngine

rows = int(os.environ.get("PROBE_ROWS", 32 * 1024 * 1024))
lib = _lib.load()
eng = GPUEngine(0)


def load(table, sort):
    paths = bd.generate_parts(rows, 16, sort=sort)[: rows // bd.PART_ROWS]
    for p in paths:
        eng.put_parquet(table, np.fromfile(p, dtype=np.uint8))


def run(table, name, kind, f, groups, aggs):
    scan = GPUScan(eng, table, f, kind, groups, aggs)
    q, keep = scan.prepare()
    ms, wall = [], []
    for i in range(4):
        res = C.c_void_p()
        t0 = time.perf_counter()
        _lib.check(lib.fgpu_query_execute(eng.handle, q, eng.table_watermark(table), C.byref(res)))
        wall.append((time.perf_counter() - t0) * 1e3)
        st = eng.stats(res)
        lib.fgpu_result_free(res)
        ms.append(st["scan_kernel_ms"])
    k = min(ms[1:])
    print(f"{name:46s} kernel_ms {k:8.3f} exec_ms {min(wall[1:]):8.3f} rows/s {rows / k * 1e3:.3e} GB/s {st['algorithmic_bytes'] / k / 1e6:8.1f} "
          f"sel {st['rows_selected']:>9d} groups {st['groups']:>8d} runs_rg {st['row_groups_runs']}/{st['row_groups']}", flush=True)
    lib.fgpu_query_free(q)


load("s", True)
ts, val = lp.Col("timestamp"), lp.Col("value")
for s in (0.001, 0.01, 0.1, 0.5, 0.9):
    f = ts.Lt(lp.Literal(bd.T0 + int(s * rows)))
    run("s", f"rows: timestamp < {s:.3f} -> timestamp,value", _lib.PLAN_FILTER, f, [ts, val], [])
for s in (0.001, 0.1, 0.5):
    f = val.Lt(lp.Literal(int(s * 1000)))   # spread over every row group: no pruning
    run("s", f"rows: value < {int(s*1000)} -> timestamp,value", _lib.PLAN_FILTER, f, [ts, val], [])
f = lp.Col("labels.l02").Eq(lp.Literal("v000003"))
run("s", "rows: l02 == v000003 -> timestamp,value", _lib.PLAN_FILTER, f, [ts, val], [])
run("s", "agg: l02 == v000003, sum by l00,l01", _lib.PLAN_AGGREGATE, f, [lp.Col("labels.l00"), lp.Col("labels.l01")], [lp.Sum(val)])
run("s", "agg: value < 500, sum by l00,l01", _lib.PLAN_AGGREGATE, val.Lt(lp.Literal(500)), [lp.Col("labels.l00"), lp.Col("labels.l01")], [lp.Sum(val)])
run("s", "agg: sum by l02 (nullable, short runs)", _lib.PLAN_AGGREGATE, None, [lp.Col("labels.l02")], [lp.Sum(val)])
run("s", "agg: sum,min,max by l00,l01", _lib.PLAN_AGGREGATE, None, [lp.Col("labels.l00"), lp.Col("labels.l01")], [lp.Sum(val), lp.Min(val), lp.Max(val)])
run("s", "agg: sum by l00,l01,l03,l05 (hash table)", _lib.PLAN_AGGREGATE, None, [lp.Col(f"labels.l{k:02d}") for k in (0, 1, 3, 5)], [lp.Sum(val)])
run("s", "distinct l00,l01", _lib.PLAN_DISTINCT, None, [lp.Col("labels.l00"), lp.Col("labels.l01")], [])
eng.drop_table("s")
load("u", False)
run("u", "unsorted: sum by l00,l01", _lib.PLAN_AGGREGATE, None, [lp.Col("labels.l00"), lp.Col("labels.l01")], [lp.Sum(val)])
run("u", "unsorted: ts range 50%, sum by l00,l01", _lib.PLAN_AGGREGATE, lp.And(ts.GtEq(lp.Literal(bd.T0 + rows // 4)), ts.Lt(lp.Literal(bd.T0 + 3 * rows // 4))),
    [lp.Col("labels.l00"), lp.Col("labels.l01")], [lp.Sum(val)])
eng.close()
