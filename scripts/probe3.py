"""Kernel-time probe of the tile-aggregate kernel (k_tile_agg) on 32 Mi rows: unsorted parts, short-run / nullable
keys on sorted parts, reducers, ring shapes (runs on the GPU box)."""
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


ONLY = os.environ.get("PROBE_ONLY")


def run(table, name, kind, f, groups, aggs, env=None):
    env = env or {}
    if ONLY and ONLY not in name + " " + " ".join(f"{a}={b}" for a, b in env.items()):
        return
    os.environ.update(env)
    try:
        scan = GPUScan(eng, table, f, kind, groups, aggs)
        q, keep = scan.prepare()
        ms, wall = [], []
        for i in range(5):
            res = C.c_void_p()
            t0 = time.perf_counter()
            _lib.check(lib.fgpu_query_execute(eng.handle, q, eng.table_watermark(table), C.byref(res)))
            wall.append((time.perf_counter() - t0) * 1e3)
            st = eng.stats(res)
            lib.fgpu_result_free(res)
            ms.append(st["scan_kernel_ms"])
        k = min(ms[1:])
        tag = " ".join(f"{a.replace('FROSTGPU_', '')}={b}" for a, b in env.items())
        print(f"{name:40s} {tag:28s} kernel_ms {k:7.3f} exec_ms {min(wall[1:]):7.3f} rows/s {rows / k * 1e3:.3e} GB/s {st['algorithmic_bytes'] / k / 1e6:7.1f} "
              f"sel {st['rows_selected']:>9d} groups {st['groups']:>7d} rg runs/tiles/all {st['row_groups_runs']}/{st['row_groups_tiles']}/{st['row_groups']}", flush=True)
        lib.fgpu_query_free(q)
    finally:
        for a in env:
            os.environ.pop(a)


ts, val = lp.Col("timestamp"), lp.Col("value")
K01 = [lp.Col("labels.l00"), lp.Col("labels.l01")]
AGG = _lib.PLAN_AGGREGATE
load("u", False)
run("u", "unsorted: sum by l00,l01", AGG, None, K01, [lp.Sum(val)])
for t, s in ((3072, 4), (3072, 3), (3072, 2), (6144, 2)):
    run("u", "unsorted: sum by l00,l01", AGG, None, K01, [lp.Sum(val)], {"FROSTGPU_TA_TILE": str(t), "FROSTGPU_TA_STAGES": str(s)})
run("u", "unsorted: sum by l00,l01", AGG, None, K01, [lp.Sum(val)], {"FROSTGPU_TA_GLOBAL": "1"})
run("u", "unsorted: sum by l00,l01", AGG, None, K01, [lp.Sum(val)], {"FROSTGPU_NO_TILE": "1"})
run("u", "unsorted: sum,count by l00,l01", AGG, None, K01, [lp.Sum(val), lp.Count(val)])
run("u", "unsorted: count by l00,l01", AGG, None, K01, [lp.Count(val)])
F50 = lp.And(ts.GtEq(lp.Literal(bd.T0 + rows // 4)), ts.Lt(lp.Literal(bd.T0 + 3 * rows // 4)))
run("u", "unsorted: ts range 50%, sum by l00,l01", AGG, F50, K01, [lp.Sum(val)])
run("u", "unsorted: value < 500, sum by l00,l01", AGG, val.Lt(lp.Literal(500)), K01, [lp.Sum(val)])
run("u", "unsorted: sum,min,max by l00,l01", AGG, None, K01, [lp.Sum(val), lp.Min(val), lp.Max(val)])
run("u", "unsorted: sum by l02 (17 slots)", AGG, None, [lp.Col("labels.l02")], [lp.Sum(val)])
run("u", "unsorted: min,max by l02 (17 slots)", AGG, None, [lp.Col("labels.l02")], [lp.Min(val), lp.Max(val)])
run("u", "unsorted: sum by l05 (129 slots)", AGG, None, [lp.Col("labels.l05")], [lp.Sum(val)])
run("u", "unsorted: l02 == v000003, sum by l00,l01", AGG, lp.Col("labels.l02").Eq(lp.Literal("v000003")), K01, [lp.Sum(val)])
run("u", "unsorted: sum by l00,l01,l02 (284k slots)", AGG, None, K01 + [lp.Col("labels.l02")], [lp.Sum(val)])
run("u", "unsorted: sum (no keys)", AGG, None, [], [lp.Sum(val)])
run("u", "unsorted: distinct l00,l01", _lib.PLAN_DISTINCT, None, K01, [])
eng.drop_table("u")
load("s", True)
run("s", "sorted: sum by l02 (short runs)", AGG, None, [lp.Col("labels.l02")], [lp.Sum(val)])
run("s", "sorted: sum by l02 (short runs)", AGG, None, [lp.Col("labels.l02")], [lp.Sum(val)], {"FROSTGPU_NO_TILE": "1"})
run("s", "sorted: sum by l00,l02", AGG, None, [lp.Col("labels.l00"), lp.Col("labels.l02")], [lp.Sum(val)])
run("s", "sorted: sum by l00,l01 (k_runs)", AGG, None, K01, [lp.Sum(val)])
run("s", "sorted: sum by l00,l01 (tiles)", AGG, None, K01, [lp.Sum(val)], {"FROSTGPU_NO_RUNS": "1"})
run("s", "sorted: sum by l05", AGG, None, [lp.Col("labels.l05")], [lp.Sum(val)])
eng.close()
