"""Parity at the benchmark's own shapes: parts of 4 Mi rows written by bench_data.generate_parts (1 Mi-row row groups,
sorted in compaction order or left in arrival order), the headline query and its neighbours, compared bit for bit with
the oracle on the same files.  The parity tests elsewhere use row groups of a few thousand rows; this one reaches the
4096-row tiles, the ticketed chunks and the multi-part statistics pruning the 100 M-row benchmark runs with."""
import ctypes as C

import numpy as np
import pytest

import bench
import bench_data as bd
from frostdb_b200 import _lib
from frostdb_b200 import logicalplan as lp
from frostdb_b200.physicalplan import GPUScan

pytestmark = pytest.mark.gpu

ROWS = 8 * 1024 * 1024


@pytest.mark.parametrize("sort", [True, False])
def test_bench_shaped_parts_match_the_oracle(store, sort):
    eng = store.engine
    table = f"bench_scale_{int(sort)}"
    paths = bd.generate_parts(ROWS, 16, sort=sort)
    bufs = bench.load_files(paths)
    for b in bufs:
        eng.put_parquet(table, b)
    try:
        ts, val = lp.Col("timestamp"), lp.Col("value")
        k01 = [lp.Col("labels.l00"), lp.Col("labels.l01")]
        f50 = lp.And(ts.GtEq(lp.Literal(bd.T0 + ROWS // 4)), ts.Lt(lp.Literal(bd.T0 + 3 * ROWS // 4)))
        funny = lp.And(ts.GtEq(lp.Literal(bd.T0 + 1_000_003)), ts.Lt(lp.Literal(bd.T0 + 6_000_001)))  # cuts row groups in the middle
        cases = [
            (f50, k01, [lp.Sum(val), lp.Count(val)]),                      # the headline
            (None, k01, [lp.Sum(val), lp.Count(val)]),                     # cfg 3 as BASELINE.json states it
            (funny, k01, [lp.Sum(val), lp.Count(val)]),                    # row filter inside the boundary row groups
            (None, [lp.Col("labels.l02")], [lp.Sum(val)]),                 # short runs / nullable key: tile aggregate
            (lp.And(funny, val.Lt(lp.Literal(500))), [lp.Col("labels.l00")], [lp.Sum(val), lp.Min(val), lp.Max(val)]),
            (lp.Col("labels.l02").Eq(lp.Literal("v000003")), [], [lp.Sum(val), lp.Count(val)]),
        ]
        for filt, groups, aggs in cases:
            scan = GPUScan(eng, table, filt, _lib.PLAN_AGGREGATE, groups, aggs)
            q, keep = scan.prepare()
            lib = _lib.load()
            exp, _, _ = bench.oracle_rows(bufs, scan, [a.Name() for a in aggs], threads=8)
            for _ in range(2):  # the second Execute runs from the plan's execution cache
                res = C.c_void_p()
                _lib.check(lib.fgpu_query_execute(eng.handle, q, eng.table_watermark(table), C.byref(res)))
                got = bench.result_rows(list(eng.drain(res)))
                lib.fgpu_result_free(res)
                par = bench.parity_of(got, exp)
                assert par["mismatches"] == 0 and par["groups"] == par["gpu_groups"], (filt.Name() if filt else None, [g.Name() for g in groups], par)
            lib.fgpu_query_free(q)
    finally:
        eng.drop_table(table)
