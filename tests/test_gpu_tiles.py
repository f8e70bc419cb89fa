"""Tile-aggregate kernel (k_tile_agg: TMA-staged tiles, CTA-private shared-memory table, flat code arrays from
k_flatten) against the oracle, and against the general scan kernel with the tile path switched off
(FROSTGPU_NO_TILE): unsorted parts, nullable keys, dictionary leaves, Min / Max / float64 reducers, absent
dynamic columns, L0 Arrow records, ragged row groups and tiles."""
import ctypes as C
import os

import numpy as np
import pyarrow as pa
import pytest

from frostdb_b200 import _lib
from frostdb_b200 import dynparquet as dp
from frostdb_b200 import logicalplan as lp
from frostdb_b200.physicalplan import GPUScan
from tests.test_gpu_parity import Pair, assert_same
from tests.util import make_columns

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pair(store):
    made = []

    def make(name, schema=None):
        p = Pair(store, name, schema or dp.SampleDefinitionWithFloat())
        made.append(p)
        return p
    yield make
    for p in made:
        p.close()
    for k in ("FROSTGPU_NO_TILE", "FROSTGPU_TA_TILE", "FROSTGPU_TA_STAGES", "FROSTGPU_TA_GLOBAL", "FROSTGPU_TA_CHUNK", "FROSTGPU_TA_CARRY"):
        os.environ.pop(k, None)


def stats_of(p, filt, aggs, groups, kind=_lib.PLAN_AGGREGATE):
    eng = p.store.engine
    scan = GPUScan(eng, p.name, filt, kind, groups, aggs)
    q, keep = scan.prepare()
    lib = _lib.load()
    res = C.c_void_p()
    _lib.check(lib.fgpu_query_execute(eng.handle, q, eng.table_watermark(p.name), C.byref(res)))
    st = eng.stats(res)
    lib.fgpu_result_free(res)
    lib.fgpu_query_free(q)
    return st


def run_both(p, build, float_cols=()):
    """tile path (default), then the general scan kernel: both equal to the oracle"""
    got, exp = p.run(build)
    assert_same(got, exp, float_cols)
    os.environ["FROSTGPU_NO_TILE"] = "1"
    try:
        got2, _ = p.run(build)
    finally:
        os.environ.pop("FROSTGPU_NO_TILE")
    assert_same(got2, exp, float_cols)


SUMCOUNT = [lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))]


@pytest.mark.parametrize("rows,rg,page", [(70_000, 30_000, 4096), (9_001, 4_099, 1024), (130_000, 130_000, 65536)])
def test_unsorted_two_keys(pair, rows, rg, page):
    p = pair("tiles2")
    for i in range(3):
        p.insert(make_columns(rows, 300 + i, {"a": (64, 0.0), "b": (256, 0.0), "c": (16, 0.1)}, t0=i * rows), sort=False,
                 row_group_size=rg, data_page_size=page)
    keys = [lp.Col("labels.a"), lp.Col("labels.b")]
    run_both(p, lambda q: q.Aggregate(SUMCOUNT, keys))
    st = stats_of(p, None, SUMCOUNT, keys)
    assert st["row_groups_tiles"] == st["row_groups"] > 0  # every row group of an unsorted part takes the tile kernel
    assert st["rows_selected"] == 3 * rows
    # ring shapes and the global-table variant give the same records
    for env in ({"FROSTGPU_TA_TILE": "3072", "FROSTGPU_TA_STAGES": "2"}, {"FROSTGPU_TA_TILE": "6144", "FROSTGPU_TA_STAGES": "3"},
                {"FROSTGPU_TA_GLOBAL": "1"}, {"FROSTGPU_TA_CHUNK": "3"}, {"FROSTGPU_TA_CARRY": "1"}):
        os.environ.update(env)
        try:
            got, exp = p.run(lambda q: q.Aggregate(SUMCOUNT, keys))
        finally:
            for k in env:
                os.environ.pop(k)
        assert_same(got, exp)


def test_nullable_and_absent_keys(pair):
    p = pair("tiles_null")
    p.insert(make_columns(50_000, 1, {"a": (5, 0.3), "b": (40, 0.1)}), sort=False, row_group_size=17_000)
    p.insert(make_columns(50_000, 2, {"b": (40, 0.0), "c": (6, 0.5)}), sort=True, row_group_size=50_000)   # short runs of c
    p.insert(make_columns(20_000, 3, {"a": (5, 0.0), "c": (6, 0.0), "d": (2, 0.0)}), sort=False)
    run_both(p, lambda q: q.Aggregate(SUMCOUNT, [lp.Col("labels.a"), lp.Col("labels.c")]))
    run_both(p, lambda q: q.Aggregate(SUMCOUNT, [lp.Col("labels.c")]))
    run_both(p, lambda q: q.Aggregate(SUMCOUNT, [lp.DynCol("labels")]))
    run_both(p, lambda q: q.Aggregate([lp.Count(lp.Col("value"))], []))
    run_both(p, lambda q: q.Distinct(lp.Col("labels.a"), lp.Col("labels.b")))
    st = stats_of(p, None, SUMCOUNT, [lp.Col("labels.c")])
    assert st["row_groups_tiles"] > 0


def test_reducers_min_max_float(pair):
    p = pair("tiles_red")
    for i in range(2):
        cols = make_columns(60_000, 40 + i, {"a": (7, 0.0), "b": (300, 0.1)}, with_float=True, t0=i * 60_000)
        cols["value"] = cols["value"] * 7_000_000_019 - 3_000_000_000_000  # sums beyond 32 bits, negative values
        p.insert(cols, sort=False, row_group_size=25_000)
    aggs = [lp.Sum(lp.Col("value")), lp.Min(lp.Col("value")), lp.Max(lp.Col("value")), lp.Count(lp.Col("value"))]
    run_both(p, lambda q: q.Aggregate(aggs, [lp.Col("labels.a"), lp.Col("labels.b")]))
    run_both(p, lambda q: q.Aggregate(aggs, [lp.Col("labels.a")]))  # few slots: replicated cells
    faggs = [lp.Sum(lp.Col("floatvalue")), lp.Min(lp.Col("floatvalue")), lp.Max(lp.Col("floatvalue")), lp.Sum(lp.Col("timestamp"))]
    run_both(p, lambda q: q.Aggregate(faggs, [lp.Col("labels.b")]), float_cols=("sum(floatvalue)",))
    st = stats_of(p, None, faggs, [lp.Col("labels.b")])
    assert st["row_groups_tiles"] == st["row_groups"]


def test_filters_range_and_dictionary_leaves(pair):
    p = pair("tiles_filter")
    n = 40_000
    for i in range(4):
        p.insert(make_columns(n, 70 + i, {"a": (9, 0.0), "b": (33, 0.2), "c": (4, 0.4)}, with_float=True, t0=i * n), sort=False,
                 row_group_size=15_000)
    p.insert(make_columns(n, 99, {"a": (9, 0.0)}, with_float=True, t0=4 * n), sort=False)  # b, c absent
    ts, v, f = lp.Col("timestamp"), lp.Col("value"), lp.Col("floatvalue")
    keys = [lp.Col("labels.a"), lp.Col("labels.b")]
    filters = [
        lp.And(ts.GtEq(lp.Literal(n // 2)), ts.Lt(lp.Literal(3 * n + 11))),
        lp.And(ts.GtEq(lp.Literal(n)), v.Lt(lp.Literal(400))),
        v.NotEq(lp.Literal(7)),
        f.Gt(lp.Literal(250.5)),
        v.Gt(lp.Literal(99.5)),                                   # int column against a float literal
        lp.Col("labels.b").Eq(lp.Literal("v000003")),
        lp.Col("labels.b").NotEq(lp.Literal("v000003")),
        lp.Col("labels.c").Eq(lp.Literal(None)),                   # selects NULL rows; absent column: all rows
        lp.Col("labels.c").NotEq(lp.Literal(None)),
        lp.And(lp.Col("labels.b").RegexMatch("v00000[1-4]"), ts.Lt(lp.Literal(4 * n))),
        lp.And(lp.Col("labels.c").Contains("002"), v.GtEq(lp.Literal(10))),
        lp.Col("labels.zz").Eq(lp.Literal("x")),                   # column absent everywhere: no row
        lp.Col("labels.zz").NotEq(lp.Literal("x")),                # ... every row
        ts.Gt(lp.Literal(100 * n)),                                # nothing
    ]
    for filt in filters:
        run_both(p, lambda q: q.Filter(filt).Aggregate(SUMCOUNT, keys))
    st = stats_of(p, filters[5], SUMCOUNT, keys)
    assert st["row_groups_tiles"] > 0


def test_l0_arrow_records(pair):
    """Fresh records (parts/arrow.go): 32-bit index streams, validity bitmaps as level streams."""
    from tests.test_gpu_arrow_parts import insert_both
    p = pair("tiles_l0")
    for i in range(3):
        insert_both(p, make_columns(20_000 + 777 * i, 900 + i, {"a": (11, 0.2), "b": (50, 0.0)}, with_float=True, t0=i * 100_000))
    keys = [lp.Col("labels.a"), lp.Col("labels.b")]
    run_both(p, lambda q: q.Aggregate(SUMCOUNT, keys))
    run_both(p, lambda q: q.Filter(lp.Col("labels.a").Eq(lp.Literal(None))).Aggregate(SUMCOUNT, [lp.Col("labels.b")]))
    st = stats_of(p, None, SUMCOUNT, keys)
    assert st["row_groups_tiles"] == st["row_groups"] == 3
