"""The CPU oracle pinned beyond the logictest files: its Parquet decode against pyarrow, and the
engine-level expectations of the reference's Go tests (aggregate_test.go, db_test.go)."""
import pyarrow as pa
import pytest

from frostdb_b200 import dynparquet as dp
from frostdb_b200 import logicalplan as lp
from oracle import oracle as orc
from tests.oracle_scan import OracleEngine, OracleTableHandle, oracle_query
from tests.util import make_columns, rows_of


@pytest.mark.parametrize("version", ["2.0", "1.0"])
def test_oracle_decode_matches_pyarrow(version):
    cols = make_columns(7000, 3, {"a": (9, 0.0), "b": (700, 0.05), "c": (2, 0.9)}, with_float=True, float_null_p=0.4)
    buf = dp.write_part(dp.SampleDefinitionWithFloat(), cols, row_group_size=2500, data_page_size=1024, data_page_version=version)
    t = orc.OracleTable()
    t.add_parquet(buf)
    ref = dp.read_part(buf)
    row = 0
    for rg in range(3):
        n = min(2500, 7000 - row)
        for name in ("labels.a", "labels.b", "labels.c", "timestamp", "value", "floatvalue", "stacktrace"):
            exp = ref[name].slice(row, n)
            if pa.types.is_dictionary(exp.type):
                exp = exp.cast(pa.string())
            assert t.decode_column(0, rg, name) == exp.to_pylist(), (rg, name)
        row += n
    t.close()


def _samples_inconsistent_schema(eng):
    """aggregate_test.go:40-73: three single-row inserts with different dynamic label columns."""
    t = OracleTableHandle(eng, "test", dp.SampleDefinition())
    for labels, ts, v in (({"label1": "value1"}, 1, 1), ({"label2": "value2"}, 2, 2), ({"label2": "value2"}, 3, 3)):
        cols = {"example_type": [""], "stacktrace": ["s"], "timestamp": [ts], "value": [v]}
        for k, val in labels.items():
            cols["labels." + k] = [val]
        t.Insert(cols)


@pytest.mark.parametrize("fn,alias,expected", [
    (lp.Sum, "value_sum", [5, 1]), (lp.Min, "value_min", [2, 1]), (lp.Max, "value_max", [3, 1]),
    (lp.Count, "value_count", [2, 1]), (lp.Avg, "value_avg", [2, 1])])
def test_aggregate_inconsistent_schema(fn, alias, expected):
    """TestAggregateInconsistentSchema, aggregate_test.go:23-148 (expected values :85-110)."""
    eng = OracleEngine(threads=2)
    try:
        _samples_inconsistent_schema(eng)
        out = []
        oracle_query(eng, "test").Aggregate([fn(lp.Col("value"))], [lp.Col("labels.label2")]) \
            .Project(fn(lp.Col("value")).Alias(alias)).Execute(None, lambda c, r: out.append(r))
        assert len(out) == 1 and out[0].schema.names == [alias]
        assert sorted(out[0].column(0).to_pylist(), reverse=True) == expected
    finally:
        eng.close()


def test_aggregation_projection_union_of_dynamic_columns():
    """TestAggregationProjection, aggregate_test.go:150-258: group by stacktrace + every label column;
    the result is the union of the dynamic columns, NULL where a group never saw one (:246-257)."""
    eng = OracleEngine(threads=3)
    try:
        t = OracleTableHandle(eng, "test", dp.SampleDefinition())
        for labels, ts, v in (({"label1": "value1"}, 1, 1), ({"label2": "value2"}, 2, 2), ({"label2": "value2", "label3": "x"}, 3, 3)):
            cols = {"example_type": [""], "stacktrace": ["s"], "timestamp": [ts], "value": [v]}
            for k, val in labels.items():
                cols["labels." + k] = [val]
            t.Insert(cols)
        out = []
        oracle_query(eng, "test").Aggregate([lp.Sum(lp.Col("value")), lp.Max(lp.Col("value"))],
                                            [lp.DynCol("labels"), lp.Col("stacktrace")]).Execute(None, lambda c, r: out.append(r))
        names = ["labels.label1", "labels.label2", "labels.label3", "stacktrace", "sum(value)", "max(value)"]
        assert rows_of(out, names) == sorted([("value1", None, None, "s", 1, 1), (None, "value2", None, "s", 2, 2),
                                              (None, "value2", "x", "s", 3, 3)], key=lambda r: tuple((x is None, x) for x in r))
    finally:
        eng.close()


def test_row_group_filter_never_changes_a_result(monkeypatch):
    """The oracle's port of LSM.Scan's TrueNegativeFilter (index/lsm.go:401-454, expr/binaryscalarexpr.go)
    drops row groups by their chunk statistics; the records must be those of a scan of every row group."""
    eng = OracleEngine(threads=2)
    try:
        t = OracleTableHandle(eng, "t", dp.SampleDefinitionWithFloat())
        n = 5_000
        for i in range(4):
            t.Insert(make_columns(n, 900 + i, {"a": (4, 0.2), "b": (9, 0.0)}, with_float=True, float_null_p=0.3, t0=i * n),
                     row_group_size=2_000)
        ts, v, f = lp.Col("timestamp"), lp.Col("value"), lp.Col("floatvalue")
        filters = [
            lp.And(ts.GtEq(lp.Literal(n + 3)), ts.Lt(lp.Literal(3 * n - 1))), ts.Eq(lp.Literal(2 * n)), ts.Eq(lp.Literal(40 * n)),
            ts.NotEq(lp.Literal(7)), ts.LtEq(lp.Literal(-1)), ts.Gt(lp.Literal(4 * n - 2)), v.Gt(lp.Literal(998)), v.Lt(lp.Literal(0)),
            f.GtEq(lp.Literal(999.5)), f.Lt(lp.Literal(250.25)), v.Lt(lp.Literal(10.5)),
            lp.Or(ts.Lt(lp.Literal(10)), ts.GtEq(lp.Literal(4 * n - 10))), lp.And(ts.Lt(lp.Literal(n)), lp.Col("labels.a").Eq(lp.Literal("v000001"))),
            lp.Col("labels.a").Eq(lp.Literal("zzz")), lp.Col("labels.a").Eq(lp.Literal(None)), lp.Col("labels.zz").NotEq(lp.Literal(None)),
            lp.Col("nope").Lt(lp.Literal(4)),
        ]
        scanned = []
        for flt in filters:
            outs = []
            for off in (False, True):
                if off:
                    monkeypatch.setenv("FROST_ORACLE_NO_PRUNE", "1")
                else:
                    monkeypatch.delenv("FROST_ORACLE_NO_PRUNE", raising=False)
                out = []
                oracle_query(eng, "t").Filter(flt).Aggregate([lp.Sum(v), lp.Count(v), lp.Max(f)], [lp.Col("labels.b")]) \
                    .Execute(None, lambda c, r: out.append(r))
                outs.append(rows_of(out))
            assert outs[0] == outs[1], flt.Name()
            scanned.append(len(outs[0]))
        assert scanned[2] == 0 and scanned[0] > 0
    finally:
        eng.close()


def test_oracle_row_group_filter_asks_bloom_filters():
    """expr/binaryscalarexpr.go:104-118: a chunk WITH a bloom filter answers `column == v` from the filter alone (the
    bounds are not consulted), one without from its bounds.  The oracle's XXH64 + split-block check against the
    plain-Python implementation of tests/bloom_file.py on a file the tests write (pyarrow reads it back)."""
    import numpy as np
    from oracle import oracle as orc
    from tests import bloom_file as bf
    rng = np.random.default_rng(9)
    xs = sorted(set(int(v) for v in rng.integers(0, 500_000, 3_000)))
    buf = bf.write_int64_file({"x": xs})
    n_bytes = max(32, ((len(xs) * 10 // 8) + 31) // 32 * 32)
    bits = bytearray(n_bytes)
    for v in xs:
        bf.sbbf_insert(bits, bf.xxh64(v.to_bytes(8, "little", signed=True)))
    for v in xs[::29]:
        assert orc.parquet_rowgroup_may_match_eq_i64(buf, 0, "x", v)
    out = 0
    for v in (int(v) for v in rng.integers(-1000, 501_000, 400)):
        expect = bf.sbbf_check(bytes(bits), bf.xxh64(v.to_bytes(8, "little", signed=True)))  # also OUTSIDE the bounds: the filter alone decides
        assert orc.parquet_rowgroup_may_match_eq_i64(buf, 0, "x", v) == expect, v
        out += 0 if expect else 1
    assert out > 300
    assert not orc.parquet_rowgroup_may_match_eq_i64(buf, 0, "x", None)   # required column: `== NULL` finds nothing
    assert not orc.parquet_rowgroup_may_match_eq_i64(buf, 0, "nope", 3) and orc.parquet_rowgroup_may_match_eq_i64(buf, 0, "nope", None)
