"""L0 parts: Arrow records handed to the engine as they are (fgpu_part_put_arrow; parts/arrow.go:14-55,
table.go:806-814).  The oracle gets the same rows as an unsorted Parquet part."""
import numpy as np
import pyarrow as pa
import pytest

from frostdb_b200 import _lib
from frostdb_b200 import dynparquet as dp
from frostdb_b200 import logicalplan as lp
from tests.test_gpu_parity import Pair, assert_same
from tests.util import make_columns, rows_of

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


def as_record(buf, *, index_type=None, plain=(), offset=0):
    """The rows of a part as ONE Arrow record (dictionary label columns, as pqarrow builds them)."""
    t = dp.read_part(buf).combine_chunks()
    cols, names = [], []
    for name in t.schema.names:
        a = t.column(name).chunk(0) if t.num_rows else pa.array([], t.schema.field(name).type)
        if pa.types.is_string(a.type) or pa.types.is_binary(a.type) or pa.types.is_large_string(a.type):
            a = a.cast(pa.binary()).dictionary_encode()
        if pa.types.is_dictionary(a.type):
            if name in plain:
                a = a.dictionary_decode().cast(pa.binary())
            else:
                a = pa.DictionaryArray.from_arrays(a.indices.cast(index_type or pa.uint32()), a.dictionary.cast(pa.binary()))
        cols.append(a)
        names.append(name)
    rb = pa.RecordBatch.from_arrays(cols, names=names)
    return rb.slice(offset) if offset else rb


def insert_both(p, cols, **rec_opts):
    """unsorted Parquet part into the oracle, the same rows as an Arrow record into the GPU engine"""
    off = rec_opts.pop("offset", 0)
    buf = dp.write_part(p.schema, cols, sort=False)
    rb = as_record(buf, **rec_opts)
    if off:
        n = rb.num_rows
        pad = pa.RecordBatch.from_arrays([pa.concat_arrays([c.slice(0, off), c]) for c in rb.columns], names=rb.schema.names)
        rb = pad.slice(off, n)  # children carry a non-zero offset
    p.gt.InsertRecord(rb)
    p.ot.InsertParquet(buf)


KEYS = [lp.Col("labels.a"), lp.Col("labels.b")]


@pytest.mark.parametrize("index_type,offset", [(pa.uint32(), 0), (pa.int32(), 37), (pa.int8(), 0), (pa.uint16(), 129)])
def test_aggregates_over_arrow_parts(pair, index_type, offset):
    p = pair("l0_agg")
    for i in range(3):
        cols = make_columns(20_011, 700 + i, {"a": (7, 0.0), "b": (90, 0.15), "c": (3, 0.6)}, with_float=True, float_null_p=0.25, t0=i * 20_011)
        insert_both(p, cols, index_type=index_type, offset=offset)
    v, fv = lp.Col("value"), lp.Col("floatvalue")
    got, exp = p.run(lambda q: q.Aggregate([lp.Sum(v), lp.Min(v), lp.Max(v), lp.Count(v), lp.Sum(fv), lp.Max(fv)], KEYS))
    assert_same(got, exp, ("sum(floatvalue)",))
    got, exp = p.run(lambda q: q.Aggregate([lp.Sum(v), lp.Count(fv)], [lp.DynCol("labels")]))
    assert_same(got, exp)
    got, exp = p.run(lambda q: q.Distinct(lp.Col("labels.c"), lp.Col("labels.a")))
    assert_same(got, exp)


def test_filters_and_projection_over_arrow_parts(pair):
    p = pair("l0_filter")
    n = 30_000
    for i in range(2):
        insert_both(p, make_columns(n, 710 + i, {"a": (6, 0.3), "b": (9, 0.0)}, with_float=True, float_null_p=0.3, t0=i * n))
    ts, v, fv, a = lp.Col("timestamp"), lp.Col("value"), lp.Col("floatvalue"), lp.Col("labels.a")
    filters = [
        lp.And(ts.GtEq(lp.Literal(n // 2)), ts.Lt(lp.Literal(n + 100))), ts.Gt(lp.Literal(10 * n)), v.Lt(lp.Literal(250)),
        fv.GtEq(lp.Literal(500.5)), a.Eq(lp.Literal("v000002")), a.NotEq(lp.Literal("v000002")), a.Eq(lp.Literal(None)),
        a.NotEq(lp.Literal(None)), a.RegexMatch("v00000[1-3]$"), a.Contains("0004"), lp.Col("labels.zz").Eq(lp.Literal("")),
        lp.Or(lp.And(a.Eq(lp.Literal("v000001")), v.Lt(lp.Literal(300))), ts.Gt(lp.Literal(2 * n - 50))),
    ]
    for f in filters:
        try:
            got, exp = p.run(lambda q: q.Filter(f).Aggregate([lp.Sum(v), lp.Count(v)], [lp.Col("labels.b")]))
            assert_same(got, exp)
        except AssertionError as e:
            raise AssertionError(f"filter {f.Name()}: {e}") from e
    names = ["timestamp", "value", "labels.a"]
    got, exp = p.run(lambda q: q.Filter(filters[0]).Project(ts, v, a))
    assert rows_of(got, names) == rows_of(exp, names) and len(rows_of(got, names)) > 0


def test_mixed_levels_and_plain_string_columns(pair):
    """a compacted (sorted Parquet) part next to fresh L0 records, one of them with non-dictionary strings"""
    p = pair("l0_mixed", dp.SampleDefinition())
    p.insert(make_columns(25_000, 720, {"a": (5, 0.0), "b": (11, 0.0)}, t0=0), row_group_size=10_000)
    insert_both(p, make_columns(9_999, 721, {"a": (5, 0.1), "b": (11, 0.0)}, t0=25_000))
    insert_both(p, make_columns(1, 722, {"a": (5, 0.0), "b": (11, 0.0)}, t0=40_000))
    insert_both(p, make_columns(4_097, 723, {"a": (5, 0.0), "d": (4, 0.5)}, t0=50_000), plain=("labels.a",))
    v = lp.Col("value")
    got, exp = p.run(lambda q: q.Aggregate([lp.Sum(v), lp.Count(v)], KEYS))
    assert_same(got, exp)
    f = lp.Col("timestamp").GtEq(lp.Literal(20_000))
    got, exp = p.run(lambda q: q.Filter(f).Aggregate([lp.Sum(v), lp.Max(v)], [lp.DynCol("labels")]))
    assert_same(got, exp)


def test_unsupported_arrow_layouts_fail_loudly(store):
    eng = store.engine
    rb = pa.RecordBatch.from_arrays([pa.array([True, False])], names=["flag"])
    with pytest.raises(_lib.FrostGPUError):
        eng.put_arrow("l0_bad", rb)
    rb = pa.RecordBatch.from_arrays([pa.array([[1], [2]])], names=["nested"])
    with pytest.raises(_lib.FrostGPUError):
        eng.put_arrow("l0_bad", rb)
    eng.drop_table("l0_bad")
