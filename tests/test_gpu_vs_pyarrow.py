"""GPU results against an independent library (pyarrow/Acero) on random tables: a second opinion
next to the oracle-based parity tests."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from frostdb_b200 import dynparquet as dp
from frostdb_b200 import logicalplan as lp
from frostdb_b200 import query
from frostdb_b200 import _lib
from tests.util import make_columns, rows_of

pytestmark = pytest.mark.gpu


def _collect(qb):
    out = []
    qb.Execute(None, lambda ctx, r: out.append(r))
    return out


def _setup(store, name, parts, schema=None):
    db = store.DB(None, "test")
    store.engine.drop_table(name) if name in db.tables else None
    db.tables.pop(name, None)
    t = db.Table(name, schema or dp.SampleDefinitionWithFloat())
    bufs = []
    for cols, opts in parts:
        buf = dp.write_part(t.schema, cols, **opts)
        bufs.append(buf)
        t.InsertParquet(buf)
    ref = pa.concat_tables([dp.read_part(b) for b in bufs], promote_options="default")
    return db, ref


@pytest.mark.parametrize("sort", [True, False])
@pytest.mark.parametrize("page_version", ["2.0", "1.0"])
def test_groupby_sum_count_min_max(store, sort, page_version):
    n = 50_000
    parts = []
    for p in range(3):
        cols = make_columns(n, 100 + p, {"a": (7, 0.0), "b": (300, 0.1), "c": (3, 0.5)}, with_float=True,
                            float_null_p=0.2, t0=1_000_000 + p * n)
        parts.append((cols, dict(sort=sort, row_group_size=17_000, data_page_size=8192, data_page_version=page_version)))
    db, ref = _setup(store, "t_gb", parts)
    eng = query.NewEngine(None, db.TableProvider())
    got = _collect(eng.ScanTable("t_gb").Aggregate(
        [lp.Sum(lp.Col("value")), lp.Count(lp.Col("value")), lp.Min(lp.Col("value")), lp.Max(lp.Col("value")),
         lp.Sum(lp.Col("floatvalue")), lp.Min(lp.Col("timestamp"))],
        [lp.Col("labels.a"), lp.Col("labels.b")]))
    names = ["labels.a", "labels.b", "sum(value)", "count(value)", "min(value)", "max(value)", "min(timestamp)"]
    rows = rows_of(got, names)
    exp_t = ref.group_by(["labels.a", "labels.b"], use_threads=False).aggregate(
        [("value", "sum"), ("value", "count"), ("value", "min"), ("value", "max"), ("timestamp", "min")])
    exp = rows_of([b for b in exp_t.select(["labels.a", "labels.b", "value_sum", "value_count", "value_min",
                                            "value_max", "timestamp_min"]).to_batches()])
    assert rows == exp
    # float sums within 1e-9 relative (NULL floatvalue contributes 0)
    fs = {r[:2]: r[2] for r in rows_of(got, ["labels.a", "labels.b", "sum(floatvalue)"])}
    exp_f = ref.group_by(["labels.a", "labels.b"], use_threads=False).aggregate([("floatvalue", "sum")])
    for a, b, s in rows_of(exp_f.select(["labels.a", "labels.b", "floatvalue_sum"]).to_batches()):
        s = 0.0 if s is None else s
        assert fs[(a, b)] == pytest.approx(s, rel=1e-9, abs=1e-9)


def test_filter_int_and_dict(store):
    n = 40_000
    cols = make_columns(n, 7, {"a": (5, 0.0), "b": (50, 0.3)}, t0=0)
    db, ref = _setup(store, "t_f", [(cols, dict(row_group_size=9_000, data_page_size=4096))], dp.SampleDefinition())
    eng = query.NewEngine(None, db.TableProvider())
    f = lp.And(lp.Col("timestamp").GtEq(lp.Literal(10_000)), lp.Col("timestamp").Lt(lp.Literal(30_000)),
               lp.Or(lp.Col("labels.b").Eq(lp.Literal("v000007")), lp.Col("labels.b").Eq(lp.Literal(None))))
    got = _collect(eng.ScanTable("t_f").Filter(f).Aggregate([lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))],
                                                           [lp.Col("labels.a")]))
    rows = rows_of(got, ["labels.a", "sum(value)", "count(value)"])
    ts, b = ref["timestamp"], ref["labels.b"].cast(pa.string())
    mask = pc.and_(pc.and_(pc.greater_equal(ts, 10_000), pc.less(ts, 30_000)),
                   pc.or_(pc.fill_null(pc.equal(b, "v000007"), False), pc.is_null(b)))
    sub = ref.filter(mask)
    exp_t = sub.group_by(["labels.a"], use_threads=False).aggregate([("value", "sum"), ("value", "count")])
    exp = rows_of(exp_t.select(["labels.a", "value_sum", "value_count"]).to_batches())
    assert rows == exp


def test_decode_column_matches_pyarrow(store):
    n = 30_000
    cols = make_columns(n, 3, {"a": (9, 0.0), "b": (70_000, 0.05), "c": (2, 0.9)}, with_float=True, float_null_p=0.4)
    db, ref = _setup(store, "t_dec", [(cols, dict(row_group_size=11_000, data_page_size=2048))])
    for name in ["labels.a", "labels.b", "labels.c", "timestamp", "value", "floatvalue", "stacktrace"]:
        got = store.engine.decode_column("t_dec", 0, name)
        if pa.types.is_dictionary(got.type):
            got = got.dictionary_decode().cast(pa.string())
        exp = ref[name].combine_chunks()
        if pa.types.is_dictionary(exp.type):
            exp = exp.dictionary_decode()
        assert got.to_pylist() == exp.cast(got.type).to_pylist(), name


def test_int64_dictionary_pages_on_the_gpu(store):
    """north_star: RLE-dictionary pages of int64 columns (CK_DICT64), decoded and aggregated on the GPU: the column
    decode equals pyarrow's read, Sum / Min / Max over it equal pyarrow's, also as a filter column."""
    import io
    import pyarrow.parquet as pq
    n = 40_000
    rng = np.random.default_rng(3)
    ts = np.sort(rng.integers(0, 500, n)).astype(np.int64) * 1000            # long runs
    val = rng.integers(-50, 50, n).astype(np.int64)                            # bit-packed indices
    val_arr = pa.array(val, mask=rng.random(n) < 0.1)
    lab = pa.array([f"v{j:03d}" for j in rng.integers(0, 9, n)]).dictionary_encode()
    t = pa.table({"labels.a": lab, "timestamp": pa.array(ts), "value": val_arr})
    t = t.cast(pa.schema([pa.field("labels.a", lab.type, nullable=True), pa.field("timestamp", pa.int64(), nullable=False),
                          pa.field("value", pa.int64(), nullable=True)]))
    sink = io.BytesIO()
    pq.write_table(t, sink, compression="NONE", use_dictionary=True, data_page_version="2.0", store_schema=False, row_group_size=15_000,
                   data_page_size=4096)
    eng = store.engine
    name = "dict64"
    eng.drop_table(name)
    try:
        pid = eng.put_parquet(name, sink.getvalue())
        for col in ("timestamp", "value"):
            got = eng.decode_column(name, pid, col)
            assert got.to_pylist() == t.column(col).combine_chunks().to_pylist()

        class _P:
            def gpu_engine(self):
                return eng
        q = query.NewEngine(None, _P()).ScanTable(name)
        out = _collect(q.Filter(lp.Col("timestamp").GtEq(lp.Literal(100_000))).Aggregate(
            [lp.Sum(lp.Col("timestamp")), lp.Max(lp.Col("timestamp")), lp.Min(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.Col("labels.a")]))
        rows = rows_of(out, ["labels.a", "sum(timestamp)", "max(timestamp)", "min(value)", "count(value)"])
        ref = t.filter(pc.greater_equal(t["timestamp"], 100_000))
        ref = ref.set_column(ref.schema.get_field_index("value"), "value", pc.fill_null(ref["value"], 0))  # NULL slots hold 0 (optbuilders.go:337-340)
        exp_t = ref.group_by(["labels.a"], use_threads=False).aggregate([("timestamp", "sum"), ("timestamp", "max"), ("value", "min"), ("value", "count")])
        exp = rows_of(exp_t.select(["labels.a", "timestamp_sum", "timestamp_max", "value_min", "value_count"]).to_batches())
        assert rows == exp
    finally:
        eng.drop_table(name)


def test_converted_parts_give_the_reference_dictionaries_and_indices(store):
    """pqarrow/arrow_test.go:17-145 (TestDifferentSchemasToArrow): five one-row buffers whose dynamic label columns
    differ, converted one after the other.  The reference's record holds, per label column, ONE dictionary in
    first-seen order and the indices [0 1 2 0 0] / [0 0 0 0 0] / [null 0 null null 0] / [null null 0 null null];
    timestamps and values [1 2 3 2 3].  Here the parts are decoded on the GPU column by column (K1) against the
    table's global dictionaries, which must come out exactly like that."""
    from frostdb_b200 import dynparquet as dp
    schema = dp.SampleDefinition()
    db = store.DB(None, "golden_arrow")
    name = "t_arrow_test"
    if name in db.tables:
        store.engine.drop_table(name)
        db.tables.pop(name)
    t = db.Table(name, schema)
    samples = [({"label1": "value1", "label2": "value2"}, 1, 1),
               ({"label1": "value2", "label2": "value2", "label3": "value3"}, 2, 2),
               ({"label1": "value3", "label2": "value2", "label4": "value4"}, 3, 3),
               ({"label1": "value1", "label2": "value2"}, 2, 2),
               ({"label1": "value1", "label2": "value2", "label3": "value3"}, 3, 3)]
    pids = []
    for labels, ts, val in samples:
        cols = {"example_type": [""], "stacktrace": ["s"], "timestamp": [ts], "value": [val]}
        cols.update({"labels." + k: [v] for k, v in labels.items()})
        pids.append(t.Insert(cols))
    expect = {"labels.label1": ([b"value1", b"value2", b"value3"], [0, 1, 2, 0, 0]),
              "labels.label2": ([b"value2"], [0, 0, 0, 0, 0]),
              "labels.label3": ([b"value3"], [None, 0, None, None, 0]),
              "labels.label4": ([b"value4"], [None, None, 0, None, None]),
              "example_type": ([b""], [0, 0, 0, 0, 0])}
    for col, (dictionary, indices) in expect.items():
        assert store.engine.dict_export(name, col) == dictionary, col
        got = []
        for i, pid in enumerate(pids):
            try:
                a = store.engine.decode_column(name, pid, col)
            except _lib.FrostGPUError:
                a = None  # the part has no such column: NULL for its rows (arrow.go:485-597)
            if a is None or a.null_count == len(a):
                got.append(None)
                continue
            assert pa.types.is_dictionary(a.type)
            value = a.dictionary_decode()[0].as_py()
            value = value if isinstance(value, bytes) else value.encode()
            got.append(dictionary.index(value))
        assert got == indices, col
    for col in ("timestamp", "value"):
        got = [store.engine.decode_column(name, pid, col)[0].as_py() for pid in pids]
        assert got == [1, 2, 3, 2, 3], col
    store.engine.drop_table(name)
    db.tables.pop(name)
