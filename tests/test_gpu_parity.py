"""Parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on identical seeded
inserts.  Integer results bit-exact; float64 sums within 1e-9 relative (north_star tolerance)."""
import numpy as np
import pyarrow as pa
import pytest

from frostdb_b200 import dynparquet as dp
from frostdb_b200 import logicalplan as lp
from frostdb_b200 import query
from tests.oracle_scan import OracleEngine, OracleTableHandle, oracle_query
from tests.util import make_columns, rows_of

pytestmark = pytest.mark.gpu
FLOAT_RTOL = 1e-9


class Pair:
    """The same parts in the GPU engine and in the oracle."""

    def __init__(self, store, name, schema):
        self.store, self.name, self.schema = store, name, schema
        self.db = store.DB(None, "parity")
        if name in self.db.tables:
            store.engine.drop_table(name)
            self.db.tables.pop(name)
        self.gt = self.db.Table(name, schema)
        self.oe = OracleEngine(threads=4)
        self.ot = OracleTableHandle(self.oe, name, schema)

    def insert(self, cols, **opts):
        buf = dp.write_part(self.schema, cols, **opts)
        self.gt.InsertParquet(buf)
        self.ot.InsertParquet(buf)

    def run(self, build):
        got, exp = [], []
        build(query.NewEngine(None, self.db.TableProvider()).ScanTable(self.name)).Execute(None, lambda c, r: got.append(r))
        build(oracle_query(self.oe, self.name)).Execute(None, lambda c, r: exp.append(r))
        return got, exp

    def close(self):
        self.oe.close()
        self.store.engine.drop_table(self.name)
        self.db.tables.pop(self.name, None)


def assert_same(got, exp, float_cols=()):
    names = sorted(set(n for b in exp for n in b.schema.names) | set(n for b in got for n in b.schema.names))
    exact = [n for n in names if n not in float_cols]
    g, e = rows_of(got, exact), rows_of(exp, exact)
    assert g == e
    if float_cols:
        keys = [n for n in exact if not n.startswith(("sum(", "min(", "max(", "count("))]
        gf = {r[:len(keys)]: r[len(keys):] for r in rows_of(got, keys + list(float_cols))}
        ef = {r[:len(keys)]: r[len(keys):] for r in rows_of(exp, keys + list(float_cols))}
        assert gf.keys() == ef.keys()
        for k in ef:
            for a, b in zip(gf[k], ef[k]):
                assert a == pytest.approx(b, rel=FLOAT_RTOL, abs=1e-12), k


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


@pytest.mark.parametrize("sort,page_version,page_size", [(True, "2.0", 8192), (False, "2.0", 1024), (True, "1.0", 65536)])
def test_aggregate_all_functions(pair, sort, page_version, page_size):
    p = pair("agg_all")
    for i in range(3):
        cols = make_columns(60_000, 11 + i, {"a": (7, 0.0), "b": (300, 0.1), "c": (3, 0.5)}, with_float=True, float_null_p=0.25,
                            t0=i * 60_000)
        p.insert(cols, sort=sort, row_group_size=25_000, data_page_size=page_size, data_page_version=page_version)
    got, exp = p.run(lambda q: q.Aggregate(
        [lp.Sum(lp.Col("value")), lp.Count(lp.Col("value")), lp.Min(lp.Col("value")), lp.Max(lp.Col("value")),
         lp.Sum(lp.Col("floatvalue")), lp.Min(lp.Col("floatvalue")), lp.Max(lp.Col("floatvalue")), lp.Count(lp.Col("floatvalue"))],
        [lp.Col("labels.a"), lp.Col("labels.b"), lp.Col("labels.c")]))
    assert_same(got, exp, float_cols=("sum(floatvalue)",))


def test_group_by_dynamic_columns_with_inconsistent_schemas(pair):
    """Parts with different sets of dynamic columns: absent == NULL, output is the union (aggregate.go:509-511,560-578)."""
    p = pair("dyn")
    p.insert(make_columns(20_000, 1, {"a": (5, 0.0), "b": (4, 0.2)}))
    p.insert(make_columns(20_000, 2, {"b": (4, 0.0), "c": (6, 0.3)}))
    p.insert(make_columns(20_000, 3, {"a": (5, 0.5), "c": (6, 0.0), "d": (2, 0.0)}))
    got, exp = p.run(lambda q: q.Aggregate([lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.DynCol("labels")]))
    assert_same(got, exp)
    got, exp = p.run(lambda q: q.Aggregate([lp.Max(lp.Col("value"))], [lp.Col("labels.a"), lp.Col("labels.d")]))
    assert_same(got, exp)


def test_high_cardinality_hash_table(pair):
    """Enough groups to leave the dense table: exact-key open addressing."""
    p = pair("hash")
    for i in range(2):
        p.insert(make_columns(150_000, 50 + i, {"a": (3000, 0.05), "b": (2500, 0.0), "c": (40, 0.1)}, with_float=False),
                 row_group_size=40_000)
    got, exp = p.run(lambda q: q.Aggregate([lp.Sum(lp.Col("value")), lp.Count(lp.Col("value")), lp.Min(lp.Col("timestamp"))],
                                           [lp.Col("labels.a"), lp.Col("labels.b"), lp.Col("labels.c")]))
    assert_same(got, exp)


def test_group_by_int64_column(pair):
    p = pair("intkey", dp.SampleDefinition())
    cols = make_columns(50_000, 5, {"a": (4, 0.0)}, value_mod=37)
    cols["value"] = cols["value"] + 1  # avoid the 0 == NULL merge quirk of int64 keys (hashed.go:254-262)
    p.insert(cols, row_group_size=20_000)
    got, exp = p.run(lambda q: q.Aggregate([lp.Count(lp.Col("timestamp")), lp.Sum(lp.Col("timestamp"))],
                                           [lp.Col("value"), lp.Col("labels.a")]))
    assert_same(got, exp)


@pytest.mark.parametrize("sel", [0.001, 0.1, 0.5, 0.9])
def test_filter_selectivity_then_aggregate(pair, sel):
    p = pair("fsel", dp.SampleDefinition())
    n = 200_000
    p.insert(make_columns(n, 8, {"a": (16, 0.1), "b": (100, 0.0)}, t0=0), row_group_size=64_000, data_page_size=32_768)
    f = lp.And(lp.Col("timestamp").Lt(lp.Literal(int(sel * n))), lp.Col("labels.a").NotEq(lp.Literal("v000003")))
    got, exp = p.run(lambda q: q.Filter(f).Aggregate([lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.Col("labels.b")]))
    assert_same(got, exp)


def test_filter_operator_matrix(pair):
    p = pair("fops")
    p.insert(make_columns(40_000, 21, {"a": (6, 0.3), "b": (9, 0.0)}, with_float=True, float_null_p=0.3, t0=0), row_group_size=15_000)
    leaves = [
        lp.Col("timestamp").Eq(lp.Literal(777)), lp.Col("timestamp").NotEq(lp.Literal(777)), lp.Col("timestamp").LtEq(lp.Literal(20_000)),
        lp.Col("value").Gt(lp.Literal(500)), lp.Col("floatvalue").Lt(lp.Literal(250.5)), lp.Col("floatvalue").GtEq(lp.Literal(100)),
        lp.Col("value").Lt(lp.Literal(99.5)),
        lp.Col("labels.a").Eq(lp.Literal("v000002")), lp.Col("labels.a").NotEq(lp.Literal("v000002")),
        lp.Col("labels.a").Eq(lp.Literal(None)), lp.Col("labels.a").NotEq(lp.Literal(None)),
        lp.Col("labels.a").Contains("0004"), lp.Col("labels.a").ContainsNot("0004"),
        lp.Col("labels.a").RegexMatch("v00000[1-3]$"), lp.Col("labels.a").RegexNotMatch("v00000[1-3]$"),
        lp.Col("labels.zz").Eq(lp.Literal("")), lp.Col("labels.zz").Eq(lp.Literal("x")), lp.Col("labels.zz").NotEq(lp.Literal("x")),
        lp.Col("labels.zz").NotEq(lp.Literal(None)), lp.Col("labels.zz").Eq(lp.Literal(None)), lp.Col("labels.zz").RegexMatch(""),
        lp.Col("labels.zz").RegexNotMatch("foo"), lp.Col("labels.zz").RegexMatch("foo"), lp.Col("nope").Lt(lp.Literal(4)),
        lp.Or(lp.And(lp.Col("labels.a").Eq(lp.Literal("v000001")), lp.Col("value").Lt(lp.Literal(300))),
              lp.And(lp.Col("labels.b").Eq(lp.Literal("v000007")), lp.Or(lp.Col("timestamp").Gt(lp.Literal(30_000)), lp.Col("labels.a").Eq(lp.Literal(None))))),
    ]
    for f in leaves:
        got, exp = p.run(lambda q: q.Filter(f).Aggregate([lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.Col("labels.b")]))
        try:
            assert_same(got, exp)
        except AssertionError as e:
            raise AssertionError(f"filter {f.Name()}: {e}") from e


def test_arithmetic_aggregate_expression(pair):
    p = pair("arith", dp.SampleDefinition())
    p.insert(make_columns(30_000, 31, {"a": (5, 0.0)}, t0=-15_000), row_group_size=10_000)
    got, exp = p.run(lambda q: q.Aggregate(
        [lp.Sum(lp.Mul(lp.Col("value"), lp.Col("timestamp"))), lp.Max(lp.Sub(lp.Col("timestamp"), lp.Mul(lp.Col("value"), lp.Literal(3)))),
         lp.Sum(lp.Div(lp.Col("timestamp"), lp.Col("value")))], [lp.Col("labels.a")]))
    assert_same(got, exp)


def test_distinct(pair):
    p = pair("dist", dp.SampleDefinition())
    p.insert(make_columns(30_000, 41, {"a": (5, 0.2), "b": (3, 0.0)}))
    p.insert(make_columns(30_000, 42, {"a": (5, 0.0), "c": (4, 0.5)}))
    for exprs in ([lp.Col("labels.a")], [lp.Col("labels.a"), lp.Col("labels.b")], [lp.DynCol("labels")]):
        got, exp = p.run(lambda q: q.Distinct(*exprs))
        assert_same(got, exp)
    got, exp = p.run(lambda q: q.Filter(lp.Col("labels.b").Eq(lp.Literal("v000001"))).Distinct(lp.Col("labels.a")))
    assert_same(got, exp)


def test_empty_and_ragged_inputs(pair):
    p = pair("ragged", dp.SampleDefinition())
    # no parts at all
    got, exp = p.run(lambda q: q.Aggregate([lp.Sum(lp.Col("value"))], [lp.Col("labels.a")]))
    assert rows_of(got) == rows_of(exp) == []
    p.insert(make_columns(1, 1, {"a": (2, 0.0)}))            # single row
    p.insert(make_columns(2047, 2, {"a": (2, 0.0)}))         # one row short of a tile
    p.insert(make_columns(2049, 3, {"a": (2, 1.0)}))         # one row past a tile, all-NULL label
    p.insert(make_columns(4096, 4, {"a": (2, 0.0)}), row_group_size=2048)
    got, exp = p.run(lambda q: q.Aggregate([lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.Col("labels.a")]))
    assert_same(got, exp)
    # filter that selects nothing
    got, exp = p.run(lambda q: q.Filter(lp.Col("timestamp").Lt(lp.Literal(-1))).Aggregate([lp.Sum(lp.Col("value"))], [lp.Col("labels.a")]))
    assert rows_of(got) == rows_of(exp) == []


def test_missing_column_leaves_follow_the_row_group_filter(pair):
    """Parquet parts pass LSM.Scan's row-group filter first, whose rules for a column that is missing from the row
    group (query/expr/binaryscalarexpr.go:47-73) differ from the physical plan's (physicalplan/binaryscalarexpr.go:47-73):
    `missing == 5`, `missing != 5`, `missing != ""` drop the row group."""
    p = pair("missing_rules")
    n = 20_000
    p.insert(make_columns(n, 1, {"a": (5, 0.0), "b": (4, 0.2)}, t0=0))
    p.insert(make_columns(n, 2, {"a": (5, 0.1)}, t0=n))              # labels.b missing in this part
    m, ts = lp.Col("labels.b"), lp.Col("timestamp")
    zz = lp.Col("labels.zz")                                           # missing everywhere
    filters = [m.NotEq(lp.Literal("")), m.Eq(lp.Literal("")), zz.Eq(lp.Literal(5)), zz.NotEq(lp.Literal(5)),
               zz.NotEq(lp.Literal("")), zz.Eq(lp.Literal("")), zz.Eq(lp.Literal(None)), zz.NotEq(lp.Literal(None)), zz.Lt(lp.Literal(3)),
               m.NotEq(lp.Literal("v000001")), m.Eq(lp.Literal(None)), m.NotEq(lp.Literal(None)),
               lp.Or(zz.Eq(lp.Literal(5)), ts.Lt(lp.Literal(100))), lp.Or(m.NotEq(lp.Literal("")), ts.GtEq(lp.Literal(2 * n - 10))),
               lp.And(zz.NotEq(lp.Literal("x")), ts.Lt(lp.Literal(n + 50)))]
    for f in filters:
        try:
            got, exp = p.run(lambda q: q.Filter(f).Aggregate([lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.Col("labels.a")]))
            assert_same(got, exp)
        except AssertionError as e:
            raise AssertionError(f"filter {f.Name()}: {e}") from e


def test_key_column_only_in_pruned_parts(store):
    """A dynamic key column that exists only in row groups the filter rules out (lazily built parts): no dictionary was
    ever built for it; the query must neither fail nor name the column."""
    import numpy as np
    eng = store.engine
    name = "pruned_keys"
    eng.drop_table(name)
    n = 10_000
    bufs = [np.frombuffer(dp.write_part(dp.SampleDefinition(), make_columns(n, 7, {"a": (3, 0.0)}, t0=0)), dtype=np.uint8),
            np.frombuffer(dp.write_part(dp.SampleDefinition(), make_columns(n, 8, {"a": (3, 0.0), "x": (4, 0.0)}, t0=10 * n)), dtype=np.uint8)]
    try:
        for b in bufs:
            eng.put_parquet(name, b, borrow=True)
        got = []

        class _P:
            def gpu_engine(self):
                return eng
        q = query.NewEngine(None, _P()).ScanTable(name)
        q.Filter(lp.Col("timestamp").Lt(lp.Literal(n))).Aggregate([lp.Count(lp.Col("value"))], [lp.DynCol("labels")]).Execute(None, lambda c, r: got.append(r))
        rows = rows_of(got)
        assert sum(r[-1] for r in rows) == n
    finally:
        eng.drop_table(name)


def test_computed_group_keys_time_buckets(pair):
    """The Parca "Range" query shape (bench_test.go:325-349; logictest exec/aggregate/window): group by a bucket
    computed from the timestamp, alone and next to a dictionary key, with a filter in front."""
    p = pair("buckets", dp.SampleDefinition())
    n = 60_000
    for i in range(3):
        p.insert(make_columns(n, 400 + i, {"a": (6, 0.1), "b": (30, 0.0)}, t0=1_000_000 + i * n), sort=(i != 1), row_group_size=25_000)
    ts, v = lp.Col("timestamp"), lp.Col("value")

    def bucket(k):
        return lp.Mul(lp.Div(ts, lp.Literal(k)), lp.Literal(k)).Alias("timestamp_bucket")
    aggs = [lp.Sum(v), lp.Count(v), lp.Max(v)]
    for k in (1000, 7, 1_000_000_000):
        got, exp = p.run(lambda q: q.Project(v, bucket(k)).Aggregate(aggs, [lp.Col("timestamp_bucket")]))
        assert_same(got, exp)
    got, exp = p.run(lambda q: q.Filter(ts.GtEq(lp.Literal(1_030_000))).Project(v, lp.Col("labels.a"), bucket(5000))
                     .Aggregate(aggs, [lp.Col("labels.a"), lp.Col("timestamp_bucket")]))
    assert_same(got, exp)
    got, exp = p.run(lambda q: q.Project(v, lp.Sub(ts, v).Alias("d")).Aggregate([lp.Count(v)], [lp.Col("d"), lp.Col("labels.b")]))
    assert_same(got, exp)


def test_read_transaction_hides_later_parts(pair):
    """Table.View -> LSM.Scan skips parts whose tx is above the read transaction (index/lsm.go:416): a query with an
    older watermark must not see them, in the engine and in the oracle alike; the compiled plan is per watermark."""
    import ctypes as C
    from frostdb_b200 import _lib
    from frostdb_b200.physicalplan import GPUScan
    p = pair("watermark", dp.SampleDefinition())
    bufs = [dp.write_part(p.schema, make_columns(10_000, 600 + i, {"a": (4, 0.0)}, t0=i * 10_000)) for i in range(3)]
    eng = p.store.engine
    for i, b in enumerate(bufs):
        eng.put_parquet(p.name, b, tx=10 * (i + 1))
        p.oe.put_parquet(p.name, b, tx=10 * (i + 1))
    aggs, keys = [lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.Col("labels.a")]
    lib = _lib.load()
    scan = GPUScan(eng, p.name, None, _lib.PLAN_AGGREGATE, keys, aggs)
    q, keep = scan.prepare()
    plan, keep2 = scan._plan()
    names = ["labels.a", "sum(value)", "count(value)"]
    for tx, n_parts in ((5, 0), (10, 1), (25, 2), (30, 3), (10, 1), (1000, 3)):
        res = C.c_void_p()
        _lib.check(lib.fgpu_query_execute(eng.handle, q, tx, C.byref(res)))
        got = list(eng.drain(res))
        lib.fgpu_result_free(res)
        ores = p.oe.tables[p.name].execute(plan, tx=tx, threads=2)
        exp = [ores.to_batch([a.Name() for a in aggs])] if ores.n_groups else []
        ores.close()
        assert rows_of(got, names) == rows_of(exp, names), tx
        assert sum(r[2] for r in rows_of(got, names)) == 10_000 * n_parts
    lib.fgpu_query_free(q)


def test_concurrent_executes_on_one_context(store):
    """Several host threads on one fgpu_ctx (the Go shim calls from many goroutines): calls are serialised inside
    the library, every thread gets its own complete result."""
    import threading
    p = Pair(store, "concurrent", dp.SampleDefinition())
    try:
        for i in range(2):
            p.insert(make_columns(30_000, 650 + i, {"a": (5, 0.0), "b": (7, 0.1)}, t0=i * 30_000), sort=bool(i))
        plans = [lambda q: q.Aggregate([lp.Sum(lp.Col("value"))], [lp.Col("labels.a")]),
                 lambda q: q.Filter(lp.Col("value").Lt(lp.Literal(500))).Aggregate([lp.Count(lp.Col("value"))], [lp.Col("labels.b")]),
                 lambda q: q.Distinct(lp.Col("labels.a"), lp.Col("labels.b"))]
        expected = []
        for b in plans:
            exp = []
            b(oracle_query(p.oe, p.name)).Execute(None, lambda c, r: exp.append(r))
            expected.append(rows_of(exp))
        errors = []

        def worker(k):
            try:
                for it in range(20):
                    j = (k + it) % len(plans)
                    got = []
                    plans[j](query.NewEngine(None, p.db.TableProvider()).ScanTable(p.name)).Execute(None, lambda c, r: got.append(r))
                    if rows_of(got) != expected[j]:
                        errors.append((k, it, j))
            except Exception as ex:  # noqa: BLE001
                errors.append((k, repr(ex)))
        ths = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        assert not errors, errors[:3]
    finally:
        p.close()
