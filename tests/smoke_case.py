"""__graft_entry__.smoke(): one small scan -> filter -> group-by on cuda:0 through the C-ABI, checked
against the CPU oracle."""
from frostdb_b200 import dynparquet as dp
from frostdb_b200 import logicalplan as lp
from frostdb_b200 import query
from frostdb_b200.store import ColumnStore
from tests.oracle_scan import OracleEngine, OracleTableHandle, oracle_query
from tests.util import make_columns, rows_of


def run() -> None:
    cs = ColumnStore(0)
    oe = OracleEngine(threads=2)
    try:
        db = cs.DB(None, "smoke")
        schema = dp.SampleDefinition()
        gt = db.Table("smoke", schema)
        ot = OracleTableHandle(oe, "smoke", schema)
        for p in range(2):
            cols = make_columns(30_000, 900 + p, {"job": (64, 0.0), "pod": (200, 0.1)}, t0=p * 30_000)
            buf = dp.write_part(schema, cols, row_group_size=10_000, data_page_size=16_384)
            gt.InsertParquet(buf)
            ot.InsertParquet(buf)
        f = lp.And(lp.Col("timestamp").Gt(lp.Literal(15_000)), lp.Col("timestamp").Lt(lp.Literal(45_000)))

        def q(b):
            return b.Filter(f).Aggregate([lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))], [lp.Col("labels.job")])
        got, exp = [], []
        q(query.NewEngine(None, db.TableProvider()).ScanTable("smoke")).Execute(None, lambda c, r: got.append(r))
        q(oracle_query(oe, "smoke")).Execute(None, lambda c, r: exp.append(r))
        names = ["labels.job", "sum(value)", "count(value)"]
        assert rows_of(got, names) == rows_of(exp, names), "GPU result differs from the oracle"
        assert len(rows_of(got, names)) == 64
    finally:
        oe.close()
        cs.Close()
