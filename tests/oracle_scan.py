"""Test-side plumbing that runs the SAME logical plans through the CPU oracle: an OracleStore with the
Table.Insert API of frostdb_b200.store and an OracleScan that takes GPUScan's place in the physical
plan (so the host operators behind the aggregate are shared and only the scan differs)."""
from __future__ import annotations

from typing import Dict, List

from frostdb_b200 import _lib
from frostdb_b200 import dynparquet as dp
from frostdb_b200 import physicalplan as pp
from frostdb_b200 import query
from oracle import oracle as orc


class OracleScan(pp.GPUScan):
    def __init__(self, engine, *a, **kw):
        super().__init__(engine, *a, **kw)
        self.threads = getattr(engine, "threads", 1)

    def Execute(self, ctx, pool=None) -> None:
        plan, keep = self._plan()
        table: orc.OracleTable = self.engine.tables.setdefault(self.table_name, orc.OracleTable())
        try:
            res = table.execute(plan, threads=self.threads)
        except orc.OracleError as e:
            raise _lib.FrostGPUError(-1, str(e))
        try:
            batch = res.to_batch([a.Name() for a in self.agg_exprs])
            self.last_stats = {"rows_scanned": res.rows_scanned, "rows_selected": res.rows_selected, "groups": res.n_groups}
        finally:
            res.close()
        del keep
        self.next.Callback(ctx, self.renamed(batch))
        self.next.Finish(ctx)


class OracleEngine:
    """Stands where GPUEngine stands: holds the tables' parts."""

    def __init__(self, threads: int = 1):
        self.tables: Dict[str, orc.OracleTable] = {}
        self.threads = threads

    def put_parquet(self, table: str, buf: bytes, tx=None) -> None:
        self.tables.setdefault(table, orc.OracleTable()).add_parquet(buf, tx)

    def close(self):
        for t in self.tables.values():
            t.close()
        self.tables = {}


class OracleTableHandle:
    def __init__(self, engine: OracleEngine, name: str, schema: dp.Schema):
        self.engine, self.name, self.schema = engine, name, schema

    def Insert(self, columns, **opts):
        self.InsertParquet(dp.write_part(self.schema, columns, **opts))

    def InsertParquet(self, buf: bytes):
        self.engine.put_parquet(self.name, buf)


class OracleProvider:
    def __init__(self, engine: OracleEngine):
        self.engine = engine

    def gpu_engine(self):
        return self.engine


class OracleQueryBuilder(query.LocalQueryBuilder):
    def _with(self, b):
        return OracleQueryBuilder(self.engine, b)

    def buildPhysical(self):
        plan = self.planBuilder.Build()
        provider = plan.chain()[0].TableScan.TableProvider
        return pp.Build(provider.gpu_engine(), plan, scan_factory=OracleScan)


def oracle_query(engine: OracleEngine, table: str) -> OracleQueryBuilder:
    from frostdb_b200 import logicalplan as lp
    le = query.NewEngine(None, OracleProvider(engine))
    return OracleQueryBuilder(le, lp.Builder().Scan(le.table_provider, table))
