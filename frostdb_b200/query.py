"""Mirror of `query.NewEngine` / `query.Builder` (query/engine.go:16-25,48-64,73-80,158-196) on top
of the GPU physical plan."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

from . import logicalplan as lp
from . import physicalplan as pp


class LocalEngine:
    def __init__(self, pool, table_provider, **options):
        self.pool = pool
        self.table_provider = table_provider
        self.options = options

    def ScanTable(self, name: str) -> "LocalQueryBuilder":  # engine.go:73-80
        return LocalQueryBuilder(self, lp.Builder().Scan(self.table_provider, name))


def NewEngine(pool, table_provider, **options) -> LocalEngine:  # engine.go:48-64
    return LocalEngine(pool, table_provider, **options)


class LocalQueryBuilder:
    def __init__(self, engine: LocalEngine, builder: lp.Builder):
        self.engine = engine
        self.planBuilder = builder
        self.last_plan: Optional[pp.OutputPlan] = None

    def _with(self, b: lp.Builder) -> "LocalQueryBuilder":
        return LocalQueryBuilder(self.engine, b)

    def Aggregate(self, agg_exprs: Sequence[lp.AggregationFunction], group_exprs: Sequence[lp.Expr]):
        return self._with(self.planBuilder.Aggregate(agg_exprs, group_exprs))

    def Filter(self, expr: Optional[lp.Expr]):
        return self._with(self.planBuilder.Filter(expr))

    def Distinct(self, *exprs: lp.Expr):
        return self._with(self.planBuilder.Distinct(*exprs))

    def Project(self, *exprs: lp.Expr):
        return self._with(self.planBuilder.Project(*exprs))

    def Limit(self, expr: Optional[lp.Expr]):
        return self._with(self.planBuilder.Limit(expr))

    def buildPhysical(self) -> pp.OutputPlan:  # engine.go:178-196
        plan = self.planBuilder.Build()
        provider = plan.chain()[0].TableScan.TableProvider
        return pp.Build(provider.gpu_engine(), plan)

    def Execute(self, ctx, callback: Callable) -> None:  # engine.go:158-168
        phy = self.buildPhysical()
        phy.SetNextCallback(callback)
        self.last_plan = phy
        phy.Execute(ctx, self.engine.pool)

    def Explain(self, ctx=None) -> str:  # engine.go:170-176
        return self.buildPhysical().DrawString()
