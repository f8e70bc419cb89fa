"""The GPU engine against the reference's golden vectors, through the same host operators as the
oracle run (tests/test_oracle_golden.py)."""
import pytest

from frostdb_b200 import _lib, query
from tests import golden_runner as gr
from tests.golden.logictest_cases import CASES

pytestmark = pytest.mark.gpu

_SUPPORTED = {"aggregate", "distinct", "aggregate_limit_subset:1", "filter"}


@pytest.mark.parametrize("case", CASES, ids=[c["source"].split("/exec/")[1] for c in CASES])
def test_gpu_reproduces_reference_goldens(store, case):
    db = store.DB(None, "golden")
    name = "g_" + case["source"].split("/")[-1]
    if name in db.tables:
        store.engine.drop_table(name)
        db.tables.pop(name)

    def new_table(schema):
        return db.Table(name, schema)

    def new_query():
        return query.NewEngine(None, db.TableProvider()).ScanTable(name)

    ran = 0
    for step_kind in ("run",):
        try:
            for ex, got, exp in gr.run_case(case, new_table, new_query, supports=lambda k: k in _SUPPORTED):
                gr.check(ex, got, exp)
                ran += 1
        except _lib.FrostGPUError as e:
            if e.code == _lib.FGPU_ERR_UNSUPPORTED and "not implemented" in e.msg:
                pytest.skip(e.msg)
            raise
    assert ran > 0
