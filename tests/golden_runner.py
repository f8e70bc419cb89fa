"""Runs the transcribed logictest cases (tests/golden/logictest_cases.py) against an engine and
formats rows the way logictest/runner.go:337-465 does."""
from __future__ import annotations

from typing import Callable, Dict, List

import pyarrow as pa

from frostdb_b200 import dynparquet as dp

SCHEMAS = {"default": dp.SampleDefinitionWithFloat, "bytes": dp.BytesDefinition}


def format_rows(batches: List[pa.RecordBatch]) -> List[List[str]]:
    rows: List[List[str]] = []
    for b in batches:
        cols = []
        for a in b.columns:
            if pa.types.is_dictionary(a.type):
                a = a.dictionary_decode()
            vals = a.to_pylist()
            if pa.types.is_floating(a.type):
                cols.append(["null" if v is None else "%f" % v for v in vals])
            elif pa.types.is_boolean(a.type):
                cols.append(["null" if v is None else ("true" if v else "false") for v in vals])
            else:
                cols.append(["null" if v is None else (v.decode() if isinstance(v, bytes) else str(v)) for v in vals])
        for i in range(b.num_rows):
            rows.append([c[i] for c in cols])
    return rows


def columns_of_insert(schema: dp.Schema, names: List[str], rows: List[List[str]]) -> Dict[str, list]:
    """stringToValue (runner.go:286-312): "null" -> NULL, else parsed by the column's type."""
    cols: Dict[str, list] = {}
    for j, n in enumerate(names):
        t = schema.definition_for(n).type
        vals = []
        for r in rows:
            s = r[j]
            if s == "null":
                vals.append(None)
            elif t == dp.TYPE_INT64:
                vals.append(int(s))
            elif t == dp.TYPE_DOUBLE:
                vals.append(float(s))
            else:
                vals.append(s)
        cols[n] = vals
    # columns of the schema that the insert does not list: non-nullable strings take "" (Go zero
    # value), ints 0 — the reference's insert always lists what it uses, so only example_type hits this
    n_rows = len(rows)
    for c in schema.columns:
        if c.dynamic or c.name in cols or c.nullable:
            continue
        cols[c.name] = [""] * n_rows if c.type == dp.TYPE_STRING else [0] * n_rows
    return cols


def run_case(case: dict, new_table: Callable, new_query: Callable, supports: Callable[[str], bool] = lambda kind: True):
    """new_table(schema) -> object with Insert(columns); new_query() -> query builder (ScanTable result).
    Yields (exec step, got rows, expected rows) for every exec the engine supports."""
    schema = SCHEMAS[case["schema"]]()
    table = new_table(schema)
    for step in case["steps"]:
        if step[0] == "insert":
            _, names, rows = step
            table.Insert(columns_of_insert(schema, names, rows))
            continue
        ex = step[1]
        if not supports(ex["kind"]):
            continue
        out: List[pa.RecordBatch] = []
        ex["build"](new_query()).Execute(None, lambda ctx, r: out.append(r))
        got = format_rows(out)
        exp = ex["expected"]
        yield ex, got, exp


def check(ex: dict, got: List[List[str]], exp: List[List[str]]) -> None:
    where = f"{ex['sql']!r} (line {ex['line']})"
    if ex["kind"].startswith("aggregate_limit_subset"):
        k = int(ex["kind"].split(":")[1])
        assert len(got) == k, where
        for r in got:
            assert r in exp, where
        return
    if ex["unordered"]:
        assert sorted(got) == sorted(exp), where
    else:
        assert got == exp, where
