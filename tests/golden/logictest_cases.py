"""Golden vectors transcribed from the reference's own end-to-end tests.

Source: /root/reference/logictest/testdata/exec/{aggregate,filter,distinct}/* (cockroach
`datadriven` files run by logictest/logic_test.go:158-179 through logictest/runner.go).  Each
`createtable` / `insert` / `exec` block is restated here as data; the SQL text is kept for
traceability and the plan is written with the logicalplan builder exactly the way
sqlparse/visitor.go:44-140 translates that SQL (Filter -> Project(pre) -> Aggregate -> Project(post)
-> Limit; Distinct -> Builder.Distinct; plain selects -> Filter -> Project).  Expected rows are the
reference's expected output verbatim (tab-separated cells, floats printed with %f, NULL printed
"null", runner.go:406-465).

One insert block = one InsertRecord = one part (runner.go:256-279).
"""
from frostdb_b200.logicalplan import (And, Avg, Col, Count, Div, DynCol, Literal, Max, Min, Mul, Or, Sub, Sum)

# schema names of logictest/logic_test.go:41-157
DEFAULT = "default"  # dynparquet.SampleDefinitionWithFloat()


def _c(*names):
    return [Col(n) for n in names]


CASES = []


def case(source, schema=DEFAULT):
    c = {"source": source, "schema": schema, "steps": []}
    CASES.append(c)
    return c


def insert(c, cols, rows):
    c["steps"].append(("insert", cols, [r.split() for r in rows]))


def exec_(c, line, sql, build, expected, unordered=False, kind="aggregate"):
    c["steps"].append(("exec", {"line": line, "sql": sql, "build": build, "expected": [r.split() for r in expected],
                                "unordered": unordered, "kind": kind}))


# ---------------------------------------------------------------------------------------------------
# exec/aggregate/aggregate
# ---------------------------------------------------------------------------------------------------
_agg = case("logictest/testdata/exec/aggregate/aggregate")
_cols8 = ["labels.label1", "labels.label2", "labels.label3", "labels.label4", "stacktrace", "timestamp", "value", "floatvalue"]
insert(_agg, _cols8, ["value1 value2 null null stack1 1 1 1.1",
                      "value2 value2 value3 null stack1 2 2 2.2",
                      "value3 value2 null value4 stack1 3 3 3.3"])
insert(_agg, _cols8, ["value4 value2 null null stack1 4 4 4.4",
                      "value5 value2 value3 null stack1 5 5 5.5",
                      "value6 value2 null value4 stack1 6 6 6.6"])

exec_(_agg, 16, "select sum(value) as value_sum, labels.label2 group by labels.label2",
      lambda q: q.Project(*_c("value", "labels.label2")).Aggregate([Sum(Col("value"))], _c("labels.label2"))
      .Project(Sum(Col("value")).Alias("value_sum"), Col("labels.label2")),
      ["21 value2"])
exec_(_agg, 21, "select labels.label2, sum(floatvalue) as float_value_sum group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "floatvalue")).Aggregate([Sum(Col("floatvalue"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Sum(Col("floatvalue")).Alias("float_value_sum")),
      ["value2 23.100000"])
exec_(_agg, 26, "select labels.label2, max(value) as value_max group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value")).Aggregate([Max(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Max(Col("value")).Alias("value_max")),
      ["value2 6"])
exec_(_agg, 31, "select labels.label2, max(floatvalue) as value_max group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "floatvalue")).Aggregate([Max(Col("floatvalue"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Max(Col("floatvalue")).Alias("value_max")),
      ["value2 6.600000"])
exec_(_agg, 36, "select labels.label2, min(floatvalue) as value_min group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "floatvalue")).Aggregate([Min(Col("floatvalue"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Min(Col("floatvalue")).Alias("value_min")),
      ["value2 1.100000"])
exec_(_agg, 41, "select labels.label2, count(value) as value_count group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value")).Aggregate([Count(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Count(Col("value")).Alias("value_count")),
      ["value2 6"])
exec_(_agg, 46, "select labels.label2, avg(value) group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value")).Aggregate([Avg(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Avg(Col("value"))),
      ["value2 3"])
exec_(_agg, 51, "select labels.label2, avg(value) as value_avg group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value")).Aggregate([Avg(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Avg(Col("value")).Alias("value_avg")),
      ["value2 3"])
exec_(_agg, 56, "select labels.label2, avg(floatvalue) as value_avg group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "floatvalue")).Aggregate([Avg(Col("floatvalue"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Avg(Col("floatvalue")).Alias("value_avg")),
      ["value2 3.850000"])
exec_(_agg, 61, "select labels.label4, avg(value) as value_avg group by labels.label4",
      lambda q: q.Project(*_c("labels.label4", "value")).Aggregate([Avg(Col("value"))], _c("labels.label4"))
      .Project(Col("labels.label4"), Avg(Col("value")).Alias("value_avg")),
      ["null 3", "value4 4"])
exec_(_agg, 67, "select stacktrace, sum(value), count(value) group by stacktrace",
      lambda q: q.Project(*_c("stacktrace", "value")).Aggregate([Sum(Col("value")), Count(Col("value"))], _c("stacktrace"))
      .Project(Col("stacktrace"), Sum(Col("value")), Count(Col("value"))),
      ["stack1 21 6"])
exec_(_agg, 72, "select stacktrace, sum(floatvalue), count(floatvalue) group by stacktrace",
      lambda q: q.Project(*_c("stacktrace", "floatvalue")).Aggregate([Sum(Col("floatvalue")), Count(Col("floatvalue"))], _c("stacktrace"))
      .Project(Col("stacktrace"), Sum(Col("floatvalue")), Count(Col("floatvalue"))),
      ["stack1 23.100000 6"])
exec_(_agg, 77, "select stacktrace, sum(value) as value_sum, count(value) as value_count group by stacktrace",
      lambda q: q.Project(*_c("stacktrace", "value")).Aggregate([Sum(Col("value")), Count(Col("value"))], _c("stacktrace"))
      .Project(Col("stacktrace"), Sum(Col("value")).Alias("value_sum"), Count(Col("value")).Alias("value_count")),
      ["stack1 21 6"])
exec_(_agg, 82, "select labels.label2, sum(value), count(value), min(value), max(value) group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value"))
      .Aggregate([Sum(Col("value")), Count(Col("value")), Min(Col("value")), Max(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Sum(Col("value")), Count(Col("value")), Min(Col("value")), Max(Col("value"))),
      ["value2 21 6 1 6"])
exec_(_agg, 87, "select labels.label2, sum(floatvalue), count(floatvalue), min(floatvalue), max(floatvalue) group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "floatvalue"))
      .Aggregate([Sum(Col("floatvalue")), Count(Col("floatvalue")), Min(Col("floatvalue")), Max(Col("floatvalue"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Sum(Col("floatvalue")), Count(Col("floatvalue")), Min(Col("floatvalue")), Max(Col("floatvalue"))),
      ["value2 23.100000 6 1.100000 6.600000"])
exec_(_agg, 92, "select labels.label1, sum(value) as value_sum where timestamp >= 1 group by labels.label1",
      lambda q: q.Filter(Col("timestamp").GtEq(Literal(1))).Project(*_c("labels.label1", "value"))
      .Aggregate([Sum(Col("value"))], _c("labels.label1"))
      .Project(Col("labels.label1"), Sum(Col("value")).Alias("value_sum")),
      ["value1 1", "value2 2", "value3 3", "value4 4", "value5 5", "value6 6"], unordered=True)
exec_(_agg, 102, "select labels, sum(value) as value_sum group by labels",
      lambda q: q.Project(DynCol("labels"), Col("value")).Aggregate([Sum(Col("value"))], [DynCol("labels")])
      .Project(DynCol("labels"), Sum(Col("value")).Alias("value_sum")),
      ["value1 value2 null null 1", "value2 value2 value3 null 2", "value3 value2 null value4 3",
       "value4 value2 null null 4", "value5 value2 value3 null 5", "value6 value2 null value4 6"], unordered=True)
# limits: the reference relies on its ordered aggregation for WHICH rows survive a limit; only the
# row count and membership are engine independent, so the expectation here is the full set + a limit.
exec_(_agg, 113, "select sum(value) as value_sum, labels.label3 group by labels.label3 limit 3",
      lambda q: q.Project(*_c("value", "labels.label3")).Aggregate([Sum(Col("value"))], _c("labels.label3"))
      .Project(Sum(Col("value")).Alias("value_sum"), Col("labels.label3")).Limit(Literal(3)),
      ["14 null", "7 value3"], unordered=True)
exec_(_agg, 120, "select sum(value) as value_sum, labels.label3 group by labels.label3 limit 1",
      lambda q: q.Project(*_c("value", "labels.label3")).Aggregate([Sum(Col("value"))], _c("labels.label3"))
      .Project(Sum(Col("value")).Alias("value_sum"), Col("labels.label3")).Limit(Literal(1)),
      ["14 null", "7 value3"], unordered=True, kind="aggregate_limit_subset:1")

# ---------------------------------------------------------------------------------------------------
# exec/aggregate/aggregate_nulls
# ---------------------------------------------------------------------------------------------------
_nulls = case("logictest/testdata/exec/aggregate/aggregate_nulls")
insert(_nulls, ["labels.label1", "labels.label2", "stacktrace", "timestamp", "value"],
       ["value1 null stack1 1 1", "null value2 stack1 2 2", "null value2 stack1 3 3"])
exec_(_nulls, 12, "select labels.label2, sum(value) as value_sum group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value")).Aggregate([Sum(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Sum(Col("value")).Alias("value_sum")),
      ["value2 5", "null 1"], unordered=True)
exec_(_nulls, 18, "select labels.label2, max(value) as value_max group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value")).Aggregate([Max(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Max(Col("value")).Alias("value_max")),
      ["value2 3", "null 1"], unordered=True)
exec_(_nulls, 24, "select labels.label2, count(value) as value_count group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value")).Aggregate([Count(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Count(Col("value")).Alias("value_count")),
      ["value2 2", "null 1"], unordered=True)
exec_(_nulls, 30, "select labels.label2, sum(value) as value_sum, count(value) as value_count group by labels.label2",
      lambda q: q.Project(*_c("labels.label2", "value")).Aggregate([Sum(Col("value")), Count(Col("value"))], _c("labels.label2"))
      .Project(Col("labels.label2"), Sum(Col("value")).Alias("value_sum"), Count(Col("value")).Alias("value_count")),
      ["value2 5 2", "null 1 1"], unordered=True)

# ---------------------------------------------------------------------------------------------------
# exec/aggregate/math (the group-by blocks, lines 98-130) and exec/projection/math_projection:18-22
# ---------------------------------------------------------------------------------------------------
_math = case("logictest/testdata/exec/aggregate/math")
insert(_math, ["labels.label1", "timestamp", "value"], ["value1 1 2", "value1 3 4", "value1 5 6", "value1 11 0"])
exec_(_math, 98, "select labels.label1, sum(value), count(value) group by labels.label1",
      lambda q: q.Project(*_c("labels.label1", "value")).Aggregate([Sum(Col("value")), Count(Col("value"))], _c("labels.label1"))
      .Project(Col("labels.label1"), Sum(Col("value")), Count(Col("value"))),
      ["value1 12 4"])
exec_(_math, 103, "select max(value) - min(value), sum(value) / count(value) group by labels.label1",
      lambda q: q.Project(Col("value"), Col("labels.label1"))
      .Aggregate([Max(Col("value")), Min(Col("value")), Sum(Col("value")), Count(Col("value"))], _c("labels.label1"))
      .Project(Sub(Max(Col("value")), Min(Col("value"))), Div(Sum(Col("value")), Count(Col("value")))),
      ["6 3"])
exec_(_math, 108, "select labels.label1, max(value) - min(value), sum(value) / count(value) group by labels.label1",
      lambda q: q.Project(Col("labels.label1"), Col("value"))
      .Aggregate([Max(Col("value")), Min(Col("value")), Sum(Col("value")), Count(Col("value"))], _c("labels.label1"))
      .Project(Col("labels.label1"), Sub(Max(Col("value")), Min(Col("value"))), Div(Sum(Col("value")), Count(Col("value")))),
      ["value1 6 3"])
exec_(_math, 113, "select labels.label1, (max(value) - min(value)) /2, sum(value) / count(value) group by labels.label1",
      lambda q: q.Project(Col("labels.label1"), Col("value"))
      .Aggregate([Max(Col("value")), Min(Col("value")), Sum(Col("value")), Count(Col("value"))], _c("labels.label1"))
      .Project(Col("labels.label1"), Div(Sub(Max(Col("value")), Min(Col("value"))), Literal(2)),
               Div(Sum(Col("value")), Count(Col("value")))),
      ["value1 3 3"])
exec_(_math, 118, "select labels.label1, (max(value) - min(value)) /2, sum(value) / count(value), (sum(value) / count(value)) * 2 group by labels.label1",
      lambda q: q.Project(Col("labels.label1"), Col("value"))
      .Aggregate([Max(Col("value")), Min(Col("value")), Sum(Col("value")), Count(Col("value"))], _c("labels.label1"))
      .Project(Col("labels.label1"), Div(Sub(Max(Col("value")), Min(Col("value"))), Literal(2)),
               Div(Sum(Col("value")), Count(Col("value"))), Mul(Div(Sum(Col("value")), Count(Col("value"))), Literal(2))),
      ["value1 3 3 6"])

_mproj = case("logictest/testdata/exec/projection/math_projection")
insert(_mproj, ["labels.label1", "stacktrace", "timestamp", "value"], ["value1 stack1 1 2", "value1 stack1 3 4", "value1 stack2 5 6"])
exec_(_mproj, 17, "select stacktrace, sum(value * timestamp) group by stacktrace",
      lambda q: q.Project(Col("stacktrace"), Mul(Col("value"), Col("timestamp")))
      .Aggregate([Sum(Mul(Col("value"), Col("timestamp")))], _c("stacktrace"))
      .Project(Col("stacktrace"), Sum(Mul(Col("value"), Col("timestamp")))),
      ["stack1 14", "stack2 30"], unordered=True)

# ---------------------------------------------------------------------------------------------------
# exec/filter/filter — plain selects: Filter -> Project (rows, order preserved)
# ---------------------------------------------------------------------------------------------------
_filter = case("logictest/testdata/exec/filter/filter")
insert(_filter, ["labels.label1", "labels.label2", "labels.label3", "labels.label4", "stacktrace", "timestamp", "value"],
       ["value1 value2 null null stack1 1 1", "value2 value2 value3 null stack1 2 2", "value3 value2 null value4 stack1 3 3"])
_R1 = "value1 value2 null null stack1 1 1"
_R2 = "value2 value2 value3 null stack1 2 2"
_R3 = "value3 value2 null value4 stack1 3 3"
_ALL = [DynCol("labels"), Col("stacktrace"), Col("timestamp"), Col("value")]


def _sel(line, where_sql, expr, expected, limit=None):
    def build(q, expr=expr, limit=limit):
        q = q.Filter(expr).Project(*_ALL)
        return q.Limit(Literal(limit)) if limit is not None else q
    exec_(_filter, line, "select labels, stacktrace, timestamp, value where " + where_sql, build, expected, kind="filter")


_sel(11, "timestamp = 2", Col("timestamp").Eq(Literal(2)), [_R2])
_sel(16, "timestamp != 2", Col("timestamp").NotEq(Literal(2)), [_R1, _R3])
_sel(22, "timestamp < 2", Col("timestamp").Lt(Literal(2)), [_R1])
_sel(27, "timestamp <= 2", Col("timestamp").LtEq(Literal(2)), [_R1, _R2])
_sel(33, "timestamp <= 2 limit 1", Col("timestamp").LtEq(Literal(2)), [_R1], limit=1)
_sel(38, "timestamp > 2", Col("timestamp").Gt(Literal(2)), [_R3])
_sel(43, "timestamp >= 2", Col("timestamp").GtEq(Literal(2)), [_R2, _R3])
_sel(49, "labels.label4 = 'value4'", Col("labels.label4").Eq(Literal("value4")), [_R3])
_sel(54, "labels.label1 = 'value1' or labels.label2 = 'value2'",
     Or(Col("labels.label1").Eq(Literal("value1")), Col("labels.label2").Eq(Literal("value2"))), [_R1, _R2, _R3])
_sel(61, "labels.label5 != 'value4'", Col("labels.label5").NotEq(Literal("value4")), [_R1, _R2, _R3])
_sel(68, "labels.label5 = ''", Col("labels.label5").Eq(Literal("")), [_R1, _R2, _R3])
_sel(75, "labels.label1 regexp 'value.' and labels.label2 = 'value2'",
     And(Col("labels.label1").RegexMatch("value."), Col("labels.label2").Eq(Literal("value2"))), [_R1, _R2, _R3])
_sel(82, "labels.label5 regexp ''", Col("labels.label5").RegexMatch(""), [_R1, _R2, _R3])
_sel(89, "labels.label5 not regexp 'foo'", Col("labels.label5").RegexNotMatch("foo"), [_R1, _R2, _R3])
_sel(96, "labels.label3 regexp 'value.' and labels.label5 regexp '' and labels.label2 = 'value2'",
     And(Col("labels.label3").RegexMatch("value."), Col("labels.label5").RegexMatch(""), Col("labels.label2").Eq(Literal("value2"))), [_R2])
_sel(101, "labels.label1 regexp 'value.' and labels.label2 = 'value2' and labels.label1 != 'value3'",
     And(Col("labels.label1").RegexMatch("value."), Col("labels.label2").Eq(Literal("value2")), Col("labels.label1").NotEq(Literal("value3"))),
     [_R1, _R2])
_sel(108, "labels.label1 regexp 'value.'", Col("labels.label1").RegexMatch("value."), [_R1, _R2, _R3])
_sel(116, "labels.label1 regexp 'values.'", Col("labels.label1").RegexMatch("values."), [])
_sel(120, "labels.label1 regexp 'value.' and labels.label5 = ''",
     And(Col("labels.label1").RegexMatch("value."), Col("labels.label5").Eq(Literal(""))), [_R1, _R2, _R3])
_sel(127, "labels.label3 regexp 'value.' and (labels.label1 = 'value1' or labels.label1 = 'value2')",
     And(Col("labels.label3").RegexMatch("value."), Or(Col("labels.label1").Eq(Literal("value1")), Col("labels.label1").Eq(Literal("value2")))),
     [_R2])
_sel(132, "labels.label4 = 'value4' or (labels.label2 regexp 'value.' and labels.label1 = 'value2')",
     Or(Col("labels.label4").Eq(Literal("value4")), And(Col("labels.label2").RegexMatch("value."), Col("labels.label1").Eq(Literal("value2")))),
     [_R2, _R3])
_sel(138, "labels.label4 = null", Col("labels.label4").Eq(Literal(None)), [_R1, _R2])
_sel(144, "labels.label4 != null", Col("labels.label4").NotEq(Literal(None)), [_R3])
# filter column that doesn't exist (lines 148-163): every ordering comparison selects nothing
for _ln, _op in ((149, "Gt"), (153, "Lt"), (157, "GtEq"), (161, "LtEq")):
    exec_(_filter, _ln, f"select labels, timestamp, value where doesntexist {_op} 4",
          lambda q, _op=_op: q.Filter(getattr(Col("doesntexist"), _op)(Literal(4))).Project(DynCol("labels"), Col("timestamp"), Col("value")),
          [], kind="filter")
for _ln, _sql, _e, _exp in (
        (165, "stacktrace like 'ack'", Col("stacktrace").Contains("ack"), ["stack1 1", "stack1 2", "stack1 3"]),
        (172, "stacktrace like 'ack2'", Col("stacktrace").Contains("ack2"), []),
        (176, "stacktrace not like 'ack'", Col("stacktrace").ContainsNot("ack"), []),
        (180, "stacktrace not like 'ack2'", Col("stacktrace").ContainsNot("ack2"), ["stack1 1", "stack1 2", "stack1 3"]),
        (187, "labels.label1 not like 'ue2' and stacktrace like 'ack'",
         And(Col("labels.label1").ContainsNot("ue2"), Col("stacktrace").Contains("ack")), ["stack1 1", "stack1 3"])):
    exec_(_filter, _ln, "select stacktrace, value where " + _sql,
          lambda q, _e=_e: q.Filter(_e).Project(Col("stacktrace"), Col("value")), _exp, kind="filter")

# exec/filter/filter_contains: LIKE / NOT LIKE on a non-dynamic string column (schema=bytes)
_fcont = case("logictest/testdata/exec/filter/filter_contains", schema="bytes")
insert(_fcont, ["labels.label1", "labels.label2", "labels.label3", "labels.label4", "timestamp", "value"],
       ["value1 value2 null null 1 foo", "value2 value2 value3 null 2 bar", "value3 value2 null value4 3 baz"])
for _ln, _sql, _e, _exp in (
        (10, "timestamp = 2", Col("timestamp").Eq(Literal(2)), ["value2 value2 value3 null 2 bar"]),
        (15, "value LIKE 'a'", Col("value").Contains("a"), ["value2 value2 value3 null 2 bar", "value3 value2 null value4 3 baz"]),
        (21, "value NOT LIKE 'a'", Col("value").ContainsNot("a"), ["value1 value2 null null 1 foo"])):
    exec_(_fcont, _ln, "select labels, timestamp, value where " + _sql,
          lambda q, _e=_e: q.Filter(_e).Project(DynCol("labels"), Col("timestamp"), Col("value")), _exp, kind="filter")

# exec/filter/filter_projection
_fproj = case("logictest/testdata/exec/filter/filter_projection")
insert(_fproj, ["labels.label1", "labels.label2", "labels.label3", "labels.label4", "stacktrace", "timestamp", "value"],
       ["value1 value2 null null stack1 1 1", "value2 value2 value3 null stack1 2 2", "value3 value2 null value4 stack1 3 3"])
exec_(_fproj, 11, "select labels where timestamp >= 2",
      lambda q: q.Filter(Col("timestamp").GtEq(Literal(2))).Project(DynCol("labels")),
      ["value2 value2 value3 null", "value3 value2 null value4"], kind="filter")
exec_(_fproj, 18, "select timestamp where timestamp >= 2",
      lambda q: q.Filter(Col("timestamp").GtEq(Literal(2))).Project(Col("timestamp")), ["2", "3"], kind="filter")
exec_(_fproj, 25, "select value where labels.label5 = null and labels.label3 != null",
      lambda q: q.Filter(And(Col("labels.label5").Eq(Literal(None)), Col("labels.label3").NotEq(Literal(None)))).Project(Col("value")),
      ["2"], kind="filter")
exec_(_fproj, 31, "select value where labels.label5 != null and labels.label3 != null",
      lambda q: q.Filter(And(Col("labels.label5").NotEq(Literal(None)), Col("labels.label3").NotEq(Literal(None)))).Project(Col("value")),
      [], kind="filter")
exec_(_fproj, 36, "select value where (labels.label3 = 'value3' and labels.label5 = null) or (labels.label3 = null and labels.label5 = 'a')",
      lambda q: q.Filter(Or(And(Col("labels.label3").Eq(Literal("value3")), Col("labels.label5").Eq(Literal(None))),
                            And(Col("labels.label3").Eq(Literal(None)), Col("labels.label5").Eq(Literal("a"))))).Project(Col("value")),
      ["2"], kind="filter")

# ---------------------------------------------------------------------------------------------------
# exec/distinct/distinct
# ---------------------------------------------------------------------------------------------------
_dist = case("logictest/testdata/exec/distinct/distinct")
insert(_dist, ["labels.label1", "labels.label2", "labels.label3", "labels.label4", "labels.label5"],
       ["value1 value1 null null value1", "value2 value2 value3 null value1", "value3 value1 null value4 value1"])


def _d(line, sql, exprs, expected, where=None):
    def build(q, exprs=exprs, where=where):
        if where is not None:
            q = q.Filter(where)
        return q.Distinct(*exprs)
    exec_(_dist, line, sql, build, expected, unordered=True, kind="distinct")


_d(10, "select distinct(labels.label1)", _c("labels.label1"), ["value1", "value2", "value3"])
_d(17, "select distinct(labels.label2)", _c("labels.label2"), ["value1", "value2"])
_d(23, "select distinct(labels.label3)", _c("labels.label3"), ["null", "value3"])
_d(29, "select distinct(labels.label1, labels.label2)", _c("labels.label1", "labels.label2"),
   ["value1 value1", "value2 value2", "value3 value1"])
_d(38, "select distinct(labels.label1, labels.label2, labels.label3)", _c("labels.label1", "labels.label2", "labels.label3"),
   ["value1 value1 null", "value2 value2 value3", "value3 value1 null"])
_d(45, "select distinct(labels.label1, labels.label2, labels.label4)", _c("labels.label1", "labels.label2", "labels.label4"),
   ["value1 value1 null", "value2 value2 null", "value3 value1 value4"])
_d(52, "select distinct(labels.label1, labels.label2, labels.label5)", _c("labels.label1", "labels.label2", "labels.label5"),
   ["value1 value1 value1", "value2 value2 value1", "value3 value1 value1"])
_d(59, "select distinct(labels.label1, labels.label2, labels.label3, labels.label4)",
   _c("labels.label1", "labels.label2", "labels.label3", "labels.label4"),
   ["value1 value1 null null", "value2 value2 value3 null", "value3 value1 null value4"])
_d(67, "select distinct(labels)", [DynCol("labels")],
   ["value1 value1 null null value1", "value2 value2 value3 null value1", "value3 value1 null value4 value1"])
_d(75, "select distinct(labels.label1) where labels.label2 = 'value1' and labels.label4 = 'value4'", _c("labels.label1"), ["value3"],
   where=And(Col("labels.label2").Eq(Literal("value1")), Col("labels.label4").Eq(Literal("value4"))))


# ---------------------------------------------------------------------------------------------------
# exec/aggregate/window: group by a computed bucket, `(timestamp / k) * k as timestamp_bucket`.  sqlparse
# (visitor.go:62-130) plans Project(pre: aggregate inputs + the aliased bucket) -> Aggregate(by Col(alias)) ->
# Project(post); plan/aggregate/window shows the same chain.
# ---------------------------------------------------------------------------------------------------
_win = case("logictest/testdata/exec/aggregate/window")
insert(_win, ["labels.label1", "stacktrace", "timestamp", "value"],
       ["value1 stack1 120000 1", "value2 stack1 121000 2", "value3 stack1 122000 3", "value4 stack1 123000 4"])


def _bucket(k):
    return Mul(Div(Col("timestamp"), Literal(k)), Literal(k)).Alias("timestamp_bucket")


def _win_sum(line, k, expected):
    exec_(_win, line, f"select sum(value) as value_sum, (timestamp/{k})*{k} as timestamp_bucket group by timestamp_bucket",
          lambda q, k=k: q.Project(Col("value"), _bucket(k)).Aggregate([Sum(Col("value"))], _c("timestamp_bucket"))
          .Project(Sum(Col("value")).Alias("value_sum"), Col("timestamp_bucket")),
          expected, unordered=True)


_win_sum(14, 1000, ["1 120000", "2 121000", "3 122000", "4 123000"])
_win_sum(22, 2000, ["3 120000", "7 122000"])
_win_sum(28, 3000, ["6 120000", "4 123000"])
exec_(_win, 34, "select sum(value) as value_sum, count(value) as value_count, (timestamp/3000)*3000 as timestamp_bucket group by timestamp_bucket",
      lambda q: q.Project(Col("value"), _bucket(3000)).Aggregate([Sum(Col("value")), Count(Col("value"))], _c("timestamp_bucket"))
      .Project(Sum(Col("value")).Alias("value_sum"), Count(Col("value")).Alias("value_count"), Col("timestamp_bucket")),
      ["6 3 120000", "4 1 123000"], unordered=True)
_win_sum(40, 4000, ["10 120000"])
exec_(_win, 45, "select labels.label1, (timestamp/5000)*5000 as timestamp_bucket, sum(value) as value_sum group by labels.label1, timestamp_bucket",
      lambda q: q.Project(Col("labels.label1"), _bucket(5000), Col("value")).Aggregate([Sum(Col("value"))], _c("labels.label1", "timestamp_bucket"))
      .Project(Col("labels.label1"), Col("timestamp_bucket"), Sum(Col("value")).Alias("value_sum")),
      ["value1 120000 1", "value2 120000 2", "value3 120000 3", "value4 120000 4"], unordered=True)
exec_(_win, 54, "select (timestamp/2000)*2000 as timestamp_bucket, sum(value) as value_sum, count(timestamp) as timestamp_count group by timestamp_bucket",
      lambda q: q.Project(_bucket(2000), Col("value"), Col("timestamp")).Aggregate([Sum(Col("value")), Count(Col("timestamp"))], _c("timestamp_bucket"))
      .Project(Col("timestamp_bucket"), Sum(Col("value")).Alias("value_sum"), Count(Col("timestamp")).Alias("timestamp_count")),
      ["120000 3 2", "122000 7 2"], unordered=True)
exec_(_win, 60, "select (timestamp/3000)*3000 as timestamp_bucket, count(timestamp) as timestamp_count group by timestamp_bucket",
      lambda q: q.Project(_bucket(3000), Col("timestamp")).Aggregate([Count(Col("timestamp"))], _c("timestamp_bucket"))
      .Project(Col("timestamp_bucket"), Count(Col("timestamp")).Alias("timestamp_count")),
      ["120000 3", "123000 1"], unordered=True)
