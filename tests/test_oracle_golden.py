"""The CPU oracle against the reference's own golden vectors (no GPU needed).  This is what pins
the oracle: every exec block of the transcribed logictest files must reproduce the reference's
expected output."""
import pytest

from tests import golden_runner as gr
from tests.golden.logictest_cases import CASES
from tests.oracle_scan import OracleEngine, OracleTableHandle, oracle_query


@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("case", CASES, ids=[c["source"].split("/exec/")[1] for c in CASES])
def test_oracle_reproduces_reference_goldens(case, threads):
    eng = OracleEngine(threads=threads)
    try:
        n = 0
        for ex, got, exp in gr.run_case(case, lambda schema: OracleTableHandle(eng, "t", schema),
                                        lambda: oracle_query(eng, "t")):
            gr.check(ex, got, exp)
            n += 1
        assert n > 0
    finally:
        eng.close()


def test_row_group_filter_against_reference_vectors():
    """TestBinaryScalarOperation (query/expr/binaryscalarexpr_test.go:56-212): the oracle's port of the row-group
    filter gives the reference's own expected answers."""
    from oracle import oracle as orc
    from tests.golden import rowgroup_filter_cases as g
    for name, mn, mx, right, nulls, op, expect in g.CASES:
        has_bounds = nulls != g.NUM_VALUES  # upstream: all-NULL chunks carry NULL bounds
        assert orc.chunk_may_match_i64(has_bounds, mn, mx, nulls, g.NUM_VALUES, op, right) == expect, name
