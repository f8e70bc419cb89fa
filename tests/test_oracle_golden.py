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
