"""Golden vectors of the row-group (TrueNegative) filter, transcribed from the reference's
query/expr/binaryscalarexpr_test.go:56-212 (TestBinaryScalarOperation; numValues = 10 in every case).
right = None stands for the NULL literal (`right: -1` upstream); min/max default to 0 as Go's zero values do.
When null_count == num_values upstream sets min and max to NULL (no bounds)."""
NUM_VALUES = 10
OP_EQ, OP_GT = 1, 5  # logicalplan.Op values (expr.go:13-33), same numbers as FGPU_OP_*

CASES = [
    # name, min, max, right, null_count, op, expect_satisfies
    ("OpEqValueContained", 1, 10, 5, 0, OP_EQ, True),
    ("OpEqValueGt", 1, 10, 11, 0, OP_EQ, False),
    ("OpEqValueLt", 1, 10, 0, 0, OP_EQ, False),
    ("OpEqMaxBound", 1, 10, 10, 0, OP_EQ, True),
    ("OpEqMinBound", 1, 10, 1, 0, OP_EQ, True),
    ("OpEqNullValueNoMatch", 0, 0, None, 0, OP_EQ, False),
    ("OpEqNullValueMatch", 0, 0, None, 1, OP_EQ, True),
    ("OpEqNullColumn", 0, 0, 1, 1, OP_EQ, False),
    ("OpEqFullNullColumn", 0, 0, 1, 10, OP_EQ, False),
    ("OpGtFullNullColumn", 0, 0, 1, 10, OP_GT, False),
    ("OpGtNullValueMatch", 0, 0, None, 0, OP_GT, True),
    ("OpGtNullValueNoMatch", 0, 0, None, 1, OP_GT, True),
    ("OpGtWithSomeNullValuesNoMatch", 1, 10, 11, 1, OP_GT, False),
    ("OpGtWithSomeNullValuesMatch", 1, 10, 5, 1, OP_GT, True),
]
