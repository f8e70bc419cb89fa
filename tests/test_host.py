"""CPU-only tests of the host half of libfrostgpu: the C-ABI surface, the Parquet parser / run
directories / chunk seeds (checked against pyarrow's independent decode), error behaviour."""
import ctypes as C
import os
import re

import numpy as np
import pyarrow as pa
import pytest

from frostdb_b200 import _lib
from frostdb_b200 import dynparquet as dp
from tests.util import make_columns

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "frostgpu.h")).read()
    declared = set(re.findall(r"\b(fgpu_[a-z_0-9]+)\s*\(", hdr)) - {"fgpu_match_fn"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(built_lib, name), f"{name} is declared in include/frostgpu.h but not exported"
        assert name in _lib._SIGNATURES, f"{name} has no ctypes signature in frostdb_b200/_lib.py"
    assert built_lib.fgpu_abi_version() == _lib.ABI_VERSION


def test_init_without_device_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = _lib.Config(abi_version=_lib.ABI_VERSION, device=0, tile_rows=0, flags=0, staging_bytes=0)
    h = C.c_void_p()
    rc = built_lib.fgpu_init(C.byref(cfg), C.byref(h))
    assert rc == _lib.FGPU_ERR_NO_DEVICE
    assert b"no CUDA device" in built_lib.fgpu_last_error()
    # and the Python mirror refuses to exist without the engine: there is no CPU fallback
    from frostdb_b200.store import ColumnStore
    with pytest.raises(_lib.FrostGPUError):
        ColumnStore(0)


def _check_against_pyarrow(buf):
    d = _lib.describe_parquet(buf)
    ref = dp.read_part(buf)
    row = 0
    for rg in d["row_groups"]:
        for name, c in rg["columns"].items():
            assert "error" not in c, (name, c)
            assert c["tile_index_ok"], name
            exp = ref.column(name).slice(row, rg["n_rows"])
            if pa.types.is_dictionary(exp.type):
                exp = exp.cast(pa.string())
            assert c["decoded"] == exp.to_pylist(), name
        row += rg["n_rows"]
    return d


@pytest.mark.parametrize("version", ["2.0", "1.0"])
@pytest.mark.parametrize("sort", [True, False])
@pytest.mark.parametrize("page_size", [256, 4096, 1 << 20])
def test_run_directories_and_seeds_match_pyarrow(built_lib, version, sort, page_size):
    cols = make_columns(9000, 5, {"a": (6, 0.0), "b": (500, 0.07), "c": (2, 0.95), "d": (40_000, 0.01)}, with_float=True,
                        float_null_p=0.35)
    buf = dp.write_part(dp.SampleDefinitionWithFloat(), cols, sort=sort, row_group_size=3100, data_page_size=page_size,
                        data_page_version=version)
    d = _check_against_pyarrow(buf)
    kinds = {n: c["kind"] for n, c in d["row_groups"][0]["columns"].items()}
    assert kinds["timestamp"] == 1 and kinds["labels.a"] == 2  # PLAIN64 / DICT_STR


def test_edge_shapes(built_lib):
    schema = dp.SampleDefinitionWithFloat()
    # single row, all-NULL dynamic column, rows == chunk boundary +-1
    for n in (1, 127, 128, 129, 2047, 2048, 2049):
        cols = make_columns(n, n, {"a": (3, 0.0), "z": (2, 1.0)}, with_float=True, float_null_p=1.0 if n % 2 else 0.5)
        _check_against_pyarrow(dp.write_part(schema, cols, row_group_size=max(1, n // 2 + 1)))


def test_int64_dictionary_encoded_pages(built_lib):
    """north_star: RLE-dictionary for int64 timestamp columns (pyarrow's default dictionary encoding)."""
    import io
    import pyarrow.parquet as pq
    n = 5000
    t = pa.table({"timestamp": pa.array(np.arange(n) // 7, type=pa.int64()),
                  "value": pa.array((np.arange(n) * 31) % 11, type=pa.int64())})
    t = t.cast(pa.schema([pa.field("timestamp", pa.int64(), nullable=False), pa.field("value", pa.int64(), nullable=True)]))
    sink = io.BytesIO()
    pq.write_table(t, sink, compression="NONE", use_dictionary=True, data_page_version="2.0", store_schema=False)
    d = _check_against_pyarrow(sink.getvalue())
    assert d["row_groups"][0]["columns"]["timestamp"]["kind"] == 3  # DICT64


def test_rejects_what_it_cannot_read(built_lib):
    with pytest.raises(_lib.FrostGPUError) as e:
        _lib.describe_parquet(b"not a parquet file at all")
    assert e.value.code == _lib.FGPU_ERR_PARQUET
    # compressed chunks: parsed, but flagged per column so that only a query projecting them fails
    import io
    import pyarrow.parquet as pq
    sink = io.BytesIO()
    pq.write_table(pa.table({"value": pa.array([1, 2, 3], type=pa.int64())}), sink, compression="SNAPPY")
    d = _lib.describe_parquet(sink.getvalue())
    assert "compressed" in d["row_groups"][0]["columns"]["value"]["error"]
    # truncated file
    good = dp.write_part(dp.SampleDefinition(), make_columns(100, 1, {"a": (3, 0.0)}))
    with pytest.raises(_lib.FrostGPUError):
        _lib.describe_parquet(good[: len(good) // 2])


def test_dictionary_values_of_a_file(built_lib):
    cols = make_columns(3000, 9, {"a": (17, 0.2)})
    buf = dp.write_part(dp.SampleDefinition(), cols, row_group_size=1000)
    vals = _lib.parquet_dict_values(buf, "labels.a")
    ref = set(x for x in dp.read_part(buf)["labels.a"].cast(pa.string()).to_pylist() if x is not None)
    assert set(v.decode() for v in vals) == ref and len(vals) == len(ref)
    assert _lib.parquet_dict_values(buf, "labels.nope") == []


def test_footer_statistics_and_run_length_directories_of_sorted_parts(built_lib):
    """What row-group pruning and the sorted-run kernel rely on: the int64 bounds and null counts of the
    footer survive into the column description, and a sorted dictionary column's directory holds one
    run per distinct value (the writer's literal groups at run boundaries are unpacked on the host)."""
    n = 20_000
    cols = make_columns(n, 77, {"a": (5, 0.0), "b": (37, 0.0), "n": (3, 0.25)}, t0=-500)
    buf = dp.write_part(dp.SampleDefinition(), cols, sort=True, row_group_size=n, data_page_size=2048)
    d = _check_against_pyarrow(buf)
    c = d["row_groups"][0]["columns"]
    assert (c["timestamp"]["min"], c["timestamp"]["max"], c["timestamp"]["null_count"]) == (-500, n - 501, 0)
    v = np.asarray(cols["value"])
    assert (c["value"]["min"], c["value"]["max"]) == (int(v.min()), int(v.max()))
    assert c["labels.n"]["null_count"] == int((np.asarray(cols["labels.n"][0]) < 0).sum())
    assert c["labels.a"]["n_runs"] == 5              # first sort key: one run per value, across all pages
    assert c["labels.b"]["n_runs"] <= 5 * 37         # second key: one run per (a, b) group
    buf = dp.write_part(dp.SampleDefinition(), cols, sort=True, row_group_size=n, write_statistics=False)
    c = _lib.describe_parquet(buf)["row_groups"][0]["columns"]
    assert "min" not in c["timestamp"]
    # nullable sort keys: the two run-length streams (levels, indices) merge into one directory in row space
    cols = make_columns(n, 78, {"a": (5, 0.2), "b": (37, 0.3), "z": (2, 1.0)}, t0=0)
    d = _check_against_pyarrow(dp.write_part(dp.SampleDefinition(), cols, sort=True, row_group_size=7_001, data_page_size=2048))
    seen = 0
    for rg in d["row_groups"]:
        for name in ("labels.a", "labels.b"):
            c = rg["columns"][name]
            assert c["row_runs_ok"] and c["n_row_runs"] <= c["n_runs"] + 2 * c["n_defruns"] + 1, (name, c["n_row_runs"])
            seen += c["has_nulls"]
    assert seen > 0


def test_row_group_pruning_decisions_against_reference_vectors(built_lib):
    """The engine's own statistics decision (fgpu_rowgroup_leaf_mode, the function compile() uses) replayed on
    TestBinaryScalarOperation (query/expr/binaryscalarexpr_test.go:56-212): it must never rule out a row group
    the reference keeps, it must rule out every bounded chunk the reference rules out, and "every row passes" needs
    bounds inside the range and no NULLs."""
    from tests.golden import rowgroup_filter_cases as g
    ruled_out = 0
    for name, mn, mx, right, nulls, op, expect in g.CASES:
        if right is None:
            continue  # numeric comparison with the NULL literal is decided at row level (nothing is selected)
        mode = C.c_int32(-1)
        has_bounds = nulls != g.NUM_VALUES
        assert built_lib.fgpu_rowgroup_leaf_mode(op, right, int(has_bounds), mn, mx, nulls, g.NUM_VALUES, C.byref(mode)) == 0
        if mode.value == 2:
            assert not expect, name
            ruled_out += 1
        if not expect and has_bounds:
            assert mode.value == 2, name
        if mode.value == 1:
            assert expect and nulls == 0, name
    assert ruled_out >= 4
    # ranges: <, <=, >=, != against [5, 9]
    cases = [(3, 5, 2), (3, 6, 0), (3, 10, 1), (4, 4, 2), (4, 5, 0), (4, 9, 1), (6, 10, 2), (6, 9, 0), (6, 5, 1), (2, 7, 0), (2, 4, 1)]
    for op, lit, want in cases:  # logicalplan.Op: 2 !=, 3 <, 4 <=, 5 >, 6 >=
        mode = C.c_int32(-1)
        assert built_lib.fgpu_rowgroup_leaf_mode(op, lit, 1, 5, 9, 0, 10, C.byref(mode)) == 0
        assert mode.value == want, (op, lit, mode.value)
        if want == 1:  # NULLs in the chunk: "every row passes" is off the table
            assert built_lib.fgpu_rowgroup_leaf_mode(op, lit, 1, 5, 9, 3, 10, C.byref(mode)) == 0 and mode.value == 0
    mode = C.c_int32(-1)
    assert built_lib.fgpu_rowgroup_leaf_mode(2, 7, 1, 7, 7, 2, 10, C.byref(mode)) == 0 and mode.value == 2   # != 7 on a chunk of 7s and NULLs


# ---- split-block bloom filters (N2: expr/binaryscalarexpr.go:104-118, dynparquet/schema.go:1111-1157) ---------------
def test_xxhash64_known_answers_and_python_reference():
    """XXH64 (seed 0) known answers from the xxHash specification, then the library against the plain-Python
    implementation of tests/bloom_file.py over every tail length."""
    import ctypes as C
    from tests import bloom_file as bf
    lib = _lib.load()
    def h(data, seed=0):
        out = C.c_uint64(0)
        assert lib.fgpu_xxhash64(data, len(data), seed, C.byref(out)) == 0
        return out.value
    assert h(b"") == 0xEF46DB3751D8E999
    assert h(b"a") == 0xD24EC4F1A98C6E5B
    assert h(b"abc") == 0x44BC2CF5AD770999
    rng = np.random.default_rng(5)
    for n in list(range(0, 70)) + [127, 128, 129, 1000]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for seed in (0, 0x9E3779B185EBCA87):
            assert h(data, seed) == bf.xxh64(data, seed), (n, seed)


def test_split_block_filter_matches_the_format_description():
    """Insert through the library, check in Python and the other way round (block selection, the eight salts, bit
    positions: Parquet BloomFilter.md); values that were never inserted are mostly rejected, inserted ones never."""
    import ctypes as C
    from tests import bloom_file as bf
    lib = _lib.load()
    rng = np.random.default_rng(6)
    for n_bytes in (32, 64, 1024, 4096 + 32):
        a = bytearray(n_bytes)           # written by Python
        b = (C.c_uint8 * n_bytes)()      # written by the library
        hashes = [int(x) for x in rng.integers(0, 2**63, 200, dtype=np.int64)] + [0, 2**64 - 1, 2**32, 2**32 - 1]
        for hv in hashes:
            bf.sbbf_insert(a, hv)
            assert lib.fgpu_bloom_insert(b, n_bytes, hv) == 0
        assert bytes(a) == bytes(b)
        a_c = (C.c_uint8 * n_bytes).from_buffer_copy(bytes(a))
        out = C.c_int32(0)
        for hv in hashes:
            assert lib.fgpu_bloom_check(a_c, n_bytes, hv, C.byref(out)) == 0 and out.value == 1
        others = [int(x) for x in rng.integers(0, 2**63, 500, dtype=np.int64)]
        agree = 0
        for hv in others:
            assert lib.fgpu_bloom_check(a_c, n_bytes, hv, C.byref(out)) == 0
            assert bool(out.value) == bf.sbbf_check(bytes(a), hv)
            agree += 1 - out.value
        if n_bytes >= 1024:
            assert agree > 300  # a filter with >= 40 bits per value rejects nearly every stranger
    assert lib.fgpu_bloom_check(a_c, 48, 1, C.byref(out)) != 0  # not a multiple of the 32-byte block


def test_row_group_filter_asks_the_bloom_filter_for_equality():
    """A Parquet file whose chunks carry bloom filters (tests/bloom_file.py: pyarrow cannot write them, it reads the
    file back): `x == v` for a value inside [min, max] that is not in the chunk is ruled out by the filter
    (expr/binaryscalarexpr.go:104-118), a value that is there never is; outside the bounds and `== NULL` on a required
    column are ruled out as before.  The answers are compared with the plain-Python filter over 300 absent values."""
    import ctypes as C
    import io
    import pyarrow.parquet as pq
    from tests import bloom_file as bf
    rng = np.random.default_rng(7)
    xs = sorted(set(int(v) for v in rng.integers(0, 1_000_000, 2_000)))
    ys = [int(v) for v in rng.integers(-50, 50, len(xs))]
    buf = bf.write_int64_file({"x": xs, "y": ys})
    assert pq.read_table(io.BytesIO(buf)).to_pydict() == {"x": xs, "y": ys}
    d = _lib.describe_parquet(buf)
    assert d["row_groups"][0]["n_rows"] == len(xs)
    lib = _lib.load()
    def may(col, v, lit_type=_lib.SCALAR_INT64):
        out = C.c_int32(-1)
        assert lib.fgpu_parquet_rowgroup_may_match_eq(buf, len(buf), 0, col.encode(), lit_type, int(v or 0), 0.0, None, 0, C.byref(out)) == 0
        return bool(out.value)
    for v in xs[::37]:
        assert may("x", v)
    # the bitset the writer produced, rebuilt here to predict the filter's answer for absent values
    n_bytes = max(32, ((len(xs) * 10 // 8) + 31) // 32 * 32)
    bits = bytearray(n_bytes)
    for v in xs:
        bf.sbbf_insert(bits, bf.xxh64(v.to_bytes(8, "little", signed=True)))
    present, ruled_out = set(xs), 0
    for v in (int(v) for v in rng.integers(xs[0] + 1, xs[-1], 300)):
        if v in present:
            continue
        expect = bf.sbbf_check(bytes(bits), bf.xxh64(v.to_bytes(8, "little", signed=True)))
        assert may("x", v) == expect, v
        ruled_out += 0 if expect else 1
    assert ruled_out > 250  # 10 bits per value: ~1 % false positives
    assert not may("x", xs[-1] + 5) and not may("x", xs[0] - 1)    # outside the bounds
    assert not may("x", None, _lib.SCALAR_NULL)                    # required column: no NULL to find
    assert may("y", ys[3]) and not may("y", 77)
    assert not may("nope", 1) and may("nope", None, _lib.SCALAR_NULL)  # missing column (:47-73)
