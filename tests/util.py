"""Synthetic FrostDB tables for the parity tests (seeded, deterministic)."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import pyarrow as pa

from frostdb_b200 import dynparquet as dp


def label_values(card: int) -> List[str]:
    return [f"v{i:06d}" for i in range(card)]


def make_columns(n: int, seed: int, labels: Dict[str, tuple], *, with_float: bool = False, t0: int = 1_000_000,
                 float_null_p: float = 0.0, value_mod: int = 1000) -> Dict[str, object]:
    """labels: name -> (cardinality, null probability).  Columns of dynparquet.SampleDefinition."""
    rng = np.random.default_rng(seed)
    cols: Dict[str, object] = {
        "example_type": (np.zeros(n, np.int32), ["cpu"]),
        "stacktrace": (rng.integers(0, 37, n).astype(np.int32), [f"stack{i:02d}" for i in range(37)]),
        "timestamp": t0 + np.arange(n, dtype=np.int64),
        "value": rng.integers(0, value_mod, n).astype(np.int64),
    }
    for name, (card, pnull) in labels.items():
        idx = rng.integers(0, card, n).astype(np.int32)
        if pnull > 0:
            idx[rng.random(n) < pnull] = -1
        cols[f"labels.{name}"] = (idx, label_values(card))
    if with_float:
        f = rng.integers(0, 10**6, n).astype(np.float64) / 1000.0
        if float_null_p > 0:
            mask = rng.random(n) < float_null_p
            cols["floatvalue"] = pa.array(f, mask=mask)
        else:
            cols["floatvalue"] = pa.array(f)
    return cols


def rows_of(batches: List[pa.RecordBatch], names: Optional[List[str]] = None) -> List[tuple]:
    """Result rows as sorted tuples (dictionary columns decoded, None for NULL / absent column)."""
    out = []
    for b in batches:
        cols = names if names is not None else b.schema.names
        arrays = []
        for n in cols:
            i = b.schema.get_field_index(n)
            if i < 0:
                arrays.append([None] * b.num_rows)
                continue
            a = b.column(i)
            if pa.types.is_dictionary(a.type):
                a = a.dictionary_decode()
            vals = a.to_pylist()
            arrays.append([v.decode() if isinstance(v, bytes) else v for v in vals])
        out.extend(zip(*arrays) if arrays else [])
    return sorted(out, key=lambda r: tuple((x is None, x) for x in r))
