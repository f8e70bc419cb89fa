"""Where an Execute of a prepared query spends its time on the host (runs on the GPU box): the C call alone, + stats,
+ draining the record into pyarrow; FROSTGPU_PROFILE=1 adds the library's own phase times on stderr."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_data as bd
from frostdb_b200 import _lib, logicalplan as lp
from frostdb_b200.physicalplan import GPUScan
from frostdb_b200.store import GPUEngine

rows = int(os.environ.get("PROBE_ROWS", 100_000_000))
lib = _lib.load()
eng = GPUEngine(0)
for p in bd.generate_parts(rows, 16, sort=True):
    eng.put_parquet("s", np.fromfile(p, dtype=np.uint8))
ts, val = lp.Col("timestamp"), lp.Col("value")
K01 = [lp.Col("labels.l00"), lp.Col("labels.l01")]
F50 = lp.And(ts.GtEq(lp.Literal(bd.T0 + rows // 4)), ts.Lt(lp.Literal(bd.T0 + 3 * rows // 4)))
scan = GPUScan(eng, "s", F50, _lib.PLAN_AGGREGATE, K01, [lp.Sum(val), lp.Count(val)])
q, keep = scan.prepare()
tx = eng.table_watermark("s")


def loop(n, mode):
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 0.0
    for _ in range(n):
        res = C.c_void_p()
        _lib.check(lib.fgpu_query_execute(eng.handle, q, tx, C.byref(res)))
        if mode >= 1:
            st = eng.stats(res)
            k += st["scan_kernel_ms"]
        if mode >= 2:
            b = list(eng.drain(res))
        lib.fgpu_result_free(res)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, k / n


loop(5, 2)
for mode, name in ((0, "execute + free"), (1, "+ stats"), (2, "+ drain into pyarrow")):
    ms, k = loop(50, mode)
    print(f"{name:24s} {ms:.4f} ms/step   (scan kernel {k:.4f})", flush=True)
os.environ["FROSTGPU_PROFILE"] = "1"
