#!/usr/bin/env python
"""bench.py — rows/s of scan + filter + hash-aggregate on the 100M-row Parca schema (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W [--impl frostgpu|reference]

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE).  A step is one execution of

    ScanTable(t).Filter(timestamp in the middle 50% of the range)
                .Aggregate([Sum(value), Count(value)], [labels.l00, labels.l01])

over this rank's parts: 100M rows per GPU (weak scaling), 16 dynamic label columns, parts of 4Mi
rows sorted the way compaction leaves them, row groups of 1Mi rows.  N > 1 adds the one exchange
step the path has: an all-gather of the per-rank partial aggregate tables and the merge kernel.

`value`  parts resident in HBM before the timed region (fgpu_query_execute only).
`e2e`    the same query through the public C-ABI from HOST Parquet buffers: every step puts the
         parts (host parse + H2D), executes, reads the result record back and drops the parts.
`--impl reference` times the CPU restatement of the reference path (oracle/, a port: the Go
         engine cannot be built here) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_data as bd  # noqa: E402
from frostdb_b200 import logicalplan as lp  # noqa: E402

METRIC = "rows/sec scan+filter+hash-agg (100M-row Parca schema)"
N_LABELS = 16
TABLE = "bench"


def env_int(name, default):
    return int(os.environ.get(name, default))


def headline_query_exprs(total_first_row: int, rows: int):
    lo = bd.T0 + total_first_row + rows // 4
    hi = bd.T0 + total_first_row + (3 * rows) // 4
    filt = lp.And(lp.Col("timestamp").GtEq(lp.Literal(lo)), lp.Col("timestamp").Lt(lp.Literal(hi)))
    aggs = [lp.Sum(lp.Col("value")), lp.Count(lp.Col("value"))]
    groups = [lp.Col("labels.l00"), lp.Col("labels.l01")]
    return filt, aggs, groups


def load_files(paths):
    return [np.fromfile(p, dtype=np.uint8) for p in paths]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args, rows_per_gpu):
    """CPU restatement of the reference path on all host cores, bounded sample (rank 0 only)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from frostdb_b200.physicalplan import GPUScan
    from frostdb_b200 import _lib
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    threads = min(cores, env_int("FROSTGPU_REF_THREADS", cores))
    sample_rows = min(rows_per_gpu, env_int("FROSTGPU_REF_SAMPLE_ROWS", 128 * bd.RG_ROWS))
    paths = bd.generate_parts(rows_per_gpu, N_LABELS)
    need_parts = (sample_rows + bd.PART_ROWS - 1) // bd.PART_ROWS
    bufs = load_files(paths[:need_parts])
    table = orc.OracleTable()
    for b in bufs:
        table.add_pinned(b.ctypes.data, b.nbytes, b)
    filt, aggs, groups = headline_query_exprs(0, rows_per_gpu)
    scan = GPUScan(None, TABLE, filt, _lib.PLAN_AGGREGATE, groups, aggs)
    plan, keep = scan._plan()
    times, scanned = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = table.execute(plan, threads=threads, max_rows=sample_rows)
        dt = time.perf_counter() - t0
        scanned = res.rows_scanned
        res.close()
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = scanned * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(rows_per_gpu, args.gpus),
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": f"{scanned} rows ({scanned // bd.RG_ROWS} row groups of the same parts) per step, C port of the "
                                   "reference's decode->filter->hash-aggregate chain (oracle/frost_oracle.c), one chain per thread"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(rows_per_gpu, n_gpus):
    return {"workload": f"cfg3+filter: {rows_per_gpu} rows per GPU x {n_gpus} GPU(s), Parca SampleDefinition with {N_LABELS} dynamic "
                        "label columns, Filter(timestamp in the middle 50% of the rank's own time range) + Sum(value),Count(value) GROUP BY labels.l00,labels.l01 "
                        "(<=16705 groups); parts of 4Mi rows sorted in compaction order, 1Mi-row row groups, uncompressed, DataPageV2",
            "rows_per_gpu": rows_per_gpu, "label_columns": N_LABELS, "part_rows": bd.PART_ROWS, "row_group_rows": bd.RG_ROWS,
            "l2": "inputs (>=1.6 GB projected per step per GPU) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"parts sharded one range per GPU x{n_gpus}, one NCCL all-gather of the partial aggregate tables + k_merge"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="frostgpu", choices=["frostgpu", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rows_per_gpu = env_int("FROSTGPU_BENCH_ROWS", 100_000_000)
    if args.impl == "reference":
        run_reference(args, rows_per_gpu)
        return

    import torch
    import torch.distributed as dist
    from frostdb_b200 import _lib
    from frostdb_b200.physicalplan import GPUScan
    from frostdb_b200.store import GPUEngine
    import ctypes as C

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the frostgpu arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    first_row = rank * rows_per_gpu
    t_gen = time.perf_counter()
    paths = bd.generate_parts(rows_per_gpu, N_LABELS, first_row=first_row)
    bufs = load_files(paths)
    t_gen = time.perf_counter() - t_gen
    file_bytes = sum(b.nbytes for b in bufs)

    eng = GPUEngine(local)
    # cross-rank dictionary ids: union of every rank's dictionary entries, in rank order, preloaded
    # before the parts are put (one-time, at upload)
    key_cols = ["labels.l00", "labels.l01"]
    unions = {}
    if world > 1:
        for col in key_cols:
            seen, mine = set(), []
            for b in bufs:
                for v in _lib.parquet_dict_values(b, col):
                    if v not in seen:
                        seen.add(v)
                        mine.append(v)
            allv = [None] * world
            dist.all_gather_object(allv, mine)
            seen, union = set(), []
            for vs in allv:
                for v in vs:
                    if v not in seen:
                        seen.add(v)
                        union.append(v)
            unions[col] = union
            eng.dict_preload(TABLE, col, union)
    t_up = time.perf_counter()
    for b in bufs:
        eng.put_parquet(TABLE, b)
    t_up = time.perf_counter() - t_up

    filt, aggs, groups = headline_query_exprs(first_row, rows_per_gpu)
    scan = GPUScan(eng, TABLE, filt, _lib.PLAN_AGGREGATE, groups, aggs)
    q, keep = scan.prepare()
    tx = eng.table_watermark(TABLE)
    q_main, tx_main = q, tx

    def step_single():
        res = C.c_void_p()
        _lib.check(lib.fgpu_query_execute(eng.handle, q, tx, C.byref(res)))
        st = eng.stats(res)
        batches = list(eng.drain(res))
        lib.fgpu_result_free(res)
        return st, batches

    gathered = {}

    def step_multi(q=None, tx=None):
        q = q_main if q is None else q
        tx = tx_main if tx is None else tx
        res, ptr, nbytes = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _lib.check(lib.fgpu_query_execute_partial(eng.handle, q, tx, C.byref(res), C.byref(ptr), C.byref(nbytes)))
        n8 = nbytes.value // 8
        # zero-copy view of the library's partial table (the scan has completed on its stream)
        mine = torch.as_tensor(_DevMem(ptr.value, nbytes.value), device="cuda")
        additive = C.c_int32(0)
        _lib.check(lib.fgpu_result_partial_is_additive(res, C.byref(additive)))
        if additive.value and os.environ.get("FROSTGPU_BENCH_ALLREDUCE"):
            # dense table of counts and integer sums: the partial -> final step can be one in-place all-reduce
            # (measured at N=2: 0.74 ms/step against 0.63 for all-gather + k_merge, so the gather stays the default)
            dist.all_reduce(mine, op=dist.ReduceOp.SUM)
            torch.cuda.current_stream().synchronize()
            _lib.check(lib.fgpu_result_merge_partials(eng.handle, res, None, 0, 0))
        else:
            if "buf" not in gathered or gathered["buf"].numel() != n8 * world:
                gathered["buf"] = torch.empty(n8 * world, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(gathered["buf"], mine)
            torch.cuda.current_stream().synchronize()
            _lib.check(lib.fgpu_result_merge_partials(eng.handle, res, gathered["buf"].data_ptr(), nbytes.value, world))
        st = eng.stats(res)
        batches = list(eng.drain(res))
        lib.fgpu_result_free(res)
        return st, batches

    step = step_multi if world > 1 else step_single

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        st, batches = step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sync_all()
    t0 = time.perf_counter()
    scan_ms, launches, alg_bytes = [], 0, 0
    for _ in range(args.steps):
        st, batches = step()
        scan_ms.append(st["scan_kernel_ms"])
        launches += st["kernel_launches"]
        alg_bytes = st["algorithmic_bytes"]
    sync_all()
    dt = time.perf_counter() - t0
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()
    tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item())
    total_rows = rows_per_gpu * world
    value = total_rows * args.steps / dt

    # ---- e2e: host Parquet buffers -> result record, every step --------------------------------------
    # The parts sit in page-locked host memory (where the Go side would have written them); a step
    # registers them (footer + page-header parse), executes the query — which builds and uploads only
    # the projected columns, PLAIN pages DMA'd straight from the host buffers — reads the result
    # record back and drops the parts again.
    e2e = None
    if not os.environ.get("FROSTGPU_SKIP_E2E"):
        from frostdb_b200.store import PinnedBuffer
        e2e_steps = max(1, min(args.steps, env_int("FROSTGPU_E2E_STEPS", 3)))
        E2E = "bench_e2e"
        pinned = []
        for b in bufs:
            pb = PinnedBuffer(b.nbytes)
            pb.array[:] = b
            pinned.append(pb)
        scan_e = GPUScan(eng, E2E, filt, _lib.PLAN_AGGREGATE, groups, aggs)
        h2d = {"bytes": 0}

        def e2e_step():
            for col, union in unions.items():  # (multi-rank) same dictionary ids on every rank
                eng.dict_preload(E2E, col, union)
            for pb in pinned:
                eng.put_parquet(E2E, pb.array, borrow=True)
            if world == 1:
                out = []
                scan_e.SetNext(_Collect(out))
                scan_e.Execute(None)
                h2d["bytes"] = scan_e.last_stats["h2d_bytes"]
                d2h = sum(x.nbytes for x in out)
            else:  # partial table per rank -> all-gather -> merge, as in the resident run
                qe, keep_e = scan_e.prepare()
                st_e, batches_e = step_multi(qe, eng.table_watermark(E2E))
                lib.fgpu_query_free(qe)
                h2d["bytes"] = st_e["h2d_bytes"]
                d2h = sum(x.nbytes for x in batches_e)
            eng.drop_table(E2E)
            return d2h

        e2e_step()  # warm
        sync_all()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(e2e_steps):
            d2h = e2e_step()
        sync_all()
        de = time.perf_counter() - t0
        tde = torch.tensor([de], dtype=torch.float64, device="cuda")
        tb = torch.tensor([float(h2d["bytes"])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tde, op=dist.ReduceOp.MAX)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        de, h2d["bytes"] = float(tde.item()), int(tb.item())
        e2e = {"value": rows_per_gpu * world * e2e_steps / de, "unit": "rows/s", "h2d_bytes_per_step": int(h2d["bytes"]),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": 1000.0 * de / e2e_steps,
               "note": "every step: fgpu_part_put_parquet(BORROW_PINNED) of all parts from page-locked host memory (footer/page "
                       "parse), fgpu_query_execute (builds + uploads the projected columns, then the scan), result record read "
                       "back, fgpu_table_drop"}
        for pb in pinned:
            pb.close()

    # ---- CPU baseline (rank 0, N == 1) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not os.environ.get("FROSTGPU_SKIP_CPU"):
        from oracle import oracle as orc
        cores = os.cpu_count() or 1
        sample_rows = min(rows_per_gpu, env_int("FROSTGPU_REF_SAMPLE_ROWS", 128 * bd.RG_ROWS))
        need_parts = (sample_rows + bd.PART_ROWS - 1) // bd.PART_ROWS
        table = orc.OracleTable()
        for b in bufs[:need_parts]:
            table.add_pinned(b.ctypes.data, b.nbytes, b)
        plan, keep2 = scan._plan()
        best, scanned = None, 0
        for _ in range(3):
            t0 = time.perf_counter()
            r = table.execute(plan, threads=cores, max_rows=sample_rows)
            d = time.perf_counter() - t0
            scanned = r.rows_scanned
            r.close()
            best = d if best is None else min(best, d)
        table.close()
        cpu = {"value": scanned / best, "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": f"{scanned} rows of the same parts, best of 3, C port of the reference chain (oracle/frost_oracle.c)"}

    if rank == 0:
        peak, peak_src = measured_peak()
        avg_scan_ms = float(np.mean(scan_ms))
        achieved = alg_bytes / (avg_scan_ms * 1e-3) / 1e9 if avg_scan_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel: from the committed ncu --set full capture of this same
        # workload (profiles/r1_traffic.json), only when the launch processes the same bytes
        traffic = None
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")) as f:
                tr = json.load(f)
            if tr["workload_rows_per_gpu"] == rows_per_gpu and abs(tr["algorithmic_bytes"] - alg_bytes) <= 0.01 * alg_bytes:
                traffic = int(tr["dram_bytes_read"] + tr["dram_bytes_write"])
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": workload_config(rows_per_gpu, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "fgpu::k_runs (sorted-run scan; fgpu::k_scan takes unsorted / nullable-key row groups)", "kernel_ms": avg_scan_ms, "algorithmic_bytes": int(alg_bytes),
                         "peak_source": peak_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
            "groups": int(st["groups"]), "rows_selected_per_gpu": int(st["rows_selected"]),
            "setup": {"generate_s": round(t_gen, 2), "upload_s": round(t_up, 2), "parquet_bytes_per_gpu": int(file_bytes)},
        }
        print(json.dumps(line), flush=True)
    lib.fgpu_query_free(q)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


class _Collect:
    def __init__(self, out):
        self.out = out

    def Callback(self, ctx, r):
        self.out.append(r)

    def Finish(self, ctx):
        return None


class _DevMem:
    """__cuda_array_interface__ wrapper so torch can view library-owned device memory without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8", "data": (ptr, False), "version": 2}


if __name__ == "__main__":
    main()
