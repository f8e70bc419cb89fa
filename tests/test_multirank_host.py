"""N > 1 host logic on CPU (gloo, world_size 2): every rank contributes the dictionary entries of its
own parts, all ranks derive the same union in the same order — the list each rank preloads with
fgpu_dict_preload before putting its parts, which is what makes partial tables mergeable."""
import os
import tempfile

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from frostdb_b200 import _lib
from frostdb_b200 import dynparquet as dp
from tests.util import make_columns


def union_in_rank_order(lists):
    seen, out = set(), []
    for vs in lists:
        for v in vs:
            if v not in seen:
                seen.add(v)
                out.append(v)
    return out


def _worker(rank, world, init_file, out_dir):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        # rank r owns labels from a shifted window so the dictionaries really differ
        cols = make_columns(4000, 100 + rank, {"a": (10 + 5 * rank, 0.1)})
        buf = dp.write_part(dp.SampleDefinition(), cols, row_group_size=1500)
        mine = _lib.parquet_dict_values(buf, "labels.a")
        allv = [None] * world
        dist.all_gather_object(allv, mine)
        union = union_in_rank_order(allv)
        with open(os.path.join(out_dir, f"union_{rank}.txt"), "wb") as f:
            f.write(b"\n".join(union))
        assert set(mine) <= set(union)
    finally:
        dist.destroy_process_group()


def test_dictionary_union_is_identical_on_every_rank(built_lib):
    world = 2
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rdzv")
        mp.spawn(_worker, args=(world, init_file, d), nprocs=world, join=True)
        unions = [open(os.path.join(d, f"union_{r}.txt"), "rb").read() for r in range(world)]
        assert unions[0] == unions[1] and len(unions[0]) > 0
        assert len(unions[0].split(b"\n")) == 15  # rank 1's window is a superset: v000000..v000014
