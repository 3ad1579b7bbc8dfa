"""CPU, world_size 2 over gloo: the multi-GPU host logic (batch sharding + single all-gather of the per-sequence
representations, esm_b200/extract.py) without a GPU.  Rendezvous on 127.0.0.1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from esm_b200.extract import all_gather_rows, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 255, 256, 257):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_items, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_items * 5, dtype=torch.float32).view(n_items, 5) * 0.5 + 1.0
        s, e = shard_range(n_items, world, rank)
        got = all_gather_rows(full[s:e].clone(), n_items)
        ret[rank] = bool(torch.equal(got, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7, 1])
def test_all_gather_rows_world2_gloo(n_items):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
