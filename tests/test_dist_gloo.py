"""CPU, world_size 2 over gloo: the N > 1 host path (contiguous image shards, one all-gather of fixed-size result
records, global order restored, ragged last shard)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from betapose_b200 import _lib, dist as D

    lo, hi = D.shard_range(n_total, rank, world)
    rec = np.zeros(hi - lo, dtype=np.dtype([("image_index", "<i4"), ("status", "<i4"), ("pad", "u1", _lib.RECORD_BYTES - 8)]))
    rec["image_index"] = np.arange(lo, hi)
    rec["status"] = 1
    rec["pad"][:, 0] = rank + 1
    local = D.records_to_bytes(rec)  # also for an empty shard (n_total < world)
    allrec = D.gather_records(local, n_total)
    got = allrec.numpy().view(rec.dtype).reshape(-1)
    ok = (got["image_index"].tolist() == list(range(n_total))) and bool((got["status"] == 1).all())
    owners = got["pad"][:, 0].tolist()
    ok = ok and owners == [r + 1 for r in range(world) for _ in range(D.shard_sizes(n_total, world)[r])]
    q.put((rank, ok, allrec.shape[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_gather_records_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and n == n_total for _, ok, n in res), res
