"""World-size-2 test (gloo, CPU) of the case scheduler: partition + final gather reproduce the single-rank table."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from contact_b200 import scheduler


def test_partition_covers_all_cases():
    for n in (0, 1, 7, 148, 1184, 4096):
        for w in (1, 2, 3, 8):
            parts = scheduler.partition(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, ncase, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = scheduler.my_range(ncase, rank, world)
    # per-case "results": case index, its square, rank that solved it
    idx = torch.arange(lo, hi, dtype=torch.float64)
    local = torch.stack([idx, idx * idx, torch.full_like(idx, float(rank))], dim=1)
    table = scheduler.gather_case_results(local, ncase)
    if rank == 0:
        q.put(table.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ncase", [5, 16])
def test_two_rank_gather_matches_single_rank(ncase):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ncase, q)) for r in range(2)]
    for p in procs:
        p.start()
    table = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert table.shape == (ncase, 3)
    assert (table[:, 0] == range(ncase)).all() and (table[:, 1] == table[:, 0] ** 2).all()
    parts = scheduler.partition(ncase, 2)
    assert (table[:parts[0][1], 2] == 0).all() and (table[parts[0][1]:, 2] == 1).all()
