"""world_size-2 gloo tests of the multi-GPU host logic (row sharding, variable-length all-gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from qsft_b200.dist import DistContext, bin_range
    ctx = DistContext()
    # row gather: 5 rows over 2 ranks -> 3 per rank, buffer 6 rows
    per = 3
    buf = torch.zeros((per * world, 4), dtype=torch.complex64)
    lo, hi = rank * per, min(5, (rank + 1) * per)
    for r in range(lo, hi):
        buf[r] = torch.full((4,), complex(r + 1, -r))
    ctx.all_gather_rows_(buf, per)
    want = torch.zeros((6, 4), dtype=torch.complex64)
    for r in range(5):
        want[r] = torch.full((4,), complex(r + 1, -r))
    ok1 = torch.equal(buf, want)
    # variable length gather
    cap = 8
    cnt = 3 if rank == 0 else 5
    a = torch.arange(cap, dtype=torch.int64) + 100 * rank
    k = (torch.arange(cap * 2, dtype=torch.int8).reshape(cap, 2) + rank)
    z = torch.arange(cap, dtype=torch.float32).to(torch.complex64) * (1j if rank else 1)
    (ga, gk, gz), counts = ctx.all_gather_var([a, k, z], cnt, cap, extra=rank + 1)
    ok2 = counts == [3, 5] and ctx.last_extra_sum == 3 and ga.tolist() == [0, 1, 2, 100, 101, 102, 103, 104] and gk.shape == (8, 2) \
        and gz[3:].tolist() == [complex(0, v) for v in range(5)]
    ok3 = bin_range(10, 0, 2) == (0, 5) and bin_range(10, 1, 2) == (5, 10) and bin_range(3, 1, 4) == (1, 2) \
        and bin_range(3, 3, 4) == (3, 3)
    # peel placement policy: replicated while U is small, sharded above the threshold or on request
    # (gloo: no symmetric memory, so "sharded" falls back to the host-driven exchange)
    ok4 = ctx.shard_peel(1 << 30) == "" and ctx.shard_peel(9 << 30) == "host" and DistContext(peel_mode="sharded").shard_peel(8) == "host" \
        and DistContext(peel_mode="sharded_host").shard_peel(8) == "host" \
        and DistContext(peel_mode="replicated").shard_peel(1 << 40) == "" and not ctx.symmetric
    # agreement of host-side randomness: equal arrays pass, rank-dependent ones raise on every rank; rank 0's indices win
    ctx.assert_same("test arrays", np.arange(6).reshape(2, 3), np.ones(4, dtype=np.int8))
    try:
        ctx.assert_same("test arrays", np.arange(6) + rank)
        ok5 = False
    except RuntimeError as exc:
        ok5 = "differ between ranks" in str(exc)
    ok5 = ok5 and ctx.from_rank0(np.array([3 + rank, 1, 2 * rank])).tolist() == [3, 1, 0]
    ctx.post_check("equal", np.arange(5))                   # deferred form: queued now, read in verify()
    ctx.verify()
    ctx.post_check("unequal", np.arange(5) * (rank + 1))
    try:
        ctx.verify()
        ok5 = False
    except RuntimeError as exc:
        ok5 = ok5 and "differ between ranks" in str(exc)
    ctx.verify()                                            # nothing pending any more
    out[rank] = int(ok1 and ok2 and ok3 and ok4 and ok5)
    td.destroy_process_group()


def test_gloo_world2_collectives():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
