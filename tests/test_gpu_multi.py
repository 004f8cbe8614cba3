"""Real multi-GPU parity: two NCCL ranks (one process per GPU) run ONE transform with the delay rows sharded over the ranks
and must return exactly what a single GPU returns -- same keys in the same first-seen order, values to 1e-5 -- for both
peel placements (bin-sharded on-device loop with the finds exchanged inside the kernel, bin-sharded with one NCCL
all-gather per round, and replicated).  Skipped on a box with one GPU."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # n, q, S, b, C, R, chan, noise_sd
    (20, 4, 1000, 7, 3, 1, "nso", 0.0),          # lattice GEMM path (q = 4, b >= 7), noiseless
    (14, 4, 300, 5, 3, 2, "nso", 0.05),          # small b: K1 + K2 path, repeats, device noise shared through rank 0's seed
    (12, 3, 60, 4, 3, 1, "identity", 0.0),       # odd q: plain (non-TMA) peel tiles
]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _transform(case, dist, peel_mode=None):
    import qsft_b200
    n, q, S, b, C, R, chan, noise_sd = case
    qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
          "delays_method_channel": chan, "num_repeat": R, "b": b}
    np.random.seed(77)
    kw = {"dist": dist} if dist is not None else {}
    sig = qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=noise_sd, query_args=dict(qa),
                                                  noise_rng="device", **kw)
    res = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel=chan).transform(sig, output="arrays")
    return res["locations"].copy(), res["values"].copy()


def _worker(rank, world, port, out):
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from qsft_b200.dist import DistContext
    ok, msg = True, ""
    try:
        for ci, case in enumerate(CASES):
            want_k, want_v = _transform(case, None)                       # this rank alone, same seed
            for mode in ("replicated", "sharded", "sharded_host"):
                k, v = _transform(case, DistContext(peel_mode=mode))
                same = k.shape == want_k.shape and np.array_equal(k, want_k)
                err = float(np.max(np.abs(v - want_v))) if same and len(v) else 0.0
                if not same or err > 1e-5 * max(1.0, float(np.max(np.abs(want_v)))) or len(k) < 0.9 * case[2]:
                    ok, msg = False, f"case {ci} mode {mode}: same={same} err={err} found={len(k)}"
        # the support staged by the ranks together (each packs and uploads its slice, one all-gather): same result
        from qsft_b200 import ops as _ops
        keep, _ops.SHARD_PACK_MIN_ROWS = _ops.SHARD_PACK_MIN_ROWS, 1
        try:
            for ci in (0, 2):                                              # S = 1000 (even split) and S = 60 over two ranks
                want_k, want_v = _transform(CASES[ci], None)
                k, v = _transform(CASES[ci], DistContext(peel_mode="sharded"))
                if not (k.shape == want_k.shape and np.array_equal(k, want_k) and float(np.max(np.abs(v - want_v))) <= 1e-5):
                    ok, msg = False, f"sharded support staging, case {ci}: result differs"
        finally:
            _ops.SHARD_PACK_MIN_ROWS = keep
        # asynchronous result (nothing read back inside the call): two transforms queued, both report the right support size
        import qsft_b200 as qb
        n, q, S, b, C, R, chan, noise_sd = CASES[0]
        qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
              "delays_method_channel": chan, "num_repeat": R, "b": b}
        dc = DistContext(peel_mode="sharded")
        pend = []
        for seed in (5, 6):
            np.random.seed(seed)
            sig = qb.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=0.0, query_args=dict(qa), dist=dc)
            pend.append((len(sig.signal_w), qb.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                                                    reconstruct_method_channel=chan).transform(sig, output="device_async")))
        for want_n, h in pend:
            st = h.wait()
            if st["distinct"] != want_n:
                ok, msg = False, f"device_async: {st} but the support has {want_n} coefficients"
        # ranks seeded differently must be refused, not silently mixed
        np.random.seed(1000 + rank)
        import qsft_b200
        n, q, S, b, C, R, chan, noise_sd = CASES[0]
        qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
              "delays_method_channel": chan, "num_repeat": R, "b": b}
        try:
            dc = DistContext()
            qsft_b200.get_random_subsampled_signal(n=n, q=q, sparsity=S, a_min=1, a_max=1, noise_sd=0, query_args=dict(qa), dist=dc)
            dc.verify()                       # the comparison is queued at construction and read here / at the end of transform
            ok, msg = False, "differently seeded ranks were not refused"
        except RuntimeError as exc:
            if "differ between ranks" not in str(exc):
                raise
    except Exception as exc:                                              # report instead of hanging the peer
        ok, msg = False, repr(exc)
    out[rank] = (ok, msg)
    td.barrier()
    td.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_one_gpu():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: (True, ""), 1: (True, "")}, dict(out)
