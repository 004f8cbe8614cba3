"""Host-side profile of the multi-GPU transform (run under torchrun): where do the non-kernel milliseconds go?"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsft_b200  # noqa: E402
from qsft_b200.dist import DistContext  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
dist = DistContext()
q, n, b, S, C, R = 4, 40, 10, 100_000, 3, 1
qa = {"query_method": "complex", "num_subsample": C, "delays_method_source": "identity", "subsampling_method": "qsft",
      "delays_method_channel": "nso", "num_repeat": R, "b": b}


def one(seed, output):
    np.random.seed(seed)
    sw, locq, st = qsft_b200.generate_signal_w(n, q, S, 1, 1, 0, full=False)
    Ms, Ds = qsft_b200.get_Ms_and_Ds(n, q, **qa)
    td.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    sig = qsft_b200.SyntheticSubsampledSignal(signal_w=sw, locq=locq, strengths=st, noise_sd=0.0, n=n, q=q,
                                              query_args=dict(qa), Ms=Ms, Ds=Ds, noise_rng="device", dist=dist, device=dev)
    torch.cuda.synchronize()
    t1 = time.time()
    res = qsft_b200.QSFT(num_subsample=C, num_repeat=R, b=b, reconstruct_method_source="identity",
                         reconstruct_method_channel="nso").transform(sig, output=output)
    torch.cuda.synchronize()
    t2 = time.time()
    return round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)


probe = dist.symm_acquire(1024, dev)
if rank == 0:
    print("symmetric memory:", probe is not None, "multicast_ptr:", hex(probe[3]) if probe else None, flush=True)
dist.symm_release(probe)
for s in range(3):
    r = one(s, "device")
    if rank == 0:
        print("warm construct/transform ms", r)
pr = cProfile.Profile()
pr.enable()
out = one(5, "device")
pr.disable()
arrs = [one(6 + i, "arrays") for i in range(4)]          # collective: every rank takes part
pr2 = cProfile.Profile()
pr2.enable()
arrs.append(one(11, "arrays"))
pr2.disable()
print(f"rank {rank}: arrays result construct/transform ms", arrs, flush=True)
td.barrier()
if rank == 0:
    print("timed (device result)", out)
    pstats.Stats(pr).sort_stats("cumtime").print_stats(25)
    print("---- arrays result ----")
    pstats.Stats(pr2).sort_stats("cumtime").print_stats(30)
td.destroy_process_group()
