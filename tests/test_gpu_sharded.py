"""Bin-sharded peeling (qsft_b200.dist.peel_sharded) on ONE GPU: W ranks are simulated by W threads that exchange
their finds through an in-process stand-in for the NCCL all-gather.  The sharded result must equal the single-GPU
peel (support bit-exact, values to 1e-5)."""
import threading

import numpy as np
import pytest
import torch

from conftest import case_params, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import qsft_b200
    from qsft_b200 import ops
    from qsft_b200.dist import peel_sharded
    DEV = torch.device("cuda", 0)


class ThreadDist:
    """Same interface as DistContext, collectives implemented with a barrier and shared slots."""

    def __init__(self, rank, world, shared):
        self.rank, self.world_size, self.sh, self.group = rank, world, shared, None

    def _exchange(self, obj):
        self.sh["slots"][self.rank] = obj
        self.sh["bar"].wait()
        out = list(self.sh["slots"])
        self.sh["bar"].wait()
        return out

    def all_gather_var(self, tensors, count, cap, extra=0):
        torch.cuda.synchronize()
        if isinstance(count, torch.Tensor):
            count, extra = (int(v) for v in count.cpu().tolist())
        parts = self._exchange(([t[:count].clone() for t in tensors], count, int(extra)))
        counts = [p[1] for p in parts]
        self.last_extra_sum = sum(p[2] for p in parts)
        outs = [torch.cat([p[0][i] for p in parts], dim=0) for i in range(len(tensors))]
        return outs, counts

    def all_reduce_sum(self, value, device):
        return sum(self._exchange(int(value)))


def _run_sharded(world, make_problem, U0):
    shared = {"slots": [None] * world, "bar": threading.Barrier(world)}
    results = [None] * world
    distinct = [None] * world
    errors = []

    def work(rank):
        try:
            torch.cuda.set_device(0)
            # all simulated ranks share the default stream: tensors handed between threads stay allocator-safe
            prob = make_problem()
            results[rank] = peel_sharded(prob, U0.clone(), ThreadDist(rank, world, shared))
            distinct[rank] = prob.distinct()
            torch.cuda.synchronize()
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
            shared["bar"].abort()

    torch.cuda.synchronize()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    return results, distinct


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["cfg1_q4_n10_b4_identity", "cfg2r_q4_n14_b5_nso_noisy", "q3_n12_b4_lowweight_nso"])
def test_sharded_peel_equals_single_gpu(name, world):
    g = load_golden(name)
    p = case_params(g)
    q, n = p["q"], p["n"]
    U0 = torch.from_numpy(np.ascontiguousarray(g["mdu_Us"].reshape(p["trC"], -1, q ** p["trb"])).astype(np.complex64)).to(DEV)
    D = g["mdu_Ds"].reshape(p["trC"], -1, n)
    cutoff = 1e-9 + 1.5 * p["noise_sd"] ** 2 / q ** p["trb"]
    Ms = [np.array(M) for M in g["mdu_Ms"]]                  # read the archive here: NpzFile is not thread safe

    def make_problem():
        return ops.PeelProblem(q, n, p["trb"], Ms, D, p["P_src"], p["chan"], p["src"], cutoff, DEV)

    single = make_problem()
    single.alloc(4 * U0.shape[0] * U0.shape[2])
    U1 = U0.clone()
    nf, nr = single.peel(U1)
    want, _ = qsft_b200.QSFT._finds_to_dict(*single.finds(nf))
    results, distinct = _run_sharded(world, make_problem, U0)
    for dk, dv, dc in distinct:
        assert [tuple(int(v) for v in r) for r in dk] == list(want.keys())
        assert max(abs(v - want[key]) for v, key in zip(dv, want)) < 1e-5
    for cj, k, rho, rnd, rounds in results:
        got, _ = qsft_b200.QSFT._finds_to_dict(cj.cpu().numpy(), k.cpu().numpy(), rho.cpu().numpy(), rnd.cpu().numpy())
        assert rounds == nr
        assert list(got.keys()) == list(want.keys())
        assert max(abs(got[key] - want[key]) for key in want) < 1e-5
    ref_keys = [tuple(int(v) for v in kk) for kk in g["res_keys"]]
    assert list(want.keys()) == ref_keys
