"""Stand-alone timings of the HBM-bound kernels (K1, K3, K4) at config-5 sizes: GB/s against MEASURED_PEAKS.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qsft_b200  # noqa: E402
from qsft_b200 import ops, utils  # noqa: E402

ONLY = set(sys.argv[sys.argv.index("--only") + 1].split(",")) if "--only" in sys.argv else {"k1", "k3", "k4"}
dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
q, n, b, P = 4, 40, 10, 41
B = q ** b


def timeit(fn, iters=10, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)     # 256 MB > L2
out = {}
x = torch.randn((P, B, 2), device=dev).view(torch.float32)
xc = torch.view_as_complex(x.view(P, B, 2))
ms = timeit(lambda: ops.gwht_batch_(xc, q, b), flush=flush)
out["k3_gwht 41 x 4^10"] = {"ms": ms, "GBps_algorithmic(16B/elem)": 16 * P * B / ms / 1e6, "frac": 16 * P * B / ms / 1e6 / peak}
if "k3sweep" in ONLY:                       # ring depth / fence cost of the TMA pipeline, and the register-staged kernels, same box
    for st in ("2", "3", "4", "5", "6"):
        for nf in ("0", "1"):
            os.environ["QSFT_K3_STAGES"], os.environ["QSFT_K3_NOFENCE"] = st, nf
            ms = timeit(lambda: ops.gwht_batch_(xc, q, b), flush=flush)
            out[f"k3_gwht 41 x 4^10, stages={st} nofence={nf}"] = {"ms": ms, "frac": 16 * P * B / ms / 1e6 / peak}
    os.environ.pop("QSFT_K3_STAGES")
    os.environ.pop("QSFT_K3_NOFENCE")
    os.environ["QSFT_K3_IMPL"] = "1"
    ms = timeit(lambda: ops.gwht_batch_(xc, q, b), flush=flush)
    out["k3_gwht 41 x 4^10, register-staged kernels (QSFT_K3_IMPL=1)"] = {"ms": ms, "frac": 16 * P * B / ms / 1e6 / peak}
    os.environ.pop("QSFT_K3_IMPL")
for bb, rows in ([(7, 1024), (8, 512), (9, 128), (6, 4096), (12, 4)] if "k3" in ONLY else []):
    y = torch.view_as_complex(torch.randn((rows, q ** bb, 2), device=dev))
    ms = timeit(lambda: ops.gwht_batch_(y, q, bb), flush=flush)
    out[f"k3_gwht {rows} x 4^{bb}"] = {"ms": ms, "GBps": 16 * rows * q ** bb / ms / 1e6, "frac": 16 * rows * q ** bb / ms / 1e6 / peak}
if "k3" in ONLY:
    y = torch.view_as_complex(torch.randn((64, 3 ** 12, 2), device=dev))
    ms = timeit(lambda: ops.gwht_batch_(y, 3, 12), flush=flush)
    out["k3_gwht 64 x 3^12"] = {"ms": ms, "GBps": 16 * 64 * 3 ** 12 / ms / 1e6, "frac": 16 * 64 * 3 ** 12 / ms / 1e6 / peak}

rng = np.random.default_rng(0)
M, D = rng.integers(0, q, (n, b)), rng.integers(0, q, (P, n))
if "k1" in ONLY:
    ms = timeit(lambda: ops.query_lattice(M, D, q, device=dev, want_idx=True, want_digits=False), flush=flush)
    out["k1_lattice idx only (16 B/index)"] = {"ms": ms, "GBps": 16 * P * B / ms / 1e6, "frac": 16 * P * B / ms / 1e6 / peak}
    ms = timeit(lambda: ops.query_lattice(M, D, q, device=dev, want_idx=False, want_digits=True), flush=flush)
    out["k1_lattice digits only (64 B/row)"] = {"ms": ms, "GBps": 64 * P * B / ms / 1e6, "frac": 64 * P * B / ms / 1e6 / peak}
if "k4" not in ONLY:
    print(json.dumps(out, indent=1))
    sys.exit(0)

# K4: one classification round on config-5 bins (C=3, P=41), ~10 % singletons
np.random.seed(1)
S, C = 100_000, 3
sw, locq, strengths = qsft_b200.generate_signal_w(n, q, S, 1, 1, 0, full=False)
Ms, Ds = qsft_b200.get_Ms_and_Ds(n, q, query_method="complex", num_subsample=C, delays_method_source="identity",
                                 delays_method_channel="nso", num_repeat=1, b=b)
ld = utils.padded_ld(n)
loc = ops.pad_digits(locq.T, ld, dev)
a = torch.from_numpy(strengths.astype(np.complex64)).to(dev)
U0 = torch.stack([ops.closed_form_bins(Ms[c], Ds[c][0], q, loc, a) for c in range(C)]).contiguous()
Dall = np.stack([np.vstack(Ds[c]) for c in range(C)])
prob = ops.PeelProblem(q, n, b, Ms, Dall, n + 1, "nso", "identity", 1e-9, dev)
prob.alloc(4 * C * B, C * B)


def classify_round1():
    prob.counters.zero_()
    prob.classify(U0, 0, B, 1)


ms = timeit(classify_round1, flush=flush)
out["k4_classify_kernel round 1 (stand-alone classification, C=3,P=41,B=4^10)"] = {"ms": ms, "GBps": 8 * C * P * B / ms / 1e6, "frac": 8 * C * P * B / ms / 1e6 / peak}


def peel_sig(Uin):
    nf, nr = prob.peel(Uin)
    torch.cuda.synchronize()
    nf = int((prob.find_cj[:nf] >= 0).sum())                 # valid finds (slots are handed out in chunks)
    nu = prob.n_uniq
    k = prob.uniq_k[:nu].to(torch.int64)
    w = torch.arange(1, k.shape[1] + 1, device=dev, dtype=torch.int64)
    return nf, nr, nu, int(((k * w).sum(dim=1) * (prob.uniq_key[:nu] % 1000003 + 1)).sum()), float(prob.uniq_sum[:nu].abs().sum())


try:
    round_bytes = 8 * C * P * B
    # the whole loop in one persistent kernel (default); U is not modified
    sig_dev = peel_sig(U0)
    ms = timeit(lambda: prob.peel(U0), flush=flush)
    out["k4 peel, device loop (default)"] = {"ms": ms, "rounds": sig_dev[1], "finds": sig_dev[0], "distinct": sig_dev[2],
                                             "GBps(8 B x rounds)": round_bytes * sig_dev[1] / ms / 1e6,
                                             "frac": round_bytes * sig_dev[1] / ms / 1e6 / peak}
    os.environ["QSFT_K4_MAX_ROUNDS"] = "1"
    ms = timeit(lambda: prob.peel(U0), flush=flush)
    out["k4 peel, device loop, round 1 only (QSFT_K4_MAX_ROUNDS=1)"] = {"ms": ms, "GBps": round_bytes / ms / 1e6, "frac": round_bytes / ms / 1e6 / peak}
    os.environ.pop("QSFT_K4_MAX_ROUNDS")
    Uz = torch.zeros_like(U0)
    ms = timeit(lambda: prob.peel(Uz), flush=flush)
    out["k4 peel, device loop, all-zeroton bins (1 round)"] = {"ms": ms, "GBps": round_bytes / ms / 1e6, "frac": round_bytes / ms / 1e6 / peak}
    os.environ["QSFT_K4_NO_TMA"] = "1"
    sig_nt = peel_sig(U0)
    ms = timeit(lambda: prob.peel(U0), flush=flush)
    out["k4 peel, device loop without TMA (QSFT_K4_NO_TMA=1)"] = {"ms": ms, "same_result": sig_nt[:4] == sig_dev[:4]}
    os.environ.pop("QSFT_K4_NO_TMA")
    # host-driven rounds (cross-check path): modifies U in place, so every run starts from a copy
    os.environ["QSFT_K4_IMPL"] = "1"
    U = U0.clone()
    sig_host = peel_sig(U)
    ms = timeit(lambda: (U.copy_(U0), prob.peel(U)), flush=flush)
    ms_copy = timeit(lambda: U.copy_(U0), flush=flush)
    out["k4 peel, host-driven rounds (QSFT_K4_IMPL=1)"] = {"ms": ms - ms_copy, "rounds": sig_host[1], "finds": sig_host[0],
                                                          "same_result_as_device_loop": sig_host[:4] == sig_dev[:4],
                                                          "sum_abs_rel_diff": abs(sig_host[4] - sig_dev[4]) / max(sig_host[4], 1e-30)}
    os.environ.pop("QSFT_K4_IMPL")
    out["copy U (torch) reference"] = {"ms": ms_copy, "GBps": 16 * C * P * B / ms_copy / 1e6}
except Exception as exc:
    out["error in the peel timings"] = repr(exc)
print(json.dumps(out, indent=1))
