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

ONLY = set(sys.argv[sys.argv.index("--only") + 1].split(",")) if "--only" in sys.argv else {"k1", "k3", "k3lag", "k4"}
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
# ticket order of the two-pass kernel: 0 = plain block-by-block order, k = contiguous pass k blocks ahead (default: auto)
for lag in (("0", "1", "2", "3", "4") if "k3lag" in ONLY else ()):
    os.environ["QSFT_K3_LAG"] = lag
    ms = timeit(lambda: ops.gwht_batch_(xc, q, b), flush=flush)
    out[f"k3_gwht 41 x 4^10, QSFT_K3_LAG={lag}"] = {"ms": ms, "frac": 16 * P * B / ms / 1e6 / peak}
os.environ.pop("QSFT_K3_LAG", None)
if "k3lag" in ONLY:                                          # 3 CTAs per SM (80 registers) instead of 4 (64 registers)
    os.environ["QSFT_K3_CTAS"] = "3"
    for lag in ("0", "2", "3"):
        os.environ["QSFT_K3_LAG"] = lag
        ms = timeit(lambda: ops.gwht_batch_(xc, q, b), flush=flush)
        out[f"k3_gwht 41 x 4^10, QSFT_K3_CTAS=3 QSFT_K3_LAG={lag}"] = {"ms": ms, "frac": 16 * P * B / ms / 1e6 / peak}
    os.environ.pop("QSFT_K3_LAG", None)
    os.environ.pop("QSFT_K3_CTAS", None)
for bb, rows in ([(7, 1024), (8, 512), (6, 4096), (12, 4)] if "k3" in ONLY else []):
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
out[f"k4_classify round 1 (C=3,P=41,B=4^10), QSFT_K4_IMPL={os.environ.get('QSFT_K4_IMPL', '1')} FASTDET={os.environ.get('QSFT_K4_FASTDET', '0')}"] = {"ms": ms, "GBps": 8 * C * P * B / ms / 1e6, "frac": 8 * C * P * B / ms / 1e6 / peak}
if "--k4-variants" in sys.argv:
    # the opt-in classification variants in the same process (the knobs are read on every call): time of round 1 and a
    # cheap parity signal -- number of singletons / multitons and an order-independent checksum of the finds
    def signature():
        classify_round1()
        torch.cuda.synchronize()
        nf, nm = int(prob.counters[0]), int(prob.counters[1])
        cj = prob.find_cj[:nf]
        k = prob.find_k[:nf].to(torch.int64)
        w = torch.arange(1, k.shape[1] + 1, device=dev, dtype=torch.int64)
        return nf, nm, int(cj.sum()), int(((k * w).sum(dim=1) * (cj % 1000003 + 1)).sum())

    base_sig = signature()
    for impl, fast in (("1", "1"), ("2", "0"), ("2", "1")):
        os.environ["QSFT_K4_IMPL"], os.environ["QSFT_K4_FASTDET"] = impl, fast
        try:
            sig = signature()
            ms = timeit(classify_round1, flush=flush)
            out[f"k4_classify round 1, QSFT_K4_IMPL={impl} FASTDET={fast}"] = {
                "ms": ms, "frac": 8 * C * P * B / ms / 1e6 / peak, "same_finds_as_default": sig == base_sig,
                "finds": sig[0], "multitons": sig[1]}
        except Exception as exc:
            out[f"k4_classify round 1, QSFT_K4_IMPL={impl} FASTDET={fast}"] = {"error": repr(exc)}
    os.environ.pop("QSFT_K4_IMPL", None)
    os.environ.pop("QSFT_K4_FASTDET", None)
try:                                     # (a faulting opt-in variant above would have poisoned the context: keep what we have)
    Uz = torch.zeros_like(U0)
    ms = timeit(lambda: (prob.counters.zero_(), prob.classify(Uz, 0, B, 1)), flush=flush)
    out["k4_classify all-zeroton round"] = {"ms": ms, "GBps": 8 * C * P * B / ms / 1e6, "frac": 8 * C * P * B / ms / 1e6 / peak}
    U = U0.clone()
    ms = timeit(lambda: (U.copy_(U0), prob.peel(U)), flush=flush)
    ms_copy = timeit(lambda: U.copy_(U0), flush=flush)
    out["k4 full peel loop (3 rounds incl. host syncs)"] = {"ms": ms - ms_copy, "rounds": prob.peel(U0.clone())[1]}
    out["copy U (torch) reference"] = {"ms": ms_copy, "GBps": 16 * C * P * B / ms_copy / 1e6}
except Exception as exc:
    out["error after the variants"] = repr(exc)
print(json.dumps(out, indent=1))
