"""The reference's synt_exp/quick_example.py flow on the B200 engine: only the imports change.

    python examples/quick_example.py [--q 3 --n 40 --sparsity 100 --b 4 --noise-sd 1 --t 4 --coded]

Needs a CUDA device (there is no CPU fallback)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsft_b200 import QSFT, get_random_subsampled_signal, get_reed_solomon_dec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", type=int, default=3)
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--sparsity", type=int, default=100)
    ap.add_argument("--b", type=int, default=4)
    ap.add_argument("--noise-sd", type=float, default=1.0)
    ap.add_argument("--num-subsample", type=int, default=3)
    ap.add_argument("--num-repeat", type=int, default=1)
    ap.add_argument("--t", type=int, default=4, help="maximum Hamming weight of the support (and RS error capability)")
    ap.add_argument("--coded", action="store_true", help="Reed-Solomon source delays (prime q) instead of identity")
    ap.add_argument("--channel", default="identity", choices=["identity", "nso"])
    ap.add_argument("--seed", type=int, default=20)
    a = ap.parse_args()

    np.random.seed(a.seed)
    source = "coded" if a.coded else "identity"
    query_args = {"query_method": "complex", "num_subsample": a.num_subsample, "delays_method_source": source,
                  "subsampling_method": "qsft", "delays_method_channel": a.channel, "num_repeat": a.num_repeat,
                  "b": a.b, "t": a.t}
    qsft_args = {"num_subsample": a.num_subsample, "num_repeat": a.num_repeat, "reconstruct_method_source": source,
                 "reconstruct_method_channel": a.channel, "b": a.b, "noise_sd": a.noise_sd,
                 "source_decoder": get_reed_solomon_dec(a.n, a.t, a.q) if a.coded else None}
    signal = get_random_subsampled_signal(n=a.n, q=a.q, sparsity=a.sparsity, a_min=1, a_max=1, noise_sd=a.noise_sd,
                                          query_args=query_args, max_weight=a.t)
    result = QSFT(**qsft_args).transform(signal, verbosity=1, timing_verbose=True, report=True, sort=True)

    gwht = result["gwht"]
    diff = dict(signal.signal_w)
    for k, v in gwht.items():
        diff[k] = diff.get(k, 0) - v
    nmse = np.sum(np.abs(list(diff.values())) ** 2) / np.sum(np.abs(list(signal.signal_w.values())) ** 2)
    print(f"found {len(gwht)} of {len(signal.signal_w)} non-zero coefficients")
    print("Total samples = ", result["n_samples"])
    print("Total sample ratio = ", result["n_samples"] / a.q ** a.n)
    print("NMSE = ", nmse)
    print("AVG Hamming Weight of Nonzero Locations = ", result["avg_hamming_weight"])
    print("Max Hamming Weight of Nonzero Locations = ", result["max_hamming_weight"])


if __name__ == "__main__":
    main()
