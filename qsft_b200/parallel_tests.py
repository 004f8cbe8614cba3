"""Parameter sweeps over the decoder settings (mirror of qsft/parallel_tests.py:12-114).

run_tests() evaluates one method of a TestHelper for every combination of (num_subsample, num_repeat, b, noise_sd,
iteration) and returns the reference's result table (one pandas row per combination).  The reference fans the
combinations out over a process pool; here every combination is a few GPU kernel launches on the helper's already
sampled signals, so they run back to back in this process (`parallel` is accepted and ignored: a forked pool cannot
share the CUDA context that holds the samples)."""
from __future__ import annotations

import itertools

import numpy as np
import pandas as pd

RESULT_COLUMNS = ["n", "q", "runtime", "found_sparsity", "n_samples", "ratio_samples", "max_hamming_weight", "nmse", "method"]


def run_one(helper, method, num_subsample, num_repeat, b, noise_sd):
    """One decoder run + test NMSE (qsft/parallel_tests.py:12-55)."""
    model_kwargs = {"num_subsample": num_subsample, "num_repeat": num_repeat, "b": b, "noise_sd": noise_sd,
                    "n_samples": num_subsample * (helper.q ** b) * num_repeat * (helper.n + 1)}
    model = helper.compute_model(method=method, model_kwargs=model_kwargs, report=True, verbosity=0)
    beta = model.get("gwht")
    return {"n": helper.n, "q": helper.q, "runtime": model.get("runtime"), "found_sparsity": len(beta),
            "n_samples": model.get("n_samples"), "ratio_samples": model.get("n_samples") / (helper.q ** helper.n),
            "max_hamming_weight": model.get("max_hamming_weight"),
            "nmse": helper.test_model(method=method, beta=beta), "method": method}


def run_tests(test_method, helper, iters, num_subsample_list, num_repeat_list, b_list, noise_sd_list, parallel=True):
    """qsft/parallel_tests.py:58-114: the sweep table joined with the per-run results."""
    params = list(itertools.product(num_subsample_list, num_repeat_list, b_list, noise_sd_list, range(iters)))
    test_df = pd.DataFrame(data=params, columns=["num_subsample", "num_repeat", "b", "noise_sd", "iter"])
    rows = [run_one(helper, test_method, int(C), int(R), int(b), noise_sd) for (C, R, b, noise_sd, _) in params]
    results_df = pd.DataFrame(data=rows, columns=RESULT_COLUMNS if not rows else None)
    return pd.concat([test_df, results_df], axis=1, join="inner")
