import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared library is a build artefact (git-ignored): build it once if a fresh checkout lacks it
    lib = os.path.join(ROOT, "qsft_b200", "libqsft_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "qsft_b200", "csrc"), "-j8"], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    """GPU-marked tests are skipped (not failed) on a host without CUDA, so a plain `pytest tests` works anywhere."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def u128_to_ints(hi, lo):
    return [(int(h) << 64) | int(l) for h, l in zip(np.asarray(hi).ravel(), np.asarray(lo).ravel())]


FULL_CASES = ["cfg1_q4_n10_b4_identity", "cfg2r_q4_n14_b5_nso_noisy", "q3_n12_b4_lowweight_nso",
              "q2_n12_b4_simple", "q4_n10_allbs_subselect", "q5_n6_b3_identity_noisy"]
WIDE_FULL_CASES = ["cfg4r_q4_n50_b4_lowweight_nso_noisy",    # BASELINE config 4 shape (100-bit indices), reduced
                   "q2_n100_b5_identity_wide"]               # q = 2, 100-bit indices, P_src = 101
NSO2_CASES = ["q4_n10_b4_nso2_noisy", "q3_n9_b3_nso2"]      # reference run with nso_subtype="nso2" (see gen_golden.py)
INDEX_CASES = ["idx_q4_n40_b3", "idx_q4_n50_b2", "idx_q3_n45_b3", "idx_q7_n22_b2", "idx_q2_n100_b5"]


def case_params(g):
    seed, n, q, S, b, C, R, P_src = (int(v) for v in g["meta"])
    mw = int(g["max_weight"])
    qa = {"query_method": str(g["query_method"]), "num_subsample": C, "delays_method_source": str(g["src"]),
          "subsampling_method": "qsft", "delays_method_channel": str(g["chan"]), "num_repeat": R, "b": b}
    if len(g["all_bs"]) > 1:
        qa["all_bs"] = [int(v) for v in g["all_bs"]]
    trC, trR, trb = (int(v) for v in g["tr"])
    return dict(seed=seed, n=n, q=q, S=S, b=b, C=C, R=R, P_src=P_src, max_weight=None if mw < 0 else mw,
                query_args=qa, noise_sd=float(g["noise_sd"]), src=str(g["src"]), chan=str(g["chan"]),
                trC=trC, trR=trR, trb=trb)


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()
