"""Model-quality evaluation on held-out samples (mirror of TestHelper._test_qary, qsft/test_helper.py:235-260).

The reference re-evaluates the recovered sparse model  y_hat[m] = sum_k beta[k] w^<m, k>  on test queries with a dense
NumPy matrix product; here it is one call of the evaluation kernel (K2, arbitrary non-lattice queries)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from . import ops
from .utils import NpEncoder, index_limbs, ints_to_limbs, padded_ld


def evaluate_model(beta, sample_idx_dec, q, n, device=None):
    """y_hat for decimal query indices (Python ints, up to 128 bits) under the sparse model `beta` {tuple(k): coef}."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if len(beta) == 0 or len(sample_idx_dec) == 0:
        return np.zeros(len(sample_idx_dec), dtype=complex)
    ld = padded_ld(n)
    keys = np.array(list(beta.keys()), dtype=np.int8).reshape(len(beta), n)
    vals = np.array(list(beta.values())).astype(np.complex64)
    loc = ops.pad_digits(keys, ld, device)
    a = torch.from_numpy(vals).to(device)
    limbs = ints_to_limbs(sample_idx_dec, index_limbs(q, n))
    idx = torch.from_numpy(limbs.view(np.int64)).to(device)
    dig = ops.dec_to_qary(idx, q, n, ld)
    return ops.eval_synth(dig, loc, a, q, n).cpu().numpy().astype(complex)


def test_nmse(beta, sample_idx_dec, samples, q, n, device=None):
    """|| y_hat - y ||^2 / || y ||^2 on the test set; 1 for an empty model (qsft/test_helper.py:235-260)."""
    if len(beta) == 0:
        return 1
    samples = np.asarray(samples)
    y_hat = evaluate_model(beta, list(sample_idx_dec), q, n, device)
    return float(np.linalg.norm(y_hat - samples) ** 2 / np.linalg.norm(samples) ** 2)


test_nmse.__test__ = False   # not a pytest test


class TestHelper:
    """Experiment harness around the transform path (mirror of qsft/test_helper.py:9-288): builds the training signal(s)
    and a uniformly sampled noiseless test signal under `exp_dir` (train/, train_coded/, test/ -- the reference's cache
    layout, so a directory produced by either implementation can be reused by the other), runs the decoder
    (`compute_model`) and scores a recovered model on the test samples (`test_model`).

    Methods: "qsft" (identity source + NSO channel delays) and "qsft_coded" (Reed-Solomon source delays + NSO).
    "lasso", "gwht" (dense transform) and "qsft_binary" are outside the transform path and raise NotImplementedError
    (in the reference "qsft_binary" loads no data at all, qsft/test_helper.py:89-104).  Subclasses provide
    generate_signal(signal_args), e.g. qsft_b200.synthetic_helper.SyntheticHelper."""
    __test__ = False   # not a pytest class

    # method -> (attribute holding its training signal, sub-folder, delays_method_source)
    _TRAIN = {"qsft": ("train_signal", "train", "identity"), "qsft_coded": ("train_signal_coded", "train_coded", "coded")}

    def __init__(self, signal_args, methods, subsampling_args, test_args, exp_dir, subsampling=True):
        self.n, self.q = signal_args["n"], signal_args["q"]
        self.exp_dir = Path(exp_dir)
        self.subsampling = subsampling
        self.signal_args, self.subsampling_args, self.test_args = signal_args, subsampling_args, test_args
        config_path = self.exp_dir / "config.json"
        if not config_path.is_file():                        # qsft/test_helper.py:19-25
            with open(config_path, "w") as f:
                json.dump({"query_args": subsampling_args}, f, cls=NpEncoder)
        if not subsampling:
            raise NotImplementedError("full (dense q^n) signals are outside the subsampled transform path")
        unsupported = set(methods) - set(self._TRAIN) - {"qsft_binary"}
        if unsupported:
            raise NotImplementedError(f"methods {sorted(unsupported)} are not part of the q-SFT transform path")
        for method, (attr, _, _) in self._TRAIN.items():
            if method in methods:
                setattr(self, attr, self._load_train(method))
        if "qsft_binary" in methods:
            self.train_signal_binary = None
        self.test_signal = self.load_test_data()

    def generate_signal(self, signal_args):
        raise NotImplementedError

    def _load_train(self, method):
        _, folder, source = self._TRAIN[method]
        signal_args = dict(self.signal_args)
        query_args = dict(self.subsampling_args)
        query_args.update({"subsampling_method": "qsft", "query_method": "complex", "delays_method_source": source,
                           "delays_method_channel": "nso"})
        if source == "coded":
            query_args["t"] = signal_args["t"]
        signal_args["folder"] = self.exp_dir / folder
        signal_args["query_args"] = query_args
        return self.generate_signal(signal_args)

    def load_train_data(self):
        return self._load_train("qsft")

    def load_train_data_coded(self):
        return self._load_train("qsft_coded")

    def load_test_data(self):
        """Noiseless uniformly sampled test set (qsft/test_helper.py:115-121)."""
        signal_args = dict(self.signal_args)
        (self.exp_dir / "test").mkdir(exist_ok=True)
        signal_args["query_args"] = {"subsampling_method": "uniform", "n_samples": self.test_args.get("n_samples")}
        signal_args["folder"] = self.exp_dir / "test"
        signal_args["noise_sd"] = 0
        return self.generate_signal(signal_args)

    def compute_model(self, method, model_kwargs, report=False, verbosity=0):
        """Runs QSFT.transform on the method's training signal (qsft/test_helper.py:127-139, 156-196)."""
        from .qsft import QSFT
        from .query import get_reed_solomon_dec
        if method not in self._TRAIN:
            raise NotImplementedError(f"method {method!r} is not part of the q-SFT transform path")
        attr, _, source = self._TRAIN[method]
        if verbosity >= 1:
            print("Estimating GWHT coefficients with QSFT")
        kwargs = dict(reconstruct_method_source=source, reconstruct_method_channel="nso",
                      num_subsample=model_kwargs["num_subsample"], num_repeat=model_kwargs["num_repeat"],
                      b=model_kwargs["b"])
        if source == "coded":
            kwargs["source_decoder"] = get_reed_solomon_dec(self.signal_args["n"], self.signal_args["t"], self.signal_args["q"])
        signal = getattr(self, attr)
        signal.noise_sd = model_kwargs["noise_sd"]           # the decoder's threshold follows the signal (qsft.py:125)
        out = QSFT(**kwargs).transform(signal, verbosity=verbosity, timing_verbose=(verbosity >= 1), report=report)
        if verbosity >= 1:
            print("Found GWHT coefficients")
        return out

    def test_model(self, method, **kwargs):
        if method in ("qsft", "qsft_coded", "lasso"):
            return self._test_qary(**kwargs)
        raise NotImplementedError()

    def _test_qary(self, beta):
        """NMSE of the sparse model `beta` on the test samples (qsft/test_helper.py:235-260), evaluated by K2."""
        signal_t = self.test_signal.signal_t
        return test_nmse(beta, list(signal_t.keys()), list(signal_t.values()), self.q, self.n,
                         getattr(self.test_signal, "device", None))
