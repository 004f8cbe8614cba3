"""Base signal class (mirror of qsft/input_signal.py:9-96, only what the transform path needs)."""
from __future__ import annotations


class Signal:
    """Holds n, q, noise_sd and (optionally) the sparse spectrum `signal_w`.  Full time-domain signals (q^n values)
    are outside the accelerated path; subclasses implement sampling."""

    def __init__(self, **kwargs):
        self._set_params(**kwargs)
        self._init_signal()

    def _set_params(self, **kwargs):
        self.n = kwargs.get("n")
        self.q = kwargs.get("q")
        self.noise_sd = kwargs.get("noise_sd", 0)
        self.N = self.q ** self.n
        self.signal_t = kwargs.get("signal_t")
        self.signal_w = kwargs.get("signal_w")
        self.foldername = kwargs.get("folder")
        self.is_synt = False

    def _init_signal(self):
        raise NotImplementedError("only subsampled signals are supported by the B200 engine")

    def shape(self):
        return tuple(self.q for _ in range(self.n))
