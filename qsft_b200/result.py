"""The transform's result container: the reference's {tuple(k): complex} dict (qsft/qsft.py:247-260), filled on first use.

QSFT.transform hands back the distinct coefficients as two host arrays (locations (K, n) int8 in the reference's first-seen
order, values (K,) complex128).  Turning them into K tuples of n Python ints plus K complex objects costs ~0.6 us per
digit in CPython -- 60 ms for K = 1e5, n = 40, as long as the whole sample + FFT + peel on the GPU -- and many callers only
iterate once, look a few keys up, or want the arrays anyway.  SparseSpectrum therefore IS a dict (isinstance, ==, pickle,
iteration order all behave like the reference's result) whose entries are created the first time anything looks at them;
`.locations` / `.coefficients` expose the arrays without ever building the tuples."""
from __future__ import annotations

import itertools

import numpy as np


def _build(loc, val):
    n = loc.shape[1]
    if n == 0:
        return {(): complex(v) for v in val[:1]}
    return zip(itertools.batched(np.ascontiguousarray(loc).astype(np.uint8).tobytes(), n), val.tolist())


class SparseSpectrum(dict):
    """dict {tuple(k): complex}; keys in first-seen order (round, then (group, bin)), values averaged over duplicate finds."""
    __slots__ = ("_loc", "_val", "_pending")

    def __init__(self, locations, values):
        dict.__init__(self)
        self._loc = np.asarray(locations)
        self._val = np.asarray(values)
        self._pending = True

    # -- array views (no tuple construction) -----------------------------------------------------------------
    @property
    def locations(self):
        """(K, n) int8 digits of the recovered k, first-seen order (valid until the dict is modified)."""
        return self._loc

    @property
    def coefficients(self):
        """(K,) complex128 coefficient values matching `locations`."""
        return self._val

    def _fill(self):
        if self._pending:
            self._pending = False
            dict.update(self, _build(self._loc, self._val))
        return self

    # -- everything that looks at or changes the entries fills first ----------------------------------------------
    def __len__(self):
        return len(self._loc) if self._pending else dict.__len__(self)

    def __bool__(self):
        return len(self) > 0

    def __iter__(self):
        return dict.__iter__(self._fill())

    def __reversed__(self):
        return dict.__reversed__(self._fill())

    def __contains__(self, key):
        return dict.__contains__(self._fill(), key)

    def __getitem__(self, key):
        return dict.__getitem__(self._fill(), key)

    def __setitem__(self, key, value):
        dict.__setitem__(self._fill(), key, value)

    def __delitem__(self, key):
        dict.__delitem__(self._fill(), key)

    def __eq__(self, other):
        if isinstance(other, SparseSpectrum):
            other._fill()
        return dict.__eq__(self._fill(), other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return dict.__repr__(self._fill())

    def __or__(self, other):
        return dict(self._fill()) | other

    def __ror__(self, other):
        return other | dict(self._fill())

    def __ior__(self, other):
        dict.update(self._fill(), other)
        return self

    def __reduce__(self):
        return (dict, (dict(self._fill()),))

    def __sizeof__(self):
        return dict.__sizeof__(self._fill())

    def keys(self):
        return dict.keys(self._fill())

    def values(self):
        return dict.values(self._fill())

    def items(self):
        return dict.items(self._fill())

    def get(self, key, default=None):
        return dict.get(self._fill(), key, default)

    def pop(self, *args):
        return dict.pop(self._fill(), *args)

    def popitem(self):
        return dict.popitem(self._fill())

    def setdefault(self, key, default=None):
        return dict.setdefault(self._fill(), key, default)

    def update(self, *args, **kwargs):
        dict.update(self._fill(), *args, **kwargs)

    def clear(self):
        self._pending = False
        dict.clear(self)

    def copy(self):
        return dict(self._fill())


class PendingSpectrum:
    """Result of `QSFT.transform(..., output="device_async")`: the peel has been QUEUED on the device, nothing was read back.

    `wait()` blocks until this transform is done, raises what the synchronous call would have raised (find buffer too
    small, ranks seeded differently) and returns the statistics; `device()` then gives the distinct-k list as device
    tensors like output="device".  The tensors live in the peel workspace that the next transform of the same shape on this
    device reuses: read them (or queue the reads on the stream) before starting that one -- the statistics stay valid."""

    def __init__(self, prob, n, dist):
        import torch
        self._prob, self._n, self._dist = prob, n, dist
        self._host = torch.empty(8, dtype=torch.int64, pin_memory=True)
        self._host.copy_(prob.counters, non_blocking=True)          # stream ordered: the counters of THIS transform
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(prob.device))
        self._buffers = (prob.uniq_k, prob.uniq_sum, prob.uniq_cnt, prob.uniq_key, prob.max_uniq)
        self.stats = None

    def done(self):
        return self._event.query()

    def wait(self):
        if self.stats is None:
            self._event.synchronize()
            c = self._host.tolist()
            if c[6]:
                raise RuntimeError(f"find buffer too small: {c[0]} finds")
            if c[4] > self._buffers[4]:
                raise RuntimeError(f"unique buffer too small: {c[4]} > {self._buffers[4]}")
            if self._dist is not None:
                self._dist.verify()
            self.stats = {"rounds": int(c[5]), "finds": int(c[7]), "distinct": int(c[4])}
        return self.stats

    def device(self):
        nu = self.wait()["distinct"]
        uk, us, uc, ukey, _ = self._buffers
        return {"k": uk[:nu, :self._n], "sum": us[:nu], "count": uc[:nu], "key": ukey[:nu]}
