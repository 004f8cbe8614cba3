"""QSFT: the peeling decoder front end (drop-in for qsft/qsft.py:13-281).

transform() gathers (Ms, Ds, Us) from the signal exactly like the reference (so the NumPy RNG is consumed in the
same order), hands the bins to the on-device peeling loop (csrc/k4_peel.cu) and converts the list of finds into
the reference's result dict {tuple(k): complex}."""
from __future__ import annotations

import itertools
import time

import numpy as np
import torch

from . import _lib, ops
from .input_signal_subsampled import SubsampledSignal
from .result import PendingSpectrum, SparseSpectrum
from .utils import calc_hamming_weight, sort_qary_vecs


class QSFT:
    """kwargs: num_subsample, num_repeat, b, reconstruct_method_source ("identity" | "coded"),
    reconstruct_method_channel ("identity" | "nso"), source_decoder (from get_reed_solomon_dec; needed for
    "coded"), noise_sd (accepted and ignored like the reference: the signal's noise_sd sets the threshold).
    Extension: nso_subtype ("nso1" default = what the reference hard-codes at qsft.py:171; "nso2" = the hard-decision
    detector reconstruct.py:116-129)."""

    def __init__(self, **kwargs):
        self.reconstruct_method_source = kwargs.get("reconstruct_method_source")
        self.reconstruct_method_channel = kwargs.get("reconstruct_method_channel")
        self.num_subsample = kwargs.get("num_subsample")
        self.num_repeat = kwargs.get("num_repeat")
        self.b = kwargs.get("b")
        self.source_decoder = kwargs.get("source_decoder", None)
        self.nso_subtype = kwargs.get("nso_subtype", "nso1")
        self.last_stats = {}

    def transform(self, signal, verbosity=0, report=False, timing_verbose=False, **kwargs):
        q, n, b = signal.q, signal.n, self.b
        if not isinstance(signal, SubsampledSignal):
            raise NotImplementedError("QSFT currently only supports signals that inherit from SubsampledSignal")
        # the transform seconds are CUDA-event times: asking for them waits for the device, so only when they are reported
        want_times = bool(report or timing_verbose)
        mdu = signal.get_MDU(self.num_subsample, self.num_repeat, b, trans_times=want_times)
        Ms, Ds, Us = mdu[0], mdu[1], mdu[2]
        transform_time = float(np.sum(mdu[3])) if want_times else 0.0
        if timing_verbose:
            print(f"Transform Time:{transform_time}", flush=True)
        peeling_start = time.time()
        dev = signal.device
        C, P, B = len(Us), sum(int(u.shape[0]) for u in Us[0]), int(Us[0][0].shape[-1])
        if any(sum(int(u.shape[0]) for u in us) != P for us in Us):
            raise ValueError("every subsampling group must carry the same number of delay rows")

        def stacked():
            # (C, P, B) bins in one array -- a private copy: the host-driven rounds subtract in place (the reference vstacks
            # copies too, qsft.py:115-121).  One copy straight into place: the bins are 1-3 GB at the large configurations.
            U = torch.empty((C, P, B), dtype=torch.complex64, device=dev)
            for i, us in enumerate(Us):
                row = 0
                for u in us:
                    u = torch.as_tensor(u, device=dev)
                    U[i, row:row + u.shape[0]].copy_(u)
                    row += u.shape[0]
            return U

        D = np.stack([np.vstack(d) for d in Ds])
        cutoff = 1e-9 + (1 + 0.5) * (signal.noise_sd ** 2) / (q ** b)   # noise threshold, qsft.py:124-125
        cutoff = kwargs.get("cutoff", cutoff)
        if verbosity >= 2:
            print("cutoff = ", cutoff, flush=True)
        source = self.reconstruct_method_source or "identity"
        channel = self.reconstruct_method_channel or "identity"
        rs = None
        if source == "coded":
            rs = getattr(self.source_decoder, "__self__", None)
            if rs is None or not hasattr(rs, "device_tables"):
                raise ValueError("reconstruct_method_source='coded' needs source_decoder=get_reed_solomon_dec(n, t, q)")
        prob = ops.PeelProblem(q, n, b, Ms, D, signal.get_source_parity(), channel, source, cutoff, dev, rs=rs,
                               nso_subtype=self.nso_subtype)
        dist = getattr(signal, "dist", None)
        shard = dist.shard_peel(C * P * B * 8) if dist is not None and dist.world_size > 1 else ""
        n_rounds = None
        output = kwargs.get("output", "dict")
        # output="device_async": nothing is read back -- the peel is queued behind the sampling and transform kernels and the
        # call returns a PendingSpectrum; a stream of transforms then keeps the GPU busy while the host prepares the next one
        wait = output != "device_async"
        # find list: 4 C B slots cover every case seen; the reference has no limit (up to C B singletons per round, 15 rounds),
        # so a peel that runs out of slots is repeated once with the true bound -- the on-device loop never modifies the bins
        # (asynchronous calls report the overflow from PendingSpectrum.wait() instead)
        small, full = min(15 * C * B, max(4096, 4 * C * B)), 15 * C * B
        small = min(full, int(kwargs.get("max_finds", small)))       # (kwarg: first allocation of the find list)

        def with_retry(run):
            try:
                prob.alloc(max_finds=small, max_uniq=max(4096, C * B), reuse=True)
                return run()
            except _lib.QsftError as exc:
                if "find buffer too small" not in str(exc) or small == full or not wait:
                    raise
                prob.alloc(max_finds=full, max_uniq=max(4096, C * B), reuse=True)
                return run()

        if shard == "device":
            blocks = [u if isinstance(u, torch.Tensor) else torch.as_tensor(u, device=dev) for us in Us for u in us]
            if all(t.is_cuda and t.dtype == torch.complex64 and t.is_contiguous() for t in blocks):
                n_rounds = with_retry(lambda: prob.peel_blocks_sharded(blocks, dist, wait=wait))
            else:
                prob.alloc(max_finds=small, max_uniq=max(4096, C * B), reuse=True)
            n_finds = -1
            if n_rounds is None:                         # shape / platform does not fit: every rank agrees (same inputs)
                if not getattr(signal, "Us_complete", True):
                    raise RuntimeError("the bins were scattered for the bin-sharded on-device peel, which does not take this "
                                       "shape; build the signal with DistContext(peel_mode='replicated' or 'sharded_host')")
                shard = "host" if dist.peel_mode == "sharded" else ""
        if shard == "host":
            from .dist import peel_sharded
            n_rounds = peel_sharded(prob, stacked(), dist)[4]
            n_finds = -1
        elif n_rounds is None:
            # the on-device loop reads the blocks get_MDU returned where they lie (no copy)
            blocks = [u if isinstance(u, torch.Tensor) else torch.as_tensor(u, device=dev) for us in Us for u in us]
            fits = all(t.is_cuda and t.dtype == torch.complex64 and t.is_contiguous() for t in blocks)
            prob.alloc(max_finds=small, max_uniq=max(4096, C * B), reuse=True)
            done = with_retry(lambda: prob.peel_blocks(blocks, wait=wait)) if fits else None
            if done is None and not wait:
                raise ValueError("output='device_async' needs bins the on-device loop can read in place (C * R <= 16, "
                                 "contiguous complex64 CUDA blocks)")
            n_finds, n_rounds = done if done is not None else prob.peel(stacked())
        if not wait:
            if shard == "host":
                raise ValueError("output='device_async' is not available with the host-driven sharded peel")
            self.last_stats = {"cutoff": float(cutoff)}
            return PendingSpectrum(prob, n, dist if dist is not None and dist.world_size > 1 else None)
        if dist is not None and dist.world_size > 1:
            dist.verify()                                 # deferred rank-agreement checks (Ms / Ds, selections, noise seed)
        self.last_stats = {"rounds": int(n_rounds), "finds": int(n_finds), "distinct": int(prob.n_uniq),
                           "cutoff": float(cutoff)}
        if output == "device":
            # distinct k left in HBM (no host copy): k int8 (K, n), sum of rho complex64, find counts, first-seen keys
            nu = prob.n_uniq
            return {"k": prob.uniq_k[:nu, :n], "sum": prob.uniq_sum[:nu], "count": prob.uniq_cnt[:nu],
                    "key": prob.uniq_key[:nu]}
        loc_arr, values, counts = prob.distinct()          # first-seen order, mean over duplicate finds (qsft.py:247-255)
        if output == "arrays":
            gwht = {"locations": loc_arr, "values": values, "counts": counts}
        else:
            # the reference's {tuple(k): complex}; the tuples are created on first use (result.py)
            gwht = SparseSpectrum(loc_arr, values)
        peeling_time = time.time() - peeling_start
        if timing_verbose:
            print(f"Peeling Time:{peeling_time}", flush=True)
        if not report:
            return gwht
        n_samples = C * P * B
        if len(loc_arr) > 0:
            loc = list(itertools.batched(loc_arr.astype(np.uint8).tobytes(), n))
            if kwargs.get("sort", False):
                loc = sort_qary_vecs(loc)
            hw = calc_hamming_weight(loc_arr)
            avg_hamming_weight, max_hamming_weight = np.mean(hw), np.max(hw)
        else:
            loc, avg_hamming_weight, max_hamming_weight = [], 0, 0
        return {
            "gwht": gwht,
            "runtime": transform_time + peeling_time,
            "n_samples": n_samples,
            "locations": loc,
            "avg_hamming_weight": avg_hamming_weight,
            "max_hamming_weight": max_hamming_weight,
        }

    @staticmethod
    def _finds_to_dict(cj, k, rho, rnd):
        """Finds -> {tuple(k): mean of all rho found for k} in the reference's first-seen order
        (result list order = round, then (i, j); averaging qsft.py:247-255).  k (finds, n) small non-negative ints."""
        nf = len(cj)
        n = k.shape[1] if k.ndim == 2 else 0
        if nf == 0:
            return {}, np.zeros((0, n), dtype=np.int64)
        order = np.lexsort((cj, rnd))
        k8 = np.ascontiguousarray(k[order], dtype=np.uint8)
        rho = rho[order].astype(np.complex128)
        # group equal rows: 64-bit words of the zero-padded digit rows, mixed into one hash; ties verified exactly
        nw = (n + 7) // 8
        padded = np.zeros((nf, nw * 8), dtype=np.uint8)
        padded[:, :n] = k8
        words = padded.view(np.uint64)                              # (finds, nw)
        mult = (np.arange(nw, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0xD6E8FEB86659FD93)) | np.uint64(1)
        with np.errstate(over="ignore"):
            h = np.bitwise_xor.reduce((words * mult) ^ ((words * mult) >> np.uint64(29)), axis=1)
        srt = np.argsort(h, kind="stable")
        ws = words[srt]
        same = np.zeros(nf, dtype=bool)
        same[1:] = np.all(ws[1:] == ws[:-1], axis=1)
        if np.any((h[srt][1:] == h[srt][:-1]) & ~same[1:]):          # hash collision between different k: exact fallback
            srt = np.lexsort(tuple(words[:, i] for i in range(nw - 1, -1, -1)))
            ws = words[srt]
            same[1:] = np.all(ws[1:] == ws[:-1], axis=1)
        gid_sorted = np.cumsum(~same) - 1
        inv = np.empty(nf, dtype=np.int64)
        inv[srt] = gid_sorted
        ngroups = int(gid_sorted[-1]) + 1
        sums = np.zeros(ngroups, dtype=np.complex128)
        np.add.at(sums, inv, rho)
        cnt = np.bincount(inv, minlength=ngroups)
        mean = sums / cnt
        first = np.full(ngroups, nf, dtype=np.int64)
        np.minimum.at(first, inv, np.arange(nf))
        seen = np.argsort(first, kind="stable")
        keys8 = k8[first[seen]]
        vals = mean[seen]
        gwht = dict(zip(itertools.batched(keys8.tobytes(), n), vals.tolist()))     # tuple(k) -> complex
        return gwht, keys8.astype(np.int64)
