"""QSFT: the peeling decoder front end (drop-in for qsft/qsft.py:13-281).

transform() gathers (Ms, Ds, Us) from the signal exactly like the reference (so the NumPy RNG is consumed in the
same order), hands the bins to the on-device peeling loop (csrc/k4_peel.cu) and converts the list of finds into
the reference's result dict {tuple(k): complex}."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ops
from .input_signal_subsampled import SubsampledSignal
from .utils import calc_hamming_weight, sort_qary_vecs


class QSFT:
    """kwargs: num_subsample, num_repeat, b, reconstruct_method_source ("identity" | "coded"),
    reconstruct_method_channel ("identity" | "nso"), source_decoder (from get_reed_solomon_dec; needed for
    "coded"), noise_sd (accepted and ignored like the reference: the signal's noise_sd sets the threshold)."""

    def __init__(self, **kwargs):
        self.reconstruct_method_source = kwargs.get("reconstruct_method_source")
        self.reconstruct_method_channel = kwargs.get("reconstruct_method_channel")
        self.num_subsample = kwargs.get("num_subsample")
        self.num_repeat = kwargs.get("num_repeat")
        self.b = kwargs.get("b")
        self.source_decoder = kwargs.get("source_decoder", None)
        self.last_stats = {}

    def transform(self, signal, verbosity=0, report=False, timing_verbose=False, **kwargs):
        q, n, b = signal.q, signal.n, self.b
        if not isinstance(signal, SubsampledSignal):
            raise NotImplementedError("QSFT currently only supports signals that inherit from SubsampledSignal")
        Ms, Ds, Us, Ts = signal.get_MDU(self.num_subsample, self.num_repeat, b, trans_times=True)
        transform_time = float(np.sum(Ts))
        if timing_verbose:
            print(f"Transform Time:{transform_time}", flush=True)
        peeling_start = time.time()
        dev = signal.device
        # (C, P, B) bins; a private copy because peeling subtracts in place (the reference vstacks copies too)
        U = torch.stack([torch.cat([torch.as_tensor(u, device=dev) for u in us], dim=0) for us in Us]).contiguous()
        if U.dtype != torch.complex64:
            U = U.to(torch.complex64)
        D = np.stack([np.vstack(d) for d in Ds])
        C, P, B = U.shape
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
        prob = ops.PeelProblem(q, n, b, Ms, D, signal.get_source_parity(), channel, source, cutoff, dev, rs=rs)
        dist = getattr(signal, "dist", None)
        if dist is not None and dist.world_size > 1:
            from .dist import peel_sharded
            finds = peel_sharded(prob, U, dist)
            if kwargs.get("device_result", False):
                self.last_stats = {"rounds": int(finds[4]), "finds": int(len(finds[0])), "cutoff": float(cutoff)}
                return {"find_cj": finds[0], "find_k": finds[1], "find_rho": finds[2], "find_round": finds[3]}
        else:
            prob.alloc(max_finds=min(15 * C * B, max(4096, 4 * C * B)))
            n_finds, n_rounds = prob.peel(U)
            if kwargs.get("device_result", False):
                # raw finds left in HBM (no host copy, no dict): used by the device-resident benchmark loop
                self.last_stats = {"rounds": n_rounds, "finds": n_finds, "cutoff": float(cutoff)}
                return {"find_cj": prob.find_cj[:n_finds], "find_k": prob.find_k[:n_finds, :n],
                        "find_rho": prob.find_rho[:n_finds], "find_round": prob.find_round[:n_finds]}
            finds = (prob.find_cj[:n_finds].cpu().numpy(), prob.find_k[:n_finds, :n].cpu().numpy(),
                     prob.find_rho[:n_finds].cpu().numpy(), prob.find_round[:n_finds].cpu().numpy(), n_rounds)
        gwht, loc_arr = self._finds_to_dict(*finds[:4])
        self.last_stats = {"rounds": int(finds[4]), "finds": int(len(finds[0])), "cutoff": float(cutoff)}
        peeling_time = time.time() - peeling_start
        if timing_verbose:
            print(f"Peeling Time:{peeling_time}", flush=True)
        if not report:
            return gwht
        n_samples = C * P * B
        if len(loc_arr) > 0:
            loc = [tuple(r) for r in loc_arr.tolist()]
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
        (result list order = round, then (i, j); averaging qsft.py:247-255)."""
        if len(cj) == 0:
            return {}, np.zeros((0, k.shape[1] if k.ndim == 2 else 0), dtype=np.int64)
        order = np.lexsort((cj, rnd))
        k, rho = k[order], rho[order].astype(np.complex128)
        # group equal k: pack the digits into two uint64 words (>= 64 bits each is plenty: n <= 128 digits of < 2^7
        # would not fit, so hash-free exact packing uses as many words as needed)
        n = k.shape[1]
        per_word = max(1, 64 // max(1, int(np.ceil(np.log2(max(2, int(k.max()) + 1))))))
        bits = 64 // per_word
        words = []
        for w0 in range(0, n, per_word):
            blk = k[:, w0:min(n, w0 + per_word)].astype(np.uint64)
            weights = (np.uint64(1) << (np.uint64(bits) * np.arange(blk.shape[1] - 1, -1, -1, dtype=np.uint64)))
            words.append(blk @ weights)
        srt = np.lexsort(tuple(words[::-1]))
        same = np.ones(len(k), dtype=bool)
        for w in words:
            ws = w[srt]
            same[1:] &= ws[1:] == ws[:-1]
        same[0] = False
        gid_sorted = np.cumsum(~same) - 1                  # group id along the sorted order
        inv = np.empty(len(k), dtype=np.int64)
        inv[srt] = gid_sorted
        ngroups = int(gid_sorted[-1]) + 1
        sums = np.zeros(ngroups, dtype=np.complex128)
        np.add.at(sums, inv, rho)
        cnt = np.bincount(inv, minlength=ngroups)
        mean = sums / cnt
        first = np.full(ngroups, len(k), dtype=np.int64)
        np.minimum.at(first, inv, np.arange(len(k)))
        seen = np.argsort(first, kind="stable")
        keys = k[first[seen]].astype(np.int64)
        vals = mean[seen]
        gwht = dict(zip(map(tuple, keys.tolist()), vals.tolist()))
        return gwht, keys
