"""Singleton detectors as public functions (drop-in for qsft/reconstruct.py:12-168).

Inside QSFT.transform detection is fused into the classification kernel (csrc/k4_peel.cu); this module keeps the
reference's function-level API callable for users who drive the detectors themselves.  Every function accepts one
column `U_slice` (P,) like the reference, or a batch (P, N) of columns, as NumPy arrays or CUDA tensors; the
arithmetic runs on the GPU through qsft_singleton_detect / qsft_detect_mle.  There is no CPU path."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _cols(U_slice, device=None):
    """(P,) or (P, N) array / tensor -> ((N, P) complex64 CUDA tensor, was_single)."""
    if isinstance(U_slice, torch.Tensor):
        t = U_slice
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(U_slice, dtype=np.complex64)))
    single = t.dim() == 1
    if single:
        t = t[:, None]
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("qsft_b200.reconstruct needs a CUDA device (no CPU fallback)")
        t = t.to(device or "cuda")
    return t.to(torch.complex64).t().contiguous(), single


def _ret(k, single):
    k = k.t().cpu().numpy().astype(int)               # (n, N) like the oracle / the reference's per-column (n,)
    return k[:, 0] if single else k


def singleton_detection_noiseless(U_slice, **kwargs):
    """reconstruct.py:12-31."""
    cols, single = _cols(U_slice)
    return _ret(ops.singleton_detect(cols, kwargs.get("q"), cols.shape[1], "identity"), single)


def singleton_detection_nso(U_slice, **kwargs):
    """reconstruct.py:87-97 (nso_subtype "nso1" soft / "nso2" hard decision)."""
    cols, single = _cols(U_slice)
    return _ret(ops.singleton_detect(cols, kwargs.get("q"), kwargs.get("source_parity"), "nso",
                                     nso_subtype=kwargs.get("nso_subtype", "nso1")), single)


def singleton_detection_nso1(U_slice, **kwargs):
    return singleton_detection_nso(U_slice, **{**kwargs, "nso_subtype": "nso1"})


def singleton_detection_nso2(U_slice, **kwargs):
    return singleton_detection_nso(U_slice, **{**kwargs, "nso_subtype": "nso2"})


def singleton_detection_mle(U_slice, **kwargs):
    """reconstruct.py:54-84: returns (selection[k_sel], S_slice[:, k_sel]) for one column; for a batch (P, N) the same
    pair with a leading / trailing batch axis (selection[k_sel] (N,), S_slice[:, k_sel] (P, N))."""
    selection, S_slice = kwargs.get("selection"), kwargs.get("S_slice")
    cols, single = _cols(U_slice)
    S = torch.from_numpy(np.ascontiguousarray(np.asarray(S_slice, dtype=np.complex64))).to(cols.device) \
        if not isinstance(S_slice, torch.Tensor) else S_slice.to(cols.device, torch.complex64).contiguous()
    k_sel, _ = ops.detect_mle(cols, S)
    k_sel = k_sel.cpu().numpy()
    sel = np.asarray(selection)[k_sel]
    sig = (S_slice.cpu().numpy() if isinstance(S_slice, torch.Tensor) else np.asarray(S_slice))[:, k_sel]
    return (sel[0], sig[:, 0]) if single else (sel, sig)


def singleton_detection_coded(k, **kwargs):
    """reconstruct.py:34-51: syndrome symbols -> decoded k through the host decoder returned by get_reed_solomon_dec."""
    decoder = kwargs.get("source_decoder")
    dec = decoder(list(k))
    return np.array(dec[0][0, :], dtype=np.int32)


def singleton_detection(U_slice, method_source="identity", method_channel="identity", **kwargs):
    """reconstruct.py:132-168: channel stage {"mle", "nso", "identity"} then source stage {"identity", "coded"}.
    With method_source="coded" the Reed-Solomon decode also runs on the GPU (source_decoder must come from
    qsft_b200.get_reed_solomon_dec)."""
    if method_channel == "mle":
        k = singleton_detection_mle(U_slice, **kwargs)
        if method_source != "identity":
            raise NotImplementedError("mle returns (selection entry, signature), not syndrome symbols")
        return k
    if method_channel not in ("nso", "identity"):
        raise TypeError(f"unknown method_channel {method_channel!r}")   # the reference fails with 'NoneType' is not callable
    if method_source == "identity":
        fn = singleton_detection_nso if method_channel == "nso" else singleton_detection_noiseless
        return fn(U_slice, **kwargs)
    if method_source != "coded":
        raise TypeError(f"unknown method_source {method_source!r}")
    rs = getattr(kwargs.get("source_decoder"), "__self__", None)
    if rs is None or not hasattr(rs, "device_tables"):
        raise ValueError("method_source='coded' needs source_decoder=get_reed_solomon_dec(n, t, q)")
    cols, single = _cols(U_slice)
    p1 = kwargs.get("source_parity") if method_channel == "nso" else cols.shape[1]
    k = ops.singleton_detect(cols, kwargs.get("q"), p1, method_channel, source="coded", n=rs.ns, rs=rs,
                             nso_subtype=kwargs.get("nso_subtype", "nso1"))
    return _ret(k, single).astype(np.int32)
