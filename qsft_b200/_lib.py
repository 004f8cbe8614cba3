"""ctypes binding of libqsft_b200.so (the C ABI declared in include/qsft_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqsft_b200.so")

EXPORTS = [
    "qsft_last_error", "qsft_version", "qsft_launch_count", "qsft_reset_launch_count", "qsft_host_pack_digits",
    "qsft_query_lattice", "qsft_dec_to_qary", "qsft_qary_to_dec", "qsft_eval_synth", "qsft_gwht_batch", "qsft_gwht_batch_bcast", "qsft_gwht_batch_mcast", "qsft_gwht_batch_scatter",
    "qsft_eval_lattice_supported", "qsft_eval_synth_lattice", "qsft_eval_synth_lattice_ex",
    "qsft_peel_classify", "qsft_peel_apply", "qsft_peel_reduce", "qsft_peel_distinct", "qsft_peel", "qsft_peel_blocks", "qsft_peel_blocks_sharded", "qsft_peel_sharded_workspace_bytes", "qsft_closed_form_bins",
    "qsft_singleton_detect", "qsft_detect_mle", "qsft_k3_ticket_decode", "qsft_add_noise",
]


class QsftError(RuntimeError):
    pass


class Shard(C.Structure):
    """Mirror of qsft_shard."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("peers", C.POINTER(C.c_void_p)), ("epoch", C.c_uint32)]


class PeelDesc(C.Structure):
    """Mirror of qsft_peel_desc."""
    _fields_ = [
        ("q", C.c_int), ("n", C.c_int), ("b", C.c_int),
        ("C", C.c_int), ("P", C.c_int), ("P_src", C.c_int),
        ("channel", C.c_int), ("source", C.c_int),
        ("rs_t", C.c_int), ("rs_s", C.c_int),
        ("ld", C.c_int),
        ("cutoff", C.c_float),
        ("MT", C.c_void_p), ("D", C.c_void_p),
        ("rs_exp", C.c_void_p), ("rs_log", C.c_void_p),
    ]


class Uniq(C.Structure):
    """Mirror of qsft_uniq."""
    _fields_ = [("seen0", C.c_void_p), ("uniq_k", C.c_void_p), ("uniq_sum", C.c_void_p), ("uniq_cnt", C.c_void_p),
                ("uniq_key", C.c_void_p), ("uniq_next", C.c_void_p), ("max_uniq", C.c_int64)]


_lib = None


def build(verbose=False):
    """Compile the CUDA sources in-tree (nvcc, sm_100a).  Used by __graft_entry__.build()."""
    import subprocess
    src = os.path.join(_HERE, "csrc")
    res = subprocess.run(["make", "-C", src, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise QsftError("building libqsft_b200.so failed")
    return LIB_PATH


def lib():
    """Load (once) and return the shared library with argument types set."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QsftError(f"{LIB_PATH} not found: build it with `make -C qsft_b200/csrc` "
                        f"(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.qsft_last_error.restype = C.c_char_p
    L.qsft_last_error.argtypes = []
    L.qsft_version.restype = i32
    L.qsft_launch_count.restype = i64
    L.qsft_reset_launch_count.restype = None
    L.qsft_host_pack_digits.argtypes = [vp, i32, i64, i32, i64, i64, vp, i32, i32]
    L.qsft_query_lattice.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, vp]
    L.qsft_dec_to_qary.argtypes = [vp, i32, i64, i32, i32, vp, i32, vp]
    L.qsft_qary_to_dec.argtypes = [vp, i32, i64, i32, i32, vp, i32, vp]
    L.qsft_eval_synth.argtypes = [vp, i64, vp, vp, i64, i32, i32, i32, vp, i32, vp]
    L.qsft_gwht_batch.argtypes = [vp, i64, i32, i32, vp]
    L.qsft_gwht_batch_bcast.argtypes = [vp, i64, i32, i32, C.POINTER(vp), i32, vp]
    L.qsft_gwht_batch_mcast.argtypes = [vp, i64, i32, i32, vp, vp]
    L.qsft_gwht_batch_scatter.argtypes = [vp, i64, i32, i32, C.POINTER(vp), i32, i32, i64, vp]
    L.qsft_eval_lattice_supported.argtypes = [i32, i32, i32, i32, i64]
    L.qsft_eval_synth_lattice.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp, vp]
    L.qsft_eval_synth_lattice_ex.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp, i32, vp]
    pd = C.POINTER(PeelDesc)
    L.qsft_peel_classify.argtypes = [pd, vp, i64, i64, vp, vp, vp, vp, vp, i64, i32, vp, vp]
    L.qsft_peel_apply.argtypes = [pd, vp, i64, i64, vp, vp, vp, vp, i64, i64, i32, vp, vp]
    pu = C.POINTER(Uniq)
    L.qsft_peel_reduce.argtypes = [pd, vp, vp, vp, vp, i64, i64, i32, pu, vp, vp]
    L.qsft_peel_distinct.argtypes = [pu, vp, i64, i32, i32, vp, vp, vp, vp]
    L.qsft_peel.argtypes = [pd, vp, vp, vp, vp, vp, vp, i64, vp, pu, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32), vp]
    L.qsft_peel_blocks.argtypes = [pd, C.POINTER(vp), i64, vp, vp, vp, vp, vp, i64, vp, pu, C.POINTER(i64), C.POINTER(i64),
                                   C.POINTER(i32), vp]
    L.qsft_peel_sharded_workspace_bytes.argtypes = [pd, i64]
    L.qsft_peel_sharded_workspace_bytes.restype = i64
    L.qsft_peel_blocks_sharded.argtypes = [pd, C.POINTER(vp), i64, C.POINTER(Shard), i64, vp, pu, C.POINTER(i64), C.POINTER(i32), vp]
    L.qsft_closed_form_bins.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, vp, i64, vp, vp]
    L.qsft_singleton_detect.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp]
    L.qsft_detect_mle.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp]
    L.qsft_k3_ticket_decode.argtypes = [C.c_uint32, i64, i32, i32, i32, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    L.qsft_add_noise.argtypes = [vp, i64, C.c_float, C.c_uint64, C.c_uint64, vp]
    for name in EXPORTS:
        fn = getattr(L, name)  # raises AttributeError if a declared symbol is missing
        if name not in ("qsft_last_error", "qsft_version", "qsft_launch_count", "qsft_reset_launch_count", "qsft_host_pack_digits",
                        "qsft_peel_sharded_workspace_bytes"):
            fn.restype = i32
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().qsft_last_error().decode("utf-8", "replace")
        raise QsftError(f"libqsft_b200 error {rc}: {msg}")


def launch_count():
    return int(lib().qsft_launch_count())


def reset_launch_count():
    lib().qsft_reset_launch_count()
