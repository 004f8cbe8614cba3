"""Import shim that makes the UNMODIFIED reference at /root/reference importable in this container.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Used by oracle/gen_golden.py to produce the committed
fixtures under tests/golden/.  /root/reference does not exist on the GPU box, so nothing at test/bench time
imports this module.

What it patches:
  * numpy >= 1.24 removed the aliases np.complex / np.int that the reference still uses
    (qsft/input_signal_subsampled.py:136, synt_exp/synt_src/synthetic_signal.py:34,129, qsft/utils.py:58)
  * the reference imports group_lasso / matplotlib / galois at module top level (qsft/utils.py:6,16,
    qsft/ReedSolomon.py:1-2); none are installed -> empty stub modules.  The Reed-Solomon ("coded") path
    therefore cannot be run through the reference here: parity for it is UNPINNED.
"""
import sys
import types

import numpy as np

import os

REFERENCE_ROOT = "/root/reference"
# the unmodified copy made by oracle/make_ref.py (git-ignored; what exists on the GPU box)
REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """/root/reference where it exists (this container), else the copy under oracle/_ref, else None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "qsft")):
        return REFERENCE_ROOT
    if os.path.isdir(os.path.join(REF_COPY, "qsft")):
        return REF_COPY
    return None


def install(root=None):
    if not hasattr(np, "complex"):
        np.complex = complex
    if not hasattr(np, "int"):
        np.int = int

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("group_lasso", GroupLasso=object)
    stub("group_lasso._fista", ConvergenceWarning=Warning)
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    try:
        import galois  # noqa: F401
    except Exception:
        g = stub("galois", ReedSolomon=object, GF=lambda *a, **k: None)
        g._codes = stub("galois._codes")
        g._codes._reed_solomon = stub("galois._codes._reed_solomon", decode_jit=lambda *a, **k: None)
    root = root or reference_root() or REFERENCE_ROOT
    if root not in sys.path:
        sys.path.insert(0, root)
