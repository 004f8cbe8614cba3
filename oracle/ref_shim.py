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

REFERENCE_ROOT = "/root/reference"


def install():
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
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
