"""qsft_b200 -- B200-native engine for the q-SFT transform path (drop-in for basics-lab/qsft's QSFT /
SubsampledSignal / query_args).  Python host code; all numerics run in libqsft_b200.so (hand-written sm_100a CUDA,
C ABI in include/qsft_b200.h).  There is no CPU fallback."""
from ._lib import QsftError, lib, build  # noqa: F401
from .qsft import QSFT  # noqa: F401
from .input_signal_subsampled import SubsampledSignal  # noqa: F401
from .synthetic_signal import (SyntheticSubsampledSignal, generate_signal_w,  # noqa: F401
                               get_random_subsampled_signal)
from .query import get_Ms, get_D, get_Ms_and_Ds, get_reed_solomon_dec  # noqa: F401
from .reed_solomon import ReedSolomon  # noqa: F401
from . import reconstruct  # noqa: F401
from .test_helper import TestHelper  # noqa: F401
from .synthetic_helper import SyntheticHelper  # noqa: F401
from .reconstruct import singleton_detection  # noqa: F401

__all__ = ["QSFT", "SubsampledSignal", "SyntheticSubsampledSignal", "generate_signal_w",
           "get_random_subsampled_signal", "get_Ms", "get_D", "get_Ms_and_Ds", "get_reed_solomon_dec",
           "ReedSolomon", "reconstruct", "singleton_detection", "TestHelper", "SyntheticHelper", "QsftError", "lib", "build"]
