"""SyntheticHelper: TestHelper over synthetic sparse signals (mirror of synt_exp/synt_src/synthetic_helper.py)."""
from .synthetic_signal import SyntheticSubsampledSignal
from .test_helper import TestHelper


class SyntheticHelper(TestHelper):
    def generate_signal(self, signal_args):
        return SyntheticSubsampledSignal(**signal_args)
