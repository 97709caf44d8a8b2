"""densecap_b200: B200-native (sm_100a) implementation of the DenseCap inference hot path.

The compute lives in ``libdensecap_b200.so`` (hand-written CUDA, flat C ABI declared in
``include/densecap_b200.h``); this package is the thin host-side mirror of the
reference's operator API (DenseCapModel / nn.LocalizationLayer / box_utils).
There is no CPU fallback: every operator raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
