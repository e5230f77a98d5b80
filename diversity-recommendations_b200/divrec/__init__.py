"""divrec (B200-native hot path) -- the import surface of amtsyplov/diversity-recommendations for the
score -> mask -> top-k -> ILD / coverage path and the BPR step, with the arithmetic done by
libdivrec_b200.so (hand-written sm_100a CUDA behind the C ABI in include/divrec_b200.h).

There is no CPU path and no PyTorch fallback: the classes construct anywhere (so state_dicts load and
`.to("cuda")` works), but computing with CPU tensors raises TypeError and a missing shared library
raises ImportError at first use.
"""
__version__ = "0.1.0"
