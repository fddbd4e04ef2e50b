"""spadot_b200 — B200-native (sm_100a) implementation of SpaDOT's optimal-transport hot path.

Python host code + hand-written CUDA behind a C ABI (include/spadot_b200.h).  No CPU fallback.
"""
__version__ = "0.1.0"
