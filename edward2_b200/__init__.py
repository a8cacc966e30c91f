"""edward2_b200 — B200-native (sm_100a) implementation of Edward2's stochastic-weight dense layers and SNGP head.

`edward2_b200.layers` mirrors `ed.layers` (TF/Keras signatures), `edward2_b200.nn` mirrors `edward2.jax.nn`.
All arithmetic runs in libed2b200.so (hand-written CUDA behind the C ABI of include/ed2b200.h).
"""
__version__ = "0.1.0"

from . import constraints, initializers, layers, nn, regularizers  # noqa: E402,F401
