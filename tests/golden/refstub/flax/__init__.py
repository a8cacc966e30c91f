"""NumPy-backed stand-in for the slice of `flax` the reference's edward2/jax/nn files use (see ../README.md)."""
from . import core, linen  # noqa: F401
