"""JAX / Flax call signatures of the reference (`edward2.jax.nn.*`) on libed2b200 (filled in by nn/*.py)."""
