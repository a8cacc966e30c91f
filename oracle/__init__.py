"""CPU oracle: a float64 NumPy restatement of the reference's arithmetic for the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `edward2_b200/` may import this package; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` do.

PARITY STATUS: **parity unpinned for sampled values**.  The reference (google/edward2 0.0.3) is pure Python over
TensorFlow / TFP / JAX / Flax, none of which is installed here or on the GPU box, so it cannot be executed to
produce golden vectors, and its own tests hold no golden vectors or RNG-stream pins for this path
(SURVEY.md §8c).  What *is* pinned:
  * the Philox4x32-10 generator against the three Random123 known-answer vectors (tests/golden/philox_kat.json);
  * `mean_field_logits` against the reference's known-answer tests (edward2/tensorflow/layers/utils_test.py:201-241,
    edward2/jax/nn/utils_test.py:28-68);
  * every algebraic identity the reference's tests assert (vectorised == per-member loop, exact / momentum
    precision updates, (1+ridge)·I on orthogonal data, diag == diag(full), spectral norm -> c, KL >= 0 and linear
    in scale_factor, softplus(-3) initial scale) — restated in tests/test_oracle_*.py;
  * closed-form gradients against torch.autograd in float64.
Every function cites the reference file:line it follows; all randomness is an explicit argument.
"""
