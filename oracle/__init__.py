"""CPU oracle: a float64 NumPy restatement of the reference's arithmetic for the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `edward2_b200/` may import this package; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` do.

PARITY STATUS: **pinned to the reference for the deterministic arithmetic; unpinned for sampled VALUES** (the reference
holds no RNG-stream pins for this path).  The reference (google/edward2 0.0.3) is pure Python over TensorFlow / TFP /
JAX / Flax, none of which is installed here or on the GPU box.  Its JAX files are pure `jnp` / `lax.linalg` code, so
`tests/golden/make_reference_golden.py` EXECUTES THEM UNMODIFIED under NumPy-backed stand-ins of jax / flax
(`tests/golden/refstub/`) and commits their outputs as `tests/golden/reference_vectors.npz`:
  * `edward2/jax/nn/dense.py` DenseBatchEnsemble, `random_feature.py` RandomFourierFeatures /
    LaplaceRandomFeatureCovariance / RandomFeatureGaussianProcess / MCSigmoidDenseFASNGP(BE), `normalization.py`
    SpectralNormalization, `utils.py` mean_field_logits, `heteroscedastic_lib.py` MCSoftmaxDenseFA / MCSigmoidDenseFA
    (with the draws of their `jax.random.normal` calls recorded) — `tests/test_oracle.py::test_reference_*` checks this
    oracle against those vectors at 1e-12, `tests/test_gpu_reference_golden.py` checks the CUDA path against them.
  * `edward2/tensorflow/layers/dense.py` (with `layers/utils.py`, `initializers.py`, `regularizers.py`, `constraints.py`)
    is executed the same way under NumPy-backed tensorflow / tfp stand-ins (`tests/golden/refstub_tf/`,
    `tests/golden/make_reference_golden_tf.py` -> `reference_vectors_tf.npz`): DenseReparameterization, DenseFlipout
    (with its Uniform[-1, 3) sign_output), DenseBatchEnsemble rank 1 / rank 3 and DenseRank1 (clip bounds, additive
    perturbations, sampled bias) with every draw recorded, plus TrainableNormal / Softplus / NormalKLDivergence; and
    `layers/normalization.py` SpectralNormalization (in-place TF form), `layers/random_feature.py`
    RandomFeatureGaussianProcess / LaplaceRandomFeatureCovariance (momentum and exact updates, three likelihoods, the
    covariance inverse), `layers/utils.py` mean_field_logits — `tests/test_oracle.py::test_reference_tf_*` checks
    `oracle/dense.py` and `oracle/sngp.py` against them at 1e-12 (1e-10 through the matrix inverse).
    `layers/convolutional.py` (Conv2DReparameterization / Flipout / BatchEnsemble / Rank1) and `layers/recurrent.py`
    (LSTMCellReparameterization / Flipout, Keras implementations 1 and 2, and LSTMCellRank1, multiplicative and additive)
    run the same way and pin `oracle/conv.py` and `oracle/recurrent.py`; `layers/heteroscedastic.py` (MCSoftmaxDenseFA,
    MCSigmoidDenseFA: one diagonal draw per class) and `layers/hetsngp.py` (HeteroscedasticSNGPLayer, training and
    evaluation calls) pin the TF form of `oracle/heteroscedastic.py`.
For the VALUES of the native samplers (the reference has no RNG-stream contract) what is pinned is:
  * the Philox4x32-10 generator against the three Random123 known-answer vectors (tests/golden/philox_kat.json);
  * `mean_field_logits` against the reference's known-answer tests (edward2/tensorflow/layers/utils_test.py:201-241,
    edward2/jax/nn/utils_test.py:28-68);
  * every algebraic identity the reference's tests assert (vectorised == per-member loop, exact / momentum
    precision updates, (1+ridge)·I on orthogonal data, diag == diag(full), spectral norm -> c, KL >= 0 and linear
    in scale_factor, softplus(-3) initial scale, DenseRank1 with deterministic fast weights == DenseBatchEnsemble);
  * closed-form gradients against torch.autograd in float64.
Every function cites the reference file:line it follows; all randomness is an explicit argument.
"""
