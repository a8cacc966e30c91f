"""JAX / Flax call signatures of the reference (`edward2.jax.nn.*`) on libed2b200.

The JAX backend of the reference exports DenseBatchEnsemble, RandomFeatureGaussianProcess, RandomFourierFeatures,
LaplaceRandomFeatureCovariance, SpectralNormalization and utils (edward2/jax/nn/__init__.py:18-53); it has no
DenseRank1 / DenseFlipout / DenseReparameterization.  MCSoftmaxDenseFA mirrors
edward2/jax/nn/heteroscedastic_lib.py and MCSigmoidDenseFASNGP(BE) edward2/jax/nn/random_feature.py:407-735
(SURVEY.md §8f-2).
"""
from . import utils
from .dense import DenseBatchEnsemble
from .heteroscedastic import MCSigmoidDenseFA, MCSoftmaxDenseFA
from .module import Module
from .normalization import Dense, SpectralNormalization
from .random_feature import (LaplaceRandomFeatureCovariance, MCSigmoidDenseFASNGP, MCSigmoidDenseFASNGPBE,
                             RandomFeatureGaussianProcess, RandomFourierFeatures)
from .utils import make_sign_initializer, mean_field_logits

__all__ = ["Module", "Dense", "DenseBatchEnsemble", "SpectralNormalization", "LaplaceRandomFeatureCovariance",
           "RandomFeatureGaussianProcess", "RandomFourierFeatures", "MCSigmoidDenseFA", "MCSigmoidDenseFASNGP", "MCSigmoidDenseFASNGPBE", "MCSoftmaxDenseFA", "make_sign_initializer", "mean_field_logits", "utils"]
