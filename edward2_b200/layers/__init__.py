"""TF / Keras call signatures of the reference (`ed.layers.*`) on libed2b200."""
from . import utils
from .base import Layer
from .dense import Dense, DenseBatchEnsemble, DenseFlipout, DenseRank1, DenseReparameterization
from .normalization import SpectralNormalization
from .random_feature import LaplaceRandomFeatureCovariance, RandomFeatureGaussianProcess

__all__ = [
    "Layer", "Dense", "DenseBatchEnsemble", "DenseFlipout", "DenseRank1", "DenseReparameterization",
    "SpectralNormalization", "LaplaceRandomFeatureCovariance", "RandomFeatureGaussianProcess", "utils",
]
