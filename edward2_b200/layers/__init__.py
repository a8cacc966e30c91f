"""TF / Keras call signatures of the reference (`ed.layers.*`) on libed2b200."""
from . import utils
from .base import Layer
from .convolutional import Conv2DBatchEnsemble, Conv2DFlipout, Conv2DRank1, Conv2DReparameterization
from .heteroscedastic import HeteroscedasticSNGPLayer, MCSigmoidDenseFA, MCSoftmaxDenseFA
from .dense import Dense, DenseBatchEnsemble, DenseFlipout, DenseRank1, DenseReparameterization
from .normalization import SpectralNormalization
from .recurrent import LSTMCellFlipout, LSTMCellRank1, LSTMCellReparameterization
from .random_feature import LaplaceRandomFeatureCovariance, RandomFeatureGaussianProcess

__all__ = [
    "Layer", "Conv2DBatchEnsemble", "Conv2DFlipout", "Conv2DRank1", "Conv2DReparameterization", "Dense", "DenseBatchEnsemble", "DenseFlipout", "DenseRank1", "DenseReparameterization", "HeteroscedasticSNGPLayer", "LSTMCellFlipout", "LSTMCellRank1", "LSTMCellReparameterization", "MCSigmoidDenseFA", "MCSoftmaxDenseFA",
    "SpectralNormalization", "LaplaceRandomFeatureCovariance", "RandomFeatureGaussianProcess", "utils",
]
