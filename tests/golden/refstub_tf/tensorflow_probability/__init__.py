"""`tensorflow_probability`: the two distributions the reference's trainable-normal weights are made of
(tfp.distributions.Normal, Independent), restated on NumPy: sample = loc + scale * eps with eps from the stand-in
tensorflow's `random._rng` (recorded by the fixture generator), mean, stddev and the closed-form Normal-Normal KL
(TFP `_kl_normal_normal`: 0.5 (mu_a - mu_b)^2 / s_b^2 + 0.5 expm1(2 log(s_a / s_b)) - log(s_a / s_b))."""
import sys
import types

import numpy as np
import tensorflow as tf


class Distribution:
    pass


class Normal(Distribution):
    def __init__(self, loc, scale, **kwargs):
        self.loc, self.scale = np.asarray(tf._t(loc)), np.asarray(tf._t(scale))

    @property
    def batch_shape(self):
        return tf.TensorShape(np.broadcast(self.loc, self.scale).shape)

    event_shape = tf.TensorShape(())

    def mean(self):
        return tf._t(np.broadcast_to(self.loc, self.batch_shape))

    def stddev(self):
        return tf._t(np.broadcast_to(self.scale, self.batch_shape))

    def sample(self, sample_shape=(), seed=None):
        lead = tuple(np.atleast_1d(np.asarray(sample_shape, dtype=np.int64)).tolist())
        shape = lead + tuple(self.batch_shape)
        # TensorFlow: the same (global seed, op seed) pair reproduces the same values — the reference relies on it to
        # evaluate the predictive variance on the samples of the predictive mean (heteroscedastic.py:159-160, 228-253)
        rng = tf.random._rng if seed is None else np.random.default_rng([int(seed)] + [int(d) for d in shape])
        eps = rng.standard_normal(shape)
        recorded.append(eps)
        return tf._t(self.loc + self.scale * eps)

    def kl_divergence(self, other):
        ratio = self.scale / other.scale
        return tf._t(0.5 * ((self.loc - other.loc) / other.scale) ** 2 + 0.5 * np.expm1(2.0 * np.log(ratio)) - np.log(ratio))


class Independent(Distribution):
    def __init__(self, distribution, reinterpreted_batch_ndims=None, **kwargs):
        self.distribution = distribution
        n = len(distribution.batch_shape) if reinterpreted_batch_ndims is None else int(reinterpreted_batch_ndims)
        self.reinterpreted_batch_ndims = n

    @property
    def event_shape(self):
        bs = self.distribution.batch_shape
        return tf.TensorShape(bs[len(bs) - self.reinterpreted_batch_ndims:])

    @property
    def batch_shape(self):
        bs = self.distribution.batch_shape
        return tf.TensorShape(bs[:len(bs) - self.reinterpreted_batch_ndims])

    def mean(self):
        return self.distribution.mean()

    def stddev(self):
        return self.distribution.stddev()

    def sample(self, sample_shape=(), seed=None):
        return self.distribution.sample(sample_shape, seed)

    def kl_divergence(self, other):
        kl = np.asarray(self.distribution.kl_divergence(other.distribution))
        axes = tuple(range(kl.ndim - self.reinterpreted_batch_ndims, kl.ndim))
        return tf._t(kl.sum(axis=axes))


recorded = []  # standard-normal draws of every Normal.sample call, in call order (read by the fixture generator)
distributions = types.ModuleType("tensorflow_probability.distributions")
class Logistic(Distribution):   # named as default arguments by heteroscedastic.py; not sampled by the fixtures
    pass


class Gumbel(Distribution):
    pass


distributions.__dict__.update(Distribution=Distribution, Normal=Normal, Independent=Independent, Logistic=Logistic,
                              Gumbel=Gumbel)
sys.modules["tensorflow_probability.distributions"] = distributions
