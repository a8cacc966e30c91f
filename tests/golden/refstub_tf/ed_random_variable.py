"""Stand-ins for `edward2.tensorflow.random_variable` and `edward2.tensorflow.generated_random_variables` (the reference's
own versions are built on TensorFlow's tensor-conversion registry and on TFP class introspection): an ed.RandomVariable is
a distribution plus ONE sampled value drawn at construction (random_variable.py:62-110) that behaves as that value in
arithmetic."""
import numpy as np
import tensorflow as tf
import tensorflow_probability as tfp


class RandomVariable:
    def __init__(self, distribution, sample_shape=(), value=None):
        self.distribution = distribution
        self.value = tf._t(distribution.sample(sample_shape)) if value is None else tf._t(value)

    @property
    def shape(self):
        return self.value.shape

    @property
    def dtype(self):
        return tf.float32

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def __add__(self, o): return self.value + tf._t(o)
    def __radd__(self, o): return tf._t(o) + self.value
    def __sub__(self, o): return self.value - tf._t(o)
    def __rsub__(self, o): return tf._t(o) - self.value
    def __mul__(self, o): return self.value * tf._t(o)
    def __rmul__(self, o): return tf._t(o) * self.value
    def __neg__(self): return -self.value
    def __getitem__(self, i): return self.value[i]


def Normal(loc, scale, **kwargs):  # noqa: N802
    return RandomVariable(tfp.distributions.Normal(loc, scale))


def Independent(distribution, reinterpreted_batch_ndims=None, **kwargs):  # noqa: N802
    return RandomVariable(tfp.distributions.Independent(distribution, reinterpreted_batch_ndims))
