"""`tensorflow`: a NumPy-backed stand-in (float64) with just the surface that the reference files

    edward2/tensorflow/layers/dense.py, layers/utils.py, initializers.py, regularizers.py, constraints.py

touch when DenseReparameterization / DenseFlipout / DenseBatchEnsemble / DenseRank1 are built and called.  TensorFlow is
not installable here; the stand-in exists so that tests/golden/make_reference_golden_tf.py can EXECUTE THOSE FILES
UNMODIFIED and record their outputs (see ../README.md).  Third-party behaviour restated here, and only here:
tf.keras.layers.Dense (`matmul(inputs, kernel)`, tensordot for rank > 2, bias_add, activation), the Keras initializers
that the reference falls back to by name, and the op semantics of reshape / tile / transpose / clip_by_value / ...
(their NumPy equivalents).  Randomness comes from `random._rng` (a seeded numpy Generator); tf.random.uniform keeps
TensorFlow's dtype-dependent behaviour — integer dtypes draw integers in [minval, maxval), float dtypes draw reals in
[minval, maxval) — which is what gives DenseFlipout.sign_output its Uniform[-1, 3) values (dense.py:266-270)."""
import contextlib
import sys
import types

import numpy as np

float32, float64, int32, int64 = "float32", "float64", "int32", "int64"
newaxis = None


class TensorShape(tuple):
    def __new__(cls, dims=()):
        if isinstance(dims, TensorShape):
            return dims
        return super().__new__(cls, tuple(int(d) for d in dims))

    @property
    def ndims(self):
        return len(self)

    rank = ndims

    def as_list(self):
        return list(self)

    def __getitem__(self, i):
        out = tuple.__getitem__(self, i)
        return TensorShape(out) if isinstance(i, slice) else out


class Tensor(np.ndarray):
    """ndarray whose `.shape` is a TensorShape (the reference reads `inputs.shape.ndims`); `assign` makes it double as a
    tf.Variable (in place, so aliases such as SpectralNormalization's `self.w = self.layer.kernel` keep aliasing)."""

    @property
    def shape(self):
        return TensorShape(np.ndarray.shape.__get__(self))

    def assign(self, value):
        self[...] = np.asarray(_t(value))
        return self


def _t(x, dtype=None):
    if hasattr(x, "value") and hasattr(x, "distribution"):  # an ed.RandomVariable stand-in
        x = x.value
    a = np.asarray(x)
    if dtype is not None and "bool" in str(dtype):
        kind = np.bool_
    elif dtype is not None and "int" in str(dtype):
        kind = np.int64
    elif dtype is None and a.dtype.kind in "iub":
        kind = a.dtype
    else:
        kind = np.float64
    return np.array(a, dtype=kind).view(Tensor)


convert_to_tensor = _t
bool = "bool"  # noqa: A001


class DType:
    pass


def constant(value, dtype=None, **kwargs):
    return _t(value, dtype)


def Variable(initial_value=None, dtype=None, trainable=True, name=None, **kwargs):  # noqa: N802
    return _t(initial_value, dtype)


class VariableSynchronization:
    AUTO, ON_READ, ON_WRITE, NONE = "auto", "on_read", "on_write", "none"


class VariableAggregation:
    NONE, SUM, MEAN, ONLY_FIRST_REPLICA = "none", "sum", "mean", "only_first_replica"


def cond(pred, true_fn, false_fn):
    return true_fn() if np.asarray(pred).item() else false_fn()


def eye(n, dtype=None, **kwargs):
    return _t(np.eye(int(n)))


def ones(shape, dtype=None):
    return _t(np.ones([int(d) for d in shape]))


def sigmoid(x):
    return _t(1.0 / (1.0 + np.exp(-_t(x))))


def is_tensor(x):
    return isinstance(x, np.ndarray)


def get_static_value(x):
    return np.asarray(x)


def shape(x):
    return np.array(np.shape(np.asarray(_t(x))), dtype=np.int64)


def cast(x, dtype):
    return _t(x, dtype)


def reshape(x, shp):
    return _t(np.reshape(_t(x), [int(s) for s in np.asarray(shp).tolist()]))


def tile(x, multiples):
    return _t(np.tile(_t(x), [int(m) for m in np.asarray(multiples).tolist()]))


def expand_dims(x, axis):
    return _t(np.expand_dims(_t(x), axis))


def transpose(x, perm=None):
    return _t(np.transpose(_t(x), perm))


def clip_by_value(x, lo, hi):
    return _t(np.clip(_t(x), lo, hi))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    return _t(np.matmul(a.T if transpose_a else a, b.T if transpose_b else b))


def tensordot(a, b, axes):
    return _t(np.tensordot(_t(a), _t(b), axes))


def concat(values, axis):
    return np.concatenate([np.atleast_1d(np.asarray(v)) for v in values], axis)


def stack(values, axis=0):
    return _t(np.stack([_t(v) for v in values], axis))


def zeros(shp, dtype=None):
    return _t(np.zeros([int(s) for s in np.asarray(shp).tolist()]))


def zeros_like(x):
    return _t(np.zeros_like(_t(x)))


def broadcast_to(x, shp):
    return _t(np.broadcast_to(np.asarray(x, dtype=np.float64), tuple(int(s) for s in shp)))


def reduce_sum(x, axis=None, keepdims=False):
    return _t(np.sum(_t(x), axis=axis, keepdims=keepdims))


def reduce_mean(x, axis=None, keepdims=False):
    return _t(np.mean(_t(x), axis=axis, keepdims=keepdims))


def square(x):
    return _t(np.square(_t(x)))


def sqrt(x):
    return _t(np.sqrt(_t(x)))


def exp(x):
    return _t(np.exp(_t(x)))


def abs(x):  # noqa: A001
    return _t(np.abs(_t(x)))


def maximum(a, b):
    return _t(np.maximum(_t(a), b))


def multiply(a, b):
    return _t(_t(a) * _t(b))


@contextlib.contextmanager
def name_scope(name):
    yield


def einsum(equation, *operands):
    return _t(np.einsum(equation, *[np.asarray(_t(o)) for o in operands]))


def executing_eagerly():
    return True


def zeros_initializer():
    return _Zeros()


def squeeze(x, axis=None):
    return _t(np.squeeze(_t(x), axis=axis))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ---- tf.math / tf.nn / tf.linalg / tf.random
def _softplus(x):
    x = _t(x)
    return _t(np.where(x > 30, x, np.log1p(np.exp(np.minimum(x, 30)))))


math = _module("tensorflow.math", log=lambda x: _t(np.log(_t(x))), reduce_mean=reduce_mean, softplus=_softplus,
               cos=lambda x: _t(np.cos(_t(x))), sigmoid=lambda x: sigmoid(x), exp=lambda x: exp(x))
nn = _module("tensorflow.nn", relu=lambda x: _t(np.maximum(_t(x), 0.0)), softplus=_softplus,
             bias_add=lambda v, b, data_format=None: _t(_t(v) + _t(b)),
             softmax=lambda x, axis=-1: _t(np.exp(_t(x) - np.max(_t(x), axis, keepdims=True))
                                           / np.sum(np.exp(_t(x) - np.max(_t(x), axis, keepdims=True)), axis, keepdims=True)))
nn.l2_normalize = lambda x, axis=None, epsilon=1e-12: _t(_t(x) / np.sqrt(np.maximum(np.sum(np.square(_t(x)), axis=axis, keepdims=True), epsilon)))  # noqa: E731,E501
linalg = _module("tensorflow.linalg", matmul=matmul, inv=lambda a: _t(np.linalg.inv(_t(a))),
                 diag_part=lambda a: _t(np.diagonal(_t(a)).copy()))


def _uniform(shape, minval=0, maxval=None, dtype=float32, seed=None):
    shp = tuple(int(s) for s in np.asarray(shape).tolist())
    if "int" in str(dtype):
        return random._rng.integers(minval, maxval, size=shp).view(Tensor)
    return _t(random._rng.uniform(minval, 1.0 if maxval is None else maxval, size=shp))


def _normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None):
    shp = tuple(int(s) for s in np.asarray(shape).tolist())
    return _t(mean + stddev * random._rng.standard_normal(shp))


random = _module("tensorflow.random", uniform=_uniform, normal=_normal, _rng=np.random.default_rng(0),
                 set_seed=lambda seed: None)
dtypes = _module("tensorflow.dtypes", DType=DType)


# ---- tf.keras
class _Initializer:
    def get_config(self):
        return {}


class _Zeros(_Initializer):
    def __call__(self, shape, dtype=None, **kw):
        return zeros(shape)


class _Ones(_Initializer):
    def __call__(self, shape, dtype=None, **kw):
        return _t(np.ones([int(s) for s in shape]))


class _TruncatedNormal(_Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None, **kw):
        shp = tuple(int(s) for s in shape)
        z = random._rng.standard_normal(shp)
        while np.any(np.abs(z) > 2.0):  # Keras: redraw beyond two standard deviations
            bad = np.abs(z) > 2.0
            z[bad] = random._rng.standard_normal(int(bad.sum()))
        return _t(self.mean + self.stddev * z)


class _VarianceScaling(_Initializer):
    def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None):
        self.scale, self.mode, self.distribution, self.seed = scale, mode, distribution, seed

    def __call__(self, shape, dtype=None, **kw):
        shp = tuple(int(s) for s in shape)
        fan_in, fan_out = (shp[-2], shp[-1]) if len(shp) >= 2 else (shp[0], shp[0])
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[self.mode]
        if self.distribution == "uniform":
            lim = np.sqrt(3.0 * self.scale / n)
            return _t(random._rng.uniform(-lim, lim, shp))
        return _t(np.sqrt(self.scale / n) * random._rng.standard_normal(shp))


class _Orthogonal(_Initializer):
    """tf.keras.initializers.Orthogonal: QR of a normal matrix, sign-fixed by diag(R), transposed when rows < cols."""

    def __init__(self, gain=1.0, seed=None):
        self.gain, self.seed = gain, seed

    def __call__(self, shape, dtype=None, **kw):
        shp = tuple(int(d) for d in shape)
        rows, cols = int(np.prod(shp[:-1])), shp[-1]
        a = random._rng.standard_normal((max(rows, cols), min(rows, cols)))
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diagonal(r))
        if rows < cols:
            q = q.T
        return _t(self.gain * q.reshape(shp))

    def get_config(self):
        return {"gain": self.gain, "seed": self.seed}


class _RandomUniform(_Initializer):
    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval, self.seed = minval, maxval, seed

    def __call__(self, shape, dtype=None, **kw):
        return _t(random._rng.uniform(self.minval, self.maxval, tuple(int(d) for d in shape)))


class _RandomNormal(_Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None, **kw):
        return _t(self.mean + self.stddev * random._rng.standard_normal(tuple(int(d) for d in shape)))


random_uniform_initializer = _RandomUniform
initializers = _module("tensorflow.initializers", random_normal=_RandomNormal)


def _get_initializer(identifier):
    if identifier is None or callable(identifier):
        return identifier
    table = {"zeros": _Zeros, "zero": _Zeros, "ones": _Ones,
             "glorot_uniform": lambda: _VarianceScaling(1.0, "fan_avg", "uniform"),
             "he_normal": lambda: _VarianceScaling(2.0, "fan_in", "truncated_normal"),
             "glorot_normal": lambda: _VarianceScaling(1.0, "fan_avg", "truncated_normal"),
             "orthogonal": lambda: _Orthogonal()}
    return table[str(identifier)]()


class _Regularizer:
    pass


class _Constraint:
    pass


def _get_plain(identifier):
    if identifier is None or callable(identifier):
        return identifier
    raise ValueError(f"stand-in keras has no object named {identifier!r}")


def _activation(identifier):
    if identifier is None or identifier == "linear":
        return None if identifier is None else (lambda x: x)
    if callable(identifier):
        return identifier
    return {"relu": nn.relu}[identifier]


def _serialize(obj):
    return None if obj is None else {"class_name": type(obj).__name__, "config": obj.get_config() if hasattr(obj, "get_config") else {}}


def _deserialize(config, module_objects=None, custom_objects=None, printable_module_name="object"):
    name = config["class_name"]
    cls = (module_objects or {}).get(name)
    if cls is None or not callable(cls):
        raise ValueError(f"Unknown {printable_module_name}: {name}")
    return cls(**config.get("config", {}))


class Layer:
    def __init__(self, trainable=True, name=None, dtype=None, **kwargs):
        self.trainable, self.name, self.dtype, self.built = trainable, name, dtype or float32, False
        self.compute_dtype = self.dtype
        self._loss_fns = []

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True,
                   constraint=None, **kwargs):
        if initializer is None:  # Keras default: glorot_uniform for floating dtypes, zeros otherwise
            initializer = "zeros" if (dtype is not None and ("bool" in str(dtype) or "int" in str(dtype))) or len(shape) < 1 \
                else "glorot_uniform"
        w = _t(_get_initializer(initializer)(shape, dtype), dtype if dtype is not None and "bool" in str(dtype) else None)
        if regularizer is not None:
            self._loss_fns.append(lambda: regularizer(w))
        return w

    def add_loss(self, fn):
        self._loss_fns.append(fn)

    def add_update(self, op):
        pass

    def compute_output_shape(self, input_shape):
        return TensorShape(input_shape)

    @property
    def losses(self):
        return [f() if callable(f) else f for f in self._loss_fns]

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            self.build(_t(inputs).shape)
            self.built = True
        return self.call(inputs, *args, **kwargs)

    def get_config(self):
        return {"name": self.name, "dtype": self.dtype}


class Dense(Layer):
    """tf.keras.layers.Dense, restated: kernel [in, units], bias [units]; call = matmul / tensordot, bias_add, activation."""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self.units, self.activation, self.use_bias = int(units), _activation(activation), use_bias
        self.kernel_initializer, self.bias_initializer = _get_initializer(kernel_initializer), _get_initializer(bias_initializer)
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.activity_regularizer = activity_regularizer
        self.kernel_constraint, self.bias_constraint = kernel_constraint, bias_constraint

    def build(self, input_shape):
        last = int(TensorShape(input_shape)[-1])
        self.kernel = self.add_weight("kernel", shape=[last, self.units], initializer=self.kernel_initializer,
                                      regularizer=self.kernel_regularizer, constraint=self.kernel_constraint,
                                      dtype=self.dtype, trainable=True)
        self.bias = self.add_weight("bias", shape=[self.units], initializer=self.bias_initializer,
                                    regularizer=self.bias_regularizer, constraint=self.bias_constraint, dtype=self.dtype,
                                    trainable=True) if self.use_bias else None
        self.built = True

    def call(self, inputs):
        x = _t(inputs)
        out = matmul(x, self.kernel) if x.ndim == 2 else tensordot(x, self.kernel, [[x.ndim - 1], [0]])
        if self.use_bias:
            out = nn.bias_add(out, self.bias)
        return self.activation(out) if self.activation is not None else out


class LayerNormalization(Layer):
    """tf.keras.layers.LayerNormalization defaults: axis -1, epsilon 1e-3, trainable gamma (ones) / beta (zeros)."""

    def __init__(self, axis=-1, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        d = int(TensorShape(input_shape)[-1])
        self.gamma, self.beta = _t(np.ones(d)), _t(np.zeros(d))
        self.built = True

    def call(self, inputs):
        x = _t(inputs)
        mean = x.mean(-1, keepdims=True)
        var = np.square(x - mean).mean(-1, keepdims=True)
        return _t((x - mean) / np.sqrt(var + self.epsilon) * self.gamma + self.beta)


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super().__init__(**kwargs)
        self.function = function

    def call(self, inputs):
        return self.function(inputs)


class Wrapper(Layer):
    def __init__(self, layer, **kwargs):
        super().__init__(**kwargs)
        self.layer = layer

    def build(self, input_shape=None):
        if not self.layer.built:
            self.layer.build(input_shape)
            self.layer.built = True
        self.built = True


def _dense_output_shape(self, input_shape):
    return TensorShape(tuple(TensorShape(input_shape)[:-1]) + (self.units,))


Dense.compute_output_shape = _dense_output_shape


def _conv2d_nhwc(x, k, strides, padding):
    """Direct NHWC cross-correlation (what tf.nn.convolution computes): 'VALID' or TensorFlow's 'SAME' padding (extra
    padding at the bottom / right)."""
    x, k = np.asarray(_t(x)), np.asarray(_t(k))
    kh, kw, cin, cout = k.shape
    sh, sw = strides
    B, H, W, _ = x.shape
    if str(padding).upper() == "SAME":
        ho, wo = -(-H // sh), -(-W // sw)
        ph, pw = max((ho - 1) * sh + kh - H, 0), max((wo - 1) * sw + kw - W, 0)
        x = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
    else:
        ho, wo = (H - kh) // sh + 1, (W - kw) // sw + 1
    out = np.zeros((B, ho, wo, cout))
    for i in range(kh):
        for j in range(kw):
            patch = x[:, i:i + (ho - 1) * sh + 1:sh, j:j + (wo - 1) * sw + 1:sw, :]
            out += patch @ k[i, j]
    return _t(out)


def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else tuple(int(a) for a in v)


nn.convolution = lambda input, filters, strides=None, padding="VALID", data_format=None, dilations=None: _conv2d_nhwc(  # noqa: E731,A002
    input, filters, _pair(1 if strides is None else strides), padding)


class Conv2D(Layer):
    """tf.keras.layers.Conv2D, restated for channels_last / dilation 1: kernel [kh, kw, Cin, filters], bias [filters];
    call = convolution_op(inputs, kernel), bias_add, activation."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", data_format=None, dilation_rate=(1, 1),
                 groups=1, activation=None, use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self.rank, self.filters, self.kernel_size, self.strides = 2, int(filters), _pair(kernel_size), _pair(strides)
        self.padding, self.data_format = padding, data_format or "channels_last"
        self.dilation_rate, self.groups = _pair(dilation_rate), groups
        self.activation, self.use_bias = _activation(activation), use_bias
        self.kernel_initializer, self.bias_initializer = _get_initializer(kernel_initializer), _get_initializer(bias_initializer)
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.activity_regularizer = activity_regularizer
        self.kernel_constraint, self.bias_constraint = kernel_constraint, bias_constraint

    def build(self, input_shape):
        cin = int(TensorShape(input_shape)[-1])
        self.kernel = self.add_weight(name="kernel", shape=self.kernel_size + (cin, self.filters),
                                      initializer=self.kernel_initializer, regularizer=self.kernel_regularizer,
                                      constraint=self.kernel_constraint, trainable=True, dtype=self.dtype)
        self.bias = self.add_weight(name="bias", shape=(self.filters,), initializer=self.bias_initializer,
                                    regularizer=self.bias_regularizer, constraint=self.bias_constraint, trainable=True,
                                    dtype=self.dtype) if self.use_bias else None
        self.built = True

    def convolution_op(self, inputs, kernel):
        return _conv2d_nhwc(inputs, kernel, self.strides, self.padding)

    def call(self, inputs):
        out = self.convolution_op(inputs, self.kernel)
        if self.use_bias:
            out = nn.bias_add(out, self.bias)
        return self.activation(out) if self.activation is not None else out


class LSTMCell(Layer):
    """tf.keras.layers.LSTMCell, restated (dropout 0): kernel [in, 4 units], recurrent_kernel [units, 4 units], gate order
    i | f | c | o; implementation 1 works on the four gate slices through `_compute_carry_and_output`, implementation 2 on
    the fused pre-activation through `_compute_carry_and_output_fused`; h = o * activation(c)."""

    def __init__(self, units, activation="tanh", recurrent_activation="sigmoid", use_bias=True,
                 kernel_initializer="glorot_uniform", recurrent_initializer="orthogonal", bias_initializer="zeros",
                 unit_forget_bias=True, kernel_regularizer=None, recurrent_regularizer=None, bias_regularizer=None,
                 kernel_constraint=None, recurrent_constraint=None, bias_constraint=None, dropout=0.0,
                 recurrent_dropout=0.0, implementation=2, **kwargs):
        super().__init__(**kwargs)
        acts = {"tanh": lambda x: _t(np.tanh(_t(x))), "sigmoid": sigmoid}
        self.units = int(units)
        self.activation = acts.get(activation, activation)
        self.recurrent_activation = acts.get(recurrent_activation, recurrent_activation)
        self.use_bias, self.unit_forget_bias = use_bias, unit_forget_bias
        self.kernel_initializer, self.recurrent_initializer = kernel_initializer, recurrent_initializer
        self.bias_initializer = _get_initializer(bias_initializer)
        self.kernel_regularizer, self.recurrent_regularizer, self.bias_regularizer = (kernel_regularizer,
                                                                                      recurrent_regularizer, bias_regularizer)
        self.kernel_constraint, self.recurrent_constraint, self.bias_constraint = (kernel_constraint, recurrent_constraint,
                                                                                   bias_constraint)
        self.dropout, self.recurrent_dropout, self.implementation = dropout, recurrent_dropout, implementation
        self.state_size = [self.units, self.units]

    def get_dropout_mask_for_cell(self, inputs, training, count=1):
        return None

    def get_recurrent_dropout_mask_for_cell(self, inputs, training, count=1):
        return None

    def get_initial_state(self, inputs=None, batch_size=None, dtype=None):
        if inputs is not None:
            batch_size = np.shape(inputs)[0]
        return [zeros([batch_size, self.units]), zeros([batch_size, self.units])]

    def _compute_carry_and_output(self, x, h_tm1, c_tm1):
        x_i, x_f, x_c, x_o = x
        h_i, h_f, h_c, h_o = h_tm1
        u, rk = self.units, _t(self.recurrent_kernel)
        i = self.recurrent_activation(x_i + matmul(h_i, rk[:, :u]))
        f = self.recurrent_activation(x_f + matmul(h_f, rk[:, u:2 * u]))
        c = f * c_tm1 + i * self.activation(x_c + matmul(h_c, rk[:, 2 * u:3 * u]))
        o = self.recurrent_activation(x_o + matmul(h_o, rk[:, 3 * u:]))
        return c, o

    def _compute_carry_and_output_fused(self, z, c_tm1):
        z0, z1, z2, z3 = z
        i, f = self.recurrent_activation(z0), self.recurrent_activation(z1)
        c = f * c_tm1 + i * self.activation(z2)
        return c, self.recurrent_activation(z3)

    def call(self, inputs, states, training=None):
        h_tm1, c_tm1 = states[0], states[1]
        if self.implementation == 1:
            k = split(_t(self.kernel), 4, axis=1)
            x = [matmul(inputs, ki) for ki in k]
            if self.use_bias:
                x = [nn.bias_add(xi, bi) for xi, bi in zip(x, split(_t(self.bias), 4, axis=0))]
            c, o = self._compute_carry_and_output(tuple(x), (h_tm1,) * 4, c_tm1)
        else:
            z = matmul(inputs, self.kernel) + matmul(h_tm1, self.recurrent_kernel)
            if self.use_bias:
                z = nn.bias_add(z, self.bias)
            c, o = self._compute_carry_and_output_fused(split(z, 4, axis=1), c_tm1)
        h = o * self.activation(c)
        return h, [h, c]

    def __call__(self, inputs, states, training=None):
        if not self.built:
            self.build(_t(inputs).shape)
            self.built = True
        return self.call(inputs, states, training)


def split(value, num_or_size_splits, axis=0):
    return [_t(a) for a in np.split(np.asarray(_t(value)), num_or_size_splits, axis=axis)]


class Conv1D(Layer):
    pass


class DepthwiseConv2D(Layer):
    pass


class InputSpec:
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class _L2(_Regularizer):
    def __init__(self, l2=0.01):
        self.l2 = l2

    def __call__(self, w):
        return self.l2 * np.sum(np.square(_t(w)))


def _layer_serialize(layer):
    return {"class_name": type(layer).__name__}


_layers = _module("tensorflow.keras.layers", serialize=_layer_serialize, Layer=Layer, Dense=Dense, LayerNormalization=LayerNormalization,
                  Lambda=Lambda, Wrapper=Wrapper, experimental=types.SimpleNamespace(), Conv2D=Conv2D, Conv1D=Conv1D,
                  DepthwiseConv2D=DepthwiseConv2D, InputSpec=InputSpec, LSTMCell=LSTMCell)
_initializers = _module("tensorflow.keras.initializers", Initializer=_Initializer, Zeros=_Zeros, Ones=_Ones,
                        TruncatedNormal=_TruncatedNormal, VarianceScaling=_VarianceScaling, Orthogonal=_Orthogonal,
                        get=_get_initializer, he_normal=lambda seed=None: _VarianceScaling(2.0, "fan_in", "truncated_normal"),
                        glorot_normal=lambda seed=None: _VarianceScaling(1.0, "fan_avg", "truncated_normal"),
                        GlorotNormal=lambda seed=None: _VarianceScaling(1.0, "fan_avg", "truncated_normal"),
                        glorot_uniform=lambda seed=None: _VarianceScaling(1.0, "fan_avg", "uniform"),
                        zeros=_Zeros, ones=_Ones, Constant=lambda value=0.0: (lambda shape, dtype=None, **kw: _t(np.full([int(s) for s in shape], value))))
_regularizers = _module("tensorflow.keras.regularizers", Regularizer=_Regularizer, get=_get_plain, l2=_L2, L2=_L2)
_constraints = _module("tensorflow.keras.constraints", Constraint=_Constraint, get=_get_plain)
_activations = _module("tensorflow.keras.activations", get=_activation, serialize=lambda a: getattr(a, "__name__", None),
                       relu=nn.relu, linear=lambda x: x)
_backend = _module("tensorflow.keras.backend", epsilon=lambda: 1e-7, learning_phase=lambda: 0, dot=matmul,
                  bias_add=lambda v, b: _t(_t(v) + _t(b)),
                  concatenate=lambda ts, axis=-1: _t(np.concatenate([np.asarray(_t(a)) for a in ts], axis)))
_legacy = _module("tensorflow.keras.utils.legacy", serialize_keras_object=_serialize, deserialize_keras_object=_deserialize)
_utils = _module("tensorflow.keras.utils", legacy=_legacy)
keras = _module("tensorflow.keras", layers=_layers, initializers=_initializers, regularizers=_regularizers,
                constraints=_constraints, activations=_activations, backend=_backend, utils=_utils)
class _Dimension:
    pass


_v1 = _module("tensorflow.compat.v1", Dimension=_Dimension)
sys.modules["tensorflow.compat.v2"] = sys.modules[__name__]   # `import tensorflow.compat.v2 as tf`
compat = _module("tensorflow.compat", v1=_v1, v2=sys.modules[__name__])
