"""`flax.linen`: a minimal Module system (dataclass fields, setup / @compact, param / variable collections, init /
apply with `mutable`), Dense, LayerNorm and the initializers the reference names.  Plumbing only: the arithmetic that
runs inside a reference module is the reference's own code."""
import dataclasses
import functools
import types
import zlib
from typing import Any, Optional

import numpy as np

from . import core as _core

# ------------------------------------------------------------------------------------------------- initializers


def _rng(key):
    return np.random.default_rng([int(k) for k in np.asarray(key).reshape(-1)])


def _zeros(key, shape, dtype=None):
    return np.zeros(tuple(shape))


def _ones(key, shape, dtype=None):
    return np.ones(tuple(shape))


def _normal(stddev=1e-2):
    return lambda key, shape, dtype=None: _rng(key).standard_normal(tuple(shape)) * stddev


def _uniform(scale=1e-2):
    return lambda key, shape, dtype=None: _rng(key).uniform(0.0, scale, tuple(shape))


def _variance_scaling(scale, mode, distribution, in_axis=-2, out_axis=-1):
    def init(key, shape, dtype=None):
        shape = tuple(shape)
        fan_in, fan_out = shape[in_axis], shape[out_axis]
        denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
        var = scale / denom
        g = _rng(key)
        if distribution == "normal":
            return g.standard_normal(shape) * np.sqrt(var)
        if distribution == "truncated_normal":
            z = g.standard_normal(shape)
            while (bad := np.abs(z) > 2).any():  # rejection sampling of N(0,1) truncated to [-2, 2]
                z[bad] = g.standard_normal(int(bad.sum()))
            return z * np.sqrt(var) / 0.87962566103423978
        if distribution == "uniform":
            lim = np.sqrt(3 * var)
            return g.uniform(-lim, lim, shape)
        raise ValueError(distribution)

    return init


initializers = types.SimpleNamespace(
    zeros=_zeros, ones=_ones, normal=_normal, uniform=_uniform, variance_scaling=_variance_scaling,
    lecun_normal=lambda: _variance_scaling(1.0, "fan_in", "truncated_normal"),
    he_normal=lambda: _variance_scaling(2.0, "fan_in", "truncated_normal"),
    zeros_init=lambda: _zeros, ones_init=lambda: _ones,
)


# flax.linen.linear.default_kernel_init (heteroscedastic_lib.py:90)
linear = types.SimpleNamespace(default_kernel_init=initializers.lecun_normal())


def relu(x):
    return np.maximum(x, 0.0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))


def softplus(x):
    return np.logaddexp(0.0, np.asarray(x, dtype=np.float64))


def _promote_dtype(*args, dtype=None, inexact=True):
    return [None if a is None else np.asarray(a, dtype=np.float64) for a in args]


dtypes = types.SimpleNamespace(promote_dtype=_promote_dtype)

# ------------------------------------------------------------------------------------------------- Module system


class _Scope:
    def __init__(self, variables, mutable, rngs):
        self.variables = variables          # {collection: nested dict by module path}
        self.mutable = mutable              # True (all) or a set of collection names
        self.rngs = dict(rngs or {})
        self.rng_count = 0

    def is_mutable(self, col):
        return self.mutable is True or col in self.mutable

    def node(self, col, path, create):
        d = self.variables.get(col)
        if d is None:
            if not create:
                return None
            d = self.variables.setdefault(col, {})
        for p in path:
            nxt = d.get(p)
            if nxt is None:
                if not create:
                    return None
                nxt = d.setdefault(p, {})
            d = nxt
        return d


class Variable:
    def __init__(self, scope, col, path, name):
        self._scope, self._col, self._path, self._name = scope, col, path, name

    @property
    def value(self):
        return self._scope.node(self._col, self._path, False)[self._name]

    @value.setter
    def value(self, v):
        if not self._scope.is_mutable(self._col):
            raise RuntimeError(f"collection {self._col!r} is immutable")
        self._scope.node(self._col, self._path, True)[self._name] = v


_stack = []  # modules whose setup / compact method is currently running


def compact(fn):
    fn._compact = True
    return fn


def _wrap_call(fn):
    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        if self._scope is None:
            raise RuntimeError(f"{type(self).__name__} is not bound: use .init / .apply")
        self._ensure_setup()
        _stack.append(self)
        self._auto = {}
        try:
            return fn(self, *args, **kwargs)
        finally:
            _stack.pop()

    return wrapped


class Module:
    name: Optional[str] = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        ann = dict(cls.__dict__.get("__annotations__", {}))
        ann.pop("name", None)
        ann["name"] = Optional[str]
        cls.__annotations__ = ann
        cls.name = dataclasses.field(default=None, kw_only=True)
        if "__call__" in cls.__dict__:
            cls.__call__ = _wrap_call(cls.__dict__["__call__"])
        dataclasses.dataclass(cls, eq=False, repr=False)

    def __post_init__(self):
        object.__setattr__(self, "_scope", None)
        object.__setattr__(self, "_path", ())
        object.__setattr__(self, "_setup_done", False)
        object.__setattr__(self, "_in_setup", False)
        object.__setattr__(self, "_auto", {})
        if _stack:  # created inside a parent's setup / compact method: becomes its child
            parent = _stack[-1]
            if not parent._in_setup:
                name = self.name
                if name is None:
                    i = parent._auto.get(type(self).__name__, 0)
                    parent._auto[type(self).__name__] = i + 1
                    name = f"{type(self).__name__}_{i}"
                self._bind(parent._scope, parent._path + (name,))

    def __setattr__(self, k, v):
        if isinstance(v, Module) and getattr(self, "_in_setup", False):
            v._bind(self._scope, self._path + (v.name or k,))
        object.__setattr__(self, k, v)

    # -- binding -----------------------------------------------------------------------------------------
    def _bind(self, scope, path):
        object.__setattr__(self, "_scope", scope)
        object.__setattr__(self, "_path", tuple(path))

    def _ensure_setup(self):
        if not self._setup_done:
            object.__setattr__(self, "_setup_done", True)
            object.__setattr__(self, "_in_setup", True)
            _stack.append(self)
            try:
                self.setup()
            finally:
                _stack.pop()
                object.__setattr__(self, "_in_setup", False)

    def setup(self):
        pass

    def _clone_unbound(self):
        kw = {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.init}
        saved = list(_stack)
        _stack.clear()  # a top-level clone has no parent
        try:
            return type(self)(**kw)
        finally:
            _stack.extend(saved)

    def clone(self, **updates):
        kw = {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.init}
        kw.update(updates)
        saved = list(_stack)
        _stack.clear()
        try:
            return type(self)(**kw)
        finally:
            _stack.extend(saved)

    # -- public API --------------------------------------------------------------------------------------
    def init(self, rngs, *args, **kwargs):
        if not isinstance(rngs, dict):
            rngs = {"params": rngs}
        m = self._clone_unbound()
        scope = _Scope({}, True, rngs)
        m._bind(scope, ())
        saved = list(_stack)
        _stack.clear()
        try:
            m(*args, **kwargs)
        finally:
            _stack.extend(saved)
        return scope.variables

    def apply(self, variables, *args, mutable=False, rngs=None, method=None, **kwargs):
        if mutable is True:
            mut = True
        elif not mutable:
            mut = set()
        else:
            mut = {mutable} if isinstance(mutable, str) else set(mutable)
        import copy

        vars_ = copy.deepcopy(_core.unfreeze(variables))
        scope = _Scope(vars_, mut, rngs)
        m = self._clone_unbound()
        m._bind(scope, ())
        saved = list(_stack)
        _stack.clear()
        try:
            if method is None:
                out = m(*args, **kwargs)
            else:
                m._ensure_setup()
                _stack.append(m)
                try:
                    out = getattr(m, method if isinstance(method, str) else method.__name__)(*args, **kwargs)
                finally:
                    _stack.pop()
        finally:
            _stack.extend(saved)
        if mutable is False or mutable is None or mutable == [] or mutable == ():
            return out
        cols = vars_.keys() if mut is True else mut
        return out, {c: vars_[c] for c in cols if c in vars_}

    # -- inside a module ---------------------------------------------------------------------------------
    def is_mutable_collection(self, col):
        return self._scope.is_mutable(col)

    def is_initializing(self):
        return self._scope.mutable is True

    def make_rng(self, name):
        base = self._scope.rngs[name]
        self._scope.rng_count += 1
        # deterministic across processes (Python's str hash is salted): the fixture can be regenerated bit for bit
        mix = zlib.crc32(repr((self._path, self._scope.rng_count)).encode()) & 0xFFFFFFFF
        return np.array([int(np.asarray(base).reshape(-1)[-1]), mix], dtype=np.uint32)

    def has_variable(self, col, name):
        node = self._scope.node(col, self._path, False)
        return node is not None and name in node

    def param(self, name, init_fn, *init_args):
        node = self._scope.node("params", self._path, False)
        if node is not None and name in node:
            return node[name]
        if not self._scope.is_mutable("params"):
            raise KeyError(f"parameter {'/'.join(self._path + (name,))} is missing")
        val = init_fn(self.make_rng("params"), *init_args)
        self._scope.node("params", self._path, True)[name] = val
        return val

    def variable(self, col, name, init_fn=None, *init_args):
        node = self._scope.node(col, self._path, False)
        if node is None or name not in node:
            if not self._scope.is_mutable(col):
                raise KeyError(f"variable {col}/{'/'.join(self._path + (name,))} is missing")
            self._scope.node(col, self._path, True)[name] = init_fn(*init_args)
        return Variable(self._scope, col, self._path, name)

    def sow(self, col, name, value):
        if self._scope.is_mutable(col) and self._scope.mutable is not True:
            self._scope.node(col, self._path, True)[name] = (value,)
        return True


# ------------------------------------------------------------------------------------------------- stock layers


class Dense(Module):
    features: int
    use_bias: bool = True
    dtype: Any = None
    param_dtype: Any = None
    precision: Any = None
    kernel_init: Any = initializers.lecun_normal()
    bias_init: Any = _zeros

    @compact
    def __call__(self, inputs):
        inputs = np.asarray(inputs, dtype=np.float64)
        kernel = self.param("kernel", self.kernel_init, (inputs.shape[-1], self.features), np.float64)
        y = np.tensordot(inputs, np.asarray(kernel, dtype=np.float64), axes=((inputs.ndim - 1,), (0,)))
        if self.use_bias:
            bias = self.param("bias", self.bias_init, (self.features,), np.float64)
            y = y + np.asarray(bias, dtype=np.float64)
        return y


class LayerNorm(Module):
    epsilon: float = 1e-6
    dtype: Any = None
    use_bias: bool = True
    use_scale: bool = True
    bias_init: Any = _zeros
    scale_init: Any = _ones

    @compact
    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        mean = x.mean(-1, keepdims=True)
        var = np.maximum((x * x).mean(-1, keepdims=True) - mean * mean, 0.0)  # flax: use_fast_variance=True
        y = (x - mean) / np.sqrt(var + self.epsilon)
        if self.use_scale:
            y = y * self.param("scale", self.scale_init, (x.shape[-1],), np.float64)
        if self.use_bias:
            y = y + self.param("bias", self.bias_init, (x.shape[-1],), np.float64)
        return y
