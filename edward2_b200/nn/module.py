"""A minimal Flax-shaped functional module system: `init(rng, *args)` -> variables, `apply(variables, *args,
mutable=...)`; collections `params`, `random_features`, `laplace_covariance`, `spectral_stats`, `intermediates`.

`rng` is an int seed or a `torch.Generator` (the reference passes a jax PRNGKey); tensors are torch CUDA tensors and
gradients flow through torch autograd to the leaves stored in `variables['params']`.
"""
from __future__ import annotations

import contextlib
import copy

import torch


class PRNG:
    """Stand-in for a jax PRNGKey: deterministic, splittable by name."""

    def __init__(self, seed=0):
        if isinstance(seed, PRNG):
            seed = seed.seed
        self.seed = int(seed)

    def fold(self, name: str) -> "PRNG":
        h = self.seed
        for ch in str(name):
            h = (h * 1000003 + ord(ch)) % (2**31 - 1)
        return PRNG(h + 1)

    def generator(self) -> torch.Generator:
        return torch.Generator(device="cpu").manual_seed(self.seed)


class _Scope:
    def __init__(self, variables, mutable, initializing, rng, path=(), rngs=None):
        self.variables, self.mutable, self.initializing, self.rng, self.path = variables, mutable, initializing, rng, path
        self.rngs = rngs or {}  # named streams passed to init / apply (flax: rngs={'dropout': key, ...})

    def collection(self, col, create=False):
        d = self.variables.setdefault(col, {}) if create else self.variables.get(col, {})
        for p in self.path:
            if create:
                d = d.setdefault(p, {})
            else:
                d = d.get(p, {})
        return d

    def child(self, name):
        return _Scope(self.variables, self.mutable, self.initializing, self.rng.fold(name), self.path + (name,), self.rngs)


class Module:
    """Subclasses declare fields as class attributes with defaults and implement __call__."""

    name = None

    def __init__(self, *args, **kwargs):
        fields = [k for k, v in vars(type(self)).items()]
        anns = {}
        for klass in reversed(type(self).__mro__):
            anns.update(getattr(klass, "__annotations__", {}))
        names = [n for n in anns if n != "name"] + ["name"]
        if len(args) > len(names):
            raise TypeError("too many positional arguments")
        for n, a in zip(names, args):
            kwargs[n] = a
        for n in names:
            if n in kwargs:
                setattr(self, n, kwargs.pop(n))
            elif not hasattr(type(self), n):
                raise TypeError(f"missing required field {n!r}")
            else:
                # copy the class-level default onto the instance: a function default must stay a plain function
                # (looked up through the instance it would be bound as a method)
                setattr(self, n, type(self).__dict__.get(n, getattr(type(self), n)) if not callable(getattr(type(self), n))
                        else _unbound(type(self), n))
        if kwargs:
            raise TypeError(f"unexpected fields {sorted(kwargs)}")
        self._scope = None
        self.setup()

    def setup(self):
        pass

    # -- Flax surface ------------------------------------------------------------------------------------------
    def init(self, rngs, *args, device="cuda", **kwargs):
        rng = rngs.get("params", 0) if isinstance(rngs, dict) else rngs
        variables = {}
        named = {k: v for k, v in rngs.items() if k != "params"} if isinstance(rngs, dict) else {}
        with self._bind(_Scope(variables, mutable=True, initializing=True, rng=PRNG(rng), rngs=named), device):
            self(*args, **kwargs)
        variables.pop("intermediates", None)
        return variables

    def apply(self, variables, *args, mutable=False, rngs=None, **kwargs):
        if mutable is True:
            mut = True
        elif not mutable:
            mut = set()
        else:
            mut = set([mutable] if isinstance(mutable, str) else mutable)
        work = {k: _copy_tree(v) for k, v in variables.items()}
        dev = _first_device(variables)
        with self._bind(_Scope(work, mutable=mut, initializing=False, rng=PRNG(0), rngs=dict(rngs or {})), dev):
            out = self(*args, **kwargs)
        if mutable is False or mutable is None:
            return out
        cols = work.keys() if mut is True else mut
        return out, {c: work.get(c, {}) for c in cols}

    @contextlib.contextmanager
    def _bind(self, scope, device):
        prev, prev_dev = self._scope, getattr(self, "_device", None)
        self._scope, self._device = scope, device
        try:
            yield
        finally:
            self._scope, self._device = prev, prev_dev

    def sub(self, module: "Module", name: str):
        """Run a child module inside this module's scope under `name`."""
        module._scope = self._scope.child(name)
        module._device = self._device
        return module

    def is_mutable_collection(self, col: str) -> bool:
        m = self._scope.mutable
        return m is True or (m is not False and col in m)

    def is_initializing(self) -> bool:
        return self._scope.initializing

    def make_rng(self, name="params") -> PRNG:
        """A named stream passed as rngs={name: seed} to init / apply (folded with the module path), else derived from
        the module's own stream."""
        if name in self._scope.rngs:
            base = PRNG(self._scope.rngs[name])
            for p in self._scope.path:
                base = base.fold(p)
            return base.fold(name)
        return self._scope.rng.fold(name)

    def param(self, name, init_fn, *init_args):
        d = self._scope.collection("params", create=True)
        if name not in d:
            if not self._scope.initializing:
                raise KeyError(f"parameter {'/'.join(self._scope.path + (name,))} is missing from variables['params']")
            t = init_fn(self._scope.rng.fold(name), *init_args)
            d[name] = _to_device(t, self._device).requires_grad_(True)
        return d[name]

    def variable(self, col, name, init_fn, *init_args):
        d = self._scope.collection(col, create=True)
        if name not in d:
            t = init_fn(*init_args)
            d[name] = _to_device(t, self._device)
        return _Variable(d, name)

    def sow(self, col, name, value):
        if self._scope.initializing or self.is_mutable_collection(col):
            self._scope.collection(col, create=True)[name] = value


def _unbound(klass, name):
    for k in klass.__mro__:
        if name in k.__dict__:
            return k.__dict__[name]
    return getattr(klass, name)


class _Variable:
    def __init__(self, d, name):
        self._d, self._name = d, name

    @property
    def value(self):
        return self._d[self._name]

    @value.setter
    def value(self, v):
        self._d[self._name] = v


def _to_device(t, device):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.to(device=device).detach().clone()


def _copy_tree(v):
    if isinstance(v, dict):
        return {k: _copy_tree(x) for k, x in v.items()}
    return v


def _first_device(tree):
    if isinstance(tree, torch.Tensor):
        return tree.device
    if isinstance(tree, dict):
        for v in tree.values():
            d = _first_device(v)
            if d is not None:
                return d
    return None


# -- initializers with the flax signature init(key, shape, dtype) ------------------------------------------------
def ones(key, shape, dtype=None):
    return torch.ones(tuple(shape))


def zeros(key, shape, dtype=None):
    return torch.zeros(tuple(shape))


def normal(stddev=1e-2):
    def init(key, shape, dtype=None):
        return torch.randn(tuple(shape), generator=PRNG(key).generator()) * stddev
    return init


def uniform(scale=1e-2):
    def init(key, shape, dtype=None):
        return torch.rand(tuple(shape), generator=PRNG(key).generator()) * scale
    return init


def variance_scaling(scale, mode, distribution):
    def init(key, shape, dtype=None):
        fan_in, fan_out = shape[-2], shape[-1]
        denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
        var = scale / denom
        g = PRNG(key).generator()
        if distribution == "normal":
            return torch.randn(tuple(shape), generator=g) * var ** 0.5
        if distribution == "truncated_normal":
            t = torch.empty(tuple(shape))
            std = var ** 0.5 / 0.87962566103423978
            torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=g)
            return t
        lim = (3 * var) ** 0.5
        return (torch.rand(tuple(shape), generator=g) * 2 - 1) * lim
    return init


lecun_normal = lambda: variance_scaling(1.0, "fan_in", "truncated_normal")  # noqa: E731
