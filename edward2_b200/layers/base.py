"""Keras-shaped base class for the `ed.layers` mirrors: lazy `build`, `call`, `.losses`, `get_config`."""
from __future__ import annotations

import itertools

import torch
from torch import nn

_layer_uid = itertools.count(1)

_DTYPES = {
    None: torch.float32, "float32": torch.float32, "bfloat16": torch.bfloat16,
    torch.float32: torch.float32, torch.bfloat16: torch.bfloat16,
}

GOLDEN = 0x9E3779B97F4A7C15
MASK64 = (1 << 64) - 1


def resolve_dtype(dtype) -> torch.dtype:
    if dtype not in _DTYPES:
        raise ValueError(f"dtype must be float32 or bfloat16 (the compute dtypes of libed2b200), got {dtype!r}")
    return _DTYPES[dtype]


class Layer(nn.Module):
    """Mirrors the parts of tf.keras.layers.Layer the reference layers rely on."""

    def __init__(self, dtype=None, name=None, seed=None, trainable=True, **kwargs):
        super().__init__()
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        self.compute_dtype = resolve_dtype(dtype)
        self.uid = next(_layer_uid)
        self.layer_name = name or f"{type(self).__name__.lower()}_{self.uid}"
        # Philox key of this layer's native sampler; the per-call offset is folded into the key
        self.seed = int(seed) if seed is not None else (0xED2B200 ^ (self.uid * GOLDEN)) & MASK64
        self.built = False
        self.trainable = trainable
        self._calls = 0
        self._injected = {}
        self._extra_losses = []
        self.step_counter = None  # optional int64 device scalar advanced by the caller once per step
        # Data-parallel placement of this process's batch rows (edward2_b200.dist.shard_noise sets it on every layer):
        # (rank, world).  Per-example noise is then drawn for the GLOBAL row, so the sharded run reproduces the
        # single-device global-batch draw.
        self.data_shard = None

    # -- Keras surface -------------------------------------------------------------------------------------
    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        raise NotImplementedError

    def forward(self, inputs, *args, **kwargs):
        if not self.built:
            shape = inputs.shape if hasattr(inputs, "shape") else None
            self._device = inputs.device if hasattr(inputs, "device") else None
            self.build(tuple(shape))
        out = self.call(inputs, *args, **kwargs)
        self._calls += 1
        return out

    @property
    def losses(self):
        return list(self._extra_losses)

    def get_config(self):
        return {"name": self.layer_name, "dtype": str(self.compute_dtype).replace("torch.", "")}

    # -- randomness ----------------------------------------------------------------------------------------
    def inject(self, **noise):
        """Inject noise tensors by name (parity tests): the next calls use them instead of the Philox stream."""
        self._injected = dict(noise)
        return self

    def _row_map(self, per_example):
        """ed2_noise row mapping of a per-example noise tensor whose local groups hold `per_example` rows each (one
        group = one ensemble member; ungrouped layers pass the local batch): with data_shard = (rank, world) this
        process holds rows [rank * per_example, (rank + 1) * per_example) of every group of the global batch."""
        if per_example is None or self.data_shard is None:
            return None
        rank, world = self.data_shard
        if world <= 1:
            return None
        return (rank * per_example, per_example, per_example * world)

    def _noise(self, name: str, stream: int, per_example: int | None = None):
        from .. import functional as F

        if name in self._injected and self._injected[name] is not None:
            return F.Randomness(self._injected[name])
        rows = self._row_map(per_example)
        if self.step_counter is not None:  # device-side step: fresh noise on every CUDA-graph replay
            return F.Randomness(seed=self.seed, stream=stream, step=self.step_counter, rows=rows)
        key = (self.seed + self._calls * GOLDEN) & MASK64
        return F.Randomness(seed=key, stream=stream, rows=rows)


def flatten_batch(x: torch.Tensor):
    """[..., features] -> ([rows, features], leading shape)."""
    lead = x.shape[:-1]
    return x.reshape(-1, x.shape[-1]), lead
