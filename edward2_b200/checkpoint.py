"""Reference-compatible variable names for the layers of this path (the on-disk format either side of
the path).  A converter between a reference checkpoint and these layers is a rename: Keras nests variable names as
`<layer>/<sublayer>/<variable>:0`; trainable-normal weights are `<layer>/kernel/mean:0` and `<layer>/kernel/stddev:0`
(edward2/tensorflow/layers/utils_test.py:36-39), BatchEnsemble / rank-1 fast weights `alpha`, `gamma`, `ensemble_bias`
(dense.py:527-545), spectral-norm state `u`, `v` (normalization.py:340,348), the Laplace state `gp_precision_matrix`,
`gp_covariance_matrix`, `update_gp_covariance` (random_feature.py:346-378) and the RFGP sub-layers
`gp_input_normalization`, `gp_random_feature`, `gp_covariance`, `gp_output_weights`, `gp_output_bias`
(random_feature.py:169-200).  Shapes need no transposition: kernels are [in, out] on both sides.

Checkpoint file I/O (TF checkpoints, msgpack) stays out of scope — these functions map names <-> tensors only."""
from __future__ import annotations

import torch
from torch import nn

from .layers.base import Layer

_RENAMED = {
    "precision_matrix": "gp_precision_matrix",
    "covariance_matrix": "gp_covariance_matrix",
    "if_update_covariance": "update_gp_covariance",
    "_gp_output_bias": "gp_output_bias",
}


# constructor argument `<x>_initializer` -> the weight it initialises, where the two names differ
# (recurrent.py:100-112: `recurrent_initializer` creates the variable `recurrent_kernel`)
_WEIGHT_OF_INITIALIZER = {"recurrent": "recurrent_kernel"}


def _walk(module: nn.Module, path: str):
    for store in (module._parameters, module._buffers):
        for name, t in store.items():
            if t is not None:
                yield f"{path}/{_RENAMED.get(name, name)}:0", store, name
    for cname, child in module._modules.items():
        if child is None:
            continue
        if cname.endswith("_initializer"):  # TrainableNormal & co.: <layer>/kernel/mean:0
            weight = cname[:-len("_initializer")]
            sub = f"{path}/{_WEIGHT_OF_INITIALIZER.get(weight, weight)}"
        elif isinstance(child, Layer):
            sub = f"{path}/{child.layer_name}"
        else:
            sub = f"{path}/{cname}"
        yield from _walk(child, sub)


def reference_variables(layer: Layer, name: str | None = None) -> dict:
    """{reference variable name: tensor (detached, on its device)} of a built layer."""
    if not getattr(layer, "built", False):
        raise ValueError("build the layer (call it once) before exporting its variables")
    return {n: store[k].detach() for n, store, k in _walk(layer, name or layer.layer_name)}


def load_reference_variables(layer: Layer, variables: dict, name: str | None = None, strict: bool = True) -> list:
    """Copy `variables` ({reference name: array / tensor}) into a built layer; returns the names that were loaded.
    strict: every variable of the layer must be present and every given name must belong to the layer."""
    if not getattr(layer, "built", False):
        raise ValueError("build the layer (call it once) before loading variables into it")
    own = {n: (store, k) for n, store, k in _walk(layer, name or layer.layer_name)}
    if strict:
        missing, unexpected = sorted(set(own) - set(variables)), sorted(set(variables) - set(own))
        if missing or unexpected:
            raise KeyError(f"missing variables {missing}; unexpected variables {unexpected}")
    loaded = []
    with torch.no_grad():
        for n, (store, k) in own.items():
            if n not in variables:
                continue
            src = torch.as_tensor(variables[n])
            dst = store[k]
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{n}: shape {tuple(src.shape)} does not match {tuple(dst.shape)}")
            dst.copy_(src.to(device=dst.device, dtype=dst.dtype))
            loaded.append(n)
    return loaded
