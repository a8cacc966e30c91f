"""Shared helpers for the parity tests."""
import numpy as np
import torch

# Parity metric: tensor-scale relative error max|a-b| / max|b|.
TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.float64).cpu().numpy()


def dev(a, dtype, device, requires_grad=False):
    t = torch.tensor(np.asarray(a), dtype=torch.float32, device=device).to(dtype)
    if requires_grad:
        t.requires_grad_(True)
    return t


def round_to(a, dtype):
    """Round a float64 array to `dtype` and back: the oracle is evaluated on the same rounded inputs the kernel sees."""
    if dtype == torch.float32:
        return np.asarray(a, dtype=np.float32).astype(np.float64)
    t = torch.tensor(np.asarray(a), dtype=torch.float32).to(torch.bfloat16).to(torch.float64)
    return t.numpy()
