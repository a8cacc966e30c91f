"""Timing probe (not a test): ed2_spd_inverse against torch.linalg.inv (cuSOLVER/MAGMA) at the SNGP feature widths."""
import json
import numpy as np
import torch
from edward2_b200 import ops

out = {}
for n in (1024, 2048, 4096):
    rng = np.random.default_rng(n)
    phi = rng.standard_normal((2 * n, n)).astype(np.float32) * np.sqrt(2.0 / n)
    a = torch.tensor((phi.T @ phi + np.eye(n)).astype(np.float32), device="cuda")
    for name, fn in (("ed2", lambda: ops.spd_inverse(a)[0]), ("torch", lambda: torch.linalg.inv(a))):
        for _ in range(3):
            x = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            x = fn()
        e1.record(); torch.cuda.synchronize()
        out[f"{name}_{n}_ms"] = e0.elapsed_time(e1) / 5
    ref = np.linalg.inv(a.cpu().numpy().astype(np.float64))
    out[f"ed2_{n}_relerr"] = float(np.abs(ops.spd_inverse(a)[0].cpu().numpy() - ref).max() / np.abs(ref).max())
    out[f"torch_{n}_relerr"] = float(np.abs(torch.linalg.inv(a).cpu().numpy() - ref).max() / np.abs(ref).max())
print(json.dumps(out))
