"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def rel_err(a, b):
    """max |a-b| / max(|b|) -- relative to the tensor's scale (SURVEY 8(d) tolerances)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / den)


def torch_t_vals(n_samples=32, n_surface=16, device="cpu"):
    """The reference's own linspace calls (Renderer.py:127-128, 153)."""
    import torch
    t32 = torch.linspace(0., 1., steps=n_samples, device=device)
    t64 = torch.linspace(0., 1., steps=n_surface).double().to(device)
    return t32, t64
