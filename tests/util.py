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


def elem_err(a, b, floor=2e-2):
    """ELEMENTWISE relative error with an absolute floor:  max_i |a_i - b_i| / (|b_i| + floor * max|b|).

    ``rel_err`` is relative to the tensor's largest entry, so a small entry could be badly off and pass; here every entry
    is held relative to its own size.  The floor: a gradient entry is a sum over ~5e4 float32 terms whose rounding error
    scales with the TERMS (eps * sqrt(N) * rms ~ 1e-5 of the tensor's largest entry, measured 7e-6 between the kernels and
    the float32 oracle), not with the entry, which cancellation can make arbitrarily small -- in the reference's own
    atomics as in the kernels'.  With floor = 2e-2 an entry of 2 % of the largest must be right to 5e-4 of ITSELF at
    tolerance 1e-3, and no entry may be off by more than 2e-5 of the largest."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b) + floor * max(np.abs(b).max(), 1e-30)
    return float((np.abs(a - b) / den).max())


def torch_t_vals(n_samples=32, n_surface=16, device="cpu"):
    """The reference's own linspace calls (Renderer.py:127-128, 153)."""
    import torch
    t32 = torch.linspace(0., 1., steps=n_samples, device=device)
    t64 = torch.linspace(0., 1., steps=n_surface).double().to(device)
    return t32, t64
