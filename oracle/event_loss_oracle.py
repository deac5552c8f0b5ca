"""CPU oracle for the blurred-L2 event loss (numpy restatement).  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/Tracker.py:204-224 (identically Mapper.py:593-615):
    loss = ((gt - pred)**2).sum()
    for ks, w in zip(kernel_sizes, kernel_weights):
        loss += w * ((gaussian_blur(gt, ks) - gaussian_blur(pred, ks))**2).sum()
    loss *= balancer
with ``gaussian_blur`` = torchvision.transforms.functional.gaussian_blur on a (C,H,W) tensor (third-party arithmetic,
environment.yaml pins torchvision 0.12; the tensor path is unchanged in the 0.26 of this image): reflect pad by ks // 2,
depthwise conv2d with outer(k1d, k1d), k1d = normalised exp(-0.5 (x / sigma)^2) on linspace(-(ks-1)/2, (ks-1)/2, ks),
sigma = 0.15 ks + 0.35, all float32.

The gradient is the analytic transpose: 2 d, plus per kernel 2 w pad^T(conv^T(blur(d))) with d = gt - pred.

Parity pinning: ``tests/test_event_loss_cpu.py`` checks value and gradient against the lines above executed with
torch + torchvision autograd in this container.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def gaussian_kernel1d(ks: int) -> np.ndarray:
    sigma = ks * 0.15 + 0.35
    half = (ks - 1) * 0.5
    x = np.linspace(-half, half, ks).astype(F32)
    pdf = np.exp(F32(-0.5) * (x / F32(sigma)) ** 2).astype(F32)
    return (pdf / pdf.sum(dtype=F32)).astype(F32)


def _blur(img: np.ndarray, ks: int) -> np.ndarray:
    """img [H,W,C] float32 -> blurred [H,W,C]"""
    r = ks // 2
    k1 = gaussian_kernel1d(ks)
    k2 = np.outer(k1, k1).astype(F32)
    pad = np.pad(img, ((r, r), (r, r), (0, 0)), mode="reflect")
    H, W, _ = img.shape
    out = np.zeros_like(img)
    for v in range(ks):
        for u in range(ks):
            out += k2[v, u] * pad[v:v + H, u:u + W]
    return out


def _blur_transpose(b: np.ndarray, ks: int) -> np.ndarray:
    """transpose of _blur: conv^T into the padded domain, then the reflect pad's fold"""
    r = ks // 2
    k1 = gaussian_kernel1d(ks)
    k2 = np.outer(k1, k1).astype(F32)
    H, W, C = b.shape
    gp = np.zeros((H + 2 * r, W + 2 * r, C), F32)
    for v in range(ks):
        for u in range(ks):
            gp[v:v + H, u:u + W] += k2[v, u] * b
    g = gp[r:r + H, r:r + W].copy()
    for j in range(1, r + 1):                      # rows: padded row r-j mirrors row j; row H-1+j mirrors H-1-j
        g[j] += gp[r - j, r:r + W]
        g[H - 1 - j] += gp[r + H - 1 + j, r:r + W]
    for i in range(1, r + 1):                      # columns, including the corners already folded into rows above
        colL = gp[:, r - i].copy(); colR = gp[:, r + W - 1 + i].copy()
        foldL = colL[r:r + H].copy(); foldR = colR[r:r + H].copy()
        for j in range(1, r + 1):
            foldL[j] += colL[r - j]; foldL[H - 1 - j] += colL[r + H - 1 + j]
            foldR[j] += colR[r - j]; foldR[H - 1 - j] += colR[r + H - 1 + j]
        g[:, i] += foldL
        g[:, W - 1 - i] += foldR
    return g


def event_loss(gt: np.ndarray, pred: np.ndarray, kernel_sizes=(9,), kernel_weights=(1.0,), balancer: float = 1.0):
    """-> (loss float64, parts float64[1 + n] = unblurred and blurred sums, gradient wrt pred float32 [H,W,C])"""
    gt = gt.astype(F32); pred = pred.astype(F32)
    d = gt - pred
    parts = [float((d.astype(np.float64) ** 2).sum())]
    grad = 2.0 * d
    total = parts[0]
    for ks, w in zip(kernel_sizes, kernel_weights):
        b = _blur(gt, ks) - _blur(pred, ks)
        s = float((b.astype(np.float64) ** 2).sum())
        parts.append(s)
        total += w * s
        grad = grad + F32(2.0 * w) * _blur_transpose(b, ks)
    return balancer * total, np.array(parts), (-F32(balancer) * grad).astype(F32)
